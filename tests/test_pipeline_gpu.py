"""GPU: the TTS façade mirror end to end on tiny synthetic checkpoints in the reference's on-disk layout:
text -> stage-1 tokens -> de-interleave -> stage-2 codes -> EnCodec decoder -> PCM16 wav file."""
import os
import wave

import numpy as np
import pytest
import torch

from mvb200 import synth

pytestmark = pytest.mark.gpu


def test_tts_synthesise_writes_a_wav(tmp_path):
    from mvb200.fast_inference import TTS, normalize_text
    snap = tmp_path / "snapshot"
    snap.mkdir()
    # stage-1 checkpoint whose text prompt makes it emit audio tokens: random weights rarely hit EOA, so cap via block
    torch.save(synth.stage1_checkpoint(synth.TINY, 0), str(snap / "first_stage.pt"))
    torch.save(synth.stage2_checkpoint(synth.S2_TINY, 1), str(snap / "second_stage.pt"))
    _, enc_sd = synth.encodec_model_and_state_dict(0)
    spk = snap / "spk.pt"
    torch.save(synth.synthetic_speaker(), str(spk))
    tts = TTS(str(snap), output_dir=str(tmp_path / "out"), encodec_state_dict=enc_sd, device="cuda:0")
    # bound the run: the tiny random model does not emit end-of-audio, the engine stops at the context limit
    tts.model._cfg.max_new = 256
    path = tts.synthesise("Hello, what's up?", str(spk), top_p=0.95, guidance_scale=3.0, temperature=1.0)
    assert os.path.isfile(path) and path.endswith(".wav")
    with wave.open(path) as w:
        assert w.getframerate() == 24000 and w.getnchannels() == 1 and w.getsampwidth() == 2
        n = w.getnframes()
    assert n >= 9600 and n % 320 == 0
    assert normalize_text("a’b \t c\n") == "a'b c"
    with pytest.raises(ValueError):
        normalize_text("snow☃man")
    with pytest.raises(FileNotFoundError):
        tts.synthesise("hi", str(snap / "missing.pt"))


def test_tts_wav_speaker_reference_and_long_form(tmp_path, monkeypatch):
    """Reference-shaped call: a >= 30 s wav as the speaker reference (embedded by the on-device speaker encoder, cached
    on disk like inference.py:419-435), and long-form text synthesised as a continuous batch of <= 220-char chunks."""
    from mvb200 import audio_out as A
    from mvb200.fast_inference import TTS
    monkeypatch.setenv("HOME", str(tmp_path))             # keep ~/.cache/fam inside the test directory
    snap = tmp_path / "snapshot"
    snap.mkdir()
    torch.save(synth.stage1_checkpoint(synth.TINY, 0), str(snap / "first_stage.pt"))
    torch.save(synth.stage2_checkpoint(synth.S2_TINY, 1), str(snap / "second_stage.pt"))
    torch.save({"model_state": synth.speaker_encoder_state_dict(3)}, str(snap / "speaker_encoder.pt"))
    _, enc_sd = synth.encodec_model_and_state_dict(0)
    ref = A.audio_write_wav(str(tmp_path / "speaker_ref"), torch.from_numpy(synth.synthetic_waveform(31.0, 22050, seed=9))[None],
                            22050, strategy="clip")
    short = A.audio_write_wav(str(tmp_path / "short_ref"), torch.from_numpy(synth.synthetic_waveform(3.0, 22050, seed=9))[None],
                              22050, strategy="clip")
    from mvb200.mbd import MBDSettings, UnetSettings
    from oracle import mbd_port as M
    small = M.MBDConfig(n_models=2, unet=M.UnetCfg(hidden=16, depth=2, growth=2.0), proc_bands=4, eq_bands=8, step_list=[999, 499, 0])
    mbd_settings = MBDSettings(n_models=2, unet=UnetSettings(hidden=16, depth=2, growth=2.0), proc_bands=4, eq_bands=8, step_list=[999, 499, 0])
    tts = TTS(str(snap), output_dir=str(tmp_path / "out"), encodec_state_dict=enc_sd, device="cuda:0", max_utts=3,
              mbd_checkpoint=synth.mbd_checkpoint(small, 0), mbd_settings=mbd_settings)      # vocoder = EnCodec decode + MBD refinement
    assert tts.mbd is not None
    tts.model._cfg.max_new = 256
    with pytest.raises(Exception, match="too short"):
        tts.synthesise("hi", short)
    path = tts.synthesise("Hello there.", ref)
    assert os.path.isfile(path)
    cached = [f for f in os.listdir(tmp_path / ".cache" / "fam") if f.startswith("embedding_")]
    assert len(cached) == 1
    e = torch.load(str(tmp_path / ".cache" / "fam" / cached[0]))
    assert e.shape == (1, 256) and abs(float(e.norm()) - 1.0) < 1e-4
    text = " ".join(["The quick brown fox jumps over the lazy dog near the quiet river bank."] * 8)   # 5 chunks > 3 slots
    long_path = tts.synthesise_long(text, ref)
    with wave.open(long_path) as w:
        n = w.getnframes()
    assert n >= 5 * 9600
