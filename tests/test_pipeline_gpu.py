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
