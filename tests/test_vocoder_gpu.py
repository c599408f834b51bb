"""GPU: EnCodec decode path (RVQ + SEANet decoder incl. the persistent LSTM kernel) through the C ABI vs the CPU
oracle (oracle/vocoder_port.py, itself pinned to transformers.EncodecModel) on the same seeded checkpoint."""
import numpy as np
import pytest
import torch

from mvb200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T", [40, 375])
def test_encodec_decode_vs_oracle(T):
    from mvb200.vocoder import EncodecDecodeEngine
    from oracle import vocoder_port as V
    _, sd = synth.encodec_model_and_state_dict(seed=0)
    eng = EncodecDecodeEngine(sd, device="cuda:0", max_frames=512)
    o = V.EncodecDecodeOracle(sd)
    g = torch.Generator().manual_seed(T)
    codes = torch.randint(0, 1024, (8, T), generator=g)
    lat = eng.decode_latent(codes).cpu()
    assert torch.allclose(lat, o.decode_latent(codes[None])[0], atol=1e-5)
    wav = eng.decode(codes).cpu()
    ref = o.decode(codes[None])[0, 0]
    assert wav.shape == ref.shape == (T * 320,)
    err = float((wav - ref).abs().max() / ref.abs().max())
    print(f"EnCodec decode T={T}: rel err {err:.2e}")
    assert err < 1e-3


def test_frame_count_limits():
    from mvb200.vocoder import EncodecDecodeEngine
    _, sd = synth.encodec_model_and_state_dict(seed=0)
    eng = EncodecDecodeEngine(sd, device="cuda:0", max_frames=64)
    with pytest.raises(ValueError):
        eng.decode(torch.zeros(8, 100, dtype=torch.int64))


def test_audio_post_on_device_matches_torchaudio_meter_and_host_writer(tmp_path):
    """Row N2: loudness normalisation (-14 LUFS, BS.1770-4 == torchaudio.functional.loudness, the function audiocraft
    calls), tanh compressor and PCM16 conversion on the device vs the host restatement (mvb200/audio_out.py)."""
    import torchaudio
    from mvb200 import audio_out as A
    for secs, amp, seed in ((5.0, 0.3, 1), (1.7, 0.9, 2), (12.0, 0.02, 3)):
        wav = torch.from_numpy(synth.synthetic_waveform(secs, 24000, seed=seed)) * (amp / 0.3)
        ref_lkfs = float(torchaudio.functional.loudness(wav[None], 24000))
        blob, lkfs, gain = A.wav_bytes_on_device(wav.cuda(), 24000, return_stats=True)
        assert abs(lkfs - ref_lkfs) < 2e-3, (lkfs, ref_lkfs)
        p = A.audio_write_wav(str(tmp_path / f"h{seed}"), wav[None], 24000, strategy="loudness", loudness_compressor=True)
        host = open(p, "rb").read()
        assert blob[:44] == host[:44] and len(blob) == len(host)
        a = np.frombuffer(blob[44:], dtype="<i2").astype(np.int32)
        b = np.frombuffer(host[44:], dtype="<i2").astype(np.int32)
        assert np.abs(a - b).max() <= 2          # one LSB of PCM16 from the 2e-3 dB meter difference, plus rounding
    quiet = torch.from_numpy(synth.synthetic_waveform(2.0, 24000, seed=4)) * 1e-3     # below the energy floor: untouched
    blob, lkfs, gain = A.wav_bytes_on_device(quiet.cuda(), 24000, return_stats=True)
    assert gain == 1.0
    assert np.array_equal(np.frombuffer(blob[44:], dtype="<i2"), (quiet.clamp(-1, 1).numpy() * 32767.0).round().astype("<i2"))
