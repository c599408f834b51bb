"""GPU: EnCodec decode path (RVQ + SEANet decoder incl. the persistent LSTM kernel) through the C ABI vs the CPU
oracle (oracle/vocoder_port.py, itself pinned to transformers.EncodecModel) on the same seeded checkpoint."""
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
