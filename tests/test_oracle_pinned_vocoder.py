"""CPU: the EnCodec decode-path restatement (oracle/vocoder_port.py) against transformers.EncodecModel, the
implementation audiocraft 1.2.0 delegates the 24 kHz codec to.  (The MBD diffusion stage is unpinned: no source.)"""
import pytest
import torch

from mvb200 import synth
from oracle import vocoder_port as V


def test_encodec_decode_port_matches_transformers_implementation():
    m, sd = synth.encodec_model_and_state_dict(seed=0)
    o = V.EncodecDecodeOracle(sd)
    g = torch.Generator().manual_seed(1)
    codes = torch.randint(0, 1024, (1, 8, 40), generator=g)
    with torch.no_grad():
        ref = m.decode(codes[None], [None]).audio_values
    got = o.decode(codes)
    assert got.shape == ref.shape == (1, 1, 40 * 320)
    assert (got - ref).abs().max() / ref.abs().max() < 1e-4
    lat = o.decode_latent(codes)
    with torch.no_grad():
        ref_lat = m.quantizer.decode(codes.transpose(0, 1))
    assert torch.allclose(lat, ref_lat, atol=1e-6)
