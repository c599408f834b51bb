"""CPU: host-side pieces of the boundary that need no GPU (checkpoint layout, arena packing, tokenizer)."""
import torch
import pytest

from mvb200 import synth
from mvb200.fast_model import ModelArgs, pack_arena, transformer_configs
from mvb200.tokenise import TrainedBPETokeniser


def test_model_args_match_reference_config():
    a = ModelArgs.from_name("metavoice-1B")
    assert (a.n_layer, a.n_head, a.dim, a.vocab_size, a.head_dim, a.intermediate_size, a.block_size) == \
        (24, 16, 2048, 2562, 128, 5632, 2048)
    assert synth.FULL.n_params() == 1_248_438_272


def test_arena_packing_is_byte_exact():
    d = synth.TINY
    sd = synth.stage1_state_dict(d, 1)
    sd = {"_orig_mod." + k: v for k, v in sd.items()}  # compiled-module prefix must be stripped
    arena, off = pack_arena(sd, d.n_layer)
    assert len(off) == 5 + 7 * d.n_layer and all(o % 256 == 0 for o in off)
    w = sd["_orig_mod.transformer.h.1.mlp.c_proj.weight"]
    o = off[5 + 7 * 1 + 6]
    got = arena[o:o + w.numel() * 2].view(torch.bfloat16).view_as(w)
    assert torch.equal(got, w)


def test_synthetic_tokenizer_layout():
    tok = TrainedBPETokeniser(**synth.synthetic_tokenizer_meta())
    ids = tok.encode("Hello, what's up?")
    assert ids[-1] == 2561 == tok.eot_token and all(2049 <= i <= 2561 for i in ids)
    assert tok.decode(ids[:-1]) == "Hello, what's up?"


def test_checkpoint_container_layout():
    ck = synth.stage1_checkpoint(synth.TINY, 0)
    assert set(["model", "model_args", "config", "meta"]) <= set(ck)
    assert ck["config"]["causal"] is True and ck["meta"]["tokenizer"]["offset"] == 2049
    assert len(ck["model"]) == 5 + 7 * synth.TINY.n_layer


def test_audio_write_loudness_normalisation(tmp_path):
    """a18 (decoders.py:40-47 -> audiocraft audio_write, strategy="loudness", compressor): the pre-compressor signal
    sits at -14 LUFS (torchaudio's BS.1770 meter, the function audiocraft itself calls), quiet signals pass through,
    the file is PCM16 mono at the requested rate."""
    import torchaudio
    from mvb200 import audio_out as A
    g = torch.Generator().manual_seed(0)
    t = torch.arange(48000) / 24000.0
    wav = (0.05 * torch.sin(2 * torch.pi * 220 * t) + 0.01 * torch.randn(48000, generator=g))[None]
    out = A.normalize_loudness(wav, 24000, loudness_compressor=False)
    assert abs(torchaudio.functional.loudness(out, 24000).item() + 14.0) < 1e-3
    comp = A.normalize_loudness(wav, 24000, loudness_compressor=True)
    assert torch.allclose(comp, torch.tanh(out), atol=1e-6) and comp.abs().max() < 1.0
    quiet = 1e-4 * wav
    assert A.normalize_loudness(quiet, 24000, loudness_compressor=True) is quiet          # below the energy floor
    p = A.audio_write_wav(str(tmp_path / "x"), wav[0], 24000)
    back, sr = A.read_wav_pcm16(p)
    assert p.endswith(".wav") and sr == 24000 and back.shape == (1, 48000)
    assert (back - comp.clamp(-1, 1)).abs().max() < 1.0 / 32767 + 1e-6
    with pytest.raises(ValueError):
        A.audio_write_wav(str(tmp_path / "y"), torch.zeros(1, 2, 100), 24000)
