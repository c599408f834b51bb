"""CPU: host-side pieces of the boundary that need no GPU (checkpoint layout, arena packing, tokenizer)."""
import torch

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
