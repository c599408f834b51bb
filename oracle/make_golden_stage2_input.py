"""ORACLE (test infrastructure): tests/golden/stage2_input.npz -- the (b, 2, block_size) input tensor the REFERENCE'S OWN
``Model.non_causal_sample`` (fam/llm/inference.py:248-338) builds, captured by calling that very method (unbound) on a
stand-in ``self`` whose ``model.generate`` records ``in_x``.  ``fam.llm.inference`` imports ``fam.llm.decoders``, which
needs audiocraft/julius (absent): both are stubbed as empty modules -- no line of the input-building code touches them.
Build container only:  python oracle/make_golden_stage2_input.py
"""
import contextlib
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "metavoice-src_b200"))
from mvb200 import synth  # noqa: E402
from oracle import ref_harness as R  # noqa: E402


def reference_model_class():
    R._import_reference()

    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _MBD:
        @staticmethod
        def get_mbd_24khz(bw=6):
            return None

    stub("julius")
    stub("audiocraft")
    stub("audiocraft.data")
    stub("audiocraft.data.audio", audio_read=None, audio_write=None)
    stub("audiocraft.models", MultiBandDiffusion=_MBD)
    stub("df")
    stub("df.enhance", enhance=None, init_df=None, load_audio=None, save_audio=None)
    stub("pydub", AudioSegment=None)
    for name in ("huggingface_hub", "tyro"):
        try:
            __import__(name)
        except Exception:
            stub(name, snapshot_download=None, cli=lambda *a, **k: None)
    import fam.llm.inference as inf
    return inf.Model


def reference_in_x(Model, tokenizer, texts, encodec_tokens, block_size):
    """Run the reference's own input-building statements and return the tensor it hands to ``GPT.generate``."""
    captured = {}

    class _Gen:
        def generate(self, in_x, *a, **k):
            captured["in_x"] = in_x.clone()
            return torch.zeros((in_x.shape[0], 6, in_x.shape[2]), dtype=torch.long)

    class _Dec:
        def decode(self, tokens, causal):
            return None

    fake = types.SimpleNamespace(
        tokenizer=tokenizer, config=types.SimpleNamespace(device="cpu", num_samples=1), _num_encodec_codebooks=8,
        _encodec_codes_pad_token=1024, _encodec_ctx_window=block_size, speaker_cond=True, _ctx=contextlib.nullcontext(),
        model=_Gen(), decoder=_Dec())
    Model.non_causal_sample(fake, texts=texts, encodec_tokens=encodec_tokens, batch_size=len(texts), top_k=200,
                            temperature=1.0, speaker_embs=torch.zeros(len(texts), 1, 256))
    return captured["in_x"]


def cases(block_size):
    g = torch.Generator().manual_seed(123)
    mk = lambda n: torch.randint(0, 1024, (1, 2, n), generator=g)
    return [("Hello, what's up?", mk(60)),                       # padded on the right
            ("ok", mk(block_size - 4)),                          # text + codes + pad exceed block_size by one: cut
            ("This one is far too long for the window.", mk(block_size + 50)),   # truncated
            ("exact", mk(block_size - 1 - 6))]                   # may land exactly on block_size


def main():
    from mvb200.tokenise import TrainedBPETokeniser
    Model = reference_model_class()
    meta = synth.synthetic_tokenizer_meta(n_text_tokens=512, offset=1025)     # stage-2 text ids: 1025.. (App. D)
    tok = TrainedBPETokeniser(**meta)
    out = {}
    for tag, bs in (("tiny", synth.S2_TINY.block_size), ("full", synth.S2_FULL.block_size)):
        cs = cases(bs)
        in_x = reference_in_x(Model, tok, [c[0] for c in cs], [c[1] for c in cs], bs)
        out[f"{tag}_in_x"] = in_x.numpy().astype(np.int32)
        for i, (text, codes) in enumerate(cs):
            out[f"{tag}_text_{i}"] = np.asarray(tok.encode(text), np.int32)
            out[f"{tag}_codes_{i}"] = codes[0].numpy().astype(np.int32)
        out[f"{tag}_block"] = np.int32(bs)
        print(tag, in_x.shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "stage2_input.npz"), **out)


if __name__ == "__main__":
    main()
