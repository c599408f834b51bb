"""ORACLE (test infrastructure): tests/golden/stage2.npz from the REFERENCE'S OWN ``fam.llm.model.GPT`` (causal=False)
run on CPU in fp32 on a seeded synthetic second_stage checkpoint.  Build container only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "metavoice-src_b200"))
from mvb200 import synth  # noqa: E402
from oracle import ref_harness as R, stage2_port as P  # noqa: E402


def reference_gpt(d, sd):
    R._import_reference()
    from fam.llm.model import GPT, GPTConfig
    m = GPT(GPTConfig(**synth.stage2_model_args(d)), speaker_emb_dim=d.speaker_emb_dim).eval()
    m.load_state_dict({k: v.float() for k, v in sd.items()})
    return m


def main():
    out = {}
    for tag, d, frames in (("tiny", synth.S2_TINY, 60), ("full", synth.S2_FULL, 375)):
        sd = synth.stage2_state_dict(d, 1)
        m = reference_gpt(d, sd)
        text, cb0, cb1 = synth.synthetic_stage2_input(d, frames)
        idx = P.build_input(text, cb0, cb1, d.block_size)[None]
        spk = synth.synthetic_speaker(seed=21)[None]                       # (b=1, 1, 256) as fast_inference.py:144
        with torch.no_grad():
            logits, _ = m(idx, speaker_embs=spk)
        torch.manual_seed(4242)
        y = m.generate(idx, None, temperature=1.0, top_k=200, top_p=None, speaker_embs=spk, batch_size=1, guidance_scale=None)
        # torch.multinomial(n=1) == argmax(p / Exp(1)) drawn per (hierarchy, batch row): reproduce it from the seed
        torch.manual_seed(4242)
        noise = [torch.empty(1, d.block_size, v).exponential_(1) for v in d.target_vocab_sizes]
        y2 = P.non_causal_sample([l.clone() for l in logits], 1.0, 200, noise)
        assert torch.equal(y, y2), "multinomial != exp-race with the same generator stream"
        keep = [0, 1, len(text) - 1, len(text), len(text) + frames // 2, len(text) + frames, d.block_size - 1]
        out[f"{tag}_idx"] = idx.numpy().astype(np.int32)
        out[f"{tag}_spk"] = spk.numpy()
        out[f"{tag}_keep"] = np.asarray(keep, np.int32)
        out[f"{tag}_logits"] = torch.stack([l[0, keep] for l in logits]).numpy()      # [6, n_keep, V]
        out[f"{tag}_tokens"] = y[0].numpy().astype(np.int32)                          # [6, t]
        out[f"{tag}_frames"] = np.int32(frames)
        out[f"{tag}_ntext"] = np.int32(len(text))
        out[f"{tag}_checksum"] = np.float64(synth.state_dict_checksum(sd))
        print(tag, "ok", y.shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "stage2.npz"), **out)


if __name__ == "__main__":
    main()
