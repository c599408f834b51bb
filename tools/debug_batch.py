"""GPU debug: persistent-kernel logits for a batch of 3 utterances vs each utterance alone (same state)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "metavoice-src_b200"))
import torch  # noqa: E402

from mvb200 import _lib, synth  # noqa: E402
from mvb200.fast_model import ModelArgs, Transformer  # noqa: E402

d = synth.TINY
sd = synth.stage1_state_dict(d, 0)
cfg = ModelArgs(block_size=d.block_size, vocab_size=d.vocab_size, n_layer=d.n_layer, n_head=d.n_head, dim=d.dim)
kv = sys.argv[1] if len(sys.argv) > 1 else "bf16"
lens = [5, 17, 9]
prompts = [synth.synthetic_prompt(T, seed=20 + i) for i, T in enumerate(lens)]
spks = [synth.synthetic_speaker(seed=30 + i) for i in range(3)]


def run(n_slots, which):
    m = Transformer.from_state_dict(sd, cfg, device="cuda:0")
    m.setup_caches(2 * n_slots, d.block_size, kv_dtype=kv)
    lib, h, st = m._lib, m.handle, m._stream()
    sp = _lib.Sampling(2.0, 1.0, 0.9, 0, 9999, 1)
    outs = []
    for slot, i in enumerate(which):
        idx = prompts[i].view(1, -1).repeat(2, 1).cuda()
        m.forward(idx, spks[i].cuda(), torch.arange(lens[i]), utt=slot)
    for step in range(4):
        for slot, i in enumerate(which):
            _lib.check(lib.mvb_s1_begin(h, slot, 100 + 7 * i + step, lens[i] + step, C.byref(sp), None, None, st))
        lg = torch.empty(2 * len(which), d.vocab_size, device="cuda")
        _lib.check(lib.mvb_s1_step_logits(h, len(which), lg.data_ptr(), st))
        outs.append(lg.cpu())
    return outs


batch = run(3, [0, 1, 2])
for i in range(3):
    single = run(1, [i])
    for step in range(4):
        a, b = batch[step][2 * i:2 * i + 2], single[step]
        print(f"utt {i} step {step}: max|batch-single| = {(a - b).abs().max():.3e}  (|logit| max {b.abs().max():.3f})")
# run-to-run determinism of the batch itself
batch2 = run(3, [0, 1, 2])
print("batch run-to-run max diff", max(float((x - y).abs().max()) for x, y in zip(batch, batch2)))
