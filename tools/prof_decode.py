"""Short decode workload for ncu: full-size engine, prefill 48, then ONE persistent launch of N decode positions.
Usage: python tools/prof_decode.py [n_utts] [n_positions]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "metavoice-src_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from mvb200 import synth  # noqa: E402

utts = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
model, _, _ = bench.build_engine(dev, utts, 0, 1)
lens = [bench.T_PROMPT] * utts
prompts = [synth.synthetic_prompt(T, seed=7 + u) for u, T in enumerate(lens)]
spk = torch.cat([synth.synthetic_speaker(seed=11 + u) for u in range(utts)])
d_idx = [p.view(1, -1).repeat(2, 1).to(dev).contiguous() for p in prompts]
d_spk = [spk[u].to(dev).contiguous() for u in range(utts)]
bench.resident_pass(model, d_idx, d_spk, lens, steps, 1)
torch.cuda.synchronize()
print("launches", model._lib.mvb_s1_launch_count(model.handle), "positions", steps)
