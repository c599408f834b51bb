"""GPU debug: per-phase time breakdown of the persistent decode kernel (clock64 stamps of every CTA).
Writes gpurun_out/trace_<n>.npy and prints medians per event interval."""
import ctypes as C
import os
import sys

os.environ["MVB_PC_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "metavoice-src_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from mvb200 import _lib, synth  # noqa: E402
from mvb200.fast_model import ModelArgs, Transformer, pack_arena  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
L_mid = int(sys.argv[2]) if len(sys.argv) > 2 else 423
for kv in (sys.argv[3].split("+") if len(sys.argv) > 3 and sys.argv[3] else []):   # e.g. MVB_PC_WB=0+MVB_PF_MODE=2
    k, v = kv.split("=")
    os.environ[k] = v
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
d = synth.FULL
cfg = ModelArgs.from_name("metavoice-1B")
arena, offsets = pack_arena(synth.stage1_state_dict(d, 0), d.n_layer)
m = Transformer(cfg, arena.to(dev), offsets, device=dev)
m.setup_caches(2 * n, cfg.block_size, kv_dtype="bf16")
lib, h, st = m._lib, m.handle, m._stream()
for u in range(n):
    sp = _lib.Sampling(3.0, 1.0, 0.95, 0, 9999, 5 + u)
    spk = synth.synthetic_speaker(seed=u).to(dev).reshape(-1).contiguous()
    _lib.check(lib.mvb_s1_set_speaker(h, u, spk.data_ptr(), st))
    _lib.check(lib.mvb_s1_begin(h, u, 100 + u, L_mid - 30, C.byref(sp), None, None, st))
_lib.check(lib.mvb_s1_decode(h, n, 30, st))
torch.cuda.synchronize()
buf = np.zeros((148, 512), dtype=np.int64)
got = lib.mvb_s1_trace_fetch(h, buf.ctypes.data_as(C.c_void_p), 148, st)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", f"trace_{n}.npy"), buf)
t = buf[:, :2 + 14 * 24 + 4].astype(np.float64)
t = (t - t[:, :1]) / 1.9e3   # ~us at 1.9 GHz
names = ["qkv.wait", "qkv.stage", "qkv.epi", "att.wait", "att.work", "o.wait", "o.stage", "o.epi", "w13.wait", "w13.stage",
         "w13.epi", "w2.wait", "w2.stage", "w2.epi"]
dt = np.diff(t, axis=1)          # interval ending at event i+1
print("total us (median over CTAs):", np.median(t[:, -1]))
per = dt[:, 1:1 + 14 * 24].reshape(148, 24, 14)   # skip start->embed
for j, nm in enumerate(names):
    v = per[:, 2:, j]
    print(f"{nm:10s} median {np.median(v):7.2f}  mean {v.mean():7.2f}  p90 {np.percentile(v, 90):7.2f}  max {v.max():7.2f}")
print("per-layer sum (median CTA):", np.median(per[:, 2:, :].sum(axis=2)))
sub = buf[:, 380:380 + 5 * 24].astype(np.float64).reshape(148, 24, 5) / 1.9e3
if sub.any():      # library built with -DPC_ATT_PROF: cumulative attention sub-phase clocks per layer
    for j, nm in enumerate(["att.begin(q load)", "att.append_cur", "att.kv_wait", "att.tile_math", "att.unit_end"]):
        v = sub[:, 2:, j]
        print(f"{nm:18s} median {np.median(v):7.2f}  mean {v.mean():7.2f}  p90 {np.percentile(v, 90):7.2f}  max {v.max():7.2f}")
