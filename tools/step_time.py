"""Decode-step time (CUDA events) for several engine configurations in one process.
Usage: python tools/step_time.py "A1:1,B:8,C:1,C:1:MVB_PC_WB=0+MVB_PF_MODE=2+MVB_PF_AHEAD=16,C:8" [context_len]
  A1 = CUDA-core path with PDL, A0 = without, B = tensor-core rows path, C = persistent fused kernel;
  second field = utterances; optional third field = '+'-separated environment switches read by mvb_s1_create."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "metavoice-src_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from mvb200 import _lib, synth  # noqa: E402
from mvb200.fast_model import ModelArgs, Transformer, pack_arena  # noqa: E402

spec = sys.argv[1] if len(sys.argv) > 1 else "A1:1,A0:1,B:1,B:8"
L_mid = int(sys.argv[2]) if len(sys.argv) > 2 else 423
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
d = synth.FULL
cfg = ModelArgs.from_name("metavoice-1B")
arena, offsets = pack_arena(synth.stage1_state_dict(d, 0), d.n_layer)
arena = arena.to(dev)
peak, _ = bench.measured_peaks()
res = []
for item in spec.split(","):
    parts = item.split(":")
    kind, n = parts[0], int(parts[1])
    extra = dict(kv.split("=") for kv in parts[2].split("+")) if len(parts) > 2 and parts[2] else {}
    for k in ("MVB_PC_WB", "MVB_PC_FUSED", "MVB_PF_MODE", "MVB_PF_AHEAD", "MVB_PC_EPI", "MVB_PC_KVPF", "MVB_PC_SPLITS", "MVB_PC_ATT_MMA"):
        os.environ.pop(k, None)
    os.environ.update(extra)
    os.environ["MVB_PDL"] = "0" if kind.endswith("0") else "1"
    os.environ["MVB_DECODE_B_MIN"] = "1" if kind.startswith("B") else "9999"
    os.environ["MVB_PATHC"] = "1" if kind[0] in "CS" else "0"
    m = Transformer(cfg, arena, offsets, device=dev)
    m.setup_caches(2 * n, cfg.block_size, kv_dtype="bf16")
    lib, h, st = m._lib, m.handle, m._stream()
    for u in range(n):
        sp = _lib.Sampling(3.0, 1.0, 0.95, 0, 9999, 5 + u)
        spk = synth.synthetic_speaker(seed=u).to(dev).reshape(-1).contiguous()
        _lib.check(lib.mvb_s1_set_speaker(h, u, spk.data_ptr(), st))
        _lib.check(lib.mvb_s1_begin(h, u, 100 + u, L_mid - 100, C.byref(sp), None, None, st))
    _lib.check(lib.mvb_s1_decode(h, n, 20, st))   # warm-up + graph capture
    torch.cuda.synchronize()
    reps = 160
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.mvb_s1_decode(h, n, reps, st))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    by = bench.W_BYTES + n * bench.KV_BYTES_PER_POS * (L_mid + 1)
    r = {"cfg": item, "ms_per_step": round(ms, 4), "tok_per_s": round(n / ms * 1e3, 1), "GBps": round(by / ms / 1e6, 1),
         "roofline_frac": round(by / ms / 1e6 / peak, 3)}
    print(json.dumps(r), flush=True)
    res.append(r)
    m.close()
    del m
    torch.cuda.empty_cache()
