"""Host-side calibration for the CPU baseline: effective core count (affinity / cgroup quota) and decode-step time of
the oracle port for a few thread counts and dtypes, each bounded.  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "metavoice-src_b200"))
import torch  # noqa: E402

from mvb200 import synth  # noqa: E402
from oracle import stage1_port as P  # noqa: E402


def cgroup_quota():
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        return None if q == "max" else float(q) / float(p)
    except Exception:
        return None


info = {"nproc": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)), "cgroup_cpus": cgroup_quota(), "runs": []}
d = synth.FULL
sd = synth.stage1_state_dict(d, 0)
spk = synth.synthetic_speaker()
for dtype in (torch.bfloat16, torch.float32):
    m = P.Stage1Oracle(sd, d.n_head, d.norm_eps, dtype, faithful_full_cache=True)
    m.setup_caches()
    for thr in (8, 16, 32, 64, 128):
        if thr > (os.cpu_count() or 1):
            continue
        torch.set_num_threads(thr)
        tok = torch.tensor([[5], [5]], dtype=torch.int32)
        t0 = time.perf_counter()
        n = 0
        while n < 4 and time.perf_counter() - t0 < 6.0:
            m.forward(tok, spk.to(dtype), torch.tensor([10 + n]))
            n += 1
        dt = (time.perf_counter() - t0) / max(n, 1)
        info["runs"].append({"dtype": str(dtype), "threads": thr, "s_per_step": round(dt, 4)})
        print(info["runs"][-1], flush=True)
print(json.dumps(info))
