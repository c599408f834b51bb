"""Multi-band-diffusion workload for ncu launch lists: default parametrised config, N seconds of audio, K sampler calls.
Usage: python tools/prof_mbd.py [seconds] [n_models] [n_calls]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "metavoice-src_b200"))
import torch  # noqa: E402

from mvb200 import synth  # noqa: E402
from mvb200.mbd import MBDSettings, MultiBandDiffusionEngine  # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
n_models = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n_calls = int(sys.argv[3]) if len(sys.argv) > 3 else 2
steps = [999 - 50 * i for i in range(n_calls)] + [0]
s = MBDSettings(n_models=n_models, step_list=steps)
eng = MultiBandDiffusionEngine(synth.mbd_checkpoint(s, 0), s, device="cuda:0", max_seconds=secs + 1)
frames = int(secs * 75)
T = frames * 320
cond = torch.randn(128, frames, device="cuda")
wav = torch.randn(T, device="cuda") * 0.1
eng.tokens_to_wav(cond, wav, seed=1)
torch.cuda.synchronize()
t0 = time.perf_counter()
eng.tokens_to_wav(cond, wav, seed=2)
torch.cuda.synchronize()
print(f"mbd {secs}s audio, {n_models} models x {n_calls} calls: {(time.perf_counter() - t0) * 1e3:.1f} ms")
