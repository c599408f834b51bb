"""Stage-2 + EnCodec decoder workload for ncu (launch lists and --set full captures of the downstream stages):
one 5 s utterance (375 frames), the same calls bench.py's pipeline leg makes.  Usage: python tools/prof_downstream.py [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "metavoice-src_b200"))
import torch  # noqa: E402

from mvb200 import synth  # noqa: E402
from mvb200.second_stage import SecondStage  # noqa: E402
from mvb200.vocoder import EncodecDecodeEngine  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = "cuda:0"
s2 = SecondStage(synth.stage2_checkpoint(synth.S2_FULL, 1), device=dev, max_batch=1)
codec = EncodecDecodeEngine(synth.encodec_model_and_state_dict(0)[1], device=dev, max_frames=1024)
frames = 375
g = torch.Generator().manual_seed(3)
text_ids = torch.randint(1025, 1537, (11,), generator=g).tolist() + [1537]
cb = [torch.randint(0, 1024, (frames,), generator=g).tolist() for _ in range(2)]
spk = synth.synthetic_speaker(seed=0).to(dev).reshape(1, -1)
for k in range(reps + 1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx = s2.build_input(text_ids, cb)[None]
    out = s2.forward_tokens(idx, spk, 1.0, 200, seed=k)
    codes8 = torch.cat([idx[0, :, len(text_ids):len(text_ids) + frames].to(dev), out[0, :, len(text_ids):len(text_ids) + frames]]).clamp_(0, 1023)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    wav = codec.decode(codes8)
    cond = codec.decode_latent(codes8)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    if k:
        print(f"stage-2 {1e3 * (t1 - t0):.2f} ms, EnCodec decode + latent {1e3 * (t2 - t1):.2f} ms, wav {tuple(wav.shape)}")
