"""ORACLE (test infrastructure): generate tests/golden/*.npz by executing the REFERENCE'S OWN code
(/root/reference, CPU) on seeded synthetic checkpoints.  Run in the build container only:

    python oracle/make_golden.py [--skip-full]

The reference ships no golden vectors for this path, so these files are the pin:
they hold what ``fam.llm.fast_model.Transformer`` / ``fam.llm.fast_inference_utils.{generate,sample}``
produce, and both ``oracle/stage1_port.py`` and the CUDA engine are tested against them.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "metavoice-src_b200"))

from mvb200 import synth  # noqa: E402
from oracle import ref_harness as R  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class _Recorder:
    """Records what the reference's sampler draws/produces without changing its arithmetic:
    the replacement performs the identical two statements of fast_inference_utils.py:61-65."""

    def __init__(self, fiu):
        self.fiu = fiu
        self.q, self.probs, self.idx = [], [], []
        self._orig = fiu.multinomial_sample_one_no_sync

    def __enter__(self):
        rec = self

        def recording(probs_sort):
            q = torch.empty_like(probs_sort).exponential_(1)
            out = torch.argmax(probs_sort / q, dim=-1, keepdim=True).to(dtype=torch.int)
            rec.q.append(q.float().clone()); rec.probs.append(probs_sort.float().clone()); rec.idx.append(int(out))
            return out

        self.fiu.multinomial_sample_one_no_sync = recording
        return self

    def __exit__(self, *a):
        self.fiu.multinomial_sample_one_no_sync = self._orig


def _run_generate(fiu, model, prompt, spk, n_new, seed, sampling, record_logits):
    logits_log = []
    hook = model.register_forward_hook(lambda m, i, o: logits_log.append(o[:, -1].float().clone()))
    torch.manual_seed(seed)
    with _Recorder(fiu) as rec:
        y = fiu.generate(model, prompt, spk, max_new_tokens=n_new, end_of_audio_token=9999, **sampling)
    hook.remove()
    return y, logits_log, rec


def golden_stage1(dims, tag, T, n_new, keep_steps, weight_seed=0, with_bf16=True):
    fiu = R.reference_functions()
    sd = synth.stage1_state_dict(dims, weight_seed)
    prompt = synth.synthetic_prompt(T)
    spk = synth.synthetic_speaker()
    sampling = dict(temperature=torch.tensor(1.0), top_p=torch.tensor(0.95),
                    guidance_scale=torch.tensor(3.0), top_k=None)
    t0 = time.time()
    model = R.build_reference_model(sd, dims, torch.float32)
    y, logits_log, rec = _run_generate(fiu, model, prompt, spk, n_new, 1337, sampling, True)
    # cross-check: the recorder must not perturb the reference (same seed, unpatched code)
    if dims.n_layer <= 4:
        model_b = R.build_reference_model(sd, dims, torch.float32)
        torch.manual_seed(1337)
        y_plain = fiu.generate(model_b, prompt, spk, max_new_tokens=n_new, end_of_audio_token=9999, **sampling)
        assert torch.equal(y, y_plain)
    print(f"[{tag}] reference fp32 generate {n_new} tokens: {time.time()-t0:.1f}s")
    gen = y[T:].to(torch.int32)
    assert len(logits_log) == n_new and gen.numel() == n_new
    out = dict(
        prompt=prompt.numpy(), spk=spk.numpy(), tokens=gen.numpy(),
        steps=np.asarray(keep_steps, dtype=np.int32),
        logits=torch.stack([logits_log[s] for s in keep_steps]).numpy(),       # [n,2,V] pre-sampling
        noise=torch.stack([rec.q[s] for s in keep_steps]).numpy(),
        probs=torch.stack([rec.probs[s] for s in keep_steps]).numpy(),
        weight_seed=np.int64(weight_seed), weight_checksum=np.float64(synth.state_dict_checksum(sd)),
        n_layer=np.int32(dims.n_layer), n_head=np.int32(dims.n_head), dim=np.int32(dims.dim),
        guidance=np.float32(3.0), temperature=np.float32(1.0), top_p=np.float32(0.95),
    )
    if with_bf16:
        # the reference's own production precision, teacher-forced with the fp32 run's tokens
        t0 = time.time()
        torch.set_grad_enabled(False)
        mb = R.build_reference_model(sd, dims, torch.bfloat16)
        lb = []
        spk_b = spk.to(torch.bfloat16)
        lg = mb(prompt.view(1, -1).repeat(2, 1), spk_b, torch.arange(T))
        lb.append(lg[:, -1].float().clone())
        last = max(keep_steps)
        for s in range(1, last + 1):
            tok = gen[s - 1].view(1, 1).repeat(2, 1)
            lg = mb(tok, spk_b, torch.tensor([T + s - 1]))
            lb.append(lg[:, -1].float().clone())
        out["logits_ref_bf16"] = torch.stack([lb[s] for s in keep_steps]).numpy()
        print(f"[{tag}] reference bf16 teacher-forced: {time.time()-t0:.1f}s")
    np.savez_compressed(os.path.join(GOLD, f"stage1_{tag}.npz"), **out)
    rel = []
    if with_bf16:
        for a, b in zip(out["logits"], out["logits_ref_bf16"]):
            rel.append(float(np.abs(a - b).max() / np.abs(a).max()))
        print(f"[{tag}] reference bf16-vs-fp32 max-norm rel gap per kept step: {rel}")


def golden_sampler(n_cases=8, V=2562):
    """Known-answer vectors for fast_inference_utils.py:61-120 straight from the reference."""
    fiu = R.reference_functions()
    g = torch.Generator().manual_seed(99)
    cases = []
    combos = [(3.0, 1.0, 0.95, None), (1.0, 1.0, 0.9, None), (2.0, 0.7, 0.99, None), (3.0, 0.0, 0.95, None),
              (1.5, 1.3, None, None), (3.0, 1.0, 0.95, 50), (2.5, 0.5, None, 1), (3.0, 1.0, 0.5, None)]
    logits_all, q_all, probs_all, idx_all, par_all = [], [], [], [], []
    for c in range(n_cases):
        gs, temp, tp, tk = combos[c % len(combos)]
        scale = [0.05, 1.0, 4.0][c % 3]
        logits = torch.randn(2, 1, V, generator=g) * scale
        torch.manual_seed(1000 + c)
        with _Recorder(fiu) as rec:
            idx, probs = fiu.sample(logits, guidance_scale=torch.tensor(gs), temperature=torch.tensor(temp),
                                    top_p=None if tp is None else torch.tensor(tp), top_k=tk)
        logits_all.append(logits[:, 0]); q_all.append(rec.q[0]); probs_all.append(probs.float())
        idx_all.append(int(idx)); par_all.append([gs, temp, -1.0 if tp is None else tp, 0 if tk is None else tk])
    np.savez_compressed(os.path.join(GOLD, "sampler.npz"),
                        logits=torch.stack(logits_all).numpy(), noise=torch.stack(q_all).numpy(),
                        probs=torch.stack(probs_all).numpy(), idx=np.asarray(idx_all, np.int32),
                        params=np.asarray(par_all, np.float32))
    print(f"[sampler] {n_cases} cases")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-full", action="store_true")
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    golden_sampler()
    golden_stage1(synth.TINY, "tiny", T=12, n_new=24, keep_steps=[0, 1, 2, 3, 8, 23])
    if not a.skip_full:
        golden_stage1(synth.FULL, "full", T=48, n_new=256, keep_steps=[0, 1, 8, 64, 255])
