"""Row N4: text chunker (CPU) and continuous batching into free KV slots (GPU, through the C ABI)."""
import pytest
import torch

from mvb200 import synth
from mvb200.serving import MAX_CHARS, chunk_text


def test_chunker_respects_limit_and_preserves_text():
    text = ("This is the first sentence. Here is a second one, with a clause; and a third part: really! "
            "Is it a question? " + "word " * 70 + "end. " + "x" * 500 + " tail.")
    chunks = chunk_text(text)
    assert all(0 < len(c) <= MAX_CHARS for c in chunks)
    assert " ".join(chunks).replace(" ", "") == " ".join(text.split()).replace(" ", "")
    assert chunk_text("short text") == ["short text"] and chunk_text("   ") == []
    # sentence boundaries are preferred over mid-sentence cuts
    two = chunk_text("A" * 150 + ". " + "B" * 150 + ".")
    assert two == ["A" * 150 + ".", "B" * 150 + "."]
    # 60 s of speech at ~15 chars/s is ~900 characters -> a handful of chunks (BASELINE configs[3])
    long = " ".join(["The quick brown fox jumps over the lazy dog near the quiet river bank."] * 13)
    assert 4 <= len(chunk_text(long)) <= 7


@pytest.mark.gpu
@pytest.mark.parametrize("tc", ["", "BC"], ids=["cuda_core_exact", "persistent_kernel"])
def test_continuous_batching_equals_solo_runs(tc):
    """5 requests of different lengths through 2 KV slots: requests are admitted as slots free up, every utterance gets
    the tokens it gets when decoded alone (exact on the deterministic CUDA-core path; the persistent kernel's split-K
    red.add order may flip a near-tie, so there at most one sequence may differ)."""
    from mvb200 import fast_inference_utils as U
    from mvb200.fast_model import ModelArgs, Transformer
    from mvb200.serving import ContinuousBatcher
    d = synth.TINY
    sd = synth.stage1_state_dict(d, 0)
    cfg = ModelArgs(block_size=d.block_size, vocab_size=d.vocab_size, n_layer=d.n_layer, n_head=d.n_head, dim=d.dim)

    def mk(utts):
        m = Transformer.from_state_dict(sd, cfg, device="cuda:0")
        m.setup_caches(2 * utts, d.block_size, kv_dtype="bf16", tensor_core_path=tc)
        return m

    lens, news = [5, 17, 9, 30, 12], [40, 7, 70, 33, 20]
    prompts = [synth.synthetic_prompt(T, seed=100 + i) for i, T in enumerate(lens)]
    spks = [synth.synthetic_speaker(seed=200 + i) for i in range(5)]
    noise = [torch.empty(n, d.vocab_size).exponential_(1, generator=torch.Generator().manual_seed(300 + i)) for i, n in enumerate(news)]
    kw = dict(guidance_scale=2.5, temperature=1.0, top_p=0.9, end_of_audio_token=9999)
    solo = mk(1)
    want = [U.generate_batch(solo, [prompts[i]], spks[i], max_new_tokens=news[i], noise=noise[i][None], **kw)[0] for i in range(5)]
    solo.close()
    cb = ContinuousBatcher(mk(2), burst=16)
    ids = [cb.submit(prompts[i], spks[i], max_new_tokens=news[i], noise=noise[i], **kw) for i in range(3)]
    first = cb.step()                                   # requests 0 and 1 occupy the two slots; request 1 (7 tokens) finishes
    assert first == [ids[1]] and len(cb._active) == 1 and len(cb._pending) == 1
    ids += [cb.submit(prompts[i], spks[i], max_new_tokens=news[i], noise=noise[i], **kw) for i in (3, 4)]   # late arrivals
    got = cb.run_until_done()
    assert sorted(got) == sorted(ids) and cb.idle
    diff = [i for i in range(5) if got[ids[i]].tolist() != want[i].tolist()]
    assert all(len(got[ids[i]]) == news[i] for i in range(5))
    assert len(diff) <= (0 if tc == "" else 1), diff
