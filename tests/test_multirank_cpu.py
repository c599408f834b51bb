"""CPU, world_size 2 over gloo: the N>1 host logic (arena broadcast, utterance sharding, max-over-ranks timing)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "metavoice-src_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mvb200 import synth
    from mvb200.distributed import broadcast_arena, gather_token_lists, max_over_ranks, shard_utterances
    from mvb200.fast_model import pack_arena
    from oracle import stage1_port as P
    d = synth.TINY
    arena = offsets = None
    if rank == 0:
        arena, offsets = pack_arena(synth.stage1_state_dict(d, 0), d.n_layer)
    arena, offsets, ms = broadcast_arena(arena, offsets, "cpu")
    # every rank rebuilds the state dict view from the replicated bytes and decodes ITS utterances with the oracle
    sd_ref = synth.stage1_state_dict(d, 0)
    keys = ["transformer.wtes.0.weight", "transformer.wpe.weight", "speaker_cond_pos.weight", "transformer.ln_f.weight", "lm_heads.0.weight"]
    for k, o in zip(keys, offsets[:5]):
        t = sd_ref[k]
        assert torch.equal(arena[o:o + t.numel() * 2].view(torch.bfloat16).view_as(t), t), f"rank {rank}: {k} differs after broadcast"
    mine = shard_utterances(5, rank, world)
    m = P.Stage1Oracle(sd_ref, d.n_head, d.norm_eps, torch.float32, faithful_full_cache=False)
    local = []
    for u in mine:
        m.setup_caches()
        torch.manual_seed(100 + u)
        y = P.generate(m, synth.synthetic_prompt(6 + u, seed=u), synth.synthetic_speaker(seed=u), max_new_tokens=4,
                       end_of_audio_token=9999, guidance_scale=3.0, temperature=1.0, top_p=0.95)
        local.append((u, y.tolist()))
    allt = gather_token_lists(local, world)
    worst = max_over_ranks([float(rank + 1), ms], "cpu")
    if rank == 0:
        q.put((mine, [u for u, _ in allt], worst[0]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_replica_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    mine0, all_ids, worst = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert mine0 == [0, 2, 4] and all_ids == [0, 1, 2, 3, 4] and worst == 2.0
