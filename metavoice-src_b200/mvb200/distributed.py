"""Multi-GPU plumbing: replicas.  One process per GPU; rank 0 builds the bf16 weight arena, ONE
broadcast (NCCL over NVLink on GPUs, gloo in the CPU tests) replicates it, utterances are sharded across ranks and
there is no steady-state collective.  The reference itself has no distributed code at all."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_utterances(n_total: int, rank: int, world: int) -> List[int]:
    """Round-robin utterance ids of this rank (cf. the length-balanced batching of mixins/causal.py:290-338)."""
    return list(range(rank, n_total, world))


def broadcast_arena(arena: Optional[torch.Tensor], offsets: Optional[Sequence[int]], device, src: int = 0) -> Tuple[torch.Tensor, List[int], float]:
    """Replicate (arena bytes, offset table) from `src`; returns (arena on `device`, offsets, broadcast ms)."""
    import time
    meta = [None if offsets is None else list(offsets), None if arena is None else int(arena.numel())]
    dist.broadcast_object_list(meta, src=src)
    offsets, nbytes = meta
    if arena is None:
        arena = torch.empty(nbytes, dtype=torch.uint8, device=device)
    else:
        arena = arena.to(device)
    dist.barrier()
    use_cuda = torch.device(device).type == "cuda"
    if use_cuda:
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(arena, src=src)
        e1.record()
        torch.cuda.synchronize(device)
        ms = e0.elapsed_time(e1)
    else:
        t0 = time.perf_counter()
        dist.broadcast(arena, src=src)
        ms = (time.perf_counter() - t0) * 1e3
    return arena, offsets, ms


def max_over_ranks(values: Sequence[float], device) -> List[float]:
    """Device-side timings are reported as the max over ranks (never wall clock of one rank)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def gather_token_lists(local: List[Tuple[int, List[int]]], world: int) -> List[Tuple[int, List[int]]]:
    """Collect (utterance id, tokens) from every rank on rank 0 (used for verification, not on the timed path)."""
    out = [None] * world
    dist.all_gather_object(out, local)
    return sorted([x for part in out for x in part])
