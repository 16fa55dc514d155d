"""Work partitioning for the encode path: length-bucketed batching and rank sharding.

The reference parallelises only by shard id, one independent SLURM job per GPU
(REF/yodas2-mimi/submit/job_template.sh:6-10); utterances are encoded independently, so there is no
exchange step on the hot path. Here: one process per GPU, batches dealt round-robin to ranks, and a single
NCCL/gloo all_reduce of a small counter vector at the end of a run.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def bucket_batches(lengths: Sequence[int], batch_size: int, bucket_width: int = 2 * 24000) -> List[List[int]]:
    """Group item indices into batches of <= batch_size items of similar length.

    Items are keyed by ``length // bucket_width`` (2 s buckets at 24 kHz by default), buckets are walked
    from short to long and, inside a bucket, in the original (file) order, so padding waste per batch is
    bounded by one bucket width except where a batch straddles two buckets."""
    order = sorted(range(len(lengths)), key=lambda i: (lengths[i] // bucket_width, i))
    return [order[i:i + batch_size] for i in range(0, len(order), batch_size)]


def shard_for_rank(n_units: int, rank: int, world_size: int) -> List[int]:
    """Unit j (a batch, a shard id, a sub-shard) belongs to rank ``j % world_size``."""
    return list(range(rank, n_units, world_size))


def padding_waste(lengths: Sequence[int], batches: Sequence[Sequence[int]]) -> float:
    """Fraction of padded samples over all batches (0 = no padding)."""
    padded = sum(max(lengths[i] for i in b) * len(b) for b in batches if b)
    real = sum(lengths[i] for b in batches for i in b)
    return 0.0 if padded == 0 else 1.0 - real / padded


def reduce_counters(counters: Dict[str, float], device: torch.device | str = "cpu") -> Dict[str, float]:
    """SUM-all_reduce a dict of end-of-run counters (audio seconds, frames, bytes, mismatches); keys ending
    in ``_max`` are MAX-reduced (e.g. elapsed time). No-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(counters)
    keys = sorted(counters)
    sums = torch.tensor([float(counters[k]) for k in keys if not k.endswith("_max")], dtype=torch.float64, device=device)
    maxs = torch.tensor([float(counters[k]) for k in keys if k.endswith("_max")], dtype=torch.float64, device=device)
    if sums.numel():
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    if maxs.numel():
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    out, si, mi = {}, 0, 0
    for k in keys:
        if k.endswith("_max"):
            out[k] = float(maxs[mi]); mi += 1
        else:
            out[k] = float(sums[si]); si += 1
    return out
