# coding: utf-8
"""
Multi-GPU plumbing of the front-end: one process per GPU, utterances sharded across ranks with no
data-path collective (utterance CMVN and SpecAugment are per utterance), and exactly one exchange
step for **global CMVN**: an all-reduce (SUM) of 161 float64 — per-bin sum[80], sum of
squares[80], frame count — over NCCL/NVLink (``torch.distributed``; ``gloo`` in the CPU tests).

The reference's only collectives live in ``joeynmt/helpers_for_ddp.py`` (none on the audio path);
its rank-strided sharding (``helpers_for_ddp.py:319``: ``indices[rank::world]``) is reused here.
The statistics → (mean, 1/std) formula is the reference CMVN's (``data_augmentation.py:98-105``).
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

NUM_MEL = 80
ACCUM_LEN = 2 * NUM_MEL + 1


def shard_utterances(costs: Sequence[int], rank: int, world: int,
                     mode: str = "strided") -> List[int]:
    """Indices of the utterances this rank processes.

    ``strided``: ``range(n)[rank::world]`` like the reference's ``DistributedSubsetSampler``;
    ``balanced``: greedy longest-first assignment on ``costs`` (frames or samples), deterministic.
    """
    n = len(costs)
    if mode == "strided":
        return list(range(n))[rank::world]
    if mode == "balanced":
        order = sorted(range(n), key=lambda i: (-int(costs[i]), i))
        loads = [0] * world
        owner = [0] * n
        for i in order:
            r = min(range(world), key=lambda k: (loads[k], k))
            owner[i] = r
            loads[r] += int(costs[i])
        return [i for i in range(n) if owner[i] == rank]
    raise ValueError(f"unknown sharding mode {mode!r}")


def new_accumulator(device=None) -> torch.Tensor:
    """Zeroed (161,) float64 accumulator: sum[80] | sumsq[80] | frames."""
    return torch.zeros(ACCUM_LEN, dtype=torch.float64, device=device)


def allreduce_global_stats(accum: torch.Tensor, group=None) -> torch.Tensor:
    """The path's single collective: in-place SUM all-reduce of the 161 float64 statistics,
    enqueued on the current stream (NCCL) — no host synchronisation."""
    assert accum.dtype == torch.float64 and accum.numel() == ACCUM_LEN
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(accum, op=dist.ReduceOp.SUM, group=group)
    return accum


def stats_to_mean_istd(accum, norm_means: bool = True,
                       norm_vars: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """(mean[80], 1/std[80]) in float64 from the all-reduced statistics —
    ``var = sumsq/n - mean**2; std = sqrt(max(var, 1e-10))`` (data_augmentation.py:98-105)."""
    a = accum.detach().cpu().numpy() if isinstance(accum, torch.Tensor) else np.asarray(accum)
    a = a.astype(np.float64)
    n = a[2 * NUM_MEL]
    if n <= 0:
        raise ValueError("global CMVN statistics are empty")
    mu = a[:NUM_MEL] / n
    var = a[NUM_MEL:2 * NUM_MEL] / n - mu**2
    std = np.sqrt(np.maximum(var, 1e-10))
    mean = mu if norm_means else np.zeros(NUM_MEL)
    istd = 1.0 / std if norm_vars else np.ones(NUM_MEL)
    return mean, istd


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world) from the torchrun environment; initialises the process group
    when WORLD_SIZE > 1 (NCCL on GPUs, gloo otherwise)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, local_rank, world
