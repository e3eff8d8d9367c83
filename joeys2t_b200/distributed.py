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


class NcclCommunicator:
    """``ncclComm_t`` owned by the C ABI (``js2t_nccl_comm_create``), for the path's one collective.

    The 128-byte NCCL unique id is drawn by rank 0 and broadcast over the existing
    ``torch.distributed`` process group (the only thing torch is used for here); the communicator
    itself and the all-reduce are the library's own ``ncclCommInitRank`` / ``ncclAllReduce`` calls.
    """

    def __init__(self, device: int, group=None):
        import ctypes
        from joeys2t_b200 import _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("NcclCommunicator needs an initialised torch.distributed process group")
        self._lib = _lib.load()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = int(device)
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (ctypes.c_ubyte * 128)()
            _lib.check(self._lib.js2t_nccl_unique_id(buf))
            uid = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
        on_gpu = dist.get_backend(group) == "nccl"
        t = uid.to(f"cuda:{self.device}") if on_gpu else uid
        dist.broadcast(t, src=0, group=group)
        raw = bytes(t.cpu().numpy().tobytes())
        self._h = ctypes.c_void_p()
        _lib.check(self._lib.js2t_nccl_comm_create(raw, self.world, self.rank, self.device,
                                                   ctypes.byref(self._h)))

    def allreduce_global_stats(self, accum: torch.Tensor) -> torch.Tensor:
        """In-place SUM all-reduce of the (161,) float64 accumulator on the current stream."""
        from joeys2t_b200 import _lib
        assert accum.is_cuda and accum.dtype == torch.float64 and accum.numel() == ACCUM_LEN
        stream = torch.cuda.current_stream(accum.device).cuda_stream
        _lib.check(self._lib.js2t_global_stats_allreduce(self._h, accum.data_ptr(), stream))
        return accum

    def close(self):
        if getattr(self, "_h", None):
            self._lib.js2t_nccl_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass


def allreduce_global_stats(accum: torch.Tensor, group=None,
                           comm: Optional["NcclCommunicator"] = None) -> torch.Tensor:
    """The path's single collective: in-place SUM all-reduce of the 161 float64 statistics,
    enqueued on the current stream — no host synchronisation.  With ``comm`` it goes through the C ABI
    (``js2t_global_stats_allreduce`` -> ``ncclAllReduce``); without, through ``torch.distributed``
    (the route of the reference's own ``ddp_reduce``, helpers_for_ddp.py:157-174; ``gloo`` in the CPU
    tests)."""
    assert accum.dtype == torch.float64 and accum.numel() == ACCUM_LEN
    if comm is not None:
        return comm.allreduce_global_stats(accum)
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
