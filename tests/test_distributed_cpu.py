# coding: utf-8
"""
N > 1 host logic on CPU (gloo, world_size 2): utterance sharding and the path's single exchange
step, the all-reduce of the 161 float64 global-CMVN statistics.  The per-rank statistics come from
the oracle here (no GPU); the GPU test ``test_global_cmvn_two_pass`` covers the device side.
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from joeys2t_b200 import distributed as D  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, _, w = D.init_from_env(backend="gloo")
    z = np.load(ROOT / "tests" / "golden" / "ref_fbank.npz")
    feats = [z[f"fbank{i}"] for i in range(10)]
    mine = D.shard_utterances([f.shape[0] for f in feats], r, w, mode=mode)
    acc = D.new_accumulator()
    for i in mine:
        x = feats[i].astype(np.float64)
        acc[:80] += torch.from_numpy(x.sum(0))
        acc[80:160] += torch.from_numpy((x * x).sum(0))
        acc[160] += x.shape[0]
    D.allreduce_global_stats(acc)
    mean, istd = D.stats_to_mean_istd(acc)
    q.put((r, mine, mean, istd, float(acc[160])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["strided", "balanced"])
def test_global_cmvn_allreduce_world2(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    z = np.load(ROOT / "tests" / "golden" / "ref_fbank.npz")
    allx = np.concatenate([z[f"fbank{i}"] for i in range(10)], 0).astype(np.float64)
    mu = allx.mean(0)
    std = np.sqrt(np.maximum((allx**2).mean(0) - mu**2, 1e-10))
    shards = sorted(sum((g[1] for g in got), []))
    assert shards == list(range(10))  # a partition: every utterance exactly once
    for _, _, mean, istd, n in got:
        assert n == allx.shape[0]
        np.testing.assert_allclose(mean, mu, rtol=0, atol=1e-9)
        np.testing.assert_allclose(istd, 1.0 / std, rtol=1e-9)
    # both ranks hold bit-identical statistics after the all-reduce
    assert np.array_equal(got[0][2], got[1][2]) and np.array_equal(got[0][3], got[1][3])


def test_sharding_modes():
    costs = [5, 1, 9, 3, 7, 2, 8]
    assert D.shard_utterances(costs, 1, 3) == [1, 4]
    parts = [D.shard_utterances(costs, r, 3, mode="balanced") for r in range(3)]
    assert sorted(sum(parts, [])) == list(range(7))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(costs)
    with pytest.raises(ValueError):
        D.shard_utterances(costs, 0, 2, mode="nope")


def test_stats_to_mean_istd_flags():
    acc = np.zeros(161)
    acc[:80] = 10.0
    acc[80:160] = 60.0
    acc[160] = 5.0
    mean, istd = D.stats_to_mean_istd(acc)
    np.testing.assert_allclose(mean, 2.0)
    np.testing.assert_allclose(istd, 1.0 / np.sqrt(12.0 - 4.0))
    mean, istd = D.stats_to_mean_istd(acc, norm_means=False, norm_vars=False)
    assert (mean == 0).all() and (istd == 1).all()
    with pytest.raises(ValueError):
        D.stats_to_mean_istd(np.zeros(161))
