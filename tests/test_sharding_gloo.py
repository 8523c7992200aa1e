"""CPU, world_size 2, gloo: the N>1 host logic of bench.py / multi-GPU predict -- slabs cover the query
set exactly once, the timing reduction is the max over ranks, results gathered in rank order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from inference_tools_b200.sharding import round_robin, shard_range


def test_shard_range_properties():
    for n in (0, 1, 7, 1 << 20, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sorted(sum((round_robin(11, r, 4) for r in range(4)), [])) == list(range(11))
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_block_cyclic_plan_tiles_the_matrix_once():
    """host side of the distributed Cholesky (dist.cu): every block column has exactly one owner, owners are cyclic,
    and the per-rank panel storage adds up to the (augmented) lower block triangle."""
    from inference_tools_b200 import _lib
    for n, block, world in ((131072, 1024, 8), (5000, 256, 3), (700, 128, 2), (128, 128, 4)):
        plans = [_lib.dist_plan(n, block, world, r) for r in range(world)]
        nblk = plans[0]["n_blocks"]
        npad = (n + 127) // 128 * 128
        assert nblk == -(-npad // block)
        assert all(p["owners"] == [j % world for j in range(nblk)] for p in plans)
        assert sum(p["n_owned"] for p in plans) == nblk
        assert max(p["n_owned"] for p in plans) - min(p["n_owned"] for p in plans) <= 1
        # each panel: its rows of the augmented lower block triangle + the inverted 128-blocks of its diagonal block
        expect = sum((npad + 256 - j * block) * block + block * 128 for j in range(nblk))   # 256 augmented rows
        assert sum(p["panel_doubles"] for p in plans) == expect
        assert all(p["staging_doubles"] == 2 * ((npad + 256) * block + block * 128) for p in plans)
    big = _lib.dist_plan(131072, 1024, 8, 0)
    assert (big["panel_doubles"] + big["staging_doubles"]) * 8 < 12e9          # fits one B200 many times over
    with pytest.raises(_lib.EngineError):
        _lib.dist_plan(100, 100, 2, 0)


def _worker(rank, world, port, m, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(m, rank, world)
    q = np.arange(m, dtype=np.float64)
    local = torch.from_numpy(q[lo:hi] * 2.0)                       # stand-in for this rank's predictions
    sizes = [shard_range(m, r, world) for r in range(world)]
    bufs = [torch.empty(h - l, dtype=torch.float64) for l, h in sizes]
    dist.all_gather(bufs, local) if len({b.numel() for b in bufs}) == 1 else None
    t = torch.tensor([0.5 + rank], dtype=torch.float64)            # per-rank elapsed time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    # distributed fit without hyperpars (gp/regression.py): every rank derives the optimiser's start points from the
    # replicated targets alone, so the ranks -- whose numpy global generators differ -- take identical steps
    from inference_tools_b200.gp import GpRegressor
    np.random.seed(1000 + rank)
    holder = type("Replica", (), {"y": np.sin(np.arange(50.0))})()
    starts = np.random.RandomState(GpRegressor._lockstep_seed(holder)).random_sample(size=6)
    gathered_starts = [None] * world
    dist.all_gather_object(gathered_starts, starts.tolist())
    if rank == 0:
        out.put((float(t), int(cnt), [b.tolist() for b in bufs] if len({b.numel() for b in bufs}) == 1 else None, gathered_starts))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_predict_bookkeeping():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    m = 64
    procs = [ctx.Process(target=_worker, args=(r, 2, port, m, out)) for r in range(2)]
    [p.start() for p in procs]
    tmax, total, gathered, starts = out.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert tmax == 1.5 and total == m
    assert np.allclose(np.concatenate(gathered), np.arange(m) * 2.0)
    assert starts[0] == starts[1] and len(starts[0]) == 6              # lockstep start points agree across ranks
