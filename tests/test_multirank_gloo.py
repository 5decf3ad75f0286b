"""CPU suite, part 4: the N > 1 host logic under torch.distributed (gloo, world_size 2 and 4, CPU tensors):
row-block partitioning + in-place position all-gather, ensemble job assignment, and that a row-decomposed force
evaluation assembled from the ranks' row blocks equals the single-rank result (here the per-rank compute is the
oracle restatement standing in for the GPU kernel: the test is about the sharding plumbing, which is device-agnostic)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mdqtplasmasims_b200 import sharding
    from oracle import pyoracle as po
    orc = po.Oracle()
    L = (n * 4 * np.pi / 3) ** (1. / 3)
    lDeb = 1 / np.sqrt(0.3)
    rng = np.random.default_rng(42)           # same seed on every rank: the global initial state
    R0 = rng.uniform(0, L, size=(3, n))
    row0, rows = sharding.row_block(n, world, rank)
    ld = ((n + 31) // 32) * 32
    # each rank only "knows" its own rows after a local update; the others are stale (NaN) until the all-gather
    R = torch.full((3, ld), float("nan"), dtype=torch.float64)
    R[:, row0:row0 + rows] = torch.from_numpy(R0[:, row0:row0 + rows] + 0.001 * (rank + 1))
    sharding.allgather_positions(R, n, world, rank, dist)
    Rg = R[:, :n].numpy()
    expect = R0.copy()
    for r in range(world):
        a, b = sharding.row_block(n, world, r)
        expect[:, a:a + b] += 0.001 * (r + 1)
    assert np.array_equal(Rg, expect)
    # row-decomposed forces: rank computes rows [row0,row0+rows) against ALL j
    F_full = orc.forces_su(np.ascontiguousarray(Rg), L, lDeb)
    mine = torch.from_numpy(np.ascontiguousarray(F_full[:, row0:row0 + rows]))
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    F_asm = np.concatenate([g.numpy() for g in gathered], axis=1)
    assert np.array_equal(F_asm, F_full)
    # partial observables summed over ranks (output steps)
    ekin_part = float((Rg[:, row0:row0 + rows] ** 2).sum())
    tot = sharding.allreduce_scalars([ekin_part, rows], dist)
    assert abs(tot[0] - (Rg ** 2).sum()) < 1e-9 * (Rg ** 2).sum() and tot[1] == n
    # output() observables of a row-decomposed run: two small all-reduces (a numpy stand-in plays the engine's partial sums)
    Vg = np.random.default_rng(7).normal(size=(3, n))

    class _Rows:
        def diag_partial(self, vx_mean=None):
            m = 0.0 if vx_mean is None else float(vx_mean)
            v = Vg[:, row0:row0 + rows]
            return np.array([v[0].sum(), 0.5 * ((v[0] - m) ** 2).sum(), 0.5 * (v[1] ** 2).sum(), 0.5 * (v[2] ** 2).sum(), rows / n])

        def vel_dist_partial(self, vx_mean):
            return np.full((3, 2001), float(rows))

    d = sharding.distributed_diagnostics(_Rows(), n, dist, want_vel_dist=True)
    assert abs(d["vx_avg"] - Vg[0].mean()) < 1e-15 and abs(d["ekin_x"] - 0.5 * ((Vg[0] - Vg[0].mean()) ** 2).mean()) < 1e-14
    assert abs(d["ekin_y"] - 0.5 * (Vg[1] ** 2).mean()) < 1e-14 and abs(d["epot"] - 1.0) < 1e-15 and (d["pvel"] == n).all()
    jobs = sharding.ensemble_jobs(10, world, rank)
    all_jobs = [None] * world
    dist.all_gather_object(all_jobs, jobs)
    assert sorted(sum(all_jobs, [])) == list(range(1, 11))
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_row_decomposition_and_ensembles_under_gloo(world, tmp_path):
    n = 256
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


def test_partition_helpers():
    from mdqtplasmasims_b200 import sharding
    assert sharding.row_block(1000000, 8, 3) == (375000, 125000)
    with pytest.raises(ValueError):
        sharding.row_block(10, 4, 0)
    assert sharding.padded_ions(3500, 8) == 3504
    assert [len(sharding.ensemble_jobs(512, 8, r)) for r in range(8)] == [64] * 8
    assert sharding.ensemble_jobs(5, 2, 0) == [1, 2, 3] and sharding.ensemble_jobs(5, 2, 1) == [4, 5]


def _worker_exchange(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mdqtplasmasims_b200 import sharding
    rng = np.random.default_rng(7)
    truth = rng.integers(-2 ** 62, 2 ** 62, size=(3, n), dtype=np.int64)   # fixed-point positions after the step, all ranks' rows
    row0, n_rows, R = sharding.row_block_ceil(n, world, rank)
    X = np.full((3, ((n + 31) // 32) * 32), -1, dtype=np.int64)             # this rank knows only its own rows
    X[:, row0:row0 + n_rows] = truth[:, row0:row0 + n_rows]
    sharding.exchange_rows(X, n, world, rank, dist)
    ok = np.array_equal(X[:, :n], truth) and np.all(X[:, n:] == -1)
    blocks = [sharding.row_block_ceil(n, world, r)[:2] for r in range(world)]
    cover = sum(b[1] for b in blocks) == n and all(blocks[r][0] + blocks[r][1] == (blocks[r + 1][0] if r + 1 < world else n) or blocks[r][1] == 0
                                                    for r in range(world))
    np.save(os.path.join(out_dir, "x%d.npy" % rank), np.array([ok, cover]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1000), (3, 1000), (4, 1201)])
def test_row_exchange_protocol_with_unequal_blocks(tmp_path, world, n):
    """The exchange protocol of csrc/mdqt_comm.cu (pack own rows -> ONE all-gather of padded [3][R] blocks -> unpack the remote rows),
    mirrored in sharding.py: with R = ceil(N / world) and a shorter last block, every rank ends with everybody's rows and nothing
    else is written. gloo, CPU tensors, int64 payload (the fixed-point positions)."""
    port = _free_port()
    mp.spawn(_worker_exchange, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(os.path.join(str(tmp_path), "x%d.npy" % r)).all()
