"""Host-side multi-GPU logic on the CPU: gloo backend, world_size 2 (the NCCL path runs the same code on the GPU box).
Covers the column sharding of bench.py and the one-collective-per-step reduction fold."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpp_b200 import parallel as PL


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ncol_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        c0, c1 = PL.shard_range(ncol_total, rank, world)
        d = bench.shard_inputs(c0, c1, chunk=64)
        # what the library would leave in its reduction buffer for this shard (mass sums, maxima, worst reason)
        mass = d["watsat"].sum(axis=1)
        local = torch.tensor([mass.sum(), 2.0 * mass.sum(), float(d["infil"].sum()), 0.0,
                              float(mass.max()), float(3 + rank), float(rank == 1), float(rank), 3.0 - 8.0 * rank], dtype=torch.float64)
        gr = PL.GlobalReductions(local)
        for _ in range(3):                                   # repeated steps reuse the same buffers
            res = gr.step()
        q.put((rank, c0, c1, d["watsat"].sum(), float(d["infil"].sum()), float(mass.max()), res.tolist(), gr.as_dict()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ncol_total", [256, 250 + 7])
def test_sharding_and_reductions_world2_gloo(ncol_total):
    import bench
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ncol_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shards tile the batch exactly and reproduce the single-process batch bit for bit
    assert outs[0][1] == 0 and outs[0][2] == outs[1][1] and outs[1][2] == ncol_total
    full = bench.shard_inputs(0, ncol_total, chunk=64)
    assert outs[0][3] + outs[1][3] == pytest.approx(full["watsat"].sum(), rel=1e-15)
    mass = full["watsat"].sum(axis=1)
    for rank, c0, c1, wsum, infil, mmax, res, dd in outs:
        assert res[0] == pytest.approx(mass.sum(), rel=1e-14) and res[1] == pytest.approx(2 * mass.sum(), rel=1e-14)
        assert res[2] == pytest.approx(full["infil"].sum(), rel=1e-13)
        assert res[4] == mass.max() and res[5] == 4.0 and res[6] == 1.0 and res[7] == 1.0
        assert res[8] == -5.0                               # worst SNES reason over ranks (a diverged rank wins)
        assert dd["any_diverged"] is True and dd["worst_reason"] == -5 and dd["max_newton_its"] == 4
    assert outs[0][6] == outs[1][6]                         # every rank holds the same folded result


def test_shard_range_edges():
    assert PL.shard_range(10, 0, 1) == (0, 10)
    covered = []
    for r in range(8):
        c0, c1 = PL.shard_range(4194304, r, 8)
        assert c1 - c0 == 524288
        covered.append((c0, c1))
    assert covered[0][0] == 0 and covered[-1][1] == 4194304 and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    sizes = [PL.shard_range(10, r, 4) for r in range(4)]
    assert [b - a for a, b in sizes] == [2, 3, 2, 3] and sizes[-1][1] == 10
    with pytest.raises(ValueError):
        PL.shard_range(10, 4, 4)


def test_fold_single_process():
    g = torch.tensor([[1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 0.0, 1.0, 3.0], [10.0, 20.0, 30.0, 40.0, 4.0, 7.0, 1.0, 0.0, 4.0]], dtype=torch.float64)
    out = PL.fold(g).tolist()
    assert out == [11.0, 22.0, 33.0, 44.0, 5.0, 7.0, 1.0, 1.0, 3.0]
    gr = PL.GlobalReductions(g[0].clone())
    assert gr.step().tolist() == g[0].tolist()
