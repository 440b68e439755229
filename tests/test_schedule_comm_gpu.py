"""Scheduling and reduction plumbing that must not change results: the cost-ordered launch (mppgpu_set_column_ordering) and the
global reductions behind the C ABI (mppgpu_comm_init / mppgpu_global_mass_balance)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import problems as PB
from mpp_b200 import constants as K

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mpp():
    import mpp_b200
    from mpp_b200._lib import lib
    assert lib().mppgpu_device_count() > 0, "no CUDA device: the gpu tests must run on the B200 box"
    return mpp_b200


def _run(mpp, d, ordering, nsteps, coupled_chunks=None):
    p, ids = PB.build_elm_vsfm(mpp.VSFM, d)
    p.set_column_ordering(ordering)
    outs, stats = [], []
    for s in range(nsteps):
        if coupled_chunks is None:
            conv, reason, out = PB.elm_vsfm_step(p, ids, d, 1800.0, s + 1)
        else:
            ins = [(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, ids[n], np.ascontiguousarray(d[n])) for n in ("infil", "et", "dew", "drain", "snow", "sublim")]
            ins.append((K.AUXVAR_INTERNAL, K.VAR_FRAC_LIQ_SAT, 1, d["frac_liq"]))
            out = {k: np.empty(d["ncol"] * d["nlev"]) for k in ("sat", "mass", "smp", "pressure")}
            olist = [(K.AUXVAR_INTERNAL, v, 1, out[k]) for k, v in (("sat", K.VAR_LIQ_SAT), ("mass", K.VAR_MASS), ("smp", K.VAR_SOIL_MATRIX_POT), ("pressure", K.VAR_PRESSURE))]
            conv, reason = p.coupled_step(1800.0, s + 1, ins, olist, coupled_chunks)
            p.post_step_dt()
        outs.append({k: v.copy() for k, v in out.items()})
        stats.append({k: v.copy() for k, v in p.stats().items()})
    sums, maxs = p.mass_balance()
    p.close()
    return outs, stats, sums, maxs


@pytest.mark.parametrize("zwt_min", [1.0, 2.0])
def test_cost_ordered_launch_is_bit_identical_per_column(mpp, zwt_min):
    """Columns are independent: visiting them grouped by last step's residual-evaluation count changes which warp a column shares,
    never its arithmetic.  5000 columns x 5 steps (the order changes every step), every output and every per-column counter."""
    d = PB.elm_vsfm_inputs(5000, 15, zwt_min=zwt_min)
    o0, s0, sums0, maxs0 = _run(mpp, d, 0, 5)
    o1, s1, sums1, maxs1 = _run(mpp, d, 1, 5)
    for a, b in zip(o0, o1):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    for a, b in zip(s0, s1):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(maxs0, maxs1)                      # maxima are order-independent; the sums differ by summation order only
    assert np.allclose(sums0, sums1, rtol=1e-13, atol=0.0)


@pytest.mark.parametrize("satfunc", ["smooth_brooks_corey_bz3", "van_genuchten"])
def test_cost_ordered_launch_th_is_bit_identical_per_column(mpp, satfunc):
    """The TH register kernel visits its columns in the same cost order (two columns per warp): same solution, same counters, column
    for column.  van Genuchten from 1 m water tables brings dt cuts and very uneven costs into the batch."""
    d = PB.elm_th_inputs(3000, 15, satfunc=satfunc, zwt_min=(2.0 if satfunc != "van_genuchten" else 1.0))
    res = []
    for ordering in (0, 1):
        p, ids = PB.build_elm_th(mpp.TH, d)
        p.set_column_ordering(ordering)
        p.set_step_budget(3000)                              # (bounds the van Genuchten stragglers; same budget on both sides)
        outs, stats = [], []
        for s in range(4):
            conv, reason, out = PB.elm_th_step(p, ids, d, 1800.0, s + 1)
            outs.append(out); stats.append({k: v.copy() for k, v in p.stats().items()})
        res.append((outs, stats))
        p.close()
    for a, b in zip(res[0][0], res[1][0]):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    for a, b in zip(res[0][1], res[1][1]):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    assert max(st["nfuncs"].max() for st in res[0][1]) > 8   # the costs really differ


def test_cost_ordered_launch_in_the_chunked_coupling_step(mpp):
    d = PB.elm_vsfm_inputs(6144, 15, zwt_min=2.0)
    o0, s0, _, _ = _run(mpp, d, 0, 4, coupled_chunks=3)
    o1, s1, _, _ = _run(mpp, d, 1, 4, coupled_chunks=3)
    o2, s2, _, _ = _run(mpp, d, 1, 4)                       # unchunked separate calls
    for a, b, c in zip(o0, o1, o2):
        for k in a:
            assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k]), k


def test_global_mass_balance_single_rank_equals_local(mpp):
    d = PB.elm_vsfm_inputs(700, 15, zwt_min=2.0)
    p, ids = PB.build_elm_vsfm(mpp.VSFM, d)
    with pytest.raises(mpp.MPPError):
        p.global_mass_balance()                              # comm_init first
    p.comm_init(1, 0, None)
    PB.elm_vsfm_step(p, ids, d, 1800.0, 1)
    sums, maxs = p.mass_balance()
    g = p.global_mass_balance()
    assert [g["mass_begin"], g["mass_end"], g["source_dt"], g["boundary_exchanged"]] == list(sums)
    assert g["max_abs_mass_error"] == maxs[0] and g["max_newton_its"] == int(maxs[1]) and g["worst_reason"] in (2, 3, 4)
    assert g["max_abs_mass_error"] < 1e-5                    # the reference's gate (MPPVSFMALM_Driver.F90:140)
    p.close()


_WORKER = r"""
import os, sys, numpy as np
sys.path.insert(0, sys.argv[4]); sys.path.insert(0, os.path.join(sys.argv[4], "tests"))
rank, world, idfile = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
import time, json
import mpp_b200
from mpp_b200 import problems as PB, parallel as PL
if rank == 0:
    uid = mpp_b200.comm_unique_id()
    open(idfile + ".tmp", "wb").write(uid); os.rename(idfile + ".tmp", idfile)
else:
    while not os.path.exists(idfile):
        time.sleep(0.05)
    uid = open(idfile, "rb").read()
ncol = 4096
c0, c1 = PL.shard_range(ncol, rank, world)
d = PB.shard_inputs(c0, c1, chunk=512)
p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d, device=rank)
p.comm_init(world, rank, uid)
res = []
for s in range(2):
    PB.elm_vsfm_step(p, ids, d, 1800.0, s + 1)
    p.global_reduce_async()
    res.append(p.global_mass_balance())
sums, maxs = p.mass_balance()
json.dump({"global": res[-1], "local_sums": list(sums), "local_maxs": list(maxs)}, open(idfile + ".out%d" % rank, "w"))
p.close()
"""


def test_global_mass_balance_two_ranks_nccl_without_torch_distributed(mpp, tmp_path):
    """World size 2, one process per GPU, NO torch.distributed / gloo anywhere: the unique id travels through a file (an MPI host
    model would MPI_Bcast it) and the all-gather + fold run inside libmppgpu.so.  Both ranks must report the same global figures,
    equal to the fold of the two local ones and to a single-GPU run over all columns (sums to summation-order round-off)."""
    from mpp_b200._lib import lib
    if lib().mppgpu_device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import json
    idfile = str(tmp_path / "nccl_id")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", idfile, ROOT]) for r in range(2)]
    for pr in procs:
        assert pr.wait(timeout=300) == 0
    r0, r1 = (json.load(open(idfile + ".out%d" % r)) for r in range(2))
    assert r0["global"] == r1["global"]
    g = r0["global"]
    assert g["mass_begin"] == r0["local_sums"][0] + r1["local_sums"][0] and g["mass_end"] == r0["local_sums"][1] + r1["local_sums"][1]
    assert g["max_abs_mass_error"] == max(r0["local_maxs"][0], r1["local_maxs"][0])
    d = PB.shard_inputs(0, 4096, chunk=512)
    p, ids = PB.build_elm_vsfm(mpp.VSFM, d)
    p.comm_init(1, 0, None)
    for s in range(2):
        PB.elm_vsfm_step(p, ids, d, 1800.0, s + 1)
    one = p.global_mass_balance()
    p.close()
    assert abs(one["mass_end"] - g["mass_end"]) <= 1e-12 * abs(g["mass_end"])
    assert one["max_abs_mass_error"] == g["max_abs_mass_error"] and one["max_newton_its"] == g["max_newton_its"]
    assert one["worst_reason"] == g["worst_reason"] and not g["any_diverged"]


def test_entry_points_hand_the_callers_device_back(mpp):
    """A handle lives on the device it was created on; every call selects it and restores whatever device the caller had current
    (a host model, or torch in the same process, may be working on another one).  With one GPU the check is trivial but still runs."""
    import ctypes
    from mpp_b200._lib import lib
    rt = ctypes.CDLL("libcudart.so.12")
    cur = ctypes.c_int(-1)
    ndev = lib().mppgpu_device_count()
    mine, other = 0, (1 if ndev >= 2 else 0)
    assert rt.cudaSetDevice(mine) == 0
    d = PB.elm_vsfm_inputs(64)
    p, ids = PB.build_elm_vsfm(mpp.VSFM, d, device=other)
    assert rt.cudaGetDevice(ctypes.byref(cur)) == 0 and cur.value == mine
    conv, reason, out = PB.elm_vsfm_step(p, ids, d)
    assert conv
    assert rt.cudaGetDevice(ctypes.byref(cur)) == 0 and cur.value == mine
    p.close()
    assert rt.cudaGetDevice(ctypes.byref(cur)) == 0 and cur.value == mine
    if ndev >= 2:
        # same columns on the other device: same numbers
        q, qids = PB.build_elm_vsfm(mpp.VSFM, d, device=mine)
        conv2, reason2, out2 = PB.elm_vsfm_step(q, qids, d)
        assert np.array_equal(out["pressure"], out2["pressure"])
        q.close()
