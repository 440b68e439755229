#!/usr/bin/env python
"""bench.py -- headline benchmark of the VSFM hot path (BASELINE.json: soil column-timesteps/s, fp64, Newton-converged).

    python bench.py --gpus N --steps K --warmup W                    # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm on the host cores

Headline workload (BASELINE.json configs[3]): 4 Mi synthetic ELM-like soil columns x 15 layers PER GPU, van Genuchten curves,
Tanaka density, six COND_MASS_RATE source/sinks, dt = 1800 s, SNES tolerances = reference defaults.  Columns are independent and
sharded contiguously: rank r solves columns [r, r+1) x 4 Mi of ONE seeded global batch of 4N Mi columns (weak scaling with
DISTINCT columns on every rank; `mpp_b200/problems.py:shard_inputs` explains the batch and why its water tables start at 2 m).
One "step" = one ELM coupling step = PreStepDT + StepDT + PostStepDT over the whole shard (MPPVSFMALM_Driver.F90:603-935) + the
global mass-balance / convergence reduction (one ncclAllGather of 9 doubles per rank, issued by the library itself).

Rank 0 prints ONE JSON line.  `value` is timed with all inputs resident in HBM; `e2e` repeats the same steps through the C ABI
with pinned HOST buffers (SetDataFromCLM x7 in, GetDataForCLM x4 out, as MPPVSFMALM_Solve does).  `multi_gpu` adds, at every N,
the other two configurations BASELINE.json names: `strong_4Mi_total` (configs[3] literally: 4 Mi columns in total split N ways)
and `th_2Mi_total` (configs[4]: coupled thermal-hydrology, 2 Mi columns x 15 in total split N ways), each with its own roofline.
The reference itself (Fortran + PETSc + MPI) cannot be built in this image, so `cpu_baseline` / `--impl reference` time the
oracle -- the C restatement of the reference algorithm (oracle/, kind "port") -- on the box's host cores, compiled on the box
with -O3 -march=native for timing (the golden-vector build keeps -ffp-contract=off).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mpp_b200 import constants as K  # noqa: E402
from mpp_b200 import problems as PB  # noqa: E402

NLEV = 15
DT = 1800.0
CHUNK = PB.BENCH_CHUNK
ALG_BYTES_PER_COLSTEP = 1224       # SURVEY.md section 8(d), VSFM-VG base variant
ALG_BYTES_TH = 1824                # SURVEY.md section 8(d), TH with both source arrays
SS_NAMES = ("infil", "et", "dew", "drain", "snow", "sublim")
ZWT_MIN = PB.BENCH_ZWT_MIN


def shard_inputs(c0, c1, chunk=CHUNK, zwt_min=None):
    """Columns [c0, c1) of the global seeded VSFM batch (mpp_b200/problems.py:shard_inputs)."""
    return PB.shard_inputs(c0, c1, chunk=chunk, nlev=NLEV, zwt_min=ZWT_MIN if zwt_min is None else zwt_min)


def shard_inputs_th(c0, c1, chunk=CHUNK, satfunc="smooth_brooks_corey_bz3"):
    return PB.shard_inputs(c0, c1, chunk=chunk, nlev=NLEV,
                           builder=lambda n, nl, seed: PB.elm_th_inputs(n, nl, seed=seed, zwt_min=ZWT_MIN, satfunc=satfunc))


def measured_kernel_facts():
    """ncu-measured per-kernel figures written by tools/profile_summary.py --json (DRAM bytes per column-step, fp64 pipe share);
    nothing is typed into this file: a kernel without a committed capture reports null."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "roofline_measured.json")))
    except Exception:
        return {}


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled from a thread every few
    milliseconds (nvidia-smi -lms needs about a second to start, longer than the timed region of a short run)."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index, period_s=0.004):
        import threading
        self.sm, self.mask, self.smax, self.ok = [], 0, None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else gpu_index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            return
        self.period = period_s
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.smax, "reasons": [], "samples": 0}
        if not self.ok:
            return out
        self._stop.set()
        self.t.join(timeout=2)
        if self.sm:
            out.update(sm_mhz=float(np.median(self.sm)), reasons=[n for b, n in self.REASONS if self.mask & b], samples=len(self.sm))
        return out


def bind_to_gpu_numa_node(gpu_index):
    """Pin this process to the CPUs closest to its GPU (NVML topology) so that the pinned host buffers of the end-to-end
    path are allocated on the GPU's NUMA node; harmless when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[gpu_index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else gpu_index
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
    except Exception:
        pass


def set_forcing_host(p, ids, d):
    for name in SS_NAMES:
        p.set_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, ids[name], d[name])
    p.set_data(K.AUXVAR_INTERNAL, K.VAR_FRAC_LIQ_SAT, 1, d["frac_liq"])


def warp_waste(nf, group=4):
    """sum over warps of group * max(nf) / sum of nf in LAUNCH order `nf`: what a warp's columns waiting for its slowest one costs."""
    nf = np.asarray(nf, dtype=np.float64)
    n = (nf.size // group) * group
    return float(group * nf[:n].reshape(-1, group).max(axis=1).sum() / max(nf[:n].sum(), 1.0))


# --------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle, built on this box for speed
# --------------------------------------------------------------------------------------------------------------------
def timing_oracle():
    """oracle/ compiled here and now with -O3 -march=native (timing only; contraction allowed) into oracle/_native/: the golden build
    (oracle/Makefile, -ffp-contract=off, no -march) stays the checker.  Falls back to the golden build if gcc is unavailable."""
    from oracle import oracle as O
    odir = os.path.join(ROOT, "oracle")
    out = os.path.join(odir, "_native", "libmpp_oracle_native.so")
    flags = ["-O3", "-march=native", "-fPIC", "-fopenmp", "-shared"]
    try:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        srcs = sorted(os.path.join(odir, f) for f in os.listdir(odir) if f.endswith(".c"))
        subprocess.check_call(["gcc"] + flags + ["-o", out] + srcs + ["-lm"], stderr=subprocess.DEVNULL)
        O.use_library(out)
        return O, "gcc " + " ".join(flags) + " (built on this box)"
    except Exception as e:                                   # noqa: BLE001
        return O, "oracle/Makefile golden build (-O3 -ffp-contract=off); native build failed: %r" % (e,)


def cpu_baseline(steps, warmup, target_seconds=20.0):
    """The oracle (reference algorithm, per-column SNES) on a bounded column sample with every host thread."""
    O, build = timing_oracle()
    cores = os.cpu_count() or 1
    ncol = 4096
    d = shard_inputs(0, ncol)
    o, ids = PB.build_elm_vsfm(O.OracleVSFM, d, per_column=True, nthreads=cores)
    t0 = time.perf_counter()
    PB.elm_vsfm_step(o, ids, d, DT, 1)
    probe = time.perf_counter() - t0                       # first (most expensive) step on 4096 columns
    # size the sample so that warmup + steps take about target_seconds
    ncol = int(min(1 << 20, max(4096, 4096 * target_seconds / max(probe, 1e-3) / (steps + warmup) * 2.0)))
    ncol = (ncol // 4096) * 4096
    d = shard_inputs(0, ncol)
    o, ids = PB.build_elm_vsfm(O.OracleVSFM, d, per_column=True, nthreads=cores)
    for s in range(warmup):
        PB.elm_vsfm_step(o, ids, d, DT, s + 1)
    t0 = time.perf_counter()
    conv = True
    for s in range(steps):
        cv, reason, _ = PB.elm_vsfm_step(o, ids, d, DT, warmup + s + 1)
        conv = conv and cv
    el = time.perf_counter() - t0
    return {"value": ncol * steps / el, "unit": "column-timesteps/s", "cores": cores, "kind": "port", "sample_ncol": ncol,
            "sample": "%d of the benchmark's columns (global columns 0..%d of the same seeded batch), same %d warm-up + %d timed steps, "
                      "oracle/ C restatement of the reference algorithm (per-column SNES newtonls+bt, Thomas), OpenMP over columns; "
                      "the Fortran+PETSc reference cannot be built in this image" % (ncol, ncol - 1, warmup, steps),
            "build": build, "seconds": el, "converged": bool(conv)}, ncol, el


def workload_config(ncol_per_gpu, ngpus):
    """Identical for the CUDA arm and the reference arm (the reference arm times a bounded SAMPLE of this workload and says so in
    `cpu_baseline.sample` / `sample_ncol`)."""
    return {"workload": "VSFM Richards (BASELINE.json configs[3]): %d synthetic ELM-like soil columns x %d layers per GPU, van Genuchten-Mualem, "
                        "Tanaka density, 6 COND_MASS_RATE source/sinks, initial water table 2-20 m, dt=%.0f s, SNES rtol 1e-8 / stol 1e-10 / "
                        "max_it 50 (reference defaults), per-column Newton + bt line search + tridiagonal solve" % (ncol_per_gpu, NLEV, DT),
            "ncol_total": ncol_per_gpu * ngpus, "ncol_per_gpu": ncol_per_gpu, "nlev": NLEV, "dt_s": DT, "satfunc": "van_genuchten",
            "zwt_min_m": ZWT_MIN,
            "parallelism": "columns sharded contiguously over %d GPU(s): rank r solves columns [r, r+1) x %d of one seeded global batch; no "
                           "data-path collective; one ncclAllGather of 9 mass-balance/convergence doubles per rank per step "
                           "(mppgpu_global_reduce_async)" % (ngpus, ncol_per_gpu),
            "cache": "inputs larger than L2: about %.1f GB of HBM-resident arrays touched per step per GPU vs 126 MB L2"
                     % (ncol_per_gpu * 1870.0 / 1e9)}


def run_reference(args, real_stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, ncol, el = cpu_baseline(args.steps, args.warmup, target_seconds=30.0)
    line = {"impl": "reference", "metric": "soil_column_timesteps_per_sec", "value": cb["value"], "unit": "column-timesteps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.ncol, args.gpus), "sample_ncol": ncol, "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "column-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(real_stdout, line)


# --------------------------------------------------------------------------------------------------------------------
# side measurements at N = 1
# --------------------------------------------------------------------------------------------------------------------
def other_workloads(device, stream):
    """BASELINE.json configs[1] (soil thermal, 1 Mi columns x 15), ELM's default curve, the snow + standing-water thermal system and
    the two ELM driver calls on this GPU, device-resident, CUDA-event time of the StepDT kernels: reported beside the headline."""
    import mpp_b200
    out = {}
    # thermal: one ELM coupling step loads the per-step data, then StepDT chained on the device
    ncol = 1 << 20
    d = PB.elm_thermal_inputs(ncol, NLEV)
    p, ids = PB.build_elm_thermal(mpp_b200.Thermal, d, device=device)
    p.set_stream(stream)
    PB.elm_thermal_step(p, ids, d, d["T0"], DT, 1)
    ms = []
    for s in range(13):
        p.step_dt(DT, s + 2); ms.append(p.last_step_ms())
    m = float(np.mean(ms[3:]))
    out["thermal_1Mi_x15"] = {"column_timesteps_per_sec": ncol / (m * 1e-3), "ms_per_step": m, "kernel": "thermal_step2_kernel<8>",
                              "roofline": {"bound": "hbm", "algorithmic_bytes_per_column_step": 1224, "achieved": 1224 * ncol / (m * 1e-3) / 1e9, "unit": "GB/s"}}
    p.close()
    # the VSFM batch once more with ELM's default saturation curve, smooth_brooks_corey_bz3 (+32 B per cell of parameters)
    d = shard_inputs(0, ncol)
    d["satfunc"] = "smooth_brooks_corey_bz3"
    p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d, device=device)
    p.set_stream(stream)
    set_forcing_host(p, ids, d)
    ms, conv = [], True
    for s in range(9):
        p.pre_step_dt(); cv, _ = p.step_dt(DT, s + 1); p.post_step_dt(); ms.append(p.last_step_ms()); conv = conv and cv
    m = float(np.median(ms[3:]))
    st = p.stats()
    out["vsfm_sbc_bz3_1Mi_x15"] = {"column_timesteps_per_sec": ncol / (m * 1e-3), "ms_per_step_median": m, "kernel": "vsfm_step2_kernel<8,SBC,noBC>",
                                   "converged_all": bool(conv), "newton_its_mean": float(st["newton_its"].mean()), "residual_evals_mean": float(st["nfuncs"].mean()),
                                   "roofline": {"bound": "hbm", "algorithmic_bytes_per_column_step": 1224 + 32 * NLEV,
                                                "achieved": (1224 + 32 * NLEV) * ncol / (m * 1e-3) / 1e9, "unit": "GB/s"}}
    p.close()
    # the survey's original batch (water tables from 1 m): the dt-cut / failing columns of tests/golden/hard_columns.json are in it
    d = shard_inputs(0, ncol, zwt_min=1.0)
    p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d, device=device)
    p.set_stream(stream)
    set_forcing_host(p, ids, d)
    ms, nfail, ncut = [], [], []
    for s in range(9):
        p.pre_step_dt(); p.step_dt(DT, s + 1); p.post_step_dt(); ms.append(p.last_step_ms())
        st = p.stats(); nfail.append(int((st["reasons"] < 0).sum())); ncut.append(int((st["dt_cuts"] > 0).sum()))
    out["vsfm_vg_1Mi_x15_water_table_from_1m"] = {
        "column_timesteps_per_sec_median": ncol / (float(np.median(ms[3:])) * 1e-3), "ms_per_step_all": [round(x, 2) for x in ms],
        "columns_failed_per_step": nfail, "columns_with_dt_cuts_per_step": ncut,
        "note": "SURVEY.md 8d's original draw (zwt 1-20 m): root-zone ET drags barely saturated van Genuchten cells through pc = 0, where the "
                "reference algorithm cuts dt or fails (mpp_b200/problems.py, tests/golden/hard_columns.json); not part of the headline"}
    p.close()
    # ELM's real thermal column: 5 snow layers (variable active count) + standing surface water + 15 soil layers
    base = 4096
    d0 = PB.elm_snow_thermal_inputs(base, NLEV, 5)
    d, o = PB.tile_snow_thermal(d0, PB.pack_elm_snow_thermal(d0), ncol // base)
    p = PB.build_elm_snow_thermal(mpp_b200.ThermalSnow, d, device=device)
    p.set_stream(stream)
    PB.elm_snow_thermal_step(p, o, DT, 1)
    ms = []
    for s in range(13):
        p.step_dt(DT, s + 2); ms.append(p.last_step_ms())
    m = float(np.mean(ms[3:]))
    out["thermal_snow_ssw_soil_1Mi_x21"] = {"column_timesteps_per_sec": ncol / (m * 1e-3), "ms_per_step": m, "kernel": "thermal_snow_step3_kernel<8>",
                                            "roofline": {"bound": "hbm", "algorithmic_bytes_per_column_step": 2356,
                                                         "achieved": 2356 * ncol / (m * 1e-3) / 1e9, "unit": "GB/s"}}
    # the whole MPPThermalTBasedALM_Solve on the same columns with ELM's (c, j) HOST arrays, page-locked in place once
    e0 = PB.elm_thermal_raw_arrays(d0)
    reps = ncol // base
    e = PB.page_aligned_state({k: (np.tile(v, (1, reps)) if v.ndim == 2 else np.tile(v, reps)) for k, v in e0.items()})
    for v in e.values():
        mpp_b200.host_register(v)
    wall, wall_all_rows = [], []
    for s in range(4):
        t0 = time.perf_counter(); p.elm_solve(DT, e, s + 20); wall_all_rows.append(time.perf_counter() - t0)
    p.elm_set_pipeline(0, static_soil_geometry=True)        # ELM's soil grid is fixed: z / dz / zi soil rows go up with the first solve only
    for s in range(5):
        t0 = time.perf_counter(); p.elm_solve(DT, e, s + 30); wall.append(time.perf_counter() - t0)
    for v in e.values():
        mpp_b200.host_unregister(v)
    nsno_, nl_ = 5, 5 + NLEV
    full = int(sum(v.nbytes for v in e.values()))
    out["thermal_snow_ssw_soil_1Mi_x21"]["elm_solve_host_arrays_page_locked"] = {
        "column_timesteps_per_sec": ncol / float(np.median(wall[2:])), "ms_per_solve": [round(w * 1e3, 2) for w in wall[1:]],
        "ms_per_solve_every_row_every_solve": [round(w * 1e3, 2) for w in wall_all_rows],
        "host_bytes_per_solve": full - (2 * (nl_ - nsno_) + (nl_ + 1 - nsno_ - 1)) * 8 * ncol, "host_bytes_per_solve_every_row": full,
        "api": "mppgpu_thermal_elm_solve, pipelined over 8 column chunks on 3 streams (mppgpu_elm_set_pipeline, static soil geometry): "
               "elm_thermal_pack_kernel + thermal_snow_step3_kernel + elm_thermal_unpack_kernel per chunk between its copies"}
    p.close()
    # MPPVSFMALM_Solve with ELM's raw column arrays: packing, StepDT, per-column retry loop, unpacking on the device.
    # With ELM's own default curve (smooth_brooks_corey_bz3, mpp_varctl.F90:17): the driver drains water from the saturated layers below
    # the water table, which pulls cells through pc = 0 every step -- with van Genuchten that is exactly where the reference algorithm
    # halves dt a dozen times (one column of 1 Mi then needs 10^4 - 10^6 residual evaluations per solve, oracle and GPU alike)
    d = shard_inputs(0, ncol)
    d["satfunc"] = "smooth_brooks_corey_bz3"
    p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d, device=device)
    p.set_stream(stream)
    st0 = PB.elm_vsfm_raw_state(p, d, patches=True)
    p.elm_set_geometry(st0["zi"], st0["dz"], st0["nlevsoi"], ids)
    sp = PB.page_aligned_state(st0)
    o = p.elm_solve(DT, sp, 1)
    op = PB.page_aligned_state(o)
    locked = [v for v in list(sp.values()) + list(op.values()) if isinstance(v, np.ndarray) and v.nbytes]
    for v in locked:
        mpp_b200.host_register(v)
    wall, msd, att, nf = [], [], [], []
    for s in range(4):
        t0 = time.perf_counter(); r = p.elm_solve(DT, sp, s + 2, out=op); wall.append(time.perf_counter() - t0)
        msd.append(round(p.last_step_ms(), 2)); att.append(r["nattempts"]); nf.append(r["nfailed"])
    for v in locked:
        mpp_b200.host_unregister(v)
    out["vsfm_elm_solve_1Mi_x15"] = {"column_timesteps_per_sec_host_arrays_page_locked": ncol / float(np.median(wall[1:])),
                                     "ms_per_solve_host_arrays_page_locked": [round(w * 1e3, 2) for w in wall],
                                     "ms_per_solve_device_span_first_upload_to_last_download": msd,
                                     "host_bytes_per_solve": int(ncol * (NLEV * 8 * (4 + 5) + 10 * 8 + 4 + 5 * 8 + 2 * 4)
                                                                 + sum(sp[k].nbytes for k in ("col_pfti", "col_npfts", "pft_active", "pft_wtcol", "rootr_pft", "qflx_tran_veg_pft") if k in sp)),
                                     "stepdt_calls": att, "columns_failed": nf,
                                     "api": "mppgpu_vsfm_elm_solve, pipelined over 8 column chunks on 3 streams (mppgpu_elm_set_pipeline)",
                                     "kernels": "per chunk: elm_pack_kernel<16> + vsfm_step2_kernel + elm_decide_kernel<16>; then the RETRY specialisation on the columns that need it",
                                     "satfunc": "smooth_brooks_corey_bz3",
                                     "note": "ELM's default curve; no step budget: reference behaviour; host arrays page-locked in place with mppgpu_host_register"}
    p.close()
    return out


# --------------------------------------------------------------------------------------------------------------------
def _claim_stdout():
    """The contract is ONE JSON line on stdout.  NCCL and friends print banners to the C-level stdout, so fd 1 is
    pointed at stderr for the whole run and the JSON line goes to the saved descriptor."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(real_stdout_fd, line):
    os.write(real_stdout_fd, (json.dumps(line) + "\n").encode())


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mpp_b200")
    ap.add_argument("--ncol", type=int, default=4 * 1024 * 1024, help="columns per GPU of the headline (weak scaling)")
    ap.add_argument("--zwt-min", type=float, default=PB.BENCH_ZWT_MIN, help="shallowest initial water table [m] (1.0 = SURVEY.md 8d's original draw)")
    ap.add_argument("--ordering", type=int, default=1, help="mppgpu_set_column_ordering (1 = by the previous step's cost, 0 = batch order)")
    ap.add_argument("--step-budget", type=int, default=0, help="mppgpu_set_step_budget (0 = off = reference behaviour)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chunks", type=int, default=16, help="column chunks of the pipelined coupling step (0 = library default)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the thermal / driver-call side measurements (N = 1)")
    ap.add_argument("--no-multi", action="store_true", help="skip the strong-scaling and TH configurations")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3                                      # timing rules: W >= 3
    global ZWT_MIN
    ZWT_MIN = args.zwt_min
    if args.impl == "reference":
        return run_reference(args, real_stdout)

    import torch
    import torch.distributed as dist
    import mpp_b200
    from mpp_b200 import parallel as PL

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    bind_to_gpu_numa_node(local_rank)            # before any pinned allocation: host staging buffers land next to the GPU
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus or world == 1
    # a dedicated (non-default) stream: the library (kernels and its own NCCL all-gather) and the CUDA events all run on it
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    facts = measured_kernel_facts()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_source = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(xs):
        t = torch.tensor(list(xs), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        return t.tolist()

    def timed(step_fn, steps, warmup, first_step=1):
        """W untimed + K timed steps; CUDA events on the library's stream; barrier + synchronize on both sides; max over ranks."""
        for s in range(warmup):
            step_fn(first_step + s)
        barrier()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        marks[0].record(stream)
        for s in range(steps):
            step_fn(first_step + warmup + s)
            marks[s + 1].record(stream)              # per-step times (no synchronisation): shows which steps carried a slow column
        barrier()
        per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
        return max_over_ranks(marks[0].elapsed_time(marks[steps])), per_step

    def solver_summary(p, glob):
        st = p.stats()
        nf = st["nfuncs"]
        tot = sum_over_ranks([float((st["reasons"] < 0).sum()), float((st["dt_cuts"] > 0).sum()), float(nf.sum()), float(st["newton_its"].sum()), float(nf.size)])
        return {"converged_all": not glob["any_diverged"], "worst_reason": glob["worst_reason"],
                "newton_its_mean": tot[3] / tot[4], "newton_its_max": glob["max_newton_its"], "residual_evals_mean": tot[2] / tot[4],
                "residual_evals_max_rank0": int(nf.max()), "max_abs_mass_error_kg": glob["max_abs_mass_error"], "max_dt_cuts": glob["max_dt_cuts"],
                "columns_failed_last_step": int(tot[0]), "columns_with_dt_cuts_last_step": int(tot[1]),
                "warp_waste_batch_order_rank0": warp_waste(nf), "global_reductions_last_step": glob}

    # ------------------------------------------------------------------ VSFM: device-resident throughput of one shard layout
    def run_vsfm(c0, c1, steps, warmup, e2e):
        d = shard_inputs(c0, c1)
        ncol = c1 - c0

        def fresh():
            p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d, device=local_rank)
            p.set_stream(stream.cuda_stream)
            p.set_column_ordering(args.ordering)
            if args.step_budget:
                p.set_step_budget(args.step_budget)
            set_forcing_host(p, ids, d)
            PL.init_comm(p)
            return p, ids

        p, ids = fresh()

        def step(nstep):
            p.pre_step_dt()
            p.step_dt_async(DT, nstep)
            p.post_step_dt()
            p.global_reduce_async()                      # global mass-balance / convergence reductions

        l0 = [0]

        def step_counted(nstep):
            if nstep == warmup + 1:
                l0[0] = p.launch_count()
            step(nstep)
        sampler = ClockSampler(local_rank)
        ms, per_step = timed(step_counted, steps, warmup)
        clocks = sampler.stop()
        launches = p.launch_count() - l0[0]
        glob = p.global_mass_balance()
        res = {"ms_total": ms, "per_step_ms": [round(x, 3) for x in per_step], "clocks": clocks, "launches": int(launches),
               "last_kernel_ms": p.last_step_ms(), "solver": solver_summary(p, glob), "ncol": ncol}
        p.close()
        del p
        res["e2e"] = None
        if e2e:
            # end to end through the C ABI with pinned host buffers: one ELM coupling step (MPPVSFMALM_Driver.F90:379-463, 603, 642,
            # 674-705): 7 inputs host->device, PreStepDT + StepDT, 4 outputs device->host, pipelined over column chunks
            p, ids = fresh()
            pin = {k: torch.from_numpy(d[k]).pin_memory().numpy() for k in SS_NAMES + ("frac_liq",)}
            outs = {k: torch.empty(ncol * NLEV, dtype=torch.float64).pin_memory().numpy() for k in ("sat", "mass", "smp", "pressure")}
            h2d, d2h = sum(v.nbytes for v in pin.values()), sum(v.nbytes for v in outs.values())
            ins = [(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, ids[name], pin[name]) for name in SS_NAMES]
            ins.append((K.AUXVAR_INTERNAL, K.VAR_FRAC_LIQ_SAT, 1, pin["frac_liq"]))
            olist = [(K.AUXVAR_INTERNAL, var, 1, outs[key]) for key, var in
                     (("sat", K.VAR_LIQ_SAT), ("mass", K.VAR_MASS), ("smp", K.VAR_SOIL_MATRIX_POT), ("pressure", K.VAR_PRESSURE))]
            conv = [True]

            def e2e_step(nstep):
                cv, rs = p.coupled_step(DT, nstep, ins, olist, args.chunks)
                p.post_step_dt()
                p.global_reduce_async()
                conv[0] = conv[0] and cv
            for s in range(warmup):
                e2e_step(s + 1)
            barrier()
            conv[0] = True
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            t0 = time.perf_counter()
            for s in range(steps):
                e2e_step(warmup + s + 1)
            e1.record(stream)
            barrier()
            wall = time.perf_counter() - t0
            ems = max_over_ranks(max(e0.elapsed_time(e1), wall * 1e3))     # host staging happens off-stream: take the larger clock
            glob = p.global_mass_balance()
            res["e2e"] = {"ms_total": ems, "h2d": h2d, "d2h": d2h, "converged": bool(conv[0]) and not glob["any_diverged"],
                          "max_abs_mass_error_kg": glob["max_abs_mass_error"]}
            p.close()
        return res

    # ------------------------------------------------------------------ TH (configs[4]): 2 Mi columns in total over the ranks
    def run_th(ncol_total, steps, warmup):
        c0, c1 = PL.shard_range(ncol_total, rank, world)
        d = shard_inputs_th(c0, c1)
        p, ids = PB.build_elm_th(mpp_b200.TH, d, device=local_rank)
        p.set_stream(stream.cuda_stream)
        PL.init_comm(p)
        p.set_data(K.AUXVAR_BC, K.VAR_BC_SS_CONDITION, ids["T_top"], d["T_top"], ieqn=2)
        p.set_data(K.AUXVAR_BC, K.VAR_PRESSURE, ids["T_top"], d["P_top_bc"], ieqn=2)
        p.set_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, ids["infil"], d["infil"], ieqn=1)
        p.set_data(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, ids["heat"], d["heat"], ieqn=2)

        def step(nstep):
            p.pre_step_dt()
            p.step_dt_async(DT, nstep)
            p.post_step_dt()
            p.global_reduce_async()
        ms, per_step = timed(step, steps, warmup)
        glob = p.global_mass_balance()
        st = p.stats()
        tot = sum_over_ranks([float(st["nfuncs"].sum()), float(st["newton_its"].sum()), float(st["nfuncs"].size), float((st["dt_cuts"] > 0).sum())])
        m = ms / steps
        ach = ALG_BYTES_TH * (c1 - c0) / (m * 1e-3) / 1e9
        kname = "th_step2_kernel<16,SBC,TGDPB01,const>"
        f = facts.get(kname, {})
        out = {"column_timesteps_per_sec": ncol_total * steps / (ms * 1e-3), "ms_per_step": m, "ms_per_step_all_rank0": [round(x, 2) for x in per_step],
               "ncol_total": ncol_total, "ncol_per_gpu": c1 - c0, "scaling": "strong",
               "workload": "coupled thermal-hydrology (BASELINE.json configs[4]): %d columns x %d layers in total over %d GPU(s); VSFM soils + csol, tkdry; "
                           "ELM's default curve smooth_brooks_corey_bz3 (mpp_varctl.F90:17; with van Genuchten the reference algorithm stalls on the cell "
                           "that holds the water table, mpp_b200/problems.py); Tanaka density + constant heat capacity; Dirichlet surface temperature, "
                           "mass-rate infiltration, heat-rate source; 2x2 block-tridiagonal Newton system per column" % (ncol_total, NLEV, world),
               "converged_all": not glob["any_diverged"], "worst_reason": glob["worst_reason"], "max_dt_cuts": glob["max_dt_cuts"],
               "columns_with_dt_cuts_last_step": int(tot[3]), "newton_its_mean": tot[1] / tot[2], "residual_evals_mean": tot[0] / tot[2],
               "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "kernel": kname,
                            "algorithmic_bytes_per_column_step": ALG_BYTES_TH,
                            "traffic": (f["dram_bytes_per_unit"] * (c1 - c0)) if f.get("dram_bytes_per_unit") else None,
                            "fp64_issue_frac": f.get("fp64_pipe_frac")},
               "limiter": "fp64 issue / dependent latency of the 2x2 block Newton loop; the step ends with its slowest column (dt cuts of the reference algorithm)"}
        p.close()
        return out

    # ================================================================== headline: weak scaling, distinct columns per rank
    c0, c1 = rank * args.ncol, (rank + 1) * args.ncol
    ncol_total = args.ncol * world
    R = run_vsfm(c0, c1, args.steps, args.warmup, e2e=not args.no_e2e)
    per_launch_ms = R["ms_total"] / args.steps
    value = ncol_total * args.steps / (R["ms_total"] * 1e-3)
    achieved = ALG_BYTES_PER_COLSTEP * args.ncol / (per_launch_ms * 1e-3) / 1e9
    kname = "vsfm_step2_kernel<8,VG,noBC>"
    f = facts.get(kname, {})

    multi = {}
    if not args.no_multi:
        # configs[3] literally: 4 Mi columns in total, split over the ranks
        if world == 1:
            multi["strong_4Mi_total"] = {"column_timesteps_per_sec": value, "ms_per_step": per_launch_ms, "ncol_total": args.ncol, "ncol_per_gpu": args.ncol,
                                         "scaling": "strong", "same_run_as": "headline (N = 1)",
                                         "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "kernel": kname}}
        else:
            s0, s1 = PL.shard_range(args.ncol, rank, world)
            S = run_vsfm(s0, s1, args.steps, args.warmup, e2e=False)
            sm = S["ms_total"] / args.steps
            ach = ALG_BYTES_PER_COLSTEP * (s1 - s0) / (sm * 1e-3) / 1e9
            multi["strong_4Mi_total"] = {"column_timesteps_per_sec": args.ncol * args.steps / (S["ms_total"] * 1e-3), "ms_per_step": sm,
                                         "ncol_total": args.ncol, "ncol_per_gpu": s1 - s0, "scaling": "strong",
                                         "ms_per_step_all_rank0": S["per_step_ms"], "converged_all": S["solver"]["converged_all"],
                                         "max_abs_mass_error_kg": S["solver"]["max_abs_mass_error_kg"],
                                         "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "kernel": kname,
                                                      "algorithmic_bytes_per_column_step": ALG_BYTES_PER_COLSTEP},
                                         "limiter": "per-GPU share of %d columns is %d warps per SM-slot wave: launch + reduction + all-gather latency "
                                                    "(~30 us) and the tail of the last wave weigh more as the share shrinks" % (s1 - s0, (s1 - s0) // 4 // (148 * 16))}
        multi["th_2Mi_total"] = run_th(2 * 1024 * 1024, args.steps, args.warmup)

    if rank == 0:
        e2e = None
        if R["e2e"]:
            E = R["e2e"]
            e2e = {"value": ncol_total * args.steps / (E["ms_total"] * 1e-3), "unit": "column-timesteps/s",
                   "h2d_bytes_per_step": int(E["h2d"] * world), "d2h_bytes_per_step": int(E["d2h"] * world),
                   "ms_per_step": E["ms_total"] / args.steps, "converged": E["converged"], "max_abs_mass_error_kg": E["max_abs_mass_error_kg"],
                   "api": "mppgpu_vsfm_coupled_step (= SetDataFromCLM x7 from pinned host + PreStepDT + StepDT + GetDataForCLM x4 to pinned host, "
                          "software-pipelined over %d column chunks on 3 streams) + mppgpu_post_step_dt + mppgpu_global_reduce_async" % (args.chunks or 16),
                   "limiter": "host DMA: %.2f GB cross PCIe per step per GPU; all GPUs of this box hang off one host NUMA node" % ((E["h2d"] + E["d2h"]) / 1e9)}
        sol = R["solver"]
        line = {
            "metric": "soil_column_timesteps_per_sec", "value": value, "unit": "column-timesteps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_launch_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.ncol, world),
            "clocks": {"sm_mhz": R["clocks"]["sm_mhz"], "sm_max_mhz": R["clocks"]["sm_max_mhz"], "reasons": R["clocks"]["reasons"], "samples": R["clocks"]["samples"]},
            "e2e": e2e, "gpu_launches": R["launches"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (f["dram_bytes_per_unit"] * args.ncol) if f.get("dram_bytes_per_unit") else None,
                         "traffic_source": f.get("source"),
                         "kernel": kname, "peak_source": peak_source,
                         "algorithmic_bytes_per_column_step": ALG_BYTES_PER_COLSTEP,
                         "fp64_issue_frac": f.get("fp64_pipe_frac"),
                         "fp64_issue_frac_source": ("sm__pipe_fp64_cycles_active of the same kernel, " + f["source"]) if f.get("source") else None,
                         "note": "fp64-issue / dependent-latency bound, not HBM bound (2 log + 2 exp + 3 reciprocals per cell per residual evaluation, "
                                 "%.1f evaluations and %.1f Newton iterations per column-step); see DESIGN.md" % (sol["residual_evals_mean"], sol["newton_its_mean"])},
            "solver": dict(sol, last_step_kernel_ms=R["last_kernel_ms"], ms_per_step_all=R["per_step_ms"], column_ordering=args.ordering,
                           step_budget=args.step_budget),
            "multi_gpu": multi,
        }
        if world == 1 and not args.no_other:
            line["other_workloads"] = other_workloads(local_rank, stream.cuda_stream)
            for v in line["other_workloads"].values():
                if "roofline" in v:
                    v["roofline"]["peak"] = peak; v["roofline"]["frac"] = v["roofline"]["achieved"] / peak
        if not args.no_cpu and world == 1:
            cb, _, _ = cpu_baseline(args.steps, args.warmup)
            line["cpu_baseline"] = cb
        _emit(real_stdout, line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
