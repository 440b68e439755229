#!/usr/bin/env python
"""bench.py -- headline benchmark of the VSFM hot path (BASELINE.json: soil column-timesteps/s, fp64, Newton-converged).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm on the host cores

Workload (BASELINE.json configs[3]): 4 Mi synthetic ELM-like soil columns x 15 layers PER GPU, van Genuchten curves,
Tanaka density, six COND_MASS_RATE source/sinks, dt = 1800 s, SNES tolerances = reference defaults.  Columns are
independent; weak scaling (default): every rank solves the SAME 4 Mi-column batch the 1-GPU run solves, so the work per
GPU is identical by construction (`--scaling weak-distinct` gives every rank its own 4 Mi columns of one 4N Mi global
batch -- then the step time is the slowest column among 4N Mi, see DESIGN.md "hard columns"; `--scaling strong` splits
one 4 Mi batch over the ranks).  One "step" = one ELM coupling step =
PreStepDT + StepDT + PostStepDT over the whole batch (MPPVSFMALM_Driver.F90:603-935) + one NCCL all-gather of the
9 mass-balance / convergence doubles of every rank.

Prints ONE JSON line (rank 0).  `value` is timed with all inputs resident in HBM; `e2e` repeats the same steps
through the C ABI with HOST buffers (SetDataFromCLM x7 in, GetDataForCLM x4 out, as MPPVSFMALM_Solve does).
The reference itself (Fortran + PETSc + MPI) cannot be built in this image, so `cpu_baseline` / `--impl reference`
time the oracle -- the C restatement of the reference algorithm (oracle/, kind "port") -- on the box's host cores.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import problems as PB  # noqa: E402
from mpp_b200 import constants as K  # noqa: E402

NLEV = 15
DT = 1800.0
CHUNK = 65536                      # columns per seeded chunk (shards are unions of chunks)
ALG_BYTES_PER_COLSTEP = 1224       # SURVEY.md section 8(d), VSFM-VG base variant
# DRAM bytes per column-step of vsfm_step2_kernel measured by ncu (dram__bytes_read.sum + dram__bytes_write.sum over
# 1 Mi columns, profiles/r1_vsfm_v12.md: 1.3087 GB + 0.6562 GB); the kernel also reads the six source arrays unsummed
# and writes mass, smp and the committed solution, which the 1224 B base variant does not count
TRAFFIC_BYTES_PER_COLSTEP = (1.308749e9 + 0.656228e9) / 1048576
SS_NAMES = ("infil", "et", "dew", "drain", "snow", "sublim")


def shard_inputs(c0, c1, chunk=CHUNK):
    """Columns [c0, c1) of the global seeded batch: chunk k uses seed SEED + k so any rank builds only its shard."""
    CHUNK = chunk
    parts = []
    k0, k1 = c0 // CHUNK, (c1 - 1) // CHUNK
    for k in range(k0, k1 + 1):
        d = PB.elm_vsfm_inputs(CHUNK, NLEV, seed=PB.SEED + k)
        lo, hi = max(c0, k * CHUNK) - k * CHUNK, min(c1, (k + 1) * CHUNK) - k * CHUNK
        parts.append((d, lo, hi))
    out = {"ncol": c1 - c0, "nlev": NLEV, "satfunc": "van_genuchten"}
    for key in ("dz", "watsat", "hksat", "bsw", "sucsat", "residual_sat"):
        out[key] = np.concatenate([d[key][lo:hi] for d, lo, hi in parts], axis=0)
    for key in ("area", "infil", "dew", "snow", "sublim"):
        out[key] = np.concatenate([d[key][lo:hi] for d, lo, hi in parts])
    for key in ("press_ic", "et", "drain", "frac_liq"):
        out[key] = np.concatenate([d[key].reshape(CHUNK, NLEV)[lo:hi].reshape(-1) for d, lo, hi in parts])
    return out


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


def cpu_baseline(steps, warmup, target_seconds=20.0):
    """The oracle (reference algorithm, per-column SNES) on a bounded column sample with every host thread."""
    from oracle import oracle as O
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
    for s in range(steps):
        conv, reason, _ = PB.elm_vsfm_step(o, ids, d, DT, warmup + s + 1)
    el = time.perf_counter() - t0
    return {"value": ncol * steps / el, "unit": "column-timesteps/s", "cores": cores, "kind": "port",
            "sample": "%d of the benchmark's columns (global columns 0..%d), same %d warm-up + %d timed steps, "
                      "oracle/ C restatement of the reference algorithm (per-column SNES newtonls+bt, Thomas), OpenMP over columns; "
                      "the Fortran+PETSc reference cannot be built in this image" % (ncol, ncol - 1, warmup, steps),
            "seconds": el, "converged": bool(conv)}, ncol, el


def other_workloads(device, stream):
    """BASELINE.json configs[1] (soil thermal, 1 Mi columns x 15) and configs[4] (TH, 2 Mi columns x 15 over 8 GPUs = 256 Ki per
    GPU) on this GPU, device-resident, CUDA-event time of the StepDT kernels: reported beside the headline, not part of it."""
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
    out["thermal_1Mi_x15"] = {"column_timesteps_per_sec": ncol / (m * 1e-3), "ms_per_step": m, "kernel": "thermal_step_kernel<16>",
                              "roofline": {"bound": "hbm", "algorithmic_bytes_per_column_step": 1224, "achieved": 1224 * ncol / (m * 1e-3) / 1e9, "unit": "GB/s"}}
    p.close()
    # the VSFM batch once more with ELM's default saturation curve, smooth_brooks_corey_bz3 (+32 B per cell of parameters)
    d = shard_inputs(0, ncol)
    d["satfunc"] = "smooth_brooks_corey_bz3"
    p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d, device=device)
    p.set_stream(stream)
    set_forcing_host(p, ids, d)
    ms = []
    for s in range(9):
        p.pre_step_dt(); p.step_dt(DT, s + 1); p.post_step_dt(); ms.append(p.last_step_ms())
    m = float(np.median(ms[3:]))
    st = p.stats()
    out["vsfm_sbc_bz3_1Mi_x15"] = {"column_timesteps_per_sec": ncol / (m * 1e-3), "ms_per_step_median": m, "kernel": "vsfm_step2_kernel<8,SBC,noBC>",
                                   "newton_its_mean": float(st["newton_its"].mean()), "residual_evals_mean": float(st["nfuncs"].mean()),
                                   "roofline": {"bound": "hbm", "algorithmic_bytes_per_column_step": 1224 + 32 * NLEV,
                                                "achieved": (1224 + 32 * NLEV) * ncol / (m * 1e-3) / 1e9, "unit": "GB/s"}}
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
    wall, dms = [], []
    for s in range(4):
        t0 = time.perf_counter(); p.elm_solve(DT, e, s + 20); wall.append(time.perf_counter() - t0); dms.append(p.last_step_ms())
    for v in e.values():
        mpp_b200.host_unregister(v)
    out["thermal_snow_ssw_soil_1Mi_x21"]["elm_solve_host_arrays_page_locked"] = {
        "column_timesteps_per_sec": ncol / float(np.median(wall[1:])), "ms_per_solve": [round(w * 1e3, 2) for w in wall],
        "host_bytes_per_solve": int(sum(v.nbytes for v in e.values())),
        "api": "mppgpu_thermal_elm_solve: elm_thermal_pack_kernel + thermal_snow_step3_kernel + elm_thermal_unpack_kernel between the copies"}
    p.close()
    # MPPVSFMALM_Solve with ELM's raw column arrays: packing, StepDT, per-column retry loop, unpacking on the device.
    # Synthetic forcing is not state-aware (ELM would cut infiltration into a saturated column), so a column may fail every retry;
    # the opt-in step budget keeps such a column from dominating the timing (it otherwise burns ~1e6 residual evaluations per call).
    ncol = 1 << 20
    d = shard_inputs(0, ncol)
    p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d, device=device)
    p.set_stream(stream)
    st = PB.elm_vsfm_raw_state(p, d, patches=True)
    p.elm_set_geometry(st["zi"], st["dz"], st["nlevsoi"], ids)
    p.set_step_budget(2000)
    st0 = PB.copy_state(st)
    ms, wall, att, nf = [], [], [], []
    for s in range(3):
        t0 = time.perf_counter(); o = p.elm_solve(DT, st, s + 1); wall.append(time.perf_counter() - t0)
        ms.append(p.last_step_ms()); att.append(o["nattempts"]); nf.append(o["nfailed"])
    m = float(np.median(ms))                           # (the first solve also pays the lazy load of the driver kernels)
    # the same solves with the caller's arrays page-locked in place once (mppgpu_host_register), as a host model would at start-up
    # (a fresh problem from the same initial state: the same three solves)
    p.close()
    p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d, device=device)
    p.set_stream(stream)
    p.elm_set_geometry(st0["zi"], st0["dz"], st0["nlevsoi"], ids)
    p.set_step_budget(2000)
    sp, op = PB.page_aligned_state(st0), PB.page_aligned_state(o)
    locked = [v for v in list(sp.values()) + list(op.values()) if isinstance(v, np.ndarray) and v.nbytes]
    for v in locked:
        mpp_b200.host_register(v)
    wall_locked, ms_locked = [], []
    for s in range(3):
        t0 = time.perf_counter(); p.elm_solve(DT, sp, s + 1, out=op); wall_locked.append(time.perf_counter() - t0)
        ms_locked.append(round(p.last_step_ms(), 2))
    for v in locked:
        mpp_b200.host_unregister(v)
    out["vsfm_elm_solve_1Mi_x15"] = {"column_timesteps_per_sec_device": ncol / (m * 1e-3), "ms_per_solve_device": m,
                                     "column_timesteps_per_sec_host_arrays": ncol / float(np.median(wall)),
                                     "column_timesteps_per_sec_host_arrays_page_locked": ncol / float(np.median(wall_locked)),
                                     "ms_per_solve_host_arrays": [round(w * 1e3, 2) for w in wall], "ms_per_solve_host_arrays_page_locked": [round(w * 1e3, 2) for w in wall_locked],
                                     "ms_per_solve_device_all": [round(x, 2) for x in ms], "ms_per_solve_device_page_locked_run": ms_locked,
                                     "stepdt_calls": att, "columns_failed": nf,
                                     "kernels": "elm_pack_kernel<16> + vsfm_step2_kernel + elm_decide_kernel<16> (+ RETRY specialisation on the columns that need it)",
                                     "note": "step budget 2000 residual evaluations per column per StepDT; host arrays: pageable numpy buffers, then the same arrays page-locked with mppgpu_host_register"}
    p.close()
    # TH: Tanaka density + constant heat capacity (the throughput variant of SURVEY.md section 8d)
    ncol = 1 << 18
    d = PB.elm_th_inputs(ncol, NLEV)
    p, ids = PB.build_elm_th(mpp_b200.TH, d, device=device)
    p.set_stream(stream)
    ms, conv = [], True
    for s in range(8):
        cv, reason, _ = PB.elm_th_step(p, ids, d, DT, s + 1); ms.append(p.last_step_ms()); conv = conv and cv
    st = p.stats()
    m = float(np.median(ms[2:]))
    out["th_256Ki_x15"] = {"column_timesteps_per_sec": ncol / (m * 1e-3), "ms_per_step_median": m, "ms_per_step_all": [round(x, 2) for x in ms],
                           "column_timesteps_per_sec_best_step": ncol / (min(ms[1:]) * 1e-3),
                           "kernel": "th_step2_kernel<16,VG,TGDPB01,const>", "converged_all": bool(conv),
                           "newton_its_mean": float(st["newton_its"].mean()), "residual_evals_mean": float(st["nfuncs"].mean()),
                           "roofline": {"bound": "hbm", "algorithmic_bytes_per_column_step": 1824, "achieved": 1824 * ncol / (m * 1e-3) / 1e9, "unit": "GB/s"},
                           "note": "step time is set by the slowest column of the batch (dt cuts of the reference algorithm), hence the median"}
    p.close()
    return out


def run_reference(args, real_stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, ncol, el = cpu_baseline(args.steps, args.warmup, target_seconds=30.0)
    line = {"impl": "reference", "metric": "soil_column_timesteps_per_sec", "value": cb["value"], "unit": "column-timesteps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.scaling == "strong" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.ncol if args.scaling == "strong" else args.ncol * args.gpus, args.gpus), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "column-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(real_stdout, line)


def workload_config(ncol, ngpus):
    return {"workload": "VSFM Richards (BASELINE.json configs[3]): %d synthetic ELM-like soil columns x %d layers per GPU, van Genuchten-Mualem, "
                        "Tanaka density, 6 COND_MASS_RATE source/sinks, dt=%.0f s, SNES rtol 1e-8 / stol 1e-10 / max_it 50 "
                        "(reference defaults), per-column Newton + bt line search + tridiagonal solve" % (ncol // ngpus, NLEV, DT),
            "ncol_total": ncol, "ncol_per_gpu": ncol // ngpus, "nlev": NLEV, "dt_s": DT, "satfunc": "van_genuchten",
            "parallelism": "columns sharded contiguously over %d GPU(s), no data-path collective; one NCCL all-gather of 9 "
                           "mass-balance/convergence doubles per rank per step" % ngpus,
            "cache": "inputs larger than L2: %.1f GB of HBM-resident arrays touched per step per GPU vs 126 MB L2"
                     % (ncol / ngpus * TRAFFIC_BYTES_PER_COLSTEP / 1e9)}


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
    ap.add_argument("--ncol", type=int, default=4 * 1024 * 1024, help="columns per GPU (weak scaling) / in total (strong scaling)")
    ap.add_argument("--scaling", default="weak", choices=("weak", "weak-distinct", "strong"))
    ap.add_argument("--step-budget", type=int, default=0, help="mppgpu_set_step_budget (0 = off = reference behaviour)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chunks", type=int, default=16, help="column chunks of the pipelined coupling step (0 = library default)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the thermal / TH side measurements")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3                                      # timing rules: W >= 3
    if args.impl == "reference":
        return run_reference(args, real_stdout)

    import torch
    import torch.distributed as dist
    import mpp_b200

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

    from mpp_b200 import parallel as PL
    ncol_total = args.ncol if args.scaling == "strong" else args.ncol * world
    if args.scaling == "weak":
        c0, c1 = 0, args.ncol                       # every rank solves the same 4 Mi-column batch
    else:
        c0, c1 = PL.shard_range(ncol_total, rank, world)
    d = shard_inputs(c0, c1)
    ncol = c1 - c0
    # a dedicated (non-default) stream: the library, the CUDA events and NCCL all run on it
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)

    def fresh():
        p, ids = PB.build_elm_vsfm(mpp_b200.VSFM, d, device=local_rank)
        p.set_stream(stream.cuda_stream)
        if args.step_budget:
            p.set_step_budget(args.step_budget)
        set_forcing_host(p, ids, d)
        return p, ids

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident throughput
    p, ids = fresh()
    dev = torch.device("cuda", local_rank)
    gred = PL.GlobalReductions(PL.device_view(p.reduction_buffer_ptr(), PL.NRED, dev))

    def step(nstep):
        p.pre_step_dt()
        p.step_dt_async(DT, nstep)
        p.post_step_dt()
        gred.step()                                  # global mass-balance / convergence reductions

    for s in range(args.warmup):
        step(s + 1)
    barrier()
    l0 = p.launch_count()
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev0.record(stream)
    marks[0].record(stream)
    for s in range(args.steps):
        step(args.warmup + s + 1)
        marks[s + 1].record(stream)                  # per-step times (no synchronisation): shows which steps carried a slow column
    ev1.record(stream)
    barrier()
    per_step_ms = [round(marks[i].elapsed_time(marks[i + 1]), 3) for i in range(args.steps)]
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = p.launch_count() - l0
    last_kernel_ms = p.last_step_ms()
    conv, reason = p.step_result()
    sums, maxs = p.mass_balance(DT)
    st = p.stats()
    glob = gred.as_dict()
    nfailed = torch.tensor([float((st["reasons"] < 0).sum()), float((st["dt_cuts"] > 0).sum())], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(nfailed)
    nfailed = nfailed.tolist()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = ncol_total * args.steps / (ms_max * 1e-3)
    its_mean, its_max, nf_mean = float(st["newton_its"].mean()), int(st["newton_its"].max()), float(st["nfuncs"].mean())
    p.close()
    del p

    # ------------------------------------------------------------------ end to end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        p, ids = fresh()
        pin = {k: torch.from_numpy(d[k]).pin_memory().numpy() for k in SS_NAMES + ("frac_liq",)}
        outs = {k: torch.empty(ncol * NLEV, dtype=torch.float64).pin_memory().numpy() for k in ("sat", "mass", "smp", "pressure")}
        h2d = sum(pin[k].nbytes for k in pin)
        d2h = sum(outs[k].nbytes for k in outs)

        ins = [(K.AUXVAR_SS, K.VAR_BC_SS_CONDITION, ids[name], pin[name]) for name in SS_NAMES]
        ins.append((K.AUXVAR_INTERNAL, K.VAR_FRAC_LIQ_SAT, 1, pin["frac_liq"]))
        olist = [(K.AUXVAR_INTERNAL, var, 1, outs[key]) for key, var in
                 (("sat", K.VAR_LIQ_SAT), ("mass", K.VAR_MASS), ("smp", K.VAR_SOIL_MATRIX_POT), ("pressure", K.VAR_PRESSURE))]

        def e2e_step(nstep):
            # one ELM coupling step through the C ABI with host buffers (MPPVSFMALM_Driver.F90:379-463, 603, 642, 674-705):
            # 7 inputs host->device, PreStepDT + StepDT, 4 outputs device->host, pipelined over column chunks
            cv, rs = p.coupled_step(DT, nstep, ins, olist, args.chunks)
            p.post_step_dt()
            gred2.step()
            return cv

        gred2 = PL.GlobalReductions(PL.device_view(p.reduction_buffer_ptr(), PL.NRED, dev))
        for s in range(args.warmup):
            e2e_step(s + 1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for s in range(args.steps):
            cv = e2e_step(args.warmup + s + 1)
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ems = max(e0.elapsed_time(e1), wall * 1e3)           # host staging happens off-stream: take the larger clock
        t = torch.tensor([ems], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": ncol_total * args.steps / (float(t.item()) * 1e-3), "unit": "column-timesteps/s",
               "h2d_bytes_per_step": int(h2d * world), "d2h_bytes_per_step": int(d2h * world),
               "ms_per_step": float(t.item()) / args.steps, "converged": bool(cv),
               "api": "mppgpu_vsfm_coupled_step (= SetDataFromCLM x7 from pinned host + PreStepDT + StepDT + GetDataForCLM x4 to pinned host, "
                      "software-pipelined over %d column chunks on 3 streams) + mppgpu_post_step_dt + NCCL all-gather" % (args.chunks or 16)}
        p.close()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        per_launch_ms = ms_max / args.steps
        achieved = ALG_BYTES_PER_COLSTEP * ncol / (per_launch_ms * 1e-3) / 1e9
        line = {
            "metric": "soil_column_timesteps_per_sec", "value": value, "unit": "column-timesteps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_launch_ms,
            "higher_is_better": True, "scaling": "strong" if args.scaling == "strong" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(ncol_total, world), shards={"weak": "every rank solves the same seeded 4 Mi-column batch", "weak-distinct": "rank r solves columns [r, r+1) x 4 Mi of one seeded global batch", "strong": "one 4 Mi batch split over the ranks"}[args.scaling], step_budget=args.step_budget),
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"], "samples": clocks["samples"]},
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": TRAFFIC_BYTES_PER_COLSTEP * ncol, "traffic_unit": "bytes per launch (ncu dram__bytes, per column x columns of this launch)",
                         "kernel": "vsfm_step2_kernel<8,VG,noBC>",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s",
                         "algorithmic_bytes_per_column_step": ALG_BYTES_PER_COLSTEP,
                         "fp64_issue_frac": 0.53, "fp64_issue_frac_source": "sm__pipe_fp64_cycles_active of the same kernel under ncu (profiles/r1_vsfm_v13.md); the second fraction SURVEY.md 8d asks for",
                         "note": "fp64-latency bound, not HBM bound (2 log + 2 exp + 3 reciprocals per cell per residual evaluation, "
                                 "%.1f evaluations and %.1f Newton iterations per column-step; fp64 pipe ~53 %% busy); see DESIGN.md" % (nf_mean, its_mean)},
            "solver": {"converged_all": not glob["any_diverged"], "worst_reason": glob["worst_reason"], "newton_its_mean": its_mean, "newton_its_max": its_max,
                       "residual_evals_mean": nf_mean, "max_abs_mass_error_kg": float(maxs[0]), "last_step_kernel_ms": last_kernel_ms,
                       "ms_per_step_all": per_step_ms,
                       "columns_failed_last_step": int(nfailed[0]), "columns_with_dt_cuts_last_step": int(nfailed[1]),
                       "global_reductions_last_step": glob,
                       "note": "the reference algorithm at its default tolerances cuts dt / fails on a handful of the 4 Mi synthetic columns; "
                               "the oracle reproduces every cut and failure (tests/golden/hard_columns.json)"},
        }
        if world == 1 and not args.no_other:
            peak_ = peak
            line["other_workloads"] = other_workloads(local_rank, stream.cuda_stream)
            for v in line["other_workloads"].values():
                if "roofline" in v:
                    v["roofline"]["peak"] = peak_; v["roofline"]["frac"] = v["roofline"]["achieved"] / peak_
        if not args.no_cpu and world == 1:
            cb, _, _ = cpu_baseline(args.steps, args.warmup)
            line["cpu_baseline"] = cb
        _emit(real_stdout, line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
