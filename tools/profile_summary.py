"""Turn an `ncu --set full --import-source on` report into the markdown summary kept under profiles/.

    python tools_profile_summary.py gpurun_out/prof.ncu-rep "title" [algorithmic_bytes_per_launch] > profiles/rN_name.md

Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv --print-source cuda,sass`
(B200_PROFILING.md recipe); needs no GPU."""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

RAW_KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of 64 warps)"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of ncu peak)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("inst_executed", "warp instructions executed"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC per SM (max 4)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active (%)"),
    ("sass__inst_executed_register_spilling", "spill instructions"),
    ("sass__inst_executed_local_loads", "local loads"), ("sass__inst_executed_local_stores", "local stores"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock during capture"),
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def to_float(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    rep, title = sys.argv[1], sys.argv[2]
    alg_bytes = float(sys.argv[3]) if len(sys.argv) > 3 else None
    # optional: --json "<kernel key of bench.py>" <units (columns) per launch>: record the measured DRAM bytes per unit and the fp64 pipe
    # share of this capture in profiles/roofline_measured.json, which bench.py reads (nothing is typed into bench.py)
    jkey = junits = None
    if "--json" in sys.argv:
        i = sys.argv.index("--json"); jkey, junits = sys.argv[i + 1], float(sys.argv[i + 2])
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    print("# %s\n" % title)
    print("Source: `%s` (`ncu --set full --clock-control none --import-source on`, one launch per row below)." % os.path.basename(rep))
    print("Numbers under a profiler are never bench values; they explain the CUDA-event timings in bench.py.\n")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    for r in rows:
        g = lambda k: r[hdr.index(k)] if k in hdr else None
        u = lambda k: units[hdr.index(k)] if k in hdr else ""
        print("## `%s`\n" % g("Kernel Name"))
        print("| metric | value |\n|---|---|")
        for k, name in RAW_KEYS:
            if g(k) is not None:
                print("| %s (`%s`) | %s %s |" % (name, k, g(k), u(k)))
        dur, rd, wr = to_float(g("gpu__time_duration.sum") or ""), to_float(g("dram__bytes_read.sum") or ""), to_float(g("dram__bytes_write.sum") or "")
        scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
        bscale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        if dur and rd is not None and wr is not None:
            t = dur * scale.get(u("gpu__time_duration.sum"), 1e-9)
            traffic = rd * bscale.get(u("dram__bytes_read.sum"), 1.0) + wr * bscale.get(u("dram__bytes_write.sum"), 1.0)
            print("| DRAM traffic (read + write) | %.4g GB per launch -> %.0f GB/s under ncu |" % (traffic / 1e9, traffic / t / 1e9))
            if alg_bytes:
                print("| algorithmic bytes per launch (DESIGN.md) | %.4g GB (traffic / algorithmic = %.2f) |" % (alg_bytes / 1e9, traffic / alg_bytes))
                if peaks.get("hbm_gbs"):
                    print("| algorithmic GB/s under ncu / measured HBM peak %.0f GB/s | %.0f GB/s = %.3f |" % (peaks["hbm_gbs"], alg_bytes / t / 1e9, alg_bytes / t / 1e9 / peaks["hbm_gbs"]))
            if jkey:
                jp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "roofline_measured.json")
                try:
                    facts = json.load(open(jp))
                except Exception:
                    facts = {}
                fp64 = to_float(g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") or "")
                facts[jkey] = {"dram_bytes_per_unit": traffic / junits, "fp64_pipe_frac": (fp64 / 100.0) if fp64 is not None else None,
                               "kernel_ms_under_ncu": t * 1e3, "units_per_launch": junits, "source": "profiles/" + os.path.basename(rep).replace(".ncu-rep", ".md")}
                json.dump(facts, open(jp, "w"), indent=1, sort_keys=True)
        print()
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    cur, h = None, None
    lines, ops, stalls = collections.OrderedDict(), collections.Counter(), collections.Counter()
    tot = tots = 0

    def I(x):
        try:
            return int(x)
        except ValueError:
            return 0
    for r in src:
        if len(r) == 2 and r[0] == "File Path":
            cur = os.path.basename(r[1]); continue
        if len(r) > 5 and r[0] == "Line No":
            h = r; iI, iS = h.index("Instructions Executed"), h.index("# Samples")
            iSass = [i for i, x in enumerate(h) if x == "Source"][1]
            continue
        if not h or len(r) <= iI:
            continue
        if r[0] != "":
            key = (cur, I(r[0]))
            a = lines.get(key, (0, 0, ""))
            lines[key] = (a[0] + I(r[iI]), a[1] + I(r[iS]), r[1])
    # opcode mix and stall reasons from the plain SASS page (the cuda,sass view repeats instructions under inlined frames)
    sass = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    h2 = None
    for r in sass:
        if len(r) > 5 and r[0] == "Address":
            h2 = r; jI, jS, jSrc = h2.index("Instructions Executed"), h2.index("# Samples"), h2.index("Source")
            st_cols = [i for i, x in enumerate(h2) if x.startswith("stall_") and "Not Issued" not in x]
            continue
        if not h2 or len(r) < len(h2):
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[jSrc])
        if m:
            ops[m.group(2)] += I(r[jI])
        tot += I(r[jI]); tots += I(r[jS])
        for i in st_cols:
            stalls[h2[i]] += I(r[i])
    if tot:
        print("## Instruction mix (SASS opcodes, share of %d executed warp instructions)\n" % tot)
        print("| opcode | share |\n|---|---|")
        for op, n in ops.most_common(16):
            print("| %s | %.1f %% |" % (op, 100.0 * n / tot))
        print("\n## Warp-stall reasons (share of %d samples)\n" % tots)
        print("| reason | share |\n|---|---|")
        for k, v in stalls.most_common(8):
            print("| %s | %.1f %% |" % (k, 100.0 * v / max(tots, 1)))
        print("\n## Hottest source lines (-lineinfo)\n")
        print("| file:line | instructions | samples | source |\n|---|---|---|---|")
        for (f, l), (n, s, text) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:25]:
            print("| %s:%d | %.2f %% | %.2f %% | `%s` |" % (f, l, 100.0 * n / tot, 100.0 * s / max(tots, 1), text.strip()[:110].replace("|", "\\|")))


if __name__ == "__main__":
    main()
