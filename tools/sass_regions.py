"""Development aid: per-region executed-instruction and stall-sample totals of one kernel from an ncu report's SASS source page.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > /tmp/s.csv ; python tools/sass_regions.py /tmp/s.csv [nbins]

Regions are split at backward-branch targets and at large jumps in the per-instruction execution count, which is what separates
prologue / Newton set-up / residual evaluation / line-search logic / epilogue in the step kernels."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; rows = rows[2:]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ilsb, iwait, inoi = hdr.index("stall_long_sb"), hdr.index("stall_wait"), hdr.index("stall_no_inst")
base = int(rows[0][ia], 16)
tot = sum(int(r[iex]) for r in rows); tots = sum(int(r[ismp]) for r in rows)
print("instructions executed %d, samples %d, static %d" % (tot, tots, len(rows)))
# regions: consecutive instructions whose execution counts are within 25 % of the region's first
regs = []; cur = None
for k, r in enumerate(rows):
    ex = int(r[iex])
    if cur is None or not (0.75 * cur["ex0"] <= ex <= 1.33 * cur["ex0"]) :
        cur = {"k0": k, "ex0": max(ex, 1), "n": 0, "ex": 0, "smp": 0, "ops": collections.Counter(), "lsb": 0, "wait": 0, "noi": 0}
        regs.append(cur)
    cur["n"] += 1; cur["ex"] += ex; cur["smp"] += int(r[ismp]); cur["lsb"] += int(r[ilsb]); cur["wait"] += int(r[iwait]); cur["noi"] += int(r[inoi])
    op = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
    cur["ops"][op.split(".")[0]] += ex
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
for g in regs:
    if g["ex"] / tot < minshare and g["smp"] / max(tots, 1) < minshare:
        continue
    top = ", ".join("%s %.0f%%" % (o, 100.0 * c / g["ex"]) for o, c in g["ops"].most_common(6))
    print("@%05x n=%4d per-inst exec %9d  inst %5.1f%%  samples %5.1f%% (long_sb %4.1f wait %4.1f no_inst %4.1f) | %s" % (
        int(rows[g["k0"]][ia], 16) - base, g["n"], g["ex0"], 100.0 * g["ex"] / tot, 100.0 * g["smp"] / max(tots, 1),
        100.0 * g["lsb"] / max(tots, 1), 100.0 * g["wait"] / max(tots, 1), 100.0 * g["noi"] / max(tots, 1), top))
