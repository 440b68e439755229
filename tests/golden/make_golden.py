"""Collect the reference's own golden vectors for the hot path into tests/golden/reference_baselines.json.

Run here (where /root/reference is mounted); the JSON it writes is committed so that the tests never need
the reference tree at run time.  Sources (paths relative to the reference root):
  regression_tests/vsfm/vsfm_celia1990.regression.baseline      (+ tolerance regression_tests/vsfm/vsfm.cfg:4-5)
  regression_tests/thermal/thermal_mms.regression.baseline      (+ regression_tests/thermal/thermal.cfg)
  regression_tests/th/mass_and_heat.regression.baseline, th_mms.regression.baseline (+ regression_tests/th/th.cfg)
  src/tests/test_eos_{constant,tgdp01,ifc67}_density.F90:16-25  (known answers, typed in below)
  src/driver/standalone/vsfm/vsfm_sy1991_problem.F90:17-106     (the two 200-value initial-pressure tables of the Srivastava & Yeh (1991)
                                                                 driver: input DATA of that problem, extracted into sy1991_ic.json)
"""
import re
import json
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def parse(path):
    out, cur = {}, None
    for line in open(path):
        line = line.strip()
        if not line:
            continue
        if line.startswith("["):
            cur = line[1:-1]
            out[cur] = {}
        elif "=" in line:
            k, v = [t.strip() for t in line.split("=")]
            out[cur][k] = v if k == "category" else float(v)
    return out


def main():
    g = {"_source": "MPP-LSM/MPP regression baselines and src/tests known answers; made by tests/golden/make_golden.py"}
    for name, rel in (("vsfm_celia1990", "regression_tests/vsfm/vsfm_celia1990.regression.baseline"),
                      ("thermal_mms", "regression_tests/thermal/thermal_mms.regression.baseline"),
                      ("mass_and_heat", "regression_tests/th/mass_and_heat.regression.baseline"),
                      ("th_mms", "regression_tests/th/th_mms.regression.baseline")):
        g[name] = parse(os.path.join(REF, rel))
    # src/tests/test_eos_*_density.F90: Density(p = 120000 Pa, T = 300 K)
    g["eos_density"] = {
        "p": 120000.0, "t_K": 300.0,
        "constant": {"den": 55.508250191225926, "dden_dp": 0.0, "dden_dT": 0.0},
        "tgdpb01": {"den": 55.317560635066179, "dden_dp": 2.4884914247886521e-8, "dden_dT": -1.5203176216371761e-2},
        "ifc67": {"den": 55.323696656461536, "dden_dp": 2.4854904480147891e-8, "dden_dT": -1.5298638598102345e-2},
        "tol": {"den": 1e-11, "dden_dp": 1e-16, "dden_dT": 1e-15},
    }
    # initial conditions of vsfm_sy1991_problem.F90 (data tables, not code)
    txt = open(os.path.join(REF, "src/driver/standalone/vsfm/vsfm_sy1991_problem.F90")).read()
    ic = {}
    for name in ("press_ic_wetting", "press_ic_drying"):
        m = re.search(r"%s\(200\)\s*=\s*\(/(.*?)/\)" % name, txt, flags=re.S)
        vals = [float(v.replace("d", "e")) for v in re.findall(r"[-+]?\d+\.\d*d[-+]?\d+", m.group(1))]
        assert len(vals) == 200, (name, len(vals))
        ic[name] = vals
    ic["_source"] = "src/driver/standalone/vsfm/vsfm_sy1991_problem.F90:17-106, extracted by tests/golden/make_golden.py"
    with open(os.path.join(HERE, "sy1991_ic.json"), "w") as f:
        json.dump(ic, f)
    with open(os.path.join(HERE, "reference_baselines.json"), "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(HERE, "reference_baselines.json"))


if __name__ == "__main__":
    main()
