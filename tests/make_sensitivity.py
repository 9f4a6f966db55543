"""Measure the unmodified reference's own sensitivity to last-bit changes of its
input (oracle/_ref/ref_dump, `perturb:INST,EPS` multiplies every state value by
1 + EPS r with a fixed pseudo-random r in [-1, 1]) on the cases whose parity
bounds are looser than rounding, and write tests/golden/sensitivity.json.

Two properties of the reference make a few quantities ill-conditioned
(DESIGN.md section 4): the implicit Jacobian carries sign(xi-dot), which is
rounding noise where the wind is exactly zero (JW equator and poles, the bubble
away from the anomaly), and the positivity filter of the tracers divides by the
non-negative mass of an element without a guard.  The tests bound the device's
deviation from the reference by a small multiple of the spread the reference
shows against itself; this script pins that spread.

Needs /root/reference (run `make -C oracle` first); the GPU box reads the JSON.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cases  # noqa: E402
import refdump  # noqa: E402

OUT = os.path.join(refdump.GOLDEN, "sensitivity.json")
EPS = ["1e-15", "-1e-15", "2e-15", "-3e-15"]


def checksum_spread(case, flags, steps, npatch=6):
    """Relative spread of Grid::Checksum after `steps` steps over the perturbed
    starts, per component, relative to the unperturbed checksum."""
    def run(eps):
        pre = "" if eps is None else "perturb:0,%s;" % eps
        d = refdump.run_ref_dump("/tmp/tb200_sens.bin", case,
                                 pre + "step:%d;checksum:cs" % steps,
                                 list(flags) + ["--nogeometry", "1"], npatch=npatch)
        return d["cs.checksum"]
    base = run(None)
    runs = [run(e) for e in EPS]
    spread = np.max([np.abs(r - base) for r in runs], axis=0)
    return dict(checksum=base.tolist(), abs_spread=spread.tolist(),
                rel_spread=(spread / np.maximum(np.abs(base), 1e-300)).tolist())


def tracer_spread(name):
    """Per tracer: max-norm spread of the final tracer fields relative to the
    largest value, and the relative change of the area-weighted tracer mass."""
    c = cases.CASES[name]
    geo = cases.load_case(name)

    def run(eps):
        pre = "addw:0,20000;dss:0;" + ("" if eps is None else "perturb:0,%s;" % eps)
        return refdump.run_ref_dump("/tmp/tb200_sens.bin", c["case"],
                                    pre + "step:2;dump:st,0", c["flags"])
    base = run(None)
    runs = [run(e) for e in EPS]
    ntr = refdump.scalar(base, "grid.ntracers")
    npatch = refdump.scalar(base, "grid.npatch")
    out = {"field_rel_spread": [], "mass_rel_spread": [], "mass": []}
    for t in range(ntr):
        num = den = 0.0
        mass = np.zeros(1 + len(runs))
        for n in range(npatch):
            area = geo["patch%d.elementareanode" % n][1:-1, 1:-1]
            ref = base["st.patch%d.inst0.tracers" % n][t][1:-1, 1:-1]
            den = max(den, np.abs(ref).max())
            mass[0] += (ref * area).sum()
            for q, r in enumerate(runs):
                v = r["st.patch%d.inst0.tracers" % n][t][1:-1, 1:-1]
                num = max(num, np.abs(v - ref).max())
                mass[1 + q] += (v * area).sum()
        out["field_rel_spread"].append(num / den)
        out["mass"].append(mass[0])
        out["mass_rel_spread"].append(np.abs(mass[1:] - mass[0]).max() / abs(mass[0]))
    return out


def state_spread(name, steps=2):
    """Per state component (u, v, rho theta, w, rho): max-norm spread of the
    fields after `steps` steps relative to the largest value, and the spread
    of the implicit stage (`vimp` from the recorded stage sequence of
    cases._STAGES_VO) relative to the largest change the stage makes."""
    c = cases.CASES[name]
    where = [("node", 0), ("node", 1), ("node", 2), ("redge", 3), ("node", 4)]

    def run(eps):
        pre = "addw:0,20000;dss:0;"
        per = "" if eps is None else "perturb:%d,%s;"
        return refdump.run_ref_dump(
            "/tmp/tb200_sens.bin", c["case"],
            pre + "copy:0,1;hexp:0,1,50;vexp:0,1,50;dss:1;dump:dss,1;copy:1,2;"
            + (per % (2, eps) if eps else "") + "vimp:2,2,30;dump:vi,2;"
            + "copy:0,1;copy:0,2;copy:0,3;copy:0,4;" + (per % (0, eps) if eps else "")
            + "step:%d;dump:st,0" % steps, c["flags"])
    base = run(None)
    runs = [run(e) for e in EPS]
    npatch = refdump.scalar(base, "grid.npatch")
    out = {"field_rel_spread": [], "implicit_stage_spread": []}
    for loc, cc in where:
        num = den = inum = iden = 0.0
        for n in range(npatch):
            ref = base["st.patch%d.inst0.%s" % (n, loc)][cc]
            vi = base["vi.patch%d.inst2.%s" % (n, loc)][cc]
            bef = base["dss.patch%d.inst1.%s" % (n, loc)][cc]
            den = max(den, np.abs(ref).max())
            iden = max(iden, np.abs(vi - bef).max())
            for r in runs:
                num = max(num, np.abs(r["st.patch%d.inst0.%s" % (n, loc)][cc] - ref).max())
                inum = max(inum, np.abs(r["vi.patch%d.inst2.%s" % (n, loc)][cc] - vi).max())
        out["field_rel_spread"].append(num / den)
        out["implicit_stage_spread"].append(inum / iden if iden > 0 else 0.0)
    return out


ENTRIES = {
    # the configuration of tests/test_dropin.py::test_nonhydro_dropin
    # (integration/b200_driver.cpp defaults: ztop = 10 km)
    "jw_ne8_l10_strang_3steps": lambda: checksum_spread(
        "jw", ["--resolution", "8", "--levels", "10", "--dt", "200s", "--ztop", "10000",
               "--timescheme", "strang"], 3),
    "jw_ne8_l10_ars343_3steps": lambda: checksum_spread(
        "jw", ["--resolution", "8", "--levels", "10", "--dt", "200s", "--ztop", "10000",
               "--timescheme", "ars343"], 3),
    # tests/test_dropin.py::test_cartesian_bubble_dropin
    "bubble_r36_l72_20steps": lambda: checksum_spread(
        "bubble", ["--resolution", "36", "--resy", "1", "--levels", "72", "--dt", "10000u",
                   "--nohypervis"], 20, npatch=1),
    # tests/test_configs.py::test_config3 (configuration 3 itself)
    "jw_ne30_l30_strang_2steps": lambda: checksum_spread(
        "jw", ["--resolution", "30", "--levels", "30", "--dt", "200s"], 2),
    # tests/test_parity.py::test_tracers_ars343
    "jwtr_ne2_l6_ars343_2steps": lambda: tracer_spread("jwtr_ne2_l6_ars343"),
    # tests/test_parity.py::test_vertical_order_above_one
    "jw_ne2_l12_vo2_2steps": lambda: state_spread("jw_ne2_l12_vo2"),
    "jw_ne2_l24_vo3_2steps": lambda: state_spread("jw_ne2_l24_vo3"),
    "jw_ne2_l24_vo4_2steps": lambda: state_spread("jw_ne2_l24_vo4"),
    # the same at vertical order 1, for comparison
    "jw_ne2_l6_strang_2steps": lambda: state_spread("jw_ne2_l6_strang"),
}


def load():
    with open(OUT) as f:
        return json.load(f)


if __name__ == "__main__":
    # `make_sensitivity.py` measures the entries the file does not hold yet,
    # `make_sensitivity.py --all` every entry
    res = {"eps": EPS}
    if os.path.exists(OUT) and "--all" not in sys.argv:
        res = load()
        assert res["eps"] == EPS
    for key, fn in ENTRIES.items():
        if key not in res:
            res[key] = fn()
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print(json.dumps(res, indent=1, sort_keys=True))
