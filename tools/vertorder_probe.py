"""Probe a vertical order that has no committed fixture: run the unmodified
reference (oracle/_ref/ref_dump, needs /root/reference) on the JW case at
`--vertorder VO --levels L`, compare the explicit stages and the implicit stage
of the emulation build with it, and measure the reference's own spread of the
implicit stage under 1-ulp perturbations of its input.

    python tools/vertorder_probe.py VO L [thread|warp]

Results of this round: order 3, L = 24: implicit stage
within 4e-11 of the largest change; order 5, L = 30: 3e-8 / 2e-7 / 4e-8
(rho theta / rho / w) against a reference spread of 3e-8 / 2e-7 / 4e-8."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import cases  # noqa: E402
import dumpctx  # noqa: E402
import refdump  # noqa: E402
from test_parity import tendency_errors  # noqa: E402

vo, L = sys.argv[1], sys.argv[2]
os.environ["TB200_COLUMN_KERNEL"] = sys.argv[3] if len(sys.argv) > 3 else "warp"
flags = ["--resolution", "2", "--levels", L, "--vertorder", vo, "--dt", "200s"]
HEAD = "addw:0,20000;dss:0;dump:ic,0;copy:0,1;hexp:0,1,50;dump:h1,1;vexp:0,1,50;dump:v1,1;" \
       "dss:1;dump:dss,1;copy:1,2;"


def run(eps):
    pre = HEAD + ("perturb:2,%s;" % eps if eps else "")
    return refdump.run_ref_dump("/tmp/tb200_voprobe.bin", "jw", pre + "vimp:2,2,30;dump:vi,2", flags)


d = run(None)
ctx = dumpctx.context_from_dump(d, library=dumpctx.EMU_LIBRARY)
dumpctx.upload_tag(ctx, d, "ic")
ctx.copy(0, 1)
ctx.h_step_explicit(0, 1, 50.0)
print("h1", max(tendency_errors(ctx, d, 1, "h1", "ic", 0, [0, 1, 2, 4], [3]).values()))
ctx.v_step_explicit(0, 1, 50.0)
print("v1", max(tendency_errors(ctx, d, 1, "v1", "ic", 0, [0, 1, 2, 4], [3]).values()))
dumpctx.upload_tag(ctx, d, "dss", instances=[1])
ctx.copy(1, 2)
ctx.v_step_implicit(2, 2, 30.0)
ctx.check_errors()
print("implicit stage, device vs reference:",
      tendency_errors(ctx, d, 2, "vi", "dss", 1, [2, 4], [3], skip_poles=True))
npatch = refdump.scalar(d, "grid.npatch")
spread = {}
for eps in ("1e-15", "-1e-15", "2e-15"):
    r = run(eps)
    for loc, cc in (("node", 2), ("node", 4), ("redge", 3)):
        num = den = 0.0
        for n in range(npatch):
            a = d["vi.patch%d.inst2.%s" % (n, loc)][cc]
            b = r["vi.patch%d.inst2.%s" % (n, loc)][cc]
            bef = d["dss.patch%d.inst1.%s" % (n, loc)][cc]
            num = max(num, np.abs(a - b).max())
            den = max(den, np.abs(a - bef).max())
        spread[(loc, cc)] = max(spread.get((loc, cc), 0.0), num / den)
print("implicit stage, reference vs perturbed reference:", spread)
