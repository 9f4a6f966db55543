#!/usr/bin/env python
"""Development aid: state after a few steps with the DSS fused into the stage /
hyperdiffusion kernels against the separate DSS pass, bit for bit, for a given
build of the library:  python tools/fuse_check.py [--lib path.so] [--ne 30]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=None)
    ap.add_argument("--ne", type=int, default=30)
    ap.add_argument("--levels", type=int, default=30)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--npatch", type=int, default=6)
    args = ap.parse_args()
    from tempestmodel_b200 import grid as G
    from tempestmodel_b200 import testcases as TC
    from tempestmodel_b200.model import Model
    res = []
    for fused in ("0", "1"):
        os.environ["TB200_DSS_FUSED"] = fused
        grid = G.GridCSGLL(args.ne, args.levels, npatch=args.npatch, ztop=30000.0)
        model = Model(grid, TC.BaroclinicWaveJWTest(ztop=30000.0, perturbation="exp"),
                      timescheme="strang", dt=200.0 * 20.0 / args.ne, library=args.lib)
        model.device_setup = True
        model.initialize()
        model.step(args.steps, check=False)
        res.append(model.download_state(0) if True else None)
        model.ctx.close()
    worst = 0.0
    nbad = 0
    for idx in res[0]:
        for loc in (0, 1):
            a, b = res[0][idx][loc], res[1][idx][loc]
            bad = (a != b) & ~(np.isnan(a) & np.isnan(b))
            nbad += int(bad.sum())
            if bad.any():
                worst = max(worst, float(np.nanmax(np.abs(a - b)[bad] / (np.abs(a[bad]) + 1e-300))))
    print("fuse_check lib=%s ne=%d: differing values %d, worst relative %.3e"
          % (args.lib, args.ne, nbad, worst))


if __name__ == "__main__":
    main()
