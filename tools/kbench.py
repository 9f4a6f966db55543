#!/usr/bin/env python
"""Kernel micro-benchmark (development aid; bench.py is the contract):

    python tools/kbench.py [--ne 120] [--levels 30] [--lib path.so] [--reps 5]

Times each C-ABI operation of one time step alone with CUDA events on the
context's stream (JW baroclinic wave, device-resident state) and prints
ms per call and the algorithmic GB/s of SURVEY 8(d)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ne", type=int, default=120)
    ap.add_argument("--levels", type=int, default=30)
    ap.add_argument("--lib", default=None)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--timescheme", default="strang")
    ap.add_argument("--tag", default="")
    ap.add_argument("--only", default=None, help="run the operations whose name contains this")
    ap.add_argument("--tracers", type=int, default=0,
                    help="carry N analytic tracers (S = 5 + N)")
    args = ap.parse_args()

    import torch
    from tempestmodel_b200 import grid as G
    from tempestmodel_b200 import testcases as TC
    from tempestmodel_b200.model import Model

    ne, L = args.ne, args.levels
    dt = 200.0 * 20.0 / ne
    torch.cuda.set_device(0)
    t0 = time.time()
    grid = G.GridCSGLL(ne, L, npatch=6, ztop=30000.0)
    if args.tracers > 0:
        test = TC.BaroclinicWaveJWTracerTest(ntracers=args.tracers, ztop=30000.0,
                                             perturbation="exp")
    else:
        test = TC.BaroclinicWaveJWTest(ztop=30000.0, perturbation="exp")
    model = Model(grid, test,
                  timescheme=args.timescheme, dt=dt, device=0, library=args.lib)
    ctx = model.ctx
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    model.device_setup = True
    model.initialize()
    model._host = {}
    model._host_tracers = {}
    ctx.sync()
    setup = time.time() - t0
    fast = ctx.fast_path() if hasattr(ctx, "fast_path") else None
    model.step(2)
    ctx.check_errors()

    nodes = ctx.column_count * L
    S = (5 + args.tracers) * 8

    ninst = ctx.cfg.ninstances

    def timeit(fn, reps=args.reps):
        # every instance holds a valid state (work instances carry Laplacians
        # after a step)
        for m in range(1, ninst):
            ctx.copy(0, m)
        fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ops = [
        ("stage(in=base)+HV 2S", lambda: ctx.hv_step_explicit_combine([0.0, 0.0, 1.0, 0.0], 2, 3, 1e-6), 2),
        ("stage(copy0)+HV 3S", lambda: ctx.hv_step_explicit_combine([1.0, 0.0, 0.0, 0.0], 2, 3, 1e-6), 3),
        ("stage(2src)+HV 4S", lambda: ctx.hv_step_explicit_combine([-0.25, 1.25, 0.0, 0.0, 0.0], 2, 4, 1e-6), 4),
        ("dss 1.5S", lambda: ctx.dss(3), 1.5),
        ("stage 2S + dss (substage)", lambda: ctx.hv_step_explicit_combine_dss([0.0, 0.0, 1.0, 0.0], 2, 3, 1e-6), 3.5),
        ("stage 3S + dss (substage)", lambda: ctx.hv_step_explicit_combine_dss([1.0, 0.0, 0.0, 0.0], 2, 3, 1e-6), 4.5),
        ("stage 4S + dss (substage)", lambda: ctx.hv_step_explicit_combine_dss([-0.25, 1.25, 0.0, 0.0, 0.0], 2, 4, 1e-6), 5.5),
        ("implicit 2S", lambda: ctx.v_step_implicit(3, 3, dt * 1e-3), 2),
        ("hyperdiffusion 8S", lambda: ctx.h_step_after_subcycle(4, 1, 2, dt * 1e-3), 8),
        ("lincomb(2src) 3S", lambda: ctx.lincomb([0.5, 0.5], 1), 3),
        ("copy 2S", lambda: ctx.copy(1, 2), 2),
        # (instance 1 is the Strang carry-over increment: zero, not a copy of the state)
        ("full step 35.5S", lambda: (ctx.zero(1), model.step(1)), 35.5),
    ]
    print("fast path:", fast, flush=True)
    out = {"tag": args.tag, "ne": ne, "L": L, "setup_s": setup, "fast_path": fast, "ops": {}}
    for name, fn, s in ops:
        if args.only is not None and args.only not in name:
            continue
        ms = timeit(fn)
        gbs = nodes * s * S / (ms * 1e-3) / 1e9
        out["ops"][name] = {"ms": round(ms, 4), "alg_GBs": round(gbs, 1)}
        print("%-24s %9.3f ms   %8.1f GB/s (algorithmic)" % (name, ms, gbs), flush=True)
    ctx.check_errors()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
