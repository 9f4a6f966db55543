"""Development aid: which operation of a step first gives different bits on 6 and
on 24 patches (one rank)?  python tools/decomp_ops.py NE L [library]"""
import sys, numpy as np
sys.path.insert(0,'/root/repo')
from tempestmodel_b200 import grid as G, testcases as TC
from tempestmodel_b200.model import Model
ne, L = int(sys.argv[1]), int(sys.argv[2])
lib = sys.argv[3] if len(sys.argv) > 3 else '/root/repo/tests/emu/libtb200_emu.so'
if lib == 'cuda':
    lib = None
dt = 200.0 * 20 / ne

def panels(model, grid, inst):
    out = {}
    for p in model.local:
        node = np.zeros((5, p.wa, p.wb, L)); redge = np.zeros((5, p.wa, p.wb, L + 1))
        model.ctx.download_state(p.index, inst, node, redge, None, False)
        sl = (slice(None), slice(4*p.ea0, 4*(p.ea0+p.nea)), slice(4*p.eb0, 4*(p.eb0+p.neb)))
        a = out.setdefault(p.panel, (np.zeros((5, 4*ne, 4*ne, L)), np.zeros((5, 4*ne, 4*ne, L+1))))
        a[0][sl] = node[:, 1:-1, 1:-1]; a[1][sl] = redge[:, 1:-1, 1:-1]
    return out

rec = []
for npatch in (6, 6, 24):
    grid = G.GridCSGLL(ne, L, npatch=npatch, ztop=30000.0)
    model = Model(grid, TC.BaroclinicWaveJWTest(ztop=30000.0, perturbation="exp"), timescheme="strang", dt=dt, library=lib)
    model.initialize()
    ctx = model.ctx
    r = {}
    r["ic"] = panels(model, grid, 0)
    for m in range(1, ctx.cfg.ninstances): ctx.copy(0, m)
    ctx.hv_step_explicit_combine([1.0, 0.0], 0, 1, dt / 5)
    r["stage"] = panels(model, grid, 1)
    ctx.dss(1)
    r["dss"] = panels(model, grid, 1)
    ctx.copy(1, 2)
    ctx.v_step_implicit(2, 2, 0.5 * dt)
    r["vimp"] = panels(model, grid, 2)
    ctx.h_step_after_subcycle(1, 3, 4, dt)
    r["hasc"] = panels(model, grid, 3)
    ctx.copy(0, 1)
    for s in range(3):
        model.step(1)
        r["step%d" % (s + 1)] = panels(model, grid, 0)
    rec.append(r)
    ctx.close()
for (i0, i1, what) in ((0, 1, "6 vs 6 patches (same run twice)"), (0, 2, "6 vs 24 patches")):
  for k in rec[0]:
    worst = 0.0
    where = None
    for p in range(6):
        for loc in (0, 1):
            a, b = rec[i0][k][p][loc], rec[i1][k][p][loc]
            for c in range(5):
                s = np.abs(a[c]).max()
                if s > 0:
                    dd = np.abs(a[c] - b[c])
                    if dd.max() / s > worst:
                        worst = dd.max() / s
                        where = (p, loc, c) + tuple(int(v) for v in np.unravel_index(dd.argmax(), dd.shape))
    print("%-8s %s: %.3e at (panel, loc, comp, ia, ib, k) = %s" % (k, what, worst, where))
