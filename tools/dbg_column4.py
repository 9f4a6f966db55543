import sys, ctypes
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, cases, dumpctx, tempestmodel_b200
d = cases.load_case("jw_ne2_l6")
lib = tempestmodel_b200.PRODUCT_LIBRARY if len(sys.argv) < 2 else dumpctx.EMU_LIBRARY
L = 6; n = 3*(L+1); total = 24*(L+1) + 2*n + n*13
names = "snU snV snP snW snR seU seV seW seR seP exn dPe xdn xde mfe dmfn pfe dpfn ken dkee dUa dUb ddW aux".split()
def getws(ctx, inst, col, solve=False):
    ws = np.zeros(total)
    ctx._ck(ctx.lib.tb200_debug_column_assembly(ctx._h, inst, 30.0, col, ctypes.c_void_p(ws.ctypes.data), -total if solve else total))
    return ws
# B: computed input
ctxB = dumpctx.context_from_dump(d, library=lib)
dumpctx.upload_tag(ctxB, d, "ic")
ctxB.copy(0, 1); ctxB.hv_step_explicit(0, 1, 50.0); ctxB.dss(1)
ctxC = dumpctx.context_from_dump(d, library=lib)
dumpctx.upload_tag(ctxC, d, "dss", instances=[1])
worst = (0, None)
for col in range(294):
    a = getws(ctxB, 1, col, True); b = getws(ctxC, 1, col, True)
    o = 24*(L+1)
    da = a[o+n:o+2*n]; db = b[o+n:o+2*n]
    e = np.abs(da-db).max()/max(np.abs(db).max(), 1e-300)
    if e > worst[0]: worst = (e, col)
print('worst column', worst)
col = worst[1]
a = getws(ctxB, 1, col); b = getws(ctxC, 1, col)
for i, nm in enumerate(names):
    sa, sb = a[i*(L+1):(i+1)*(L+1)], b[i*(L+1):(i+1)*(L+1)]
    print(nm, np.abs(sa-sb).max(), '\n   B', sa, '\n   C', sb)
o = 24*(L+1)
for nm, ln in (('x0', n), ('F', n), ('DG', n*13)):
    sa, sb = a[o:o+ln], b[o:o+ln]; o += ln
    print(nm, np.abs(sa-sb).max(), np.abs(sb).max())
    if nm == 'F': print('  B', sa, '\n  C', sb)
