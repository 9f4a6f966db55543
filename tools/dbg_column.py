import sys, ctypes
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, cases, dumpctx, tempestmodel_b200
from scipy.linalg import lapack
d = cases.load_case("jw_ne2_l6")
L = 6; n = 3*(L+1); total = 24*(L+1) + 2*n + n*13
libs = {'emu': dumpctx.EMU_LIBRARY}
import torch
if torch.cuda.is_available(): libs['cuda'] = tempestmodel_b200.PRODUCT_LIBRARY
for name, lib in libs.items():
    ctx = dumpctx.context_from_dump(d, library=lib)
    dumpctx.upload_tag(ctx, d, "dss", instances=[1])
    for col in [0, 17, 100]:
        ws = np.zeros(total); ws2 = np.zeros(total)
        ctx._ck(ctx.lib.tb200_debug_column_assembly(ctx._h, 1, 30.0, col, ctypes.c_void_p(ws.ctypes.data), total))
        ctx._ck(ctx.lib.tb200_debug_column_assembly(ctx._h, 1, 30.0, col, ctypes.c_void_p(ws2.ctypes.data), -total))
        o = 24*(L+1)
        x0 = ws[o:o+n]; F = ws[o+n:o+2*n]; DG = ws[o+2*n:o+2*n+n*13].reshape(n, 13)
        delta = ws2[o+n:o+2*n]
        lub, piv, xr, info = lapack.dgbsv(4, 4, DG.T.copy(order='F'), F.copy())
        xb = ctx.test_band_solve(DG[None], F[None], 4, 4)[0]
        print(name, col, 'info', info, 'kernel-vs-lapack', np.abs(delta-xr).max()/np.abs(xr).max(),
              'standalone-vs-lapack', np.abs(xb-xr).max()/np.abs(xr).max(), 'max delta', np.abs(xr).max())
        if np.abs(delta-xr).max() > 1e-8*np.abs(xr).max():
            print('   delta', delta[:12]); print('   ref  ', xr[:12])
