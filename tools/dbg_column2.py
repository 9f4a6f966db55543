import sys, ctypes
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, cases, dumpctx, tempestmodel_b200
d = cases.load_case("jw_ne2_l6")
import torch
libs = {'emu': dumpctx.EMU_LIBRARY}
if torch.cuda.is_available(): libs['cuda'] = tempestmodel_b200.PRODUCT_LIBRARY
res = {}
for name, lib in libs.items():
    ctx = dumpctx.context_from_dump(d, library=lib)
    dumpctx.upload_tag(ctx, d, "dss", instances=[1])
    ctx.copy(1, 2)
    ctx.v_step_implicit(2, 2, 30.0); ctx.check_errors()
    res[name] = dumpctx.download(ctx, d, 2)
    ref = {n: (d['vi.patch%d.inst2.node' % n], d['vi.patch%d.inst2.redge' % n]) for n in range(6)}
    for n in range(6):
        for loc, c in ((0, 2), (0, 4), (1, 3)):
            a = res[name][n][loc][c, 1:-1, 1:-1]; r = ref[n][loc][c, 1:-1, 1:-1]
            e = np.abs(a - r); sc = np.abs(r).max()
            bad = np.argwhere(e > 1e-9 * sc)
            if len(bad):
                print(name, 'patch', n, 'loc', loc, 'comp', c, 'nbad', len(bad), 'of', e.size, 'first', bad[:8].tolist())
                i, j, k = bad[0]
                print('    dev', a[i, j], '\n    ref', r[i, j])
