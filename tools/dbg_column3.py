import sys, ctypes
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, cases, dumpctx, tempestmodel_b200
d = cases.load_case("jw_ne2_l6")
lib = tempestmodel_b200.PRODUCT_LIBRARY
def werr(ctx):
    return dumpctx.compare(ctx, d, 2, "vi", [2, 4], [3])
# A: uploaded input, download before implicit
ctx = dumpctx.context_from_dump(d, library=lib)
dumpctx.upload_tag(ctx, d, "dss", instances=[1])
dumpctx.download(ctx, d, 1)
ctx.copy(1, 2); ctx.v_step_implicit(2, 2, 30.0); ctx.check_errors()
print('A uploaded+download', werr(ctx))
# B: computed input, no download
ctx = dumpctx.context_from_dump(d, library=lib)
dumpctx.upload_tag(ctx, d, "ic")
ctx.copy(0, 1); ctx.hv_step_explicit(0, 1, 50.0); ctx.dss(1)
ctx.copy(1, 2); ctx.v_step_implicit(2, 2, 30.0); ctx.check_errors()
print('B computed, no download', werr(ctx))
print('B inst1 vs dss', dumpctx.compare(ctx, d, 1, "dss", [0, 1, 2, 4], [3]))
# C: computed input, then overwrite with upload
ctx = dumpctx.context_from_dump(d, library=lib)
dumpctx.upload_tag(ctx, d, "ic")
ctx.copy(0, 1); ctx.hv_step_explicit(0, 1, 50.0); ctx.dss(1)
dumpctx.upload_tag(ctx, d, "dss", instances=[1])
ctx.copy(1, 2); ctx.v_step_implicit(2, 2, 30.0); ctx.check_errors()
print('C computed then uploaded', werr(ctx))
# D: computed input; exact bitwise comparison of device inst1 with dss dump incl. location of max diff
ctx = dumpctx.context_from_dump(d, library=lib)
dumpctx.upload_tag(ctx, d, "ic")
ctx.copy(0, 1); ctx.hv_step_explicit(0, 1, 50.0); ctx.dss(1)
got = dumpctx.download(ctx, d, 1)
for n in range(6):
    for loc, nm in ((0, 'node'), (1, 'redge')):
        r = d['dss.patch%d.inst1.%s' % (n, nm)][:, 1:-1, 1:-1]; a = got[n][loc][:, 1:-1, 1:-1]
        for c in range(5):
            if (loc == 0) == (c == 3): continue
            e = np.abs(a[c] - r[c]); i = np.unravel_index(e.argmax(), e.shape)
            print('D patch', n, nm, c, 'maxabs', e.max(), 'at', i, 'ref', r[c][i], 'scale', np.abs(r[c]).max())
