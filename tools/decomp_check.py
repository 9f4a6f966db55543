"""Development aid: the same run on 6 and on 24 patches (one rank) must give the
same bits.  python tools/decomp_check.py NE L NSTEPS [library] [device_setup]"""
import sys, numpy as np
sys.path.insert(0,'/root/repo')
from tempestmodel_b200 import grid as G, testcases as TC
from tempestmodel_b200.model import Model
ne, L, nsteps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
lib = sys.argv[4] if len(sys.argv) > 4 else '/root/repo/tests/emu/libtb200_emu.so'
if lib == 'cuda':
    lib = None
devset = len(sys.argv) > 5 and sys.argv[5] == '1'
res=[]
for npatch in (6,24):
    grid = G.GridCSGLL(ne, L, npatch=npatch, ztop=30000.0)
    model = Model(grid, TC.BaroclinicWaveJWTest(ztop=30000.0, perturbation="exp"), timescheme="strang", dt=200.0*20/ne, library=lib)
    model.device_setup = devset
    model.initialize()
    model.step(nsteps)
    st = model.download_state(0)
    nodes = {p: np.zeros((5, 4*ne, 4*ne, L)) for p in range(6)}
    for p in grid.patches:
        node, redge = st[p.index]
        sl = (slice(None), slice(4*p.ea0, 4*(p.ea0+p.nea)), slice(4*p.eb0, 4*(p.eb0+p.neb)))
        nodes[p.panel][sl] = node[:, 1:-1, 1:-1]
    res.append(nodes); print(npatch, model.checksum(0)); model.ctx.close()
for c in (0,1,2,4):
    m = max(np.abs(res[0][p][c]-res[1][p][c]).max() for p in range(6)); s = max(np.abs(res[0][p][c]).max() for p in range(6))
    print("comp", c, "maxdiff", m, "rel", m/s)
