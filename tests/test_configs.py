"""The reference's own configurations through the Python driver (numpy grid and
initial conditions + device kernels), against the golden checksums the
unmodified reference prints for them (regenerated with
oracle/_ref/BaroclinicWaveJWTest / SWTest2)."""
import numpy as np
import pytest

from tempestmodel_b200 import grid as G
from tempestmodel_b200 import testcases as TC
from tempestmodel_b200.model import Model

# BaroclinicWaveJWTest --resolution 30 --levels 30 --dt 200s --endtime 400s
#   --pert Exp --ztop 30000 --output_none : checksums after 2 steps
CONFIG3 = dict(U=7.520878775555240e+26, V=4.418289088264086e+21,
               RhoTheta=1.741943050815948e+21, W=4.476383903565237e+21,
               Rho=5.127026948774204e+18)


@pytest.mark.gpu
def test_config3_jw_ne30_l30_checksums(cuda_library):
    grid = G.GridCSGLL(30, 30, npatch=6, ztop=30000.0)
    model = Model(grid, TC.BaroclinicWaveJWTest(ztop=30000.0, perturbation="exp"),
                  timescheme="strang", dt=200.0, library=cuda_library)
    model.initialize()
    assert model.ctx.fast_path()[0]
    model.step(2, last=False)
    model.ctx.check_errors()
    cs = model.checksum(0)
    # mass and rho-theta: conserved quantities, independent of the sign noise of
    # the implicit Jacobian at zero wind
    assert abs(cs[4] - CONFIG3["Rho"]) <= 1e-12 * abs(CONFIG3["Rho"])
    assert abs(cs[2] - CONFIG3["RhoTheta"]) <= 1e-12 * abs(CONFIG3["RhoTheta"])
    assert abs(cs[0] - CONFIG3["U"]) <= 1e-7 * abs(CONFIG3["U"])
    model.ctx.close()

# SWTest2 --resolution 20 --order 4 --output_none (dt = 200 s, 1 step, strang,
# hypervis 4, nu = 1e15): checksums after the step
CONFIG1 = dict(U=7.114413176809416e+22, V=1.651200000000000e+06, H=1.205365996298435e+18)


@pytest.mark.gpu
def test_config1_sw2_ne20_checksums(cuda_library):
    grid = G.GridCSGLL(20, 1, npatch=6, ztop=1.0)
    model = Model(grid, TC.ShallowWaterTestCase2(), timescheme="strang", dt=200.0,
                  library=cuda_library)
    model.initialize()
    model.step(1, last=True)
    model.ctx.check_errors()
    cs = model.checksum(0)
    assert abs(cs[0] - CONFIG1["U"]) <= 1e-12 * abs(CONFIG1["U"])
    assert abs(cs[2] - CONFIG1["H"]) <= 1e-13 * abs(CONFIG1["H"])
    # V sums to rounding noise of U-sized terms
    assert abs(cs[1] - CONFIG1["V"]) <= 1e-12 * abs(CONFIG1["U"])
    model.ctx.close()
