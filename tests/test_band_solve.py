"""Pin the device band solver (the LAPACK dgbsv = dgbtf2 + dgbtrs restatement
used by the implicit column solve, tb200_column.cuh) against LAPACK itself.

The arithmetic of dgbsv is third party (reference LAPACK as shipped in the
OpenBLAS the oracle links; version unpinned by the reference,
mk/system/*.make: -llapack; call site src/base/LinearAlgebra.cpp:196 <-
VerticalDynamicsFEM.cpp:1457).  No reference test pins results at that
boundary, so the solver is checked against scipy's dgbsv to a tolerance set by
the conditioning, not bitwise."""
import numpy as np
import pytest
from scipy.linalg import lapack

from tempestmodel_b200 import DeviceContext

BACKENDS = [pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def library(request):
    if request.param == "emu":
        return request.getfixturevalue("emu_library")
    return request.getfixturevalue("cuda_library")


def _ctx(library):
    return DeviceContext(
        library=library, np=4, nlev=4, vertical_order=1, ncomp=3, ntracers=0,
        ninstances=2, eqn_type=1, cartesian_xz=0, device=-1, g=9.8, R=287.0,
        cp=1004.5, cv=717.5, p0=1e5, omega=0.0, earth_radius=1.0, ztop=1.0,
        ref_length=1.0, hypervis_order=0, nu_scalar=0.0, nu_div=0.0,
        nu_vort=0.0, fully_explicit=0, off_centering=0.0)


@pytest.mark.parametrize("n,kl,ku,ncols", [(93, 4, 4, 40), (219, 4, 4, 7), (30, 1, 1, 65), (12, 9, 9, 3)])
def test_band_solve_matches_lapack(library, n, kl, ku, ncols):
    rng = np.random.default_rng(1234 + n)
    ldab = 2 * kl + ku + 1
    # reference storage: row-major [n][ldab] handed to Fortran as AB(ldab, n)
    ab = np.zeros((ncols, n, ldab))
    ab[:, :, kl:] = rng.standard_normal((ncols, n, ldab - kl))
    # weak diagonal so that partial pivoting really swaps rows
    ab[:, :, kl + ku] *= 0.05
    b = rng.standard_normal((ncols, n))
    ctx = _ctx(library)
    x = ctx.test_band_solve(ab, b, kl, ku)
    ctx.close()
    for c in range(ncols):
        # scipy wants AB(ldab, n) column-major == our [n][ldab] transposed
        lub, piv, xr, info = lapack.dgbsv(kl, ku, ab[c].T.copy(order="F"), b[c].copy())
        assert info == 0
        scale = np.abs(xr).max()
        assert np.abs(x[c] - xr).max() <= 1e-9 * scale
        # residual of our solution in the original system
        A = np.zeros((n, n))
        for j in range(n):
            for i in range(max(0, j - ku), min(n, j + kl + 1)):
                A[i, j] = ab[c, j, kl + ku + i - j]
        r = A @ x[c] - b[c]
        assert np.abs(r).max() <= 1e-10 * max(1.0, np.abs(A).max() * scale)
