"""Output-side interpolation on the device: tb200_interpolate against
Grid::ReduceInterpolate of the unmodified reference (fixture jwtr_ne2_l6_interp:
state at both locations and tracers on a 12 x 7 latitude-longitude grid and five
uniform REta levels, with and without the conversion to primitive variables).

The element search and the Lagrangian coefficients are the caller's part of the
interface (GridPatchCSGLL.cpp:1605-1641, PolynomialInterp.cpp:26-47); they are
restated here.  1e-12 of each field."""
import numpy as np
import pytest

import cases
import dumpctx
from tempestmodel_b200._lib import DATA_STATE, DATA_TRACERS
from test_parity import BACKENDS


@pytest.fixture(params=BACKENDS)
def library(request):
    if request.param == "emu":
        return request.getfixturevalue("emu_library")
    return request.getfixturevalue("cuda_library")


def lagrangian_coeffs(x, xs):
    """PolynomialInterp::LagrangianPolynomialCoeffs."""
    n = len(x)
    c = np.ones(n)
    for i in range(n):
        for j in range(n):
            if i != j:
                c[i] *= (xs - x[j]) / (x[i] - x[j])
    return c


def locate(d, tag, np_=4):
    """Patch-local element and coefficients of every point (GridPatchCSGLL.cpp:1594-1641)."""
    alpha, beta, ipatch = d[tag + ".alpha"], d[tag + ".beta"], d[tag + ".ipatch"]
    byindex = {dumpctx.S(d, "patch%d.index" % n): n for n in range(dumpctx.S(d, "grid.npatch"))}
    ea, eb, ca, cb = [], [], [], []
    for a, b, p in zip(alpha, beta, ipatch):
        n = byindex[int(p)]
        anode, bnode = d["patch%d.anode" % n], d["patch%d.bnode" % n]
        halo = dumpctx.S(d, "patch%d.halo" % n)
        da, db = dumpctx.S(d, "patch%d.delta_a" % n), dumpctx.S(d, "patch%d.delta_b" % n)
        nea, neb = dumpctx.S(d, "patch%d.nelem_a" % n), dumpctx.S(d, "patch%d.nelem_b" % n)
        ia = min(max(int((a - anode[halo]) / da), 0), nea - 1)
        ib = min(max(int((b - bnode[halo]) / db), 0), neb - 1)
        ea.append(ia)
        eb.append(ib)
        ca.append(lagrangian_coeffs(anode[halo + ia * np_: halo + (ia + 1) * np_], a))
        cb.append(lagrangian_coeffs(bnode[halo + ib * np_: halo + (ib + 1) * np_], b))
    return alpha, beta, ipatch, np.array(ea), np.array(eb), np.array(ca), np.array(cb)


def vop(d, tag, name):
    return (d["%s.%s.coeff" % (tag, name)], d["%s.%s.begin" % (tag, name)],
            d["%s.%s.end" % (tag, name)])


@pytest.mark.parametrize("tag,primitive", [("raw", False), ("prim", True)])
def test_interpolate_state_and_tracers(library, tag, primitive):
    d = cases.load_case("jwtr_ne2_l6_interp")
    ctx = dumpctx.context_from_dump(d, library=library)
    dumpctx.upload_tag(ctx, d, "ic")
    alpha, beta, ipatch, ea, eb, ca, cb = locate(d, tag)
    nout = len(d[tag + ".reta"])
    vn, ve = vop(d, tag, "vop_node"), vop(d, tag, "vop_redge")
    node = ctx.interpolate(0, DATA_STATE, 0, ipatch, ea, eb, ca, cb, alpha, beta, nout,
                           vop_node=vn, vop_redge=ve, primitive=primitive)
    redge = ctx.interpolate(0, DATA_STATE, 1, ipatch, ea, eb, ca, cb, alpha, beta, nout,
                            vop_node=vn, vop_redge=ve, primitive=primitive)
    trac = ctx.interpolate(0, DATA_TRACERS, -1, ipatch, ea, eb, ca, cb, alpha, beta, nout,
                           vop_node=vn, vop_redge=ve, primitive=primitive)
    ref_node, ref_redge, ref_tr = d[tag + ".node"], d[tag + ".redge"], d[tag + ".tracers"]
    # components on levels: u, v, rho theta, rho; on interfaces: w
    scale_uv = max(np.abs(ref_node[0]).max(), np.abs(ref_node[1]).max())
    for c in (0, 1):
        assert np.abs(node[c] - ref_node[c]).max() <= 1e-12 * scale_uv, c
    for c in (2, 4):
        assert np.abs(node[c] - ref_node[c]).max() <= 1e-12 * np.abs(ref_node[c]).max(), c
    assert np.all(node[3] == 0.0) and np.all(ref_node[3] == 0.0)
    assert np.abs(redge[3] - ref_redge[3]).max() <= 1e-12 * np.abs(ref_redge[3]).max()
    for c in (0, 1, 2, 4):
        assert np.all(redge[c] == 0.0)
    for c in range(ref_tr.shape[0]):
        assert np.abs(trac[c] - ref_tr[c]).max() <= 1e-12 * np.abs(ref_tr[c]).max(), c
    ctx.close()


def test_interpolate_identity_operator_returns_the_levels(library):
    """Without a column operator the rows come back on the model levels: a point on
    a GLL node reproduces the node's column."""
    d = cases.load_case("jwtr_ne2_l6_interp")
    ctx = dumpctx.context_from_dump(d, library=library)
    dumpctx.upload_tag(ctx, d, "ic")
    n = ctx.local_patches[0]
    idx = dumpctx.S(d, "patch%d.index" % n)
    halo = dumpctx.S(d, "patch%d.halo" % n)
    anode, bnode = d["patch%d.anode" % n], d["patch%d.bnode" % n]
    L = dumpctx.S(d, "grid.nlev")
    # node (element 1, i = 2; element 0, j = 1) of the patch
    a, b = anode[halo + 4 + 2], bnode[halo + 1]
    ca = lagrangian_coeffs(anode[halo + 4: halo + 8], a)[None, :]
    cb = lagrangian_coeffs(bnode[halo: halo + 4], b)[None, :]
    out = ctx.interpolate(0, DATA_STATE, 0, [idx], [1], [0], ca, cb, [a], [b], L,
                          primitive=False)
    ref = d["ic.patch%d.inst0.node" % n][:, halo + 6, halo + 1, :]
    for c in (0, 2, 4):
        assert np.abs(out[c, :, 0] - ref[c]).max() <= 1e-13 * np.abs(ref[c]).max(), c
    ctx.close()


def test_interpolate_derived_fields(emu_library):
    """Relative vorticity, divergence (GridPatchCSGLL::ComputeCurlAndDiv) and
    temperature (GridPatch::ComputeTemperature) computed on the device and
    interpolated, against the reference's ComputeVorticityDivergence /
    ComputeTemperature + ReduceInterpolate.  Vorticity and divergence are
    differences of u-sized terms: held to 1e-11 of the field.  Emulation backend
    (written after the GPU budget of the round was spent)."""
    from tempestmodel_b200._lib import DATA_DIVERGENCE, DATA_TEMPERATURE, DATA_VORTICITY
    d = cases.load_case("jwtr_ne2_l6_interp")
    ctx = dumpctx.context_from_dump(d, library=emu_library)
    dumpctx.upload_tag(ctx, d, "ic")
    alpha, beta, ipatch, ea, eb, ca, cb = locate(d, "raw")
    nout = len(d["raw.reta"])
    vn, ve = vop(d, "raw", "vop_node"), vop(d, "raw", "vop_redge")
    with pytest.raises(Exception, match="output fields not computed"):
        ctx.interpolate(0, DATA_VORTICITY, -1, ipatch, ea, eb, ca, cb, alpha, beta, nout,
                        vop_node=vn, vop_redge=ve)
    ctx.compute_output_fields(0)
    for kind, key, tol in ((DATA_VORTICITY, "vorticity", 1e-11),
                           (DATA_DIVERGENCE, "divergence", 1e-11),
                           (DATA_TEMPERATURE, "temperature", 1e-12)):
        got = ctx.interpolate(0, kind, -1, ipatch, ea, eb, ca, cb, alpha, beta, nout,
                              vop_node=vn, vop_redge=ve)
        ref = d["raw." + key]
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= tol * np.abs(ref).max(), key
    ctx.close()
