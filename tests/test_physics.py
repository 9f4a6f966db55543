"""Column physics as device workflow steps: Held-Suarez forcing
against HeldSuarezPhysics::Perform of the unmodified reference
(src/atm/HeldSuarezPhysics.cpp:62-301; fixture jw_ne2_l30_hs, two applications
with different forcing intervals on the JW state)."""
import numpy as np
import pytest

import cases
import dumpctx
from test_parity import BACKENDS, assert_below, tendency_errors


@pytest.fixture(params=BACKENDS)
def library(request):
    if request.param == "emu":
        return request.getfixturevalue("emu_library")
    return request.getfixturevalue("cuda_library")


def test_held_suarez(library):
    d = cases.load_case("jw_ne2_l30_hs")
    ctx = dumpctx.context_from_dump(d, library=library)
    dumpctx.upload_tag(ctx, d, "ic")
    for n in ctx.local_patches:
        idx = dumpctx.S(d, "patch%d.index" % n)
        ctx.upload_held_suarez(idx, d["patch%d.lat" % n], d["hs.patch%d.surface_product" % n])
    # the forcing changes u, v (boundary-layer friction) and rho-theta (relaxation):
    # held to 1e-12 of the largest change of the component (libm of the device
    # against glibc: exp, log, pow, sin, cos agree to an ulp or two)
    ctx.held_suarez(1800.0)
    assert_below(tendency_errors(ctx, d, 0, "hs1", "ic", 0, [0, 1, 2]), 1e-12)
    assert_below(dumpctx.compare(ctx, d, 0, "hs1", [4], [3]), 0.0)     # rho, w untouched
    ctx.held_suarez(250.5)
    assert_below(tendency_errors(ctx, d, 0, "hs2", "hs1", 0, [0, 1, 2]), 1e-12)
    ctx.close()


def test_held_suarez_needs_its_inputs(library):
    d = cases.load_case("jw_ne2_l30_hs")
    ctx = dumpctx.context_from_dump(d, library=library)
    dumpctx.upload_tag(ctx, d, "ic")
    with pytest.raises(Exception, match="Held-Suarez inputs not uploaded"):
        ctx.held_suarez(1800.0)
    ctx.close()


def _moist_tracers(d, n, rho):
    """Test data: a moist column set on the JW state - water vapour near and
    above saturation in the lower troposphere, patches of cloud and rain water."""
    lon = d["patch%d.lon" % n][:, :, None]
    lat = d["patch%d.lat" % n][:, :, None]
    zs = d["patch%d.topography" % n][:, :, None]
    reta = d["grid.retalevels"][None, None, :]
    z = zs + reta * (dumpctx.S(d, "grid.ztop") - zs)
    qv = 0.022 * np.exp(-z / 2600.0) * (1.0 + 0.4 * np.cos(lat) * np.sin(2.0 * lon))
    qc = 1.5e-3 * np.exp(-((z - 2500.0) / 1500.0) ** 2) * (np.cos(lat) ** 2) * (1.0 + np.sin(lon)) / 2
    qr = 0.8e-3 * np.exp(-((z - 1500.0) / 1200.0) ** 2) * (np.sin(lat + 0.3) ** 2)
    # a few dry / negative entries exercise the clipping of KesslerPhysics.cpp:168-183
    qr = np.where(np.abs(lat) > 1.3, -1.0e-6, qr)
    return np.stack([rho * qv, rho * qc, rho * qr]), z


def test_kessler_against_the_c_restatement(library):
    """Kessler warm-rain microphysics (tb200_kessler) on a moist test state
    against oracle/kessler_port.c, the C restatement of the reference's Fortran
    kernel and of KesslerPhysics::Perform.  PARITY UNPINNED: the reference's
    kernel cannot be compiled in this image (no Fortran); this test pins the
    device kernel to the restatement, not to the reference.  Tolerance 1e-10 of
    each field: exp / log / pow of the device library against glibc."""
    import ctypes
    import os
    port = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref",
                        "libkessler_port.so")
    if not os.path.exists(port):
        pytest.skip("oracle/_ref/libkessler_port.so not built (make -C oracle)")
    lib = ctypes.CDLL(port)
    d = cases.load_case("jwtr_ne2_l30")
    ctx = dumpctx.context_from_dump(d, library=library)
    L = dumpctx.S(d, "grid.nlev")
    dt = 200.0
    g, R, cp, p0 = (dumpctx.S(d, "phys." + k) for k in ("g", "R", "cp", "p0"))
    gamma = cp / (cp - R)
    pscale = p0 * (R / p0) ** gamma
    expect = {}
    for n in ctx.local_patches:
        idx = dumpctx.S(d, "patch%d.index" % n)
        node = d["ic.patch%d.inst0.node" % n].copy()
        redge = d["ic.patch%d.inst0.redge" % n]
        tr, z = _moist_tracers(d, n, node[4])
        ctx.upload_state(idx, 0, node, redge, np.ascontiguousarray(tr))
        I = (slice(1, -1), slice(1, -1))
        cols = [np.ascontiguousarray(a[I].reshape(-1, L)) for a in
                (node[2], node[4], tr[0], tr[1], tr[2], np.broadcast_to(z, node[4].shape))]
        ncol = cols[0].shape[0]
        precip = np.zeros(ncol)
        work = np.zeros(7 * L + 3 * L + 8)
        P = ctypes.POINTER(ctypes.c_double)
        lib.kessler_physics_batch(
            *[c.ctypes.data_as(P) for c in cols[:5]], cols[5].ctypes.data_as(P),
            ctypes.c_int(ncol), ctypes.c_int(L), ctypes.c_double(dt), ctypes.c_double(pscale),
            ctypes.c_double(gamma), ctypes.c_double(R), precip.ctypes.data_as(P),
            work.ctypes.data_as(P))
        expect[n] = (cols, precip, node[4][I].shape)
    ctx.kessler(dt)
    state = dumpctx.download(ctx, d, 0)
    tracers = dumpctx.download_tracers(ctx, d, 0)
    changed = 0.0
    for n in ctx.local_patches:
        cols, precip, shape = expect[n]
        dev = dumpctx.interior(state[n][0])
        dtr = dumpctx.interior(np.asarray(tracers[n]))
        ref0 = dumpctx.interior(d["ic.patch%d.inst0.node" % n])
        for name, got, ref in (("rhotheta", dev[2], cols[0]), ("rho", dev[4], cols[1]),
                               ("rqv", dtr[0], cols[2]), ("rqc", dtr[1], cols[3]),
                               ("rqr", dtr[2], cols[4])):
            ref = ref.reshape(shape)
            assert np.abs(got - ref).max() <= 1e-10 * np.abs(ref).max(), (n, name)
        changed = max(changed, np.abs(dev[2] - ref0[2]).max() / np.abs(ref0[2]).max())
        assert np.all(precip >= 0.0) and precip.max() > 0.0
    # the step did something: latent heating changed rho theta
    assert changed > 1e-6
    ctx.close()
