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
