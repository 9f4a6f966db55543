"""Drop-in boundary: the reference's own driver flow (Model::Go, checksum
output manager, error norms) run with the B200 plugins of integration/
(C++ shells deriving from the reference's HorizontalDynamics /
VerticalDynamics / TimestepScheme) must print the reference's checksums.

oracle/_ref/b200_driver is built in the container that has /root/reference
(`make -C oracle`); the GPU box runs the prebuilt binary."""
import json
import os
import re
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "b200_driver")


def reference_spread(key):
    """The reference's own spread under 1e-15 perturbations of its input
    (tests/golden/sensitivity.json, written by tests/make_sensitivity.py)."""
    with open(os.path.join(HERE, "golden", "sensitivity.json")) as f:
        return json.load(f)[key]


def run(mode, *flags, env=None, full=False):
    res = subprocess.run([DRIVER, "--b200", mode, "--output_none"] + list(flags),
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, timeout=900, env=dict(os.environ, **(env or {})))
    assert res.returncode == 0, res.stdout[-3000:]
    if full:
        return res.stdout
    sums = {}
    final = res.stdout[res.stdout.rindex("(Final)"):]
    for m in re.finditer(r"Checksum \((\w+)\): ([-+0-9.eE]+)", final):
        sums[m.group(1)] = float(m.group(2))
    norms = re.findall(r"^\s+(\w+)\s+([0-9.e+-]+)\s+([0-9.e+-]+)\s+([0-9.e+-]+)\s*$",
                       res.stdout, flags=re.M)
    return sums, norms


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="oracle/_ref/b200_driver not built")
def test_driver_reference_mode_runs():
    sums, norms = run("none", "--case", "sw2", "--resolution", "4", "--levels", "1")
    assert set(sums) == {"U", "V", "H"} and len(norms) == 3


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["plugins", "scheme"])
def test_shallow_water_dropin(cuda_library, mode):
    """Williamson 2, ne=8, 3 steps: checksums and error norms of the reference."""
    assert os.path.exists(DRIVER), "oracle/_ref/b200_driver missing"
    flags = ["--case", "sw2", "--resolution", "8", "--levels", "1", "--endtime", "600s"]
    ref, rn = run("none", *flags)
    got, gn = run(mode, *flags)
    scale = max(abs(v) for v in ref.values())
    for k in ref:
        # V sums to rounding noise of U-sized terms: tolerance relative to the largest sum
        assert abs(got[k] - ref[k]) <= 1e-12 * scale, (k, got, ref)
    assert [r[0] for r in rn] == [g[0] for g in gn]
    for r, g in zip(rn, gn):
        for a, b in zip(r[1:], g[1:]):
            assert abs(float(a) - float(b)) <= 1e-4 * float(a) + 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("mode,scheme", [("plugins", "strang"), ("scheme", "strang"),
                                         ("scheme", "ars343")])
def test_nonhydro_dropin(cuda_library, mode, scheme):
    """JW baroclinic wave ne=8 L10, 3 steps.  rho and rho-theta (conserved)
    must match the reference to rounding; U, V, W only to the reference's own
    last-bit sensitivity on this case."""
    assert os.path.exists(DRIVER), "oracle/_ref/b200_driver missing"
    flags = ["--case", "jw", "--resolution", "8", "--levels", "10", "--dt", "200s",
             "--endtime", "600s", "--timescheme", scheme]
    ref, _ = run("none", *flags)
    got, _ = run(mode, *flags)
    for k in ("Rho", "RhoTheta"):
        assert abs(got[k] - ref[k]) <= 1e-12 * abs(ref[k]), (k, got, ref)
    # U, W: 10 x the spread the reference shows against itself (pinned)
    sp = reference_spread("jw_ne8_l10_%s_3steps" % scheme)["rel_spread"]
    assert abs(got["U"] - ref["U"]) <= 10.0 * sp[0] * abs(ref["U"]), (got, ref, sp)
    assert abs(got["W"] - ref["W"]) <= 10.0 * sp[3] * abs(ref["W"]), (got, ref, sp)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["plugins", "scheme"])
def test_cartesian_bubble_dropin(cuda_library, mode):
    """Config 2: rising thermal bubble on the periodic Cartesian x-z slice
    (resx = 36, 72 levels, dt = 0.01 s, 20 steps, --nohypervis as in SURVEY 8c)
    through the reference's own driver flow."""
    assert os.path.exists(DRIVER), "oracle/_ref/b200_driver missing"
    flags = ["--case", "bubble", "--resolution", "36", "--resy", "1", "--levels", "72",
             "--dt", "10000u", "--endtime", "200000u", "--nohypervis"]
    ref, _ = run("none", *flags)
    got, _ = run(mode, *flags)
    for k in ("Rho", "RhoTheta"):
        assert abs(got[k] - ref[k]) <= 1e-12 * abs(ref[k]), (k, got, ref)
    # w grows from rest: the columns away from the bubble hold rounding noise
    # whose sign the reference's Jacobian depends on
    sp = reference_spread("bubble_r36_l72_20steps")["rel_spread"]
    assert abs(got["W"] - ref["W"]) <= max(10.0 * sp[3], 1e-12) * abs(ref["W"]), (got, ref, sp)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["plugins", "scheme"])
def test_uniform_diffusion_dropin(cuda_library, mode):
    """The bubble with uniform diffusion (--diffs 300 --diffv 150, the coefficients
    of the reference's other Cartesian cases) through the driver flow: the shells
    hand Grid::HasUniformDiffusion, the coefficients and the reference state to the
    device."""
    from conftest import added_after_the_gpu_budget
    added_after_the_gpu_budget(cuda_library)
    assert os.path.exists(DRIVER), "oracle/_ref/b200_driver missing"
    flags = ["--case", "bubble", "--resolution", "12", "--resy", "1", "--levels", "24",
             "--dt", "10000u", "--endtime", "100000u", "--nohypervis",
             "--diffs", "300", "--diffv", "150"]
    ref, _ = run("none", *flags)
    got, _ = run(mode, *flags)
    for k in ("Rho", "RhoTheta"):
        assert abs(got[k] - ref[k]) <= 1e-12 * abs(ref[k]), (k, got, ref)
    assert abs(got["W"] - ref["W"]) <= 1e-9 * abs(ref["W"]), (got, ref)
    # and the diffusion is not in the noise of that comparison
    plain, _ = run("none", *[f for f in flags[:-4]])
    assert abs(plain["W"] - ref["W"]) > 1e-6 * abs(ref["W"]), (plain, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("vdisc", ["FE", "FV"])
def test_vertical_order_two_dropin(cuda_library, vdisc):
    """--vertorder 2 with the finite-element and the finite-volume column operators
    through the driver flow (the shells pass Grid::GetVerticalOrder and
    GetVerticalDiscretization on): conserved sums of the reference."""
    from conftest import added_after_the_gpu_budget
    added_after_the_gpu_budget(cuda_library)
    assert os.path.exists(DRIVER), "oracle/_ref/b200_driver missing"
    flags = ["--case", "jw", "--resolution", "4", "--levels", "12", "--vertorder", "2",
             "--vdisc", vdisc, "--dt", "200s", "--endtime", "600s"]
    ref, _ = run("none", *flags)
    got, _ = run("scheme", *flags)
    for k in ("Rho", "RhoTheta"):
        assert abs(got[k] - ref[k]) <= 1e-12 * abs(ref[k]), (k, got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("case,order", [("sw2", 3), ("jw", 5)])
def test_horizontal_order_dropin(cuda_library, case, order):
    """--order 3 (shallow water) and 5 (JW) through the driver flow: the shells
    pass Grid::GetHorizontalOrder on and the general kernels of that order run."""
    from conftest import added_after_the_gpu_budget
    added_after_the_gpu_budget(cuda_library)
    assert os.path.exists(DRIVER), "oracle/_ref/b200_driver missing"
    flags = ["--case", case, "--resolution", "4", "--order", str(order), "--endtime", "600s"]
    flags += ["--levels", "1"] if case == "sw2" else ["--levels", "10", "--dt", "200s"]
    ref, _ = run("none", *flags)
    got, _ = run("scheme", *flags)
    for k in (("H",) if case == "sw2" else ("Rho", "RhoTheta")):
        assert abs(got[k] - ref[k]) <= 1e-12 * abs(ref[k]), (k, got, ref)
    if case == "sw2":
        assert abs(got["U"] - ref["U"]) <= 1e-12 * abs(ref["U"]), (got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,scheme", [("plugins", "ark232"), ("scheme", "ark232"),
                                         ("scheme", "gark2"), ("scheme", "ssp3_332")])
def test_more_schemes_dropin(cuda_library, mode, scheme):
    """GARK2, SSP3(3,3,2) and ARK232 through the driver flow; under the reference's
    own ARK232 the vertical shell also serves StepImplicitTermsExplicitly."""
    from conftest import added_after_the_gpu_budget
    added_after_the_gpu_budget(cuda_library)
    assert os.path.exists(DRIVER), "oracle/_ref/b200_driver missing"
    flags = ["--case", "jw", "--resolution", "4", "--levels", "10", "--dt", "200s",
             "--endtime", "400s", "--timescheme", scheme]
    ref, _ = run("none", *flags)
    got, _ = run(mode, *flags)
    for k in ("Rho", "RhoTheta"):
        assert abs(got[k] - ref[k]) <= 1e-12 * abs(ref[k]), (k, got, ref)


@pytest.mark.gpu
def test_lazy_instance0_residency(cuda_library):
    """TimestepSchemeB200 keeps instance 0 on the device between steps unless an
    output manager fires (SURVEY 8b call-order contract, Model.cpp:477-509): a
    run that only comes back to the host at the end, one that comes back for an
    output every second step and the eager run (both ways every step) print the
    same checksums."""
    assert os.path.exists(DRIVER), "oracle/_ref/b200_driver missing"
    flags = ["--case", "jw", "--resolution", "4", "--levels", "10", "--dt", "200s",
             "--endtime", "1200s"]
    eager, _ = run("scheme", "--b200eager", "1", *flags)
    lazy, _ = run("scheme", *flags)
    lazy_out, _ = run("scheme", "--outputtime", "400s", *flags)
    for k in eager:
        assert lazy[k] == eager[k], (k, lazy, eager)
        assert lazy_out[k] == eager[k], (k, lazy_out, eager)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["plugins", "scheme"])
def test_function_timer_groups(cuda_library, mode):
    """With TB200_TIMING=1 the shells open the reference's FunctionTimer groups
    around the device calls, so that Model::Go's end-of-run report
    (Model.cpp:640-688) is filled in as under the reference's own plugins."""
    assert os.path.exists(DRIVER), "oracle/_ref/b200_driver missing"
    flags = ["--case", "jw", "--resolution", "4", "--levels", "10", "--dt", "200s",
             "--endtime", "600s"]
    out = run(mode, *flags, env={"TB200_TIMING": "1"}, full=True)
    ref = run("none", *flags, full=True)
    counts = {}
    for text, store in ((out, counts), (ref, {})):
        for m in re.finditer(r"Time \[(\w+)\]: \d+ \[\d+, \d+\] \((\d+)\)", text):
            store[m.group(1)] = int(m.group(2))
        if store is not counts:
            refcounts = store
    for group in ("SNHP", "VSIm", "SaSc"):
        assert counts.get(group, 0) == refcounts[group], (group, counts, refcounts)
    assert counts.get("Comm", 0) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode,outputtime", [("plugins", None), ("scheme", None),
                                             ("scheme", "400s")])
def test_held_suarez_dropin(cuda_library, mode, outputtime):
    """Held-Suarez forcing as a workflow process: the reference's
    HeldSuarezPhysics on the host arrays under --b200 none against
    HeldSuarezPhysicsB200, which applies it to instance 0 on the device (under
    TimestepSchemeB200 without any bus traffic, also when an output manager reads
    the host copy every second step).  JW ne=4 L10, 6 steps of 200 s; the forcing
    changes rho-theta, so that checksum is no longer a conserved number: it is
    compared with the reference run like the others."""
    assert os.path.exists(DRIVER), "oracle/_ref/b200_driver missing"
    flags = ["--case", "jw", "--resolution", "4", "--levels", "10", "--dt", "200s",
             "--endtime", "1200s", "--heldsuarez", "1"]
    if outputtime is not None:
        flags += ["--outputtime", outputtime]
    ref, _ = run("none", *flags)
    plain, _ = run("none", *[f for f in flags if f not in ("--heldsuarez", "1")], "--heldsuarez", "0")
    got, _ = run(mode, *flags)
    # the forcing is seen at all (friction + relaxation change the sums) ...
    assert abs(ref["RhoTheta"] - plain["RhoTheta"]) > 1e-9 * abs(plain["RhoTheta"])
    # ... and the device applies the same one
    assert abs(got["Rho"] - ref["Rho"]) <= 1e-12 * abs(ref["Rho"]), (got, ref)
    assert abs(got["RhoTheta"] - ref["RhoTheta"]) <= 1e-11 * abs(ref["RhoTheta"]), (got, ref)
    change = abs(ref["U"] - plain["U"])
    assert abs(got["U"] - ref["U"]) <= 1e-3 * change + 1e-6 * abs(ref["U"]), (got, ref, plain)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["plugins", "scheme"])
def test_explicit_vertical_dropin(cuda_library, mode):
    """--explicitvertical through the reference's driver flow: the
    vertical plugin advances rho theta, w and rho explicitly, StepImplicit is a
    no-op.  JW ne=4 L10, dt = 1 s, 5 steps."""
    assert os.path.exists(DRIVER), "oracle/_ref/b200_driver missing"
    flags = ["--case", "jw", "--resolution", "4", "--levels", "10", "--dt", "1s",
             "--endtime", "5s", "--explicitvertical"]
    ref, _ = run("none", *flags)
    got, _ = run(mode, *flags)
    for k in ("Rho", "RhoTheta"):
        assert abs(got[k] - ref[k]) <= 1e-12 * abs(ref[k]), (k, got, ref)
    assert abs(got["U"] - ref["U"]) <= 1e-10 * abs(ref["U"]), (got, ref)
