"""Parity at L = 30, the level count every measured number is quoted on.

The kernels behind the bench line (k_nh_stage_pipe, k_hyper_pipe, k_dss_fast,
k_column_fast) take a different shape there than at the L = 6 / 8 of the other
fixtures: 120 active threads of a 128-thread block, one level tile of
TBF_KB = 32 with two idle level slots, 19.3 kB element rows, the shared-memory
carve-up of tb_pipe_smem_doubles, a 93-unknown band system in the column solve.
Fixtures: the unmodified reference at `--levels 30 --ztop 30000 --pert Exp`
(the flags of bench.py), ne = 2 on 6 patches stage by stage and over two Strang
steps, and ne = 4 on the 24 patches bench.py uses on 2, 4 and 8 GPUs.
Tolerances as in test_parity.py (FP64): explicit stages 1e-12 of the largest
tendency, DSS 1e-14 of the field, implicit stage 1e-10 of the largest change,
multi-step states 1e-10 of the field.
"""
import numpy as np
import pytest

import cases
import dumpctx
from test_parity import (BACKENDS, TOL_DSS, TOL_IMPLICIT, TOL_STAGE, TOL_STATE,
                         assert_below, tendency_errors)


@pytest.fixture(params=BACKENDS)
def library(request):
    if request.param == "emu":
        return request.getfixturevalue("emu_library")
    return request.getfixturevalue("cuda_library")


def test_stages_l30(library):
    """h1 / v1 / dss / vi / hasc against the reference, fast path enabled."""
    d = cases.load_case("jw_ne2_l30")
    ctx = dumpctx.context_from_dump(d, library=library)
    enabled, reason, dev = ctx.fast_path()
    assert enabled, reason
    assert dev <= 1e-13
    dumpctx.upload_tag(ctx, d, "ic")
    assert_below(dumpctx.compare(ctx, d, 0, "ic", [0, 1, 2, 4], [3]), 0.0)
    ctx.copy(0, 1)
    ctx.h_step_explicit(0, 1, 50.0)
    assert_below(tendency_errors(ctx, d, 1, "h1", "ic", 0, [0, 1, 2, 4], [3]), TOL_STAGE)
    ctx.v_step_explicit(0, 1, 50.0)
    assert_below(tendency_errors(ctx, d, 1, "v1", "ic", 0, [0, 1, 2, 4], [3]), TOL_STAGE)
    # the fused forms the time schemes launch: k_nh_stage_pipe<true, 0 / 1 / 2>
    a = dumpctx.download(ctx, d, 1)
    for coeff, out in (([1.0, 0.0, 0.0, 0.0], 3),          # base = input (2 S)
                       ([0.0, 0.0, 1.0, 0.0], 3),          # base = another instance (3 S)
                       ([0.75, 0.0, 0.25, 0.0, 0.0], 4)):  # two-term base (4 S)
        ctx.copy(0, 2)
        ctx.hv_step_explicit_combine(coeff, 0, out, 50.0)
        assert_below(tendency_errors(ctx, d, out, "v1", "ic", 0, [0, 1, 2, 4], [3],
                                     ref_inst=1), TOL_STAGE)
        if coeff[0] == 1.0:
            b = dumpctx.download(ctx, d, out)
            for n in ctx.local_patches:
                for loc in (0, 1):
                    ia, ib = dumpctx.interior(a[n][loc]), dumpctx.interior(b[n][loc])
                    for c in ([0, 1, 2, 4] if loc == 0 else [3]):
                        assert np.abs(ia[c] - ib[c]).max() <= 4e-16 * np.abs(ia[c]).max()
    ctx.dss(1)
    assert_below(dumpctx.compare(ctx, d, 1, "dss", [0, 1, 2, 4], [3]), TOL_DSS)
    assert_below(tendency_errors(ctx, d, 1, "dss", "ic", 0, [0, 1, 2, 4], [3],
                                 scale=("v1", 1)), 1e-10)
    ctx.copy(1, 2)
    ctx.v_step_implicit(2, 2, 30.0)
    ctx.check_errors()
    assert_below(tendency_errors(ctx, d, 2, "vi", "dss", 1, [2, 4], [3],
                                 skip_poles=True), TOL_IMPLICIT)
    assert_below(dumpctx.compare(ctx, d, 2, "vi", [0, 1, 2, 4], [3], skip_poles=True), 1e-12)
    ctx.h_step_after_subcycle(1, 3, 4, 200.0)
    assert_below(dumpctx.compare(ctx, d, 3, "hasc", [0, 1, 2, 4], [3]), 1e-13)
    assert_below(dumpctx.compare(ctx, d, 4, "hasc", [0, 1, 2, 4], [3]), 1e-11)
    ctx.close()


def test_general_kernels_l30(library, monkeypatch):
    """The general kernels on the stored 3-D metric at L = 30 (two level chunks
    of 15 in k_nh_explicit) against the same records."""
    d = cases.load_case("jw_ne2_l30")
    monkeypatch.setenv("TB200_STAGE_KERNEL", "generic")
    monkeypatch.setenv("TB200_HYPER_KERNEL", "generic")
    monkeypatch.setenv("TB200_COLUMN_KERNEL", "window")
    ctx = dumpctx.context_from_dump(d, library=library, analytic_metric=False)
    assert not ctx.fast_path()[0]
    dumpctx.upload_tag(ctx, d, "ic")
    ctx.copy(0, 1)
    ctx.h_step_explicit(0, 1, 50.0)
    assert_below(tendency_errors(ctx, d, 1, "h1", "ic", 0, [0, 1, 2, 4], [3]), TOL_STAGE)
    ctx.v_step_explicit(0, 1, 50.0)
    assert_below(tendency_errors(ctx, d, 1, "v1", "ic", 0, [0, 1, 2, 4], [3]), TOL_STAGE)
    ctx.dss(1)
    assert_below(dumpctx.compare(ctx, d, 1, "dss", [0, 1, 2, 4], [3]), TOL_DSS)
    ctx.copy(1, 2)
    ctx.v_step_implicit(2, 2, 30.0)
    ctx.check_errors()
    assert_below(tendency_errors(ctx, d, 2, "vi", "dss", 1, [2, 4], [3],
                                 skip_poles=True), TOL_IMPLICIT)
    ctx.h_step_after_subcycle(1, 3, 4, 200.0)
    assert_below(dumpctx.compare(ctx, d, 3, "hasc", [0, 1, 2, 4], [3]), 1e-13)
    ctx.close()


def test_steps_l30(library):
    """Two Strang / KGU35 steps (the bench scheme) at L = 30: state, increment
    instance and the reference's checksums."""
    d = cases.load_case("jw_ne2_l30_strang")
    ctx = dumpctx.context_from_dump(d, library=library)
    assert ctx.fast_path()[0]
    dumpctx.upload_tag(ctx, d, "ic")
    for m in range(1, ctx.cfg.ninstances):
        ctx.copy(0, m)
    ctx.step("strang", True, False, 200.0)
    ctx.step("strang", False, False, 200.0)
    ctx.check_errors()
    assert_below(dumpctx.compare(ctx, d, 0, "st", [0, 1, 2, 4], [3]), TOL_STATE)
    cs = ctx.checksum(0)
    ref = d["cs.checksum"]
    assert abs(cs[4] - ref[4]) <= 1e-13 * abs(ref[4])
    assert abs(cs[2] - ref[2]) <= 1e-13 * abs(ref[2])
    assert abs(cs[0] - ref[0]) <= 1e-10 * abs(ref[0])
    ctx.close()


def test_24_patches_l30(library):
    """ne = 4 on 24 patches (one rank): the decomposition of the multi-GPU bench
    lines - DSS across patch edges inside a panel, panel seams and cube corners
    shared by three patches - stage records and two Strang steps."""
    d = cases.load_case("jw_ne4_l30_p24")
    ctx = dumpctx.context_from_dump(d, library=library)
    assert ctx.fast_path()[0]
    dumpctx.upload_tag(ctx, d, "ic")
    ctx.hv_step_explicit_combine([1.0, 0.0], 0, 1, 50.0)
    ctx.dss(1)
    assert_below(dumpctx.compare(ctx, d, 1, "dss", [0, 1, 2, 4], [3]), TOL_DSS)
    ctx.copy(1, 2)
    ctx.v_step_implicit(2, 2, 30.0)
    ctx.check_errors()
    assert_below(tendency_errors(ctx, d, 2, "vi", "dss", 1, [2, 4], [3],
                                 skip_poles=True), TOL_IMPLICIT)
    ctx.h_step_after_subcycle(1, 3, 4, 200.0)
    assert_below(dumpctx.compare(ctx, d, 3, "hasc", [0, 1, 2, 4], [3]), 1e-13)
    for m in range(1, ctx.cfg.ninstances):
        ctx.copy(0, m)
    ctx.step("strang", True, False, 200.0)
    ctx.step("strang", False, False, 200.0)
    ctx.check_errors()
    assert_below(dumpctx.compare(ctx, d, 0, "st", [0, 1, 2, 4], [3]), TOL_STATE)
    cs = ctx.checksum(0)
    ref = d["cs.checksum"]
    assert abs(cs[4] - ref[4]) <= 1e-13 * abs(ref[4])
    assert abs(cs[2] - ref[2]) <= 1e-13 * abs(ref[2])
    ctx.close()


def test_lean_geometry_l30(library):
    """A host that uploads the 2-D metric, the topography derivatives and the
    vertical coordinate only (no 3-D metric arrays: the memory layout of the
    ne = 240, L = 60 run): same bits as with the full geometry, and the general
    kernels refuse to run."""
    from tempestmodel_b200 import TempestError
    d = cases.load_case("jw_ne2_l30_strang")
    res = []
    for lean in (False, True):
        ctx = dumpctx.context_from_dump(d, library=library, lean=lean)
        enabled, reason, dev = ctx.fast_path()
        assert enabled, reason
        dumpctx.upload_tag(ctx, d, "ic")
        for m in range(1, ctx.cfg.ninstances):
            ctx.copy(0, m)
        ctx.step("strang", True, False, 200.0)
        ctx.step("strang", False, False, 200.0)
        ctx.check_errors()
        assert_below(dumpctx.compare(ctx, d, 0, "st", [0, 1, 2, 4], [3]), TOL_STATE)
        res.append(dumpctx.download(ctx, d, 0))
        if lean:
            import os
            os.environ["TB200_COLUMN_KERNEL"] = "window"
            try:
                with pytest.raises(TempestError, match="3-D metric arrays were not uploaded"):
                    ctx.v_step_implicit(0, 0, 1.0)
            finally:
                del os.environ["TB200_COLUMN_KERNEL"]
        ctx.close()
    for n in res[0]:
        for loc in (0, 1):
            assert np.array_equal(res[0][n][loc], res[1][n][loc])


@pytest.mark.parametrize("name,strip", [("jw_ne2_l30_strang", None), ("jw_ne4_l30_p24", None),
                                        ("jw_ne4_l30_p24", "1"), ("jw_ne2_l6_strang", None)])
def test_fused_dss_same_bits(library, monkeypatch, name, strip):
    """DSS fused into the stage and hyperdiffusion kernels (in-patch averaging
    groups averaged while the values are in registers / L2, the rest by the
    group kernels) against the separate DSS pass: bit-identical state after two
    Strang steps, for any strip partition."""
    d = cases.load_case(name)
    res = []
    for fused in (False, True):
        monkeypatch.setenv("TB200_DSS_FUSED", "1" if fused else "0")
        if strip is not None:
            monkeypatch.setenv("TB200_STRIP", strip)
        ctx = dumpctx.context_from_dump(d, library=library)
        assert ctx.fast_path()[0]
        assert ctx.fused_group_count > 0
        dumpctx.upload_tag(ctx, d, "ic")
        for m in range(1, ctx.cfg.ninstances):
            ctx.copy(0, m)
        ctx.step("strang", True, False, 200.0)
        ctx.step("strang", False, False, 200.0)
        ctx.check_errors()
        res.append([dumpctx.download(ctx, d, m) for m in (0, 1, 2, 4)])
        ctx.close()
    for a, b in zip(*res):
        for n in a:
            for loc in (0, 1):
                assert np.array_equal(a[n][loc], b[n][loc])


@pytest.mark.parametrize("ne,npatch,strip", [(6, 6, None), (6, 6, "4"), (4, 24, "3")])
def test_fused_dss_same_bits_larger_patches(library, monkeypatch, ne, npatch, strip):
    """The same on patches of 6 x 6 elements (every combination of first / inner /
    last element of a strip and of an alpha-row) through the Python driver: no
    reference data needed, the fused and the separate DSS must agree bit for bit."""
    from tempestmodel_b200 import grid as G
    from tempestmodel_b200 import testcases as TC
    from tempestmodel_b200.model import Model
    res = []
    for fused in (False, True):
        monkeypatch.setenv("TB200_DSS_FUSED", "1" if fused else "0")
        if strip is not None:
            monkeypatch.setenv("TB200_STRIP", strip)
        grid = G.GridCSGLL(ne, 7, npatch=npatch, ztop=30000.0)
        model = Model(grid, TC.BaroclinicWaveJWTest(ztop=30000.0, perturbation="exp"),
                      timescheme="strang", dt=300.0, library=library)
        model.initialize()
        assert model.ctx.fast_path()[0]
        assert model.ctx.fused_group_count > 0
        model.step(2)
        res.append((model.download_state(0), model.download_state(1), model.checksum(0)))
        model.ctx.close()
    for inst in (0, 1):
        for idx in res[0][inst]:
            for loc in (0, 1):
                assert np.array_equal(res[0][inst][idx][loc], res[1][inst][idx][loc])


def test_state_does_not_depend_on_the_patch_decomposition(library):
    """6 patches (one rank) against 24 patches: the averaging groups pair their
    members by position on the panel (alpha pairs first, as GridCSGLL::ApplyDSS
    does), not by patch index, so the explicit stages, the DSS and the
    hyperdiffusion give the same bits on any decomposition, and on this grid the
    whole state after three steps does.  (On larger grids the column solve can
    differ in the last bit: a patch solves one copy of a shared column and copies
    it to the duplicates, as the reference does, and which copy that is depends
    on where the patch boundaries are - tools/decomp_ops.py,
    profiles/r2_decomposition_ops.txt.)"""
    from tempestmodel_b200 import grid as G
    from tempestmodel_b200 import testcases as TC
    from tempestmodel_b200.model import Model
    ne, L = 4, 7
    res = []
    for npatch in (6, 24):
        grid = G.GridCSGLL(ne, L, npatch=npatch, ztop=30000.0)
        model = Model(grid, TC.BaroclinicWaveJWTest(ztop=30000.0, perturbation="exp"),
                      timescheme="strang", dt=300.0, library=library)
        model.initialize()
        model.step(3)
        st = model.download_state(0)
        nodes = {p: np.zeros((5, 4 * ne, 4 * ne, L)) for p in range(6)}
        redges = {p: np.zeros((5, 4 * ne, 4 * ne, L + 1)) for p in range(6)}
        for p in grid.patches:
            node, redge = st[p.index]
            sl = (slice(None), slice(4 * p.ea0, 4 * (p.ea0 + p.nea)),
                  slice(4 * p.eb0, 4 * (p.eb0 + p.neb)))
            nodes[p.panel][sl] = node[:, 1:-1, 1:-1]
            redges[p.panel][sl] = redge[:, 1:-1, 1:-1]
        res.append((nodes, redges))
        model.ctx.close()
    for panel in range(6):
        for c in (0, 1, 2, 4):
            assert np.array_equal(res[0][0][panel][c], res[1][0][panel][c]), (panel, c)
        assert np.array_equal(res[0][1][panel][3], res[1][1][panel][3]), panel


@pytest.mark.parametrize("kernels", ["fast", "generic"])
def test_tracers_l30(library, monkeypatch, kernels):
    """Tracer transport at L = 30 (config 4's level count) on the column-constant
    path (k_tracer_stage, k_tracer_hyper, k_column_tracers with the column
    constants) and on the general kernels (TB200_TRACER_KERNEL=generic):
    horizontal transport with the element filter, DSS, implicit column
    transport with the column filter, hyperdiffusion, three Strang steps."""
    if kernels == "generic":
        monkeypatch.setenv("TB200_TRACER_KERNEL", "generic")
    d = cases.load_case("jwtr_ne2_l30")
    ctx = dumpctx.context_from_dump(d, library=library)
    dumpctx.upload_tag(ctx, d, "ic")
    ctx.copy(0, 1)
    ctx.h_step_explicit(0, 1, 50.0)
    assert_below(tendency_errors(ctx, d, 1, "h1", "ic", 0, [0, 1, 2, 4], [3]), TOL_STAGE)
    assert_below(dumpctx.compare_tracers(ctx, d, 1, "h1", before=("ic", 0)), 1e-11)
    ctx.v_step_explicit(0, 1, 50.0)
    ctx.dss(1)
    assert_below(dumpctx.compare(ctx, d, 1, "dss", [0, 1, 2, 4], [3]), TOL_DSS)
    assert_below(dumpctx.compare_tracers(ctx, d, 1, "dss"), 1e-13)
    ctx.copy(1, 2)
    ctx.v_step_implicit(2, 2, 30.0)
    ctx.check_errors()
    assert_below(tendency_errors(ctx, d, 2, "vi", "dss", 1, [2, 4], [3],
                                 skip_poles=True), TOL_IMPLICIT)
    assert_below(dumpctx.compare_tracers(ctx, d, 2, "vi"), 1e-10)
    ctx.h_step_after_subcycle(1, 3, 4, 200.0)
    assert_below(dumpctx.compare(ctx, d, 3, "hasc", [0, 1, 2, 4], [3]), 1e-13)
    assert_below(dumpctx.compare_tracers(ctx, d, 3, "hasc"), 1e-12)
    for m in range(1, ctx.cfg.ninstances):
        ctx.copy(0, m)
    ctx.step("strang", True, False, 200.0)
    ctx.step("strang", False, False, 200.0)
    ctx.step("strang", False, False, 200.0)
    ctx.check_errors()
    assert_below(dumpctx.compare(ctx, d, 0, "st", [0, 1, 2, 4], [3]), TOL_STATE)
    assert_below(dumpctx.compare_tracers(ctx, d, 0, "st"), 1e-9)
    ctx.close()


def test_tracer_stage_kernels_agree(library, monkeypatch):
    """The pipelined tracer stage kernel (persistent blocks, bulk copies one
    element ahead) against the block-per-element kernel
    (TB200_TRACER_STAGE_KERNEL=plain): same arithmetic, to rounding of the
    compiler's FMA contraction (bit-identical on the emulation)."""
    d = cases.load_case("jwtr_ne2_l30")
    res = []
    for plain in (False, True):
        if plain:
            monkeypatch.setenv("TB200_TRACER_STAGE_KERNEL", "plain")
        ctx = dumpctx.context_from_dump(d, library=library)
        dumpctx.upload_tag(ctx, d, "ic")
        for m in range(1, ctx.cfg.ninstances):
            ctx.copy(0, m)
        ctx.step("strang", True, False, 200.0)
        ctx.step("strang", False, False, 200.0)
        ctx.check_errors()
        res.append(dumpctx.download_tracers(ctx, d, 0))
        ctx.close()
    for n in res[0]:
        a, b = np.asarray(res[0][n]), np.asarray(res[1][n])
        assert np.abs(a - b).max() <= 1e-14 * np.abs(b).max(), n


def test_carry_over_with_fused_tracer_filter(library):
    """tb200_lincomb_v_filter (the start of a Strang step: instance 0 += increment,
    then the column filter of the tracers, with the tracer combination formed
    inside the filter kernel) against the two separate calls: the same bits."""
    d = cases.load_case("jwtr_ne2_l30")
    res = []
    for fused in (False, True):
        ctx = dumpctx.context_from_dump(d, library=library)
        dumpctx.upload_tag(ctx, d, "ic")
        # an "increment" with negative tracer values in instance 1
        ctx.copy(0, 1)
        ctx.lincomb([0.0, -0.37], 1)
        ctx.h_step_explicit(0, 1, 400.0)
        if fused:
            ctx.lincomb_v_filter([1.0, 1.0], 0)
        else:
            ctx.lincomb([1.0, 1.0], 0)
            ctx.v_filter_negative_tracers(0)
        res.append((dumpctx.download(ctx, d, 0), dumpctx.download_tracers(ctx, d, 0)))
        ctx.close()
    for n in res[0][0]:
        assert np.array_equal(res[0][0][n][0], res[1][0][n][0])
        assert np.array_equal(res[0][0][n][1], res[1][0][n][1])
        assert np.array_equal(np.asarray(res[0][1][n]), np.asarray(res[1][1][n]))
