"""Golden cases: reference runs through oracle/_ref/ref_dump whose outputs are
committed under tests/golden/ (regenerate with `python tests/make_golden.py`
in a container that has /root/reference; the GPU box only reads the files)."""
import os

import numpy as np

import refdump

GOLDEN = refdump.GOLDEN

KGU = "-0.25,1.25,0,0,0"

CASES = {
    # shallow water, Williamson 2 (reduced to ne=2)
    "sw2_ne2": dict(
        case="sw2", flags=["--resolution", "2"],
        script=";".join([
            "dump:ic,0", "copy:0,1", "hexp:0,1,100", "dump:h1,1", "dss:1", "dump:dss,1",
            "copy:1,4", "hasc:4,1,2,200", "dump:hasc,1,2",
            "lincomb:3,0.25,1.5,0,0.5,-1", "dump:lc,3",
            "copy:0,1", "copy:0,2", "copy:0,3", "copy:0,4",
            "step:2", "dump:st,0,1", "checksum:cs"])),
    # nonhydrostatic, Jablonowski-Williamson (config 3, reduced).  "addw" adds a
    # smooth non-zero w so that no column has exactly zero wind: the reference's
    # implicit Jacobian carries sign(xi-dot), which is rounding noise on the JW
    # equator and poles and makes those columns chaotic for the reference itself.
    "jw_ne2_l6": dict(
        case="jw", flags=["--resolution", "2", "--levels", "6"],
        script=";".join([
            "addw:0,20000", "dss:0",
            "dump:ic,0", "copy:0,1", "hexp:0,1,50", "dump:h1,1", "vexp:0,1,50",
            "dump:v1,1", "dss:1", "dump:dss,1", "copy:1,2", "vimp:2,2,30",
            "dump:vi,2", "hasc:1,3,4,200", "dump:hasc,3,4"])),
    "jw_ne2_l6_strang": dict(
        case="jw", flags=["--resolution", "2", "--levels", "6", "--dt", "200s"],
        script="addw:0,20000;dss:0;dump:ic,0;step:2;dump:st,0,1;checksum:cs"),
    "jw_ne2_l6_ars343": dict(
        case="jw", flags=["--resolution", "2", "--levels", "6", "--dt", "200s",
                          "--timescheme", "ars343"],
        script="addw:0,20000;dss:0;dump:ic,0;step:2;dump:st,0;checksum:cs"),
}

# nonhydrostatic Cartesian x-z slice, rising thermal bubble (SURVEY 8c config 2,
# reduced): periodic GridCartesianGLL, one element across y
CASES["bubble_r6_l8"] = dict(
    case="bubble", npatch=1,
    flags=["--resolution", "6", "--resy", "1", "--levels", "8", "--dt", "10000u",
           "--nu", "1e4", "--nud", "1e4", "--nuv", "1e4"],
    script=";".join([
        "addw:0,100", "dss:0",
        "dump:ic,0", "copy:0,1", "hexp:0,1,0.01", "dump:h1,1", "vexp:0,1,0.01",
        "dump:v1,1", "dss:1", "dump:dss,1", "copy:1,2", "vimp:2,2,0.01",
        "dump:vi,2", "hasc:1,3,4,0.01", "dump:hasc,3,4"]))
CASES["bubble_r6_l8_strang"] = dict(
    case="bubble", npatch=1,
    flags=["--resolution", "6", "--resy", "1", "--levels", "8", "--dt", "10000u",
           "--nu", "1e4", "--nud", "1e4", "--nuv", "1e4"],
    script="addw:0,100;dss:0;dump:ic,0;step:3;dump:st,0;checksum:cs",
    geometry_from="bubble_r6_l8")

# uniform diffusion (TestCase::GetUniformDiffusionCoeffs -> Grid::HasUniformDiffusion;
# the Cartesian cases of test/nonhydro_xz use it): the bubble and the JW case with
# distinct scalar and vector coefficients, stage by stage and over Strang steps
_STAGES_DIFF = [
    "dump:ic,0", "copy:0,1", "hexp:0,1,%(dt)s", "dump:h1,1", "vexp:0,1,%(dt)s",
    "dump:v1,1", "dss:1", "dump:dss,1", "copy:1,2", "vimp:2,2,%(dt)s", "dump:vi,2",
    "copy:0,1", "copy:0,2", "copy:0,3", "copy:0,4", "step:2", "dump:st,0", "checksum:cs"]
CASES["bubble_r6_l8_diff"] = dict(
    case="bubble", npatch=1,
    flags=["--resolution", "6", "--resy", "1", "--levels", "8", "--dt", "10000u",
           "--nu", "1e4", "--nud", "1e4", "--nuv", "1e4", "--diffs", "300", "--diffv", "150"],
    script=";".join(["addw:0,100", "dss:0"] + [s % dict(dt="0.01") for s in _STAGES_DIFF]),
    compact=True, keep=("refstatenode", "refstateredge"))
CASES["jw_ne2_l6_diff"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "6", "--dt", "200s",
                      "--diffs", "2e5", "--diffv", "1e5"],
    script=";".join(["addw:0,20000", "dss:0"] + [s % dict(dt="50") for s in _STAGES_DIFF]),
    compact=True, keep=("refstatenode", "refstateredge"))

# tracers: the JW case carrying three analytic tracer densities (oracle/ref_dump.cpp
# JWTracerTest): horizontal transport, implicit column transport, both
# positive-definite filters, DSS and hyperdiffusion of tracers
CASES["jwtr_ne2_l6"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "6", "--ntracers", "3"],
    script=";".join([
        "addw:0,20000", "dss:0",
        "dump:ic,0", "copy:0,1", "hexp:0,1,50", "dump:h1,1", "vexp:0,1,50",
        "dump:v1,1", "dss:1", "dump:dss,1", "copy:1,2", "vimp:2,2,30",
        "dump:vi,2", "hasc:1,3,4,200", "dump:hasc,3,4"]),
    geometry_from="jw_ne2_l6")
CASES["jwtr_ne2_l6_strang"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "6", "--dt", "200s", "--ntracers", "3"],
    script="addw:0,20000;dss:0;dump:ic,0;step:3;dump:st,0;checksum:cs",
    geometry_from="jw_ne2_l6_strang")

CASES["jwtr_ne2_l6_ars343"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "6", "--dt", "200s", "--ntracers", "3",
                      "--timescheme", "ars343"],
    script="addw:0,20000;dss:0;dump:ic,0;step:1;dump:s1,0,3,4;step:1;dump:st,0",
    geometry_from="jw_ne2_l6_strang")

# Rayleigh friction: the JW case with a sponge layer (oracle/ref_dump.cpp --rayleigh)
CASES["jwray_ne2_l6"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "6", "--dt", "200s", "--rayleigh", "0.01"],
    script=";".join(["addw:0,20000", "dss:0", "dump:ic,0", "hasc:0,1,2,200", "dump:hasc,1",
                     "copy:0,1", "copy:0,2", "copy:0,3", "copy:0,4",
                     "step:2", "dump:st,0"]))

# second-order viscosity (--hypervisorder 2: one Laplacian application + DSS,
# HorizontalDynamicsFEM.cpp:2671-2684; nu is not scaled with the resolution)
CASES["jwhv2_ne2_l6"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "6", "--dt", "200s",
                      "--hypervisorder", "2", "--nu", "1.0e7", "--nud", "2.0e7",
                      "--nuv", "0.5e7"],
    script=";".join(["addw:0,20000", "dss:0", "dump:ic,0", "hasc:0,1,2,200", "dump:hasc,1",
                     "copy:0,1", "copy:0,2", "copy:0,3", "copy:0,4",
                     "step:2", "dump:st,0"]),
    geometry_from="jw_ne2_l6_strang")

# stretched levels (--vstretch cubic): non-uniform vertical operator tables and
# a level-dependent layer depth (the column-constant fast path must step aside)
CASES["jw_ne2_l6_cubic"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "6", "--dt", "200s", "--vstretch", "cubic"],
    script=";".join([
        "addw:0,20000", "dss:0",
        "dump:ic,0", "copy:0,1", "hexp:0,1,50", "dump:h1,1", "vexp:0,1,50",
        "dump:v1,1", "dss:1", "dump:dss,1", "copy:1,2", "vimp:2,2,30",
        "dump:vi,2", "copy:0,1", "copy:0,2", "copy:0,3", "copy:0,4",
        "step:2", "dump:st,0"]))

# fourth-order hyperdiffusion with three different coefficients (scalar,
# divergence, vorticity): the defaults are all 1e15 and would hide a mix-up
CASES["jwhv4_ne2_l6"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "6", "--dt", "200s",
                      "--nu", "1.0e15", "--nud", "2.5e15", "--nuv", "0.4e15"],
    script=";".join(["addw:0,20000", "dss:0", "dump:ic,0", "hasc:0,1,2,200", "dump:hasc,1",
                     "copy:0,1", "copy:0,2", "copy:0,3", "copy:0,4",
                     "step:2", "dump:st,0"]),
    geometry_from="jw_ne2_l6_strang")

# Williamson 2 with the flow axis tilted (--alpha 0.7): Coriolis parameter and
# both velocity components non-trivial on every panel
CASES["sw2_ne2_alpha"] = dict(
    case="sw2", flags=["--resolution", "2", "--alpha", "0.7"],
    script=";".join([
        "dump:ic,0", "copy:0,1", "hexp:0,1,100", "dump:h1,1", "dss:1", "dump:dss,1",
        "copy:0,1", "copy:0,2", "copy:0,3", "copy:0,4",
        "step:2", "dump:st,0", "checksum:cs"]))

# more time schemes on the same grid and initial state: only the run records are
# stored, the geometry comes from the strang case (same flags)
for _scheme in ("ars222", "ars232", "ars443", "strang/ssprk53", "strang/rk4", "strang/rk3"):
    CASES["jw_ne2_l6_" + _scheme.replace("/", "_")] = dict(
        case="jw", flags=["--resolution", "2", "--levels", "6", "--dt", "200s",
                          "--timescheme", _scheme],
        script="addw:0,20000;dss:0;dump:ic,0;step:2;dump:st,0;checksum:cs",
        geometry_from="jw_ne2_l6_strang")
for _scheme in ("erk", "erk/rk4", "erk/rk3", "erk/ssprk53", "erk/fe"):
    CASES["sw2_ne2_" + _scheme.replace("/", "_")] = dict(
        case="sw2", flags=["--resolution", "2", "--timescheme", _scheme],
        script="dump:ic,0;step:2;dump:st,0;checksum:cs",
        geometry_from="sw2_ne2")

# ---- L = 30: the level count every measured number is quoted on (bench.py).
# 120-thread blocks in the pipelined kernels, level tiles of TBF_KB = 32, the
# shared-memory carve-up of tb_pipe_smem_doubles and the 93-unknown band system
# of the column solve are only exercised at this depth.
_STAGES_L30 = ";".join([
    "addw:0,20000", "dss:0",
    "dump:ic,0", "copy:0,1", "hexp:0,1,50", "dump:h1,1", "vexp:0,1,50",
    "dump:v1,1", "dss:1", "dump:dss,1", "copy:1,2", "vimp:2,2,30",
    "dump:vi,2", "hasc:1,3,4,200", "dump:hasc,3,4"])
CASES["jw_ne2_l30"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "30", "--ztop", "30000", "--pert", "Exp"],
    script=_STAGES_L30)
CASES["jw_ne2_l30_strang"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "30", "--ztop", "30000", "--pert", "Exp",
                      "--dt", "200s"],
    script="addw:0,20000;dss:0;dump:ic,0;step:2;dump:st,0,1;checksum:cs",
    geometry_from="jw_ne2_l30")
# ne = 4 on 24 patches (2 x 2 elements each): the decomposition bench.py uses
# on 2, 4 and 8 GPUs, stage by stage and over two Strang steps
CASES["jw_ne4_l30_p24"] = dict(
    case="jw", npatch=24,
    flags=["--resolution", "4", "--levels", "30", "--ztop", "30000", "--pert", "Exp",
           "--dt", "200s"],
    script=";".join([
        "addw:0,20000", "dss:0", "dump:ic,0", "copy:0,1", "hexp:0,1,50", "vexp:0,1,50",
        "dss:1", "dump:dss,1", "copy:1,2", "vimp:2,2,30", "dump:vi,2",
        "hasc:1,3,4,200", "dump:hasc,3",
        "copy:0,1", "copy:0,2", "copy:0,3", "copy:0,4", "step:2", "dump:st,0", "checksum:cs"]),
    compact=True)

# tracers at L = 30 (config 4's level count): stage records and three Strang steps
CASES["jwtr_ne2_l30"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "30", "--ztop", "30000", "--pert", "Exp",
                      "--dt", "200s", "--ntracers", "3"],
    script=";".join([
        "addw:0,20000", "dss:0",
        "dump:ic,0", "copy:0,1", "hexp:0,1,50", "dump:h1,1", "vexp:0,1,50",
        "dss:1", "dump:dss,1", "copy:1,2", "vimp:2,2,30",
        "dump:vi,2", "hasc:1,3,4,200", "dump:hasc,3",
        "copy:0,1", "copy:0,2", "copy:0,3", "copy:0,4", "step:3", "dump:st,0"]),
    geometry_from="jw_ne2_l30", compact=True)

# conservation diagnostics of the reference (Grid::ComputeTotalEnergy,
# ComputeTotalPotentialEnstrophy, ComputeTotalVerticalMomentum) on the state
# before and after two steps; only the scalars are stored, geometry and initial
# state come from the case named in geometry_from (same flags, same script head)
CASES["jw_ne2_l6_energy"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "6", "--dt", "200s"],
    script="addw:0,20000;dss:0;energy:e0,0;step:2;energy:e2,0;checksum:cs",
    geometry_from="jw_ne2_l6_strang", scalars_only=True)
CASES["jw_ne2_l30_energy"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "30", "--ztop", "30000", "--pert", "Exp",
                      "--dt", "200s"],
    script="addw:0,20000;dss:0;energy:e0,0;step:2;energy:e2,0;checksum:cs",
    geometry_from="jw_ne2_l30_strang", scalars_only=True)
CASES["sw2_ne2_energy"] = dict(
    case="sw2", flags=["--resolution", "2", "--alpha", "0.7"],
    script="energy:e0,0;step:2;energy:e2,0;checksum:cs",
    geometry_from="sw2_ne2_alpha", scalars_only=True)

# Held-Suarez forcing as a workflow step: HeldSuarezPhysics::Perform
# on the JW state, two applications with different intervals
CASES["jw_ne2_l30_hs"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "30", "--ztop", "30000", "--pert", "Exp",
                      "--dt", "200s"],
    script="addw:0,20000;dss:0;dump:ic,0;hs:1800;dump:hs1,0;hs:250.5;dump:hs2,0",
    geometry_from="jw_ne2_l30", compact=True, surface_product=True)

# --explicitvertical: VerticalDynamicsFEM::StepExplicit also advances
# rho theta, w, rho with the column tendencies (VerticalDynamicsFEM.cpp:748-793),
# StepImplicit does nothing (:1240-1242); one stage and two Strang steps at a
# time step the explicit vertical allows
CASES["jw_ne2_l6_explicitv"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "6", "--dt", "1s", "--explicitvertical"],
    script=";".join([
        "addw:0,20000", "dss:0", "dump:ic,0", "copy:0,1", "hexp:0,1,1", "vexp:0,1,1",
        "dump:v1,1", "copy:1,2", "vimp:2,2,1", "dump:vi,2",
        "copy:0,1", "copy:0,2", "copy:0,3", "copy:0,4", "step:2", "dump:st,0"]),
    geometry_from="jw_ne2_l6", compact=True)

# the same with tracers: UpdateColumnTracers in its explicit branches
# (VerticalDynamicsFEM.cpp:802-810, 4048-4171)
CASES["jwtr_ne2_l6_explicitv"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "6", "--dt", "1s", "--explicitvertical",
                      "--ntracers", "3"],
    script=";".join([
        "addw:0,20000", "dss:0", "dump:ic,0", "copy:0,1", "hexp:0,1,1", "dump:h1,1",
        "vexp:0,1,1", "dump:v1,1",
        "copy:0,1", "copy:0,2", "copy:0,3", "copy:0,4", "step:2", "dump:st,0"]),
    geometry_from="jw_ne2_l6", compact=True)

# output-side interpolation: Grid::ReduceInterpolate of the state and
# tracers to a latitude-longitude grid and uniform REta levels, with and without
# the conversion to primitive variables
CASES["jwtr_ne2_l6_interp"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "6", "--ntracers", "3"],
    script="addw:0,20000;dss:0;dump:ic,0;interp:raw,12,7,5,0;interp:prim,12,7,5,1",
    geometry_from="jw_ne2_l6", compact=True)

# vertical order > 1: the general kernels (column operators of any
# width, Jacobian band 2 * offd + 1) against the reference at --vertorder 2, 12 levels,
# and --vertorder 4, 24 levels (fewer levels than about six elements make the
# reference's own dgbsv call fail: "Matrix A has insufficient rows for DGBSV")
_STAGES_VO = ";".join([
    "addw:0,20000", "dss:0",
    "dump:ic,0", "copy:0,1", "hexp:0,1,50", "dump:h1,1", "vexp:0,1,50",
    "dump:v1,1", "dss:1", "dump:dss,1", "copy:1,2", "vimp:2,2,30",
    "dump:vi,2", "hasc:1,3,4,200", "dump:hasc,3",
    "copy:0,1", "copy:0,2", "copy:0,3", "copy:0,4", "step:2", "dump:st,0"])
CASES["jw_ne2_l12_vo2"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "12", "--vertorder", "2", "--dt", "200s"],
    script=_STAGES_VO, compact=True)
CASES["jw_ne2_l24_vo3"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "24", "--vertorder", "3", "--dt", "200s"],
    script=_STAGES_VO, compact=True)
CASES["jw_ne2_l24_vo4"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "24", "--vertorder", "4", "--dt", "200s"],
    script=_STAGES_VO, compact=True)

# --vmassfluxlevels: BuildF with the mass and rho-theta fluxes formed on levels
CASES["jw_ne2_l6_mfl"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "6", "--dt", "200s", "--vmassfluxlevels"],
    script=_STAGES_VO, compact=True)

# the remaining IMEX schemes that run with these plugins (TimestepSchemeGARK2,
# SSP3332, ARK232; ARS343b only works with HighSpeedDynamics)
for _s in ("gark2", "ssp3_332", "ark232"):
    CASES["jw_ne2_l6_%s" % _s] = dict(
        case="jw", flags=["--resolution", "2", "--levels", "6", "--dt", "200s",
                          "--timescheme", _s],
        script="addw:0,20000;dss:0;dump:ic,0;step:2;dump:st,0;checksum:cs",
        geometry_from="jw_ne2_l6_strang")

# three-dimensional periodic Cartesian box (GridCartesianGLL with fCartesianXZ =
# false, as ThermalBubbleCartesian3DTest sets its grid up): the bubble on 4 x 3
# elements; the flow stays uniform in y, what is exercised is the connectivity and
# the DSS across y and the v rows
CASES["bubble3d_r4x3_l6"] = dict(
    case="bubble", npatch=1,
    flags=["--resolution", "4", "--resy", "3", "--levels", "6", "--dt", "10000u",
           "--nu", "1e4", "--nud", "1e4", "--nuv", "1e4", "--xz", "0"],
    script=";".join([
        "addw:0,100", "dss:0",
        "dump:ic,0", "copy:0,1", "hexp:0,1,0.01", "dump:h1,1", "vexp:0,1,0.01",
        "dump:v1,1", "dss:1", "dump:dss,1", "copy:1,2", "vimp:2,2,0.01",
        "dump:vi,2", "hasc:1,3,4,0.01", "dump:hasc,3",
        "copy:0,1", "copy:0,2", "copy:0,3", "copy:0,4", "step:2", "dump:st,0"]),
    compact=True)

# shallow-water tracers (oracle/ref_dump.cpp SWTracerTest): transport inside
# StepShallowWater, the element filter, DSS and hyperdiffusion of the tracers
CASES["sw2tr_ne2"] = dict(
    case="sw2", flags=["--resolution", "2", "--alpha", "0.7", "--ntracers", "2"],
    script=";".join([
        "dump:ic,0", "copy:0,1", "hexp:0,1,100", "dump:h1,1", "dss:1", "dump:dss,1",
        "copy:1,4", "hasc:4,1,2,200", "dump:hasc,1",
        "copy:0,1", "copy:0,2", "copy:0,3", "copy:0,4",
        "step:3", "dump:st,0", "checksum:cs"]), compact=True)

# --order 3 and 5 (horizontal order np other than 4: the general kernels are
# templates on np): shallow water and the JW case, stages and two Strang steps
for _np in (3, 5, 6):
    CASES["sw2_ne2_np%d" % _np] = dict(
        case="sw2", flags=["--resolution", "2", "--order", str(_np)],
        script=";".join([
            "dump:ic,0", "energy:e0,0", "copy:0,1", "hexp:0,1,100", "dump:h1,1", "dss:1",
            "dump:dss,1", "copy:1,4", "hasc:4,1,2,200", "dump:hasc,1,2",
            "copy:0,1", "copy:0,2", "copy:0,3", "copy:0,4",
            "step:2", "dump:st,0", "checksum:cs"]), compact=True)
    if _np == 6:
        continue
    CASES["jw_ne2_l6_np%d" % _np] = dict(
        case="jw", flags=["--resolution", "2", "--levels", "6", "--order", str(_np), "--dt", "200s"],
        script=_STAGES_VO, compact=True)

# --vdisc FV (finite-volume column operators; even orders only): order 2, 12 levels
CASES["jw_ne2_l12_fv2"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "12", "--vertorder", "2", "--vdisc", "FV",
                      "--dt", "200s"],
    script=_STAGES_VO, compact=True)

# tracers at vertical order 2: the column transport of the tracers factorises
# a band of half-width 2 * order - 1 (VerticalDynamicsFEM.cpp:4028-4038)
CASES["jwtr_ne2_l12_vo2"] = dict(
    case="jw", flags=["--resolution", "2", "--levels", "12", "--vertorder", "2", "--dt", "200s",
                      "--ntracers", "3"],
    script=";".join([
        "addw:0,20000", "dss:0",
        "dump:ic,0", "copy:0,1", "hexp:0,1,50", "dump:h1,1", "vexp:0,1,50",
        "dss:1", "dump:dss,1", "copy:1,2", "vimp:2,2,30", "dump:vi,2"]),
    geometry_from="jw_ne2_l12_vo2", compact=True)

_SHARED_PREFIXES = ("patch", "op.", "grid.")


def golden_path(name):
    return os.path.join(GOLDEN, name + ".npz")


def load_case(name):
    """Golden dump of a case: the committed file, else a fresh reference run."""
    path = golden_path(name)
    if os.path.exists(path):
        with np.load(path) as z:
            d = {k: z[k] for k in z.files}
        base = CASES.get(name, {}).get("geometry_from")
        if base is not None:
            scalars = CASES[name].get("scalars_only")
            for k, v in load_case(base).items():
                if (k.startswith(_SHARED_PREFIXES) or (scalars and k.startswith("ic."))) \
                        and k not in d:
                    d[k] = v
        return d
    if not refdump.have_ref_dump():
        raise FileNotFoundError("golden file %s missing and oracle/_ref/ref_dump not built" % path)
    c = CASES[name]
    return refdump.run_ref_dump("/tmp/tb200_%s.bin" % name, c["case"], c["script"], c["flags"],
                                npatch=c.get("npatch", 6))


def _compact(d, keep=()):
    """Keep what the device path reads and the tests compare: zero the halo and
    the component slots that are not located at the array (the reference keeps
    every component at both locations, GridPatch.cpp:341-357; only the valid
    ones are uploaded and compared), drop records no test reads."""
    on_edge = [int(v) for v in d["grid.varloc"]]
    out = {}
    for k, v in d.items():
        leaf = k.rsplit(".", 1)[-1]
        if leaf in ("refstatenode", "refstateredge", "zlevels", "zinterfaces",
                    "rayleighnode", "rayleighredge") and leaf not in keep:
            continue
        if ".inst" in k and leaf == "tracers":
            v = v.copy()
            v[:, 0, :, :] = 0.0
            v[:, -1, :, :] = 0.0
            v[:, :, 0, :] = 0.0
            v[:, :, -1, :] = 0.0
        if ".inst" in k and leaf in ("node", "redge"):
            v = v.copy()
            v[:, 0, :, :] = 0.0
            v[:, -1, :, :] = 0.0
            v[:, :, 0, :] = 0.0
            v[:, :, -1, :] = 0.0
            for c in range(v.shape[0]):
                if (on_edge[c] != 0) != (leaf == "redge"):
                    v[c] = 0.0
        out[k] = v
    return out


def write_golden(name):
    c = CASES[name]
    d = refdump.run_ref_dump("/tmp/tb200_%s.bin" % name, c["case"], c["script"], c["flags"],
                             npatch=c.get("npatch", 6))
    if c.get("surface_product"):
        # the slots HeldSuarezPhysics takes its surface pressure from: rho and
        # rho-theta on the lowest interface (invalid location, dropped by _compact)
        n = 0
        while "ic.patch%d.inst0.redge" % n in d:
            e = d["ic.patch%d.inst0.redge" % n]
            d["hs.patch%d.surface_product" % n] = e[4, :, :, 0] * e[2, :, :, 0]
            n += 1
    if c.get("compact"):
        d = _compact(d, keep=c.get("keep", ()))
    if c.get("scalars_only"):
        d = {k: v for k, v in d.items() if v.size <= 16 and not k.startswith(_SHARED_PREFIXES)}
    os.makedirs(GOLDEN, exist_ok=True)
    if c.get("geometry_from") is not None:
        base = load_case(c["geometry_from"])
        for k in [k for k in d if k.startswith(_SHARED_PREFIXES)]:
            if k in base and np.array_equal(d[k], base[k]):
                del d[k]
            else:
                assert d[k].size < 64, "geometry differs from %s: %s" % (c["geometry_from"], k)
    np.savez_compressed(golden_path(name), **d)
    return golden_path(name)
