"""Build a DeviceContext from an oracle dump (tests only): the device gets the
reference's own geometry, tables and column operators bit for bit, so that
differences measured afterwards are differences of the kernels alone."""
import os

import numpy as np

import refdump
from tempestmodel_b200 import DeviceContext, cartesian as cartgrid, cubedsphere
from tempestmodel_b200._lib import OP_NAMES

EMU_LIBRARY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu",
                           "libtb200_emu.so")


def S(d, k):
    return refdump.scalar(d, k)


def context_from_dump(d, library=None, ninstances=None, owners=None, rank=0,
                      nranks=1, exchange=None, analytic_metric=True, lean=False,
                      fully_explicit=0):
    npatch = S(d, "grid.npatch")
    np_ = S(d, "grid.np")
    nlev = S(d, "grid.nlev")
    ncomp = S(d, "grid.ncomp")
    eqn = S(d, "grid.eqntype")
    hv = 0 if S(d, "run.nohypervis") else S(d, "run.hypervisorder")
    cfg = dict(
        np=np_, nlev=nlev, vertical_order=S(d, "grid.vertorder"), ncomp=ncomp,
        ntracers=S(d, "grid.ntracers"),
        ninstances=ninstances or S(d, "grid.ninstances"),
        eqn_type={1: 1, 2: 2}.get(eqn, eqn),
        cartesian_xz=S(d, "grid.xz"),
        comp_on_redge=list(d["grid.varloc"]) + [0] * (8 - ncomp),
        device=-1,
        g=S(d, "phys.g"), R=S(d, "phys.R"), cp=S(d, "phys.cp"), cv=S(d, "phys.cv"),
        p0=S(d, "phys.p0"), omega=S(d, "phys.omega"),
        earth_radius=S(d, "phys.radius"), ztop=S(d, "grid.ztop"),
        ref_length=S(d, "grid.reflength"),
        hypervis_order=hv,
        nu_scalar=0.0 if S(d, "run.nohypervis") else S(d, "run.nu_scalar"),
        nu_div=0.0 if S(d, "run.nohypervis") else S(d, "run.nu_div"),
        nu_vort=0.0 if S(d, "run.nohypervis") else S(d, "run.nu_vort"),
        fully_explicit=fully_explicit, off_centering=0.0,
    )
    ctx = DeviceContext(library=library, **cfg)
    if nranks > 1:
        ctx.set_exchange(rank, nranks, exchange)
    owners = owners or [0] * npatch
    ne = None
    for n in range(npatch):
        p = "patch%d." % n
        ctx.add_patch(S(d, p + "index"), S(d, p + "panel"), S(d, p + "nelem_a"),
                      S(d, p + "nelem_b"), S(d, p + "halo"), S(d, p + "delta_a"),
                      S(d, p + "delta_b"), owners[n])
    ctx.commit_layout()
    ctx.set_tables(d["table.dxbasis1d"], d["table.stiffness1d"],
                   d["table.gllweights1d"])
    if nlev > 1:
        for i, name in enumerate(OP_NAMES):
            key = "op.%s.coeff" % name
            if key in d:
                ctx.set_column_op(i, d[key], d["op.%s.begin" % name],
                                  d["op.%s.end" % name])
    cartesian = S(d, "grid.iscartesian")
    if not cartesian:
        # elements per panel edge
        ne = int(round((np.pi / 2) / S(d, "patch0.delta_a")))
    for n in range(npatch):
        p = "patch%d." % n
        idx = S(d, p + "index")
        if owners[n] == rank:
            geo = dict(
                jacobian2d=d[p + "jacobian2d"],
                contrametric2da=d[p + "contrametric2da"],
                contrametric2db=d[p + "contrametric2db"],
                coriolis=d[p + "coriolis"], topography=d[p + "topography"],
                jacobian=d[p + "jacobian"], jacobian_redge=d[p + "jacobianredge"])
            if lean:
                # lean geometry: no 3-D metric arrays at all
                del geo["jacobian"], geo["jacobian_redge"]
            if eqn == 2 and not lean:
                geo.update(
                    contrametrica=d[p + "contrametrica"],
                    contrametricb=d[p + "contrametricb"],
                    contrametricxi=d[p + "contrametricxi"],
                    contrametrica_redge=d[p + "contrametricaredge"],
                    contrametricb_redge=d[p + "contrametricbredge"],
                    contrametricxi_redge=d[p + "contrametricxiredge"],
                    derivr_node=d[p + "derivrnode"], derivr_redge=d[p + "derivrredge"])
            ctx.upload_geometry(idx, **geo)
            if (p + "rayleighnode") in d and (np.abs(d[p + "rayleighnode"]).max() > 0.0
                                              or np.abs(d[p + "rayleighredge"]).max() > 0.0):
                ctx.upload_rayleigh(idx, d[p + "rayleighnode"], d[p + "rayleighredge"],
                                    d[p + "refstatenode"], d[p + "refstateredge"])
            if "grid.diffs" in d and (float(np.ravel(d["grid.diffs"])[0]) != 0.0 or float(np.ravel(d["grid.diffv"])[0]) != 0.0):
                ctx.upload_reference_state(idx, d[p + "refstatenode"], d[p + "refstateredge"])
            if eqn == 2 or True:
                if (p + "elementareanode") in d:
                    ctx.upload_element_area(idx, d[p + "elementareanode"],
                                            d[p + "elementarearedge"])
        if not cartesian:
            h = S(d, p + "halo")
            ea0 = S(d, p + "a_global_begin") // np_
            eb0 = S(d, p + "b_global_begin") // np_
            nea, neb = S(d, p + "nelem_a"), S(d, p + "nelem_b")
            ctx.set_node_ids(idx, cubedsphere.node_ids(
                S(d, p + "panel"), nea, neb, ea0, eb0, ne, np_))
            if owners[n] == rank:
                an = d[p + "anode"][h:-h]
                bn = d[p + "bnode"][h:-h]
                ia, ib, sp, m = cubedsphere.seam_transforms(
                    S(d, p + "panel"), nea, neb, ea0, eb0, ne, np_, an, bn)
                ctx.set_seam_transforms(idx, ia, ib, sp, m)
    if cartesian:
        # one row of patches along alpha (GridCartesianGLL.cpp:146-229)
        ne_a = sum(S(d, "patch%d.nelem_a" % n) for n in range(npatch))
        ne_b = S(d, "patch0.nelem_b")
        for n in range(npatch):
            p = "patch%d." % n
            ctx.set_node_ids(S(d, p + "index"), cartgrid.node_ids(
                S(d, p + "nelem_a"), S(d, p + "nelem_b"),
                S(d, p + "a_global_begin") // np_, S(d, p + "b_global_begin") // np_,
                ne_a, ne_b, np_))
    if (analytic_metric and eqn == 2 and ("patch0.xnode" in d or cartesian)):
        for n in range(npatch):
            if owners[n] == rank:
                p = "patch%d." % n
                xn = d[p + "xnode"] if not cartesian else d[p + "anode"]
                yn = d[p + "ynode"] if not cartesian else d[p + "bnode"]
                ctx.set_terrain_metric(S(d, p + "index"), xn, yn,
                                       d[p + "topographyderiv"])
        ctx.set_vertical_coordinate(d["grid.retalevels"], d["grid.retainterfaces"])
    ctx.build_connectivity()
    if "grid.massfluxlevels" in d and int(np.ravel(d["grid.massfluxlevels"])[0]) != 0:
        ctx.set_mass_flux_on_levels(True)
    if "grid.vdisc_fv" in d and int(np.ravel(d["grid.vdisc_fv"])[0]) != 0:
        ctx.set_vertical_discretization(True)
    if "grid.diffs" in d and (float(np.ravel(d["grid.diffs"])[0]) != 0.0 or float(np.ravel(d["grid.diffv"])[0]) != 0.0):
        ctx.set_uniform_diffusion(float(np.ravel(d["grid.diffs"])[0]), float(np.ravel(d["grid.diffv"])[0]))
    ctx.dump = d
    ctx.local_patches = [n for n in range(npatch) if owners[n] == rank]
    return ctx


def upload_tag(ctx, d, tag, instances=None):
    """Upload the state instances recorded under `tag`."""
    ninst = ctx.cfg.ninstances
    for n in ctx.local_patches:
        idx = S(d, "patch%d.index" % n)
        for m in (instances if instances is not None else range(ninst)):
            key = "%s.patch%d.inst%d." % (tag, n, m)
            if key + "node" not in d:
                continue
            ctx.upload_state(idx, m, d[key + "node"], d.get(key + "redge"),
                             d.get(key + "tracers"))


def download(ctx, d, inst):
    """-> {patch n: (node, redge)} in the reference layout (halo zero)."""
    out = {}
    for n in ctx.local_patches:
        idx = S(d, "patch%d.index" % n)
        key = None
        for k in d:
            if k.endswith(".patch%d.inst0.node" % n):
                key = k
                break
        node = np.zeros_like(d[key])
        rk = key.replace("node", "redge")
        redge = np.zeros_like(d[rk]) if rk in d else None
        ctx.download_state(idx, inst, node, redge, None, True)
        out[n] = (node, redge)
    return out


def interior(a, halo=1):
    return a[:, halo:-halo, halo:-halo, :]


def pole_mask(d, n):
    """True where a node is NOT a pole of the sphere (interior nodes)."""
    lat = d["patch%d.lat" % n][1:-1, 1:-1]
    return np.abs(np.abs(lat) - 0.5 * np.pi) > 1e-9


def compare(ctx, d, inst, tag, comps_node, comps_redge=(), ref_inst=None,
            skip_poles=False):
    """max |dev - ref| / max |ref| per component over interior nodes."""
    got = download(ctx, d, inst)
    ref_inst = inst if ref_inst is None else ref_inst
    errs = {}
    for loc, comps in (("node", comps_node), ("redge", comps_redge)):
        for c in comps:
            num, den = 0.0, 0.0
            for n in ctx.local_patches:
                ref = d["%s.patch%d.inst%d.%s" % (tag, n, ref_inst, loc)]
                dev = got[n][0 if loc == "node" else 1]
                r = interior(ref)[c]
                v = interior(dev)[c]
                if skip_poles:
                    m = pole_mask(d, n)
                    r, v = r[m], v[m]
                num = max(num, np.abs(v - r).max())
                den = max(den, np.abs(r).max())
            errs[(loc, c)] = num / den if den > 0 else num
    return errs


def download_tracers(ctx, d, inst):
    """-> {patch n: tracers [nT][W_A][W_B][L]} in the reference layout."""
    out = {}
    for n in ctx.local_patches:
        idx = S(d, "patch%d.index" % n)
        key = [k for k in d if k.endswith(".patch%d.inst0.tracers" % n)][0]
        tr = np.zeros_like(d[key])
        ctx.download_state(idx, inst, None, None, tr, False)
        out[n] = tr
    return out


def compare_tracers(ctx, d, inst, tag, before=None):
    """max |dev - ref| per tracer over interior nodes, relative to max |ref|, or -
    with before = (tag, inst) - to the largest change since that record."""
    got = download_tracers(ctx, d, inst)
    ntr = S(d, "grid.ntracers")
    errs = {}
    for c in range(ntr):
        num = den = 0.0
        for n in ctx.local_patches:
            ref = interior(d["%s.patch%d.inst%d.tracers" % (tag, n, inst)])[c]
            dev = interior(got[n])[c]
            num = max(num, np.abs(dev - ref).max())
            if before is None:
                den = max(den, np.abs(ref).max())
            else:
                b = interior(d["%s.patch%d.inst%d.tracers" % (before[0], n, before[1])])[c]
                den = max(den, np.abs(ref - b).max())
        errs[("tracer", c)] = num / den if den > 0 else num
    return errs
