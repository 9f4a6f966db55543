"""Host-side setup mirroring the reference's grid / test-case classes.

Everything here runs once before the step loop (setup-time code, SURVEY 8
rows a-12, a-13): GLL tables, vertical column operators, cubed-sphere patch
layout, metric terms and pointwise initial conditions, evaluated with numpy
and handed to the device library through the C ABI in the reference's host
layout.  The per-timestep path never comes back here.

Class and method names follow the reference:
  GridCSGLL          src/atm/GridCSGLL.{h,cpp}, GridGLL.{h,cpp}, Grid.{h,cpp}
  GridPatchCSGLL     src/atm/GridPatchCSGLL.{h,cpp}
  PhysicalConstants  src/atm/PhysicalConstants.h
"""
import math

import numpy as np

from . import cubedsphere as cs


class PhysicalConstants:
    """Defaults of reference src/atm/PhysicalConstants.h:119-131."""

    def __init__(self):
        self.earth_radius = 6.37122e6
        self.g = 9.80616
        self.omega = 7.29212e-5
        self.alpha = 0.0
        self.R = 287.0
        self.cp = 1004.5
        self.p0 = 100000.0

    @property
    def cv(self):
        return self.cp - self.R

    @property
    def gamma(self):
        return self.cp / (self.cp - self.R)

    @property
    def pressure_scaling(self):
        return self.p0 * (self.R / self.p0) ** self.gamma

    def rho_theta_from_pressure(self, p):
        # PhysicalConstants.h:389-391
        return np.exp(np.log(p / self.pressure_scaling) / self.gamma)


# -- GLL tables (GridGLL::Initialize, reference src/atm/GridGLL.cpp:101-180) ----

def gll_tables(np_):
    """DxBasis1D[s][i] = phi'_s(x_i), Stiffness1D[m][i] = Dx[m][i] w_i / w_m and
    the GLL weights, on the reference element [0, 1]."""
    x, w = cs.gll_points(np_, 0.0, 1.0)
    dx = np.zeros((np_, np_))
    for s in range(np_):
        for i in range(np_):
            if i == s:
                dx[s, i] = sum(1.0 / (x[s] - x[m]) for m in range(np_) if m != s)
            else:
                num = 1.0
                for m in range(np_):
                    if m != s and m != i:
                        num *= (x[i] - x[m])
                den = 1.0
                for m in range(np_):
                    if m != s:
                        den *= (x[s] - x[m])
                dx[s, i] = num / den
    st = np.zeros((np_, np_))
    for m in range(np_):
        for i in range(np_):
            st[m, i] = dx[m, i] * w[i] / w[m]
    return dx, st, w


# -- vertical column operators, vertical order 1, uniform levels ----------------
# (GridGLL::InitializeVerticalCoordinate, reference src/atm/GridGLL.cpp:190-363;
#  LinearColumnInterpFEM / DiffFEM / DiffDiffFEM / DiscPenaltyFEM builders,
#  src/atm/LinearColumnOperatorFEM.cpp).  For vertical order 1 every operator
#  is the 2-3 point stencil written below; higher orders are supplied by the
#  reference's own tables through the C++ shells.

def _lagrange_weights(xs, x):
    w = np.ones(len(xs))
    for a in range(len(xs)):
        for b in range(len(xs)):
            if b != a:
                w[a] *= (x - xs[b]) / (xs[a] - xs[b])
    return w


def _lagrange_dweights(xs, x):
    n = len(xs)
    w = np.zeros(n)
    for a in range(n):
        for m in range(n):
            if m == a:
                continue
            t = 1.0 / (xs[a] - xs[m])
            for b in range(n):
                if b != a and b != m:
                    t *= (x - xs[b]) / (xs[a] - xs[b])
            w[a] += t
    return w


def _lagrange_ddweights(xs, x):
    """second derivative of the Lagrange basis through 3 points"""
    assert len(xs) == 3
    w = np.zeros(3)
    for a in range(3):
        den = 1.0
        for b in range(3):
            if b != a:
                den *= (xs[a] - xs[b])
        w[a] = 2.0 / den
    return w


def column_operators(nlev):
    """-> {name: (coeff[nout][nin], begin[nout], end[nout])} for uniform levels,
    vertical order 1."""
    L = nlev
    zn = (np.arange(L) + 0.5) / L           # levels
    ze = np.arange(L + 1) / float(L)        # interfaces
    ops = {}

    def new(nout, nin):
        return np.zeros((nout, nin)), np.zeros(nout, np.int32), np.zeros(nout, np.int32)

    # interpolation levels -> interfaces: linear through the two nearest levels
    c, b, e = new(L + 1, L)
    for k in range(L + 1):
        lo = min(max(k - 1, 0), L - 2)
        c[k, lo:lo + 2] = _lagrange_weights(zn[lo:lo + 2], ze[k])
        b[k], e[k] = lo, lo + 2
    ops["interp_n2e"] = (c, b, e)
    # interpolation interfaces -> levels
    c, b, e = new(L, L + 1)
    for k in range(L):
        c[k, k:k + 2] = _lagrange_weights(ze[k:k + 2], zn[k])
        b[k], e[k] = k, k + 2
    ops["interp_e2n"] = (c, b, e)
    # derivative levels -> levels: centred, one-sided at the ends
    c, b, e = new(L, L)
    for k in range(L):
        lo, hi = max(k - 1, 0), min(k + 2, L)
        c[k, lo:hi] = _lagrange_dweights(zn[lo:hi], zn[k])
        b[k], e[k] = lo, hi
    ops["diff_n2n"] = (c, b, e)
    # derivative levels -> interfaces: two-point, zero at both boundaries
    c, b, e = new(L + 1, L)
    for k in range(L + 1):
        if k == 0 or k == L:
            lo = 0 if k == 0 else L - 2
            b[k], e[k] = lo, lo + 2
            continue
        c[k, k - 1:k + 1] = _lagrange_dweights(zn[k - 1:k + 1], ze[k])
        b[k], e[k] = max(k - 2, 0), k + 1
    ops["diff_n2e"] = (c, b, e)
    # derivative interfaces -> levels
    c, b, e = new(L, L + 1)
    for k in range(L):
        c[k, k:k + 2] = _lagrange_dweights(ze[k:k + 2], zn[k])
        b[k], e[k] = k, k + 2
    ops["diff_e2n"] = (c, b, e)
    # derivative interfaces -> interfaces
    c, b, e = new(L + 1, L + 1)
    for k in range(L + 1):
        lo, hi = max(k - 1, 0), min(k + 2, L + 1)
        c[k, lo:hi] = _lagrange_dweights(ze[lo:hi], ze[k])
        b[k], e[k] = lo, hi
    ops["diff_e2e"] = (c, b, e)
    # second derivative levels -> levels (one-sided first-difference rows at
    # the ends, as the reference's FE operator yields for order 1)
    dz = 1.0 / L
    c, b, e = new(L, L)
    for k in range(L):
        if k == 0:
            c[k, 0:2] = [-1.0 / dz ** 2, 1.0 / dz ** 2]
            b[k], e[k] = 0, 2
        elif k == L - 1:
            c[k, L - 2:L] = [1.0 / dz ** 2, -1.0 / dz ** 2]
            b[k], e[k] = L - 2, L
        else:
            c[k, k - 1:k + 2] = _lagrange_ddweights(zn[k - 1:k + 2], zn[k])
            b[k], e[k] = k - 1, k + 2
    ops["diffdiff_n2n"] = (c, b, e)
    c, b, e = new(L + 1, L + 1)
    for k in range(L + 1):
        if k == 0:
            c[k, 0:2] = [-2.0 / dz ** 2, 2.0 / dz ** 2]
            b[k], e[k] = 0, 2
        elif k == L:
            c[k, L - 1:L + 1] = [2.0 / dz ** 2, -2.0 / dz ** 2]
            b[k], e[k] = L - 1, L + 1
        else:
            c[k, k - 1:k + 2] = _lagrange_ddweights(ze[k - 1:k + 2], ze[k])
            b[k], e[k] = k - 1, k + 2
    ops["diffdiff_e2e"] = (c, b, e)
    # discontinuous penalty (LinearColumnDiscPenaltyFEM::Initialize,
    # LinearColumnOperatorFEM.cpp:1740-1850, order 1): jump across the
    # interface above (left op) / below (right op) the level over the level depth
    c, b, e = new(L, L)
    for k in range(L - 1):
        c[k, k] = -0.5 / dz
        c[k, k + 1] = 0.5 / dz
        b[k], e[k] = k, k + 2
    ops["penalty_left"] = (c, b, e)
    c, b, e = new(L, L)
    for k in range(1, L):
        c[k, k - 1] = 0.5 / dz
        c[k, k] = -0.5 / dz
        b[k], e[k] = k - 1, k + 1
    ops["penalty_right"] = (c, b, e)
    return ops


# -- cubed-sphere geometry --------------------------------------------------------

def rll_from_xyp(X, Y, panel):
    """CubedSphereTrans::RLLFromXYP (reference CubedSphereTrans.cpp:200-266)."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    if panel < 4:
        lon = np.arctan(X) + 0.5 * math.pi * panel
        lat = np.arctan(Y / np.sqrt(1.0 + X * X))
    elif panel == 4:
        lon = np.where(np.abs(X) > np.finfo(float).eps, np.arctan2(X, -Y),
                       np.where(Y <= 0.0, 0.0, math.pi))
        lat = 0.5 * math.pi - np.arctan(np.sqrt(X * X + Y * Y))
    else:
        lon = np.where(np.abs(X) > np.finfo(float).eps, np.arctan2(X, Y),
                       np.where(Y > 0.0, 0.0, math.pi))
        lat = -0.5 * math.pi + np.arctan(np.sqrt(X * X + Y * Y))
    lon = np.where(lon < 0.0, lon + 2.0 * math.pi, lon)
    return lon, lat


def covec_abp_from_rll(X, Y, panel, ulon, ulat):
    """Covariant (u_alpha, u_beta) of a vector given by zonal / meridional
    components times the radius (CubedSphereTrans::CoVecTransABPFromRLL,
    reference CubedSphereTrans.cpp:549-636)."""
    d2 = 1.0 + X * X + Y * Y
    if panel < 4:
        lat = np.arctan(Y / np.sqrt(1.0 + X * X))
        ul = ulon / np.cos(lat)
        ua = (1.0 + X * X) / d2 * ul - X * Y * np.sqrt(1.0 + X * X) / d2 * ulat
        ub = np.sqrt(1.0 + X * X) * (1.0 + Y * Y) / d2 * ulat
        return ua, ub
    rad = np.sqrt(X * X + Y * Y)
    pole = (np.abs(X) < 1.0e-13) & (np.abs(Y) < 1.0e-13)
    rs = np.where(pole, 1.0, rad)
    sgn = 1.0 if panel == 4 else -1.0
    lat = sgn * (0.5 * math.pi - np.arctan(rad))
    cl = np.where(pole, 1.0, np.cos(lat))
    ul = ulon / cl
    ua = sgn * (-Y * (1.0 + X * X) / d2 * ul - X * (1.0 + X * X) / (d2 * rs) * ulat)
    ub = sgn * (+X * (1.0 + Y * Y) / d2 * ul - Y * (1.0 + Y * Y) / (d2 * rs) * ulat)
    ua = np.where(pole, sgn * ulon, ua)
    ub = np.where(pole, ulat, ub)
    return ua, ub


class GridPatchCSGLL:
    """One patch: index box, coordinates, metric terms in the reference's host
    layout ([iA][iB][k][m] with a one-node halo)."""

    def __init__(self, grid, index, panel, ea0, eb0, nea, neb):
        self.grid = grid
        self.index = index
        self.panel = panel
        self.ea0, self.eb0, self.nea, self.neb = ea0, eb0, nea, neb
        self.halo = 1
        np_ = grid.np
        self.wa = nea * np_ + 2
        self.wb = neb * np_ + 2
        self.delta = 0.5 * math.pi / grid.ne
        # alpha / beta of interior nodes (GridPatchCSGLL::InitializeCoordinateData,
        # GridPatchCSGLL.cpp:176-215)
        self.anode = cs.alpha_nodes(ea0, nea, grid.ne, np_)
        self.bnode = cs.alpha_nodes(eb0, neb, grid.ne, np_)
        self.X = np.tan(self.anode)
        self.Y = np.tan(self.bnode)
        XX, YY = np.meshgrid(self.X, self.Y, indexing="ij")
        self.XX, self.YY = XX, YY
        self.lon, self.lat = rll_from_xyp(XX, YY, panel)

    def _pad(self, a):
        """interior array [wa-2][wb-2][...] -> host layout with zero halo"""
        out = np.zeros((self.wa, self.wb) + a.shape[2:])
        out[1:-1, 1:-1] = a
        return out

    def node_ids(self):
        g = self.grid
        return cs.node_ids(self.panel, self.nea, self.neb, self.ea0, self.eb0, g.ne, g.np)

    def seam_transforms(self):
        g = self.grid
        return cs.seam_transforms(self.panel, self.nea, self.neb, self.ea0, self.eb0,
                                  g.ne, g.np, self.anode, self.bnode)

    def evaluate_geometric_terms(self, zs, dazs, dbzs, lean=False):
        """GridPatchCSGLL::EvaluateGeometricTerms (GridPatchCSGLL.cpp:295-574).
        zs, dazs, dbzs: topography and its (DSS'd) derivatives on interior nodes.
        lean: the 2-D terms and the element areas only (the device rebuilds the
        3-D metric from column constants)."""
        g = self.grid
        phys = g.phys
        L = g.nlev
        X, Y = self.XX, self.YY
        a = phys.earth_radius
        d2 = 1.0 + X * X + Y * Y
        d = np.sqrt(d2)
        if g.is2d:
            zs = np.zeros_like(X)
            dazs = np.zeros_like(X)
            dbzs = np.zeros_like(X)
        j2d = (1.0 + X * X) * (1.0 + Y * Y) / (d * d * d)
        j2d = j2d * (a * a)
        scale = d2 / (1.0 + X * X) / (1.0 + Y * Y) / (a * a)
        c2a = np.stack([scale * (1.0 + Y * Y), scale * X * Y], axis=-1)
        c2b = np.stack([scale * X * Y, scale * (1.0 + X * X)], axis=-1)
        out = dict(
            jacobian2d=self._pad(j2d), contrametric2da=self._pad(c2a),
            contrametric2db=self._pad(c2b),
            coriolis=self._pad(2.0 * phys.omega * np.sin(self.lat)),
            topography=self._pad(zs))
        gl, wl = cs.gll_points(g.np, 0.0, 1.0)
        wi = np.tile(wl, self.nea)[:, None]
        wj = np.tile(wl, self.neb)[None, :]

        def column(reta, warea):
            n = len(reta)
            e = reta[None, None, :]
            dxr = (g.ztop - zs)[:, :, None] * np.ones((1, 1, n))
            dar = (1.0 - e) * dazs[:, :, None]
            dbr = (1.0 - e) * dbzs[:, :, None]
            jac = dxr * j2d[:, :, None]
            area = jac * (wi * self.delta)[:, :, None] * (wj * self.delta)[:, :, None] \
                * warea[None, None, :]
            sc = scale[:, :, None]
            x, y = X[:, :, None], Y[:, :, None]
            ca2 = -sc / dxr * ((1.0 + y * y) * dar + x * y * dbr)
            cb2 = -sc / dxr * (x * y * dar + (1.0 + x * x) * dbr)
            cx2 = 1.0 / (dxr * dxr) - 1.0 / dxr * (ca2 * dar + cb2 * dbr)
            one = np.ones((1, 1, n))
            ca = np.stack([c2a[:, :, None, 0] * one, c2a[:, :, None, 1] * one, ca2], axis=-1)
            cb = np.stack([c2b[:, :, None, 0] * one, c2b[:, :, None, 1] * one, cb2], axis=-1)
            cx = np.stack([ca2, cb2, cx2], axis=-1)
            dr = np.stack([dar, dbr, dxr], axis=-1)
            return jac, area, ca, cb, cx, dr

        if lean:
            for name, reta, warea in (("area_node", g.reta_levels, g.reta_levels_area),
                                      ("area_redge", g.reta_interfaces, g.reta_interfaces_area)):
                col = ((g.ztop - zs) * j2d * (wi * self.delta) * (wj * self.delta))
                setattr(self, name, self._pad(col[:, :, None] * warea[None, None, :]))
            self.zs = zs
            return out
        jac, area, ca, cb, cx, dr = column(g.reta_levels, g.reta_levels_area)
        out.update(jacobian=self._pad(jac), contrametrica=self._pad(ca),
                   contrametricb=self._pad(cb), contrametricxi=self._pad(cx),
                   derivr_node=self._pad(dr))
        self.area_node = self._pad(area)
        jac, area, ca, cb, cx, dr = column(g.reta_interfaces, g.reta_interfaces_area)
        out.update(jacobian_redge=self._pad(jac), contrametrica_redge=self._pad(ca),
                   contrametricb_redge=self._pad(cb), contrametricxi_redge=self._pad(cx),
                   derivr_redge=self._pad(dr))
        self.area_redge = self._pad(area)
        self.zs = zs
        return out


class GridCSGLL:
    """Cubed-sphere GLL grid: parameters, tables, vertical coordinate, patches
    (GridCSGLL::SetParameters / ApplyDefaultPatchLayout, GridCSGLL.cpp:40-148;
    GridGLL::Initialize, GridGLL.cpp:101-363)."""

    def __init__(self, ne, nlev, np_=4, vertical_order=1, npatch=6, ztop=1.0,
                 phys=None, is2d=False):
        if vertical_order != 1:
            raise NotImplementedError(
                "the numpy setup path builds vertical order 1 operators only")
        self.ne, self.nlev, self.np = ne, nlev, np_
        self.vertical_order = vertical_order
        self.ztop = ztop
        self.phys = phys or PhysicalConstants()
        self.is2d = is2d
        # reference length of the hyperviscosity scaling (GridCSGLL.cpp:87)
        self.reference_length = 0.5 * math.pi / 30.0
        self.dx, self.stiffness, self.gll_weights = gll_tables(np_)
        L = nlev
        self.reta_levels = (np.arange(L) + 0.5) / L
        self.reta_interfaces = np.arange(L + 1) / float(L)
        self.reta_levels_area = np.full(L, 1.0 / L)
        wi = np.full(L + 1, 1.0 / L)
        wi[0] = wi[-1] = 0.5 / L
        self.reta_interfaces_area = wi
        self.ops = column_operators(L) if L > 1 else {}
        # ApplyDefaultPatchLayout
        k = max(int(math.isqrt(npatch // 6)), 1)
        if ne % k != 0:
            raise ValueError("elements must divide equally among patches")
        per = ne // k
        self.patches = []
        ix = 0
        for panel in range(6):
            for i in range(k):
                for j in range(k):
                    self.patches.append(GridPatchCSGLL(
                        self, ix, panel, i * per, j * per, per, per))
                    ix += 1

    @property
    def column_count(self):
        return 6 * self.ne * self.ne * self.np * self.np

    # -- topography: pointwise values, GLL derivatives, DSS of the derivatives
    #    (GridPatchCSGLL::EvaluateTopography :218-291; GridGLL.cpp:557-567)
    def evaluate_topography(self, test):
        np_ = self.np
        ids, grads, meta = [], [], []
        for p in self.patches:
            zs = test.evaluate_topography(self.phys, p.lon, p.lat)
            z4 = zs.reshape(p.nea, np_, p.neb, np_)
            da = np.einsum("si,asbj->aibj", self.dx, z4).reshape(zs.shape) / p.delta
            db = np.einsum("sj,aibs->aibj", self.dx, z4).reshape(zs.shape) / p.delta
            p._zs, p._da, p._db = zs, da, db
        if all(np.all(p._zs == 0.0) for p in self.patches):
            for p in self.patches:
                p._dazs, p._dbzs = p._da, p._db
            return
        # average the gradient covector over duplicates in a panel-independent
        # (Cartesian) representation: same map as the reference's pairwise
        # TransformTopographyDeriv + DSS.
        for p in self.patches:
            ea, eb, ca, cb = _bases(p.panel, p.XX, p.YY)
            v = p._da[..., None] * ca + p._db[..., None] * cb
            ids.append(p.node_ids().reshape(-1))
            grads.append(v.reshape(-1, 3))
            meta.append((ea, eb))
        allid = np.concatenate(ids)
        allv = np.concatenate(grads)
        uniq, inv = np.unique(allid, return_inverse=True)
        cnt = np.bincount(inv).astype(np.float64)
        # the 2 - 4 contributions of a node are summed in ascending order of their
        # values, not in the order the patches are listed: the metric (and with it
        # the state) must not depend on how a panel is cut into patches
        avg = np.empty((len(uniq), 3))
        for c in range(3):
            order = np.lexsort((allv[:, c], inv))
            starts = np.searchsorted(inv[order], np.arange(len(uniq)))
            avg[:, c] = np.add.reduceat(allv[order, c], starts) / cnt
        off = 0
        for p, (ea, eb) in zip(self.patches, meta):
            n = p.XX.size
            v = avg[inv[off:off + n]].reshape(p.XX.shape + (3,))
            off += n
            p._dazs = np.einsum("abc,abc->ab", v, ea)
            p._dbzs = np.einsum("abc,abc->ab", v, eb)

    def evaluate_geometric_terms(self):
        return {p.index: p.evaluate_geometric_terms(p._zs, p._dazs, p._dbzs)
                for p in self.patches}


def _bases(panel, X, Y):
    """covariant (e_alpha, e_beta) and contravariant (e^alpha, e^beta) basis
    vectors on the unit sphere for arrays X, Y."""
    one = np.ones_like(X)
    zero = np.zeros_like(X)
    d = {0: (one, X, Y), 1: (-X, one, Y), 2: (-one, -X, Y), 3: (X, -one, Y),
         4: (-Y, X, one), 5: (Y, X, -one)}[panel]
    dX = {0: (zero, one, zero), 1: (-one, zero, zero), 2: (zero, -one, zero),
          3: (one, zero, zero), 4: (zero, one, zero), 5: (zero, one, zero)}[panel]
    dY = {0: (zero, zero, one), 1: (zero, zero, one), 2: (zero, zero, one),
          3: (zero, zero, one), 4: (-one, zero, zero), 5: (one, zero, zero)}[panel]
    d = np.stack(d, axis=-1)
    dX = np.stack(dX, axis=-1)
    dY = np.stack(dY, axis=-1)
    r = np.sqrt((d * d).sum(-1))[..., None]
    ex = dX / r - d * (d * dX).sum(-1)[..., None] / r ** 3
    ey = dY / r - d * (d * dY).sum(-1)[..., None] / r ** 3
    ea = ex * (1.0 + X * X)[..., None]
    eb = ey * (1.0 + Y * Y)[..., None]
    gaa = (ea * ea).sum(-1)
    gab = (ea * eb).sum(-1)
    gbb = (eb * eb).sum(-1)
    det = gaa * gbb - gab * gab
    ca = (gbb[..., None] * ea - gab[..., None] * eb) / det[..., None]
    cb = (-gab[..., None] * ea + gaa[..., None] * eb) / det[..., None]
    return ea, eb, ca, cb
