"""Cartesian GLL grids (reference src/atm/GridCartesianGLL.{h,cpp}): node
identification for the averaging groups of the DSS.

The reference fills one-node halos through its exchange connectivity - periodic
boundaries wrap the global element indices (GridCartesianGLL.cpp:380-432) - and
then averages across every element edge (ApplyDSS, :508-654).  Here a node is
named by its periodic global index, so that duplicates across element edges,
patch edges and periodic boundaries land in one averaging group.  With a single
element across a periodic direction (XZ slices: one element in y) the two
edges of the same element are duplicates of each other, as in the reference.
"""
import numpy as np


def node_ids(nelem_a, nelem_b, elem_a0, elem_b0, ne_a, ne_b, np_,
             periodic_a=True, periodic_b=True):
    """Global ids [nelem_a*np][nelem_b*np] of a patch whose first element is
    (elem_a0, elem_b0) on a domain of ne_a x ne_b elements."""
    if not (periodic_a and periodic_b):
        raise NotImplementedError(
            "non-periodic Cartesian boundaries (GridPatchCartesianGLL::"
            "ApplyBoundaryConditions) are not implemented")
    ea = elem_a0 + np.arange(nelem_a * np_) // np_
    ia = np.arange(nelem_a * np_) % np_
    eb = elem_b0 + np.arange(nelem_b * np_) // np_
    jb = np.arange(nelem_b * np_) % np_
    na = ne_a * (np_ - 1)
    nb = ne_b * (np_ - 1)
    ua = (ea * (np_ - 1) + ia) % na
    ub = (eb * (np_ - 1) + jb) % nb
    A, B = np.meshgrid(ua, ub, indexing="ij")
    return (A.astype(np.int64) * nb + B).astype(np.int64)


# -- grid and patch for the Python driver ---------------------------------------
import math

from . import cubedsphere as cs
from .grid import PhysicalConstants, column_operators, gll_tables


class GridPatchCartesianGLL:
    """One patch of a Cartesian GLL grid: coordinates and the metric of flat
    terrain in the reference's host layout
    (GridPatchCartesianGLL::InitializeCoordinateData / EvaluateGeometricTerms,
    GridPatchCartesianGLL.cpp:100-460 with zs = 0)."""

    panel = 0

    def __init__(self, grid, index, ea0, eb0, nea, neb):
        self.grid = grid
        self.index = index
        self.ea0, self.eb0, self.nea, self.neb = ea0, eb0, nea, neb
        self.halo = 1
        np_ = grid.np
        self.wa = nea * np_ + 2
        self.wb = neb * np_ + 2
        self.delta_a = (grid.dims[1] - grid.dims[0]) / grid.ne_a
        self.delta_b = (grid.dims[3] - grid.dims[2]) / grid.ne_b
        self.delta = self.delta_a
        ga, _ = cs.gll_points(np_, 0.0, self.delta_a)
        gb, _ = cs.gll_points(np_, 0.0, self.delta_b)
        ea = ea0 + np.arange(nea)
        eb = eb0 + np.arange(neb)
        self.anode = (grid.dims[0] + self.delta_a * ea[:, None] + ga[None, :]).reshape(-1)
        self.bnode = (grid.dims[2] + self.delta_b * eb[:, None] + gb[None, :]).reshape(-1)
        self.X, self.Y = self.anode, self.bnode
        self.XX, self.YY = np.meshgrid(self.anode, self.bnode, indexing="ij")
        # the reference stores x and y in its longitude / latitude arrays
        self.lon, self.lat = self.XX, self.YY

    def _pad(self, a):
        out = np.zeros((self.wa, self.wb) + a.shape[2:])
        out[1:-1, 1:-1] = a
        return out

    def node_ids(self):
        g = self.grid
        return node_ids(self.nea, self.neb, self.ea0, self.eb0, g.ne_a, g.ne_b, g.np)

    def seam_transforms(self):
        z = np.zeros(0, dtype=np.int32)
        return z, z, z, np.zeros((0, 4))

    def evaluate_geometric_terms(self, zs, dazs, dbzs):
        g = self.grid
        if np.any(zs != 0.0):
            raise NotImplementedError(
                "Cartesian topography (the terrain decay of "
                "GridPatchCartesianGLL.cpp:262-330) is not restated here")
        sh = self.XX.shape
        one2 = np.ones(sh)
        zero2 = np.zeros(sh)
        gl, wl = cs.gll_points(g.np, 0.0, 1.0)
        wi = np.tile(wl, self.nea)[:, None]
        wj = np.tile(wl, self.neb)[None, :]
        out = dict(
            jacobian2d=self._pad(one2),
            contrametric2da=self._pad(np.stack([one2, zero2], axis=-1)),
            contrametric2db=self._pad(np.stack([zero2, one2], axis=-1)),
            coriolis=self._pad(zero2), topography=self._pad(zero2))

        def column(n, warea):
            one = np.ones(sh + (n,))
            zero = np.zeros(sh + (n,))
            dxz = g.ztop * one
            jac = dxz
            area = jac * (wi * self.delta_a)[:, :, None] * (wj * self.delta_b)[:, :, None] \
                * warea[None, None, :]
            ca = np.stack([one, zero, -zero / dxz], axis=-1)
            cb = np.stack([zero, one, -zero / dxz], axis=-1)
            cx = np.stack([-zero / dxz, -zero / dxz, one / (dxz * dxz)], axis=-1)
            dr = np.stack([zero, zero, dxz], axis=-1)
            return jac, area, ca, cb, cx, dr

        L = g.nlev
        jac, area, ca, cb, cx, dr = column(L, g.reta_levels_area)
        out.update(jacobian=self._pad(jac), contrametrica=self._pad(ca),
                   contrametricb=self._pad(cb), contrametricxi=self._pad(cx),
                   derivr_node=self._pad(dr))
        self.area_node = self._pad(area)
        jac, area, ca, cb, cx, dr = column(L + 1, g.reta_interfaces_area)
        out.update(jacobian_redge=self._pad(jac), contrametrica_redge=self._pad(ca),
                   contrametricb_redge=self._pad(cb), contrametricxi_redge=self._pad(cx),
                   derivr_redge=self._pad(dr))
        self.area_redge = self._pad(area)
        self.zs = zs
        return out


class GridCartesianGLL:
    """Periodic Cartesian GLL grid, optionally an x-z slice
    (GridCartesianGLL::SetParameters / ApplyDefaultPatchLayout,
    GridCartesianGLL.cpp:60-229): patches are strips along alpha."""

    is_cartesian = True

    def __init__(self, ne_a, ne_b, nlev, dims, np_=4, vertical_order=1, npatch=1,
                 xz=True, phys=None, reference_length=None):
        if vertical_order != 1:
            raise NotImplementedError(
                "the numpy setup path builds vertical order 1 operators only")
        if ne_a % npatch != 0:
            raise ValueError("elements must divide equally among patches")
        self.ne_a, self.ne_b, self.ne = ne_a, ne_b, ne_a
        self.nlev, self.np = nlev, np_
        self.vertical_order = vertical_order
        self.dims = tuple(float(x) for x in dims)
        self.ztop = self.dims[5]
        self.xz = bool(xz)
        self.is2d = False
        self.phys = phys or PhysicalConstants()
        xl = abs(self.dims[1] - self.dims[0])
        # ThermalBubbleCartesianTest.cpp:322-330
        self.reference_length = reference_length if reference_length is not None \
            else min(xl, 110000.0)
        self.dx, self.stiffness, self.gll_weights = gll_tables(np_)
        L = nlev
        self.reta_levels = (np.arange(L) + 0.5) / L
        self.reta_interfaces = np.arange(L + 1) / float(L)
        self.reta_levels_area = np.full(L, 1.0 / L)
        wi = np.full(L + 1, 1.0 / L)
        wi[0] = wi[-1] = 0.5 / L
        self.reta_interfaces_area = wi
        self.ops = column_operators(L) if L > 1 else {}
        per = ne_a // npatch
        self.patches = [GridPatchCartesianGLL(self, ix, ix * per, 0, per, ne_b)
                        for ix in range(npatch)]

    @property
    def column_count(self):
        return self.ne_a * self.ne_b * self.np * self.np

    def evaluate_topography(self, test):
        for p in self.patches:
            zs = test.evaluate_topography(self.phys, p.lon, p.lat)
            p._zs = zs
            p._da = p._db = p._dazs = p._dbzs = np.zeros_like(zs)
