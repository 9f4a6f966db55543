"""Thin object wrapper over the C ABI (include/tempest_b200.h).

Every method is one C-ABI call; arrays are numpy float64/int32/int64 in the
reference's host layout.  Errors of the library are raised as
:class:`TempestError` carrying ``tb200_last_error`` (the C++ shells raise the
reference's ``Exception`` at the same places, src/base/Exception.h:25-49).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import (DATA_ALL, DATA_STATE, DATA_TRACERS, EQN_PRIMITIVE_NONHYDRO,
                   EQN_SHALLOW_WATER, OP_NAMES, SCHEMES)


class TempestError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class DeviceContext:
    """One tb200_ctx: the device-resident dynamical core of one rank."""

    def __init__(self, library=None, **cfg):
        self.lib = _lib.load(library)
        c = _lib.Config()
        onedge = cfg.pop("comp_on_redge", [0] * 8)
        for k, v in cfg.items():
            if not hasattr(c, k):
                raise TypeError("unknown configuration field %r" % k)
            setattr(c, k, v)
        for i, v in enumerate(onedge):
            c.comp_on_redge[i] = int(v)
        self.cfg = c
        self._h = ctypes.c_void_p()
        rc = self.lib.tb200_create(ctypes.byref(c), ctypes.byref(self._h))
        if rc != 0:
            msg = self.lib.tb200_last_error(self._h).decode()
            self.lib.tb200_destroy(self._h)
            self._h = None
            raise TempestError(msg)
        self._keep = []
        self.patches = {}

    # -- plumbing ---------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise TempestError(self.lib.tb200_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.tb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        self._ck(self.lib.tb200_set_stream(self._h, ctypes.c_void_p(cuda_stream)))

    def sync(self):
        self._ck(self.lib.tb200_sync(self._h))

    def check_errors(self):
        self._ck(self.lib.tb200_check_errors(self._h))

    @property
    def launch_count(self):
        return int(self.lib.tb200_launch_count(self._h))

    @property
    def column_count(self):
        return int(self.lib.tb200_column_count(self._h))

    @property
    def fused_group_count(self):
        """Averaging groups the pipelined kernels average themselves (fused DSS)."""
        return int(self.lib.tb200_fused_group_count(self._h))

    # -- grid -----------------------------------------------------------------
    def set_exchange(self, rank, nranks, fn):
        cb = _lib.EXCHANGE_FN(fn) if fn is not None else _lib.EXCHANGE_FN()
        self._keep.append(cb)
        self._ck(self.lib.tb200_set_exchange(self._h, rank, nranks, cb, None))

    def add_patch(self, index, panel, nelem_a, nelem_b, halo, delta_a, delta_b,
                  owner_rank=0):
        self._ck(self.lib.tb200_add_patch(self._h, index, panel, nelem_a,
                                          nelem_b, halo, delta_a, delta_b,
                                          owner_rank))
        self.patches[index] = dict(panel=panel, nelem_a=nelem_a,
                                   nelem_b=nelem_b, halo=halo, owner=owner_rank)

    def commit_layout(self):
        self._ck(self.lib.tb200_commit_layout(self._h))

    def set_tables(self, dx_basis, stiffness, gll_weights):
        dx, st, w = _f64(dx_basis), _f64(stiffness), _f64(gll_weights)
        self._ck(self.lib.tb200_set_tables(self._h, _ptr(dx), _ptr(st), _ptr(w)))

    def set_column_op(self, op, coeff, begin, end):
        if isinstance(op, str):
            op = OP_NAMES.index(op)
        coeff = _f64(coeff)
        begin = np.ascontiguousarray(begin, dtype=np.int32)
        end = np.ascontiguousarray(end, dtype=np.int32)
        self._ck(self.lib.tb200_set_column_op(
            self._h, op, coeff.shape[0], coeff.shape[1], _ptr(coeff),
            _ptr(begin), _ptr(end)))

    def upload_geometry(self, patch, **arrays):
        g = _lib.Geometry()
        keep = []
        for name in _lib.GEOMETRY_FIELDS:
            a = _f64(arrays.get(name))
            keep.append(a)
            setattr(g, name, None if a is None else a.ctypes.data)
        self._ck(self.lib.tb200_upload_geometry(self._h, patch, ctypes.byref(g)))

    def set_terrain_metric(self, patch, xnode, ynode, topography_deriv):
        x, y, t = _f64(xnode), _f64(ynode), _f64(topography_deriv)
        self._ck(self.lib.tb200_set_terrain_metric(self._h, patch, _ptr(x), _ptr(y),
                                                   _ptr(t)))

    def set_vertical_coordinate(self, reta_levels, reta_interfaces):
        a, b = _f64(reta_levels), _f64(reta_interfaces)
        self._ck(self.lib.tb200_set_vertical_coordinate(self._h, _ptr(a), _ptr(b)))

    def fast_path(self):
        """-> (enabled, reason, metric deviation) of the order-1 fast kernels."""
        r = self.lib.tb200_fast_path(self._h)
        if r < 0:
            self._ck(1)
        return (r == 1, self.lib.tb200_fast_path_reason(self._h).decode(),
                self.lib.tb200_fast_path_metric_error(self._h))

    def upload_element_area(self, patch, area_node, area_redge):
        a, b = _f64(area_node), _f64(area_redge)
        self._ck(self.lib.tb200_upload_element_area(self._h, patch, _ptr(a), _ptr(b)))

    def upload_rayleigh(self, patch, strength_node, strength_redge, ref_node, ref_redge):
        a, b, c, d = (_f64(strength_node), _f64(strength_redge), _f64(ref_node),
                      _f64(ref_redge))
        self._ck(self.lib.tb200_upload_rayleigh(self._h, patch, _ptr(a), _ptr(b), _ptr(c),
                                                _ptr(d)))

    def upload_reference_state(self, patch, ref_node, ref_redge):
        """GridPatch::GetReferenceState of a local patch (state layout)."""
        a, b = _f64(ref_node), _f64(ref_redge)
        self._ck(self.lib.tb200_upload_reference_state(self._h, patch, _ptr(a), _ptr(b)))

    def set_vertical_discretization(self, finite_volume):
        """--vdisc FV: the column operators supplied are the finite-volume ones."""
        self._ck(self.lib.tb200_set_vertical_discretization(self._h, int(bool(finite_volume))))

    def set_mass_flux_on_levels(self, on):
        """--vmassfluxlevels: BuildF forms the mass and rho-theta fluxes on levels."""
        self._ck(self.lib.tb200_set_mass_flux_on_levels(self._h, int(bool(on))))

    def set_uniform_diffusion(self, scalar_coeff, vector_coeff):
        """TestCase::GetUniformDiffusionCoeffs -> Grid::HasUniformDiffusion: uniform
        second-order diffusion of the state minus the reference state."""
        self._ck(self.lib.tb200_set_uniform_diffusion(self._h, float(scalar_coeff),
                                                      float(vector_coeff)))

    # -- device-side set-up (tb200_setup.cuh) ------------------------------------
    def evaluate_geometry_cs(self, patch, radius, omega):
        """2-D metric, Coriolis parameter, longitude / latitude of a cubed-sphere
        patch from the node coordinates already on the device."""
        self._ck(self.lib.tb200_evaluate_geometry_cs(self._h, patch, radius, omega))

    def compute_output_fields(self, inst):
        """Temperature, relative vorticity and divergence of an instance on levels,
        kept on the device for interpolate()."""
        self._ck(self.lib.tb200_compute_output_fields(self._h, inst))

    def interpolate(self, inst, data_type, only_location, patch, elem_a, elem_b, ca, cb,
                    alpha, beta, nout, vop_node=None, vop_redge=None, primitive=True):
        """Grid::ReduceInterpolate on the device (tb200_output.cuh).  vop_*: (dense
        matrix, begin, end) of the LinearColumnInterpFEM operator or None (identity).
        Returns [components or tracers][nout][npts]."""
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        patch, elem_a, elem_b = i32(patch), i32(elem_a), i32(elem_b)
        ca, cb, alpha, beta = _f64(ca), _f64(cb), _f64(alpha), _f64(beta)
        npts = len(patch)
        ncomp = self.cfg.ntracers if data_type == _lib.DATA_TRACERS else self.cfg.ncomp
        if data_type in (_lib.DATA_TEMPERATURE, _lib.DATA_VORTICITY, _lib.DATA_DIVERGENCE):
            ncomp = 1
        out = np.zeros((ncomp, nout, npts))
        keep = []

        def op(v):
            if v is None:
                return None, None, None
            m, b, e = _f64(v[0]), i32(v[1]), i32(v[2])
            keep.extend([m, b, e])
            return _ptr(m), _ptr(b), _ptr(e)
        vn, ve = op(vop_node), op(vop_redge)
        self._ck(self.lib.tb200_interpolate(
            self._h, inst, data_type, only_location, npts, _ptr(patch), _ptr(elem_a),
            _ptr(elem_b), _ptr(ca), _ptr(cb), _ptr(alpha), _ptr(beta), nout,
            vn[0], vn[1], vn[2], ve[0], ve[1], ve[2], 1 if primitive else 0, _ptr(out)))
        return out

    def column_field(self, which, nelem, nn=16):
        """Per-column device array [element][np * np] (0 Jacobian2D, 1, 2
        ContraMetric2DA, 3, 4 ContraMetric2DB, 5 Coriolis, 6 topography, 7 longitude,
        8 latitude): read-back for tests."""
        out = np.zeros((nelem, nn))
        self._ck(self.lib.tb200_debug_column_field(self._h, which, _ptr(out)))
        return out

    @staticmethod
    def _jw(test, phys):
        from ._lib import JWTest
        return JWTest(eta0=test.eta0, tropopause_eta=test.tropopause_eta, t0=test.t0,
                      delta_t=test.delta_t, lapse_rate=test.lapse_rate, u0=test.u0,
                      up=test.up, pert_lon=test.pert_lon, pert_lat=test.pert_lat,
                      pert_r=test.pert_r, perturbation=1 if test.perturbation == "exp" else 0,
                      omega=phys.omega, radius=phys.earth_radius)

    def evaluate_jw_topography(self, patch, test, phys):
        import ctypes
        self._ck(self.lib.tb200_evaluate_jw_topography(self._h, patch,
                                                       ctypes.byref(self._jw(test, phys))))

    def evaluate_jw_state(self, patch, inst, test, phys):
        """Initial state of BaroclinicWaveJWTest written straight into `inst`."""
        import ctypes
        self._ck(self.lib.tb200_evaluate_jw_state(self._h, patch, inst,
                                                  ctypes.byref(self._jw(test, phys))))

    def upload_held_suarez(self, patch, latitude, surface_product):
        """Per-column inputs of HeldSuarezPhysics::Perform: latitude and the product
        of the rho and rho-theta slots on the lowest interface of instance 0
        (HeldSuarezPhysics.cpp:95-115), [iA][iB] with halo."""
        a, b = _f64(latitude), _f64(surface_product)
        self._ck(self.lib.tb200_upload_held_suarez(self._h, patch, _ptr(a), _ptr(b)))

    def held_suarez(self, dt):
        """HeldSuarezPhysics::Perform on instance 0 (device workflow step)."""
        self._ck(self.lib.tb200_held_suarez(self._h, dt))

    def kessler(self, dt):
        """KesslerPhysics::Perform on instance 0 (device workflow step; parity
        unpinned, see include/tempest_b200.h)."""
        self._ck(self.lib.tb200_kessler(self._h, dt))

    def set_node_ids(self, patch, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        self._ck(self.lib.tb200_set_node_ids(self._h, patch, _ptr(ids)))

    def set_seam_transforms(self, patch, ia, ib, src_panel, mats):
        ia = np.ascontiguousarray(ia, dtype=np.int32)
        ib = np.ascontiguousarray(ib, dtype=np.int32)
        sp = np.ascontiguousarray(src_panel, dtype=np.int32)
        m = _f64(mats)
        self._ck(self.lib.tb200_set_seam_transforms(
            self._h, patch, len(ia), _ptr(ia), _ptr(ib), _ptr(sp), _ptr(m)))

    def build_connectivity(self):
        self._ck(self.lib.tb200_build_connectivity(self._h))

    def exchange_counts(self, nranks):
        s = np.zeros(nranks, dtype=np.int64)
        r = np.zeros(nranks, dtype=np.int64)
        self._ck(self.lib.tb200_exchange_counts(self._h, _ptr(s), _ptr(r)))
        return s, r

    def peer_export(self, nranks):
        """(IPC handle bytes, first receive slot per source rank, receive slots)"""
        h = np.zeros(64, dtype=np.uint8)
        off = np.zeros(nranks, dtype=np.int64)
        tot = np.zeros(1, dtype=np.int64)
        self._ck(self.lib.tb200_peer_export(self._h, _ptr(h), _ptr(off), _ptr(tot)))
        return h.tobytes(), off.tolist(), int(tot[0])

    def peer_attach(self, handles, my_offset_at, recv_totals):
        h = np.frombuffer(b"".join(handles), dtype=np.uint8).copy()
        off = np.ascontiguousarray(my_offset_at, dtype=np.int64)
        tot = np.ascontiguousarray(recv_totals, dtype=np.int64)
        self._ck(self.lib.tb200_peer_attach(self._h, _ptr(h), _ptr(off), _ptr(tot)))

    def peer_detach(self):
        self._ck(self.lib.tb200_peer_detach(self._h))

    # -- state ------------------------------------------------------------------
    def upload_state(self, patch, inst, node=None, redge=None, tracers=None):
        n, e, t = _f64(node), _f64(redge), _f64(tracers)
        self._ck(self.lib.tb200_upload_state(self._h, patch, inst, _ptr(n),
                                             _ptr(e), _ptr(t)))

    def download_state(self, patch, inst, node=None, redge=None, tracers=None,
                       fill_derived=True):
        for a in (node, redge, tracers):
            if a is not None:
                assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        self._ck(self.lib.tb200_download_state(
            self._h, patch, inst, _ptr(node), _ptr(redge), _ptr(tracers),
            1 if fill_derived else 0))

    def upload_state_async(self, patch, inst, node=None, redge=None, tracers=None):
        """Enqueue only: the arrays (C-contiguous float64, ideally pinned) must stay
        alive and unchanged until transfer_sync()."""
        for a in (node, redge, tracers):
            if a is not None:
                assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        self._ck(self.lib.tb200_upload_state_async(self._h, patch, inst, _ptr(node),
                                                   _ptr(redge), _ptr(tracers)))

    def download_state_async(self, patch, inst, node=None, redge=None, tracers=None,
                             fill_derived=True):
        for a in (node, redge, tracers):
            if a is not None:
                assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        self._ck(self.lib.tb200_download_state_async(
            self._h, patch, inst, _ptr(node), _ptr(redge), _ptr(tracers),
            1 if fill_derived else 0))

    def transfer_sync(self):
        self._ck(self.lib.tb200_transfer_sync(self._h))

    def copy(self, src, dst, mask=DATA_ALL):
        self._ck(self.lib.tb200_copy(self._h, src, dst, mask))

    def lincomb(self, coeff, dst, mask=DATA_ALL):
        c = _f64(coeff)
        self._ck(self.lib.tb200_lincomb(self._h, _ptr(c), len(c), dst, mask))

    def lincomb_v_filter(self, coeff, dst):
        """Grid::LinearCombineData(coeff -> dst) of state and tracers followed by
        VerticalDynamics::FilterNegativeTracers(dst) (start of a Strang step)."""
        c = _f64(coeff)
        self._ck(self.lib.tb200_lincomb_v_filter(self._h, _ptr(c), len(c), dst))

    def filter_negative_tracers(self, inst):
        """HorizontalDynamicsFEM::FilterNegativeTracers (element-wise)."""
        self._ck(self.lib.tb200_filter_negative_tracers(self._h, inst))

    def v_filter_negative_tracers(self, inst):
        """VerticalDynamicsFEM::FilterNegativeTracers (column-wise)."""
        self._ck(self.lib.tb200_v_filter_negative_tracers(self._h, inst))

    def zero(self, inst, mask=DATA_ALL):
        self._ck(self.lib.tb200_zero(self._h, inst, mask))

    # -- dynamics ---------------------------------------------------------------
    def h_step_explicit(self, i_in, i_out, dt):
        self._ck(self.lib.tb200_h_step_explicit(self._h, i_in, i_out, dt))

    def v_step_explicit(self, i_in, i_out, dt):
        self._ck(self.lib.tb200_v_step_explicit(self._h, i_in, i_out, dt))

    def hv_step_explicit(self, i_in, i_out, dt):
        self._ck(self.lib.tb200_hv_step_explicit(self._h, i_in, i_out, dt))

    def hv_step_explicit_combine(self, coeff, i_in, i_out, dt):
        c = _f64(coeff)
        self._ck(self.lib.tb200_hv_step_explicit_combine(self._h, _ptr(c), len(c),
                                                         i_in, i_out, dt))

    def v_step_implicit(self, i_in, i_out, dt):
        self._ck(self.lib.tb200_v_step_implicit(self._h, i_in, i_out, dt))

    def hv_step_explicit_combine_dss(self, coeff, i_in, i_out, dt):
        """The explicit substage of every time scheme: combination, both explicit
        plugins and PostProcessSubstage(out) (the DSS fused into the stage kernel
        where the groups allow)."""
        c = np.ascontiguousarray(coeff, dtype=np.float64)
        self._ck(self.lib.tb200_hv_step_explicit_combine_dss(self._h, _ptr(c), len(c),
                                                             i_in, i_out, dt))

    def dss(self, inst, mask=DATA_ALL):
        self._ck(self.lib.tb200_dss(self._h, inst, mask))

    def h_step_after_subcycle(self, i_in, i_out, i_work, dt):
        self._ck(self.lib.tb200_h_step_after_subcycle(self._h, i_in, i_out,
                                                      i_work, dt))

    def step(self, scheme, first, last, dt):
        if isinstance(scheme, str):
            scheme = SCHEMES[scheme.lower()]
        self._ck(self.lib.tb200_step(self._h, scheme, int(first), int(last), dt))

    def checksum(self, inst):
        s = np.zeros(self.cfg.ncomp, dtype=np.float64)
        self._ck(self.lib.tb200_checksum(self._h, inst, _ptr(s)))
        return s

    def total_energy(self, inst):
        """Grid::ComputeTotalEnergy over the local patches."""
        v = np.zeros(1, dtype=np.float64)
        self._ck(self.lib.tb200_total_energy(self._h, inst, _ptr(v)))
        return float(v[0])

    def total_potential_enstrophy(self, inst, work=-1):
        """Grid::ComputeTotalPotentialEnstrophy (shallow water: `work` is a scratch
        instance for the DSS'd vorticity)."""
        v = np.zeros(1, dtype=np.float64)
        self._ck(self.lib.tb200_total_potential_enstrophy(self._h, inst, work, _ptr(v)))
        return float(v[0])

    def total_vertical_momentum(self, inst):
        v = np.zeros(1, dtype=np.float64)
        self._ck(self.lib.tb200_total_vertical_momentum(self._h, inst, _ptr(v)))
        return float(v[0])

    def test_band_solve(self, ab, b, kl, ku):
        ab = np.ascontiguousarray(ab, dtype=np.float64)
        x = np.array(b, dtype=np.float64, order="C", copy=True)
        ncols, n, _ = ab.shape
        self._ck(self.lib.tb200_test_band_solve(self._h, ncols, n, kl, ku,
                                                _ptr(ab), _ptr(x)))
        return x


__all__ = ["DeviceContext", "TempestError", "DATA_ALL", "DATA_STATE",
           "DATA_TRACERS", "EQN_SHALLOW_WATER", "EQN_PRIMITIVE_NONHYDRO"]
