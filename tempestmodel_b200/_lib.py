"""ctypes binding of the C ABI declared in include/tempest_b200.h.

The product library is ``tempestmodel_b200/libtempest_b200.so`` (nvcc, sm_100a).
There is no CPU fallback: if the library is missing, or no CUDA device is
present when a context is created, an error is raised.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_int, c_int64,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIBRARY = os.path.join(_HERE, "libtempest_b200.so")

TB200_MAX_COMPONENTS = 8
EQN_SHALLOW_WATER = 1
EQN_PRIMITIVE_NONHYDRO = 2
DATA_STATE = 1
DATA_TRACERS = 2
DATA_ALL = 3
DATA_TEMPERATURE = 4
DATA_VORTICITY = 8
DATA_DIVERGENCE = 16

OP_NAMES = ["interp_n2e", "interp_e2n", "diff_n2n", "diff_n2e", "diff_e2n",
            "diff_e2e", "diffdiff_n2n", "diffdiff_e2e", "penalty_left",
            "penalty_right", "diff_n2n_zb"]

SCHEMES = {"strang": 0, "strang/kgu35": 0, "ars343": 1, "ars232": 2,
           "ars222": 3, "ars443": 4, "strang/rk4": 5, "strang/rk3": 6,
           "strang/fe": 7, "strang/ssprk53": 8, "erk": 9, "erk/kgu35": 9,
           "erk/fe": 10, "erk/rk4": 11, "erk/rk3": 12, "erk/ssprk53": 13,
           "gark2": 14, "ssp3_332": 15, "ark232": 16}


class Config(Structure):
    _fields_ = [
        ("np", c_int), ("nlev", c_int), ("vertical_order", c_int),
        ("ncomp", c_int), ("ntracers", c_int), ("ninstances", c_int),
        ("eqn_type", c_int), ("cartesian_xz", c_int),
        ("comp_on_redge", c_int * TB200_MAX_COMPONENTS),
        ("device", c_int),
        ("g", c_double), ("R", c_double), ("cp", c_double), ("cv", c_double),
        ("p0", c_double), ("omega", c_double), ("earth_radius", c_double),
        ("ztop", c_double), ("ref_length", c_double),
        ("hypervis_order", c_int),
        ("nu_scalar", c_double), ("nu_div", c_double), ("nu_vort", c_double),
        ("fully_explicit", c_int),
        ("off_centering", c_double),
    ]


GEOMETRY_FIELDS = [
    "jacobian2d", "contrametric2da", "contrametric2db", "coriolis",
    "topography", "jacobian", "jacobian_redge", "contrametrica",
    "contrametricb", "contrametricxi", "contrametrica_redge",
    "contrametricb_redge", "contrametricxi_redge", "derivr_node",
    "derivr_redge",
]


class Geometry(Structure):
    _fields_ = [(name, c_void_p) for name in GEOMETRY_FIELDS]


class JWTest(Structure):
    """tb200_jw_test: parameters of BaroclinicWaveJWTest."""
    _fields_ = [(n, c_double) for n in ("eta0", "tropopause_eta", "t0", "delta_t",
                                         "lapse_rate", "u0", "up", "pert_lon", "pert_lat",
                                         "pert_r")] \
        + [("perturbation", c_int), ("omega", c_double), ("radius", c_double)]


EXCHANGE_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_void_p, c_void_p,
                               POINTER(c_int64), POINTER(c_int64), c_int)

_SIGNATURES = {
    "tb200_create": (c_int, [POINTER(Config), POINTER(c_void_p)]),
    "tb200_destroy": (c_int, [c_void_p]),
    "tb200_last_error": (c_char_p, [c_void_p]),
    "tb200_version": (c_char_p, []),
    "tb200_set_stream": (c_int, [c_void_p, c_void_p]),
    "tb200_sync": (c_int, [c_void_p]),
    "tb200_check_errors": (c_int, [c_void_p]),
    "tb200_add_patch": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int,
                                c_double, c_double, c_int]),
    "tb200_commit_layout": (c_int, [c_void_p]),
    "tb200_set_tables": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "tb200_set_column_op": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p,
                                    c_void_p, c_void_p]),
    "tb200_upload_geometry": (c_int, [c_void_p, c_int, POINTER(Geometry)]),
    "tb200_set_node_ids": (c_int, [c_void_p, c_int, c_void_p]),
    "tb200_set_seam_transforms": (c_int, [c_void_p, c_int, c_int, c_void_p,
                                          c_void_p, c_void_p, c_void_p]),
    "tb200_build_connectivity": (c_int, [c_void_p]),
    "tb200_upload_state": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p]),
    "tb200_download_state": (c_int, [c_void_p, c_int, c_int, c_void_p,
                                     c_void_p, c_void_p, c_int]),
    "tb200_copy": (c_int, [c_void_p, c_int, c_int, c_int]),
    "tb200_lincomb": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int]),
    "tb200_zero": (c_int, [c_void_p, c_int, c_int]),
    "tb200_h_step_explicit": (c_int, [c_void_p, c_int, c_int, c_double]),
    "tb200_v_step_explicit": (c_int, [c_void_p, c_int, c_int, c_double]),
    "tb200_hv_step_explicit": (c_int, [c_void_p, c_int, c_int, c_double]),
    "tb200_hv_step_explicit_combine": (c_int, [c_void_p, c_void_p, c_int, c_int,
                                               c_int, c_double]),
    "tb200_hv_step_explicit_combine_dss": (c_int, [c_void_p, c_void_p, c_int, c_int,
                                               c_int, c_double]),
    "tb200_set_terrain_metric": (c_int, [c_void_p, c_int, c_void_p, c_void_p,
                                         c_void_p]),
    "tb200_set_vertical_coordinate": (c_int, [c_void_p, c_void_p, c_void_p]),
    "tb200_fast_path": (c_int, [c_void_p]),
    "tb200_fast_path_reason": (c_char_p, [c_void_p]),
    "tb200_fast_path_metric_error": (c_double, [c_void_p]),
    "tb200_v_step_implicit": (c_int, [c_void_p, c_int, c_int, c_double]),
    "tb200_copy_v_step_implicit": (c_int, [c_void_p, c_int, c_int, c_double]),
    "tb200_copy_v_step_implicit_diff": (c_int, [c_void_p, c_int, c_int, c_double]),
    "tb200_v_step_implicit_inc_available": (c_int, [c_void_p]),
    "tb200_v_step_implicit_inc": (c_int, [c_void_p, c_int, c_int, c_double]),
    "tb200_dss": (c_int, [c_void_p, c_int, c_int]),
    "tb200_h_step_after_subcycle": (c_int, [c_void_p, c_int, c_int, c_int,
                                            c_double]),
    "tb200_filter_negative_tracers": (c_int, [c_void_p, c_int]),
    "tb200_v_filter_negative_tracers": (c_int, [c_void_p, c_int]),
    "tb200_lincomb_v_filter": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "tb200_evaluate_geometry_cs": (c_int, [c_void_p, c_int, c_double, c_double]),
    "tb200_compute_output_fields": (c_int, [c_void_p, c_int]),
    "tb200_interpolate": (c_int, [c_void_p, c_int, c_int, c_int, c_int] + [c_void_p] * 7
                          + [c_int] + [c_void_p] * 6 + [c_int, c_void_p]),
    "tb200_debug_column_field": (c_int, [c_void_p, c_int, c_void_p]),
    "tb200_evaluate_jw_topography": (c_int, [c_void_p, c_int, POINTER(JWTest)]),
    "tb200_evaluate_jw_state": (c_int, [c_void_p, c_int, c_int, POINTER(JWTest)]),
    "tb200_upload_held_suarez": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "tb200_held_suarez": (c_int, [c_void_p, c_double]),
    "tb200_kessler": (c_int, [c_void_p, c_double]),
    "tb200_scheme_instances": (c_int, [c_int]),
    "tb200_scheme_from_name": (c_int, [c_char_p]),
    "tb200_v_step_implicit_terms_explicitly": (c_int, [c_void_p, c_int, c_int, c_double]),
    "tb200_step": (c_int, [c_void_p, c_int, c_int, c_int, c_double]),
    "tb200_upload_element_area": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "tb200_upload_rayleigh": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p]),
    "tb200_upload_reference_state": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "tb200_set_uniform_diffusion": (c_int, [c_void_p, c_double, c_double]),
    "tb200_set_vertical_discretization": (c_int, [c_void_p, c_int]),
    "tb200_set_mass_flux_on_levels": (c_int, [c_void_p, c_int]),
    "tb200_upload_state_async": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "tb200_download_state_async": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p,
                                           c_void_p, c_int]),
    "tb200_transfer_sync": (c_int, [c_void_p]),
    "tb200_host_register": (c_int, [c_void_p, c_void_p, ctypes.c_size_t]),
    "tb200_host_unregister": (c_int, [c_void_p, c_void_p]),
    "tb200_set_timing_hooks": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "tb200_checksum": (c_int, [c_void_p, c_int, c_void_p]),
    "tb200_total_energy": (c_int, [c_void_p, c_int, c_void_p]),
    "tb200_total_potential_enstrophy": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "tb200_total_vertical_momentum": (c_int, [c_void_p, c_int, c_void_p]),
    "tb200_set_exchange": (c_int, [c_void_p, c_int, c_int, EXCHANGE_FN,
                                   c_void_p]),
    "tb200_exchange_counts": (c_int, [c_void_p, c_void_p, c_void_p]),
    "tb200_peer_export": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "tb200_peer_attach": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "tb200_peer_detach": (c_int, [c_void_p]),
    "tb200_launch_count": (c_int64, [c_void_p]),
    "tb200_column_count": (c_int64, [c_void_p]),
    "tb200_fused_group_count": (c_int64, [c_void_p]),
    "tb200_debug_column_assembly": (c_int, [c_void_p, c_int, c_double, c_int,
                                            c_void_p, c_int]),
    "tb200_test_band_solve": (c_int, [c_void_p, c_int, c_int, c_int, c_int,
                                      c_void_p, c_void_p]),
}

EXPORTED_SYMBOLS = sorted(_SIGNATURES)

_cache = {}


class LibraryMissing(RuntimeError):
    pass


def load(path=None):
    """Load the C-ABI library and attach the prototypes.

    ``path`` defaults to the product library; tests of kernel logic on a
    GPU-less host pass the emulation build explicitly."""
    path = os.path.abspath(path or PRODUCT_LIBRARY)
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise LibraryMissing(
            "%s not found: build it with `python -c 'import __graft_entry__ as g;"
            " g.build()'` (nvcc, sm_100a). tempestmodel_b200 has no CPU fallback."
            % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _cache[path] = lib
    return lib
