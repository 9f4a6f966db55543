"""The C-ABI shared library loads and exports every symbol declared in
include/tempest_b200.h (no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

from tempestmodel_b200 import PRODUCT_LIBRARY, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tempest_b200.h")).read()
    return sorted(set(re.findall(r"\b(tb200_[a-z0-9_]+)\s*\(", text)) - {"tb200_exchange_fn"})


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.EXPORTED_SYMBOLS)


def test_product_library_exports_every_symbol():
    assert os.path.exists(PRODUCT_LIBRARY), "build with __graft_entry__.build()"
    lib = ctypes.CDLL(PRODUCT_LIBRARY)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    lib.tb200_version.restype = ctypes.c_char_p
    assert b"CUDA sm_100a" in lib.tb200_version()


def test_no_cpu_fallback_without_device():
    """On a host without a CUDA device creating a context fails loudly."""
    import torch
    if torch.cuda.is_available():
        return
    from tempestmodel_b200 import DeviceContext, TempestError
    try:
        DeviceContext(np=4, nlev=1, vertical_order=1, ncomp=3, ntracers=0,
                      ninstances=2, eqn_type=1, device=-1)
    except TempestError as exc:
        assert "no CPU fallback" in str(exc) or "CUDA" in str(exc)
    else:
        raise AssertionError("context created without a GPU")
