"""Reader/runner for the oracle dump files written by oracle/ref_dump.cpp.

TEST INFRASTRUCTURE ONLY.  A dump is a flat sequence of named records
(name, dtype, dims, raw data) holding the unmodified reference's arrays in the
reference's own layout: state [c][iA][iB][k] with a one-node halo
(reference src/base/DataArray4D.h:507-530, src/atm/GridPatch.cpp:341-357).
"""
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def read_dump(path):
    """Return {record name: ndarray} (scalars as 1-element arrays)."""
    out = {}
    with open(path, "rb") as f:
        magic = f.read(8)
        if magic != b"TB2DUMP1":
            raise ValueError("not a ref_dump file: %r" % path)
        while True:
            hdr = f.read(4)
            if len(hdr) < 4:
                break
            (nname,) = struct.unpack("<I", hdr)
            name = f.read(nname).decode()
            dtype, ndim = struct.unpack("<II", f.read(8))
            dims = struct.unpack("<%dQ" % ndim, f.read(8 * ndim))
            n = int(np.prod(dims))
            if dtype == 0:
                a = np.frombuffer(f.read(8 * n), dtype="<f8").reshape(dims)
            else:
                a = np.frombuffer(f.read(4 * n), dtype="<i4").reshape(dims)
            out[name] = a.copy()
    return out


def have_ref_dump():
    return os.path.exists(REF_DUMP) and os.access(REF_DUMP, os.X_OK)


def run_ref_dump(out_path, case, script, flags=(), npatch=6, timeout=1800):
    """Run the reference through the dump hook (oracle/_ref/ref_dump)."""
    cmd = [REF_DUMP, "--case", case, "--out", out_path, "--npatch", str(npatch),
           "--output_none", "--script", script] + list(flags)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         timeout=timeout, text=True)
    if res.returncode != 0:
        raise RuntimeError("ref_dump failed:\n" + res.stdout[-4000:])
    return read_dump(out_path)


def scalar(d, key):
    return d[key].reshape(-1)[0].item()
