"""Patch decomposition over ranks: the golden Strang case on 2 ranks (3 patches
each) must reproduce the single-rank reference state.  CPU: gloo + emulation
library; GPU (needs 2 devices): nccl + product library."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(backend, nproc=2, port=29611, env=None, extra=(), case="jw_ne2_l6_strang"):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "_multirank_worker.py"),
           backend, case, "strang"] + list(extra)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, timeout=900, env=dict(os.environ, **(env or {})))
    assert res.returncode == 0, res.stdout[-3000:]
    assert "MULTIRANK worst=" in res.stdout, res.stdout[-3000:]
    return res.stdout


def test_two_ranks_gloo(emu_library):
    _run("gloo")


def test_three_ranks_gloo(emu_library):
    """Two patches per rank: every rank exchanges with two neighbours, so the
    per-source slot offsets of the receive buffer are non-trivial."""
    _run("gloo", nproc=3, port=29617)


def test_two_ranks_gloo_tracers(emu_library):
    """Tracers on two ranks: state and tracer rows cross in one exchange per DSS
    (three Strang steps of the golden tracer case)."""
    _run("gloo", port=29619, case="jwtr_ne2_l6_strang", env={"TB_WORKER_STEPS": "3"})


def test_two_ranks_gloo_order5(emu_library):
    """np = 5 on two ranks: pack / exchange / unpack of the halo nodes with node
    groups that are not the 16-node elements of the column-constant path."""
    _run("gloo", port=29621, case="jw_ne2_l6_np5")


def test_two_ranks_gloo_finite_volume_order2(emu_library):
    """--vdisc FV --vertorder 2 on two ranks (general kernels, 12 levels)."""
    _run("gloo", port=29623, case="jw_ne2_l12_fv2")


def test_two_ranks_gloo_overlap(emu_library):
    """Element-list launches of the persistent kernels (exchange-feeding elements
    first, the rest on the second stream): same state."""
    _run("gloo", port=29613, env={"TB200_OVERLAP": "1"})


@pytest.mark.gpu
def test_two_ranks_nccl(cuda_library):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run("nccl", port=29612)


@pytest.mark.gpu
def test_two_ranks_nccl_overlap(cuda_library):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run("nccl", port=29614, env={"TB200_OVERLAP": "1"})


@pytest.mark.gpu
@pytest.mark.parametrize("nproc", [2, 3])
def test_ranks_peer_memory_exchange(cuda_library, nproc):
    """Halo exchange by direct stores into the peers' receive buffers (CUDA IPC
    over NVLink) instead of the all-to-all callback: same state."""
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    _run("nccl", nproc=nproc, port=29615 + nproc, extra=["peer"])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_two_ranks_tracers_gpus(cuda_library, mode):
    """Tracers on two GPUs, peer-memory exchange and NCCL callback."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run("nccl", port=29621 + (1 if mode == "nccl" else 0), case="jwtr_ne2_l6_strang",
         env={"TB_WORKER_STEPS": "3"}, extra=["peer"] if mode == "peer" else [])


def _states(tag):
    import glob
    import numpy as np
    out = {}
    for f in glob.glob("/tmp/tb200_mr_%s.rank*.npz" % tag):
        with np.load(f) as z:
            for k in z.files:
                out[k] = z[k]
        os.remove(f)
    return out


def test_ranks_do_not_change_the_bits(emu_library):
    """The same 24 patches on one rank and on four: the state after two Strang
    steps is the same bit for bit (members of an averaging group are ordered by
    their position on the panel, remote members enter the average exactly as
    local ones do, every rank solves its own copy of a shared column from
    identical inputs).  What does change the last bit is the decomposition
    itself - which copy of a shared column a patch solves - see bench.py."""
    import numpy as np
    res = []
    for nproc, port in ((1, 29651), (4, 29652)):
        _run("gloo", nproc=nproc, port=port, case="jw_ne4_l30_p24",
             env={"TB_WORKER_DUMP": "/tmp/tb200_mr_w%d" % nproc})
        res.append(_states("w%d" % nproc))
    assert len(res[0]) == 48 and set(res[0]) == set(res[1])
    for k in res[0]:
        assert np.array_equal(res[0][k], res[1][k]), k


@pytest.mark.gpu
def test_ranks_do_not_change_the_bits_gpus(cuda_library):
    """The same on GPUs: one device against four (peer-memory exchange)."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    res = []
    for nproc, port in ((1, 29653), (4, 29654)):
        _run("nccl", nproc=nproc, port=port, case="jw_ne4_l30_p24",
             extra=["peer"] if nproc > 1 else [],
             env={"TB_WORKER_DUMP": "/tmp/tb200_mr_g%d" % nproc})
        res.append(_states("g%d" % nproc))
    assert len(res[0]) == 48 and set(res[0]) == set(res[1])
    for k in res[0]:
        assert np.array_equal(res[0][k], res[1][k]), k


# ---- the decomposition the multi-GPU bench lines use: 24 patches, L = 30 -------

@pytest.mark.parametrize("nproc", [4, 8])
def test_24_patches_gloo(emu_library, nproc):
    """ne = 4, L = 30 on 24 patches over 4 and 8 ranks (6 and 3 patches each):
    every rank has up to seven neighbours, cube corners are shared by three
    ranks; state after two Strang steps against the single-rank reference."""
    _run("gloo", nproc=nproc, port=29630 + nproc, case="jw_ne4_l30_p24")


@pytest.mark.gpu
@pytest.mark.parametrize("nproc,mode", [(2, "peer"), (4, "peer"), (8, "peer"), (8, "nccl")])
def test_24_patches_gpus(cuda_library, nproc, mode):
    """The same on 2, 4 and 8 GPUs (peer-memory exchange, and the NCCL callback
    at 8): the configuration SCALE measures."""
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    _run("nccl", nproc=nproc, port=29640 + nproc + (1 if mode == "nccl" else 0),
         extra=["peer"] if mode == "peer" else [], case="jw_ne4_l30_p24")
