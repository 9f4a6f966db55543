"""Patch decomposition over ranks and the halo-exchange callback
(one process per GPU; torch.distributed is plumbing only).

The reference distributes patch n to rank n % size (Grid::DistributePatches,
reference src/atm/Grid.cpp:1038-1062), which scatters the patches of a panel;
here patches that share edges are kept together.  Per DSS the
library packs the shared nodes into one device buffer per destination rank
and calls back; the callback issues one all-to-all on the same stream.
"""
import ctypes

import numpy as np


def assign_patches(npatch, nranks):
    """owner rank of every patch: contiguous blocks of patch indices, i.e.
    whole panels or neighbouring quadrants of a panel stay on one rank."""
    if npatch % nranks != 0:
        raise ValueError("patch count %d not divisible by %d ranks" % (npatch, nranks))
    per = npatch // nranks
    return [n // per for n in range(npatch)]


class _DevicePointer:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {
            "shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}


class Exchange:
    """Callable handed to DeviceContext.set_exchange.  `cuda` selects device
    (NCCL) or host (gloo, emulation tests) buffers."""

    def __init__(self, cuda=True):
        self.cuda = cuda
        self.calls = 0
        self.bytes_sent = 0
        self._cache = {}

    def _wrap(self, ptr, n):
        import torch
        key = (ptr, n)
        t = self._cache.get(key)
        if t is None:
            if n == 0:
                t = torch.empty(0, dtype=torch.float64,
                                device="cuda" if self.cuda else "cpu")
            elif self.cuda:
                t = torch.as_tensor(_DevicePointer(ptr, n), device="cuda")
            else:
                buf = (ctypes.c_double * n).from_address(ptr)
                t = torch.from_numpy(np.frombuffer(buf, dtype=np.float64))
            self._cache[key] = t
        return t

    def __call__(self, user, sendbuf, recvbuf, send_counts, recv_counts, nranks):
        import torch.distributed as dist
        try:
            sc = [int(send_counts[r]) for r in range(nranks)]
            rc = [int(recv_counts[r]) for r in range(nranks)]
            send = self._wrap(sendbuf or 0, sum(sc))
            recv = self._wrap(recvbuf or 0, sum(rc))
            if self.cuda:
                dist.all_to_all_single(recv, send, rc, sc)
            else:
                # gloo: pairwise exchange
                reqs = []
                so = ro = 0
                me = dist.get_rank()
                for r in range(nranks):
                    if r != me and rc[r] > 0:
                        reqs.append(dist.irecv(recv[ro:ro + rc[r]], src=r))
                    ro += rc[r]
                for r in range(nranks):
                    if r != me and sc[r] > 0:
                        reqs.append(dist.isend(send[so:so + sc[r]].clone(), dst=r))
                    so += sc[r]
                for q in reqs:
                    q.wait()
            self.calls += 1
            self.bytes_sent += 8 * sum(sc)
            return 0
        except Exception as exc:  # surfaced by the library as an error code
            import traceback
            traceback.print_exc()
            self.error = exc
            return 1


def enable_peer_exchange(ctx, rank, nranks):
    """Switch the halo exchange of `ctx` from the callback to direct stores
    into the peers' receive buffers (CUDA IPC over NVLink / NVSwitch; all ranks
    on one node).  torch.distributed only carries the 64-byte handles and the
    slot offsets at set-up.  Collective: every rank must call it after
    build_connectivity.  Returns False (and leaves the callback in place) when
    any rank could not map its peers."""
    import torch
    import torch.distributed as dist
    ok = 1
    try:
        mine = ctx.peer_export(nranks)
    except Exception:
        mine, ok = None, 0
    gathered = [None] * nranks
    dist.all_gather_object(gathered, mine)
    if ok and all(g is not None for g in gathered):
        try:
            ctx.peer_attach([g[0] for g in gathered],
                            [g[1][rank] for g in gathered],
                            [g[2] for g in gathered])
        except Exception:
            ok = 0
    else:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32,
                        device="cuda" if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        if ok:
            ctx.peer_detach()
        return False
    return True
