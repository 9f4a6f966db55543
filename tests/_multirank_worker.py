"""Worker of test_multirank.py: one rank of a 2-rank run of the golden
Strang case with patches split over ranks; backend gloo + emulation library on
a CPU host, nccl + product library on GPUs."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cases  # noqa: E402
import dumpctx  # noqa: E402
from tempestmodel_b200 import PRODUCT_LIBRARY  # noqa: E402
from tempestmodel_b200.parallel import Exchange, assign_patches  # noqa: E402


def main():
    backend = sys.argv[1]
    case = sys.argv[2]
    scheme = sys.argv[3]
    cuda = backend == "nccl"
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group(backend)
    d = cases.load_case(case)
    npatch = dumpctx.S(d, "grid.npatch")
    owners = assign_patches(npatch, world)
    ex = Exchange(cuda=cuda)
    ctx = dumpctx.context_from_dump(
        d, library=PRODUCT_LIBRARY if cuda else dumpctx.EMU_LIBRARY,
        owners=owners, rank=rank, nranks=world, exchange=ex)
    peer = len(sys.argv) > 4 and sys.argv[4] == "peer"
    if peer:
        from tempestmodel_b200.parallel import enable_peer_exchange
        assert enable_peer_exchange(ctx, rank, world), "peer-memory exchange unavailable"
    dumpctx.upload_tag(ctx, d, "ic")
    for m in range(1, ctx.cfg.ninstances):
        ctx.copy(0, m)
    nsteps = int(os.environ.get("TB_WORKER_STEPS", "2"))
    for s in range(nsteps):
        ctx.step(scheme, s == 0, False, 200.0)
    ctx.check_errors()
    errs = dumpctx.compare(ctx, d, 0, "st", [0, 1, 2, 4], [3])
    worst = max(errs.values())
    if dumpctx.S(d, "grid.ntracers") > 0:
        # tracers travel in the same exchange as the state (one DSS pass)
        terrs = dumpctx.compare_tracers(ctx, d, 0, "st")
        assert max(terrs.values()) < 1e-9, terrs
        errs.update(terrs)
    dump = os.environ.get("TB_WORKER_DUMP")
    if dump:
        # raw state of the local patches, for bit-for-bit comparisons between runs
        # on different numbers of ranks
        got = dumpctx.download(ctx, d, 0)
        np.savez(dump + ".rank%d.npz" % rank,
                 **{"p%d.%s" % (n, loc): got[n][q] for n in got for q, loc in ((0, "node"), (1, "redge"))})
    t = torch.tensor([worst], dtype=torch.float64, device="cuda" if cuda else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    s, r = ctx.exchange_counts(world)
    if rank == 0:
        print("MULTIRANK worst=%.3e exchanges=%d send_nodes=%s" % (t.item(), ex.calls, s.tolist()))
    if world > 1:
        assert s.sum() > 0
        # peer-memory exchange: the callback is never used
        assert (ex.calls == 0) if peer else (ex.calls > 0)
    assert t.item() < 1e-10, errs
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
