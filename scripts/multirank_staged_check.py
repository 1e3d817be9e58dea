"""Run under torchrun with the gloo backend: N ranks that may SHARE one GPU (the driver's 1-GPU test box) propagate + integrate
with the halo exchange supplied by the host through ecwam_b200_set_exchange (staged through pinned host buffers, moved with
torch.distributed send / recv over gloo) -- the shape of the reference's own MPEXCHNG over MPI (mpexchng.F90:164-206).  Every rank
compares its own points with the 1-rank CPU oracle; PROPAGS2 must be bit-identical to the 1-rank result."""
import os
os.environ.setdefault("ECWAM_B200_PROPAG", "exact")   # the N-rank = 1-rank bit-for-bit check needs the exact PROPAGS2 kernel
import ctypes as C
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import torch.distributed as dist

from common import CASES, make_oracle, relerr
from ecwam_b200 import lib as L, model as M


def make_exchange(rank):
    """ecwam_b200_exchange_fn over gloo: non-blocking receives first, then the sends, in rank order (one message per pair and call)."""
    def fn(user, nproc, sendbuf, so, sc, recvbuf, ro, rc):
        try:
            reqs, keep = [], []
            for q in range(nproc):
                if rc[q] > 0:
                    a = np.ctypeslib.as_array(C.cast(recvbuf + 8 * ro[q], C.POINTER(C.c_double)), shape=(rc[q],))
                    t = torch.from_numpy(a)
                    keep.append(t)
                    reqs.append(dist.irecv(t, src=q))
            for q in range(nproc):
                if sc[q] > 0:
                    a = np.ctypeslib.as_array(C.cast(sendbuf + 8 * so[q], C.POINTER(C.c_double)), shape=(sc[q],))
                    reqs.append(dist.isend(torch.from_numpy(a), dst=q))
            for r in reqs:
                r.wait()
            return 0
        except Exception as e:      # never let an exception cross the C boundary
            print("exchange callback failed on rank %d: %r" % (rank, e), flush=True)
            return 1
    return L.ExchangeFn(fn)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")
    lib = L.load()
    cb = make_exchange(rank)
    ok = True
    for case, extra in (("o48like", {}), ("o640like", dict(ifrelfmax=5, delpro_lf=225.0)), ("o48like", dict(irefra=1))):
        CASES["_mr"] = dict(CASES[case], N=28)
        g, o, f, fl = make_oracle("_mr", **extra)
        c = CASES["_mr"]
        s = M.WamSetup(g, nproc=world, nang=c["A"], nfre_red=c["Fr"], iphys=c["iphys"], nproma=c["nproma"], idelt=c["dt"],
                       idelpro=c["dt"], delpro_lf=extra.get("delpro_lf", c["dt"]), ifrelfmax=extra.get("ifrelfmax", 0),
                       irefra=extra.get("irefra", 0))
        w = M.WamIntgr(s, rank, device="cuda:%d" % dev, nccl_comm=None)
        L.check(lib.ecwam_b200_set_exchange(w.h, cb, None, 1), "set_exchange")
        w.set_static(g.depth)
        for k, v in f.items():
            w.set_field(k, v)
        w.set_fl1(fl)
        assert o.propag() == 0 and w.propag() == 0
        w.synchronize()
        same = np.array_equal(w.get_spec("fl1"), o.get_fl1()[:, :, w.own])
        o.implsch(); w.implsch()
        for _ in range(2):
            o.step(); w.step()
        w.synchronize()
        e = relerr(w.get_spec("fl1"), o.get_fl1()[:, :, w.own])
        mij_ok = bool((w.get_field("mij") == o.get_field("MIJ")[w.own]).all())
        print("rank %d/%d on cuda:%d %s %s: propag bit-exact %s, FL1 rel err after 3 steps %.2e, MIJ exact %s" % (
            rank, world, dev, case, extra, same, e, mij_ok), flush=True)
        ok = ok and same and e < 1e-12 and mij_ok
        w.close()
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIRANK STAGED OK" if int(flag.item()) == 1 else "MULTIRANK STAGED FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
