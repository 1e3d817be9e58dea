"""Run under torchrun: N ranks (one GPU each) propagate + integrate with the NCCL halo and every rank compares its own points
with the 1-rank CPU oracle.  PROPAGS2 must be bit-identical to the 1-rank result (it is order-independent per point)."""
import os
os.environ.setdefault("ECWAM_B200_PROPAG", "exact")   # the N-rank = 1-rank bit-for-bit check needs the exact PROPAGS2 kernel
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import torch.distributed as dist

from common import CASES, OUT_ICE, OUT_ITG, OUT_SEA, compare_bout, make_oracle, relerr
from ecwam_b200 import lib as L, model as M, synth


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = L.load()
    buf = C.create_string_buffer(128)
    if rank == 0:
        L.check(lib.ecwam_b200_nccl_unique_id(buf), "uid")
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone().cuda()
    dist.broadcast(t, 0)
    cm = C.c_void_p()
    L.check(lib.ecwam_b200_nccl_comm_init(bytes(t.cpu().numpy().tobytes()), world, rank, C.byref(cm)), "comm_init")
    ok = True
    host_done = False
    for case, extra in (("o48like", {}), ("o640like", dict(ifrelfmax=5, delpro_lf=225.0)), ("o48like", dict(irefra=1)), ("o48like", dict(irefra=3)),
                        ("o48like", dict(llgcbz0=1, llnormagam=1, wspmin=0.3))):
        CASES["_mr"] = dict(CASES[case], N=28)
        g, o, f, fl = make_oracle("_mr", **extra)
        c = CASES["_mr"]
        s = M.WamSetup(g, nproc=world, nang=c["A"], nfre_red=c["Fr"], iphys=c["iphys"], nproma=c["nproma"], idelt=c["dt"],
                       idelpro=c["dt"], delpro_lf=extra.get("delpro_lf", c["dt"]), ifrelfmax=extra.get("ifrelfmax", 0),
                       irefra=extra.get("irefra", 0), llgcbz0=extra.get("llgcbz0", 0), llnormagam=extra.get("llnormagam", 0),
                       wspmin=extra.get("wspmin", 1.0))
        w = M.WamIntgr(s, rank, device="cuda:%d" % local, nccl_comm=cm.value)
        w.set_static(g.depth)
        for k, v in f.items():
            w.set_field(k, v)
        w.set_fl1(fl)
        if extra.get("irefra", 0) >= 2:
            from common import synthetic_currents
            uc, vc = synthetic_currents(g)
            o.set_field("UCUR", uc); o.set_field("VCUR", vc)
            w.set_field("ucur", uc); w.set_field("vcur", vc)
        assert o.propag() == 0 and w.propag() == 0
        w.synchronize()
        same = np.array_equal(w.get_spec("fl1"), o.get_fl1()[:, :, w.own])
        o.implsch(); w.implsch()
        for _ in range(2):
            o.step(); w.step()
        w.synchronize()
        e = relerr(w.get_spec("fl1"), o.get_fl1()[:, :, w.own])
        mij_ok = bool((w.get_field("mij") == o.get_field("MIJ")[w.own]).all())
        print("rank %d %s: propag bit-exact %s, FL1 rel err after 3 steps %.2e, MIJ exact %s" % (rank, case, same, e, mij_ok), flush=True)
        ok = ok and same and e < 1e-12 and mij_ok
        if extra.get("irefra", 0) >= 2:
            w.close()
            continue
        # OUTBS + WAMNORM over the ranks: the global-order norm is the reference's reproducible one (mpminmaxavg.F90:121-153)
        b = o.outbs(OUT_ITG, OUT_ICE, OUT_SEA)
        a = w.outbs(OUT_ITG, OUT_ICE, OUT_SEA)
        compare_bout(a, b[:, w.own])
        for glob in (True, False):
            wa, wb = w.outwnorm(glob), o.outwnorm(glob)
            circ = np.array([itg in (2, 5, 13, 14, 63, 22, 27, 28) for itg in OUT_ITG])
            scale = np.maximum(np.maximum(np.abs(wb[:, 1]), np.abs(wb[:, 2])), 1e-6)[:, None]
            nerr = float((np.abs(wa[:, :3] - wb[:, :3]) / scale)[~circ].max())
            cnt_ok = bool((wa[:, 3] == wb[:, 3]).all())
            print("rank %d %s: WAMNORM global=%s max rel err %.2e, counts equal %s" % (rank, case, glob, nerr, cnt_ok), flush=True)
            ok = ok and nerr < 1e-10 and cnt_ok
        if not host_done:
            # the host-buffer entry point on N ranks (send chunks up first, halo exchange, then the band pipeline) against the
            # device-resident path on a twin handle: same kernels per point, so the spectra must be identical
            host_done = True
            wt = M.WamIntgr(s, rank, device="cuda:%d" % local, nccl_comm=cm.value)
            for n_, t_ in w.t.items():
                wt.t[n_].copy_(t_)
            host = {n_: w.t[n_].cpu().pin_memory() for n_, _ in L.Fields._fields_}
            hf = L.Fields()
            for n_, _ in L.Fields._fields_:
                setattr(hf, n_, C.cast(host[n_].data_ptr(), C.POINTER(C.c_int if n_ == "mij" else C.c_double)))
            hin, hout = C.c_longlong(), C.c_longlong()
            for _ in range(3):
                wt.step()
                L.check(lib.ecwam_b200_wamintgr_host(w.h, C.byref(hf), 1, C.byref(hin), C.byref(hout)), "wamintgr_host")
            wt.synchronize()
            same_h = bool(torch.equal(host["fl1"], wt.t["fl1"].cpu()) and torch.equal(host["mij"], wt.t["mij"].cpu()) and
                          torch.equal(host["ufric"], wt.t["ufric"].cpu()))
            print("rank %d %s: host-buffer path on %d ranks identical to the device path: %s" % (rank, case, world, same_h), flush=True)
            ok = ok and same_h
            wt.close()
        w.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and flag.item() == 1:
        print("MULTIRANK OK")
    L.check(lib.ecwam_b200_nccl_comm_destroy(cm), "comm_destroy")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
