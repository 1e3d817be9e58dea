"""world_size=2 (and 4) `gloo` test of the N>1 host logic on CPU: every rank builds its MPDECOMP tables with the product's
host builder, packs the points the plan says it must send (mpexchng.F90:120-157 order), exchanges them with
torch.distributed send/recv, and checks that each halo slot received the point it stands for."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from ecwam_b200 import synth, model as M
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = synth.make_grid(20, "continents")
        s = M.WamSetup(g, nproc=world, nang=12, nfre_red=25)
        d = s.decomp_arrays(rank)
        ijs, ijl, ninf, nsup = d["ijs"], d["ijl"], d["ninf"], d["nsup"]
        # the "field": value of a point = its global (relabelled) index
        ext = np.full(nsup + 1 - ninf + 1, -1.0)
        ext[ijs - ninf: ijl - ninf + 1] = np.arange(ijs, ijl + 1)
        ijtope = d["ijtope"].reshape(world, d["ntopemax"])          # (JH, IP) column-major -> [ip][jh]
        reqs, recvbufs = [], {}
        for q_ in range(world):
            ns, nr = int(d["ntope"][q_]), int(d["nfrompe"][q_])
            if ns:
                sb = torch.from_numpy(ext[ijtope[q_, :ns] - ninf].copy())
                reqs.append(dist.isend(sb, q_))
            if nr:
                recvbufs[q_] = torch.empty(nr, dtype=torch.float64)
                reqs.append(dist.irecv(recvbufs[q_], q_))
        for r in reqs:
            r.wait()
        for q_, buf in recvbufs.items():
            st = int(d["nijstart"][q_]) - ninf
            ext[st: st + buf.numel()] = buf.numpy()
        # every halo slot now holds a global index owned by the right rank, increasing with the slot
        halo = np.concatenate([ext[: ijs - ninf], ext[ijl - ninf + 1: nsup - ninf + 1]])
        ok = (halo > 0).all() and (np.diff(halo) > 0).all()
        owner = np.searchsorted(s.nstart, halo, side="right") - 1
        ok = ok and (owner != rank).all()
        # and the neighbour tables, mapped through the received indices, equal the 1-rank (global) tables
        s1 = M.WamSetup(g, nproc=1, nang=12, nfre_red=25)
        # s1 numbers points in the ORIGINAL order; translate through the relabelling of the world-rank setup
        d1 = s1.decomp_arrays(0)
        n1 = g.niblo
        klon_glob = d1["klon"].reshape(2, n1)
        own_orig = s.new2ij[ijs: ijl + 1]
        klon_loc = d["klon"].reshape(2, ijl - ijs + 1)
        for ic in range(2):
            ref = klon_glob[ic, own_orig - 1]                      # original numbering, n1+1 = land
            got = klon_loc[ic]
            got_glob = np.where(got == nsup + 1, -1, ext[np.minimum(got, nsup) - ninf]).astype(np.int64)
            ref_new = np.where(ref == n1 + 1, -1, s.ij2new[np.minimum(ref, n1)])
            ok = ok and np.array_equal(got_glob, ref_new)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_halo_exchange_plan_over_gloo(built, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == list(range(world))
    assert all(ok for _, ok in res), res
