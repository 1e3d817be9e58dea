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


def _restart_worker(rank, world, port, q, tmpdir):
    """Every rank writes its own points of the shared BLS / LAW files at the same time (rank 0 lays the files out first,
    then a barrier, as scripts/run_standalone.py does), reads its points back, and rank 0 compares the files with the oracle's."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import torch.distributed as dist
    from ecwam_b200 import lib as L, synth, model as M
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = L.load()
        g = synth.make_grid(20, "continents")
        s = M.WamSetup(g, nproc=world, nang=12, nfre_red=25)
        n, A, F, NREAL = s.niblo, 12, 36, 16
        rng = np.random.default_rng(11)                     # same numbers on every rank
        fl_new, r_new = rng.random((F, A, n)), rng.normal(size=(NREAL, n))
        a, b = int(s.nstart[rank]), int(s.nend[rank])
        ij = np.ascontiguousarray(s.new2ij[a: b + 1], dtype=np.int32)
        mine, rmine = np.ascontiguousarray(fl_new[:, :, a - 1: b]), np.ascontiguousarray(r_new[:, a - 1: b])
        bls, law = os.path.join(tmpdir, "BLS").encode(), os.path.join(tmpdir, "LAW").encode()
        dates = [b"20220101060000"] * 4
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)

        def write(create):
            L.check(lib.ecwam_b200_savspec(bls, n, A, F, ij.size, ij.ctypes.data_as(ip), mine.ctypes.data_as(dp), create), "savspec")
            L.check(lib.ecwam_b200_savstress(law, *dates, n, NREAL, ij.size, ij.ctypes.data_as(ip), rmine.ctypes.data_as(dp), create), "savstress")
        if rank == 0:
            write(1)
        dist.barrier()
        if rank != 0:
            write(0)
        dist.barrier()
        got, rgot = np.empty_like(mine), np.empty_like(rmine)
        L.check(lib.ecwam_b200_getspec(bls, n, A, F, ij.size, ij.ctypes.data_as(ip), got.ctypes.data_as(dp)), "getspec")
        L.check(lib.ecwam_b200_getstress(law, None, n, NREAL, ij.size, ij.ctypes.data_as(ip), rgot.ctypes.data_as(dp)), "getstress")
        ok = np.array_equal(got, mine) and np.array_equal(rgot, rmine)
        if rank == 0:
            from oracle import restart_io as R
            ij2new = np.asarray(s.ij2new[1:], dtype=np.int64)
            R.writefl(os.path.join(tmpdir, "BLS_ref"), fl_new, ij2new)
            R.writestress(os.path.join(tmpdir, "LAW_ref"), [d.decode() for d in dates], r_new, ij2new)
            for nm in ("BLS", "LAW"):
                ok = ok and open(os.path.join(tmpdir, nm), "rb").read() == open(os.path.join(tmpdir, nm + "_ref"), "rb").read()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_concurrent_restart_writes_over_gloo(built, tmp_path, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_restart_worker, args=(r, world, port, q, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == list(range(world))
    assert all(ok for _, ok in res), res
