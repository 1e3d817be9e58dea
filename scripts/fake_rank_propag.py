"""Developer probe: one process plays rank R of an NPROC-rank MPDECOMP on one GPU (the halo exchange is a no-op callback: the halo
stays zero), to time / profile the per-rank kernels of a decomposed run without a second GPU.
usage: fake_rank_propag.py <workload> <nproc> <rank> [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from ecwam_b200 import lib as L, model as M, synth
import bench as B

wl, nproc, rank = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
cfgw, nproma = B.workload_cfg(wl)
g = synth.make_grid(cfgw["N"], "continents")
s = M.WamSetup(g, nproc=nproc, nang=cfgw["nang"], nfre_red=cfgw["nfre_red"], iphys=1, nproma=nproma, idelt=cfgw["idelt"],
               idelpro=cfgw["idelpro"], delpro_lf=cfgw["delpro_lf"], ifrelfmax=cfgw["ifrelfmax"])
w = M.WamIntgr(s, rank, device="cuda:0", nccl_comm=None)
lib = w.lib
noop = L.ExchangeFn(lambda *a: 0)
L.check(lib.ecwam_b200_set_exchange(w.h, noop, None, 0), "set_exchange")
w.set_static(g.depth)
f = synth.make_forcing(g)
for k, v in f.items():
    w.set_field(k, v)
synth.jonswap_cold_start_device(w, f["WSWAVE"], f["WDWAVE"])
for mode in (os.environ.get("MODES", "exact fast").split()):
    os.environ["ECWAM_B200_PROPAG"] = mode
    for _ in range(2):
        assert w.propag() == 0
    lib.ecwam_b200_timing_reset(w.h); lib.ecwam_b200_timing_enable(w.h, 1)
    for _ in range(steps):
        assert w.propag() == 0
    w.synchronize()
    ms, cnt = C.c_double(), C.c_longlong()
    lib.ecwam_b200_timing_get(w.h, b"propags2", C.byref(ms), C.byref(cnt))
    print("rank %d/%d of %s (%d own points): PROPAG=%s propags2 %.3f ms per launch (%d launches)" % (rank, nproc, wl, w.nloc, mode, ms.value / max(cnt.value, 1), cnt.value), flush=True)
    lib.ecwam_b200_timing_enable(w.h, 0)
