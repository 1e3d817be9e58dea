"""Shared helpers of the test-suite: one synthetic case fed identically to the CPU oracle and to the CUDA path."""
import numpy as np

from ecwam_b200 import synth, model as M
from oracle import oracle as O

# (name, N, NANG, NFRE_RED, mask, iphys, nproma, dt) — spectral settings of the reference's configs on small grids
CASES = {
    "o48like": dict(N=20, A=12, Fr=25, mask="continents", iphys=1, nproma=32, dt=900.0),      # etopo1_oper_an_fc_O48.yml
    "o48_iphys0": dict(N=20, A=12, Fr=25, mask="continents", iphys=0, nproma=24, dt=900.0),   # ..._O48_iphys_0.yml
    "o320like": dict(N=16, A=24, Fr=29, mask="continents", iphys=1, nproma=64, dt=900.0),     # ..._O320.yml
    "o640like": dict(N=16, A=36, Fr=29, mask="continents", iphys=1, nproma=24, dt=450.0),     # ..._O640.yml
    "aqua": dict(N=16, A=12, Fr=25, mask="aqua", iphys=1, nproma=32, dt=900.0),
}


def make_grid(case, grid_hook=None):
    c = CASES[case]
    g = synth.make_grid(c["N"], c["mask"])
    if grid_hook is not None:
        grid_hook(g)
    return g


def make_oracle(case, npr=1, grid_hook=None, **extra):
    c = CASES[case]
    g = make_grid(case, grid_hook)
    kw = dict(nang=c["A"], nfre_red=c["Fr"], nproma=c["nproma"], npr=npr, iphys=c["iphys"], idelt=c["dt"], idelpro=c["dt"],
              delpro_lf=c["dt"])
    kw.update(extra)
    cfg = O.default_config(**kw)
    o = O.Oracle(cfg, g)
    f = synth.make_forcing(g)
    for k, v in f.items():
        o.set_field(k, v)
    fl = synth.jonswap_cold_start(f["WSWAVE"], f["WDWAVE"], c["A"], 36, c["Fr"])
    o.set_fl1(fl)
    return g, o, f, fl


def make_setup(case, nproc=1, grid_hook=None, **extra):
    c = CASES[case]
    g = make_grid(case, grid_hook)
    kw = dict(nang=c["A"], nfre_red=c["Fr"], iphys=c["iphys"], nproma=c["nproma"], idelt=c["dt"], idelpro=c["dt"], delpro_lf=c["dt"])
    kw.update(extra)
    return g, M.WamSetup(g, nproc=nproc, **kw)


def make_gpu(case, setup=None, grid=None, rank=0, comm=None, **extra):
    if setup is None:
        grid, setup = make_setup(case, **extra)
    c = CASES[case]
    w = M.WamIntgr(setup, rank, nccl_comm=comm)
    w.set_static(grid.depth)
    f = synth.make_forcing(grid)
    for k, v in f.items():
        w.set_field(k, v)
    fl = synth.jonswap_cold_start(f["WSWAVE"], f["WDWAVE"], c["A"], 36, c["Fr"])
    w.set_fl1(fl)
    return grid, setup, w


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


OUT_FIELDS = ("UFRIC", "TAUW", "TAUWDIR", "Z0M", "Z0B", "CHRNCK", "USTOKES", "VSTOKES", "TAUXD", "TAUYD", "TAUOCXD", "TAUOCYD",
              "TAUOC", "PHIOCD", "PHIEPS", "PHIAW")
