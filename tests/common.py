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
    # ..._O48_cy49r1.yml: gravity-capillary roughness + renormalised growth (WSPMIN = 0.3 with LLGCBZ0, userin.F90:913-918)
    "o48_cy49r1": dict(N=20, A=12, Fr=25, mask="continents", iphys=1, nproma=32, dt=900.0,
                       cfg=dict(llgcbz0=1, llnormagam=1, wspmin=0.3)),
    "o48_iphys0_gc": dict(N=20, A=12, Fr=25, mask="continents", iphys=0, nproma=24, dt=900.0,
                          cfg=dict(llgcbz0=1, llnormagam=1, wspmin=0.3)),
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
    kw.update(c.get("cfg", {}))
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
    kw.update(c.get("cfg", {}))
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

# ---- the steps either side of the hot path: NEWWIND, OUTBS/OUTBLOCK, OUTWNORM ------------------------------------------
# OUTBLOCK parameters built here (numbering of mpcrtbl.F90, NTRAIN = 3) with MPCRTBL's mask flags
# (IPRMINFO(:,6) sea-ice mask, IPRMINFO(:,7) shallow points to missing; mpcrtbl.F90:93-431)
OUT_PARAMS = {1: (1, 1), 2: (1, 1), 3: (1, 1), 4: (0, 1), 5: (0, 0), 6: (1, 1), 7: (0, 0), 8: (1, 1), 9: (1, 1), 10: (0, 0), 11: (1, 1),
              12: (1, 1), 13: (1, 1), 14: (1, 1), 15: (1, 1), 16: (1, 1), 20: (1, 1), 21: (1, 1), 22: (1, 1), 23: (1, 1),
              24: (1, 1), 25: (1, 1), 26: (1, 1), 27: (1, 1), 28: (1, 1), 32: (0, 1), 35: (1, 1), 36: (1, 1), 37: (0, 1),
              38: (0, 1), 39: (0, 1), 40: (0, 1), 41: (0, 1), 52: (1, 1), 53: (0, 0), 54: (0, 0), 55: (0, 1), 56: (0, 1), 62: (1, 1),
              63: (1, 1), 64: (1, 1), 65: (1, 1), 66: (1, 1), 67: (1, 1), 68: (1, 1), 69: (1, 1), 73: (0, 1), 74: (0, 1), 75: (0, 1), 76: (0, 1), 77: (0, 1)}
OUT_ITG = sorted(OUT_PARAMS)
OUT_ICE = [OUT_PARAMS[i][0] for i in OUT_ITG]
OUT_SEA = [OUT_PARAMS[i][1] for i in OUT_ITG]
OUT_DIRECTIONS = {2: 1, 13: 11, 14: 12, 63: 62, 5: None}    # direction column -> column of the matching height / magnitude
OUT_COPIES = (4, 10, 32, 35, 36, 37, 38, 39, 40, 41, 53, 54, 55, 56, 73, 74, 75, 76, 77)
ZMISS = -999.0


def next_forcing(f, k=1):
    """A second, different forcing set (what GETWND would deliver for the next wind step): rotated and rescaled winds."""
    n = f["WSWAVE"].size
    ramp = 0.5 + 1.2 * ((np.arange(n) * 7919 * k) % 101) / 100.0
    return dict(WSWAVE=np.maximum(f["WSWAVE"] * ramp, 0.3), WDWAVE=np.mod(f["WDWAVE"] + 0.9 * k, 2 * np.pi), AIRD=f["AIRD"] * 1.02,
                WSTAR=f["WSTAR"] * 0.5, CICOVER=np.clip(f["CICOVER"] * 1.1, 0, 1), CITHICK=f["CITHICK"] + 0.1,
                USTRA=0.01 * np.cos(np.arange(n)), VSTRA=0.01 * np.sin(np.arange(n)))


def compare_bout(a, b, itgs=OUT_ITG):
    """a (CUDA) vs b (oracle) BOUT[column, point]; the tolerances are the ones stated in tests/test_gpu_output.py."""
    worst = {}
    for i, itg in enumerate(itgs):
        x, y = a[i], b[i]
        assert np.array_equal(x == ZMISS, y == ZMISS), "parameter %d: missing-value pattern differs" % itg
        ok = y != ZMISS
        x, y = x[ok], y[ok]
        if itg in OUT_COPIES:            # copies of model fields: as close as the fields themselves (exactness vs the own
            scale = np.maximum(np.abs(y), 1e-8 * max(float(np.abs(y).max()), 1e-300))   # fields is checked by the caller)
            worst[itg] = float((np.abs(x - y) / scale).max())
            assert worst[itg] <= 1e-10, "parameter %d (copy of a field): relative difference %g" % (itg, worst[itg])
            continue
        if itg in OUT_DIRECTIONS:
            d = np.abs(x - y)
            d = np.minimum(d, 360.0 - d)
            ref = OUT_DIRECTIONS[itg]
            if ref is not None and ref in itgs:   # a mean direction is only defined where there is energy
                h = b[itgs.index(ref)][ok]
                d = d[h > 1e-3 * max(h.max(), 1e-300)]
            worst[itg] = float(d.max()) if d.size else 0.0
            assert worst[itg] <= 1e-7, "parameter %d: direction differs by %g deg" % (itg, worst[itg])
        elif itg in (22, 27, 28):        # sqrt(2 (1 - x)) amplifies rounding where the spectrum is narrow
            worst[itg] = float(np.abs(x - y).max())
            assert worst[itg] <= 1e-7, "parameter %d: spread differs by %g" % (itg, worst[itg])
        else:
            worst[itg] = float((np.abs(x - y) / np.maximum(np.abs(y), 1e-30)).max())
            assert worst[itg] <= 1e-10, "parameter %d: relative difference %g" % (itg, worst[itg])
    return worst


def synthetic_currents(g, amp=0.8):
    """Smooth surface currents with an area of exact zeros (= "no current data": gradi.F90:171-181 extrapolates no gradient
    there)."""
    lam, phi = np.deg2rad(g.lon), np.deg2rad(g.lat)
    u = amp * np.cos(phi) * np.sin(3 * lam) * (1 + 0.3 * np.cos(5 * phi))
    v = 0.5 * amp * np.sin(2 * phi) * np.cos(2 * lam)
    m = np.abs(g.lat) > 60
    u[m] = 0.0
    v[m] = 0.0
    return u, v


def synthetic_fieldg(g, seed=7, with_ws=True):
    """Forcing fields on the (NGY, NGX) wave grid itself plus IFROMIJ / JFROMIJ of every sea point (original order): what
    GETWND's blocking step (WAMWND + MICEP) reads.  Includes calm cells, cells with missing / out-of-range ice values, both
    signs of every component."""
    ngy, ngx = int(g.ngy), int(max(g.nlonrgg))
    rng = np.random.default_rng(seed)
    start = np.concatenate([[0], np.cumsum(g.nlonrgg)[:-1]])
    cell = np.flatnonzero(np.asarray(g.mask).astype(bool))
    jj = g.row_of[cell].astype(np.int32) + 1
    ii = (cell - start[g.row_of[cell]]).astype(np.int32) + 1
    f = dict(uwnd=rng.normal(0, 7, (ngy, ngx)), vwnd=rng.normal(0, 7, (ngy, ngx)), aird=1.1 + 0.2 * rng.random((ngy, ngx)),
             wstar=rng.random((ngy, ngx)), cicover=np.clip(rng.normal(0.3, 0.5, (ngy, ngx)), -0.2, 1.3), cithick=2.0 * rng.random((ngy, ngx)),
             ustra=rng.normal(0, 0.1, (ngy, ngx)), vstra=rng.normal(0, 0.1, (ngy, ngx)))
    calm = rng.random((ngy, ngx)) < 0.05
    f["uwnd"][calm] = 0.0; f["vwnd"][calm] = 0.0
    f["cicover"][rng.random((ngy, ngx)) < 0.05] = ZMISS
    if with_ws:
        f["wswave"] = np.where(rng.random((ngy, ngx)) < 0.2, 0.0, np.abs(rng.normal(8, 4, (ngy, ngx))))
        f["wswave"][rng.random((ngy, ngx)) < 0.05] = ZMISS
        f["wdwave"] = 2 * np.pi * rng.random((ngy, ngx))
    return f, ii, jj
