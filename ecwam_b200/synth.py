"""Synthetic inputs for the WAMINTGR hot path (GRIB forcing / ETOPO1 are not available offline).

Follows SURVEY.md 8(d):
  * grid      : octahedral O<N> as share/ecwam/scripts/ecwam_grids.py:42-72,107-114
                (NGY = 2N rows, NLONRGG = 20 + 4 j, first latitude from the lat0 table)
  * mask      : "aqua"  -> all sea except the two polar rows (src/programs/preproc.F90:337-346)
                "continents" -> seeded smooth blobs, ~34 % land, with shelves shallower than 50 m
  * depth     : BATHYMAX (998.999) for aqua, clamp(5 + 4000 s, DEPTHA=2, 998.999) otherwise
  * forcing   : analytic 10 m wind, air density, convective velocity, sea-ice cover
  * cold start: JONSWAP fetch law from the local wind (mstart.F90 / peak.F90 / jonswap.F90 / spr.F90,
                namelist values of share/ecwam/scripts/ecwam_run_preset.sh:153-206) + f^-5 tail above NFRE_RED
                (getspec.F90:650-654)

Everything is returned in the ORIGINAL global sea-point order (rows south->north, west->east inside a row,
readmdlconf.F90:107-116).  This module is pure numpy host code shared by bench.py, the tests and smoke().
"""
from __future__ import annotations

import dataclasses
import numpy as np

LAT0 = {16: 85.760587120444, 24: 87.159094555863, 32: 87.863798839233, 48: 88.572168514007, 64: 88.927735352296,
        80: 89.141519426461, 96: 89.284227532514, 128: 89.462821568577, 160: 89.570089550607, 200: 89.655964246870,
        256: 89.731148618413, 320: 89.784876907219, 400: 89.827874645894, 512: 89.865508687700, 576: 89.880445682778,
        640: 89.892396445590, 800: 89.913910432567, 1024: 89.932737928460, 1280: 89.946187715666}

# spectral / time-step settings of the reference's test configurations (tests/etopo1_oper_an_fc_O*.yml)
CONFIGS = {
    "O48": dict(N=48, nang=12, nfre_red=25, idelt=900.0, idelpro=900.0, ifrelfmax=0, delpro_lf=900.0),
    "O320": dict(N=320, nang=24, nfre_red=29, idelt=900.0, idelpro=900.0, ifrelfmax=0, delpro_lf=900.0),
    "O640": dict(N=640, nang=36, nfre_red=29, idelt=450.0, idelpro=450.0, ifrelfmax=0, delpro_lf=450.0),
    "O1280": dict(N=1280, nang=36, nfre_red=29, idelt=450.0, idelpro=450.0, ifrelfmax=5, delpro_lf=225.0),
    # profiling only: the O640 spectrum (36 x 36(29)) on a small grid, so that ncu replays stay short
    "P256": dict(N=256, nang=36, nfre_red=29, idelt=450.0, idelpro=450.0, ifrelfmax=0, delpro_lf=450.0),
}


@dataclasses.dataclass
class SynthGrid:
    N: int
    ngy: int
    nlonrgg: np.ndarray      # int32 [ngy]
    amosop: float
    amonop: float
    mask: np.ndarray         # uint8 [sum(nlonrgg)], row-major south->north
    row_of: np.ndarray       # int32 per grid cell
    lon: np.ndarray          # deg, per sea point (original order)
    lat: np.ndarray          # deg, per sea point
    depth: np.ndarray        # m, per sea point
    niblo: int


def octahedral(N: int):
    if N not in LAT0:
        # any N is accepted for tests: extrapolate the first Gaussian latitude (only the value of XDELLA changes)
        lat0 = 90.0 - 90.0 / (N + 0.5) * 0.74
    else:
        lat0 = LAT0[N]
    j = np.arange(2 * N)
    nlon = 20 + 4 * np.minimum(j, 2 * N - 1 - j)
    return nlon.astype(np.int32), -lat0, lat0


def _smooth_field(lon_deg, lat_deg, seed, nmodes=24):
    """Seeded smooth scalar field on the sphere in roughly [-1, 1] (sum of low-order plane waves on the unit sphere)."""
    rng = np.random.default_rng(seed)
    lam = np.deg2rad(lon_deg)
    phi = np.deg2rad(lat_deg)
    xyz = np.stack([np.cos(phi) * np.cos(lam), np.cos(phi) * np.sin(lam), np.sin(phi)], axis=-1)
    s = np.zeros(lon_deg.shape)
    for _ in range(nmodes):
        kvec = rng.normal(size=3) * rng.uniform(1.5, 5.0)
        s += rng.uniform(0.5, 1.0) * np.cos(xyz @ kvec + rng.uniform(0, 2 * np.pi))
    return s / np.sqrt(nmodes) * 1.6


def make_grid(N: int, mask_kind: str = "aqua", seed: int = 20230101) -> SynthGrid:
    nlon, amosop, amonop = octahedral(N)
    ngy = 2 * N
    xdella = (amonop - amosop) / (ngy - 1)
    row = np.repeat(np.arange(ngy, dtype=np.int32), nlon)
    start = np.concatenate([[0], np.cumsum(nlon)[:-1]])
    icol = np.arange(nlon.sum()) - np.repeat(start, nlon)
    lon = icol * (360.0 / np.repeat(nlon, nlon))
    lat = amosop + row * xdella
    if mask_kind == "aqua":
        mask = ((row > 0) & (row < ngy - 1)).astype(np.uint8)
        depth = np.full(lon.shape, 998.999)
    elif mask_kind == "continents":
        s = _smooth_field(lon, lat, seed)
        thr = np.quantile(s, 0.34)
        mask = ((s > thr) & (row > 0) & (row < ngy - 1)).astype(np.uint8)
        depth = np.clip(5.0 + 4000.0 * (s - thr), 2.0, 998.999)
    else:
        raise ValueError(mask_kind)
    sea = mask.astype(bool)
    return SynthGrid(N=N, ngy=ngy, nlonrgg=nlon, amosop=amosop, amonop=amonop, mask=mask, row_of=row,
                     lon=lon[sea].copy(), lat=lat[sea].copy(), depth=depth[sea].copy(), niblo=int(sea.sum()))


def write_grid_tables(g: SynthGrid, path: str, imdlgrbid_g: int = 108) -> None:
    """The grid as a `wam_grid_tables` file in the reference's binary format (OUTCOM, outcom.F90:139-144): BATHY(NGX,NGY)
    with the sea-point depths and ZMISS = -999 on land."""
    import ctypes as C
    from . import lib as L
    ngy, ngx = int(g.ngy), int(max(g.nlonrgg))
    bathy = np.full((ngy, ngx), -999.0)
    start = np.concatenate([[0], np.cumsum(g.nlonrgg)[:-1]])
    cell = np.flatnonzero(np.asarray(g.mask).astype(bool))
    bathy[g.row_of[cell], cell - start[g.row_of[cell]]] = g.depth
    nl = np.ascontiguousarray(g.nlonrgg, dtype=np.int32)
    amo = np.array([0.0, g.amosop, 360.0 - 360.0 / ngx, g.amonop, (g.amonop - g.amosop) / (ngy - 1), 360.0 / ngx])
    L.check(L.load().ecwam_b200_grid_tables_write(path.encode(), imdlgrbid_g, ngx, ngy, nl.ctypes.data_as(C.POINTER(C.c_int)), 1, 1,
                                                  amo.ctypes.data_as(C.POINTER(C.c_double)), bathy.ctypes.data_as(C.POINTER(C.c_double))),
            "grid_tables_write")


def grid_from_tables(path: str) -> SynthGrid:
    """A grid read from a `wam_grid_tables` file (READPRE's binary branch, readpre.F90:262-345): rows south -> north with
    NLONRGG points each, sea where BATHY > ZMISS (mgrid.F90:72), longitudes AMOWEP + i * 360/NLONRGG (IPER = 1)."""
    import ctypes as C
    from . import lib as L
    lib = L.load()
    v = [C.c_int() for _ in range(6)]
    L.check(lib.ecwam_b200_grid_tables_read(path.encode(), C.byref(v[0]), C.byref(v[1]), C.byref(v[2]), C.byref(v[3]), None, 0, None, None,
                                            None, None, 0), "grid_tables_read")
    ngx, ngy = v[2].value, v[3].value
    nl, amo, bathy = np.empty(ngy, np.int32), np.empty(6), np.empty((ngy, ngx))
    L.check(lib.ecwam_b200_grid_tables_read(path.encode(), C.byref(v[0]), C.byref(v[1]), C.byref(v[2]), C.byref(v[3]),
                                            nl.ctypes.data_as(C.POINTER(C.c_int)), ngy, C.byref(v[4]), C.byref(v[5]),
                                            amo.ctypes.data_as(C.POINTER(C.c_double)), bathy.ctypes.data_as(C.POINTER(C.c_double)), bathy.size),
            "grid_tables_read")
    if v[4].value != 1:
        raise ValueError("only periodic (IPER = 1) grids are handled")
    amowep, amosop, amonop, xdella = amo[0], amo[1], amo[3], amo[4]
    row = np.repeat(np.arange(ngy, dtype=np.int32), nl)
    start = np.concatenate([[0], np.cumsum(nl)[:-1]])
    icol = np.arange(int(nl.sum())) - np.repeat(start, nl)
    b = bathy[row, icol]
    sea = b > -990.0
    lon = amowep + icol * (360.0 / np.repeat(nl, nl))
    lat = amosop + row * xdella
    return SynthGrid(N=ngy // 2, ngy=ngy, nlonrgg=nl.astype(np.int64), amosop=float(amosop), amonop=float(amonop), mask=sea.astype(np.uint8),
                     row_of=row, lon=lon[sea].copy(), lat=lat[sea].copy(), depth=b[sea].copy(), niblo=int(sea.sum()))


def make_forcing(g: SynthGrid, t_hours: float = 0.0, wstar_max: float = 1.5):
    """Analytic forcing fields per sea point (FORCING_FIELDS members used by IMPLSCH)."""
    lam = np.deg2rad(g.lon)
    phi = np.deg2rad(g.lat)
    rot = 2 * np.pi * t_hours / 48.0
    wswave = 2.0 + 18.0 * np.abs(np.sin(3 * phi)) * (0.5 + 0.5 * np.cos(2 * lam + rot))
    wdwave = np.mod(np.pi * (1.0 + 0.6 * np.sin(2 * phi) + 0.4 * np.cos(lam + rot)), 2 * np.pi)
    aird = np.full_like(lam, 1.225)
    wstar = wstar_max * (0.5 + 0.5 * np.sin(lam + 2 * phi))
    cicover = np.clip((np.abs(g.lat) - 65.0) / 10.0, 0.0, 1.0)
    cithick = np.zeros_like(lam)
    return dict(WSWAVE=wswave, WDWAVE=wdwave, AIRD=aird, WSTAR=wstar, CICOVER=cicover, CITHICK=cithick)


def frequencies(nfre: int, nfre_red: int):
    """FR(m) = FR1 * 1.1**(m - IFRE1), share/ecwam/scripts/ecwam_configure.sh:52-57 + mfr.F90."""
    ifre1 = 1 if nfre_red == 25 else 3
    fr1 = 4.177248e-02
    fr = np.empty(nfre)
    fr[ifre1 - 1] = fr1
    for m in range(ifre1 - 2, -1, -1):
        fr[m] = fr[m + 1] / 1.1
    for m in range(ifre1, nfre):
        fr[m] = 1.1 * fr[m - 1]
    return fr, ifre1, fr1


def jonswap_cold_start(wswave, wdwave, nang, nfre, nfre_red, fetch=50000.0, fm=0.2, alfa=0.018, gamma=3.0,
                       sa=0.07, sb=0.09, emaxdpt=None, out=None):
    """FL1[m, k, ij] cold start: MSTART(IOPTI=1) -> PEAK -> JONSWAP x SPR, then the getspec.F90 tail fill."""
    G = 9.806
    zpi = 2 * np.pi
    zpi4gm2 = zpi ** 4 / G ** 2
    fr, _, _ = frequencies(nfre, nfre_red)
    th = (np.arange(nang) + 0.5) * zpi / nang
    n = wswave.shape[0]
    u10 = wswave
    ok = u10 > 0.1e-08
    us = np.where(ok, u10, 1.0)
    gxu = G * fetch / (us * us)
    ug = G / us
    fp = 2.84 * gxu ** (-3.0 / 10.0)
    fp = np.maximum(0.13, fp)
    fp = np.minimum(fp, fm / ug)
    alphaj = np.maximum(0.033 * fp ** (2.0 / 3.0), 0.0081)
    fp = np.where(ok, fp * ug, 0.0)
    alphaj = np.where(ok, alphaj, 0.0)
    thes = wdwave
    fl = out if out is not None else np.empty((nfre, nang, n))
    st = np.cos(th[:, None] - thes[None, :])
    st = np.where(st > 0.0, (2.0 / np.pi) * st * st, 0.0)
    st = np.where(st < 0.1e-08, 0.0, st)
    good = (alphaj != 0.0) & (fp != 0.0)
    fps = np.where(good, fp, 1.0)
    for m in range(nfre):
        frh = fr[m]
        sigma = np.where(frh > fps, sb, sa)
        earg = np.minimum(0.5 * ((frh - fps) / (sigma * fps)) ** 2, 50.0)
        fjon = gamma ** np.exp(-earg)
        fmpf = np.minimum(1.25 * (fps / frh) ** 4, 50.0)
        et = np.where(good, alphaj * (1.0 / (frh ** 5 * zpi4gm2)) * np.exp(-fmpf) * fjon, 0.0)
        fl[m] = et[None, :] * st
    fr5 = fr ** 5
    for m in range(nfre_red, nfre):
        fl[m] = fl[nfre_red - 1] * (fr5[nfre_red - 1] / fr5[m])
    return fl


def jonswap_cold_start_device(w, wswave, wdwave, fetch=50000.0, fm=0.2, alfa=0.018, gamma=3.0, sa=0.07, sb=0.09):
    """Same cold start as `jonswap_cold_start`, evaluated on the GPU straight into the NPROMA-chunked FL1 of the
    WamIntgr `w` (for grids whose (NFRE, NANG, NIBLO) host array would be tens of GB).  wswave/wdwave: numpy arrays in
    the original global order."""
    torch = w.torch
    dev = w.device
    A, F, Fr, P, C = w.A, w.F, w.Fr, w.P, w.C
    G = 9.806
    zpi = 2 * np.pi
    zpi4gm2 = zpi ** 4 / G ** 2
    fr, _, _ = frequencies(F, Fr)
    th = torch.tensor((np.arange(A) + 0.5) * zpi / A, dtype=torch.float64, device=dev)
    u10 = torch.from_numpy(np.ascontiguousarray(wswave[w.src])).to(dev)
    thes = torch.from_numpy(np.ascontiguousarray(wdwave[w.src])).to(dev)
    ok = u10 > 0.1e-08
    us = torch.where(ok, u10, torch.ones_like(u10))
    gxu = G * fetch / (us * us)
    ug = G / us
    fp = 2.84 * gxu ** (-3.0 / 10.0)
    fp = torch.clamp(fp, min=0.13)
    fp = torch.minimum(fp, fm / ug)
    alphaj = torch.clamp(0.033 * fp ** (2.0 / 3.0), min=0.0081)
    fp = torch.where(ok, fp * ug, torch.zeros_like(fp))
    alphaj = torch.where(ok, alphaj, torch.zeros_like(alphaj))
    st = torch.cos(th[:, None] - thes[None, :])
    st = torch.where(st > 0.0, (2.0 / np.pi) * st * st, torch.zeros_like(st))
    st = torch.where(st < 0.1e-08, torch.zeros_like(st), st)
    good = (alphaj != 0.0) & (fp != 0.0)
    fps = torch.where(good, fp, torch.ones_like(fp))
    fl1 = w.t["fl1"]                                     # (C, F, A, P)
    fr5 = fr ** 5
    for m in range(F):
        mm = min(m, Fr - 1)
        frh = float(fr[mm])
        sigma = torch.where(frh > fps, torch.full_like(fps, sb), torch.full_like(fps, sa))
        earg = torch.clamp(0.5 * ((frh - fps) / (sigma * fps)) ** 2, max=50.0)
        fjon = gamma ** torch.exp(-earg)
        fmpf = torch.clamp(1.25 * (fps / frh) ** 4, max=50.0)
        et = torch.where(good, alphaj * (1.0 / (frh ** 5 * zpi4gm2)) * torch.exp(-fmpf) * fjon, torch.zeros_like(fps))
        x = et[None, :] * st                              # (A, C*P)
        if m >= Fr:
            x = x * float(fr5[Fr - 1] / fr5[m])
        fl1[:, m] = x.view(A, C, P).permute(1, 0, 2)
