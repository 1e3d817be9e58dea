"""Host-side mirror of the WAMINTGR call sequence (src/ecwam/wamintgr.F90:94-146) on top of the C ABI.

`WamSetup` does what INITMDL / MPDECOMP do once (tables, decomposition, neighbour tables — host C++ inside the
library), `WamIntgr` owns one rank's NPROMA-chunked fields on one GPU (torch tensors = device memory only) and calls
PROPAG_WAM / IMPLSCH through the library.  Per-point data cross this module in the ORIGINAL global sea-point
order (rows south->north, west->east), exactly like oracle/oracle.py, so the parity tests can feed both sides
the same arrays.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import lib as L

# the members of ecwam_b200_fields by shape class
F4 = ("fl1", "xllws")
F3 = ("wavnum", "cinv", "cgroup", "xk2cg", "omosnh2kd", "stokfac", "ciwa")
F2 = ("depth", "emaxdpt", "dellam1", "cosphm1", "ucur", "vcur", "aird", "wdwave", "cicover", "wswave", "wstar", "ustra",
      "vstra", "ufric", "tauw", "tauwdir", "z0m", "z0b", "chrnck", "cithick", "wsemean", "wsfmean", "ustokes", "vstokes",
      "strnms", "tauxd", "tauyd", "tauocxd", "tauocyd", "tauoc", "tauicx", "tauicy", "phiocd", "phieps", "phiaw")
FI = ("mij",)

# OUTBLOCK parameters built by ecwam_b200_outbs -> (sea-ice mask, shallow-to-missing) flags of MPCRTBL's DEFINE_PARAMETER
# calls (mpcrtbl.F90:93-431, numbering with NTRAIN = 3)
OUTBLOCK_PARAMS = {1: (1, 1), 2: (1, 1), 3: (1, 1), 4: (0, 1), 5: (0, 0), 6: (1, 1), 7: (0, 0), 8: (1, 1), 9: (1, 1), 10: (0, 0), 11: (1, 1),
                   12: (1, 1), 13: (1, 1), 14: (1, 1), 15: (1, 1), 16: (1, 1), 20: (1, 1), 21: (1, 1), 22: (1, 1), 23: (1, 1),
                   24: (1, 1), 25: (1, 1), 26: (1, 1), 27: (1, 1), 28: (1, 1), 32: (0, 1), 35: (1, 1), 36: (1, 1), 37: (0, 1),
                   38: (0, 1), 39: (0, 1), 40: (0, 1), 41: (0, 1), 52: (1, 1), 53: (0, 0), 54: (0, 0), 55: (0, 1), 56: (0, 1), 62: (1, 1),
                   63: (1, 1), 64: (1, 1), 65: (1, 1), 66: (1, 1), 67: (1, 1), 68: (1, 1), 69: (1, 1), 73: (0, 1), 74: (0, 1), 75: (0, 1), 76: (0, 1), 77: (0, 1)}


def default_params(**kw) -> L.Params:
    """Namelist values of the reference's test configurations."""
    p = L.Params(nang=12, nfre=36, nfre_red=25, iphys=1, isnonlin=0, idamping=1, irefra=0, icase=1, llgcbz0=0, llnormagam=0,
                 llcapchnk=1, lbiwbk=1, licerun=1, lmaskice=1, lwamrsetci=1, lciwa=0, lwflux=0, lwfluxout=1, lwnemocou=0,
                 lwvflx_snl=1, lwcouast=0, icode_wnd=3, ifrelfmax=0, nproma=32, nchnk=0, idelt=900.0, idelpro=900.0,
                 delpro_lf=900.0, ximp=1.0, rnu=1.5e-5, rnum=0.11 * 1.5e-5, wspmin=1.0, cithrsh=0.3, cithrsh_tail=0.3,
                 ciblock=0.0, flmin=1e-5, bathymax=998.999, llcflcuroff=1, zalpfacx=1.0, zalpfacb=1.0, cdicwa=0.01, lwnemotauoc=0, lwnemocoustk=0,
                 lwnemocoustrn=0, lwnemocousend=1, lwcou=0, lwnemocouwrs=0, lwnemocouibr=0, zalpwrs=1.0, zibrw_thrsh=0.5)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def _np(ptr, n, dtype):
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class WamSetup:
    """Tables + grid + decomposition for `nproc` ranks (host only; no GPU needed)."""

    def __init__(self, grid, nproc: int = 1, ll1d: bool = False, **params):
        self.lib = L.load()
        self.grid = grid
        self.nproc = nproc
        self.par = default_params(**params)
        self.ifre1 = 1 if self.par.nfre_red == 25 else 3      # share/ecwam/scripts/ecwam_configure.sh:52-57
        self.fr1 = 4.177248e-02
        h = C.c_void_p()
        L.check(self.lib.ecwam_b200_host_tables_create(C.byref(self.par), self.ifre1, self.fr1, C.byref(h)), "host_tables_create")
        self._ht = h
        self.tables = self.lib.ecwam_b200_host_tables_get(h).contents
        nl = np.ascontiguousarray(grid.nlonrgg, dtype=np.int32)
        mk = np.ascontiguousarray(grid.mask, dtype=np.uint8)
        g = C.c_void_p()
        L.check(self.lib.ecwam_b200_host_grid_create(int(grid.ngy), nl.ctypes.data_as(C.POINTER(C.c_int)), float(grid.amosop),
                                                     float(grid.amonop), mk.tobytes(), nproc, int(ll1d), C.byref(g)),
                "host_grid_create")
        self._hg = g
        self.niblo = self.lib.ecwam_b200_host_grid_niblo(g)
        assert self.niblo == grid.niblo
        self.ij2new = _np(self.lib.ecwam_b200_host_grid_ij2newij(g), self.niblo + 1, np.int64)
        self.new2ij = _np(self.lib.ecwam_b200_host_grid_newij2ij(g), self.niblo + 1, np.int64)
        self.kxlt = _np(self.lib.ecwam_b200_host_grid_kxlt(g), self.niblo, np.int64)
        self.nstart = _np(self.lib.ecwam_b200_host_grid_nstart(g), nproc, np.int64)
        self.nend = _np(self.lib.ecwam_b200_host_grid_nend(g), nproc, np.int64)
        # land-point group velocity = deep-water value at BATHYMAX (initdpthflds.F90:80-88)
        deep = np.array([self.par.bathymax])
        cg = np.empty(self.par.nfre)
        dpp = C.POINTER(C.c_double)
        L.check(self.lib.ecwam_b200_host_depthprpt(C.byref(self.tables), self.par.nfre, 1, deep.ctypes.data_as(dpp), None, None,
                                                   cg.ctypes.data_as(dpp), None, None, None), "depthprpt")
        self.land_cgroup = np.ascontiguousarray(cg[: self.par.nfre_red])

    def table(self, name, n):
        return _np(getattr(self.tables, name), n, np.float64)

    def itable(self, name, n):
        return _np(getattr(self.tables, name), n, np.int32)

    def decomp(self, rank: int) -> L.Decomp:
        """Decomposition tables of 0-based `rank` (a copy of the struct with land_cgroup filled in)."""
        d = self.lib.ecwam_b200_host_grid_decomp(self._hg, rank + 1).contents
        out = L.Decomp()
        C.memmove(C.byref(out), C.byref(d), C.sizeof(L.Decomp))
        out.land_cgroup = self.land_cgroup.ctypes.data_as(C.POINTER(C.c_double))
        if getattr(self, "obstructions", None) is not None:      # LSUBGRID: own-point slices in the layout (IJS:IJL, NFRE_RED, n)
            own = self.new2ij[int(self.nstart[rank]): int(self.nend[rank]) + 1] - 1
            keep = [np.ascontiguousarray(x[:, :, own], dtype=np.float64) for x in self.obstructions]
            self._obs_keep = getattr(self, "_obs_keep", {})
            self._obs_keep[rank] = keep
            dpp = C.POINTER(C.c_double)
            out.obslon, out.obslat, out.obscor = (k.ctypes.data_as(dpp) for k in keep)
        return out

    def set_obstructions(self, lon, lat, cor):
        """Sub-grid obstruction coefficients (LSUBGRID = T) in the original point order: OBSLON / OBSLAT [ic, m, ij] (2, NFRE_RED, NIBLO),
        OBSCOR (4, NFRE_RED, NIBLO).  Call before the WamIntgr objects are created."""
        fr, n = self.par.nfre_red, self.niblo
        assert lon.shape == (2, fr, n) and lat.shape == (2, fr, n) and cor.shape == (4, fr, n)
        self.obstructions = (lon, lat, cor)

    def decomp_arrays(self, rank: int):
        d = self.decomp(rank)
        n = d.ijl - d.ijs + 1
        return dict(ijs=d.ijs, ijl=d.ijl, ninf=d.ninf, nsup=d.nsup, ntopemax=d.ntopemax,
                    klat=_np(d.klat, 4 * n, np.int32), klon=_np(d.klon, 2 * n, np.int32), kcor=_np(d.kcor, 8 * n, np.int32),
                    wlat=_np(d.wlat, 2 * n, np.float64), wcor=_np(d.wcor, 4 * n, np.float64),
                    nfrompe=_np(d.nfrompe, self.nproc, np.int32), ntope=_np(d.ntope, self.nproc, np.int32),
                    nijstart=_np(d.nijstart, self.nproc, np.int32), ijtope=_np(d.ijtope, d.ntopemax * self.nproc, np.int32))

    def close(self):
        if getattr(self, "_ht", None):
            self.lib.ecwam_b200_host_tables_free(self._ht)
            self._ht = None
        if getattr(self, "_hg", None):
            self.lib.ecwam_b200_host_grid_free(self._hg)
            self._hg = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class WamClock:
    """The dates WAMODEL / WAMINTGR step through (wamodel.F90:181-185, 228-233, 283-300; wamintgr.F90:92-186), in seconds since
    the start of the run instead of CHARACTER(14) dates: CDATE, CDTPRA / CDTPRO (start / end of the running propagation step),
    CDTIMP / CDTIMPNEXT (source-term integration), CDATEWH (date of the next forcing fields), CDATEWO (its value at the last
    source-term integration, to which WAMODEL resets CDATEWH at every advection step).  tests/test_reference_golden.py replays
    this bookkeeping against WAMODEL + WAMINTGR + NEWWIND executed from their own source."""

    def __init__(self, idelpro, idelt, idelwo):
        self.idelpro, self.idelt, self.idelwo = int(idelpro), int(idelt), int(idelwo)
        self.cdtpro = 0
        self.cdtpra = self.cdate = 0
        self.cdtimp = 0                       # wamodel.F90:185
        self.cdtimpnext = self.idelt          # wamodel.F90:182-183
        self.cdatewh = self.cdatewo = self.idelwo   # CDATEWO: the forcing date as of the last source-term integration (wamintgr.F90:181)


class WamIntgr:
    """One rank of the WAMINTGR hot path on one GPU."""

    def __init__(self, setup: WamSetup, rank: int = 0, device="cuda:0", nccl_comm=None, alloc_fields=True):
        import torch
        self.torch = torch
        self.s = setup
        self.lib = setup.lib
        self.rank = rank
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.EcwamError("the WAMINTGR hot path is CUDA only (no CPU fallback)")
        torch.cuda.set_device(self.device)
        self.ijs, self.ijl = int(setup.nstart[rank]), int(setup.nend[rank])
        self.nloc = self.ijl - self.ijs + 1
        P = min(setup.par.nproma, self.nloc)                  # mpdecomp.F90:1343-1356
        self.P = P
        self.C = (self.nloc + P - 1) // P                     # mchunk.F90:33-75
        self.par = L.Params()
        C.memmove(C.byref(self.par), C.byref(setup.par), C.sizeof(L.Params))
        self.par.nproma, self.par.nchnk = P, self.C
        self.A, self.F, self.Fr = self.par.nang, self.par.nfre, self.par.nfre_red
        # chunk slot -> original global sea-point index (padded lanes of the last chunk copy its first point)
        slot = np.arange(P * self.C)
        slot = np.where(slot < self.nloc, slot, (slot // P) * P)
        self.src = setup.new2ij[self.ijs + slot] - 1            # 0-based original index per slot
        self.own = setup.new2ij[self.ijs + np.arange(self.nloc)] - 1
        self.comm = nccl_comm
        self._dec = setup.decomp(rank)
        h = C.c_void_p()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        L.check(self.lib.ecwam_b200_create(C.byref(self.par), C.byref(setup.tables), C.byref(self._dec),
                                           C.c_void_p(nccl_comm) if nccl_comm else None, C.c_void_p(stream), C.byref(h)), "create")
        self.h = h
        self.t = {}
        if alloc_fields:
            self._alloc()

    # ---- field storage (device memory through torch)
    def _alloc(self):
        torch = self.torch
        P, Cn, A, F = self.P, self.C, self.A, self.F
        dev = self.device
        for n in F4:
            self.t[n] = torch.zeros((Cn, F, A, P), dtype=torch.float64, device=dev)
        for n in F3:
            self.t[n] = torch.zeros((Cn, F, P), dtype=torch.float64, device=dev)
        for n in F2:
            self.t[n] = torch.zeros((Cn, P), dtype=torch.float64, device=dev)
        self.t["mij"] = torch.full((Cn, P), F, dtype=torch.int32, device=dev)
        self.t["ciwa"].fill_(1.0)
        self.bind()
        if self.par.lwnemocou or self.par.lwnemocouibr:      # WAVE2OCEAN fields of the NEMO coupling (accumulators start at 0) + IBRMEM
            nf = L.NemoFields()
            for n, _ in L.NemoFields._fields_:
                self.t[n] = torch.zeros((Cn, P), dtype=torch.float64, device=dev)
                setattr(nf, n, C.cast(self.t[n].data_ptr(), C.POINTER(C.c_double)))
            self.nemo_fields = nf
            L.check(self.lib.ecwam_b200_bind_nemo(self.h, C.byref(nf)), "bind_nemo")

    def bind(self):
        f = L.Fields()
        for n, _ in L.Fields._fields_:
            ptr = self.t[n].data_ptr()
            setattr(f, n, C.cast(ptr, C.POINTER(C.c_int if n == "mij" else C.c_double)))
        self.fields = f
        L.check(self.lib.ecwam_b200_bind_fields(self.h, C.byref(f)), "bind_fields")

    def set_static(self, depth_global: np.ndarray):
        """Depth-derived fields: DEPTH, EMAXDPT, DELLAM1, COSPHM1 and DEPTHPRPT (initdpthflds.F90:57-88)."""
        torch = self.torch
        s = self.s
        d = np.ascontiguousarray(depth_global[self.src])
        ky = s.kxlt[self.ijs - 1 + np.where(np.arange(self.P * self.C) < self.nloc, np.arange(self.P * self.C),
                                           (np.arange(self.P * self.C) // self.P) * self.P)]
        zd = s.table_grid("zdello")[ky - 1]
        cosph = s.table_grid("cosph")[ky - 1]
        circ = s.tables.circ
        gam = np.where(d < 4.0, 0.8 * d / 4.0, 0.8)
        vals = dict(depth=d, emaxdpt=0.0625 * (gam * d) * (gam * d), cosphm1=1.0 / cosph, dellam1=1.0 / (zd * circ / 360.0))
        for k, v in vals.items():
            self.t[k].copy_(torch.from_numpy(np.ascontiguousarray(v)).view(self.C, self.P))
        n = self.P * self.C
        out = {k: np.empty((self.F, n)) for k in ("wavnum", "cinv", "cgroup", "xk2cg", "omosnh2kd", "stokfac")}
        dpp = C.POINTER(C.c_double)
        L.check(self.lib.ecwam_b200_host_depthprpt(C.byref(s.tables), self.F, n, d.ctypes.data_as(dpp),
                                                   *[out[k].ctypes.data_as(dpp) for k in
                                                     ("wavnum", "cinv", "cgroup", "xk2cg", "omosnh2kd", "stokfac")]), "depthprpt")
        for k, v in out.items():
            # (F, C*P) -> (C, F, P)
            self.t[k].copy_(torch.from_numpy(v).view(self.F, self.C, self.P).permute(1, 0, 2))

    def set_field(self, name: str, v_global: np.ndarray):
        self.t[name.lower()].copy_(self.torch.from_numpy(np.ascontiguousarray(v_global[self.src])).view(self.C, self.P))

    def get_field(self, name: str) -> np.ndarray:
        """Own points only, in slot order (use .own for their original global indices)."""
        return self.t[name.lower()].reshape(-1)[: self.nloc].cpu().numpy()

    def get_field3(self, name: str) -> np.ndarray:
        return self.t[name.lower()].permute(1, 0, 2).reshape(self.F, -1)[:, : self.nloc].cpu().numpy()

    def set_fl1(self, fl_global: np.ndarray):
        """fl_global[m, k, ij] in original global order."""
        torch = self.torch
        x = torch.from_numpy(np.ascontiguousarray(fl_global[:, :, self.src]))      # (F, A, C*P)
        self.t["fl1"].copy_(x.view(self.F, self.A, self.C, self.P).permute(2, 0, 1, 3))

    def get_spec(self, name="fl1") -> np.ndarray:
        """[m, k, own point] of FL1 or XLLWS."""
        x = self.t[name].permute(1, 2, 0, 3).reshape(self.F, self.A, -1)[:, :, : self.nloc]
        return x.cpu().numpy()

    # ---- GETWND's blocking step (WAMWND + MICEP): forcing grid -> FF_NEXT on the device
    def getwnd(self, fieldg: dict, ifromij, jfromij, nxs=1, nys=1, llwswave=0, llwdwave=0, lrelwind=1, iparamci=31, liceth=0, zmiss=-999.0):
        """fieldg: dict of (NY, NX) host arrays (FIELDG members); ifromij / jfromij: per sea point in the original global order.
        Returns the FF_NEXT device tensors, ready for newwind_device()."""
        torch = self.torch
        ny, nx = np.asarray(fieldg["uwnd"]).shape
        keep, g = [], L.FieldG()
        for n_, _ in L.FieldG._fields_:
            if n_ in fieldg:
                t = torch.from_numpy(np.ascontiguousarray(fieldg[n_], dtype=np.float64)).to(self.device)
                keep.append(t)
                setattr(g, n_, C.cast(t.data_ptr(), C.POINTER(C.c_double)))
        ii = torch.from_numpy(np.ascontiguousarray(np.asarray(ifromij, dtype=np.int32)[self.src])).to(self.device)
        jj = torch.from_numpy(np.ascontiguousarray(np.asarray(jfromij, dtype=np.int32)[self.src])).to(self.device)
        o = L.GetwndOpts(nxs=nxs, nxe=nxs + nx - 1, nys=nys, nye=nys + ny - 1, llwswave=llwswave, llwdwave=llwdwave, lrelwind=lrelwind,
                         iparamci=iparamci, liceth=liceth, zmiss=zmiss)
        nxt, out = L.ForcingNext(), {}
        for n_, _ in L.ForcingNext._fields_:
            out[n_] = torch.empty((self.C, self.P), dtype=torch.float64, device=self.device)
            setattr(nxt, n_, C.cast(out[n_].data_ptr(), C.POINTER(C.c_double)))
        L.check(self.lib.ecwam_b200_getwnd(self.h, C.byref(g), C.byref(o), C.c_void_p(ii.data_ptr()), C.c_void_p(jj.data_ptr()), C.byref(nxt)),
                "getwnd")
        self.synchronize()
        self._next = (nxt, out)
        return out

    def newwind_device(self):
        """NEWWIND on the FF_NEXT tensors the last getwnd() produced."""
        L.check(self.lib.ecwam_b200_newwind(self.h, C.byref(self._next[0])), "newwind")

    # ---- restart files in the reference's formats (SAVSPEC / SAVSTRESS, GETSPEC / GETSTRESS; host_io.cpp)
    LAW_FIELDS = ("wswave", "wdwave", "ufric", "tauw", "tauwdir", "z0m", "z0b", "chrnck", "aird", "wstar", "cicover", "cithick",
                  "ustra", "vstra", "ucur", "vcur")                       # savstress.F90:110-125 (NREAL = 16)

    def _ijorig(self):
        return np.ascontiguousarray(self.own + 1, dtype=np.int32)

    def _set_own(self, name, v_own):
        """v_own[..., own point] (slot order) -> the (.., C, P) device array; padded lanes repeat the chunk's first point."""
        slot = np.arange(self.P * self.C)
        slot = np.where(slot < self.nloc, slot, (slot // self.P) * self.P)
        x = self.torch.from_numpy(np.ascontiguousarray(v_own[..., slot]))
        if x.dim() == 3:
            self.t[name].copy_(x.view(self.F, self.A, self.C, self.P).permute(2, 0, 1, 3))
        else:
            self.t[name].copy_(x.view(self.C, self.P))

    def savspec(self, filename: str, create: bool = True, parallel: bool = False):
        """SAVSPEC (savspec.F90:86-166): this rank's spectra into the BLS file.  Global file: every rank writes its own points in
        place; the rank called with create=True must finish first (the caller places the barrier).  parallel=True is LRSTPARALW:
        one file per rank, FILENAME.<irank>_<nproc>."""
        self.synchronize()
        fl = np.ascontiguousarray(self.get_spec("fl1"))           # [m][k][own point] = Fortran (IJ, K, M)
        dp = C.POINTER(C.c_double)
        if parallel:
            buf = C.create_string_buffer(512)
            L.check(self.lib.ecwam_b200_restart_par_name(filename.encode(), self.rank + 1, self.s.nproc, buf, 512), "restart_par_name")
            L.check(self.lib.ecwam_b200_savspec_par(buf.value, self.nloc, self.A, self.F, fl.ctypes.data_as(dp)), "savspec_par")
            return buf.value.decode()
        ij = self._ijorig()
        L.check(self.lib.ecwam_b200_savspec(filename.encode(), self.s.niblo, self.A, self.F, self.nloc, ij.ctypes.data_as(C.POINTER(C.c_int)),
                                            fl.ctypes.data_as(dp), int(create)), "savspec")
        return filename

    def getspec(self, filename: str, parallel: bool = False):
        """GETSPEC's binary-restart branch (getspec.F90 -> READFL, readfl.F90:118-145): this rank's spectra out of the BLS file."""
        fl = np.empty((self.F, self.A, self.nloc))
        dp = C.POINTER(C.c_double)
        if parallel:
            buf = C.create_string_buffer(512)
            L.check(self.lib.ecwam_b200_restart_par_name(filename.encode(), self.rank + 1, self.s.nproc, buf, 512), "restart_par_name")
            L.check(self.lib.ecwam_b200_getspec_par(buf.value, self.nloc, self.A, self.F, fl.ctypes.data_as(dp)), "getspec_par")
        else:
            ij = self._ijorig()
            L.check(self.lib.ecwam_b200_getspec(filename.encode(), self.s.niblo, self.A, self.F, self.nloc, ij.ctypes.data_as(C.POINTER(C.c_int)),
                                                fl.ctypes.data_as(dp)), "getspec")
        self._set_own("fl1", fl)

    def savstress(self, filename: str, cdtpro: str, cdatewo: str = "", cdawifl: str = "", cdatefl: str = "", create: bool = True):
        """SAVSTRESS (savstress.F90:80-152): the 16 forcing / stress fields of the LAW file (global file, rank-wise as savspec)."""
        self.synchronize()
        r = np.ascontiguousarray(np.stack([self.get_field(n) for n in self.LAW_FIELDS]))      # [field][own point]
        ij = self._ijorig()
        L.check(self.lib.ecwam_b200_savstress(filename.encode(), cdtpro.encode(), (cdatewo or cdtpro).encode(), (cdawifl or cdtpro).encode(),
                                              (cdatefl or cdtpro).encode(), self.s.niblo, len(self.LAW_FIELDS), self.nloc,
                                              ij.ctypes.data_as(C.POINTER(C.c_int)), r.ctypes.data_as(C.POINTER(C.c_double)), int(create)),
                "savstress")

    def getstress(self, filename: str):
        """GETSTRESS -> READSTRESS (readstress.F90:97-124); returns (CDTPRO, CDATEWO, CDAWIFL, CDATEFL)."""
        r = np.empty((len(self.LAW_FIELDS), self.nloc))
        ij = self._ijorig()
        dates = C.create_string_buffer(60)
        L.check(self.lib.ecwam_b200_getstress(filename.encode(), dates, self.s.niblo, len(self.LAW_FIELDS), self.nloc,
                                              ij.ctypes.data_as(C.POINTER(C.c_int)), r.ctypes.data_as(C.POINTER(C.c_double))), "getstress")
        for i, n in enumerate(self.LAW_FIELDS):
            self._set_own(n, r[i])
        return tuple(dates.raw[15 * i: 15 * i + 14].decode() for i in range(4))

    # ---- the hot path
    def propag(self) -> int:
        return L.check(self.lib.ecwam_b200_propag(self.h), "propag")

    def implsch(self):
        L.check(self.lib.ecwam_b200_implsch_all(self.h), "implsch_all")

    def step(self) -> int:
        """One WAMINTGR sub-step with IDELPRO == IDELT."""
        return L.check(self.lib.ecwam_b200_wamintgr(self.h), "wamintgr")

    # ---- the steps either side of the hot path: NEWWIND, OUTBS, OUTWNORM
    NEXT_FIELDS = ("wswave", "wdwave", "aird", "wstar", "cicover", "cithick", "ustra", "vstra")

    def newwind(self, nxt: dict):
        """NEWWIND (newwind.F90:105-167): FF_NOW <- FF_NEXT on the device; nxt maps field names to global arrays."""
        torch = self.torch
        keep, fn = [], L.ForcingNext()
        usf = self.par.icode_wnd != 3          # the friction velocity is the forcing (nxt["UFRIC"]); FF_NEXT%WSWAVE is not read
        for n in self.NEXT_FIELDS:
            if usf and n == "wswave":
                continue
            v = nxt[n] if n in nxt else nxt[n.upper()]
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(v, dtype=np.float64)[self.src])).to(self.device)
            keep.append(t)
            setattr(fn, n, C.cast(t.data_ptr(), C.POINTER(C.c_double)))
        if usf:
            us = torch.from_numpy(np.ascontiguousarray(np.asarray(nxt["UFRIC"], dtype=np.float64)[self.src])).to(self.device)
            keep.append(us)
            L.check(self.lib.ecwam_b200_newwind_ustar(self.h, C.byref(fn), C.c_void_p(us.data_ptr())), "newwind_ustar")
        else:
            L.check(self.lib.ecwam_b200_newwind(self.h, C.byref(fn)), "newwind")
        self.synchronize()      # `keep` may be released once the kernel has read it

    def _outsel(self, itg, icemask, seamask, zmiss, llsource):
        n = len(itg)
        arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (itg, icemask, seamask)]
        sel = L.OutSel(niprmout=n, llsource=int(llsource), zmiss=float(zmiss))
        for nm, a in zip(("itg", "icemask", "seamask"), arrs):
            setattr(sel, nm, a.ctypes.data_as(C.POINTER(C.c_int)))
        return sel, arrs

    def outbs(self, itg, icemask, seamask, zmiss=-999.0, llsource=1, iodp=None):
        """OUTBS/OUTBLOCK (outbs.F90:97-122) on the device: returns BOUT[column, own point] (slot order, see .own)."""
        torch = self.torch
        sel, keep = self._outsel(itg, icemask, seamask, zmiss, llsource)
        n = len(itg)
        self.bout = torch.empty((self.C, n, self.P), dtype=torch.float64, device=self.device)
        self._sel = (list(itg), list(icemask), list(seamask), zmiss, llsource)
        io = None
        if iodp is not None:
            io = torch.from_numpy(np.ascontiguousarray(np.asarray(iodp, dtype=np.int32)[self.src])).to(self.device)
        L.check(self.lib.ecwam_b200_outbs(self.h, C.byref(sel), C.c_void_p(io.data_ptr()) if io is not None else None,
                                          C.c_void_p(self.bout.data_ptr())), "outbs")
        self.synchronize()
        return self.bout.permute(1, 0, 2).reshape(n, -1)[:, : self.nloc].cpu().numpy()

    def outwnorm(self, global_norm=True):
        """OUTWNORM/MPMINMAXAVG on the last outbs(): rows = columns, (average, minimum, maximum, count)."""
        sel, keep = self._outsel(*self._sel)
        n = len(self._sel[0])
        w = np.empty((n, 4))
        s = self.s
        ip = C.POINTER(C.c_int)
        multi = s.nproc > 1
        ij2 = np.ascontiguousarray(s.ij2new, dtype=np.int32)
        ns = np.ascontiguousarray(s.nstart, dtype=np.int32)
        ne = np.ascontiguousarray(s.nend, dtype=np.int32)
        L.check(self.lib.ecwam_b200_outwnorm(self.h, C.byref(sel), C.c_void_p(self.bout.data_ptr()), int(global_norm), int(s.niblo),
                                             ij2.ctypes.data_as(ip) if multi else None, ns.ctypes.data_as(ip) if multi else None,
                                             ne.ctypes.data_as(ip) if multi else None, w.ctypes.data_as(C.POINTER(C.c_double))),
                "outwnorm")
        return w

    def no_source(self, llsource_off: bool):
        L.check(self.lib.ecwam_b200_no_source(self.h, int(llsource_off)), "no_source")

    def wamintgr(self, clk: "WamClock", ff_next=None, llsource=True) -> int:
        """One call of WAMINTGR (wamintgr.F90:92-197) with its time bookkeeping; times are seconds since the start of the
        run instead of the reference's CHARACTER(14) dates.  Returns the CFL count of PROPAG_WAM (0 = ok).
        ff_next: callable(time) -> dict of the FF_NEXT fields, asked when NEWWIND's `CDATE >= CDATEWH` test holds."""
        cfl = 0
        if clk.cdate == clk.cdtpra:                       # :92-98  PROPAGATION TIME
            cfl = self.propag()
            clk.cdate = clk.cdtpro
        if clk.cdtimp >= clk.cdatewh:                     # :103-104 NEWWIND (newwind.F90:107-169)
            if ff_next is not None:
                self.newwind(ff_next(clk.cdatewh))
            clk.cdatewh += clk.idelwo
        if clk.cdate >= clk.cdtimpnext:                   # :108-186 IT IS TIME TO INTEGRATE THE SOURCE TERMS
            if llsource:
                self.implsch()
            else:
                self.no_source(True)
            clk.cdatewo = clk.cdatewh                     # :181
            clk.cdtimp = clk.cdtimpnext
            clk.cdtimpnext += clk.idelt
        else:                                             # :187-195 NO SOURCE TERM CONTRIBUTION
            self.no_source(False)
        return cfl

    def advection_step(self, clk: "WamClock", ff_next=None, llsource=True) -> int:
        """One iteration of WAMODEL's ADVECTION loop (wamodel.F90:228-233, 283-300): fix the end date of the propagation
        step, then call WAMINTGR until the source terms have caught up with it."""
        clk.cdtpra = clk.cdtpro
        clk.cdtpro += clk.idelpro
        clk.cdate = clk.cdtpra
        clk.cdatewh = clk.cdatewo                         # wamodel.F90:286 (differs from the running CDATEWH only when IDELT > IDELPRO)
        cfl, iloop = 0, 1
        while iloop == 1 or clk.cdtimpnext <= clk.cdtpro:
            cfl += self.wamintgr(clk, ff_next, llsource)
            iloop += 1
        return cfl

    def synchronize(self):
        L.check(self.lib.ecwam_b200_synchronize(self.h), "synchronize")

    def launch_count(self) -> int:
        return int(self.lib.ecwam_b200_launch_count(self.h))

    def timing(self, name: str):
        tot, cnt = C.c_double(), C.c_longlong()
        self.lib.ecwam_b200_timing_get(self.h, name.encode(), C.byref(tot), C.byref(cnt))
        return tot.value, cnt.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.ecwam_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _table_grid(self, name):
    d = self.lib.ecwam_b200_host_grid_decomp(self._hg, 1).contents
    return _np(getattr(d, name), d.ngy, np.float64)


WamSetup.table_grid = _table_grid


def hs_fm(setup: WamSetup, fl: np.ndarray):
    """Hs = 4 sqrt(EM) and mean frequency FM of spectra fl[m,k,n] (femean.F90, outblock.F90:223-244) — numpy, for
    diagnostics and parity metrics only."""
    F = setup.par.nfre
    dfim = setup.table("dfim", F)
    dfimofr = setup.table("dfimofr", F)
    fr = setup.table("fr", F)
    delth = setup.tables.delth
    eps = setup.tables.epsmin
    t2 = np.maximum(fl, eps).sum(axis=1)                       # (F, n)
    em = (t2 * dfim[:, None]).sum(axis=0) + setup.tables.wetail * fr[-1] * delth * t2[-1]
    fm = (t2 * dfimofr[:, None]).sum(axis=0) + setup.tables.frtail * delth * t2[-1]
    fm = np.maximum(em / fm, fr[0])
    return 4.0 * np.sqrt(em), fm
