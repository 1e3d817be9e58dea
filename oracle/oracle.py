"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE — see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class Config(C.Structure):
    """Mirror of orc::Config (oracle/oracle.h) — field order and types must match."""
    _fields_ = [
        ("nang", C.c_int), ("nfre", C.c_int), ("nfre_red", C.c_int), ("ifre1", C.c_int), ("fr1", C.c_double),
        ("iphys", C.c_int), ("isnonlin", C.c_int), ("idamping", C.c_int), ("irefra", C.c_int), ("icase", C.c_int),
        ("ipropags", C.c_int), ("llgcbz0", C.c_int), ("llnormagam", C.c_int), ("llcapchnk", C.c_int),
        ("lbiwbk", C.c_int), ("licerun", C.c_int), ("lmaskice", C.c_int), ("lwamrsetci", C.c_int),
        ("lciwa1", C.c_int), ("lciwa2", C.c_int), ("lciwa3", C.c_int), ("lciscal", C.c_int), ("lwflux", C.c_int),
        ("lwfluxout", C.c_int), ("lwnemocou", C.c_int), ("lwvflx_snl", C.c_int), ("lwcouast", C.c_int),
        ("lwcou", C.c_int), ("icode", C.c_int), ("idelt", C.c_double), ("idelpro", C.c_double),
        ("delpro_lf", C.c_double), ("ifrelfmax", C.c_int), ("ximp", C.c_double), ("rnu", C.c_double),
        ("rnum", C.c_double), ("wspmin", C.c_double), ("cithrsh", C.c_double), ("cithrsh_tail", C.c_double),
        ("ciblock", C.c_double), ("flmin", C.c_double), ("zalpfacx", C.c_double), ("zalpfacb", C.c_double), ("cdicwa", C.c_double),
        ("bathymax", C.c_double),
        ("deptha", C.c_double), ("nproma", C.c_int), ("npr", C.c_int), ("ll1d", C.c_int),
        ("store_all_weights", C.c_int), ("nthreads", C.c_int), ("llcflcuroff", C.c_int),
        ("lwnemotauoc", C.c_int), ("lwnemocoustk", C.c_int), ("lwnemocoustrn", C.c_int), ("lwnemocousend", C.c_int),
        ("lwnemocouwrs", C.c_int), ("lwnemocouibr", C.c_int), ("zalpwrs", C.c_double), ("zibrw_thrsh", C.c_double),
    ]


def default_config(**kw) -> Config:
    c = Config(nang=12, nfre=36, nfre_red=25, ifre1=3, fr1=4.177248e-02, iphys=1, isnonlin=0, idamping=1, irefra=0,
               icase=1, ipropags=2, llgcbz0=0, llnormagam=0, llcapchnk=1, lbiwbk=1, licerun=1, lmaskice=1,
               lwamrsetci=1, lciwa1=0, lciwa2=0, lciwa3=0, lciscal=0, lwflux=0, lwfluxout=1, lwnemocou=0,
               lwvflx_snl=1, lwcouast=0, lwcou=0, icode=3, idelt=900.0, idelpro=900.0, delpro_lf=900.0, ifrelfmax=0,
               ximp=1.0, rnu=1.5e-5, rnum=0.11 * 1.5e-5, wspmin=1.0, cithrsh=0.3, cithrsh_tail=0.3, ciblock=0.0,
               flmin=1e-5, zalpfacx=1.0, zalpfacb=1.0, cdicwa=0.01, bathymax=998.999, deptha=2.0, nproma=32, npr=1, ll1d=0,
               store_all_weights=0, nthreads=1, llcflcuroff=1, lwnemotauoc=0, lwnemocoustk=0, lwnemocoustrn=0, lwnemocousend=1, lwnemocouwrs=0,
               lwnemocouibr=0, zalpwrs=1.0, zibrw_thrsh=0.5)
    for k, v in kw.items():
        if not hasattr(c, k):
            raise KeyError(k)
        setattr(c, k, v)
    c.ifre1 = 1 if c.nfre_red == 25 else 3
    return c


def build(fast: bool = False) -> str:
    name = "liboracle_fast.so" if fast else "liboracle.so"
    path = os.path.join(_HERE, name)
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, name], stdout=subprocess.DEVNULL)
    return path


_libs = {}


def _lib(fast=False):
    if fast not in _libs:
        lib = C.CDLL(build(fast))
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.POINTER(Config), C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        lib.orc_destroy.argtypes = [C.c_void_p]
        lib.orc_niblo.argtypes = [C.c_void_p]
        for f in ("orc_get_table",):
            getattr(lib, f).restype = C.c_long
            getattr(lib, f).argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_long]
        for f in ("orc_get_itable", "orc_get_rank_double"):
            getattr(lib, f).restype = C.c_long
            getattr(lib, f).argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_long]
        for f in ("orc_set_field", "orc_get_field", "orc_get_field3"):
            getattr(lib, f).argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        for f in ("orc_set_fl1", "orc_get_fl1", "orc_get_xllws"):
            getattr(lib, f).argtypes = [C.c_void_p, C.c_void_p]
        lib.orc_get_hs_fm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_propag.argtypes = [C.c_void_p]
        lib.orc_newwind.argtypes = [C.c_void_p, C.c_void_p]
        lib.orc_snonlin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_term.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.orc_stresso.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_capture.argtypes = [C.c_void_p, C.c_int]
        lib.orc_get_capture.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_outbs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p]
        lib.orc_outwnorm.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.orc_implsch.argtypes = [C.c_void_p]
        lib.orc_getwnd_points.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]
        assert lib.orc_config_size() == C.sizeof(Config), "oracle Config layout mismatch"
        _libs[fast] = lib
    return _libs[fast]


FIELDG_NAMES = ("uwnd", "vwnd", "aird", "wstar", "cicover", "cithick", "ustra", "vstra", "wswave", "wdwave")
NEXT_NAMES = ("wswave", "wdwave", "aird", "wstar", "cicover", "cithick", "ustra", "vstra")


def getwnd_points(ifromij, jfromij, fieldg, nxs=1, nys=1, ucur=None, vcur=None, llwswave=0, llwdwave=0, lcorrel=0, iparamci=31, liceth=0,
                  licerun=1, lmaskice=1, wspmin=1.0, zmiss=-999.0):
    """WAMWND + MICEP (getwnd.F90:196-212) for the points (ifromij, jfromij) of the forcing grid; fieldg: dict of (NY, NX) arrays.
    Returns the FF_NEXT members as a dict."""
    lib = _lib()
    n = len(ifromij)
    ii, jj = np.ascontiguousarray(ifromij, dtype=np.int32), np.ascontiguousarray(jfromij, dtype=np.int32)
    ny, nx = np.asarray(fieldg["uwnd"]).shape
    zero = np.zeros((ny, nx))
    g = [np.ascontiguousarray(fieldg.get(k, zero), dtype=np.float64) for k in FIELDG_NAMES]
    gp = (C.c_void_p * 10)(*[a.ctypes.data for a in g])
    uc = np.ascontiguousarray(ucur if ucur is not None else np.zeros(n))
    vc = np.ascontiguousarray(vcur if vcur is not None else np.zeros(n))
    out = [np.empty(n) for _ in range(8)]
    op = (C.c_void_p * 8)(*[a.ctypes.data for a in out])
    opt = np.array([llwswave, llwdwave, lcorrel, iparamci, liceth, licerun, lmaskice], dtype=np.int32)
    ropt = np.array([wspmin, zmiss, 2.0 * np.pi])
    lib.orc_getwnd_points(n, ii.ctypes.data, jj.ctypes.data, nxs, nys, nx, gp, uc.ctypes.data, vc.ctypes.data, opt.ctypes.data, ropt.ctypes.data, op)
    return dict(zip(NEXT_NAMES, out))


class Oracle:
    def __init__(self, cfg: Config, grid, fast: bool = False):
        """grid: ecwam_b200.synth.SynthGrid (or anything with ngy/nlonrgg/amosop/amonop/mask/depth)."""
        self.lib = _lib(fast)
        self.cfg = cfg
        nl = np.ascontiguousarray(grid.nlonrgg, dtype=np.int32)
        mk = np.ascontiguousarray(grid.mask, dtype=np.uint8)
        dp = np.ascontiguousarray(grid.depth, dtype=np.float64)
        self.h = self.lib.orc_create(C.byref(cfg), int(grid.ngy), nl.ctypes.data, float(grid.amosop), float(grid.amonop),
                                     mk.ctypes.data, dp.ctypes.data)
        if not self.h:
            raise RuntimeError("orc_create failed")
        self.niblo = self.lib.orc_niblo(self.h)

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- tables
    def table(self, name, cap=1 << 16):
        buf = np.empty(cap, dtype=np.float64)
        n = self.lib.orc_get_table(self.h, name.encode(), buf.ctypes.data, cap)
        if n < 0:
            return self.table(name, -n)
        if n == 0:
            raise KeyError(name)
        return buf[:n].copy()

    def itable(self, name, rank=0, cap=1 << 16):
        buf = np.empty(cap, dtype=np.int32)
        n = self.lib.orc_get_itable(self.h, name.encode(), rank, buf.ctypes.data, cap)
        if n < 0:
            return self.itable(name, rank, -n)
        if n == 0:
            raise KeyError(name)
        return buf[:n].copy()

    def iscalar(self, name, rank=0):
        return int(self.itable(name, rank, 4)[0])

    def rank_double(self, name, rank=0, cap=1 << 20):
        buf = np.empty(cap, dtype=np.float64)
        n = self.lib.orc_get_rank_double(self.h, name.encode(), rank, buf.ctypes.data, cap)
        if n < 0:
            return self.rank_double(name, rank, -n)
        if n == 0:
            raise KeyError(name)
        return buf[:n].copy()

    # ---- per-point fields (original global order)
    def set_field(self, name, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.shape == (self.niblo,)
        if self.lib.orc_set_field(self.h, name.encode(), v.ctypes.data) != 0:
            raise KeyError(name)

    def get_field(self, name):
        v = np.empty(self.niblo)
        if self.lib.orc_get_field(self.h, name.encode(), v.ctypes.data) != 0:
            raise KeyError(name)
        return v

    def get_field3(self, name):
        v = np.empty((self.cfg.nfre, self.niblo))
        if self.lib.orc_get_field3(self.h, name.encode(), v.ctypes.data) != 0:
            raise KeyError(name)
        return v

    def set_obstructions(self, lon, lat, cor):
        """OBSLON / OBSLAT [ic, m, ij] (2, NFRE_RED, NIBLO) and OBSCOR (4, NFRE_RED, NIBLO) in the original point order (LSUBGRID = T)."""
        n, fr = self.niblo, self.cfg.nfre_red
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (lon, lat, cor)]
        assert a[0].shape == (2, fr, n) and a[1].shape == (2, fr, n) and a[2].shape == (4, fr, n)
        self.lib.orc_set_obstructions.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        if self.lib.orc_set_obstructions(self.h, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data) != 0:
            raise RuntimeError("orc_set_obstructions failed")

    def set_fl1(self, fl):
        fl = np.ascontiguousarray(fl, dtype=np.float64)
        assert fl.shape == (self.cfg.nfre, self.cfg.nang, self.niblo)
        self.lib.orc_set_fl1(self.h, fl.ctypes.data)

    def get_fl1(self):
        fl = np.empty((self.cfg.nfre, self.cfg.nang, self.niblo))
        self.lib.orc_get_fl1(self.h, fl.ctypes.data)
        return fl

    def get_xllws(self):
        fl = np.empty((self.cfg.nfre, self.cfg.nang, self.niblo))
        self.lib.orc_get_xllws(self.h, fl.ctypes.data)
        return fl

    def hs_fm(self):
        hs = np.empty(self.niblo)
        fm = np.empty(self.niblo)
        self.lib.orc_get_hs_fm(self.h, hs.ctypes.data, fm.ctypes.data)
        return hs, fm

    # ---- hot path
    def propag(self) -> int:
        rc = self.lib.orc_propag(self.h)
        if rc < 0:
            raise RuntimeError("orc_propag failed")
        return rc

    def implsch(self):
        if self.lib.orc_implsch(self.h) != 0:
            raise RuntimeError("orc_implsch failed")

    # ---- the steps either side of the hot path (oracle/orc_output.cpp)
    NEXT_FIELDS = ("WSWAVE", "WDWAVE", "AIRD", "WSTAR", "CICOVER", "CITHICK", "USTRA", "VSTRA")

    def newwind(self, nxt: dict):
        """NEWWIND (newwind.F90:105-167): FF_NOW <- FF_NEXT; nxt maps the 8 field names to global arrays (with ICODE_WND = 1, 2 the
        friction velocity nxt["UFRIC"] takes the place of WSWAVE)."""
        arrs = [np.ascontiguousarray(nxt["UFRIC" if (k == "WSWAVE" and self.cfg.icode != 3) else k], dtype=np.float64) for k in self.NEXT_FIELDS]
        ptrs = (C.c_void_p * 8)(*[a.ctypes.data for a in arrs])
        if self.lib.orc_newwind(self.h, ptrs) != 0:
            raise RuntimeError("orc_newwind failed")

    def outbs(self, itg, icemask, seamask, zmiss=-999.0, llsource=1):
        """OUTBS/OUTBLOCK (outbs.F90:97-122): returns BOUT[column, ij] in original global order."""
        itg = np.ascontiguousarray(itg, dtype=np.int32)
        im = np.ascontiguousarray(icemask, dtype=np.int32)
        sm = np.ascontiguousarray(seamask, dtype=np.int32)
        out = np.empty((len(itg), self.niblo))
        if self.lib.orc_outbs(self.h, len(itg), itg.ctypes.data, im.ctypes.data, sm.ctypes.data, float(zmiss), int(llsource),
                              out.ctypes.data) != 0:
            raise RuntimeError("orc_outbs failed (unsupported parameter?)")
        self._nout = len(itg)
        return out

    def snonlin(self):
        """SNONLIN alone (snonlin.F90): (SL, FLD)[m, k, ij] of the current spectra."""
        sl = np.empty((self.cfg.nfre, self.cfg.nang, self.niblo))
        fld = np.empty_like(sl)
        if self.lib.orc_snonlin(self.h, sl.ctypes.data, fld.ctypes.data) != 0:
            raise RuntimeError("orc_snonlin failed")
        return sl, fld

    def term(self, which):
        """One source term alone: "sinput" (NGST = 1, LLSNEG = F, stored UFRIC / Z0M), "sdissip", "sbottom", "sdiwbk" or "sdice" (the LCIWA1-3 terms that are on): (SL, FLD)[m, k, ij]."""
        sl = np.empty((self.cfg.nfre, self.cfg.nang, self.niblo))
        fld = np.empty_like(sl)
        if self.lib.orc_term(self.h, {"sinput": 1, "sdissip": 2, "sbottom": 3, "sdiwbk": 4, "sdice": 5}[which], sl.ctypes.data, fld.ctypes.data) != 0:
            raise RuntimeError("orc_term failed")
        return sl, fld

    def stresso(self):
        """The wind input of the second SINFLX call (NGST = 2, LLSNEG = T) and STRESSO + TAU_PHI_HF on it, with the stored UFRIC, Z0M, MIJ:
        (SL, SPOS)[m, k, ij] and (TAUW, TAUWDIR, PHIWA)[ij]."""
        sl = np.empty((self.cfg.nfre, self.cfg.nang, self.niblo))
        spos = np.empty_like(sl)
        out = np.empty((3, self.niblo))
        if self.lib.orc_stresso(self.h, sl.ctypes.data, spos.ctypes.data, out.ctypes.data) != 0:
            raise RuntimeError("orc_stresso failed")
        return sl, spos, out

    def capture(self, on=True):
        """Keep the IMPLSCH-internal inputs of WNFLUXES (SSOURCE, EMEAN, F1MEAN, PHIWA) of the next implsch() / step()."""
        self.lib.orc_capture(self.h, int(on))

    def captured(self):
        ss = np.empty((self.cfg.nfre, self.cfg.nang, self.niblo))
        em, f1, ph = np.empty(self.niblo), np.empty(self.niblo), np.empty(self.niblo)
        if self.lib.orc_get_capture(self.h, ss.ctypes.data, em.ctypes.data, f1.ctypes.data, ph.ctypes.data) != 0:
            raise RuntimeError("orc_get_capture: capture() was not switched on")
        return ss, em, f1, ph

    def outwnorm(self, global_norm=True):
        """OUTWNORM/MPMINMAXAVG on the last outbs(): rows = columns, (average, minimum, maximum, count)."""
        w = np.empty((self._nout, 4))
        if self.lib.orc_outwnorm(self.h, int(global_norm), w.ctypes.data) != 0:
            raise RuntimeError("orc_outwnorm failed")
        return w

    def step(self):
        """One WAMINTGR sub-step with IDELPRO == IDELT (wamintgr.F90:94-146)."""
        cfl = self.propag()
        self.implsch()
        return cfl
