"""CPU tests: the C ABI library loads and exports every declared symbol; the product's host-side builders (tables,
MPDECOMP / PROPCONNECT / halo lists) agree with the oracle bit for bit on the integer tables."""
import ctypes as C

import numpy as np
import pytest

from ecwam_b200 import lib as L, synth, model as M
from oracle import oracle as O


def test_library_exports_every_declared_symbol(built):
    lib = L.load()
    assert len(L.EXPORTS) >= 30
    for name in L.EXPORTS:
        assert hasattr(lib, name), "include/ecwam_b200.h declares %s but the library does not export it" % name
    assert lib.ecwam_b200_version() >= 100


def test_struct_mirrors_match_header(built):
    # sizes are what a C compiler would give: ints 4, doubles/pointers 8, natural alignment
    assert C.sizeof(L.Params) % 8 == 0 and C.sizeof(L.Fields) == 8 * len(L.Fields._fields_)
    assert {n for n, _ in L.Fields._fields_} >= {"fl1", "xllws", "mij", "ufric", "tauw"}


def test_create_without_gpu_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    g = synth.make_grid(8, "aqua")
    s = M.WamSetup(g, nproc=1)
    with pytest.raises(L.EcwamError):
        M.WamIntgr(s, 0, device="cpu")
    # straight through the C ABI: no CUDA device -> error code, never a CPU fallback
    par = L.Params()
    C.memmove(C.byref(par), C.byref(s.par), C.sizeof(L.Params))
    par.nproma, par.nchnk = 32, (g.niblo + 31) // 32
    dec = s.decomp(0)
    h = C.c_void_p()
    rc = s.lib.ecwam_b200_create(C.byref(par), C.byref(s.tables), C.byref(dec), None, None, C.byref(h))
    assert rc == -2 and b"no CPU fallback" in s.lib.ecwam_b200_last_error()


CONFIGS = [(20, 12, 25, "aqua", 1, 1), (20, 12, 25, "continents", 4, 1), (16, 24, 29, "continents", 8, 0),
           (24, 36, 29, "continents", 2, 1), (20, 36, 29, "continents", 3, 1), (20, 12, 25, "aqua", 5, 1),
           (20, 12, 25, "continents", 6, 1), (24, 12, 25, "continents", 7, 1)]


@pytest.mark.parametrize("N,A,Fr,mask,npr,iphys", CONFIGS)
def test_host_tables_and_decomposition_match_oracle(built, N, A, Fr, mask, npr, iphys):
    g = synth.make_grid(N, mask)
    o = O.Oracle(O.default_config(nang=A, nfre_red=Fr, nproma=32, npr=npr, iphys=iphys), g)
    s = M.WamSetup(g, nproc=npr, nang=A, nfre_red=Fr, iphys=iphys)
    F = 36
    for nm, n in [("fr", F), ("dfim", F), ("dfimofr", F), ("dfimfr", F), ("zpifr", F), ("fr5", F), ("cofrm4", F), ("flmax", F),
                  ("rhowg_dfim", F), ("dfim_sim", F), ("th", A), ("costh", A), ("sinth", A), ("swellft", 200), ("wtauhf", 19)]:
        np.testing.assert_array_equal(s.table(nm, n), o.table(nm.upper()), err_msg=nm)
    ml, lo = o.iscalar("MLSTHG"), o.iscalar("MFRSTLW")
    assert (s.tables.mlsthg, s.tables.mfrstlw, s.tables.kfrh, s.tables.nfre_odd) == (ml, lo, o.iscalar("KFRH"), o.iscalar("NFRE_ODD"))
    for nm, n in [("ikp", ml - lo + 1), ("ikp1", ml - lo + 1), ("ikm", ml - lo + 1), ("ikm1", ml - lo + 1), ("k1w", 2 * A),
                  ("k2w", 2 * A), ("k11w", 2 * A), ("k21w", 2 * A), ("inlcoef", 5 * ml)]:
        np.testing.assert_array_equal(s.itable(nm, n), o.itable(nm.upper()), err_msg=nm)     # bit-exact integer tables
    for nm, n in [("rnlcoef", 25 * ml), ("af11", ml - lo + 1)]:
        np.testing.assert_array_equal(s.table(nm, n), o.table(nm.upper()), err_msg=nm)
    if iphys == 1:
        ns = 2 * s.tables.nsdsnth + 1
        assert s.tables.nsdsnth == o.iscalar("NSDSNTH")
        np.testing.assert_array_equal(s.itable("indicessat", A * ns), o.itable("INDICESSAT"))
        np.testing.assert_array_equal(s.table("satweights", A * ns), o.table("SATWEIGHTS"))
    for nm in ("x0tauhf", "delth", "flogsprdm1", "betamaxoxkappa2", "dal1", "dal2", "tauwshelter"):
        assert getattr(s.tables, nm) == o.table(nm.upper())[0], nm
    # MPDECOMP: relabelling, rank ranges, neighbour tables, halo lists (bit-exact)
    np.testing.assert_array_equal(s.ij2new, o.itable("IJ2NEWIJ"))
    np.testing.assert_array_equal(s.nstart, o.itable("NSTART"))
    np.testing.assert_array_equal(s.nend, o.itable("NEND"))
    for r in range(npr):
        d = s.decomp_arrays(r)
        for nm in ("klat", "klon", "kcor", "nfrompe", "ntope", "nijstart"):
            np.testing.assert_array_equal(d[nm], o.itable(nm.upper(), r), err_msg="%s rank %d" % (nm, r))
        for nm in ("wlat", "wcor"):
            np.testing.assert_array_equal(d[nm], o.rank_double(nm.upper(), r), err_msg=nm)
        for nm in ("ninf", "nsup", "ijs", "ijl"):
            assert d[nm] == o.iscalar(nm.upper(), r)
        if npr > 1:
            np.testing.assert_array_equal(d["ijtope"], o.itable("IJTOPE", r))


@pytest.mark.parametrize("A,iphys,gc,ng", [(12, 1, 1, 1), (36, 1, 1, 1), (12, 0, 1, 1), (24, 1, 0, 1), (12, 1, 1, 0), (12, 1, 0, 0)])
def test_gravity_capillary_tables_match_oracle(built, A, iphys, gc, ng):
    """SETWAVPHYS for every LLGCBZ0 / LLNORMAGAM combination (setwavphys.F90:46-205), INIT_X0TAUHF's BMAXOKAP / GAMNCONST
    (init_x0tauhf.F90:65-72) and INITGC's wavenumber grid (initgc.F90:63-110): host tables of the product == oracle, bit for bit."""
    g = synth.make_grid(8, "aqua")
    kw = dict(nang=A, nfre_red=25, iphys=iphys, llgcbz0=gc, llnormagam=ng, wspmin=0.3 if gc else 1.0)
    o = O.Oracle(O.default_config(nproma=32, npr=1, **kw), g)
    s = M.WamSetup(g, nproc=1, **kw)
    for nm in ("alpha", "alphamin", "alphamax", "alphapmax", "chnkmin_u", "acdlin", "bcdlin", "bmaxokap", "gamnconst", "rn1_rn",
               "dthrn_a", "dthrn_u", "ang_gc_a", "ang_gc_b", "ang_gc_c", "sqrtgosurft", "betamaxoxkappa2", "tauwshelter", "x0tauhf",
               "z0rat", "z0tubmax", "swellf4", "swellf7", "cdis", "delta_sdis", "cdisvis"):
        assert getattr(s.tables, nm) == o.table(nm.upper())[0], nm
    n = s.tables.nwav_gc
    assert n == int(o.table("NWAV_GC")[0]) == 82
    for nm in ("xk_gc", "omega_gc", "cm_gc", "c2osqrtvg_gc", "xkmsqrtvgoc2_gc", "om3gmkm_gc", "omxkm3_gc", "delkcc_gc_ns",
               "delkcc_omxkm3_gc"):
        np.testing.assert_array_equal(s.table(nm, n), o.table(nm.upper()), err_msg=nm)
    for nm, k in (("wtauhf", 19), ("flmax", 36)):
        np.testing.assert_array_equal(s.table(nm, k), o.table(nm.upper()), err_msg=nm)
    # the dispersion relation the tables are built on: omega^2 = g k + T k^3 (gc_dispersion.h)
    k, om = s.table("xk_gc", n), s.table("omega_gc", n)
    np.testing.assert_allclose(om ** 2, 9.806 * k + 7.17e-5 * k ** 3, rtol=1e-14)


def test_ice_attenuation_table_matches_oracle(built):
    """CIGETDEAC (cigetdeac.F90:60-end): the product's host table (data include + its own statement of the two extrapolation rules)
    == the oracle's, bit for bit; dimensions and axes as in YOWICE."""
    g = synth.make_grid(8, "aqua")
    o = O.Oracle(O.default_config(nproma=32, npr=1), g)
    s = M.WamSetup(g, nproc=1)
    t = s.tables
    assert (t.nict, t.nich, t.ticmin, t.hicmin, t.dtic, t.dhic) == (16, 36, 1.0, 0.2, 1.0, 0.1)
    np.testing.assert_array_equal(s.table("cideac", 16 * 36), o.table("CIDEAC"))


def test_depthprpt_matches_oracle(built):
    g = synth.make_grid(16, "continents")
    o = O.Oracle(O.default_config(nang=12, nfre_red=25), g)
    s = M.WamSetup(g, nproc=1, nang=12, nfre_red=25)
    n = g.niblo
    out = {k: np.empty((36, n)) for k in ("wavnum", "cinv", "cgroup", "xk2cg", "omosnh2kd", "stokfac")}
    dpp = C.POINTER(C.c_double)
    d = np.ascontiguousarray(g.depth)
    rc = s.lib.ecwam_b200_host_depthprpt(C.byref(s.tables), 36, n, d.ctypes.data_as(dpp),
                                         *[out[k].ctypes.data_as(dpp) for k in ("wavnum", "cinv", "cgroup", "xk2cg", "omosnh2kd", "stokfac")])
    assert rc == 0
    for k in ("wavnum", "cinv", "cgroup", "xk2cg", "omosnh2kd", "stokfac"):
        np.testing.assert_array_equal(out[k], o.get_field3(k.upper()), err_msg=k)


def test_halo_plan_is_symmetric(built):
    """what rank p sends to q is what q expects from p (NTOPE on p == NFROMPE on q), and the sent points are the ones q's halo
    slots stand for (mpdecomp.F90:990-1176)."""
    g = synth.make_grid(24, "continents")
    npr = 6
    s = M.WamSetup(g, nproc=npr, nang=12, nfre_red=25)
    ds = [s.decomp_arrays(r) for r in range(npr)]
    for p in range(npr):
        for q in range(npr):
            assert ds[p]["ntope"][q] == ds[q]["nfrompe"][p]
        assert ds[p]["nfrompe"].sum() == (ds[p]["ijs"] - ds[p]["ninf"]) + (ds[p]["nsup"] - ds[p]["ijl"])
    # every neighbour index is inside NINF..NSUP+1
    for p in range(npr):
        for nm in ("klat", "klon", "kcor"):
            assert ds[p][nm].min() >= ds[p]["ninf"] and ds[p][nm].max() <= ds[p]["nsup"] + 1


# ---- the Fortran side of the boundary (fortran/*.F90; no Fortran compiler in the image: checked mechanically) ------------------
def _fortran_dummies(path, name):
    """Dummy-argument names of SUBROUTINE `name` in a free-form Fortran file."""
    import re
    with open(path) as f:
        src = f.read()
    m = re.search(r"SUBROUTINE\s+%s\s*\((.*?)\)\s*\n" % name, src, re.S | re.I)
    assert m, "%s not found in %s" % (name, path)
    return [a.strip().upper() for a in m.group(1).replace("&", " ").split(",")]


IMPLSCH_REF = ("KIJS KIJL FL1 WAVNUM CGROUP CIWA CINV XK2CG STOKFAC EMAXDPT DEPTH IOBND IODP IBRMEM AIRD WDWAVE CICOVER WSWAVE WSTAR USTRA "
               "VSTRA UFRIC TAUW TAUWDIR Z0M Z0B CHRNCK CITHICK NEMOUSTOKES NEMOVSTOKES NEMOSTRN NPHIEPS NTAUOC NSWH NMWP NEMOTAUX NEMOTAUY "
               "NEMOTAUICX NEMOTAUICY NEMOWSWAVE NEMOPHIF WSEMEAN WSFMEAN USTOKES VSTOKES STRNMS TAUXD TAUYD TAUOCXD TAUOCYD TAUOC TAUICX "
               "TAUICY PHIOCD PHIEPS PHIAW MIJ XLLWS").split()          # src/ecwam/implsch.F90:10-23 (ecWAM 1.5.13)
PROPAG_REF = "BLK2GLO WAVNUM CGROUP OMOSNH2KD FL1 DEPTH DELLAM1 COSPHM1 UCUR VCUR".split()   # src/ecwam/propag_wam.F90:10-11


def test_fortran_module_is_generated_from_the_header():
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.call([sys.executable, os.path.join(root, "scripts", "gen_fortran_mod.py"), "--check"]) == 0, \
        "fortran/ecwam_b200_mod.F90 is stale: run scripts/gen_fortran_mod.py"
    with open(os.path.join(root, "fortran", "ecwam_b200_mod.F90")) as f:
        mod = f.read().upper()
    for name in L.EXPORTS:      # one interface per exported function
        assert "FUNCTION %s(" % name.upper() in mod, name


def test_fortran_bodies_keep_the_reference_signatures():
    import os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert _fortran_dummies(os.path.join(root, "fortran", "implsch_b200.F90"), "IMPLSCH") == IMPLSCH_REF
    assert _fortran_dummies(os.path.join(root, "fortran", "propag_wam_b200.F90"), "PROPAG_WAM") == PROPAG_REF
    ref = "/root/reference/src/ecwam"       # present in the build container only
    if os.path.isdir(ref):
        assert _fortran_dummies(os.path.join(ref, "implsch.F90"), "IMPLSCH") == IMPLSCH_REF
        assert _fortran_dummies(os.path.join(ref, "propag_wam.F90"), "PROPAG_WAM") == PROPAG_REF
    # the C entry points take the same lists (handle first; BLK2GLO is consumed at create)
    src = re.sub(r"/\*.*?\*/", "", L._SRC, flags=re.S)
    def cargs(fn):
        m = re.search(r"int %s\((.*?)\);" % fn, src, re.S)
        return [re.split(r"[ *]", a.strip())[-1].upper() for a in m.group(1).split(",")]
    assert cargs("ecwam_b200_implsch_f") == ["H"] + IMPLSCH_REF
    assert cargs("ecwam_b200_propag_wam_f") == ["H"] + PROPAG_REF[1:]
