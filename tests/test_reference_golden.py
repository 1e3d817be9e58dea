"""Parity against the REFERENCE'S OWN SOURCE.  tests/golden/ref_implsch_*.npz hold the outputs of src/ecwam/implsch.F90 and the 43 routines
below it, executed statement by statement by tests/golden/f90run.py (a Fortran-subset -> Python translator; the image has no Fortran
compiler) on 24 grid points of 14 configurations -- made in the build container by tests/golden/make_ref_golden.py, which is the only
place that reads /root/reference.  Tolerances: spectra 1e-13 (oracle) / 1e-11 (CUDA path, whose input state is its own two steps),
MIJ and XLLWS exact, integrated fields 1e-12 / 1e-10 of the field's maximum."""
import glob
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_ref_golden as G  # noqa: E402

FILES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_implsch_*.npz")))
NAMES = [os.path.basename(f)[len("ref_implsch_"):-4] for f in FILES]


def _load(name):
    z = np.load(os.path.join(HERE, "golden", "ref_implsch_%s.npz" % name))
    return z, str(z["case"]), json.loads(str(z["kw"])), int(z["steps"]), bool(int(z["hook"])), z["pts"]


def test_the_fixtures_cover_every_switch():
    assert len(NAMES) >= 14
    kws = [json.loads(str(np.load(f)["kw"])) for f in FILES]
    for key in ("lciwa1", "lciwa2", "lciwa3", "lciscal", "lwnemocou", "lwnemocouwrs", "lwnemocouibr", "lwnemocoustrn", "isnonlin", "lwvflx_snl",
                "icode", "lwflux"):
        assert any(key in k for k in kws), key


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_the_reference_source(built, name):
    z, case, kw, steps, hook, pts = _load(name)
    g, o, f = G.prepare(case, kw, steps, hook)
    o.implsch()
    a, b = o.get_fl1()[:, :, pts], z["FL1"]
    assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()
    big = b > 1e-8 * b.max()
    assert (np.abs(a - b)[big] / b[big]).max() <= 1e-12
    np.testing.assert_array_equal(o.get_xllws()[:, :, pts], z["XLLWS"])
    np.testing.assert_array_equal(o.get_field("MIJ")[pts], z["MIJ"])
    for nm in G.OUT_CHECK:
        x, y = o.get_field(nm)[pts], z[nm]
        assert np.abs(x - y).max() <= 1e-12 * max(np.abs(y).max(), 1e-30), nm
    assert z["FL1"].max() > 1e-3 and (z["MIJ"] < o.cfg.nfre).any()            # the fixture is not trivial
    if kw.get("lwnemocou"):
        assert z["NSWH"].min() > 0 and np.abs(z["NEMOTAUX"]).max() > 0 and z["STRNMS"].max() > 0 and np.abs(z["TAUICX"]).max() > 1e-6


@pytest.mark.skipif(not os.path.isdir(G.REF), reason="the reference tree exists in the build container only")
def test_the_generator_reproduces_a_fixture(built):
    """Runs the translator on the reference source again (4 points of the first case) and compares with the committed file bit for bit."""
    from f90run import FArr, FInt, Translator
    z, case, kw, steps, hook, pts = _load("ard")
    g, o, f = G.prepare(case, kw, steps, hook)
    sub = np.arange(0, len(pts), 6)
    p = pts[sub]
    T = Translator([x + ".F90" for x in G.FILES])
    ns = T.compile(["IMPLSCH"], G.namespace(o, {}))
    K, A, NF = len(p), o.cfg.nang, o.cfg.nfre
    arg = {"FL1": FArr.of(np.ascontiguousarray(o.get_fl1()[:, :, p].transpose(2, 1, 0))), "XLLWS": FArr([(1, K), (1, A), (1, NF)])}
    for nm in G.ARGS2:
        arg[nm] = FArr.of(np.ascontiguousarray(o.get_field3(nm)[:, p].T))
    for nm in G.ARGS1_IN + G.ARGS1_OUT + G.NEMO:
        arg[nm] = FArr.of(o.get_field(nm)[p])
    arg["IOBND"] = FArr.of(np.ones(K, dtype=np.int64)); arg["IODP"] = FArr.of(np.ones(K, dtype=np.int64))
    arg["MIJ"] = FArr.of(np.full(K, NF, dtype=np.int64))
    ns["IMPLSCH"](*[FInt(1) if a == "KIJS" else FInt(K) if a == "KIJL" else arg[a] for a in T.routines["IMPLSCH"].args])
    np.testing.assert_array_equal(arg["FL1"].a.transpose(2, 1, 0), z["FL1"][:, :, sub])
    np.testing.assert_array_equal(arg["MIJ"].a, z["MIJ"][sub])
    np.testing.assert_array_equal(arg["UFRIC"].a, z["UFRIC"][sub])
    assert len(T.routines) == 44


GPU_KEYS = dict(icode="icode_wnd")


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_matches_the_reference_source(built, name):
    from common import make_gpu
    z, case, kw, steps, hook, pts = _load(name)
    gkw = {GPU_KEYS.get(k, k): v for k, v in kw.items() if not k.startswith("lciwa") and k != "lciscal"}
    lciwa = (1 if kw.get("lciwa1") else 0) | (2 if kw.get("lciwa2") else 0) | (4 if kw.get("lciwa3") else 0) | (8 if kw.get("lciscal") else 0)
    if lciwa:
        gkw["lciwa"] = lciwa
    g, s, w = make_gpu(case, grid_hook=G.shelf if hook else None, **gkw)
    from ecwam_b200 import synth
    f = synth.make_forcing(g)
    n = g.niblo
    ci = f["CICOVER"]
    w.set_field("cithick", np.where(ci > 0, 0.3 + 1.5 * ci, 0.0))
    if "ibrmem" in w.t:
        w.set_field("ibrmem", ((np.arange(n) * 7) % 5 < 2) * 1.0)
    if kw.get("icode", 3) != 3:
        us = np.sqrt(8.0e-4 + 8.0e-5 * f["WSWAVE"]) * f["WSWAVE"]
        for k, v in dict(ufric=us, tauw=0.4 * us * us, tauwdir=f["WDWAVE"], chrnck=np.full_like(us, 0.018)).items():
            w.set_field(k, v)
    for _ in range(steps):
        assert w.step() == 0
    assert w.propag() == 0
    w.implsch()
    w.synchronize()
    inv = np.empty(n, dtype=np.int64)
    inv[w.own] = np.arange(n)
    q = inv[pts]
    a, b = w.get_spec("fl1")[:, :, q], z["FL1"]
    assert np.abs(a - b).max() <= 1e-11 * np.abs(b).max()
    np.testing.assert_array_equal(w.get_spec("xllws")[:, :, q], z["XLLWS"])
    np.testing.assert_array_equal(w.get_field("mij")[q], z["MIJ"])
    for nm in G.OUT_CHECK:
        if nm.lower() not in w.t:
            continue
        x, y = w.get_field(nm.lower())[q], z[nm]
        assert np.abs(x - y).max() <= 1e-10 * max(np.abs(y).max(), 1e-12), nm


TFILES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_tables_*.npz")))


@pytest.mark.parametrize("path", TFILES, ids=[os.path.basename(f)[len("ref_tables_"):-4] for f in TFILES])
def test_tables_match_the_reference_source(built, path):
    """The one-off tables (INIWCST, MFREDIR, SETWAVPHYS, INIT_X0TAUHF, INIT_SDISS_ARDH, INISNONLIN + NLWEIGT + JAFU, INITGC, CIGETDEAC run
    from their own source by the translator) against the oracle's AND the product's host builders: integer tables exact, real ones
    1e-14 (pow() against repeated multiplication)."""
    from ecwam_b200 import model as M, synth
    from oracle import oracle as O
    z = np.load(path)
    kw = json.loads(str(z["kw"]))
    g = synth.make_grid(8, "aqua")
    o = O.Oracle(O.default_config(**kw), g)
    s = M.WamSetup(g, nproc=1, **kw)
    n_checked = 0
    # DEPTHPRPT + AKI from the reference source against the product's host routine (which test_abi_and_host.py holds bit-identical to the oracle's)
    import ctypes as C
    dep = np.ascontiguousarray(z["DEPTH_IN"])
    K, NF = dep.size, 36
    dpp = C.POINTER(C.c_double)
    outp = {k: np.empty((NF, K)) for k in ("WAVNUM", "CINV", "CGROUP", "XK2CG", "OMOSNH2KD", "STOKFAC")}
    assert s.lib.ecwam_b200_host_depthprpt(C.byref(s.tables), NF, K, dep.ctypes.data_as(dpp), *[outp[k].ctypes.data_as(dpp) for k in
                                           ("WAVNUM", "CINV", "CGROUP", "XK2CG", "OMOSNH2KD", "STOKFAC")]) == 0
    for k, v in outp.items():
        ref = z["DP_" + k].T
        assert np.abs(v - ref).max() <= 1e-13 * np.abs(ref).max(), "DEPTHPRPT " + k
    for nm in z.files:
        if nm == "kw" or nm.startswith("DP_") or nm == "DEPTH_IN":
            continue
        ref = z[nm]
        kind = G.TABLE_CHECK[nm]
        got = o.itable(nm) if kind == "i" else o.table(nm)
        assert got.shape == ref.shape, nm
        if kind == "i":
            np.testing.assert_array_equal(got, ref, err_msg=nm)
        else:
            assert np.abs(got - ref).max() <= 1e-14 * max(np.abs(ref).max(), 1e-300), nm
        n_checked += 1
        # the product's own builder (ecwam_b200_host_tables_create), where the member exists under the same name
        low = nm.lower()
        if hasattr(s.tables, low):
            v = getattr(s.tables, low)
            if isinstance(v, (int, float)):
                assert abs(v - float(ref[0])) <= 1e-14 * max(abs(float(ref[0])), 1e-300), "product " + nm
            elif kind == "i":
                np.testing.assert_array_equal(s.itable(low, ref.size), ref, err_msg="product " + nm)
            elif nm not in ("CIDEAC",) or True:
                p = s.table(low, ref.size)
                assert np.abs(p - ref).max() <= 1e-14 * max(np.abs(ref).max(), 1e-300), "product " + nm
    assert n_checked >= 80


PFILES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_propag_*.npz")))


def _prop_case(path):
    from ecwam_b200 import synth
    from oracle import oracle as O
    z = np.load(path)
    kw = json.loads(str(z["kw"]))
    g = synth.make_grid(int(z["N"]), str(z["mask"]))
    return z, kw, g


@pytest.mark.parametrize("path", PFILES, ids=[os.path.basename(f)[len("ref_propag_"):-4] for f in PFILES])
def test_ctu_weights_and_propags2_match_the_reference_source(built, path):
    """CTUWUPDT + CTUWINI + CTUWDRV + CTUW + PROPAGS2 (and, with refraction, PROPDOT + GRADI) executed from their own source (every 6th
    point of a 323-point grid is kept in the fixture): the oracle's stored weights SUMWN / WLONN / WLATN / WCORN / WKPMN / WMPMN and its
    advected spectrum are BIT-IDENTICAL, for IREFRA = 0, 1, 2, 3 and with the fast-wave split (there PROPAG_WAM is
    run whole, sub-step loop and chunk copies included: all frequencies are compared)."""
    from ecwam_b200 import synth
    from oracle import oracle as O
    z, kw, g = _prop_case(path)
    c = O.default_config(store_all_weights=1, nproma=16, **kw)
    o = O.Oracle(c, g)
    f = synth.make_forcing(g)
    for k, v in f.items():
        o.set_field(k, v)
    fl = synth.jonswap_cold_start(f["WSWAVE"], f["WDWAVE"], c.nang, 36, c.nfre_red)
    o.set_fl1(fl)
    if kw.get("irefra", 0) >= 2:
        from common import synthetic_currents
        u, v = synthetic_currents(g)
        o.set_field("UCUR", u); o.set_field("VCUR", v)
    if "obs" in z.files and int(z["obs"]):      # LSUBGRID = T
        from test_oracle import _obstructions
        o.set_obstructions(*_obstructions(g.niblo, c.nfre_red))
    assert o.propag() == 0
    n, A, FR_ = g.niblo, c.nang, c.nfre_red
    sel, new2ij, m0 = z["sel"], z["new2ij"], int(z["m0"])
    for nm, shape in (("SUMWN", (n, A, FR_)), ("WLONN", (n, A, FR_, 2)), ("WLATN", (n, A, FR_, 2, 2)), ("WCORN", (n, A, FR_, 4, 2)), ("WKPMN", (n, A, FR_, 3)),
                      ("WMPMN", (n, A, FR_, 3))):
        if nm not in z.files:
            continue
        np.testing.assert_array_equal(o.rank_double(nm).reshape(shape, order="F")[sel], z[nm], err_msg=nm)
    got = o.get_fl1()[:FR_][:, :, new2ij].transpose(2, 1, 0)[sel]
    np.testing.assert_array_equal(got[:, :, m0:], z["F3"][:, :, m0:])
    assert z["WLATN"].max() > 1e-3 and z["WKPMN"].max() > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("path", PFILES, ids=[os.path.basename(f)[len("ref_propag_"):-4] for f in PFILES])
def test_cuda_propags2_matches_the_reference_source(built, monkeypatch, path):
    """The bit-exact PROPAGS2 kernel (in-kernel CTU weights) against the reference source's advected spectrum: identical bits; the default
    (tolerance-mode) kernel within 1e-13."""
    from ecwam_b200 import model as M, synth
    z, kw, g = _prop_case(path)
    sel, new2ij, m0 = z["sel"], z["new2ij"], int(z["m0"])
    cur = kw.get("irefra", 0) >= 2
    one = cur or ("obs" in z.files and int(z["obs"]))        # currents / obstructions always run the exact kernels
    for mode, exact in ((("exact", True),) if one else (("exact", True), (None, False))):
        if mode:
            monkeypatch.setenv("ECWAM_B200_PROPAG", mode)
        else:
            monkeypatch.delenv("ECWAM_B200_PROPAG", raising=False)
        s = M.WamSetup(g, nproc=1, nproma=16, **kw)
        subgrid = "obs" in z.files and int(z["obs"])
        if subgrid:
            from test_oracle import _obstructions
            s.set_obstructions(*_obstructions(g.niblo, s.par.nfre_red))
        w = M.WamIntgr(s, 0)
        w.set_static(g.depth)
        f = synth.make_forcing(g)
        for k, v in f.items():
            w.set_field(k, v)
        fl = synth.jonswap_cold_start(f["WSWAVE"], f["WDWAVE"], w.A, 36, w.Fr)
        w.set_fl1(fl)
        if cur:
            from common import synthetic_currents
            u, v = synthetic_currents(g)
            w.set_field("ucur", u); w.set_field("vcur", v)
        assert w.propag() == 0
        w.synchronize()
        n = g.niblo
        inv = np.empty(n, dtype=np.int64)
        inv[w.own] = np.arange(n)
        got = w.get_spec("fl1")[:w.Fr][:, :, inv[new2ij]].transpose(2, 1, 0)[sel]
        if exact:
            np.testing.assert_array_equal(got[:, :, m0:], z["F3"][:, :, m0:])
        else:
            assert np.abs(got - z["F3"])[:, :, m0:].max() <= 1e-13 * np.abs(z["F3"]).max()
        w.close()


CFILES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_connect_*.npz")))


@pytest.mark.parametrize("path", CFILES, ids=[os.path.basename(f)[len("ref_connect_"):-4] for f in CFILES])
def test_grid_connectivity_matches_the_reference_source(built, path):
    """PROPCONNECT (propconnect.F90: the 2 + 4 + 8 neighbours of every sea point on the irregular lat-lon grid and their interpolation
    weights) executed from its own source: KLAT, KLON, KCOR, WLAT, WCOR of the oracle AND of the product's host builder are identical."""
    from ecwam_b200 import model as M, synth
    from oracle import oracle as O
    z = np.load(path)
    g = synth.make_grid(int(z["N"]), str(z["mask"]))
    o = O.Oracle(O.default_config(nproma=16), g)
    s = M.WamSetup(g, nproc=1, nproma=16)
    d = s.decomp_arrays(0)
    n = g.niblo
    for nm, shape, kind in (("KLAT", (n, 2, 2), "i"), ("KLON", (n, 2), "i"), ("KCOR", (n, 4, 2), "i"), ("WLAT", (n, 2), "d"), ("WCOR", (n, 4), "d")):
        got = (o.itable(nm) if kind == "i" else o.rank_double(nm)).reshape(shape, order="F")
        np.testing.assert_array_equal(got, z[nm], err_msg="oracle " + nm)
        np.testing.assert_array_equal(d[nm.lower()].reshape(shape, order="F"), z[nm], err_msg="product " + nm)
    assert (z["KLAT"] == n + 1).any() and (z["KLAT"] <= n).any()          # land and sea neighbours both occur


OFILES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_outblock_*.npz")))


def _bout_close(got, ref, itgs, rtol):
    assert np.array_equal(ref == -999.0, got == -999.0), "missing-value pattern"
    ok = ref != -999.0
    for i, itg in enumerate(itgs):
        m = ok[i]
        if m.any():
            d = np.abs(got[i][m] - ref[i][m]).max() / max(np.abs(ref[i][m]).max(), 1e-300)
            assert d <= rtol, "parameter %d: %g" % (itg, d)


@pytest.mark.parametrize("path", OFILES, ids=[os.path.basename(f)[len("ref_outblock_"):-4] for f in OFILES])
def test_outblock_matches_the_reference_source(built, path):
    """OUTBLOCK (outblock.F90) and the 24 routines below it (FEMEAN, INTPOL, SEPWISW, STHQ, MWP1/2, WDIRSPREAD + PEAKFRI + SCOSFL, OUTBETA,
    WEFLUX, SE10MEAN, SEBTMEAN, MEANSQS + _GC + _LF + HALPHAP, DOMINANT_PERIOD, OUTSETWMASK, ...) executed from their own source for the 51
    parameters the product builds: the oracle's OUTBS within 1e-12 at every point and parameter, same missing-value pattern."""
    from common import OUT_ICE, OUT_ITG, OUT_SEA
    z = np.load(path)
    kw, hook, pts = json.loads(str(z["kw"])), bool(int(z["hook"])), z["pts"]
    g, o, f = G.prepare(str(z["case"]), kw, 2, hook)
    o.implsch()
    if kw.get("irefra", 0) >= 2:
        from common import synthetic_currents
        u, v = synthetic_currents(g)
        o.set_field("UCUR", u); o.set_field("VCUR", v)
    assert list(z["itg"]) == list(OUT_ITG)
    _bout_close(o.outbs(OUT_ITG, OUT_ICE, OUT_SEA)[:, pts], z["BOUT"], OUT_ITG, 1e-12)
    assert (z["BOUT"][0] > 0.1).any()


@pytest.mark.gpu
@pytest.mark.parametrize("path", OFILES, ids=[os.path.basename(f)[len("ref_outblock_"):-4] for f in OFILES])
def test_cuda_outblock_matches_the_reference_source(built, path):
    from common import OUT_ICE, OUT_ITG, OUT_SEA, compare_bout, make_gpu
    from ecwam_b200 import synth
    z = np.load(path)
    case, kw, hook, pts = str(z["case"]), json.loads(str(z["kw"])), bool(int(z["hook"])), z["pts"]
    g, s, w = make_gpu(case, grid_hook=G.shelf if hook else None, **kw)
    f = synth.make_forcing(g)
    n = g.niblo
    ci = f["CICOVER"]
    w.set_field("cithick", np.where(ci > 0, 0.3 + 1.5 * ci, 0.0))
    for _ in range(2):
        assert w.step() == 0
    assert w.propag() == 0
    w.implsch()
    if kw.get("irefra", 0) >= 2:
        from common import synthetic_currents
        u, v = synthetic_currents(g)
        w.set_field("ucur", u); w.set_field("vcur", v)
    a = w.outbs(OUT_ITG, OUT_ICE, OUT_SEA)
    inv = np.empty(n, dtype=np.int64)
    inv[w.own] = np.arange(n)
    worst = compare_bout(a[:, inv[pts]], z["BOUT"])          # the tolerances of tests/test_gpu_output.py
    assert max(worst.values()) < 1e-7


WFILES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_getwnd_*.npz")))
WNAMES = ("wswave", "wdwave", "aird", "wstar", "cicover", "cithick", "ustra", "vstra")


def _getwnd_inputs(z):
    from common import synthetic_fieldg
    from ecwam_b200 import synth
    llwswave, llwdwave, lrelwind, irefra, iparamci, liceth, lmaskice = (int(x) for x in z["opts"])
    g = synth.make_grid(8, "continents")
    f, ii, jj = synthetic_fieldg(g, with_ws=bool(llwswave or llwdwave))
    if iparamci == 139:
        f["cicover"] = 268.0 + 10.0 * np.random.default_rng(11).random(f["uwnd"].shape)
    return g, f, ii, jj, dict(llwswave=llwswave, llwdwave=llwdwave, iparamci=iparamci, liceth=liceth), lrelwind, irefra, lmaskice


@pytest.mark.parametrize("path", WFILES, ids=[os.path.basename(f)[len("ref_getwnd_"):-4] for f in WFILES])
def test_getwnd_matches_the_reference_source(built, path):
    """WAMWND (wamwnd.F90, ICODE_WND = 3: wind from components or from a previous WAM run, relative-wind correction) + MICEP (micep.F90:
    ice fraction or SST, parametric thickness, HICMIN) executed from their own source: the oracle's FF_NEXT fields are identical."""
    from oracle import oracle as O
    z = np.load(path)
    g, f, ii, jj, opts, lrelwind, irefra, lmaskice = _getwnd_inputs(z)
    got = O.getwnd_points(ii, jj, f, ucur=z["uc"], vcur=z["vc"], lcorrel=int(lrelwind and irefra >= 2), lmaskice=lmaskice, wspmin=0.3, **opts)
    for k in WNAMES:
        np.testing.assert_array_equal(got[k], z[k], err_msg=k)
    assert (z["cicover"] > 0).any() and z["wswave"].min() >= 0.3


@pytest.mark.gpu
@pytest.mark.parametrize("path", WFILES, ids=[os.path.basename(f)[len("ref_getwnd_"):-4] for f in WFILES])
def test_cuda_getwnd_matches_the_reference_source(built, path):
    from ecwam_b200 import model as M
    z = np.load(path)
    g, f, ii, jj, opts, lrelwind, irefra, lmaskice = _getwnd_inputs(z)
    s = M.WamSetup(g, nproc=1, nproma=16, wspmin=0.3, irefra=irefra, lmaskice=lmaskice)
    w = M.WamIntgr(s, 0)
    w.set_static(g.depth)
    w.set_field("ucur", z["uc"]); w.set_field("vcur", z["vc"])
    out = w.getwnd(f, ii, jj, lrelwind=lrelwind, **opts)
    for k in WNAMES:
        a = out[k].reshape(-1)[: w.nloc].cpu().numpy()
        np.testing.assert_allclose(a, z[k][w.own], rtol=1e-13, atol=1e-14, err_msg=k)


NFILES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_newwind_*.npz")))
NOW = ("WSWAVE", "WDWAVE", "AIRD", "WSTAR", "CICOVER", "CITHICK", "USTRA", "VSTRA", "UFRIC", "TAUW")


@pytest.mark.parametrize("path", NFILES, ids=[os.path.basename(f)[len("ref_newwind_"):-4] for f in NFILES])
def test_newwind_matches_the_reference_source(built, path):
    """NEWWIND (newwind.F90:105-167) executed from its own source, the 10 m wind branch with the low-wind cap of TAUW and the
    friction-velocity branch: the oracle's FF_NOW fields are identical."""
    z = np.load(path)
    icode = int(z["icode"])
    g, o, f = G.prepare("o48like", dict(icode=icode) if icode != 3 else {}, 2, False)
    o.newwind({k[5:]: z[k] for k in z.files if k.startswith("NEXT_")})
    for k in NOW:
        np.testing.assert_array_equal(o.get_field(k), z[k], err_msg=k)
    assert (z["TAUW"] == 0).any() if icode != 3 else (z["WSWAVE"] < 4.0).any()


@pytest.mark.gpu
@pytest.mark.parametrize("path", NFILES, ids=[os.path.basename(f)[len("ref_newwind_"):-4] for f in NFILES])
def test_cuda_newwind_matches_the_reference_source(built, path):
    from common import make_gpu
    from ecwam_b200 import synth
    z = np.load(path)
    icode = int(z["icode"])
    kw = dict(icode_wnd=icode) if icode != 3 else {}
    g, s, w = make_gpu("o48like", **kw)
    f = synth.make_forcing(g)
    ci = f["CICOVER"]
    w.set_field("cithick", np.where(ci > 0, 0.3 + 1.5 * ci, 0.0))
    if icode != 3:
        us = np.sqrt(8.0e-4 + 8.0e-5 * f["WSWAVE"]) * f["WSWAVE"]
        for k, v in dict(ufric=us, tauw=0.4 * us * us, tauwdir=f["WDWAVE"], chrnck=np.full_like(us, 0.018)).items():
            w.set_field(k, v)
    for _ in range(2):
        assert w.step() == 0
    assert w.propag() == 0
    w.newwind({k[5:]: z[k] for k in z.files if k.startswith("NEXT_")})
    for k in NOW:
        a, b = w.get_field(k.lower()), z[k][w.own]
        if k in ("TAUW", "UFRIC", "WSWAVE"):      # products of the GPU's own two steps (1e-15 apart from the oracle's) enter TAUW / the cap
            assert np.abs(a - b).max() <= 1e-10 * max(np.abs(b).max(), 1e-12), k
        else:
            np.testing.assert_array_equal(a, b, err_msg=k)


DFILES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_decomp_*.npz")))


@pytest.mark.parametrize("path", DFILES, ids=[os.path.basename(f)[len("ref_decomp_"):-4] for f in DFILES])
def test_decomposition_matches_the_reference_source(built, path):
    """MPDECOMP's sector decomposition (mpdecomp.F90: NXDECOMP x NYDECOMP sectors with staggered bands, NSTART / NEND, the relabelling
    IJ2NEWIJ of the sea points and the relabelled BLK2GLO) executed from its own source for 1 - 8 ranks and the 1-D variant: the oracle's and
    the product's tables are identical."""
    from ecwam_b200 import model as M, synth
    from oracle import oracle as O
    z = np.load(path)
    npr, ll1d = int(z["npr"]), int(z["ll1d"])
    g = synth.make_grid(int(z["N"]), str(z["mask"]))
    n = g.niblo
    o = O.Oracle(O.default_config(nproma=16, npr=npr, ll1d=ll1d), g)
    s = M.WamSetup(g, nproc=npr, ll1d=bool(ll1d), nproma=16)
    np.testing.assert_array_equal(o.itable("NSTART")[:npr], z["NSTART"]); np.testing.assert_array_equal(s.nstart, z["NSTART"])
    np.testing.assert_array_equal(o.itable("NEND")[:npr], z["NEND"]); np.testing.assert_array_equal(s.nend, z["NEND"])
    np.testing.assert_array_equal(o.itable("IJ2NEWIJ")[1:n + 1], z["IJ2NEWIJ"]); np.testing.assert_array_equal(s.ij2new[1:n + 1], z["IJ2NEWIJ"])
    np.testing.assert_array_equal(o.itable("KXLT")[:n], z["KXLT"]); np.testing.assert_array_equal(s.kxlt[:n], z["KXLT"])
    np.testing.assert_array_equal(o.itable("IXLG")[:n], z["IXLG"])


HFILES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_halo_*.npz")))


@pytest.mark.parametrize("path", HFILES, ids=[os.path.basename(f)[len("ref_halo_"):-4] for f in HFILES])
def test_halo_lists_match_the_reference_source(built, path):
    """The rest of MPDECOMP's structured-grid branch (mpdecomp.F90:696-1345) executed from its own source for every rank of a 2 / 3 / 4 / 8
    rank run, the two MPL_ALLGATHERVs emulated by running all ranks twice: NINF / NSUP, KLENBOT / KLENTOP, the receive and send lists
    NFROMPE / NTOPE / NIJSTART / IJTOPE and the local addressing of KLAT / KLON / KCOR (PROPCONNECT for the rank, halo points appended below
    and above the own block, land -> NSUP+1) are identical in the oracle and in the product's host decomposition."""
    from ecwam_b200 import model as M, synth
    from oracle import oracle as O
    z = np.load(path)
    npr = int(z["npr"])
    g = synth.make_grid(int(z["N"]), str(z["mask"]))
    o = O.Oracle(O.default_config(nproma=16, npr=npr), g)
    s = M.WamSetup(g, nproc=npr, nproma=16)
    total_sent = total_recv = 0
    for r in range(npr):
        d = s.decomp_arrays(r)
        for nm in ("NINF", "NSUP"):
            assert int(o.itable(nm, r)[0]) == int(z["%s_%d" % (nm, r + 1)]) == d[nm.lower()], (nm, r)
        assert int(o.itable("NTOPEMAX", r)[0]) == int(z["NTOPEMAX_%d" % (r + 1)])
        for nm in ("KLENBOT", "KLENTOP"):
            np.testing.assert_array_equal(o.itable(nm, r), z["%s_%d" % (nm, r + 1)], err_msg="%s rank %d" % (nm, r))
        for nm in ("NTOPE", "NFROMPE", "NIJSTART", "KLAT", "KLON", "KCOR"):
            ref = z["%s_%d" % (nm, r + 1)]
            np.testing.assert_array_equal(o.itable(nm, r), ref, err_msg="oracle %s rank %d" % (nm, r))
            np.testing.assert_array_equal(d[nm.lower()], ref, err_msg="product %s rank %d" % (nm, r))
        ref = z["IJTOPE_%d" % (r + 1)]                      # IJTOPE(NTOPEMAX, NPROC): only the first NTOPE(ip) rows of a column are defined
        nmax = int(z["NTOPEMAX_%d" % (r + 1)])
        got_o, got_p = o.itable("IJTOPE", r), d["ijtope"]
        for ip in range(npr):
            k = int(z["NTOPE_%d" % (r + 1)][ip])
            np.testing.assert_array_equal(got_o[ip * nmax: ip * nmax + k], ref[ip * nmax: ip * nmax + k])
            np.testing.assert_array_equal(got_p[ip * d["ntopemax"]: ip * d["ntopemax"] + k], ref[ip * nmax: ip * nmax + k])
        total_sent += int(z["NTOPE_%d" % (r + 1)].sum()); total_recv += int(z["NFROMPE_%d" % (r + 1)].sum())
    assert total_sent == total_recv > 0


WFILES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_wnorm_*.npz")))


@pytest.mark.parametrize("path", WFILES, ids=[os.path.basename(f)[len("ref_wnorm_"):-4] for f in WFILES])
def test_wamnorm_matches_the_reference_source(built, path):
    """MPMINMAXAVG (mpminmaxavg.F90:68-195) executed from its own source on every rank of a 1 / 2 / 3 / 5-rank and a 1-D 4-rank run, both
    flavours (LLNORMWAMOUT_GLOBAL = T with MPGATHERSCFLD, = F with the three MPL_ALLREDUCEs, both emulated over the ranks): average,
    minimum, maximum and count of 8 OUTBS columns with missing values under the ice mask are identical to the oracle's OUTWNORM."""
    z = np.load(path)
    g, o, b = G.wnorm_state(int(z["npr"]), int(z["ll1d"]))
    np.testing.assert_array_equal(b, z["BOUT"])
    count = z["WNORM_GLOBAL"][:, 3]
    assert (count < g.niblo).any() and (count > 0).all()          # the masked columns miss points
    np.testing.assert_array_equal(o.outwnorm(True), z["WNORM_GLOBAL"])
    np.testing.assert_array_equal(o.outwnorm(False), z["WNORM_LOCAL"])
    np.testing.assert_array_equal(z["WNORM_GLOBAL"][:, 1:], z["WNORM_LOCAL"][:, 1:])      # min / max / count do not depend on the flavour


SFILES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_sequence_*.npz")))


class _Recorder:
    """The time bookkeeping of ecwam_b200.model.WamIntgr (wamintgr / advection_step, unbound) around recording stand-ins of the kernels."""

    def __init__(self):
        self.log = []

    def propag(self):
        self.log.append((1, 0, 0, 0, 0)); return 0

    def implsch(self):
        self.log.append((2, 0, 0, 0, 0))

    def newwind(self, t):
        self.log.append((3, t, 0, 0, 0))

    def no_source(self, off):
        pass

    def wamintgr(self, clk, ff_next=None, llsource=True):
        from ecwam_b200 import model as M
        r = M.WamIntgr.wamintgr(self, clk, ff_next, llsource)
        self.log.append((4, clk.cdate, clk.cdatewh, clk.cdtimp, clk.cdtimpnext))
        return r


@pytest.mark.parametrize("path", SFILES, ids=[os.path.basename(f)[len("ref_sequence_"):-4] for f in SFILES])
def test_time_stepping_matches_the_reference_source(path):
    """WAMODEL's ADVECTION loop body + WAMINTGR + NEWWIND executed from their own source (make_ref_golden.run_sequence): the order of
    PROPAG_WAM, new forcing and IMPLSCH and the dates CDATE / CDATEWH / CDTIMP / CDTIMPNEXT after every WAMINTGR call, for IDELPRO equal
    to, a multiple of and a fraction of IDELT and three forcing intervals -- the product's WamClock / wamintgr / advection_step
    bookkeeping produces the same sequence."""
    from ecwam_b200 import model as M
    z = np.load(path)
    idelpro, idelt, idelwo, nadv, llsource = (int(v) for v in z["cfg"])
    clk = M.WamClock(idelpro=idelpro, idelt=idelt, idelwo=idelwo)
    rec = _Recorder()
    for _ in range(nadv):
        M.WamIntgr.advection_step(rec, clk, ff_next=lambda t: t, llsource=bool(llsource))
        rec.log.append((5, clk.cdtpro, 0, 0, 0))
    got, ref = np.array(rec.log, dtype=np.int64), z["EVENTS"]
    assert got.shape == ref.shape, (got[:12], ref[:12])
    bad = np.nonzero((got != ref).any(axis=1))[0]
    assert bad.size == 0, (bad[0], got[max(bad[0] - 3, 0): bad[0] + 2], ref[max(bad[0] - 3, 0): bad[0] + 2])


@pytest.mark.skipif(not os.path.isdir(G.REF), reason="the reference tree exists in the build container only")
def test_the_generator_reproduces_the_time_stepping_fixtures(tmp_path, monkeypatch):
    """Runs WAMODEL's loop + WAMINTGR + NEWWIND from the reference source again and compares with the committed event lists."""
    monkeypatch.setattr(G, "HERE", str(tmp_path))
    for name in ("pro1800_src900_wind1800", "pro900_src1800_wind1800"):
        ev = G.run_sequence(name, *G.SEQ_CASES[name])
        np.testing.assert_array_equal(ev, np.load(os.path.join(HERE, "golden", "ref_sequence_%s.npz" % name))["EVENTS"])
