"""GPU parity of the steps either side of the hot path (pytest -m gpu): NEWWIND, OUTBS/OUTBLOCK core parameters and the
WAMNORM statistics, through the C ABI, against the CPU oracle on the same seeded inputs.

Tolerances (written in tests/common.py compare_bout):
  * missing-value pattern, norm counts, copies of model fields vs the rank's own fields .. exact
  * heights, periods, drag, fluxes .................................................. relative 1e-10
  * mean directions (where the matching height is above 1e-3 of its maximum) ....... 1e-7 degrees
  * directional spreads sqrt(2(1-x)) ................................................ absolute 1e-7
  * WAMNORM average / minimum / maximum ............................................. relative 1e-10
(the CUDA kernel sums the frequencies from NFRE down to 1 in one sweep; the reference sums upwards, routine by routine.)
"""
import ctypes as C

import numpy as np
import pytest

from common import OUT_ICE, OUT_ITG, OUT_SEA, ZMISS, compare_bout, make_gpu, make_oracle, next_forcing
from ecwam_b200 import lib as L

pytestmark = pytest.mark.gpu


def both(case, steps=3, **extra):
    g, o, f, fl = make_oracle(case, **extra)
    _, s, w = make_gpu(case, **extra)
    for _ in range(steps):
        assert o.step() == 0 and w.step() == 0
    w.synchronize()
    return g, o, f, s, w


def check_norms(w, o):
    for glob in (True, False):
        a, b = w.outwnorm(glob), o.outwnorm(glob)
        np.testing.assert_array_equal(a[:, 3], b[:, 3])                                  # non-missing points
        scale = np.maximum(np.maximum(np.abs(b[:, 1]), np.abs(b[:, 2])), 1e-6)           # an average of signed values cancels
        circ = np.array([itg in (2, 5, 13, 14, 63) for itg in OUT_ITG])                  # directions, spreads: looser, as in
        sprd = np.array([itg in (22, 27, 28) for itg in OUT_ITG])                        # compare_bout
        tight = ~circ & ~sprd
        for c in range(3):
            d = np.abs(a[:, c] - b[:, c])
            assert (d[tight] <= 1e-10 * scale[tight]).all(), (glob, c, [OUT_ITG[i] for i in np.nonzero(tight & (d > 1e-10 * scale))[0]])
            assert (d[circ] <= 1e-5).all() and (d[sprd] <= 1e-7).all(), (glob, c)


@pytest.mark.parametrize("case,irefra,amp", [("o48like", 2, 1.0), ("o640like", 3, 1.4)])
def test_outblock_with_currents_uses_intpol(built, case, irefra, amp):
    """IREFRA = 2, 3: OUTBLOCK's output spectrum is INTPOL's, on the absolute frequency axis (outblock.F90:168-169, intpol.F90:96-271:
    k_intpol, then the CUR instance of k_outblock); SEPWISW / WEFLUX keep FL1.  Currents up to 1.4 m/s so that bins leave the grid at
    both ends and directions flip (negative absolute frequency)."""
    from common import synthetic_currents
    g, o, f, fl = make_oracle(case, irefra=irefra)
    _, o0, _, _ = make_oracle(case, irefra=irefra)
    _, s, w = make_gpu(case, irefra=irefra)
    u, v = synthetic_currents(g, amp=amp)
    o.set_field("UCUR", u); o.set_field("VCUR", v)
    w.set_field("ucur", u); w.set_field("vcur", v)
    for _ in range(3):
        assert o.step() == 0 and o0.step() == 0 and w.step() == 0
    w.synchronize()
    b = o.outbs(OUT_ITG, OUT_ICE, OUT_SEA)
    a = w.outbs(OUT_ITG, OUT_ICE, OUT_SEA)
    worst = compare_bout(a, b[:, w.own])
    assert max(worst.values()) < 1e-7
    b0 = o0.outbs(OUT_ITG, OUT_ICE, OUT_SEA)               # no current: INTPOL is (nearly) the identity, the mean period differs
    i3 = OUT_ITG.index(3)
    ok = (b[i3] != ZMISS) & (b0[i3] != ZMISS)
    assert np.abs(b[i3][ok] - b0[i3][ok]).max() > 1e-3 * np.abs(b0[i3][ok]).max()


@pytest.mark.parametrize("case,extra", [("o48like", {}), ("o48_iphys0", {}), ("o640like", {}), ("o320like", dict(lmaskice=0))])
def test_outblock_and_norms_match_oracle(built, case, extra):
    g, o, f, s, w = both(case, **extra)
    b = o.outbs(OUT_ITG, OUT_ICE, OUT_SEA)
    a = w.outbs(OUT_ITG, OUT_ICE, OUT_SEA)
    worst = compare_bout(a, b[:, w.own])
    assert (b[0] != ZMISS).sum() > 100 and max(worst.values()) < 1e-7
    # pass-through columns are bit copies of the rank's own fields (outblock.F90:236-238, 290, 404-470)
    for itg, nm in ((4, "ufric"), (10, "wswave"), (32, "depth"), (35, "ustokes"), (41, "tauoc"), (53, "aird"), (73, "tauxd")):
        col, fld = a[OUT_ITG.index(itg)], w.get_field(nm)
        keep = col != ZMISS
        np.testing.assert_array_equal(col[keep], fld[keep], err_msg=nm)
    check_norms(w, o)


def test_newwind_then_steps(built):
    g, o, f, s, w = both("o640like", steps=2)
    nx = next_forcing(f)
    tauw0 = w.get_field("tauw")
    o.newwind(nx)
    w.newwind(nx)
    for k in ("WSWAVE", "WDWAVE", "AIRD", "WSTAR", "CICOVER", "CITHICK", "USTRA", "VSTRA"):
        np.testing.assert_array_equal(w.get_field(k), o.get_field(k)[w.own], err_msg=k)   # NEWWIND copies: bit-exact
    ws = nx["WSWAVE"][w.own]
    cap = (1.0 / 4.0) * (8.0e-4 + 8.0e-5 * ws) * (ws * ws * ws)                           # newwind.F90:133-139
    np.testing.assert_allclose(w.get_field("tauw"), np.where(ws < 4.0, np.minimum(tauw0, cap), tauw0), rtol=1e-15)
    np.testing.assert_allclose(w.get_field("tauw"), o.get_field("TAUW")[w.own], rtol=1e-10)
    for _ in range(2):
        assert o.step() == 0 and w.step() == 0
    w.synchronize()
    iodp = (np.arange(g.niblo) % 11 != 0).astype(np.int32)                                # some "too shallow" points
    b = o.outbs(OUT_ITG, OUT_ICE, [0] * len(OUT_ITG))
    a = w.outbs(OUT_ITG, OUT_ICE, [0] * len(OUT_ITG), iodp=None)
    compare_bout(a, b[:, w.own])
    # OUTSETWMASK's sea mask (outsetwmask.F90:72-76) with a caller-supplied IODP
    a2 = w.outbs(OUT_ITG, OUT_ICE, OUT_SEA, iodp=iodp)
    for i, itg in enumerate(OUT_ITG):
        exp = np.where((iodp[w.own] == 0) & bool(OUT_SEA[i]), ZMISS, a[i])
        np.testing.assert_array_equal(a2[i], exp, err_msg=str(itg))


def test_unsupported_parameter_is_rejected(built):
    g, o, f, s, w = both("o48like", steps=1)
    for bad in (17, 29, 42, 51, 70, 78):
        assert w.lib.ecwam_b200_outparam_supported(bad) == 0
        with pytest.raises(L.EcwamError):
            w.outbs([1, bad], [1, 1], [1, 1])
    nx = L.ForcingNext()
    assert w.lib.ecwam_b200_newwind(w.h, C.byref(nx)) == -1                               # null FF_NEXT members


@pytest.mark.parametrize("case,nproma", [("o48like", 32), ("o48_iphys0", 24)])
def test_o48_full_size_hourly_norms(built, case, nproma):
    """BASELINE configs[0] / [1] at their real size (O48, 12 directions x 36(25) frequencies, dt = 900 s, NPROMA 32 / 24): four
    steps = the first hourly output step of tests/etopo1_oper_an_fc_O48*.yml, then the seven fields those configs request
    (swh mwd mwp pp1d dwi cdww wind) and their WAMNORM lines - the quantity the reference's `validation:` blocks record."""
    from common import CASES
    CASES["_o48"] = dict(CASES[case], N=48, nproma=nproma)
    g, o, f, s, w = both("_o48", steps=4)
    assert g.niblo > 7000
    itg = [1, 2, 3, 6, 5, 7, 10]
    ice = [1, 1, 1, 1, 0, 0, 0]
    sea = [1, 1, 1, 1, 0, 0, 0]
    b = o.outbs(itg, ice, sea)
    a = w.outbs(itg, ice, sea)
    compare_bout(a, b[:, w.own], itgs=itg)
    wa, wb = w.outwnorm(True), o.outwnorm(True)
    np.testing.assert_array_equal(wa[:, 3], wb[:, 3])
    for row, name in ((0, "swh"), (2, "mwp"), (3, "pp1d"), (5, "cdww"), (6, "wind")):
        np.testing.assert_allclose(wa[row, :3], wb[row, :3], rtol=1e-11, err_msg=name)      # average, minimum, maximum


def test_wamintgr_sequencing_idelpro_twice_idelt(built):
    """WAMINTGR / WAMODEL time bookkeeping (wamintgr.F90:92-197, wamodel.F90:228-300) with IDELPRO = 2 IDELT and a new wind
    every IDELWO = IDELPRO: per advection step PROPAG_WAM once, NEWWIND when due, IMPLSCH twice.  The oracle is driven through
    the same sequence call by call."""
    from ecwam_b200 import model as M
    extra = dict(idelpro=1800.0, delpro_lf=1800.0)
    g, o, f, fl = make_oracle("o48like", **extra)
    _, s, w = make_gpu("o48like", **extra)
    clk = M.WamClock(idelpro=1800, idelt=900, idelwo=1800)
    calls = []

    def ff_next(t):
        calls.append(t)
        return next_forcing(f, k=1 + t // 1800)

    for kadv in range(3):
        assert w.advection_step(clk, ff_next) == 0
        # the same sequence on the oracle: propagation, then (NEWWIND if CDTIMP >= CDATEWH, IMPLSCH) x 2
        assert o.propag() == 0
        for sub in range(2):
            t_imp = kadv * 1800 + sub * 900
            if t_imp >= 1800 and t_imp % 1800 == 0:
                o.newwind(next_forcing(f, k=1 + t_imp // 1800))
            o.implsch()
    w.synchronize()
    assert calls == [1800, 3600] and clk.cdtpro == 5400 and clk.cdtimp == 5400 and clk.cdatewh == 5400
    assert np.abs(w.get_spec("fl1") - o.get_fl1()[:, :, w.own]).max() <= 1e-12 * o.get_fl1().max()
    assert (w.get_field("mij") == o.get_field("MIJ")[w.own]).all()


def test_wamintgr_without_source_terms(built, monkeypatch):
    """LLSOURCE = F (wamintgr.F90:163-171) and the "not yet time" branch (:188-195): MIJ = NFRE, XLLWS = 0, FL1 floored."""
    monkeypatch.setenv("ECWAM_B200_PROPAG", "exact")
    from ecwam_b200 import model as M
    g, o, f, fl = make_oracle("o48like")
    _, s, w = make_gpu("o48like")
    clk = M.WamClock(idelpro=900, idelt=900, idelwo=10 ** 9)
    assert w.advection_step(clk, llsource=False) == 0
    assert o.propag() == 0
    w.synchronize()
    np.testing.assert_array_equal(w.get_spec("fl1"), np.maximum(o.get_fl1()[:, :, w.own], 1e-33))     # PROPAGS2 is bit-exact
    assert (w.get_field("mij") == w.F).all() and (w.get_spec("xllws") == 0).all()
    w.implsch()
    w.no_source(False)
    w.synchronize()
    assert (w.get_field("mij") == w.F).all() and (w.get_spec("xllws") == 0).all()


@pytest.mark.parametrize("parallel", [False, True])
def test_restart_round_trip_is_bit_exact(built, tmp_path, parallel):
    """SAVSPEC + SAVSTRESS after two steps, GETSPEC + GETSTRESS into a fresh handle, two more steps: the same bits as the
    uninterrupted run (the BLS / LAW files carry the whole prognostic state: spectrum, forcing, u*, tau_w, z0, Charnock)."""
    from common import make_gpu
    from oracle import restart_io as R
    g, s, w = make_gpu("o48like")
    for _ in range(2):
        assert w.step() == 0
    saved = w.get_spec("fl1")
    bls, law = str(tmp_path / "BLS20220101000000_000000003000"), str(tmp_path / "LAW20220101000000_000000003000")
    w.savspec(bls, parallel=parallel)
    w.savstress(law, "20220101003000")
    for _ in range(2):
        assert w.step() == 0
    w.synchronize()
    _, _, v = make_gpu("o48like", setup=s, grid=g)
    v.t["fl1"].zero_()
    assert v.getstress(law)[0] == "20220101003000"
    v.getspec(bls, parallel=parallel)
    for _ in range(2):
        assert v.step() == 0
    v.synchronize()
    np.testing.assert_array_equal(v.get_spec("fl1"), w.get_spec("fl1"))
    for nm in ("ufric", "tauw", "z0m", "chrnck", "phiocd", "mij"):
        np.testing.assert_array_equal(v.get_field(nm), w.get_field(nm), err_msg=nm)
    if not parallel:   # the file is the reference's global layout: original point order, one record per (direction, frequency)
        np.testing.assert_array_equal(R.readfl(bls, s.niblo, w.A, w.F)[:, :, w.own], saved)


@pytest.mark.parametrize("opts,extra", [(dict(), {}), (dict(llwswave=1), {}), (dict(llwswave=1, llwdwave=1), {}), (dict(liceth=1), dict(lmaskice=0)),
                                        (dict(iparamci=139), {}), (dict(), dict(irefra=3)), (dict(llwswave=1, llwdwave=1), dict(irefra=2))])
def test_getwnd_on_device_matches_oracle(built, opts, extra):
    """ecwam_b200_getwnd = WAMWND + MICEP (getwnd.F90:196-212) from the forcing grid to FF_NEXT on the device, then NEWWIND from those
    tensors: every option branch, with the relative-wind correction when currents are on (LRELWIND, IREFRA = 2, 3)."""
    from common import make_gpu, make_oracle, synthetic_currents, synthetic_fieldg
    from oracle import oracle as O
    g, o, f0, fl = make_oracle("o48like", wspmin=0.3, **extra)
    _, s, w = make_gpu("o48like", wspmin=0.3, **extra)
    uc = vc = None
    if extra.get("irefra", 0) >= 2:
        uc, vc = synthetic_currents(g)
        w.set_field("ucur", uc); w.set_field("vcur", vc)
        o.set_field("UCUR", uc); o.set_field("VCUR", vc)
    f, ii, jj = synthetic_fieldg(g)
    if opts.get("iparamci") == 139:
        f["cicover"] = 268.0 + 8.0 * np.random.default_rng(2).random(f["uwnd"].shape)
    out = w.getwnd(f, ii, jj, **opts)
    ref = O.getwnd_points(ii, jj, f, ucur=uc, vcur=vc, lcorrel=int(extra.get("irefra", 0) >= 2), wspmin=0.3,
                          lmaskice=extra.get("lmaskice", 1), **opts)
    for k, t in out.items():
        a = t.reshape(-1)[: w.nloc].cpu().numpy()
        if k in ("wswave", "wdwave", "cithick"):
            np.testing.assert_allclose(a, ref[k][w.own], rtol=1e-13, atol=1e-14, err_msg=k)
        else:
            np.testing.assert_array_equal(a, ref[k][w.own], err_msg=k)
    # NEWWIND straight from the device tensors == the oracle's NEWWIND fed with the oracle's GETWND
    w.newwind_device()
    w.synchronize()
    o.newwind({k.upper(): v for k, v in ref.items()})
    for k in ("wswave", "wdwave", "cicover", "cithick", "tauw", "ustra"):
        np.testing.assert_allclose(w.get_field(k), o.get_field(k.upper())[w.own], rtol=1e-13, atol=1e-14, err_msg=k)
    assert w.step() == 0 and o.step() == 0            # and the step that follows runs on that forcing
    w.synchronize()
    assert np.abs(w.get_spec("fl1") - o.get_fl1()[:, :, w.own]).max() <= 1e-10 * np.abs(o.get_fl1()).max()


def test_resident_state_step(built):
    """ecwam_b200_wamintgr_forced: the state stays on the device, a step takes the FF_NEXT forcing from host memory and hands the
    integrated 1-D fields back; it must equal NEWWIND + WAMINTGR through the device entry points, and the host copies must
    equal the device fields."""
    import ctypes as C
    import torch
    g, o, f, s, w1 = both("o48like", steps=1)
    _, _, _, _, w2 = both("o48like", steps=1)
    nx = next_forcing(f)
    w1.newwind(nx)
    assert w1.step() == 0
    w1.synchronize()
    hn, ho, keep = L.ForcingNext(), L.Fields(), {}
    for n in w2.NEXT_FIELDS:
        v = nx[n] if n in nx else nx[n.upper()]
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(v, dtype=np.float64)[w2.src])).pin_memory()
        keep[n] = t
        setattr(hn, n, C.cast(t.data_ptr(), C.POINTER(C.c_double)))
    outs = ("ufric", "tauw", "tauwdir", "z0m", "chrnck", "ustokes", "phiaw", "tauoc", "mij")
    for n in outs:
        t = torch.empty(w2.t[n].shape, dtype=w2.t[n].dtype).pin_memory()
        keep["o_" + n] = t
        setattr(ho, n, C.cast(t.data_ptr(), C.POINTER(C.c_int if n == "mij" else C.c_double)))
    hin, hout = C.c_longlong(), C.c_longlong()
    L.check(w2.lib.ecwam_b200_wamintgr_forced(w2.h, C.byref(hn), C.byref(ho), C.byref(hin), C.byref(hout)), "wamintgr_forced")
    assert hin.value == 8 * 8 * w2.P * w2.C and hout.value == (8 * (len(outs) - 1) + 4) * w2.P * w2.C
    assert torch.equal(w1.t["fl1"], w2.t["fl1"]) and torch.equal(w1.t["xllws"], w2.t["xllws"])
    for n in outs:
        assert torch.equal(keep["o_" + n], w2.t[n].cpu()), n
        assert torch.equal(w1.t[n], w2.t[n]), n
