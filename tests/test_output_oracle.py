"""CPU tests of the oracle's NEWWIND / OUTBLOCK / WAMNORM restatement (oracle/orc_output.cpp): decomposition independence,
reproducible global norms (LLNORMWAMOUT_GLOBAL), and an independent numpy restatement of the simple parameters."""
import numpy as np
import pytest

from common import OUT_ICE, OUT_ITG, OUT_SEA, ZMISS, make_oracle, next_forcing


def run(case, npr, steps=3, **extra):
    g, o, f, fl = make_oracle(case, npr=npr, **extra)
    for _ in range(steps):
        assert o.step() == 0
    return g, o, f


def test_outblock_is_decomposition_independent_and_norms_reproducible(built):
    res = []
    for npr in (1, 3):
        g, o, f = run("o48like", npr)
        b = o.outbs(OUT_ITG, OUT_ICE, OUT_SEA)
        res.append((b, o.outwnorm(True), o.outwnorm(False)))
    np.testing.assert_array_equal(res[0][0], res[1][0])
    np.testing.assert_array_equal(res[0][1], res[1][1])          # global-order sum: bit-reproducible for any NPROC
    np.testing.assert_allclose(res[0][2], res[1][2], rtol=1e-12, atol=1e-13)
    b, wg, wl = res[0]
    ok = b[0] != ZMISS
    assert wg[0, 3] == ok.sum() and wl[0, 3] == ok.sum()
    assert wg[0, 1] == b[0][ok].min() and wg[0, 2] == b[0][ok].max()
    np.testing.assert_allclose(wg[:, 0], [c[c != ZMISS].mean() for c in b], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("case", ["o48like", "o640like"])
def test_outblock_against_numpy(built, case):
    g, o, f = run(case, 1)
    b = dict(zip(OUT_ITG, o.outbs(OUT_ITG, [0] * len(OUT_ITG), [0] * len(OUT_ITG))))
    fl = o.get_fl1()                                             # [m, k, ij]
    F, A, n = fl.shape
    fr, dfim, th = o.table("FR"), o.table("DFIM"), o.table("TH")
    delth = o.table("DELTH")[0]
    dfimofr = o.table("DFIMOFR")
    eps = 1e-33
    t2 = np.maximum(fl, eps).sum(axis=1)
    em = (t2 * dfim[:, None]).sum(0) + 0.25 * fr[-1] * delth * t2[-1]
    fm = np.maximum(em / ((t2 * dfimofr[:, None]).sum(0) + 0.2 * delth * t2[-1]), fr[0])
    np.testing.assert_allclose(b[1], 4 * np.sqrt(em), rtol=1e-12)
    np.testing.assert_allclose(b[3], 1 / fm, rtol=1e-12)
    tk = (fl * dfim[:, None, None]).sum(0)                        # [k, ij]
    thq = np.arctan2((np.sin(th)[:, None] * tk).sum(0), (np.cos(th)[:, None] * tk).sum(0)) % (2 * np.pi)
    d = np.abs(np.mod(np.degrees(thq) + 180.0, 360.0) - b[2])
    assert np.minimum(d, 360 - d).max() < 1e-8
    # wind sea + swell = total (sepwisw.F90: the two spectra partition FL1), heights 11 / 12 vs 1
    e1, es, ew = (b[1] / 4) ** 2, (b[12] / 4) ** 2, (b[11] / 4) ** 2
    assert np.abs(es + ew - e1).max() <= 1e-12 * e1.max()
    # period bands (sebtmean.F90): an independent trapezoid of the 1-D spectrum between the two cut-off frequencies
    f1d = fl.sum(axis=1) * delth                                 # [m, ij]
    def band(tb, tt):
        lo, hi = max(1.0 / tt, fr[0]), min(1.0 / tb, fr[-1])
        nodes = np.concatenate(([lo], fr[(fr > lo) & (fr < hi)], [hi]))
        vals = np.stack([np.interp(nodes, fr, f1d[:, j]) for j in range(0, n, max(1, n // 50))], axis=1)
        return np.trapezoid(vals, nodes, axis=0), slice(0, n, max(1, n // 50))
    for itg, (tb, tt) in ((64, (10., 12.)), (66, (14., 17.)), (52, (10., 1.0 / fr[0]))):
        e, sl = band(tb, tt)
        np.testing.assert_allclose((b[itg][sl] / 4.0) ** 2, e + 1e-33, rtol=1e-9, atol=1e-30, err_msg=str(itg))
    # simple copies
    np.testing.assert_array_equal(b[4], o.get_field("UFRIC"))
    np.testing.assert_array_equal(b[10], o.get_field("WSWAVE"))
    np.testing.assert_array_equal(b[77], np.maximum(-o.get_field("PHIOCD"), 0.0))
    sea = b[1] > 0.01                                           # ice-covered points hold the noise floor only (SETICE)
    assert (b[22] >= 0).all() and (b[22] <= np.sqrt(2.0)).all() and (b[6][sea] > 0).all() and (b[7] <= 0.01).all()


def test_masks_and_newwind(built):
    g, o, f = run("o48like", 2, steps=1)
    nx = next_forcing(f)
    tauw0, ws = o.get_field("TAUW"), nx["WSWAVE"]
    o.newwind(nx)
    for k in ("WSWAVE", "WDWAVE", "AIRD", "WSTAR", "CICOVER", "CITHICK"):
        np.testing.assert_array_equal(o.get_field(k), nx[k])
    cap = (1.0 / 4.0) * (8.0e-4 + 8.0e-5 * ws) * ws ** 3           # newwind.F90:133-139
    np.testing.assert_allclose(o.get_field("TAUW"), np.where(ws < 4.0, np.minimum(tauw0, cap), tauw0), rtol=1e-15)
    assert (ws < 4.0).any() and (o.get_field("TAUW") < tauw0).any()
    assert o.step() == 0
    b = o.outbs(OUT_ITG, OUT_ICE, OUT_SEA)
    ice = o.get_field("CICOVER") > 0.3
    assert ice.any() and not ice.all()
    for i, itg in enumerate(OUT_ITG):
        assert ((b[i] == ZMISS) == (ice & bool(OUT_ICE[i]))).all(), itg      # outsetwmask.F90:62-78 (IODP = 1 everywhere)


def test_mean_square_slope_parts(built):
    """Parameter 9, MEANSQS (meansqs.F90:80-100): resolved part (MEANSQS_LF, restated in numpy) + Phillips tail up to the
    gravity-capillary transition + MEANSQS_GC.  The total exceeds the resolved part, the excess is at most ALPHAPMAX times the
    logarithmic width of the unresolved range (2 HALP ln(f_gc/f_N) + the k^-4 gravity-capillary integral), and the slope grows
    with the wind speed."""
    from oracle import oracle as O
    g, o, f, fl = make_oracle("o48like")
    for _ in range(3):
        assert o.step() == 0
    b = o.outbs([9, 1, 10], [0, 0, 0], [0, 0, 0])
    mss = b[0]
    F = o.get_fl1()                                   # [m, k, ij]
    wn = o.get_field3("WAVNUM")                       # [m, ij]
    dfim = o.table("DFIM")
    lf = np.einsum("m,mi,mi->i", dfim, wn ** 2, F.sum(axis=1))
    assert (mss >= lf * (1 - 1e-12)).all()
    n = int(o.table("NWAV_GC")[0])
    xk = o.table("XK_GC")
    width = np.log(np.sqrt(9.806 * xk[-1] + 7.17e-5 * xk[-1] ** 3) / (2 * np.pi) / o.table("FR")[-1])
    assert ((mss - lf) <= 0.031 * (width + 2.0)).all() and n == xk.size
    sea = (b[1] > 0.2) & (f["CICOVER"] < 0.01)
    r = np.corrcoef(mss[sea], b[2][sea])[0, 1]
    assert r > 0.6, r
    assert 1e-4 < np.median(mss[sea]) < 0.08


@pytest.mark.parametrize("opts", [dict(), dict(llwswave=1), dict(llwswave=1, llwdwave=1), dict(lcorrel=1), dict(llwswave=1, llwdwave=1, lcorrel=1),
                                  dict(lmaskice=0), dict(lmaskice=0, liceth=1), dict(iparamci=139)])
def test_getwnd_blocking_step(built, opts):
    """WAMWND + MICEP (getwnd.F90:196-212, wamwnd.F90:120-300, micep.F90:84-240) against an independent numpy statement of the
    same rules: wind speed / direction of the components (optionally rescaled to the wave model's speed, optionally relative to
    half the surface current), WSPMIN floor, directions in [0, 2 pi); ice cover clipped to {0} U [0.01, 0.95] U {1} or taken from
    the SST; thickness 0.2 + 0.4 c when none is supplied, c * h with the 0.1 m cut-off when it is."""
    from common import make_grid, synthetic_fieldg
    from oracle import oracle as O
    g = make_grid("o48like")
    f, ii, jj = synthetic_fieldg(g)
    if opts.get("iparamci") == 139:
        f["cicover"] = 268.0 + 8.0 * np.random.default_rng(2).random(f["uwnd"].shape)
    rng = np.random.default_rng(4)
    uc, vc = rng.normal(0, 1, g.niblo), rng.normal(0, 1, g.niblo)
    r = O.getwnd_points(ii, jj, f, ucur=uc, vcur=vc, wspmin=0.3, **opts)
    pick = lambda a: a[jj - 1, ii - 1]
    u, v = pick(f["uwnd"]).copy(), pick(f["vwnd"]).copy()
    if opts.get("llwswave") and opts.get("llwdwave"):
        ws, wd = pick(f["wswave"]).copy(), pick(f["wdwave"]).copy()
        bad = ws <= 0
        ws[bad], wd[bad] = np.hypot(u, v)[bad], np.arctan2(u, v)[bad]
        wd[bad & (np.hypot(u, v) == 0)] = 0.0
        if opts.get("lcorrel"):
            u, v = ws * np.sin(wd) - 0.5 * uc, ws * np.cos(wd) - 0.5 * vc
            ws, wd = np.hypot(u, v), np.arctan2(u, v)
    else:
        if opts.get("llwswave"):
            w0, sp = pick(f["wswave"]), np.hypot(u, v)
            ok = (w0 != ZMISS) & (w0 > 0) & (sp > 0)
            u[ok], v[ok] = (u * w0 / np.where(sp > 0, sp, 1))[ok], (v * w0 / np.where(sp > 0, sp, 1))[ok]
        if opts.get("lcorrel"):
            u, v = u + 0.5 * uc, v + 0.5 * vc
        ws, wd = np.hypot(u, v), np.where(np.hypot(u, v) != 0, np.arctan2(u, v), 0.0)
    np.testing.assert_allclose(r["wswave"], np.maximum(ws, 0.3), rtol=1e-14)
    np.testing.assert_allclose(r["wdwave"], np.where(wd < 0, wd + 2 * np.pi, wd), rtol=1e-13, atol=1e-15)
    assert r["wswave"].min() >= 0.3 and r["wdwave"].min() >= 0.0 and r["wdwave"].max() < 2 * np.pi + 1e-12
    ci = pick(f["cicover"])
    if opts.get("iparamci") == 139:
        c = np.where(ci < 271.5, 1.0, 0.0)
    else:
        c = np.where((ci == ZMISS) | (ci < 0.01) | (ci > 1.01), 0.0, np.where(ci > 0.95, 1.0, ci))
    if opts.get("lmaskice", 1):
        h = np.zeros_like(c)
    elif not opts.get("liceth"):
        h = np.where(c > 0, 0.2 + 0.4 * c, 0.0)
    else:
        h = c * pick(f["cithick"])
        thin = (c > 0) & (h < 0.1)
        c, h = np.where(thin, 0.0, c), np.where(thin, 0.0, h)
    np.testing.assert_array_equal(r["cicover"], c)
    np.testing.assert_allclose(r["cithick"], h, rtol=1e-15)
    for k in ("aird", "wstar", "ustra", "vstra"):
        np.testing.assert_array_equal(r[k], pick(f[k]))


def test_kurtosis_family_in_the_oracle(built):
    """OUTBLOCK parameters 29-31, 33, 34, 57, 70-72 (KURTOSIS with PEAK_ANG, TRANSF_BFI, STAT_NL, H_MAX; kurtosis.F90:239-403) —
    oracle only so far (the product rejects them).  The spectral sums are restated in numpy (width nu = sqrt(m0 m2/m1^2 - 1),
    Goda's peakedness Qp = 2 int f E^2 df / (int E df)^2 over the bins above 40 % of the peak, angular width around the peak);
    the derived statistics must stay inside the clips of stat_nl.F90 / h_max.F90 and behave as extreme-value statistics do."""
    g, o, f, fl = make_oracle("o48like")
    for _ in range(4):
        assert o.step() == 0
    itg = [29, 30, 31, 33, 34, 57, 70, 71, 72, 1, 3]
    b = dict(zip(itg, o.outbs(itg, [0] * len(itg), [0] * len(itg))))
    F = o.get_fl1()
    fr, dfim, th = o.table("FR"), o.table("DFIM"), o.table("TH")
    delth = 2 * np.pi / th.size
    FF = F.sum(axis=1)                                           # [m, ij]
    eps = 10 * np.finfo(float).eps
    tail = FF[-1]
    m0 = eps + (FF * dfim[:, None]).sum(0) + 0.25 * fr[-1] * delth * tail
    m1 = (FF * (dfim * fr)[:, None]).sum(0) + (1.0 / 3.0) * delth * fr[-1] ** 2 * tail
    m2 = (FF * (dfim * fr ** 2)[:, None]).sum(0) + 0.5 * delth * fr[-1] ** 3 * tail
    xnu = np.sqrt(np.maximum(eps, m2 * m0 / m1 ** 2 - 1.0))
    sel = FF > 0.4 * FF.max(axis=0)[None, :]
    s40 = np.sqrt(eps) + np.where(sel, FF * dfim[:, None], 0.0).sum(0)
    s4 = np.where(sel, FF ** 2 * (2 * delth * dfim * fr)[:, None], 0.0).sum(0)
    qp = np.clip(s4 / s40 ** 2, 0.5, 15.0)
    hs = b[1]
    sea = hs > 0.05                                              # (the routine returns zeros where there is no energy)
    np.testing.assert_allclose(b[31][sea], qp[sea], rtol=1e-12)
    # number of waves in 20 minutes: N = nint(1200 * sqrt(2 pi) nu f_mean)  (kurtosis.F90:360-369)
    fmean = np.clip(m1 / m0, fr[0], fr[-1])
    np.testing.assert_array_equal(b[72][sea], np.rint(1200.0 * (2 * (2 * np.pi) / np.sqrt(2 * np.pi)) * xnu * fmean)[sea])
    assert (np.abs(b[29]) <= 0.25).all() and (b[57] >= 0).all() and (b[57] <= 0.25).all() and (np.abs(b[30]) <= 5).all()
    assert (b[71] >= 0).all() and (b[71] <= 16).all()
    r = b[33][sea] / hs[sea]
    assert r.min() >= 1.0 and r.max() <= 4.0 and 1.5 < np.median(r) < 2.3          # Hmax ~ 1.6 - 2.1 Hs for 100 - 300 waves
    assert (b[34][sea] >= (m0 / m1)[sea] * (1 - 1e-12)).all()                        # the period of the highest wave exceeds T_m01
    # deep-water narrow-band limit of the skewness: C3 -> 1.12 k_p sqrt(m0) (stat_nl.F90:128-130 with tanh = 1, Delta -> 0)
    deep = sea & (o.get_field("DEPTH") > 900.0) & (b[57] < 0.249)
    kp = (0.89 * 2 * np.pi * np.clip(m0 / ((FF * (dfim / fr)[:, None]).sum(0) + 0.2 * delth * tail), fr[0], fr[-1])) ** 2 / 9.806
    np.testing.assert_allclose(b[57][deep], 1.12 * kp[deep] * np.sqrt(m0[deep]), rtol=0.03)


def test_intpol_properties(built):
    """INTPOL (intpol.F90:96-271), the relative -> absolute frequency map OUTBLOCK applies with currents: (i) without current it is the
    identity (every output equals the IREFRA = 0 value of the same spectrum to rounding); (ii) it conserves the energy of a spectrum
    that stays inside the frequency grid (Hs within 1 % -- measured 6e-4 -- for currents of 0.8 m/s: only the f**-5 tail crosses the upper edge);
    (iii) waves running with the current are seen at a higher frequency from the ground, against it at a lower one."""
    from common import make_oracle, synthetic_currents, OUT_ITG, OUT_ICE, OUT_SEA, ZMISS
    g, o0, f, fl = make_oracle("o48like")
    g, o, f, fl = make_oracle("o48like", irefra=2)
    for _ in range(2):
        assert o0.step() == 0
    fl1 = o0.get_fl1()
    o.set_fl1(fl1)
    itg = [1, 3, 8]                       # Hs, mean period (-1 moment), peak period
    b0 = o0.outbs(itg, [0] * 3, [0] * 3)
    bz = o.outbs(itg, [0] * 3, [0] * 3)   # currents are zero so far
    np.testing.assert_allclose(bz[0], b0[0], rtol=1e-12)
    np.testing.assert_allclose(bz[1], b0[1], rtol=1e-12)
    n = g.niblo
    wd = f["WDWAVE"]
    u, v = 0.8 * np.sin(wd), 0.8 * np.cos(wd)            # 0.8 m/s along the wind (= wave) direction on the first half, against it on the second
    u[n // 2:] *= -1; v[n // 2:] *= -1
    o.set_field("UCUR", u); o.set_field("VCUR", v)
    bc = o.outbs(itg, [0] * 3, [0] * 3)
    sea = b0[0] > 0.3
    assert np.abs(bc[0][sea] / b0[0][sea] - 1.0).max() < 0.01
    with_cur, against = sea & (np.arange(n) < n // 2), sea & (np.arange(n) >= n // 2)
    assert with_cur.sum() > 20 and against.sum() > 20
    assert (bc[1][with_cur] < b0[1][with_cur]).all()          # shorter period from the ground
    assert (bc[1][against] > b0[1][against]).all()
    # Doppler shift of the mean: T_abs ~ T / (1 + k U / w) with deep-water k = w^2 / g  ->  dT/T ~ -2 pi U / (g T) (first order)
    rel = bc[1][with_cur] / b0[1][with_cur] - 1.0
    est = -2 * np.pi * 0.8 / (9.806 * b0[1][with_cur])
    assert np.abs(rel / est - 1.0).max() < 0.5
