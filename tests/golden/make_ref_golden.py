"""Golden vectors from the REFERENCE'S OWN SOURCE (run in the build container only: needs /root/reference).

`f90run.Translator` turns src/ecwam/implsch.F90 and the 43 routines below it into Python statement by statement and executes them
on a handful of grid points of a synthetic case; the outputs go to tests/golden/ref_implsch_<case>.npz.  tests/test_reference_golden.py
then checks the oracle (and, on a GPU, the CUDA path) against those files.  What is NOT taken from the reference source: the values of the
module variables (tables of YOWFRED / YOWINDN / ..., read from the oracle, whose table builders are checked bit for bit against the
product's independent host builders) and the inputs (the synthetic state of tests/common.py after two steps + PROPAG_WAM).

The other families, same method (each runner's docstring says what is run and what is emulated):
  ref_tables_*    the one-off table builders incl. TABU_SWELLFT + KERKEI + KZEONE          ref_propag_*   CTUW + PROPAGS2, IREFRA 0-3, LSUBGRID
  ref_connect_*   PROPCONNECT                                                              ref_outblock_* OUTBLOCK and the 24 routines below it
  ref_getwnd_*    WAMWND + MICEP                                                           ref_newwind_*  NEWWIND's field update
  ref_decomp_*    MPDECOMP's sector decomposition                                          ref_halo_*     MPDECOMP's halo lists (allgathers emulated)
  ref_wnorm_*     MPMINMAXAVG, both flavours (gather / allreduce emulated)                 ref_sequence_* WAMODEL loop + WAMINTGR + NEWWIND dates

usage: python tests/golden/make_ref_golden.py [<implsch case> | tables | propag[:name] | connect | outblock | getwnd | newwind | decomp | halo |
                                               wnorm | sequence] ...        (no argument: everything)
"""
import os
import re
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)

from f90run import REF, FArr, FInt, Translator  # noqa: E402

FILES = ("implsch sdepthlim semean fkmean sinflx airsea taut_z0 z0wave chnkmin sinput sinput_ard sinput_jan wsigstar femeanws frcutindex "
         "stresso tau_phi_hf omegagc ns_gc stress_gc halphap femean meansqs_lf sdissip sdissip_ard sdissip_jan snonlin transf transf_snl "
         "peak_ang sdiwbk sbottom sdice sdice1 sdice2 sdice3 icebreak_modify_attenuation wnfluxes imphftail setice stokestrn stokesdrift "
         "cimsstrn aki_ice").split()
MODULES = "yowfred yowpcons yowphys yowice yowcoup yowaltas yowshal yowtabl yowwind yowcurr yowparam yowstat yowindn yowcout".split()

# IMPLSCH's dummy arguments (implsch.F90:10-23) and what they are
ARGS3 = ("FL1", "XLLWS")
ARGS2 = ("WAVNUM", "CGROUP", "CIWA", "CINV", "XK2CG", "STOKFAC")
ARGS1_IN = ("EMAXDPT", "DEPTH", "IBRMEM", "AIRD", "WDWAVE", "CICOVER", "WSWAVE", "WSTAR", "USTRA", "VSTRA", "UFRIC", "TAUW", "TAUWDIR", "Z0M",
            "Z0B", "CHRNCK", "CITHICK")
NEMO = ("NEMOUSTOKES", "NEMOVSTOKES", "NEMOSTRN", "NPHIEPS", "NTAUOC", "NSWH", "NMWP", "NEMOTAUX", "NEMOTAUY", "NEMOTAUICX", "NEMOTAUICY",
        "NEMOWSWAVE", "NEMOPHIF")
ARGS1_OUT = ("WSEMEAN", "WSFMEAN", "USTOKES", "VSTOKES", "STRNMS", "TAUXD", "TAUYD", "TAUOCXD", "TAUOCYD", "TAUOC", "TAUICX", "TAUICY", "PHIOCD",
             "PHIEPS", "PHIAW")
OUT_CHECK = ("UFRIC", "TAUW", "TAUWDIR", "Z0M", "Z0B", "CHRNCK", "WSWAVE") + ARGS1_OUT + NEMO


def module_parameters():
    """PARAMETER constants of the YOW* modules, evaluated from their declarations (yowfred.F90: FRIC, WP2TAIL, ...)."""
    from f90run import RUNTIME, _logical_lines, _split_top
    ns = dict(RUNTIME)
    ns.update(JWRB=8, JWRU=8, JWIM=4, JWRO=8, JPHOOK=8)
    tr = Translator([])
    dummy = type("R", (), dict(arrays=set(), stmtfun=set(), name="MODULE"))()
    out = {}
    for m in MODULES:
        p = os.path.join(REF, m + ".F90")
        if not os.path.exists(p):
            continue
        for ln in _logical_lines(p, [REF]):
            mm = re.match(r"^(REAL|INTEGER|LOGICAL)\s*(\((?:[^()]|\([^()]*\))*\))?\s*(.*?)::\s*(.*)$", ln)
            if not mm or "DIMENSION" in mm.group(3):      # PARAMETERs and initialised module variables (G = 9.806, CIRC, ...)
                continue
            for ent in _split_top(mm.group(4)):
                em = re.match(r"^(\w+)\s*=\s*(.*)$", ent)
                if not em:
                    continue
                try:
                    v = eval(tr.expr(dummy, em.group(2)), ns)
                except Exception:
                    continue
                ns[em.group(1)] = v
                out[em.group(1)] = v
    return out


def namespace(o, cfg_extra):
    """Module variables for the translated routines: constants from the module sources, tables from the oracle, switches from its Config."""
    c = o.cfg
    ns = module_parameters()
    I = lambda v: FInt(int(v))
    F = lambda n: float(o.table(n)[0])
    for n in ("G GM1 ZPI ROWATER ROWATERM1 ZPI4GM1 ZPI4GM2 EPSMIN EPSUS EPSU10 ACD BCD CDMAX DKMAX TAUOCMIN TAUOCMAX PHIEPSMIN PHIEPSMAX "
              "WSEMEAN_MIN FRATIO WETAIL FRTAIL WP1TAIL WP2TAIL DELTH FLOGSPRDM1 XKAPPA XNLEV ZALP TAILFACTOR TAILFACTOR_PM SWELLF SWELLF2 "
              "SWELLF3 SWELLF4 SWELLF5 SWELLF6 SWELLF7 SWELLF7M1 ABMIN ABMAX SDSBR SSDSC2 SSDSC3 SSDSC4 SSDSC5 SSDSC6 MICHE SSDSBRF1 BRKPBCOEF "
              "EGRCRV AFCRV BFCRV SURFT EPS1 TICMIN HICMIN DTIC DHIC X0TAUHF BETAMAXOXKAPPA2 TAUWSHELTER ALPHA ALPHAMIN ALPHAMAX ALPHAPMAX "
              "CHNKMIN_U ACDLIN BCDLIN BMAXOKAP GAMNCONST RN1_RN DTHRN_A DTHRN_U ANG_GC_A ANG_GC_B ANG_GC_C SQRTGOSURFT Z0RAT Z0TUBMAX CDIS "
              "DELTA_SDIS CDISVIS DAL1 DAL2").split():
        ns[n] = F(n)
    for n in "ISDSDTH ISB IPSAT IAB JTOT_TAUHF NICT NICH NWAV_GC".split():
        ns[n] = I(F(n))
    for n in "NSDSNTH MFRSTLW MLSTHG KFRH NFRE_ODD".split():
        ns[n] = I(o.itable(n)[0])
    ns.update(NANG=I(c.nang), NFRE=I(c.nfre), NFRE_RED=I(c.nfre_red), IPHYS=I(c.iphys), ISNONLIN=I(c.isnonlin), IDAMPING=I(c.idamping),
              ICODE=I(c.icode), ICODE_CPL=I(c.icode), IDELT=I(int(c.idelt)), XIMP=float(c.ximp), RNU=float(c.rnu), RNUM=float(c.rnum), WSPMIN=float(c.wspmin),
              CITHRSH=float(c.cithrsh), CITHRSH_TAIL=float(c.cithrsh_tail), CIBLOCK=float(c.ciblock), FLMIN=float(c.flmin), BATHYMAX=float(c.bathymax),
              ZALPFACX=float(c.zalpfacx), ZALPFACB=float(c.zalpfacb), CDICWA=float(c.cdicwa), ZALPWRS=float(c.zalpwrs), ZIBRW_THRSH=float(c.zibrw_thrsh),
              LBIWBK=bool(c.lbiwbk), LICERUN=bool(c.licerun), LMASKICE=bool(c.lmaskice), LWAMRSETCI=bool(c.lwamrsetci), LCIWA1=bool(c.lciwa1),
              LCIWA2=bool(c.lciwa2), LCIWA3=bool(c.lciwa3), LCISCAL=bool(c.lciscal), LLGCBZ0=bool(c.llgcbz0), LLNORMAGAM=bool(c.llnormagam),
              LLCAPCHNK=bool(c.llcapchnk), LLUNSTR=False, LWCOU=bool(c.lwcou), LWCOUAST=bool(c.lwcouast), LWFLUX=bool(c.lwflux),
              LWFLUXOUT=bool(c.lwfluxout), LWNEMOCOU=bool(c.lwnemocou), LWNEMOCOUIBR=bool(c.lwnemocouibr), LWNEMOCOUSEND=bool(c.lwnemocousend),
              LWNEMOCOUSTK=bool(c.lwnemocoustk), LWNEMOCOUSTRN=bool(c.lwnemocoustrn), LWNEMOCOUWRS=bool(c.lwnemocouwrs),
              LWNEMOTAUOC=bool(c.lwnemotauoc), LWVFLX_SNL=bool(c.lwvflx_snl), LHOOK=False)
    ns.update(cfg_extra)
    A, NF = int(c.nang), int(c.nfre)
    for n in "FR DFIM DFIMOFR DFIMFR DFIMFR2 ZPIFR FR5 COFRM4 FLMAX RHOWG_DFIM DFIM_SIM".split():
        ns[n] = FArr.of(o.table(n)[:NF])
    for n in "TH COSTH SINTH".split():
        ns[n] = FArr.of(o.table(n)[:A])
    ns["SWELLFT"] = FArr.of(o.table("SWELLFT"))
    ns["WTAUHF"] = FArr.of(o.table("WTAUHF"))
    ns["CIDEAC"] = FArr.of(o.table("CIDEAC").reshape((int(ns["NICT"]), int(ns["NICH"])), order="F"))
    ng = int(ns["NWAV_GC"])
    if ng > 0:
        for n in "XK_GC XKM_GC OMEGA_GC CM_GC C2OSQRTVG_GC XKMSQRTVGOC2_GC OM3GMKM_GC OMXKM3_GC DELKCC_GC_NS DELKCC_OMXKM3_GC DELKCC_GC".split():
            ns[n] = FArr.of(o.table(n)[:ng])
    if c.iphys == 1:
        ns2 = 2 * int(ns["NSDSNTH"]) + 1
        ns["INDICESSAT"] = FArr.of(o.itable("INDICESSAT").reshape((A, ns2), order="F"))
        ns["SATWEIGHTS"] = FArr.of(o.table("SATWEIGHTS").reshape((A, ns2), order="F"))
    else:    # SDISSIP_ARD is translated with the rest of the tree but never called: its tables only have to exist
        ns["INDICESSAT"] = FArr.of(np.ones((A, 1), dtype=np.int64)); ns["SATWEIGHTS"] = FArr.of(np.zeros((A, 1)))
    lo, hi = int(ns["MFRSTLW"]), int(ns["MLSTHG"])
    for n in "IKP IKP1 IKM IKM1".split():
        ns[n] = FArr.of(o.itable(n), lb=[lo])
    for n in "AF11 FKLAP FKLAP1 FKLAM FKLAM1".split():
        ns[n] = FArr.of(o.table(n), lb=[lo])
    for n in "K1W K2W K11W K21W".split():
        ns[n] = FArr.of(o.itable(n).reshape((A, 2), order="F"))
    ns["INLCOEF"] = FArr.of(o.itable("INLCOEF").reshape((5, hi), order="F"))
    ns["RNLCOEF"] = FArr.of(o.table("RNLCOEF").reshape((25, hi), order="F"))
    return ns


def shelf(g):
    """a 3 - 80 m shelf in the northern half: depth-limited points (SDEPTHLIM, SDIWBK), finite-depth DIA scaling, bottom friction"""
    n = g.depth.size
    g.depth[n // 2:] = 3.0 + 77.0 * ((np.arange(n - n // 2) * 29) % 97) / 96.0


def prepare(case, kw, steps, hook):
    """The state IMPLSCH starts from: `steps` WAMINTGR steps of the synthetic case + PROPAG_WAM (shared with tests/test_reference_golden.py)."""
    from common import make_oracle
    g, o, f, fl = make_oracle(case, grid_hook=shelf if hook else None, **kw)
    n = g.niblo
    ci = f["CICOVER"]
    o.set_field("CITHICK", np.where(ci > 0, 0.3 + 1.5 * ci, 0.0))
    o.set_field("IBRMEM", ((np.arange(n) * 7) % 5 < 2) * 1.0)
    if kw.get("icode", 3) != 3:
        us = np.sqrt(8.0e-4 + 8.0e-5 * f["WSWAVE"]) * f["WSWAVE"]
        for k, v in dict(UFRIC=us, TAUW=0.4 * us * us, TAUWDIR=f["WDWAVE"], CHRNCK=np.full_like(us, 0.018)).items():
            o.set_field(k, v)
    for _ in range(steps):
        assert o.step() == 0
    assert o.propag() == 0
    return g, o, f


def pick_points(g, o, f, k=12):
    """a few points of every kind: strongest / weakest wind, most ice, ice edge, shallowest, deepest, highest / lowest waves + a regular stride"""
    n = g.niblo
    hs = o.hs_fm()[0]
    cand = []
    for key, arr in (("wind", f["WSWAVE"]), ("ice", f["CICOVER"]), ("depth", -g.depth), ("hs", hs)):
        order = np.argsort(arr)
        cand += [int(order[-1]), int(order[-2]), int(order[0])]
    edge = np.nonzero((f["CICOVER"] > 0.05) & (f["CICOVER"] < 0.5))[0]
    cand += [int(x) for x in edge[:2]]
    cand += [int(x) for x in np.linspace(5, n - 7, k).astype(int)]
    out = []
    for c in cand:
        if c not in out:
            out.append(c)
    return np.array(sorted(out))


def run_case(name, case, pts=None, steps=2, hook=False, **kw):
    g, o, f = prepare(case, kw, steps, hook)
    pts = pick_points(g, o, f) if pts is None else np.asarray(pts)
    K = len(pts)
    A, NF = o.cfg.nang, o.cfg.nfre
    t0 = time.time()
    T = Translator([x + ".F90" for x in FILES])
    ns = T.compile(["IMPLSCH"], namespace(o, {}))
    missing = [k for k in T.needed(["IMPLSCH"]) if k not in ns and k not in ("ENVIRONMENT", "FORCING_FIELDS", "FREQUENCY", "INTGT_PARAM_FIELDS", "WAVE2OCEAN")]
    assert not missing, missing
    arg = {}
    fl1 = o.get_fl1()[:, :, pts]                       # [m, k, ij]
    arg["FL1"] = FArr.of(np.ascontiguousarray(fl1.transpose(2, 1, 0)))
    arg["XLLWS"] = FArr([(1, K), (1, A), (1, NF)])
    for nm in ARGS2:
        arg[nm] = FArr.of(np.ascontiguousarray(o.get_field3(nm)[:, pts].T))
    for nm in ARGS1_IN + ARGS1_OUT + NEMO:
        arg[nm] = FArr.of(o.get_field(nm)[pts])
    arg["IOBND"] = FArr.of(np.ones(K, dtype=np.int64)); arg["IODP"] = FArr.of(np.ones(K, dtype=np.int64))
    arg["MIJ"] = FArr.of(np.full(K, NF, dtype=np.int64))
    inputs = {k: v.a.copy() for k, v in arg.items()}
    order = T.routines["IMPLSCH"].args
    ns["IMPLSCH"](*[FInt(1) if a == "KIJS" else FInt(K) if a == "KIJL" else arg[a] for a in order])
    print("%s: the reference source ran on %d points in %.1f s" % (name, K, time.time() - t0))
    import json
    out = dict(case=case, pts=pts, steps=steps, hook=int(hook), kw=json.dumps(kw, sort_keys=True))
    out["FL1"] = arg["FL1"].a.transpose(2, 1, 0); out["XLLWS"] = arg["XLLWS"].a.transpose(2, 1, 0); out["MIJ"] = arg["MIJ"].a
    for nm in OUT_CHECK:
        out[nm] = arg[nm].a
    # how far the oracle is from it (printed, and asserted by the test)
    o.implsch()
    a, b = o.get_fl1()[:, :, pts], out["FL1"]
    print("   FL1 max rel %.2e   MIJ equal %s   XLLWS equal %s" % (np.abs(a - b).max() / np.abs(b).max(), (o.get_field("MIJ")[pts] == out["MIJ"]).all(),
                                                                   (o.get_xllws()[:, :, pts] == out["XLLWS"]).all()))
    for nm in OUT_CHECK:
        x, y = o.get_field(nm)[pts], out[nm]
        print("   %-12s %.2e" % (nm, np.abs(x - y).max() / max(np.abs(y).max(), 1e-300)))
    np.savez_compressed(os.path.join(HERE, "ref_implsch_%s.npz" % name), **out)


TABLE_FILES = "depthprpt aki iniwcst mfredir mfr setwavphys init_x0tauhf init_sdiss_ardh inisnonlin nlweigt jafu initgc cigetdeac tabu_swellft kerkei kzeone".split()
TABLE_MODULES = MODULES + ["yowgridgen"]
# name in the oracle -> kind (t = real table, i = integer table, s = scalar)
TABLE_CHECK = dict(SWELLFT="t", DFIMOFR="t", DFIMFR="t", DFIMFR2="t", ZPIFR="t", FR5="t", COFRM4="t", FLMAX="t", RHOWG_DFIM="t", DFIM_SIM="t", FLOGSPRDM1="s",
                   FR="t", DFIM="t", GOM="t", TH="t", COSTH="t", SINTH="t", X0TAUHF="s", WTAUHF="t", SATWEIGHTS="t", INDICESSAT="i", IKP="i", IKP1="i",
                   IKM="i", IKM1="i", K1W="i", K2W="i", K11W="i", K21W="i", AF11="t", FKLAP="t", FKLAP1="t", FKLAM="t", FKLAM1="t", FRH="t",
                   INLCOEF="i", RNLCOEF="t", FTRF="t", DAL1="s", DAL2="s", ACL1="s", ACL2="s", CL11="s", CL21="s", XK_GC="t", OMEGA_GC="t",
                   CM_GC="t", C2OSQRTVG_GC="t", XKMSQRTVGOC2_GC="t", OM3GMKM_GC="t", OMXKM3_GC="t", DELKCC_GC_NS="t", DELKCC_OMXKM3_GC="t",
                   CIDEAC="t", ZPI="s", R="s", BETAMAX="s", BETAMAXOXKAPPA2="s", ZALP="s", ALPHA="s", ALPHAMIN="s", ALPHAPMAX="s", CHNKMIN_U="s",
                   TAUWSHELTER="s", TAILFACTOR="s", TAILFACTOR_PM="s", SWELLF="s", SWELLF2="s", SWELLF3="s", SWELLF4="s", SWELLF5="s", SWELLF6="s",
                   SWELLF7="s", Z0RAT="s", Z0TUBMAX="s", CDIS="s", DELTA_SDIS="s", CDISVIS="s", SDSBR="s", SSDSC2="s", SSDSC4="s", SSDSC6="s", MICHE="s",
                   EGRCRV="s", AFCRV="s", BFCRV="s", BMAXOKAP="s", GAMNCONST="s", RN1_RN="s", DTHRN_A="s", DTHRN_U="s", ANG_GC_A="s", ANG_GC_B="s",
                   ANG_GC_C="s", SQRTGOSURFT="s", ZPI4GM1="s", ZPI4GM2="s", ROWATERM1="s", DELTH="s")


def run_tables(name, **kw):
    """The one-off table builders of the reference (INIWCST, MFREDIR, SETWAVPHYS, INIT_X0TAUHF, INIT_SDISS_ARDH, INISNONLIN + NLWEIGT +
    JAFU, INITGC, CIGETDEAC) from their own source, for one spectral / physics setting; compared with the oracle's tables and stored."""
    from f90run import module_registry
    from oracle import oracle as O
    from ecwam_b200 import synth
    o = O.Oracle(O.default_config(**kw), synth.make_grid(8, "aqua"))
    c = o.cfg
    reg = module_registry(TABLE_MODULES)
    ns = module_parameters()
    for k, (t, a) in reg.items():
        if k not in ns:
            ns[k] = None
    I = lambda v: FInt(int(v))
    ns.update(NANG=I(c.nang), NFRE=I(c.nfre), NFRE_RED=I(c.nfre_red), NFRE_ODD=I(c.nfre - 1 + (c.nfre % 2)), IFRE1=I(c.ifre1), FR1=float(c.fr1),
              IPHYS=I(c.iphys), ISNONLIN=I(c.isnonlin), LLGCBZ0=bool(c.llgcbz0), LLNORMAGAM=bool(c.llnormagam), LLCAPCHNK=bool(c.llcapchnk),
              IU06=I(6), LHOOK=False, LWCOU=False, XKAPPA=float(ns.get("XKAPPA", 0.4) or 0.4), RNU=float(c.rnu), RNUM=float(c.rnum),
              ISHALLO=I(0), IRANK=I(1))
    from f90run import STATIC_DIMS
    for k, (t, dims) in STATIC_DIMS.items():
        try:
            ns[k] = FArr([(1, int(eval(d, dict(ns)))) for d in dims], t)
        except Exception:
            pass
    T = Translator([x + ".F90" for x in TABLE_FILES], registry=reg)
    # the frequency arrays INITMDL derives inline (initmdl.F90:436-503), taken as a routine of their own
    from f90run import Routine, _logical_lines
    ll = _logical_lines(os.path.join(REF, "initmdl.F90"), [REF])
    i0 = next(i for i, x in enumerate(ll) if x.replace(" ", "") == "IF(ALLOCATED(DFIMOFR))DEALLOCATE(DFIMOFR)")
    i1 = next(i for i, x in enumerate(ll) if x.replace(" ", "") == "CALLTABU_SWELLFT")
    fr = Routine("INITMDL_FREQ", "SUBROUTINE", [])
    fr.result = None
    fr.body = ll[i0:i1]
    text = " ".join(fr.body)
    fr.uses = {"MODULES": [k for k in reg if re.search(r"\b%s\b" % k, text)]}
    T.routines["INITMDL_FREQ"] = fr
    T.no_intent_outs["KZEONE"] = ("RE0", "IM0", "RE1", "IM1")       # F77-style dummies without INTENT
    T.no_intent_outs["KERKEI"] = ("KER", "KEI")
    ns = T.compile(["INIWCST", "MFREDIR", "SETWAVPHYS", "INITMDL_FREQ", "INIT_X0TAUHF", "INIT_SDISS_ARDH", "INISNONLIN", "INITGC", "CIGETDEAC",
                    "DEPTHPRPT", "TABU_SWELLFT"], ns)
    g = ns          # the functions' globals ARE this dict: module variables they assign land here
    g["INIWCST"](1.0)
    g["MFREDIR"]()
    g["DELTH"] = g["ZPI"] / float(c.nang)          # initmdl.F90:437 (the rest of INITMDL's frequency arrays is not translated)
    g["SETWAVPHYS"]()
    g["INITMDL_FREQ"]()
    g["INIT_X0TAUHF"]()
    if c.iphys == 1:
        g["INIT_SDISS_ARDH"]()
    g["INISNONLIN"]()
    g["INITGC"]()
    g["CIGETDEAC"]()
    if name == "a12_ard":          # TABU_SWELLFT + KERKEI + KZEONE (independent of the spectral setting; 20 000 Kelvin-function evaluations): once
        g["TABU_SWELLFT"]()
    else:
        g["SWELLFT"] = None
    out, worst = dict(kw=__import__("json").dumps(kw, sort_keys=True)), []
    # DEPTHPRPT + AKI (depthprpt.F90, aki.F90): the dispersion relation at the grid's own depths
    grid = synth.make_grid(8, "aqua")
    dep = np.concatenate([grid.depth[::9], [1.5, 3.0, 7.0, 15.0, 40.0, 120.0, 400.0, 998.0]])
    K, NF = dep.size, int(c.nfre)
    dp = {k: FArr([(1, K), (1, NF)]) for k in ("WAVNUM", "CINV", "CGROUP", "XK2CG", "OMOSNH2KD", "STOKFAC")}
    g["DEPTHPRPT"](I(1), I(K), FArr.of(dep), dp["WAVNUM"], dp["CINV"], dp["CGROUP"], dp["XK2CG"], dp["OMOSNH2KD"], dp["STOKFAC"])
    out["DEPTH_IN"] = dep
    for k, v in dp.items():
        out["DP_" + k] = v.a
    for nm, kind in TABLE_CHECK.items():
        v = g.get(nm)
        if v is None:
            continue
        try:
            ref = o.itable(nm) if kind == "i" else o.table(nm)
        except KeyError:
            continue
        a = np.asarray(v.a).ravel(order="F") if isinstance(v, FArr) else np.array([v])
        if kind == "i":
            a = a.astype(np.int64)
        out[nm] = a
        if a.shape != ref.shape:
            worst.append((nm, "shape %s vs %s" % (a.shape, ref.shape)))
            continue
        d = float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-300))
        worst.append((nm, d))
    bad = [(n, d) for n, d in worst if not isinstance(d, float) or d > 0.0]
    print("%s: %d tables from the reference source; not bit-identical to the oracle: %s" % (name, len(worst), bad))
    np.savez_compressed(os.path.join(HERE, "ref_tables_%s.npz" % name), **out)


PROP_FILES = "ctuwupdt ctuwini ctuwdrv ctuw propags2 propdot gradi".split()
PROP_MODULES = MODULES + ["yowubuf", "yowmap", "yowgrid", "yowrefd", "yowmpp"]


def run_propag(name, N=8, mask="continents", obs=False, **kw):
    """CTUWUPDT (+ CTUWINI, CTUWDRV, CTUW) and PROPAGS2 from their own source on a small one-rank grid: the CTU weights of every point,
    direction and frequency and one advection step, compared with the oracle's stored weights / PROPAG_WAM and stored."""
    import ctypes as C
    from f90run import STATIC_DIMS, module_registry
    from oracle import oracle as O
    from ecwam_b200 import model as M, synth
    g = synth.make_grid(N, mask)
    c = O.default_config(store_all_weights=1, nproma=16, **kw)
    o = O.Oracle(c, g)
    f = synth.make_forcing(g)
    for k, v in f.items():
        o.set_field(k, v)
    fl = synth.jonswap_cold_start(f["WSWAVE"], f["WDWAVE"], c.nang, 36, c.nfre_red)
    o.set_fl1(fl)
    n, A, FR_ = g.niblo, c.nang, c.nfre_red
    new2ij = o.itable("NEWIJ2IJ")[1:n + 1] - 1           # original index of the point with new number IJ = 1..NIBLO
    reg = module_registry(PROP_MODULES)
    ns = module_parameters()
    for k in reg:
        ns.setdefault(k, None)
    I = lambda v: FInt(int(v))
    kxlt = o.itable("KXLT")[:n]
    ixlg = o.itable("IXLG")[:n]
    ngy = int(g.ngy)
    ns.update(NANG=I(A), NFRE_RED=I(FR_), NGY=I(ngy), NPROC=I(1), IRANK=I(1), NPROMA_WAM=I(16), IREFRA=I(c.irefra), ICASE=I(1), IRGG=I(1), IPER=I(1),
              IDELPRO=I(int(c.idelpro)), DELPRO_LF=float(c.delpro_lf), IFRELFMAX=I(c.ifrelfmax), LLCFLCUROFF=bool(c.llcflcuroff), LHOOK=False,
              NULERR=I(0), IU06=I(6), ZPI=float(o.table("ZPI")[0]), R=float(o.table("R")[0]), DELTH=float(o.table("DELTH")[0]),
              XDELLA=float(o.table("XDELLA")[0]), AMOWEP=float(g.amowep) if hasattr(g, "amowep") else 0.0, AMOSOP=float(g.amosop),
              FR=FArr.of(o.table("FR")), COSTH=FArr.of(o.table("COSTH")[:A]), SINTH=FArr.of(o.table("SINTH")[:A]),
              ZDELLO=FArr.of(o.table("ZDELLO")[:ngy]), COSPH=FArr.of(o.table("COSPH")[:ngy]), SINPH=FArr.of(o.table("SINPH")[:ngy]),
              BLK2GLO_KXLT=FArr.of(kxlt.astype(np.int64)), BLK2GLO_IXLG=FArr.of(ixlg.astype(np.int64)),
              KLAT=FArr.of(o.itable("KLAT").reshape((n, 2, 2), order="F").astype(np.int64)), KLON=FArr.of(o.itable("KLON").reshape((n, 2), order="F").astype(np.int64)),
              KCOR=FArr.of(o.itable("KCOR").reshape((n, 4, 2), order="F").astype(np.int64)),
              WLAT=FArr.of(o.rank_double("WLAT").reshape((n, 2), order="F")), WCOR=FArr.of(o.rank_double("WCOR").reshape((n, 4), order="F")),
              OBSLAT=FArr.of(np.ones((n, FR_, 2))), OBSLON=FArr.of(np.ones((n, FR_, 2))), OBSCOR=FArr.of(np.ones((n, FR_, 4))))
    if obs:      # LSUBGRID = T: synthetic obstruction coefficients (tests/test_oracle.py _obstructions), in the new numbering
        from test_oracle import _obstructions
        ob = _obstructions(n, FR_)
        o.set_obstructions(*ob)
        for nm, a in zip(("OBSLON", "OBSLAT", "OBSCOR"), ob):
            ns[nm] = FArr.of(np.ascontiguousarray(a[:, :, new2ij].transpose(2, 1, 0)))
    ns["OML_GET_MAX_THREADS"] = lambda: FInt(1)
    # PROENVHALO on one rank: the fields in the new numbering + the land point NSUP + 1 (proenvhalo.F90:98-106)
    s = M.WamSetup(g, nproc=1, nang=A, nfre_red=FR_)
    cg = o.get_field3("CGROUP")[:FR_, new2ij]                        # [m, IJ]
    ext = lambda a, land: FArr.of(np.concatenate([a, [land]]))
    cg_ext = FArr.of(np.concatenate([cg.T, s.land_cgroup[None, :FR_]], axis=0))
    om = o.get_field3("OMOSNH2KD")[:FR_, new2ij]
    om_ext = FArr.of(np.concatenate([om.T, np.zeros((1, FR_))], axis=0))
    cosphm1 = 1.0 / o.table("COSPH")[kxlt - 1]
    depth = g.depth[new2ij]
    ucur, vcur = np.zeros(n), np.zeros(n)
    if c.irefra >= 2:
        from common import synthetic_currents
        u0, v0 = synthetic_currents(g)
        o.set_field("UCUR", u0); o.set_field("VCUR", v0)
        ucur, vcur = u0[new2ij], v0[new2ij]
    u_ext, v_ext, d_ext, c_ext = ext(ucur, 0.0), ext(vcur, 0.0), ext(depth, c.bathymax), ext(cosphm1, 0.0)      # the land slot: proenvhalo.F90:109-116
    ns.update(NIBLO=I(n), DELPHI=float(o.table("XDELLA")[0]) * float(ns["CIRC"]) / 360.0, DELLAM=FArr.of(o.table("DELLAM")[:ngy]))
    T = Translator([x + ".F90" for x in PROP_FILES + ["propag_wam"]], registry=reg, stubs=("PROENVHALO", "PROPAGS", "PROPAGS1", "GSTATS"),
                   externals=("MPEXCHNG",))
    ns["MPEXCHNG"] = lambda *a: None            # one rank: no halo
    ns = T.compile(["CTUWUPDT", "PROPAGS2", "PROPDOT", "PROPAG_WAM"], ns)
    t0 = time.time()
    if c.irefra != 0:     # propag_wam.F90:171-216: PROPDOT (+ GRADI) on the PROENVHALO fields, before CTUWUPDT
        land = np.array([c.bathymax])
        lw, lo_ = np.empty(c.nfre), np.empty(c.nfre)
        dpp = C.POINTER(C.c_double)
        s.lib.ecwam_b200_host_depthprpt(C.byref(s.tables), c.nfre, 1, land.ctypes.data_as(dpp), lw.ctypes.data_as(dpp), None, None, None,
                                        lo_.ctypes.data_as(dpp), None)
        wn = o.get_field3("WAVNUM")[:FR_, new2ij]
        wn_ext = FArr.of(np.concatenate([wn.T, lw[None, :FR_]], axis=0))
        om_ext = FArr.of(np.concatenate([om.T, lo_[None, :FR_]], axis=0))
        THDC, THDD, SDOT = FArr([(1, n), (1, A)]), FArr([(1, n), (1, A)]), FArr([(1, n), (1, A), (1, FR_)])
        ns["PROPDOT"](I(1), I(n), I(1), I(n), None, wn_ext, cg_ext, om_ext, c_ext, d_ext, u_ext, v_ext, THDC, THDD, SDOT)
        ns["THDC"], ns["THDD"], ns["SDOT"] = THDC, THDD, SDOT
    ns["CTUWUPDT"](I(1), I(n), I(1), I(n), None, cg_ext, om_ext, c_ext, d_ext, u_ext, v_ext)
    # PROPAGS2 on FL1_EXT (new numbering + the land slot, 0 there: propag_wam.F90:145-147)
    f1 = np.zeros((n + 1, A, FR_))
    f1[:n] = fl[:FR_][:, :, new2ij].transpose(2, 1, 0)
    F1, F3 = FArr.of(f1), FArr([(1, n + 1), (1, A), (1, FR_)])
    ns["PROPAGS2"](F1, F3, I(1), I(n), I(1), I(n), I(A), I(1), I(FR_), I(1), I(FR_))
    print("%s: CTUWUPDT + PROPAGS2 of the reference source on %d points in %.1f s" % (name, n, time.time() - t0))
    assert o.propag() == 0
    out = dict(N=N, mask=mask, obs=int(obs), kw=__import__("json").dumps(kw, sort_keys=True))
    arrays = [("SUMWN", (n, A, FR_)), ("WLONN", (n, A, FR_, 2)), ("WLATN", (n, A, FR_, 2, 2)), ("WCORN", (n, A, FR_, 4, 2)), ("WKPMN", (n, A, FR_, 3))]
    if c.irefra >= 2:
        arrays.append(("WMPMN", (n, A, FR_, 3)))
    for nm, shape in arrays:
        ref = ns[nm].a
        got = o.rank_double(nm).reshape(shape, order="F")
        out[nm] = ref
        print("   %-6s max |oracle - reference source| = %.2e (max %.3f)" % (nm, np.abs(got - ref).max(), np.abs(ref).max()))
    out["F3"] = F3.a[:n]
    got = o.get_fl1()[:FR_][:, :, new2ij].transpose(2, 1, 0)
    m0 = c.ifrelfmax if 0 < c.ifrelfmax < FR_ else 0       # the frequencies below IFRELFMAX go through further sub-steps in PROPAG_WAM
    if m0:
        # PROPAG_WAM whole (propag_wam.F90) with the weights built above (LUPDTWGHT = LLUPDTTD = F): chunks -> FL1_EXT, PROPAGS2, the
        # NSTEP_LF - 1 further sub-steps of the frequencies 1..IFRELFMAX, FL3_EXT -> chunks and the padding of the last chunk
        P = 16
        C_ = (n + P - 1) // P
        chunk = np.zeros((P, A, 36, C_))
        kijl, ijfrom = np.zeros(C_, dtype=np.int64), np.zeros((P, C_), dtype=np.int64)
        for ic in range(C_):
            k = min(P, n - ic * P)
            kijl[ic] = k
            ijfrom[:k, ic] = ic * P + 1 + np.arange(k)
            chunk[:k, :, :FR_, ic] = f1[ic * P: ic * P + k]
        FL1 = FArr.of(chunk)
        ns.update(NCHNK=I(C_), KIJL4CHNK=FArr.of(kijl), IJFROMCHNK=FArr.of(ijfrom), NINF=I(1), NSUP=I(n), NFRE=I(36), LLUNSTR=False, IPROPAGS=I(2),
                  LUPDTWGHT=False, LLUPDTTD=False, LLCHKCFL=False, LLCHKCFLA=False)
        t0 = time.time()
        ns["PROPAG_WAM"](None, None, None, None, FL1, None, None, None, None, None)
        whole = np.concatenate([FL1.a[:int(kijl[ic]), :, :FR_, ic] for ic in range(C_)], axis=0)
        assert np.array_equal(whole[:, :, m0:], out["F3"][:, :, m0:])       # the frequencies above IFRELFMAX: the single PROPAGS2 call above
        pad = FL1.a[int(kijl[-1]):, :, :FR_, -1]
        assert np.array_equal(pad, np.broadcast_to(FL1.a[0, :, :FR_, -1], pad.shape))      # propag_wam.F90:388-398
        out["F3"] = whole
        m0 = 0
        print("   PROPAG_WAM whole (%d sub-steps) in %.1f s" % (round(c.idelpro / c.delpro_lf), time.time() - t0))
    out["m0"] = m0
    print("   PROPAGS2: max |oracle - reference source| = %.2e, identical: %s" % (np.abs(got - out["F3"])[:, :, m0:].max(),
                                                                               np.array_equal(got[:, :, m0:], out["F3"][:, :, m0:])))
    out["new2ij"] = new2ij
    sel = np.arange(0, n, 6)             # every 6th point is kept in the fixture (size)
    for k in ("SUMWN", "WLONN", "WLATN", "WCORN", "WKPMN", "WMPMN", "F3"):
        if k in out:
            out[k] = out[k][sel]
    out["sel"] = sel
    np.savez_compressed(os.path.join(HERE, "ref_propag_%s.npz" % name), **out)


OUT_FILES = ("outblock femean intpol sepwisw sthq mwp1 mwp2 wdirspread peakfri scosfl outbeta weflux se10mean sebtmean meansqs meansqs_gc "
             "meansqs_lf halphap omegagc ns_gc dominant_period cimsstrn aki_ice outsetwmask chnkmin").split()
OUT_STUBS = ("KURTOSIS", "CAL_SECOND_ORDER_SPEC", "W_MAXH", "CTCOR", "IBRMEMOUT", "SEP3TR")      # not reached / results not selected


def run_outblock(name, case, hook=False, **kw):
    """OUTBLOCK (outblock.F90) and the 24 routines below it from their own source, for the 51 output parameters the product builds, on the
    points pick_points() chooses, after two full steps; compared with the oracle's OUTBS and stored."""
    from common import OUT_ICE, OUT_ITG, OUT_SEA
    from f90run import module_registry
    g, o, f = prepare(case, kw, 2, hook)
    o.implsch()
    if kw.get("irefra", 0) >= 2:
        from common import synthetic_currents
        u0, v0 = synthetic_currents(g)
        o.set_field("UCUR", u0); o.set_field("VCUR", v0)
    pts = pick_points(g, o, f)
    K, A, NF = len(pts), o.cfg.nang, o.cfg.nfre
    ns = namespace(o, {})
    reg = module_registry(MODULES)
    I = lambda v: FInt(int(v))
    JP = int(ns["JPPFLAG"])
    ipf, itob, info = np.zeros(JP, dtype=np.int64), np.zeros(JP, dtype=np.int64), np.zeros((JP, 7), dtype=np.int64)
    for col, (itg, ice, sea) in enumerate(zip(OUT_ITG, OUT_ICE, OUT_SEA)):
        ipf[itg - 1] = 1; itob[itg - 1] = col + 1; info[itg - 1, 5] = ice; info[itg - 1, 6] = sea
    for ih, (tmin, tmax) in enumerate(((10, 12), (12, 14), (14, 17), (17, 21), (21, 25), (25, 30))):      # mpcrtbl.F90:371-399 (parameters 64 - 69)
        info[63 + ih, 3] = tmin; info[63 + ih, 4] = tmax
    ng = int(ns["NWAV_GC"])
    ns.update(IPFGTBL=FArr.of(ipf), ITOBOUT=FArr.of(itob), IPRMINFO=FArr.of(info), NIPRMOUT=I(len(OUT_ITG)), NTRAIN=I(3), NTEWH=I(6), LSECONDORDER=False,
              LLPARTITION=False, LLSOURCE=True, IREFRA=I(o.cfg.irefra), ZMISS=-999.0, DEG=360.0 / float(ns["ZPI"]), CLDOMAIN="g", XKMSS_CUTOFF=float(o.table("XK_GC")[ng - 1]),
              DFIMFR_SIM=FArr.of(o.table("DFIM_SIM")[:NF] * o.table("FR")[:NF]), DFIMFR2_SIM=FArr.of(o.table("DFIM_SIM")[:NF] * o.table("FR")[:NF] ** 2),
              VG_GC=FArr.of(o.table("VG_GC")[:ng]))
    T = Translator([x + ".F90" for x in OUT_FILES], registry=reg, stubs=OUT_STUBS)
    ns = T.compile(["OUTBLOCK"], ns)
    arg = {"FL1": FArr.of(np.ascontiguousarray(o.get_fl1()[:, :, pts].transpose(2, 1, 0))),
           "XLLWS": FArr.of(np.ascontiguousarray(o.get_xllws()[:, :, pts].transpose(2, 1, 0)))}
    for nm in ("WAVNUM", "CINV", "CGROUP"):
        arg[nm] = FArr.of(np.ascontiguousarray(o.get_field3(nm)[:, pts].T))
    for nm in ("DEPTH", "UCUR", "VCUR", "IBRMEM", "USTOKES", "VSTOKES", "STRNMS", "TAUXD", "TAUYD", "TAUOCXD", "TAUOCYD", "TAUOC", "TAUICX", "TAUICY", "PHIOCD",
               "PHIEPS", "PHIAW", "AIRD", "WDWAVE", "CICOVER", "WSWAVE", "WSTAR", "UFRIC", "TAUW", "Z0M", "Z0B", "CHRNCK", "CITHICK"):
        arg[nm] = FArr.of(o.get_field(nm)[pts])
    for nm in ("ALTWH", "CALTWH", "RALTCOR", "NEMOCICOVER", "NEMOCITHICK", "NEMOUCUR", "NEMOVCUR"):
        arg[nm] = FArr.of(np.zeros(K))
    arg["IODP"] = FArr.of(np.ones(K, dtype=np.int64))
    arg["MIJ"] = FArr.of(o.get_field("MIJ")[pts].astype(np.int64))
    arg["BOUT"] = FArr([(1, K), (1, len(OUT_ITG))])
    t0 = time.time()
    # SEBTMEAN reads FR(0) and FL1(:,:,0) for a period band that lies below the first model frequency (MCUTT = MCUTB - 1 = 0,
    # sebtmean.F90:95,122-131): out of bounds in the reference, with weight WL = 0 -- any finite value gives EBT = EPSMIN.  The
    # translator returns 0 for such reads in this run only.
    FArr.oob_read_zero = True
    try:
        ns["OUTBLOCK"](*[FInt(1) if a == "KIJS" else FInt(K) if a == "KIJL" else arg[a] for a in T.routines["OUTBLOCK"].args])
    finally:
        FArr.oob_read_zero = False
    ref = arg["BOUT"].a.T.copy()                       # [column, point]
    got = o.outbs(OUT_ITG, OUT_ICE, OUT_SEA)[:, pts]
    print("%s: OUTBLOCK of the reference source, %d parameters on %d points in %.1f s" % (name, len(OUT_ITG), K, time.time() - t0))
    assert np.array_equal(ref == -999.0, got == -999.0), "missing-value pattern"
    ok = ref != -999.0
    bad = []
    for i, itg in enumerate(OUT_ITG):
        m = ok[i]
        if not m.any():
            continue
        d = np.abs(got[i][m] - ref[i][m]).max() / max(np.abs(ref[i][m]).max(), 1e-300)
        if d > 1e-12:
            bad.append((itg, float(d)))
    print("   parameters further than 1e-12 from the oracle:", bad)
    import json
    np.savez_compressed(os.path.join(HERE, "ref_outblock_%s.npz" % name), case=case, hook=int(hook), kw=json.dumps(kw, sort_keys=True), pts=pts,
                        itg=np.array(OUT_ITG), BOUT=ref)


OUT_CASES = {"ard": dict(case="o48like"), "jan_noicemask": dict(case="o48_iphys0", kw=dict(lmaskice=0)), "a36_shelf": dict(case="o640like", hook=True),
             "currents": dict(case="o48like", kw=dict(irefra=3))}


def run_getwnd(name, llwswave=0, llwdwave=0, lrelwind=0, irefra=0, iparamci=31, liceth=0, lmaskice=1):
    """GETWND's blocking step -- WAMWND (wamwnd.F90, ICODE_WND = 3) + MICEP (micep.F90, uncoupled) -- from their own source on the forcing
    fields of tests/common.py synthetic_fieldg; compared with the oracle's getwnd_points and stored."""
    from common import synthetic_fieldg
    from f90run import module_registry
    from oracle import oracle as O
    from ecwam_b200 import synth
    g = synth.make_grid(8, "continents")
    f, ii, jj = synthetic_fieldg(g, with_ws=bool(llwswave or llwdwave))
    if iparamci == 139:           # the ice field is an SST [K]: ice where it is below 271.5 K
        rng = np.random.default_rng(11)
        f["cicover"] = 268.0 + 10.0 * rng.random(f["uwnd"].shape)
    n = len(ii)
    rng = np.random.default_rng(3)
    uc, vc = rng.normal(0, 0.5, n), rng.normal(0, 0.5, n)
    ny, nx = f["uwnd"].shape
    reg = module_registry(MODULES + ["yowmap", "yowmpp", "yownemoflds"])
    ns = module_parameters()
    for k in reg:
        ns.setdefault(k, None)
    I = lambda v: FInt(int(v))
    wspmin = 0.3
    ns.update(LWCOU=False, IREFRA=I(irefra), LRELWIND=bool(lrelwind), WSPMIN=wspmin, LLWSWAVE=bool(llwswave), LLWDWAVE=bool(llwdwave),
              ZMISS=-999.0, EPSUS=1.0e-6, G=9.806, ZPI=2.0 * float(4.0 * np.arctan(1.0)), XKAPPA=0.4, IU06=I(6), LHOOK=False,
              CITHRSH=0.3, LICERUN=True, LMASKICE=bool(lmaskice), LICETH=bool(liceth), NGX=I(nx), NGY=I(ny), CLDOMAIN="g", IRANK=I(1), NPROC=I(1),
              SWAMPCITH=0.0, LWNEMOCOUCIC=False, LWNEMOCOUCIT=False, LNEMOICEREST=False, HICMIN=float(ns.get("HICMIN", 0.2) or 0.2))
    zero = np.zeros((ny, nx))
    for k in ("uwnd", "vwnd", "aird", "wstar", "cicover", "cithick", "ustra", "vstra", "wswave", "wdwave"):
        ns["FIELDG_" + k.upper()] = FArr.of(np.ascontiguousarray(f.get(k, zero).T))         # FIELDG%X(IX, JY)
    ns["FIELDG_LKFR"] = FArr.of(np.zeros((nx, ny)))
    T = Translator(["wamwnd.F90", "micep.F90"], registry=reg)
    ns = T.compile(["WAMWND", "MICEP"], ns)
    IF_, JF_ = FArr.of(ii.astype(np.int64)), FArr.of(jj.astype(np.int64))
    out = {k: FArr([(1, n)]) for k in ("U10", "US", "THW", "ADS", "WSTAR", "CITH", "USTRA", "VSTRA", "CICVR")}
    lwcur = bool(lrelwind and irefra >= 2)
    ns["WAMWND"](I(1), I(n), IF_, JF_, I(1), I(nx), I(1), I(ny), None, FArr.of(uc), FArr.of(vc), out["U10"], out["US"], out["THW"], out["ADS"],
                 out["WSTAR"], out["CITH"], out["USTRA"], out["VSTRA"], lwcur, I(3))
    ns["MICEP"](I(iparamci), I(1), I(n), IF_, JF_, I(1), I(nx), I(1), I(ny), None, out["CICVR"], out["CITH"], FArr.of(np.zeros(n)), FArr.of(np.zeros(n)))
    ref = dict(wswave=out["U10"].a, wdwave=out["THW"].a, aird=out["ADS"].a, wstar=out["WSTAR"].a, cicover=out["CICVR"].a, cithick=out["CITH"].a,
               ustra=out["USTRA"].a, vstra=out["VSTRA"].a)
    got = O.getwnd_points(ii, jj, f, ucur=uc, vcur=vc, llwswave=llwswave, llwdwave=llwdwave, lcorrel=int(lwcur), iparamci=iparamci, liceth=liceth,
                          lmaskice=lmaskice, wspmin=wspmin)
    bad = [(k, float(np.abs(got[k] - ref[k]).max())) for k in ref if not np.array_equal(got[k], ref[k])]
    print("%s: WAMWND + MICEP of the reference source on %d points; fields not identical to the oracle: %s" % (name, n, bad))
    np.savez_compressed(os.path.join(HERE, "ref_getwnd_%s.npz" % name), opts=np.array([llwswave, llwdwave, lrelwind, irefra, iparamci, liceth, lmaskice]),
                        uc=uc, vc=vc, **ref)


GETWND_CASES = {"plain": dict(), "wswave": dict(llwswave=1), "wswave_wdwave_relwind": dict(llwswave=1, llwdwave=1, lrelwind=1, irefra=3),
                "sst_liceth": dict(iparamci=139, liceth=1), "nomask_liceth": dict(liceth=1, lmaskice=0)}


def run_newwind(name, icode=3):
    """NEWWIND (newwind.F90:105-167) from its own source: FF_NOW <- FF_NEXT with the low-wind cap of the first-guess wave stress
    (ICODE_WND = 3) or the friction-velocity branch (ICODE_WND = 1); compared with the oracle and stored."""
    from common import next_forcing
    from f90run import module_registry
    kw = dict(icode=icode) if icode != 3 else {}
    g, o, f = prepare("o48like", kw, 2, False)
    n = g.niblo
    P = 16
    C_ = (n + P - 1) // P
    nxt = next_forcing(f)
    us = np.sqrt(8.0e-4 + 8.0e-5 * f["WSWAVE"]) * f["WSWAVE"]
    nxt["UFRIC"] = np.maximum(0.05, 1.3 * us * ((np.arange(n) * 13) % 7) / 6.0)
    nxt["WSWAVE"] = np.where(np.arange(n) % 3 == 0, 0.25 * nxt["WSWAVE"], nxt["WSWAVE"])      # some below WSPMIN_RESET_TAUW = 4 m/s
    reg = module_registry(MODULES + ["yowgrid", "yowwndg"])
    ns = module_parameters()
    for k in reg:
        ns.setdefault(k, None)
    I = lambda v: FInt(int(v))
    ns.update(ACD=float(o.table("ACD")[0]), BCD=float(o.table("BCD")[0]), EPSMIN=float(o.table("EPSMIN")[0]), LWCOU=False, NPROMA_WAM=I(P), NCHNK=I(C_),
              NFRE=I(36), ALPHA=float(o.table("ALPHA")[0]), IDELWO=I(3600), IU06=I(6), CDATEWL="20200101000000", CDAWIFL="20200101000000",
              CDATEFL="20300101000000", CDTNEXT="20200101010000", NSTORE=I(1), ICODE=I(icode), ICODE_CPL=I(icode), LHOOK=False)

    def chunked(v):
        a = np.zeros(P * C_)
        a[:n] = v
        return FArr.of(a.reshape((P, C_), order="F"))
    now = ("WSWAVE", "WDWAVE", "AIRD", "WSTAR", "CICOVER", "CITHICK", "USTRA", "VSTRA", "UFRIC", "TAUW", "CHRNCK")
    for k in now:
        ns["FF_NOW_" + k] = chunked(o.get_field(k))
    for k in ("WSWAVE", "WDWAVE", "AIRD", "WSTAR", "CICOVER", "CITHICK", "USTRA", "VSTRA", "UFRIC"):
        ns["FF_NEXT_" + k] = chunked(nxt[k])
    # the padded lanes of the last chunk divide by CHRNCK in the friction-velocity branch: give them a finite value
    ns["FF_NOW_CHRNCK"].a[ns["FF_NOW_CHRNCK"].a == 0.0] = 0.018
    T = Translator(["newwind.F90"], registry=reg, stubs=("INCDATE",))
    ns = T.compile(["NEWWIND"], ns)
    ns["NEWWIND"]("20200101010000", "20200101010000", False, None, None, None)
    o.newwind(nxt)
    out, bad = dict(icode=icode), []
    for k in now:
        ref = ns["FF_NOW_" + k].a.reshape(-1, order="F")[:n]
        out[k] = ref
        if not np.array_equal(o.get_field(k), ref):
            bad.append((k, float(np.abs(o.get_field(k) - ref).max())))
    for k in nxt:
        out["NEXT_" + k] = nxt[k]
    print("%s: NEWWIND of the reference source on %d points; fields not identical to the oracle: %s" % (name, n, bad))
    np.savez_compressed(os.path.join(HERE, "ref_newwind_%s.npz" % name), **out)


def run_decomp(name, N, mask, npr, ll1d=0):
    """The sector decomposition of MPDECOMP (mpdecomp.F90: the block between `NXFFS=1` and `DEALLOCATE(NEND1D)`, i.e. NXDECOMP x NYDECOMP
    sectors, NSTART / NEND and the relabelling NEWIJ2IJ / IJ2NEWIJ of the sea points) executed from its own source for NPR ranks; compared
    with the oracle's tables and stored.  (The halo lists that follow use MPL_ALLGATHERV and are not translated.)"""
    from f90run import _logical_lines, module_registry
    from oracle import oracle as O
    from ecwam_b200 import synth
    g = synth.make_grid(N, mask)
    o = O.Oracle(O.default_config(nproma=16, npr=npr, ll1d=ll1d), g)
    n, ngy = g.niblo, int(g.ngy)
    reg = module_registry(PROP_MODULES + ["yowspec", "yowunpool"])
    ns = module_parameters()
    for k in reg:
        ns.setdefault(k, None)
    I = lambda v: FInt(int(v))
    ij2new = o.itable("IJ2NEWIJ")[: n + 1]
    ixlg0 = o.itable("IXLG")[:n][ij2new[1:] - 1]            # BLK2GLO in the ORIGINAL numbering: the state MPDECOMP starts from
    kxlt0 = o.itable("KXLT")[:n][ij2new[1:] - 1]
    ns.update(NGX=I(int(np.max(g.nlonrgg))), NGY=I(ngy), NIBLO=I(n), IJS=I(1), IJL=I(n), IRANK=I(1), NPROC=I(npr), LL1D=bool(ll1d), LLUNSTR=False,
              IU06=I(6), LHOOK=False, NLONRGG=FArr.of(np.asarray(g.nlonrgg, dtype=np.int64)),
              IPER=I(1), IRGG=I(1), AMOWEP=0.0, XDELLO=360.0 / float(np.max(g.nlonrgg)), AMOEAP=360.0 - 360.0 / float(np.max(g.nlonrgg)),
              AMOSOP=float(g.amosop), AMONOP=float(g.amonop), XDELLA=float(o.table("XDELLA")[0]), ZDELLO=FArr.of(o.table("ZDELLO")[:ngy]),
              BLK2GLO_IXLG=FArr.of(ixlg0.astype(np.int64)), BLK2GLO_KXLT=FArr.of(kxlt0.astype(np.int64)),
              NEWIJ2IJ=FArr([(0, n)], int), IJ2NEWIJ=FArr([(0, n)], int))
    T = Translator(["mpdecomp.F90"], registry=reg, stubs=("FLUSH",))
    r = T.routines["MPDECOMP"]
    i0 = next(i for i, x in enumerate(r.body) if x.replace(" ", "") == "NXFFS=1")
    i1 = next(i for i, x in enumerate(r.body) if x.replace(" ", "") == "DEALLOCATE(NEND1D)")
    # the ELSE branch of `IF (LLUNSTR)` the fragment sits in is not closed inside it
    r.body = r.body[i0:i1 + 1]
    ns = T.compile(["MPDECOMP"], ns)
    ns["MPDECOMP"](I(npr), I(0), False, False)
    out = dict(N=N, mask=mask, npr=npr, ll1d=ll1d)
    bad = []
    # (NEWIJ2IJ is a local of MPDECOMP; it is the inverse of the module array IJ2NEWIJ.)  BLK2GLO comes back relabelled.
    for nm, sl in (("NSTART", slice(0, npr)), ("NEND", slice(0, npr)), ("IJ2NEWIJ", slice(1, n + 1)), ("IXLG", slice(0, n)), ("KXLT", slice(0, n))):
        ref = ns[nm].a[sl] if nm == "IJ2NEWIJ" else ns["BLK2GLO_" + nm].a if nm in ("IXLG", "KXLT") else ns[nm].a
        got = o.itable(nm)[sl]
        out[nm] = np.asarray(ref, dtype=np.int64)
        if not np.array_equal(got, ref):
            bad.append(nm)
    print("%s: MPDECOMP's sector decomposition of the reference source, %d points on %d ranks; tables not identical to the oracle: %s" % (name, n, npr, bad))
    np.savez_compressed(os.path.join(HERE, "ref_decomp_%s.npz" % name), **out)


class _GatherDone(Exception):
    pass


def run_halo(name, N, mask, npr):
    """MPDECOMP from `NXFFS=1` to the end of its structured-grid branch -- the sector decomposition, PROPCONNECT for the rank, the halo points
    it needs, the two MPL_ALLGATHERVs (emulated: every rank is run once to record what it sends, then again with the gathered arrays), the
    send / receive lists NTOPE, NFROMPE, NIJSTART, IJTOPE and the local addressing of KLAT / KLON / KCOR -- from its own source, for every rank
    of an NPR-rank run; compared with the oracle's per-rank tables and stored."""
    from f90run import module_registry
    from oracle import oracle as O
    from ecwam_b200 import synth
    g = synth.make_grid(N, mask)
    o = O.Oracle(O.default_config(nproma=16, npr=npr), g)
    n, ngy = g.niblo, int(g.ngy)
    reg = module_registry(PROP_MODULES + ["yowspec", "yowunpool", "yowunit"])
    I = lambda v: FInt(int(v))
    ij2new = o.itable("IJ2NEWIJ")[: n + 1]
    ixlg0 = o.itable("IXLG")[:n][ij2new[1:] - 1].astype(np.int64)
    kxlt0 = o.itable("KXLT")[:n][ij2new[1:] - 1].astype(np.int64)
    ngx = int(np.max(g.nlonrgg))
    ocean = np.zeros((ngx, ngy), dtype=bool)
    mk, p = np.asarray(g.mask), 0
    for k in range(ngy):
        ocean[: g.nlonrgg[k], k] = mk[p: p + g.nlonrgg[k]] != 0
        p += g.nlonrgg[k]
    T = Translator(["mpdecomp.F90", "propconnect.F90", "wam_sorti.F90", "wam_sortini.F90"], registry=reg, stubs=("FLUSH", "GSTATS"),
                   externals=("MPL_ALLGATHERV",))
    r = T.routines["MPDECOMP"]
    i0 = next(i for i, x in enumerate(r.body) if x.replace(" ", "") == "NXFFS=1")
    i1 = next(i for i, x in enumerate(r.body) if x.replace(" ", "") == "KTAG=KTAG+1")
    r.body = r.body[i0:i1]
    sent = {}

    def run_rank(rank, gather):
        ns = module_parameters()
        for k in reg:
            ns.setdefault(k, None)
        ns.update(NGX=I(ngx), NGY=I(ngy), NIBLO=I(n), IJS=I(1), IJL=I(n), IRANK=I(rank), NPROC=I(npr), LL1D=False, LLUNSTR=False, IPROPAGS=I(2),
                  IU06=I(6), LHOOK=False, NPROMA_WAM=I(16), KTAG=I(1), NLONRGG=FArr.of(np.asarray(g.nlonrgg, dtype=np.int64)), IPER=I(1), IRGG=I(1), AMOWEP=0.0,
                  XDELLO=360.0 / ngx, AMOEAP=360.0 - 360.0 / ngx, AMOSOP=float(g.amosop), AMONOP=float(g.amonop), XDELLA=float(o.table("XDELLA")[0]),
                  ZDELLO=FArr.of(o.table("ZDELLO")[:ngy]), LLOCEANMASK=FArr.of(ocean), BLK2GLO_IXLG=FArr.of(ixlg0), BLK2GLO_KXLT=FArr.of(kxlt0),
                  NEWIJ2IJ=FArr([(0, n)], int), IJ2NEWIJ=FArr([(0, n)], int), MPL_ALLGATHERV=gather)
        ns = T.compile(["MPDECOMP"], ns)
        ns["IRANK"] = I(rank)
        try:
            ns["MPDECOMP"](I(npr), I(0), False, False)
        except _GatherDone:
            pass
        return ns

    for rank in range(1, npr + 1):         # pass 1: what every rank contributes to the two gathers
        calls = []

        def record(send, recv, counts, CDSTRING=None, rank=rank, calls=calls):
            calls.append(np.array(send.a if isinstance(send, FArr) else send, copy=True))
            if len(calls) == 2:
                sent[rank] = calls
                raise _GatherDone()
        run_rank(rank, record)
    out, bad = dict(N=N, mask=mask, npr=npr), []
    for rank in range(1, npr + 1):         # pass 2: with the gathered arrays
        k = [0]

        def gather(send, recv, counts, CDSTRING=None, k=k):
            data = np.concatenate([sent[q][k[0]].ravel() for q in range(1, npr + 1)])
            recv.a.ravel()[: data.size] = data
            k[0] += 1
        ns = run_rank(rank, gather)
        for nm in ("NINF", "NSUP", "NTOPEMAX", "NFROMPEMAX"):
            ref = int(ns[nm])
            out["%s_%d" % (nm, rank)] = ref
            if int(o.itable(nm, rank - 1)[0]) != ref:
                bad.append((rank, nm))
        for nm in ("KLENBOT", "KLENTOP", "NTOPE", "NFROMPE", "NIJSTART", "IJTOPE", "KLAT", "KLON", "KCOR"):
            ref = np.asarray(ns[nm].a).ravel(order="F").astype(np.int64)
            got = o.itable(nm, rank - 1)
            out["%s_%d" % (nm, rank)] = ref
            if got.shape != ref.shape or not np.array_equal(got, ref):
                bad.append((rank, nm))
    print("%s: MPDECOMP's halo tables of the reference source, %d points, %d ranks; not identical to the oracle: %s" % (name, n, npr, bad))
    np.savez_compressed(os.path.join(HERE, "ref_halo_%s.npz" % name), **out)


WNORM_ITG = [1, 2, 3, 5, 11, 12, 16, 32]       # Hs, mean direction / frequency, U10, wind-sea Hs + direction, ... (some ice-masked)


def wnorm_state(npr, ll1d=0):
    """The oracle after one step of a cold start with some points under the ice mask, and its OUTBS columns WNORM_ITG [column, ij]."""
    from oracle import oracle as O
    from ecwam_b200 import synth
    g = synth.make_grid(12, "continents")
    o = O.Oracle(O.default_config(nang=12, nfre_red=25, nproma=16, npr=npr, ll1d=ll1d), g)
    f = synth.make_forcing(g)
    f["CICOVER"] = np.where(np.arange(g.niblo) % 7 == 0, 0.6, f["CICOVER"])       # some points under the ice mask -> ZMISS
    for k, v in f.items():
        o.set_field(k, v)
    o.set_fl1(synth.jonswap_cold_start(f["WSWAVE"], f["WDWAVE"], 12, 36, 25))
    assert o.step() == 0
    return g, o, o.outbs(WNORM_ITG, [1, 1, 1, 0, 1, 1, 1, 0], [0] * len(WNORM_ITG), -999.0)


def run_wnorm(name, npr, ll1d=0):
    """MPMINMAXAVG (mpminmaxavg.F90) from its own source on every rank of an NPR-rank run, both flavours: LLGLOBAL = T (MPGATHERSCFLD
    emulated: the receiving rank sees every rank's block of ZGLOBAL) and LLGLOBAL = F (MPL_ALLREDUCE emulated: SUM in rank order --
    MPL's reproducible path --, MIN, MAX).  The columns are the oracle's OUTBS columns of a spun-up state (they contain ZMISS under the
    ice mask); WNORM is compared with the oracle's OUTWNORM and stored together with the columns."""
    from f90run import module_registry
    g, o, b = wnorm_state(npr, ll1d)
    n, NI = g.niblo, len(WNORM_ITG)
    zmiss = -999.0
    assert (b == zmiss).any() and (b != zmiss).any()
    reg = module_registry(PROP_MODULES + ["yowspec", "yowcout", "yowgrid", "yowtest"])
    I = lambda v: FInt(int(v))
    new2ij, ij2new = o.itable("NEWIJ2IJ")[: n + 1], o.itable("IJ2NEWIJ")[: n + 1]
    nstart, nend = o.itable("NSTART")[:npr], o.itable("NEND")[:npr]
    T = Translator(["mpminmaxavg.F90"], registry=reg, externals=("MPGATHERSCFLD", "MPL_ALLREDUCE"))
    bouts, chunks = [], []
    for r in range(npr):
        nch, nproma = int(o.itable("NCHNK", r)[0]), int(o.itable("NPROMA", r)[0])
        kijl = o.itable("KIJL4CHNK", r)[:nch]
        ijfrom = o.itable("IJFROMCHNK", r).reshape(nch, nproma).T        # (NPROMA, NCHNK)
        bo = np.zeros((nproma, NI, nch))
        for ic in range(nch):
            for p in range(int(kijl[ic])):
                bo[p, :, ic] = b[:, new2ij[ijfrom[p, ic]] - 1]
        bouts.append(bo); chunks.append((nch, nproma, kijl, ijfrom))
    wn = {}
    for glob in (True, False):
        partial = {}
        for pas in (1, 2):
            for r in range(npr):
                nch, nproma, kijl, ijfrom = chunks[r]
                ns = module_parameters()
                for k in reg:
                    ns.setdefault(k, None)

                def gather(irecv, nblks, nblke, zglobal, niblo, r=r):       # MPGATHERSCFLD: the receiver gets every rank's NBLKS:NBLKE block
                    if pas == 1:
                        partial.setdefault(r, []).append(np.array(zglobal.a, copy=True))
                    else:
                        k = gather.k
                        for q in range(npr):
                            zglobal.a[nstart[q] - 1: nend[q]] = partial[q][k][nstart[q] - 1: nend[q]]
                        gather.k += 1
                gather.k = 0

                def allreduce(z, op, LDREPROD=None, CDSTRING=None, r=r):
                    if pas == 1:
                        partial.setdefault(r, []).append(np.array(z.a, copy=True))
                        if op == "MAX":
                            raise _GatherDone()
                    else:
                        k = allreduce.k
                        acc = partial[0][k].copy()
                        for q in range(1, npr):
                            acc = {"SUM": np.add, "MIN": np.minimum, "MAX": np.maximum}[op](acc, partial[q][k])
                        z.a[:] = acc
                        allreduce.k += 1
                allreduce.k = 0
                flags = np.zeros(int(ns["JPPFLAG"]), dtype=bool); itob = np.zeros(int(ns["JPPFLAG"]), dtype=np.int64)
                for col, itg in enumerate(WNORM_ITG):
                    flags[itg - 1] = True; itob[itg - 1] = col + 1
                ns.update(NIBLO=I(n), IRANK=I(r + 1), NPROC=I(npr), LL1D=bool(ll1d), LLUNSTR=False, ZMISS=zmiss, IU06=I(6), LHOOK=False,
                          NPROMA_WAM=I(nproma), NCHNK=I(nch), NIPRMOUT=I(NI), NFLAG=FArr.of(flags), ITOBOUT=FArr.of(itob),
                          IJFROMCHNK=FArr.of(np.asarray(ijfrom, dtype=np.int64)), KIJL4CHNK=FArr.of(np.asarray(kijl, dtype=np.int64)),
                          NBLKS=FArr.of(np.asarray(nstart, dtype=np.int64)), NBLKE=FArr.of(np.asarray(nend, dtype=np.int64)),
                          IJ2NEWIJ=FArr.of(np.asarray(ij2new, dtype=np.int64), lb=[0]), MPGATHERSCFLD=gather, MPL_ALLREDUCE=allreduce)
                ns = T.compile(["MPMINMAXAVG"], ns)
                w = FArr([(1, 4), (1, NI)])
                try:
                    ns["MPMINMAXAVG"](glob, I(1), True, FArr.of(bouts[r]), w)
                except _GatherDone:
                    continue
                if r == 0 or not glob:          # (pass 1 completes only for NPROC = 1; pass 2 overwrites it otherwise)
                    wn[(glob, r)] = np.array(w.a, copy=True)
    out, bad = dict(npr=npr, ll1d=ll1d, BOUT=b, ITG=np.array(WNORM_ITG), zmiss=zmiss), []
    for glob in (True, False):
        ref = o.outwnorm(glob)              # [column, 4]
        for (gl, r), w in wn.items():
            if gl == glob and not np.array_equal(w.T, ref):
                bad.append((glob, r, float(np.abs(w.T - ref).max())))
        out["WNORM_GLOBAL" if glob else "WNORM_LOCAL"] = wn[(glob, 0)].T
    print("%s: MPMINMAXAVG of the reference source, %d ranks, %d columns; not identical to the oracle: %s" % (name, npr, NI, bad))
    np.savez_compressed(os.path.join(HERE, "ref_wnorm_%s.npz" % name), **out)


SEQ_CASES = {"pro900_src900_wind900": (900, 900, 900, 8, True), "pro1800_src900_wind1800": (1800, 900, 1800, 6, True),
             "pro900_src900_wind3600": (900, 900, 3600, 10, True), "pro3600_src900_wind1800": (3600, 900, 1800, 4, True),
             "pro3600_src1200_wind7200": (3600, 1200, 7200, 5, True), "pro900_src1800_wind1800": (900, 1800, 1800, 8, True),
             "pro1800_src900_wind900_nosource": (1800, 900, 900, 4, False)}
SEQ_T0 = "20200229230000"          # across the end of February of a leap year


def run_sequence(name, idelpro, idelt, idelwo, nadv, llsource):
    """The time stepping of the hot path from its own source: WAMODEL's initialisation of CDTIMP / CDTIMPNEXT (wamodel.F90:182-185), its
    ADVECTION loop body (:232-233, :285-300) and WAMINTGR + NEWWIND whole (PROPAG_WAM / IMPLSCH are recorded instead of run, INCDATE is
    Python's datetime, NEWWIND works on a 4-point block).  Stored: the order of PROPAG_WAM / new winds / IMPLSCH and the dates after
    every WAMINTGR call, in seconds since the start."""
    import datetime
    from f90run import module_registry, Routine, _logical_lines
    reg = module_registry(MODULES + ["yowgrid", "yowwndg", "yowcoup"])
    I = lambda v: FInt(int(v))
    fmt = "%Y%m%d%H%M%S"
    t0 = datetime.datetime.strptime(SEQ_T0, fmt)
    secs = lambda c: int((datetime.datetime.strptime(c, fmt) - t0).total_seconds())
    ns = module_parameters()
    for k in reg:
        ns.setdefault(k, None)
    log = []
    src = open(os.path.join(REF, "wamintgr.F90")).read().upper()
    for nm, idx in re.findall(r"(\w+%\w+)\s*\(([^)]*)\)", src):           # the components WAMINTGR passes on: one 4-point chunk of each
        ns[nm.replace("%", "_")] = FArr([(1, 4)] * (idx.count(",") + 1), int if nm == "MIJ%PTR" else float)
    for nm in re.findall(r"(\w+%\w+)", src):
        ns.setdefault(nm.replace("%", "_"), FArr([(1, 4), (1, 1)], float))
    for k in ("WSWAVE", "WDWAVE", "AIRD", "WSTAR", "CICOVER", "CITHICK", "USTRA", "VSTRA", "UFRIC", "TAUW", "CHRNCK"):
        ns["FF_NOW_" + k] = FArr.of(np.full((4, 1), 0.5)); ns["FF_NEXT_" + k] = FArr.of(np.full((4, 1), 0.7))

    def incdate(c, shift):
        return ((datetime.datetime.strptime(c, fmt) + datetime.timedelta(seconds=int(shift))).strftime(fmt),)
    ns.update(ACD=8.0e-4, BCD=8.0e-5, EPSMIN=1e-10, LWCOU=False, NPROMA_WAM=I(4), NCHNK=I(1), NFRE=I(36), NANG=I(12), ALPHA=0.0065, IDELWO=I(idelwo),
              IDELWI=I(idelwo), IDELPRO=I(idelpro), IDELT=I(idelt), IU06=I(6), CDATEWL=SEQ_T0, CDAWIFL=SEQ_T0, CDATEFL="20300101000000",
              CDTNEXT=SEQ_T0, NSTORE=I(1), ICODE=I(3), ICODE_CPL=I(3), LHOOK=False, LLSOURCE=bool(llsource), LWNEMOCOU=False, NEMONTAU=I(0),
              TIME_PROPAG=0.0, TIME_PHYS=0.0, CDTPRO=SEQ_T0, CDATEWO=incdate(SEQ_T0, idelwo)[0], INCDATE=incdate,
              PROPAG_WAM=lambda *a: log.append(("propag",)), IMPLSCH=lambda *a: log.append(("implsch",)),
              BLK2GLO=None, WVENVI=None, WVPRPT=None, FF_NOW=None, FF_NEXT=None, INTFLDS=None, WAM2NEMO=None, MIJ=None, VARS_4D=None)
    T = Translator(["wamintgr.F90", "newwind.F90"], registry=reg, stubs=("GSTATS",), externals=("INCDATE", "PROPAG_WAM", "IMPLSCH"))
    T.external_outs["INCDATE"] = (0,)
    w = T.routines["WAMINTGR"]
    w.body = [ln for ln in w.body if "WAM_USER_CLOCK" not in ln]              # the timers
    ll = _logical_lines(os.path.join(REF, "wamodel.F90"), [REF])
    sq = lambda x: x.replace(" ", "")
    i0 = next(i for i, x in enumerate(ll) if sq(x) == "CDTIMPNEXT=CDTPRO")
    i1 = next(i for i, x in enumerate(ll) if sq(x) == "CDTPRA=CDTPRO")
    i2 = next(i for i, x in enumerate(ll) if sq(x) == "CDATE=CDTPRA")
    i3 = next(i for i, x in enumerate(ll) if sq(x) == "ILOOP=ILOOP+1")
    assert sq(ll[i0 + 2]) == "CDTIMP=CDTPRO" and sq(ll[i2 + 1]) == "CDATEWH=CDATEWO" and ll[i3 - 1].startswith("CALL WAMINTGR")
    local = ["CDATE", "CDTPRA", "CDTIMP", "CDTIMPNEXT", "CDATEWH"]            # locals of WAMODEL: kept in the namespace between the fragments
    for nm, body in (("WAMODEL_INIT", ll[i0:i0 + 3]), ("WAMODEL_STEP", ll[i1:i1 + 2] + ll[i2:i3 + 1] + ["ENDDO"])):
        fr = Routine(nm, "SUBROUTINE", [])
        fr.result = None
        fr.body = body
        fr.uses = {"MODULES": [k for k in reg if re.search(r"\b%s\b" % k, " ".join(body))] + local + ["CDTPRO", "CDATEWO", "IDELPRO", "IDELT", "BLK2GLO", "WVENVI", "WVPRPT", "FF_NOW", "FF_NEXT", "INTFLDS", "WAM2NEMO", "MIJ", "VARS_4D"]}
        T.routines[nm] = fr
    for nm in local:
        ns[nm] = SEQ_T0
    ns = T.compile(["WAMODEL_INIT", "WAMODEL_STEP", "WAMINTGR", "NEWWIND"], ns)
    wam, nw = ns["WAMINTGR"], ns["NEWWIND"]

    def newwind(cdate, cdatewh, *a):
        r = nw(cdate, cdatewh, *a)
        if r[0] != cdatewh:
            log.append(("newwind", secs(cdatewh)))
        return r

    def wamintgr(*a):
        r = wam(*a)                     # (CDATE, CDATEWH, CDTIMP, CDTIMPNEXT)
        log.append(("dates",) + tuple(secs(x) for x in r))
        return r
    ns["NEWWIND"], ns["WAMINTGR"] = newwind, wamintgr
    ns["WAMODEL_INIT"]()
    for _ in range(nadv):
        ns["WAMODEL_STEP"]()
        log.append(("step", secs(ns["CDTPRO"])))
    code = dict(propag=1, implsch=2, newwind=3, dates=4, step=5)
    ev = np.array([[code[e[0]]] + list(e[1:]) + [0] * (5 - len(e)) for e in log], dtype=np.int64)
    print("%s: %d advection steps -> %d WAMINTGR calls, %d PROPAG_WAM, %d IMPLSCH, %d new winds" %
          (name, nadv, (ev[:, 0] == 4).sum(), (ev[:, 0] == 1).sum(), (ev[:, 0] == 2).sum(), (ev[:, 0] == 3).sum()))
    np.savez_compressed(os.path.join(HERE, "ref_sequence_%s.npz" % name), EVENTS=ev, cfg=np.array([idelpro, idelt, idelwo, nadv, int(llsource)]))
    return ev


def run_connect(name, N=8, mask="continents"):
    """PROPCONNECT (propconnect.F90, 971 lines: the neighbours of every sea point on the irregular grid and their interpolation weights)
    from its own source on a one-rank grid, compared with the oracle's KLAT / KLON / KCOR / WLAT / WCOR and stored."""
    from f90run import module_registry
    from oracle import oracle as O
    from ecwam_b200 import synth
    g = synth.make_grid(N, mask)
    o = O.Oracle(O.default_config(nproma=16), g)
    n, ngy = g.niblo, int(g.ngy)
    reg = module_registry(PROP_MODULES + ["yowspec"])
    ns = module_parameters()
    for k in reg:
        ns.setdefault(k, None)
    I = lambda v: FInt(int(v))
    ngx = int(np.max(g.nlonrgg))
    ocean = np.zeros((ngx, ngy), dtype=bool)
    mk = np.asarray(g.mask)
    if mk.ndim == 1:      # flat, row by row with NLONRGG(k) entries
        p = 0
        for k in range(ngy):
            ocean[: g.nlonrgg[k], k] = mk[p: p + g.nlonrgg[k]] != 0
            p += g.nlonrgg[k]
    else:
        ocean[: mk.shape[1], :] = (mk != 0).T
    ns.update(NGX=I(ngx), NGY=I(ngy), NIBLO=I(n), IPER=I(1), IRGG=I(1), IPROPAGS=I(2), LHOOK=False, XDELLA=float(o.table("XDELLA")[0]),
              ZDELLO=FArr.of(o.table("ZDELLO")[:ngy]), NLONRGG=FArr.of(np.asarray(g.nlonrgg, dtype=np.int64)), LLOCEANMASK=FArr.of(ocean),
              BLK2GLO_IXLG=FArr.of(o.itable("IXLG")[:n].astype(np.int64)), BLK2GLO_KXLT=FArr.of(o.itable("KXLT")[:n].astype(np.int64)),
              IJ2NEWIJ=FArr.of(o.itable("IJ2NEWIJ")[: n + 1].astype(np.int64), lb=[0]),
              KLAT=FArr([(1, n), (1, 2), (1, 2)], int), KLON=FArr([(1, n), (1, 2)], int), KCOR=FArr([(1, n), (1, 4), (1, 2)], int),
              WLAT=FArr([(1, n), (1, 2)]), WCOR=FArr([(1, n), (1, 4)]), WRLAT=FArr([(1, n), (1, 2)]), WRLON=FArr([(1, n), (1, 2)]),
              KRLAT=FArr([(1, n), (1, 2), (1, 2)], int), KRLON=FArr([(1, n), (1, 2), (1, 2)], int))
    T = Translator(["propconnect.F90"], registry=reg)
    ns = T.compile(["PROPCONNECT"], ns)
    t0 = time.time()
    ns["PROPCONNECT"](I(1), I(n), FArr.of(o.itable("NEWIJ2IJ")[1: n + 1].astype(np.int64)))
    print("%s: PROPCONNECT of the reference source on %d points in %.1f s" % (name, n, time.time() - t0))
    out = dict(N=N, mask=mask)
    land = n + 1
    for nm, shape, kind in (("KLAT", (n, 2, 2), "i"), ("KLON", (n, 2), "i"), ("KCOR", (n, 4, 2), "i"), ("WLAT", (n, 2), "d"), ("WCOR", (n, 4), "d")):
        ref = ns[nm].a.copy()
        got = (o.itable(nm) if kind == "i" else o.rank_double(nm)).reshape(shape, order="F")
        if kind == "i":
            ref = np.where(ref == 0, land, ref)         # MPDECOMP sends "no neighbour" to the land point NSUP + 1 afterwards
        out[nm] = ref
        print("   %-5s identical to the oracle: %s" % (nm, np.array_equal(got, ref)))
    np.savez_compressed(os.path.join(HERE, "ref_connect_%s.npz" % name), **out)


PROP_CASES = {"a12": dict(N=8), "a12_subgrid": dict(N=8, obs=True), "a12_subgrid_irefra3": dict(N=8, obs=True, kw=dict(irefra=3)), "a12_irefra1": dict(N=8, kw=dict(irefra=1)), "a12_irefra3": dict(N=8, kw=dict(irefra=3)),
              "a12_irefra2": dict(N=8, kw=dict(irefra=2)), "a24_fastwaves": dict(N=8, kw=dict(nang=24, nfre_red=29, ifrelfmax=5, delpro_lf=225.0, idelpro=450.0, idelt=450.0))}

TABLE_CASES = {"a12_ard": dict(nang=12, nfre_red=25, iphys=1), "a24_ard": dict(nang=24, nfre_red=29, iphys=1), "a36_ard": dict(nang=36, nfre_red=29, iphys=1),
               "a12_jan": dict(nang=12, nfre_red=25, iphys=0), "a12_cy49r1": dict(nang=12, nfre_red=25, iphys=1, llgcbz0=1, llnormagam=1, wspmin=0.3),
               "a36_jan_gc": dict(nang=36, nfre_red=29, iphys=0, llgcbz0=1, llnormagam=1, wspmin=0.3)}

ICE = dict(lmaskice=0, lciwa1=1, lciwa2=1, lciwa3=1, lciscal=1, zalpfacx=0.6, zalpfacb=0.8)
CASES = {
    "ard": dict(case="o48like"),                                               # etopo1_oper_an_fc_O48.yml physics
    "jan": dict(case="o48_iphys0"),                                            # ..._O48_iphys_0.yml
    "a24": dict(case="o320like"),                                              # O320 spectral setting
    "a36_shelf": dict(case="o640like", hook=True),                             # O640 spectral setting, depth-limited points
    "cy49r1": dict(case="o48_cy49r1"),                                         # LLGCBZ0 + LLNORMAGAM
    "cy49r1_jan": dict(case="o48_iphys0_gc"),
    "ice": dict(case="o48like", kw=ICE),                                       # SDICE1 + 2 + 3 + LCISCAL under the ice
    "ice2_jan": dict(case="o48_iphys0", kw=dict(lmaskice=0, lciwa2=1)),
    "nemo": dict(case="o48like", kw=dict(lwnemocou=1, lwnemotauoc=1, lwnemocoustk=1, lwnemocoustrn=1, lwnemocouwrs=1, lwnemocouibr=1,
                                         lmaskice=0, lciwa3=1, zalpfacx=0.6, zalpwrs=0.8)),
    "snl1_shelf": dict(case="o48like", hook=True, kw=dict(isnonlin=1)),
    "snl2_shelf": dict(case="o48_iphys0", hook=True, kw=dict(isnonlin=2)),
    "noflxsnl": dict(case="o48like", kw=dict(lwvflx_snl=0)),
    "ustar": dict(case="o48like", kw=dict(icode=1)),
    "lwflux": dict(case="o48like", kw=dict(lwflux=1, lwcouast=0)),
}

if __name__ == "__main__":
    names = sys.argv[1:] or (list(CASES) + ["tables", "propag", "connect", "outblock", "getwnd", "newwind", "decomp", "halo", "wnorm", "sequence"])
    for nm in names:
        if nm == "tables":
            for t, kw in TABLE_CASES.items():
                run_tables(t, **kw)
            continue
        if nm == "outblock":
            for t, d in OUT_CASES.items():
                run_outblock(t, d["case"], hook=d.get("hook", False), **d.get("kw", {}))
            continue
        if nm == "sequence":
            for k, v in SEQ_CASES.items():
                run_sequence(k, *v)
            continue
        if nm == "wnorm":
            for npr in (1, 2, 3, 5):
                run_wnorm("npr%d" % npr, npr)
            run_wnorm("npr4_1d", 4, ll1d=1)
            continue
        if nm == "halo":
            for npr in (2, 3, 4, 8):
                run_halo("continents12_npr%d" % npr, 12, "continents", npr)
            run_halo("continents16_npr6", 16, "continents", 6)
            run_halo("aqua8_npr5", 8, "aqua", 5)
            continue
        if nm == "decomp":
            for npr in (1, 2, 3, 4, 5, 8):
                run_decomp("continents12_npr%d" % npr, 12, "continents", npr)
            run_decomp("aqua8_npr6", 8, "aqua", 6)
            run_decomp("continents12_npr4_1d", 12, "continents", 4, ll1d=1)
            continue
        if nm == "newwind":
            run_newwind("u10", 3)
            run_newwind("ustar", 1)
            continue
        if nm == "getwnd":
            for t, d in GETWND_CASES.items():
                run_getwnd(t, **d)
            continue
        if nm == "connect":
            run_connect("continents8", 8, "continents")
            run_connect("aqua6", 6, "aqua")
            continue
        if nm == "propag" or nm.startswith("propag:"):
            for t, d in PROP_CASES.items():
                if nm == "propag" or t in nm.split(":")[1:]:
                    run_propag(t, N=d.get("N", 8), obs=d.get("obs", False), **d.get("kw", {}))
            continue
        c = CASES[nm]
        run_case(nm, c["case"], hook=c.get("hook", False), **c.get("kw", {}))
