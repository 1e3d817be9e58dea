// ORACLE (test infrastructure) — dispersion fields, PROENVHALO, CTUWUPDT/CTUWINI/CTUW, PROPAGS2, PROPAG_WAM,
// and an in-process emulation of MPEXCHNG between the emulated ranks.
#include "oracle.h"
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// aki.F90:71-91
double aki(const Tables& t, double OM, double BETA) {
  const double EBS = 0.0001;
  double AKM1 = OM * OM / (4.0 * t.G);
  double AKM2 = OM / (2.0 * std::sqrt(t.G * BETA));
  double AO = std::max(AKM1, AKM2);
  for (;;) {
    double AKP = AO;
    double BO = BETA * AO;
    if (BO > t.DKMAX) return OM * OM / t.G;
    double TH = t.G * AO * std::tanh(BO);
    double STH = std::sqrt(TH);
    double ch = std::cosh(BO);
    AO = AO + (OM - STH) * STH * 2.0 / (TH / AO + t.G * BO / (ch * ch));
    if (!(std::fabs(AKP - AO) > EBS * AO)) return AO;
  }
}

// depthprpt.F90:60-81 ; arrays are (n, NFRE) column-major
void depthprpt(const Tables& t, const Config& c, long n, const double* DEPTH, double* WAVNUM, double* CINV,
               double* CGROUP, double* XK2CG, double* OMOSNH2KD, double* STOKFAC) {
  double GH = t.G / (4.0 * t.PI);
  for (int M = 1; M <= c.nfre; ++M) {
    double OM = t.ZPIFR(M);
    for (long IJ = 0; IJ < n; ++IJ) {
      size_t o = IJ + n * (size_t)(M - 1);
      double AK = aki(t, OM, DEPTH[IJ]);
      WAVNUM[o] = AK;
      double AKD = AK * DEPTH[IJ];
      if (AKD <= 10.0) {
        CGROUP[o] = 0.5 * std::sqrt(t.G * std::tanh(AKD) / AK) * (1.0 + 2.0 * AKD / std::sinh(2.0 * AKD));
        OMOSNH2KD[o] = OM / std::sinh(2.0 * AKD);
        STOKFAC[o] = 2.0 * t.G * AK * AK / (OM * std::tanh(2.0 * AKD));
      } else {
        CGROUP[o] = GH / t.FR(M);
        OMOSNH2KD[o] = 0.0;
        STOKFAC[o] = 2.0 / t.G * OM * OM * OM;
      }
      CINV[o] = WAVNUM[o] / OM;
      XK2CG[o] = WAVNUM[o] * WAVNUM[o] * CGROUP[o];
    }
  }
}

// Allocation in the NPROMA-chunked layout + depth-derived fields (mpdecomp.F90:1424-1456, initdpthflds.F90:57-88).
void alloc_fields(Model& m) {
  const Config& c = m.cfg;
  const Tables& t = m.tab;
  m.fld.assign(c.npr, Fields());
  for (int ir = 0; ir < c.npr; ++ir) {
    RankDecomp& r = m.ranks[ir];
    Fields& f = m.fld[ir];
    const int P = r.NPROMA, C = r.NCHNK, A = c.nang, F = c.nfre;
    f.FL1.alloc(1, P, 1, A, 1, F, 1, C); f.XLLWS.alloc(1, P, 1, A, 1, F, 1, C);
    for (ArrD* a : {&f.WAVNUM, &f.CINV, &f.CGROUP, &f.XK2CG, &f.OMOSNH2KD, &f.STOKFAC, &f.CIWA}) a->alloc(1, P, 1, F, 1, C);
    for (ArrD* a : {&f.DEPTH, &f.EMAXDPT, &f.DELLAM1, &f.COSPHM1, &f.UCUR, &f.VCUR, &f.IBRMEM, &f.AIRD, &f.WDWAVE,
                    &f.CICOVER, &f.WSWAVE, &f.WSTAR, &f.USTRA, &f.VSTRA, &f.UFRIC, &f.TAUW, &f.TAUWDIR, &f.Z0M, &f.Z0B,
                    &f.CHRNCK, &f.CITHICK, &f.WSEMEAN, &f.WSFMEAN, &f.USTOKES, &f.VSTOKES, &f.STRNMS, &f.TAUXD,
                    &f.TAUYD, &f.TAUOCXD, &f.TAUOCYD, &f.TAUOC, &f.TAUICX, &f.TAUICY, &f.PHIOCD, &f.PHIEPS, &f.PHIAW,
                    &f.NSWH, &f.NMWP, &f.NPHIEPS, &f.NTAUOC, &f.NEMOTAUX, &f.NEMOTAUY, &f.NEMOTAUICX, &f.NEMOTAUICY, &f.NEMOWSWAVE,
                    &f.NEMOPHIF, &f.NEMOUSTOKES, &f.NEMOVSTOKES, &f.NEMOSTRN})
      a->alloc(1, P, 1, C);
    f.MIJ.alloc(1, P, 1, C); f.INDEP.alloc(1, P, 1, C); f.IODP.alloc(1, P, 1, C); f.IOBND.alloc(1, P, 1, C);
    for (size_t i = 0; i < f.CIWA.d.size(); ++i) f.CIWA.d[i] = 1.0;
    for (size_t i = 0; i < f.IODP.d.size(); ++i) { f.IODP.d[i] = 1; f.IOBND.d[i] = 1; f.INDEP.d[i] = 1; }
    for (size_t i = 0; i < f.MIJ.d.size(); ++i) f.MIJ.d[i] = c.nfre;
    for (int ICHNK = 1; ICHNK <= C; ++ICHNK) {
      int KIJL = r.KIJL4CHNK(ICHNK);
      for (int IJ = 1; IJ <= P; ++IJ) {
        int ijb = (IJ <= KIJL) ? r.IJFROMCHNK(IJ, ICHNK) : r.IJFROMCHNK(1, ICHNK);  // padding = first point
        int JH = m.grid.KXLT(ijb);
        f.COSPHM1(IJ, ICHNK) = 1.0 / m.grid.COSPH(JH);
        f.DELLAM1(IJ, ICHNK) = 1.0 / m.grid.DELLAM(JH);
        f.DEPTH(IJ, ICHNK) = m.depth0[m.grid.NEWIJ2IJ(ijb) - 1];
        f.UCUR(IJ, ICHNK) = 0.0; f.VCUR(IJ, ICHNK) = 0.0;
        const double GAM_B_J = 0.8;
        double d = f.DEPTH(IJ, ICHNK);
        double GAM = (d < 4.0) ? GAM_B_J * d / 4.0 : GAM_B_J;
        f.EMAXDPT(IJ, ICHNK) = 0.0625 * (GAM * d) * (GAM * d);
      }
      size_t o2 = (size_t)P * (ICHNK - 1), o3 = (size_t)P * F * (ICHNK - 1);
      depthprpt(t, c, P, &f.DEPTH.d[o2], &f.WAVNUM.d[o3], &f.CINV.d[o3], &f.CGROUP.d[o3], &f.XK2CG.d[o3],
                &f.OMOSNH2KD.d[o3], &f.STOKFAC.d[o3]);
    }
    // land point (initdpthflds.F90:80-88)
    f.LAND_WAVNUM.alloc(1, F); f.LAND_CGROUP.alloc(1, F); f.LAND_OMOSNH2KD.alloc(1, F);
    std::vector<double> ci(F), xk(F), st(F);
    double deep = c.bathymax;
    depthprpt(t, c, 1, &deep, f.LAND_WAVNUM.data(), ci.data(), f.LAND_CGROUP.data(), xk.data(), f.LAND_OMOSNH2KD.data(),
              st.data());
  }
}

// mpexchng.F90:120-249, emulated: pack on every rank, then deliver, then unpack.  FLD(r) is (NINF:NSUP+1, ND2, nd3)
static void mpexchng(Model& m, std::vector<ArrD>& FLD, int NDIM2, int ND3S, int ND3E) {
  const int NPR = m.cfg.npr;
  if (NPR <= 1) return;
  // message from rank s to rank d: values FLD_s(IJTOPE_s(IH,d),K,M) in (M,K,IH) order
  for (int d = 1; d <= NPR; ++d) {
    RankDecomp& rd = m.ranks[d - 1];
    for (int INGB = 1; INGB <= rd.NGBFROMPE; ++INGB) {
      int s = rd.NFROMPELST(INGB);
      RankDecomp& rs = m.ranks[s - 1];
      int n = rs.NTOPE(d);
      if (n != rd.NFROMPE(s)) throw std::runtime_error("MPEXCHNG: NTOPE/NFROMPE mismatch");
      for (int M = ND3S; M <= ND3E; ++M)
        for (int K = 1; K <= NDIM2; ++K)
          for (int IH = 1; IH <= n; ++IH)
            FLD[d - 1](rd.NIJSTART(s) + IH - 1, K, M) = FLD[s - 1](rs.IJTOPE(IH, d), K, M);
    }
  }
}

// proenvhalo.F90:67-106 -> BUFFER_EXT(NINF:NSUP+1, 3*NFRE_RED+5)
static void proenvhalo(Model& m, std::vector<ArrD>& BUF) {
  const Config& c = m.cfg;
  const int FR = c.nfre_red;
  for (int ir = 0; ir < c.npr; ++ir) {
    RankDecomp& r = m.ranks[ir];
    Fields& f = m.fld[ir];
    BUF[ir].alloc(r.NINF, r.NSUP + 1, 1, 3 * FR + 5, 1, 1);
    for (int ICHNK = 1; ICHNK <= r.NCHNK; ++ICHNK) {
      for (int IJ = 1; IJ <= r.KIJL4CHNK(ICHNK); ++IJ) {
        int b = r.IJFROMCHNK(IJ, ICHNK);
        for (int M = 1; M <= FR; ++M) {
          BUF[ir](b, M, 1) = f.WAVNUM(IJ, M, ICHNK);
          BUF[ir](b, M + FR, 1) = f.CGROUP(IJ, M, ICHNK);
          BUF[ir](b, M + 2 * FR, 1) = f.OMOSNH2KD(IJ, M, ICHNK);
        }
        BUF[ir](b, 3 * FR + 1, 1) = f.DELLAM1(IJ, ICHNK);
        BUF[ir](b, 3 * FR + 2, 1) = f.COSPHM1(IJ, ICHNK);
        BUF[ir](b, 3 * FR + 3, 1) = f.DEPTH(IJ, ICHNK);
        BUF[ir](b, 3 * FR + 4, 1) = f.UCUR(IJ, ICHNK);
        BUF[ir](b, 3 * FR + 5, 1) = f.VCUR(IJ, ICHNK);
      }
    }
  }
  mpexchng(m, BUF, 3 * FR + 5, 1, 1);
  for (int ir = 0; ir < c.npr; ++ir) {
    RankDecomp& r = m.ranks[ir];
    Fields& f = m.fld[ir];
    int L = r.NSUP + 1;
    for (int M = 1; M <= FR; ++M) {
      BUF[ir](L, M, 1) = f.LAND_WAVNUM(M);
      BUF[ir](L, FR + M, 1) = f.LAND_CGROUP(M);
      BUF[ir](L, 2 * FR + M, 1) = f.LAND_OMOSNH2KD(M);
    }
    BUF[ir](L, 3 * FR + 1, 1) = 0.0; BUF[ir](L, 3 * FR + 2, 1) = 0.0; BUF[ir](L, 3 * FR + 3, 1) = c.bathymax;
    BUF[ir](L, 3 * FR + 4, 1) = 0.0; BUF[ir](L, 3 * FR + 5, 1) = 0.0;
  }
}

// ctuwupdt.F90:93-166 index helpers
static void ctu_index_tables(const Config& c, const Tables& t, RankDecomp& r) {
  const int NANG = c.nang, FR = c.nfre_red;
  r.MPM.alloc(1, FR, -1, 1); r.KPM.alloc(1, NANG, -1, 1); r.JXO.alloc(1, NANG, 1, 2); r.JYO.alloc(1, NANG, 1, 2);
  r.KCR.alloc(1, NANG, 1, 4);
  for (int M = 1; M <= FR; ++M) { r.MPM(M, -1) = std::max(1, M - 1); r.MPM(M, 0) = M; r.MPM(M, 1) = std::min(FR, M + 1); }
  for (int K = 1; K <= NANG; ++K) {
    int KM1 = K - 1; if (KM1 < 1) KM1 = NANG;
    r.KPM(K, -1) = KM1; r.KPM(K, 0) = K;
    int KP1 = K + 1; if (KP1 > NANG) KP1 = 1;
    r.KPM(K, 1) = KP1;
    if (t.COSTH(K) >= 0.0) {
      r.JYO(K, 1) = 1; r.JYO(K, 2) = 2;
      if (t.SINTH(K) >= 0.0) { r.JXO(K, 1) = 1; r.JXO(K, 2) = 2; r.KCR(K, 1) = 3; r.KCR(K, 2) = 2; r.KCR(K, 3) = 4; r.KCR(K, 4) = 1; }
      else { r.JXO(K, 1) = 2; r.JXO(K, 2) = 1; r.KCR(K, 1) = 2; r.KCR(K, 2) = 3; r.KCR(K, 3) = 1; r.KCR(K, 4) = 4; }
    } else {
      r.JYO(K, 1) = 2; r.JYO(K, 2) = 1;
      if (t.SINTH(K) >= 0.0) { r.JXO(K, 1) = 1; r.JXO(K, 2) = 2; r.KCR(K, 1) = 4; r.KCR(K, 2) = 1; r.KCR(K, 3) = 3; r.KCR(K, 4) = 2; }
      else { r.JXO(K, 1) = 2; r.JXO(K, 2) = 1; r.KCR(K, 1) = 1; r.KCR(K, 2) = 4; r.KCR(K, 3) = 2; r.KCR(K, 4) = 3; }
    }
  }
}

// ctuwini.F90:60-165 (modifies WLAT/WCOR near land) + ctuw.F90:110-275, :404-501, :531-690, :700-733
// (ICASE=1; obstruction coefficients from RankDecomp::OBS*, 1 when LSUBGRID=F) + ctuwdrv.F90 CFL flagging.
static void ctuwupdt(Model& m, int ir, ArrD& BUF) {
  const Config& c = m.cfg;
  const Tables& t = m.tab;
  const Grid& g = m.grid;
  RankDecomp& r = m.ranks[ir];
  const int NANG = c.nang, FR = c.nfre_red, IJS = r.IJS, IJL = r.IJL, NSUP = r.NSUP, NLAND = NSUP + 1;
  ctu_index_tables(c, t, r);
  const bool CUR = c.irefra == 2 || c.irefra == 3;
  const bool FULL = c.store_all_weights != 0 || CUR;
  r.W8.alloc(IJS, IJL, 1, NANG, 1, FR, 1, 8);
  if (CUR) r.WMPMN.alloc(IJS, IJL, 1, NANG, 1, FR, -1, 1);
  if (FULL) {
    r.SUMWN.alloc(IJS, IJL, 1, NANG, 1, FR);
    r.WLATN.alloc(IJS, IJL, 1, NANG, 1, FR, 1, 2, 1, 2);
    r.WLONN.alloc(IJS, IJL, 1, NANG, 1, FR, 1, 2);
    r.WCORN.alloc(IJS, IJL, 1, NANG, 1, FR, 1, 4, 1, 2);
    r.WKPMN.alloc(IJS, IJL, 1, NANG, 1, FR, -1, 1);
  }
  // PROPDOT / GRADI (propag_wam.F90:171-216 runs them BEFORE CTUWUPDT, i.e. with WLAT as PROPCONNECT left it): depth part
  // THDD (IREFRA = 1; gradi.F90:120-153, propdot.F90:117-164), current part THDC, SDOT (IREFRA = 2, 3; gradi.F90:167-229,
  // propdot.F90:166-200), ICASE = 1
  ArrD THDD, THDC, SDOT;
  if (c.irefra < 0 || c.irefra > 3) throw std::runtime_error("CTUW: IREFRA must be 0..3");
  if (c.irefra != 0) {
    THDD.alloc(IJS, IJL, 1, NANG);
    if (CUR) { THDC.alloc(IJS, IJL, 1, NANG); SDOT.alloc(IJS, IJL, 1, NANG, 1, FR); }
    const double DELPHI = g.XDELLA * t.CIRC / 360.0;   // readmdlconf.F90:136
    const double ONEO2DELPHI = 0.5 / DELPHI;
    const double CURRENT_GRADIENT_MAX = 0.00001;       // yowcurr.F90:19
    auto DPTHEXT = [&](int ij) { return BUF(ij, 3 * FR + 3, 1); };
    auto UEXT = [&](int ij) { return BUF(ij, 3 * FR + 4, 1); };
    auto VEXT = [&](int ij) { return BUF(ij, 3 * FR + 5, 1); };
    for (int IJ = IJS; IJ <= IJL; ++IJ) {
      double DDPHI = 0.0, DDLAM = 0.0, DUPHI = 0.0, DVPHI = 0.0, DULAM = 0.0, DVLAM = 0.0;
      const int KX = g.KXLT(IJ);
      if (c.irefra == 1 || c.irefra == 3) {
        const int IPP = r.KLAT(IJ, 2, 1), IPM = r.KLAT(IJ, 1, 1), IPP2 = r.KLAT(IJ, 2, 2), IPM2 = r.KLAT(IJ, 1, 2);
        if (IPP != NLAND && IPM != NLAND && IPP2 != NLAND && IPM2 != NLAND) {
          const double DPTP = r.WLAT(IJ, 2) * DPTHEXT(IPP) + (1.0 - r.WLAT(IJ, 2)) * DPTHEXT(IPP2);
          const double DPTM = r.WLAT(IJ, 1) * DPTHEXT(IPM) + (1.0 - r.WLAT(IJ, 1)) * DPTHEXT(IPM2);
          DDPHI = (DPTP - DPTM) * ONEO2DELPHI;
        } else if (IPP != NLAND && IPM != NLAND) {
          DDPHI = (DPTHEXT(IPP) - DPTHEXT(IPM)) * ONEO2DELPHI;
        } else if (IPP2 != NLAND && IPM2 != NLAND) {
          DDPHI = (DPTHEXT(IPP2) - DPTHEXT(IPM2)) * ONEO2DELPHI;
        } else DDPHI = 0.0;
        const int ILP = r.KLON(IJ, 2), ILM = r.KLON(IJ, 1);
        if (ILP != NLAND && ILM != NLAND) DDLAM = (DPTHEXT(ILP) - DPTHEXT(ILM)) / (2. * g.DELLAM(KX));
        else DDLAM = 0.0;
      }
      if (CUR) {
        // exact 0 means that the current field was not defined: no gradient is extrapolated (gradi.F90:171-181, 205-210)
        auto undef = [&](int ij) { return UEXT(ij) == 0.0 && VEXT(ij) == 0.0; };
        int IPP = r.KLAT(IJ, 2, 1); if (undef(IPP)) IPP = NLAND;
        int IPM = r.KLAT(IJ, 1, 1); if (undef(IPM)) IPM = NLAND;
        int IPP2 = r.KLAT(IJ, 2, 2); if (undef(IPP2)) IPP2 = NLAND;
        int IPM2 = r.KLAT(IJ, 1, 2); if (undef(IPM2)) IPM2 = NLAND;
        if (IPP != NLAND && IPM != NLAND && IPP2 != NLAND && IPM2 != NLAND) {
          const double UP = r.WLAT(IJ, 2) * UEXT(IPP) + (1.0 - r.WLAT(IJ, 2)) * UEXT(IPP2);
          const double VP = r.WLAT(IJ, 2) * VEXT(IPP) + (1.0 - r.WLAT(IJ, 2)) * VEXT(IPP2);
          const double UM = r.WLAT(IJ, 1) * UEXT(IPM) + (1.0 - r.WLAT(IJ, 1)) * UEXT(IPM2);
          const double VM = r.WLAT(IJ, 1) * VEXT(IPM) + (1.0 - r.WLAT(IJ, 1)) * VEXT(IPM2);
          DUPHI = (UP - UM) * ONEO2DELPHI; DVPHI = (VP - VM) * ONEO2DELPHI;
        } else if (IPP != NLAND && IPM != NLAND) {
          DUPHI = (UEXT(IPP) - UEXT(IPM)) * ONEO2DELPHI; DVPHI = (VEXT(IPP) - VEXT(IPM)) * ONEO2DELPHI;
        } else { DUPHI = 0.0; DVPHI = 0.0; }
        int ILP = r.KLON(IJ, 2); if (undef(ILP)) ILP = NLAND;
        int ILM = r.KLON(IJ, 1); if (undef(ILM)) ILM = NLAND;
        if (ILP != NLAND && ILM != NLAND) {
          DULAM = (UEXT(ILP) - UEXT(ILM)) / (2.0 * g.DELLAM(KX)); DVLAM = (VEXT(ILP) - VEXT(ILM)) / (2.0 * g.DELLAM(KX));
        } else { DULAM = 0.0; DVLAM = 0.0; }
        const double CGMAX = CURRENT_GRADIENT_MAX * g.COSPH(KX);
        DUPHI = sign(std::min(std::fabs(DUPHI), CGMAX), DUPHI); DVPHI = sign(std::min(std::fabs(DVPHI), CGMAX), DVPHI);
        DULAM = sign(std::min(std::fabs(DULAM), CGMAX), DULAM); DVLAM = sign(std::min(std::fabs(DVLAM), CGMAX), DVLAM);
      }
      const double DCO = BUF(IJ, 3 * FR + 2, 1);       // COSPHM1_EXT
      double OMDD = 0.0;
      if (c.irefra == 3) OMDD = VEXT(IJ) * DDPHI + UEXT(IJ) * DDLAM * DCO;
      for (int K = 1; K <= NANG; ++K) {
        const double SD = t.SINTH(K), CD = t.COSTH(K);
        THDD(IJ, K) = (c.irefra == 1 || c.irefra == 3) ? SD * DDPHI - CD * DDLAM * DCO : 0.0;
        if (CUR) {
          const double SS = SD * SD, SC = SD * CD, CC = CD * CD;
          const double S0 = -SC * DUPHI - CC * DVPHI - (SS * DULAM + SC * DVLAM) * DCO;
          THDC(IJ, K) = SS * DUPHI + SC * DVPHI - (SC * DULAM + CC * DVLAM) * DCO;
          for (int M = 1; M <= FR; ++M) SDOT(IJ, K, M) = (S0 * BUF(IJ, FR + M, 1) + OMDD * BUF(IJ, 2 * FR + M, 1)) * BUF(IJ, M, 1);
        }
      }
    }
  }
  ArrD WLATM1, WCORM1, DP;
  WLATM1.alloc(IJS, IJL, 1, 2); WCORM1.alloc(IJS, IJL, 1, 4); DP.alloc(IJS, IJL, 1, 2);
  // CTUWINI
  for (int IC = 1; IC <= 2; ++IC)
    for (int IJ = IJS; IJ <= IJL; ++IJ) {
      if (r.KLAT(IJ, IC, 1) < NLAND && r.KLAT(IJ, IC, 2) < NLAND) {
        WLATM1(IJ, IC) = 1.0 - r.WLAT(IJ, IC);
      } else if (r.KLAT(IJ, IC, 1) == NLAND) {
        if (r.WLAT(IJ, IC) <= 0.75) r.WLAT(IJ, IC) = 0.0;
        WLATM1(IJ, IC) = 1.0 - r.WLAT(IJ, IC);
      } else {
        if (r.WLAT(IJ, IC) >= 0.5) r.WLAT(IJ, IC) = 1.0;
        WLATM1(IJ, IC) = 1.0 - r.WLAT(IJ, IC);
      }
    }
  for (int ICR = 1; ICR <= 4; ++ICR)
    for (int IJ = IJS; IJ <= IJL; ++IJ) {
      if (r.KCOR(IJ, ICR, 1) < NLAND && r.KCOR(IJ, ICR, 2) < NLAND) {
        WCORM1(IJ, ICR) = 1.0 - r.WCOR(IJ, ICR);
      } else if (r.KCOR(IJ, ICR, 1) == NLAND) {
        if (r.WCOR(IJ, ICR) <= 0.75) r.WCOR(IJ, ICR) = 0.0;
        WCORM1(IJ, ICR) = 1.0 - r.WCOR(IJ, ICR);
      } else {
        if (r.WCOR(IJ, ICR) > 0.5) r.WCOR(IJ, ICR) = 1.0;
        WCORM1(IJ, ICR) = 1.0 - r.WCOR(IJ, ICR);
      }
    }
  for (int IC = 1; IC <= 2; ++IC)
    for (int IJ = IJS; IJ <= IJL; ++IJ) {
      int KY = g.KXLT(IJ);
      int KK = KY + 2 * IC - 3;
      int KKM = std::max(1, std::min(KK, g.NGY));
      DP(IJ, IC) = g.COSPH(KKM) * BUF(IJ, 3 * FR + 2, 1);
    }
  // CTUWDRV: one or two calls of CTUW depending on the fast-wave split (ctuwupdt.F90:193-235)
  auto CG = [&](int ij, int M) -> double { return BUF(ij, FR + M, 1); };
  std::vector<char> LCFLFAIL(IJL - IJS + 1, 0);
  std::vector<double> CURMASK(IJL - IJS + 1, 1.0);
  auto ctuw = [&](double DELPRO, int MSTART, int MEND) {
    const double CMTODEG = 360.0 / t.CIRC;
    const double XDELLA = g.XDELLA;
#pragma omp parallel for schedule(static)
    for (int IJ = IJS; IJ <= IJL; ++IJ) {
      const double COSPHM1 = BUF(IJ, 3 * FR + 2, 1);
      const int KY = g.KXLT(IJ);
      const double ZDELLO = g.ZDELLO(KY);
      for (int M = MSTART; M <= MEND; ++M) {
        for (int K = 1; K <= NANG; ++K) {
          double CGX[3], CGY[3], ADXP[3], ADYP[3], DXUP[3], DXDW[3], DYUP[3], DYDW[3], WEIGHT[5];
          for (int IC = 1; IC <= 2; ++IC) {
            CGX[IC] = 0.5 * (CG(IJ, M) + CG(r.KLON(IJ, IC), M)) * t.SINTH(K) * COSPHM1;
            double CGYP = r.WLAT(IJ, IC) * CG(r.KLAT(IJ, IC, 1), M) + (1.0 - r.WLAT(IJ, IC)) * CG(r.KLAT(IJ, IC, 2), M);
            CGY[IC] = 0.5 * (CG(IJ, M) + DP(IJ, IC) * CGYP) * t.COSTH(K);
            double UREL = CGX[IC], VREL = CGY[IC];
            int ISSU = 1, ISSV = 1;
            if (CUR) {   // ctuw.F90:175-181, 211-217
              auto isamesign = [](double a, double b) { return sign(1.0, a) == sign(1.0, b) ? 1 : 0; };
              const double UU = BUF(IJ, 3 * FR + 4, 1) * COSPHM1;
              UREL = CGX[IC] + UU; ISSU = isamesign(UREL, CGX[IC]);
              const double VV = BUF(IJ, 3 * FR + 5, 1) * 0.5 * (1.0 + DP(IJ, IC));
              VREL = CGY[IC] + VV; ISSV = isamesign(VREL, CGY[IC]);
            }
            double DXP = -DELPRO * UREL * CMTODEG, DYP = -DELPRO * VREL * CMTODEG;
            ADXP[IC] = std::fabs(DXP); ADYP[IC] = std::fabs(DYP);
            DXUP[IC] = ADXP[IC] * ISSU; DXDW[IC] = ADXP[IC] * (1 - ISSU);
            DYUP[IC] = ADYP[IC] * ISSV; DYDW[IC] = ADYP[IC] * (1 - ISSV);
          }
          const int JX1 = r.JXO(K, 1), JX2 = r.JXO(K, 2), JY1 = r.JYO(K, 1), JY2 = r.JYO(K, 2);
          double DXX = ZDELLO - DXUP[JX2] - DXDW[JX1];
          double DYY = XDELLA - DYUP[JY2] - DYDW[JY1];
          double GRIDAREAM1 = 1.0 / (ZDELLO * XDELLA);
          WEIGHT[JY1] = DXX * DYUP[JY1] * GRIDAREAM1;
          WEIGHT[JY2] = DXX * DYDW[JY2] * GRIDAREAM1;
          double wlatn[3][3], wlonn[3], wcorn[5][3];
          wlatn[1][1] = r.WLAT(IJ, 1) * WEIGHT[1];
          wlatn[1][2] = WLATM1(IJ, 1) * WEIGHT[1];
          wlatn[2][1] = r.WLAT(IJ, 2) * WEIGHT[2];
          wlatn[2][2] = WLATM1(IJ, 2) * WEIGHT[2];
          wlonn[JX1] = DYY * DXUP[JX1] * GRIDAREAM1;
          wlonn[JX2] = DYY * DXDW[JX2] * GRIDAREAM1;
          WEIGHT[1] = DXUP[JX1] * DYUP[JY1] * GRIDAREAM1;
          WEIGHT[2] = DXDW[JX2] * DYUP[JY1] * GRIDAREAM1;
          WEIGHT[3] = DXUP[JX1] * DYDW[JY2] * GRIDAREAM1;
          WEIGHT[4] = DXDW[JX2] * DYDW[JY2] * GRIDAREAM1;
          for (int ICR = 1; ICR <= 4; ++ICR) {
            wcorn[ICR][1] = r.WCOR(IJ, r.KCR(K, ICR)) * WEIGHT[ICR];
            wcorn[ICR][2] = WCORM1(IJ, r.KCR(K, ICR)) * WEIGHT[ICR];
          }
          double sumwn = (ZDELLO * (DYDW[JY1] + DYUP[JY2]) + XDELLA * (DXUP[JX2] + DXDW[JX1]) -
                          (DXDW[JX1] + DXUP[JX2]) * (DYDW[JY1] + DYUP[JY2])) * GRIDAREAM1;
          // weight range checks (:536-590)
          for (int a = 1; a <= 2; ++a) {
            if (wlonn[a] > 1.0 || wlonn[a] < 0.0) LCFLFAIL[IJ - IJS] = 1;
            for (int b = 1; b <= 2; ++b) if (wlatn[a][b] > 1.0 || wlatn[a][b] < 0.0) LCFLFAIL[IJ - IJS] = 1;
          }
          for (int a = 1; a <= 4; ++a)
            for (int b = 1; b <= 2; ++b) if (wcorn[a][b] > 1.0 || wcorn[a][b] < 0.0) LCFLFAIL[IJ - IJS] = 1;
          r.W8(IJ, K, M, 1) = sumwn;  // WKPMN(0) is added below
          r.W8(IJ, K, M, 2) = wlonn[JX1];
          r.W8(IJ, K, M, 3) = wlatn[JY1][1];
          r.W8(IJ, K, M, 4) = wlatn[JY1][2];
          r.W8(IJ, K, M, 5) = wcorn[1][1];
          r.W8(IJ, K, M, 6) = wcorn[1][2];
          if (FULL) {
            for (int a = 1; a <= 2; ++a) { r.WLONN(IJ, K, M, a) = wlonn[a]; for (int b = 1; b <= 2; ++b) r.WLATN(IJ, K, M, a, b) = wlatn[a][b]; }
            for (int a = 1; a <= 4; ++a) for (int b = 1; b <= 2; ++b) r.WCORN(IJ, K, M, a, b) = wcorn[a][b];
            r.SUMWN(IJ, K, M) = sumwn;
          }
          // basic CFL checks (:282-358)
          if (ADXP[1] > ZDELLO || ADYP[1] > XDELLA || ADXP[2] > ZDELLO || ADYP[2] > XDELLA) LCFLFAIL[IJ - IJS] = 1;
        }
      }
    }
    // refraction (grid only, IREFRA=0) :404-501
    const double DELTH0 = 0.25 * DELPRO / t.DELTH;
#pragma omp parallel for schedule(static)
    for (int K = 1; K <= NANG; ++K) {
      int KP1 = K + 1; if (KP1 > NANG) KP1 = 1;
      int KM1 = K - 1; if (KM1 < 1) KM1 = NANG;
      double SP = DELTH0 * (t.SINTH(K) + t.SINTH(KP1)) / t.R;
      double SM = DELTH0 * (t.SINTH(K) + t.SINTH(KM1)) / t.R;
      for (int M = MSTART; M <= MEND; ++M)
        for (int IJ = IJS; IJ <= IJL; ++IJ) {
          int JH = g.KXLT(IJ);
          double TANPH = g.SINPH(JH) / g.COSPH(JH);
          double DRGP = TANPH * SP, DRGM = TANPH * SM;
          double DRCP = 0.0, DRCM = 0.0;
          if (CUR) {                               // ctuw.F90:451-456
            DRCP = CURMASK[IJ - IJS] * (THDC(IJ, K) + THDC(IJ, KP1)) * DELTH0;
            DRCM = CURMASK[IJ - IJS] * (THDC(IJ, K) + THDC(IJ, KM1)) * DELTH0;
          }
          double DTHP, DTHM;
          if (c.irefra == 0) {                     // ctuw.F90:471-486
            DTHP = DRGP * CG(IJ, M) + DRCP;
            DTHM = DRGM * CG(IJ, M) + DRCM;
          } else {                                 // ctuw.F90:434-439, 487-501 (depth refraction only for IREFRA = 1)
            const double DRDP = c.irefra == 1 ? (THDD(IJ, K) + THDD(IJ, KP1)) * DELTH0 : 0.0;
            const double DRDM = c.irefra == 1 ? (THDD(IJ, K) + THDD(IJ, KM1)) * DELTH0 : 0.0;
            DTHP = DRGP * CG(IJ, M) + BUF(IJ, 2 * FR + M, 1) * DRDP + DRCP;
            DTHM = DRGM * CG(IJ, M) + BUF(IJ, 2 * FR + M, 1) * DRDM + DRCM;
          }
          double w0 = (DTHP + std::fabs(DTHP)) + (std::fabs(DTHM) - DTHM);
          double wp = -DTHP + std::fabs(DTHP);
          double wm = DTHM + std::fabs(DTHM);
          if (w0 > 1.0 || w0 < 0.0 || wp > 1.0 || wp < 0.0 || wm > 1.0 || wm < 0.0) LCFLFAIL[IJ - IJS] = 1;
          // SUMWN = SUMWN + WKPMN(0) (:608) and the SUMWN range check (:636)
          double s = r.W8(IJ, K, M, 1) + w0;
          r.W8(IJ, K, M, 1) = s;
          r.W8(IJ, K, M, 7) = wm;
          r.W8(IJ, K, M, 8) = wp;
          if (CUR) {   // frequency shifting due to currents (ctuw.F90:503-525) + its checks and share of SUMWN (:611-633)
            const double DELFR0 = 0.25 * DELPRO / ((t.FRATIO - 1) * t.ZPI);
            const int MP1 = std::min(FR, M + 1), MM1 = std::max(1, M - 1);
            const double DFP = DELFR0 / t.FR(M), DFM = DELFR0 / t.FR(MM1);
            const double DTP = CURMASK[IJ - IJS] * (SDOT(IJ, K, M) + SDOT(IJ, K, MP1)) * DFP;
            const double DTM = CURMASK[IJ - IJS] * (SDOT(IJ, K, M) + SDOT(IJ, K, MM1)) * DFM;
            const double f0 = (DTP + std::fabs(DTP)) + (std::fabs(DTM) - DTM);
            const double fp = (-DTP + std::fabs(DTP)) / t.FRATIO;
            const double fm = (DTM + std::fabs(DTM)) * t.FRATIO;
            if (f0 > 1.0 || f0 < 0.0 || fp > 1.0 || fp < 0.0 || fm > 1.0 || fm < 0.0) LCFLFAIL[IJ - IJS] = 1;
            r.WMPMN(IJ, K, M, 0) = f0; r.WMPMN(IJ, K, M, 1) = fp; r.WMPMN(IJ, K, M, -1) = fm;
            s = s + f0;
          }
          if (s > 1.0 || s < 0.0) LCFLFAIL[IJ - IJS] = 1;
          if (FULL) { r.WKPMN(IJ, K, M, 0) = w0; r.WKPMN(IJ, K, M, 1) = wp; r.WKPMN(IJ, K, M, -1) = wm; r.SUMWN(IJ, K, M) = s; }
        }
    }
    // the blocking coefficients go into the weights of the surrounding points, not into SUMWN (ctuw.F90:700-733); LSUBGRID = F: all 1
    if (r.OBSLON.size() > 0) {
      for (int K = 1; K <= NANG; ++K) for (int M = MSTART; M <= MEND; ++M) for (int IJ = IJS; IJ <= IJL; ++IJ) {
        const int JX1 = r.JXO(K, 1), JY1 = r.JYO(K, 1);
        r.W8(IJ, K, M, 2) = r.W8(IJ, K, M, 2) * r.OBSLON(IJ, M, JX1);
        r.W8(IJ, K, M, 3) = r.W8(IJ, K, M, 3) * r.OBSLAT(IJ, M, JY1);
        r.W8(IJ, K, M, 4) = r.W8(IJ, K, M, 4) * r.OBSLAT(IJ, M, JY1);
        r.W8(IJ, K, M, 5) = r.W8(IJ, K, M, 5) * r.OBSCOR(IJ, M, r.KCR(K, 1));
        r.W8(IJ, K, M, 6) = r.W8(IJ, K, M, 6) * r.OBSCOR(IJ, M, r.KCR(K, 1));
        if (FULL) {
          for (int IC = 1; IC <= 2; ++IC) for (int ICL = 1; ICL <= 2; ++ICL) r.WLATN(IJ, K, M, IC, ICL) = r.WLATN(IJ, K, M, IC, ICL) * r.OBSLAT(IJ, M, IC);
          for (int IC = 1; IC <= 2; ++IC) r.WLONN(IJ, K, M, IC) = r.WLONN(IJ, K, M, IC) * r.OBSLON(IJ, M, IC);
          for (int ICR = 1; ICR <= 4; ++ICR) for (int ICL = 1; ICL <= 2; ++ICL) r.WCORN(IJ, K, M, ICR, ICL) = r.WCORN(IJ, K, M, ICR, ICL) * r.OBSCOR(IJ, M, r.KCR(K, ICR));
        }
      }
    }
  };
  // CTUWDRV (ctuwdrv.F90:83-123): ICALL = 1 with the currents everywhere; with LLCFLCUROFF a second call switches the current
  // REFRACTION (not the advection by the current) off at the points that failed (CURMASK, ctuw.F90:113-127)
  std::vector<char> FAILALL(IJL - IJS + 1, 0);
  auto ctuwdrv = [&](double DELPRO, int MSTART, int MEND) {
    std::fill(LCFLFAIL.begin(), LCFLFAIL.end(), 0);
    std::fill(CURMASK.begin(), CURMASK.end(), 1.0);
    ctuw(DELPRO, MSTART, MEND);
    if (c.llcflcuroff && CUR) {
      bool any = false;
      for (char f : LCFLFAIL) any = any || f;
      if (any) {
        for (size_t i = 0; i < LCFLFAIL.size(); ++i) CURMASK[i] = LCFLFAIL[i] ? 0.0 : 1.0;
        std::fill(LCFLFAIL.begin(), LCFLFAIL.end(), 0);
        ctuw(DELPRO, MSTART, MEND);
      }
    }
    for (size_t i = 0; i < LCFLFAIL.size(); ++i) FAILALL[i] = FAILALL[i] || LCFLFAIL[i];
  };
  if (c.ifrelfmax <= 0) {
    ctuwdrv(c.idelpro, 1, FR);
  } else {
    ctuwdrv(c.delpro_lf, 1, c.ifrelfmax);
    if (c.ifrelfmax < FR) ctuwdrv(c.idelpro, c.ifrelfmax + 1, FR);
  }
  r.cfl_fail = 0;
  for (char f : FAILALL) r.cfl_fail += f;
}

// propags2.F90:99-121
static void propags2(const Config& c, const RankDecomp& r, const ArrD& F1c, ArrD& F3, int KIJS, int KIJL, int ND3S,
                     int ND3E) {
  ArrD& F1 = const_cast<ArrD&>(F1c);
  RankDecomp& rr = const_cast<RankDecomp&>(r);
  const int NANG = c.nang;
  if (c.irefra == 2 || c.irefra == 3) {   // propags2.F90:123-194 (depth and current refraction: all neighbours, frequency shift)
#pragma omp parallel for collapse(2) schedule(static)
    for (int M = ND3S; M <= ND3E; ++M) {
      for (int K = 1; K <= NANG; ++K) {
        for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
          double v = (1.0 - rr.SUMWN(IJ, K, M)) * F1(IJ, K, M);
          for (int IC = 1; IC <= 2; ++IC) v = v + rr.WLONN(IJ, K, M, IC) * F1(rr.KLON(IJ, IC), K, M);
          for (int ICL = 1; ICL <= 2; ++ICL) {
            for (int IC = 1; IC <= 2; ++IC) v = v + rr.WLATN(IJ, K, M, IC, ICL) * F1(rr.KLAT(IJ, IC, ICL), K, M);
            for (int ICR = 1; ICR <= 4; ++ICR) v = v + rr.WCORN(IJ, K, M, ICR, ICL) * F1(rr.KCOR(IJ, rr.KCR(K, ICR), ICL), K, M);
          }
          for (int IC = -1; IC <= 1; IC += 2) {
            v = v + rr.WKPMN(IJ, K, M, IC) * F1(IJ, rr.KPM(K, IC), M);
            v = v + rr.WMPMN(IJ, K, M, IC) * F1(IJ, K, rr.MPM(M, IC));
          }
          F3(IJ, K, M) = v;
        }
      }
    }
    return;
  }
#pragma omp parallel for collapse(2) schedule(static)
  for (int M = ND3S; M <= ND3E; ++M) {
    for (int K = 1; K <= NANG; ++K) {
      const int jx = rr.JXO(K, 1), jy = rr.JYO(K, 1), kc = rr.KCR(K, 1), km = rr.KPM(K, -1), kp = rr.KPM(K, 1);
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
        F3(IJ, K, M) = (1.0 - rr.W8(IJ, K, M, 1)) * F1(IJ, K, M) +
                       rr.W8(IJ, K, M, 2) * F1(rr.KLON(IJ, jx), K, M) +
                       rr.W8(IJ, K, M, 3) * F1(rr.KLAT(IJ, jy, 1), K, M) +
                       rr.W8(IJ, K, M, 4) * F1(rr.KLAT(IJ, jy, 2), K, M) +
                       rr.W8(IJ, K, M, 5) * F1(rr.KCOR(IJ, kc, 1), K, M) +
                       rr.W8(IJ, K, M, 6) * F1(rr.KCOR(IJ, kc, 2), K, M) +
                       rr.W8(IJ, K, M, 7) * F1(IJ, km, M) + rr.W8(IJ, K, M, 8) * F1(IJ, kp, M);
      }
    }
  }
}

// propag_wam.F90:105-405 for every emulated rank.
void propag_wam(Model& m) {
  const Config& c = m.cfg;
  const int NPR = c.npr, NANG = c.nang, FR = c.nfre_red;
  std::vector<ArrD> FL1_EXT(NPR), FL3_EXT(NPR);
  for (int ir = 0; ir < NPR; ++ir) {
    RankDecomp& r = m.ranks[ir];
    Fields& f = m.fld[ir];
    FL1_EXT[ir].alloc(r.NINF, r.NSUP + 1, 1, NANG, 1, FR);
    FL3_EXT[ir].alloc(r.NINF, r.NSUP + 1, 1, NANG, 1, FR);
    // :119-142 chunk -> block
#pragma omp parallel for schedule(static)
    for (int ICHNK = 1; ICHNK <= r.NCHNK; ++ICHNK) {
      int KIJL = r.KIJL4CHNK(ICHNK), IJSB = r.IJFROMCHNK(1, ICHNK);
      for (int M = 1; M <= FR; ++M)
        for (int K = 1; K <= NANG; ++K)
          for (int IJ = 1; IJ <= KIJL; ++IJ) FL1_EXT[ir](IJ - 1 + IJSB, K, M) = f.FL1(IJ, K, M, ICHNK);
    }
    // :145-147 land slot (already zero after alloc)
  }
  mpexchng(m, FL1_EXT, NANG, 1, FR);  // :166
  bool need_w = false;
  for (int ir = 0; ir < NPR; ++ir) need_w = need_w || m.ranks[ir].LUPDTWGHT;
  if (need_w) {  // :221-236
    std::vector<ArrD> BUF(NPR);
    proenvhalo(m, BUF);
    for (int ir = 0; ir < NPR; ++ir) { ctuwupdt(m, ir, BUF[ir]); m.ranks[ir].LUPDTWGHT = false; }
  }
  for (int ir = 0; ir < NPR; ++ir) {  // :245-251
    RankDecomp& r = m.ranks[ir];
    propags2(c, r, FL1_EXT[ir], FL3_EXT[ir], r.IJS, r.IJL, 1, FR);
  }
  if (c.ifrelfmax > 0 && c.ifrelfmax < FR) {  // :257-313 fast-wave sub-steps
    int NSTEP_LF = (int)nint(c.idelpro / c.delpro_lf);
    int ISUBST = 2;
    const int ND3S = 1, ND3E = c.ifrelfmax;
    while (ISUBST <= NSTEP_LF) {
      for (int ir = 0; ir < NPR; ++ir) {
        RankDecomp& r = m.ranks[ir];
        for (int M = ND3S; M <= ND3E; ++M)
          for (int K = 1; K <= NANG; ++K)
            for (int IJ = r.IJS; IJ <= r.IJL; ++IJ) FL1_EXT[ir](IJ, K, M) = FL3_EXT[ir](IJ, K, M);
      }
      mpexchng(m, FL1_EXT, NANG, ND3S, ND3E);
      for (int ir = 0; ir < NPR; ++ir) {
        RankDecomp& r = m.ranks[ir];
        propags2(c, r, FL1_EXT[ir], FL3_EXT[ir], r.IJS, r.IJL, ND3S, ND3E);
      }
      ISUBST++;
    }
  }
  for (int ir = 0; ir < NPR; ++ir) {  // :368-405 block -> chunk + padding
    RankDecomp& r = m.ranks[ir];
    Fields& f = m.fld[ir];
#pragma omp parallel for schedule(static)
    for (int ICHNK = 1; ICHNK <= r.NCHNK; ++ICHNK) {
      int KIJL = r.KIJL4CHNK(ICHNK), IJSB = r.IJFROMCHNK(1, ICHNK);
      for (int M = 1; M <= FR; ++M)
        for (int K = 1; K <= NANG; ++K) {
          for (int J = 1; J <= KIJL; ++J) f.FL1(J, K, M, ICHNK) = FL3_EXT[ir](IJSB + J - 1, K, M);
          for (int J = KIJL + 1; J <= r.NPROMA; ++J) f.FL1(J, K, M, ICHNK) = f.FL1(1, K, M, ICHNK);
        }
    }
  }
}

}  // namespace orc
