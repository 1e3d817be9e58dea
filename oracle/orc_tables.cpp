// ORACLE (test infrastructure) — one-off tables feeding the hot path.
#include "oracle.h"

namespace orc {

// iniwcst.F90 (whole routine)
static void iniwcst(Tables& t) {
  t.PI = 4.0 * std::atan(1.0);
  t.ZPI = 2.0 * t.PI;
  t.ZPI4GM1 = std::pow(t.ZPI, 4) / t.G;
  t.ZPI4GM2 = std::pow(t.ZPI, 4) / (t.G * t.G);
  t.RAD = t.PI / 180.0;
  t.DEG = 180. / t.PI;
  t.R = t.CIRC / t.ZPI * 1.0;  // PRPLRADI = 1
  t.ROWATERM1 = 1.0 / t.ROWATER;
  t.EPSU10 = std::sqrt(1.0e-3);
}

// setwavphys.F90:46-205
static void setwavphys(const Config& c, Tables& t) {
  if (c.iphys == 0) {
    t.ZALP = 0.008; t.TAILFACTOR = 2.5; t.ALPHAMIN = 0.0001; t.ALPHAPMAX = 0.03; t.TAUWSHELTER = 0.0;
    t.DELTA_THETA_RN = 0.75; t.DTHRN_A = 0.80; t.DTHRN_U = 33.0; t.RN1_RN = 0.25; t.TAILFACTOR_PM = 0.0;
    if (c.llgcbz0) {
      t.ALPHA = 0.0055; t.CHNKMIN_U = 28.; t.BETAMAX = c.llnormagam ? 1.32 : 1.25;
      t.CDIS = -1.3; t.DELTA_SDIS = 0.6; t.CDISVIS = -4.0;
    } else {
      t.ALPHA = 0.0065; t.CHNKMIN_U = 33.; t.BETAMAX = 1.20; t.CDIS = -1.33; t.DELTA_SDIS = 0.5; t.CDISVIS = 0.0;
    }
    t.EGRCRV = 1108.0; t.AFCRV = 4.0e-4; t.BFCRV = -3.0;
    // IPHYS=0 never reads the swell-dissipation constants, keep them finite
    t.SWELLF4 = 1.5e05; t.SWELLF7 = 3.6e05; t.SWELLF7M1 = 1.0 / t.SWELLF7; t.Z0RAT = 0.04; t.Z0TUBMAX = 0.0005;
    t.SSDSC5 = 0.0;
  } else if (c.iphys == 1) {
    t.ZALP = 0.008; t.TAILFACTOR = 2.5; t.TAILFACTOR_PM = 3.0; t.RN1_RN = 0.25;
    if (c.llgcbz0) {
      t.ALPHA = 0.0055; t.ALPHAMIN = 0.0001; t.CHNKMIN_U = 28.; t.ALPHAPMAX = 0.03; t.DELTA_THETA_RN = 0.75;
      t.DTHRN_A = 0.60; t.DTHRN_U = 33.0; t.Z0TUBMAX = 0.05; t.Z0RAT = 0.02; t.SWELLF4 = 1.15e05; t.SWELLF7 = 4.32e05;
      t.SWELLF7M1 = 1.0 / t.SWELLF7; t.SSDSC5 = 0.0;
      if (c.llnormagam) { t.BETAMAX = 1.39; t.TAUWSHELTER = 0.0; } else { t.BETAMAX = 1.44; t.TAUWSHELTER = 0.25; }
    } else {
      t.ALPHA = 0.0065; t.ALPHAPMAX = 0.031; t.DELTA_THETA_RN = 0.75; t.DTHRN_A = 0.60; t.DTHRN_U = 200.0;
      t.Z0TUBMAX = 0.0005; t.Z0RAT = 0.04; t.SWELLF4 = 1.5e05; t.SWELLF7 = 3.6e05; t.SWELLF7M1 = 1.0 / t.SWELLF7;
      t.SSDSC5 = 0.0;
      if (c.llnormagam) { t.BETAMAX = 1.39; t.TAUWSHELTER = 0.0; t.ALPHAMIN = 0.0005; t.CHNKMIN_U = 30.; }
      else { t.BETAMAX = 1.40; t.TAUWSHELTER = 0.25; t.ALPHAMIN = 0.0001; t.CHNKMIN_U = 33.; }
    }
    t.EGRCRV = 1065.0; t.AFCRV = 2.453e-4; t.BFCRV = -3.1236;
    t.CDIS = -1.33; t.DELTA_SDIS = 0.5; t.CDISVIS = 0.0;  // unused for IPHYS=1
  } else {
    throw std::runtime_error("SETWAVPHYS: unknown IPHYS");
  }
  if (c.nang <= 24) { t.ANG_GC_A = 0.40; t.ANG_GC_B = 0.60; t.ANG_GC_C = 3.0; }   // setwavphys.F90:52-60, 107-115
  else { t.ANG_GC_A = 0.35; t.ANG_GC_B = 0.65; t.ANG_GC_C = 3.0; }
}

// real ** integer as compilers expand it (repeated squaring), not pow()
static double powi(double x, int m) {
  unsigned n = (unsigned)(m < 0 ? -m : m);
  double y = (n & 1) ? x : 1.0;
  while (n >>= 1) { x = x * x; if (n & 1) y = y * x; }
  return m < 0 ? 1.0 / y : y;
}
// initgc.F90:63-110 with gc_dispersion.h
static void initgc(Tables& t) {
  const double KRATIO_GC = 1.2, XKS_GC = 0.006, XKL_GC = 20000.0;   // yowfred.F90:62-65
  t.SURFT = 0.0717 / t.ROWATER;                                     // iniwcst.F90:69 (GAM_SURF, mpuserin.F90:828)
  t.SQRTGOSURFT = std::sqrt(t.G / t.SURFT);
  t.NWAV_GC = (int)nint(std::log(XKL_GC / XKS_GC) / std::log(KRATIO_GC));
  const int N = t.NWAV_GC;
  for (ArrD* a : {&t.XK_GC, &t.XKM_GC, &t.OMEGA_GC, &t.OMXKM3_GC, &t.VG_GC, &t.C_GC, &t.CM_GC, &t.C2OSQRTVG_GC, &t.XKMSQRTVGOC2_GC,
                  &t.OM3GMKM_GC, &t.DELKCC_GC, &t.DELKCC_GC_NS, &t.DELKCC_OMXKM3_GC}) a->alloc(1, N);
  auto FOMEG = [&](double k) { return std::sqrt(t.G * k + t.SURFT * (k * k * k)); };
  auto FVG = [&](double k) { return 0.5 / FOMEG(k) * (t.G + 3.0 * t.SURFT * (k * k)); };
  auto FC = [&](double k) { return FOMEG(k) / k; };
  for (int I = 1; I <= N; ++I) {
    t.XK_GC(I) = XKS_GC * powi(KRATIO_GC, I - 1);
    t.XKM_GC(I) = 1.0 / t.XK_GC(I);
    t.OMEGA_GC(I) = FOMEG(t.XK_GC(I));
    t.OMXKM3_GC(I) = t.OMEGA_GC(I) * (t.XKM_GC(I) * t.XKM_GC(I) * t.XKM_GC(I));
    t.VG_GC(I) = FVG(t.XK_GC(I));
    t.C_GC(I) = FC(t.XK_GC(I));
    t.CM_GC(I) = 1.0 / t.C_GC(I);
    t.C2OSQRTVG_GC(I) = (t.C_GC(I) * t.C_GC(I)) / std::sqrt(t.VG_GC(I));
    t.XKMSQRTVGOC2_GC(I) = t.XKM_GC(I) / t.C2OSQRTVG_GC(I);
    t.OM3GMKM_GC(I) = (t.OMEGA_GC(I) * t.OMEGA_GC(I) * t.OMEGA_GC(I)) / (t.G * t.XK_GC(I));
  }
  t.DELKCC_GC(1) = 0.5 * (t.XK_GC(2) - t.XK_GC(1)) / t.C2OSQRTVG_GC(1);
  t.DELKCC_GC_NS(1) = t.DELKCC_GC(1);
  for (int I = 2; I <= N - 1; ++I) {
    t.DELKCC_GC(I) = 0.5 * (t.XK_GC(I + 1) - t.XK_GC(I - 1)) / t.C2OSQRTVG_GC(I);
    t.DELKCC_GC_NS(I) = 0.5 * (t.XK_GC(I + 1) - t.XK_GC(I)) / t.C2OSQRTVG_GC(I);
  }
  t.DELKCC_GC(N) = 0.5 * (t.XK_GC(N) - t.XK_GC(N - 1)) / t.C2OSQRTVG_GC(N);
  t.DELKCC_GC_NS(N) = t.DELKCC_GC(N);
  for (int I = 1; I <= N; ++I) t.DELKCC_OMXKM3_GC(I) = t.DELKCC_GC(I) * t.OMXKM3_GC(I);
}

// mfr.F90 + mfredir.F90:90-129
static void mfredir(const Config& c, Tables& t) {
  const int NFRE = c.nfre, NANG = c.nang;
  t.FR.alloc(1, NFRE); t.DFIM.alloc(1, NFRE); t.GOM.alloc(1, NFRE); t.C.alloc(1, NFRE);
  t.TH.alloc(1, NANG); t.COSTH.alloc(1, NANG); t.SINTH.alloc(1, NANG);
  t.FR(c.ifre1) = c.fr1;
  for (int M = c.ifre1 - 1; M >= 1; --M) t.FR(M) = t.FR(M + 1) / t.FRATIO;
  for (int M = c.ifre1 + 1; M <= NFRE; ++M) t.FR(M) = t.FRATIO * t.FR(M - 1);
  for (int M = 1; M <= NFRE; ++M) { t.GOM(M) = t.G / (4.0 * t.PI * t.FR(M)); t.C(M) = t.G / (t.ZPI * t.FR(M)); }
  t.DELTH = t.ZPI / (double)NANG;
  for (int K = 1; K <= NANG; ++K) {
    t.TH(K) = (double)(K - 1) * t.DELTH + 0.5 * t.DELTH;
    t.COSTH(K) = std::cos(t.TH(K));
    t.SINTH(K) = std::sin(t.TH(K));
  }
  double CO1 = 0.5 * (t.FRATIO - 1.0) * t.DELTH;
  t.DFIM(1) = CO1 * t.FR(1);
  for (int M = 2; M <= NFRE - 1; ++M) t.DFIM(M) = CO1 * (t.FR(M) + t.FR(M - 1));
  t.DFIM(NFRE) = CO1 * t.FR(NFRE - 1);
}

// initmdl.F90:437-503
static void initmdl_freq(const Config& c, Tables& t) {
  const int NFRE = c.nfre;
  t.DFIMOFR.alloc(1, NFRE); t.DFIMFR.alloc(1, NFRE); t.DFIMFR2.alloc(1, NFRE); t.ZPIFR.alloc(1, NFRE);
  t.FR5.alloc(1, NFRE); t.FRM5.alloc(1, NFRE); t.COFRM4.alloc(1, NFRE); t.FLMAX.alloc(1, NFRE);
  t.RHOWG_DFIM.alloc(1, NFRE); t.DFIM_SIM.alloc(1, NFRE);
  for (int M = 1; M <= NFRE; ++M) {
    t.DFIMOFR(M) = t.DFIM(M) / t.FR(M);
    t.DFIMFR(M) = t.DFIM(M) * t.FR(M);
    t.DFIMFR2(M) = t.DFIM(M) * t.FR(M) * t.FR(M);
    t.ZPIFR(M) = t.ZPI * t.FR(M);
    t.FR5(M) = std::pow(t.FR(M), 5);
    t.FRM5(M) = 1.0 / t.FR5(M);
    t.COFRM4(M) = t.COEF4 * t.G / std::pow(t.FR(M), 4);
    t.FLMAX(M) = (t.ALPHAPMAX / t.PI) / (t.ZPI4GM2 * t.FR5(M));
  }
  t.FLOGSPRDM1 = 1.0 / std::log10(t.FRATIO);
  t.XLOGFRATIO = std::log(t.FRATIO);
  t.RHOWG_DFIM(1) = 0.5 * t.ROWATER * t.G * t.DELTH * t.XLOGFRATIO * t.FR(1);
  for (int M = 2; M <= NFRE - 1; ++M) t.RHOWG_DFIM(M) = t.ROWATER * t.G * t.DELTH * t.XLOGFRATIO * t.FR(M);
  t.RHOWG_DFIM(NFRE) = 0.5 * t.ROWATER * t.G * t.DELTH * t.XLOGFRATIO * t.FR(NFRE);
  t.NFRE_ODD = NFRE - 1 + (NFRE % 2);
  t.DFIM_SIM(NFRE) = 0.0;
  t.DFIM_SIM(1) = t.DELTH * t.XLOGFRATIO * t.FR(1) / 3.0;
  for (int M = 2; M <= t.NFRE_ODD - 1; M += 2) {
    t.DFIM_SIM(M) = 4.0 * t.DELTH * t.XLOGFRATIO * t.FR(M) / 3.0;
    t.DFIM_SIM(M + 1) = 2.0 * t.DELTH * t.XLOGFRATIO * t.FR(M + 1) / 3.0;
  }
  t.DFIM_SIM(t.NFRE_ODD) = t.DELTH * t.XLOGFRATIO * t.FR(t.NFRE_ODD) / 3.0;
}

// kzeone.F90 (ACM Algorithm 484, Burrell 1974): modified Bessel K0,K1 of complex argument times exp(x)
static void kzeone(double X, double Y, double& RE0, double& IM0, double& RE1, double& IM1) {
  static const double EXSQ[8] = {0.5641003087264E0, 0.4120286874989E0, 0.1584889157959E0, 0.3078003387255E-1,
                                 0.2778068842913E-2, 0.1000044412325E-3, 0.1059115547711E-5, 0.1522475804254E-8};
  static const double TSQ[8] = {0.0E0, 3.19303633920635E-1, 1.29075862295915E0, 2.95837445869665E0,
                                5.40903159724444E0, 8.80407957805676E0, 1.34685357432515E1, 2.02499163658709E1};
  double X2, Y2, R1, R2, T1, T2, P1, P2, RTERM, ITERM, L;
  R2 = X * X + Y * Y;
  if (R2 >= 1.96E2) {
    RTERM = 1.0; ITERM = 0.0; RE0 = 1.0; IM0 = 0.0; RE1 = 1.0; IM1 = 0.0;
    P1 = 8.0 * R2; P2 = std::sqrt(R2);
    L = 3.91 + 8.12E1 / P2;
    int LL = (int)nint(L);
    R1 = 1.0; R2 = 1.0;
    int M = -8, K = 3;
    for (int N = 1; N <= LL; ++N) {
      M = M + 8; K = K - M;
      R1 = (double)(K - 4) * R1; R2 = (double)K * R2;
      T1 = (double)N * P1; T2 = RTERM;
      RTERM = (T2 * X + ITERM * Y) / T1;
      ITERM = (-T2 * Y + ITERM * X) / T1;
      RE0 += R1 * RTERM; IM0 += R1 * ITERM; RE1 += R2 * RTERM; IM1 += R2 * ITERM;
    }
    T1 = std::sqrt(P2 + X); T2 = -Y / T1;
    P1 = 8.86226925452758E-1 / P2;
    RTERM = P1 * std::cos(Y); ITERM = -P1 * std::sin(Y);
    R1 = RE0 * RTERM - IM0 * ITERM; R2 = RE0 * ITERM + IM0 * RTERM;
    RE0 = T1 * R1 - T2 * R2; IM0 = T1 * R2 + T2 * R1;
    R1 = RE1 * RTERM - IM1 * ITERM; R2 = RE1 * ITERM + IM1 * RTERM;
    RE1 = T1 * R1 - T2 * R2; IM1 = T1 * R2 + T2 * R1;
    return;
  }
  if (R2 >= 1.849E1) {
    X2 = 2.0 * X; Y2 = 2.0 * Y; R1 = Y2 * Y2;
    P1 = std::sqrt(X2 * X2 + R1); P2 = std::sqrt(P1 + X2);
    T1 = EXSQ[0] / (2.0 * P1);
    RE0 = T1 * P2; IM0 = T1 / P2; RE1 = 0.0; IM1 = 0.0;
    for (int N = 2; N <= 8; ++N) {
      T2 = X2 + TSQ[N - 1];
      P1 = std::sqrt(T2 * T2 + R1); P2 = std::sqrt(P1 + T2);
      T1 = EXSQ[N - 1] / P1;
      RE0 += T1 * P2; IM0 += T1 / P2;
      T1 = EXSQ[N - 1] * TSQ[N - 1];
      RE1 += T1 * P2; IM1 += T1 / P2;
    }
    T2 = -Y2 * IM0;
    RE1 = RE1 / R2;
    R2 = Y2 * IM1 / R2;
    RTERM = 1.41421356237309E0 * std::cos(Y);
    ITERM = -1.41421356237309E0 * std::sin(Y);
    IM0 = RE0 * ITERM + T2 * RTERM;
    RE0 = RE0 * RTERM - T2 * ITERM;
    T1 = RE1 * RTERM - R2 * ITERM;
    T2 = RE1 * ITERM + R2 * RTERM;
    RE1 = T1 * X + T2 * Y;
    IM1 = -T1 * Y + T2 * X;
    return;
  }
  X2 = X / 2.0; Y2 = Y / 2.0;
  P1 = X2 * X2; P2 = Y2 * Y2;
  T1 = -(std::log(P1 + P2) / 2.0 + 0.5772156649015329E0);
  T2 = -std::atan2(Y, X);
  X2 = P1 - P2; Y2 = X * Y2;
  RTERM = 1.0; ITERM = 0.0;
  RE0 = T1; IM0 = T2;
  T1 = T1 + 0.5;
  RE1 = T1; IM1 = T2;
  P2 = std::sqrt(R2);
  L = 2.106 * P2 + 4.4;
  if (P2 < 8.0E-1) L = 2.129 * P2 + 4.0;
  int LL = (int)nint(L);
  for (int N = 1; N <= LL; ++N) {
    P1 = N; P2 = (double)N * N;
    R1 = RTERM;
    RTERM = (R1 * X2 - ITERM * Y2) / P2;
    ITERM = (R1 * Y2 + ITERM * X2) / P2;
    T1 = T1 + 0.5 / P1;
    RE0 = RE0 + T1 * RTERM - T2 * ITERM;
    IM0 = IM0 + T1 * ITERM + T2 * RTERM;
    P1 = P1 + 1.0;
    T1 = T1 + 0.5 / P1;
    RE1 = RE1 + (T1 * RTERM - T2 * ITERM) / P1;
    IM1 = IM1 + (T1 * ITERM + T2 * RTERM) / P1;
  }
  R1 = X / R2 - 0.5 * (X * RE1 - Y * IM1);
  R2 = -Y / R2 - 0.5 * (X * IM1 + Y * RE1);
  P1 = std::exp(X);
  RE0 = P1 * RE0; IM0 = P1 * IM0; RE1 = P1 * R1; IM1 = P1 * R2;
}

// kerkei.F90
static void kerkei(double X, double& KER, double& KEI) {
  double ZR = X * 0.50 * std::sqrt(2.0), ZI = ZR, CYR, CYI, CYR1, CYI1;
  kzeone(ZR, ZI, CYR, CYI, CYR1, CYI1);
  KER = CYR / std::exp(ZR);
  KEI = CYI / std::exp(ZR);
}

// tabu_swellft.F90:64-83
static void tabu_swellft(Tables& t) {
  const int NITER = 100;
  const double ABMIN = 0.3, ABMAX = 8.0, KAPPA = 0.40;
  t.SWELLFT.alloc(1, t.IAB);
  double DZETA0 = 0.0;
  double DELAB = (ABMAX - ABMIN) / (double)t.IAB;
  double L10 = std::log(10.0);
  for (int I = 1; I <= t.IAB; ++I) {
    double ABRLOG = ABMIN + (double)I * DELAB;
    double ABR = std::exp(ABRLOG * L10);
    double FACT = 1 / ABR / (21.2 * KAPPA);
    double FSUBW = 0.05;
    for (int ITER = 1; ITER <= NITER; ++ITER) {
      double FSUBWMEMO = FSUBW, DZETA0MEMO = DZETA0, KER, KEI;
      DZETA0 = FACT * std::pow(FSUBW, -0.5);
      kerkei(2.0 * std::sqrt(DZETA0), KER, KEI);
      FSUBW = 0.08 / (KER * KER + KEI * KEI);
      FSUBW = 0.5 * (FSUBWMEMO + FSUBW);
      DZETA0 = 0.5 * (DZETA0MEMO + DZETA0);
    }
    t.SWELLFT(I) = FSUBW;
  }
}

// init_x0tauhf.F90:65-100
static void init_x0tauhf(const Config& c, Tables& t) {
  t.BETAMAXOXKAPPA2 = t.BETAMAX / (t.XKAPPA * t.XKAPPA);
  t.BMAXOKAP = t.DELTA_THETA_RN * t.BETAMAXOXKAPPA2 / t.XKAPPA;
  t.BMAXOKAPDTH = t.BMAXOKAP * t.DELTH;
  t.GAMNCONST = t.BMAXOKAP * 0.5 * std::pow(t.ZPI, 4) * std::pow(t.GM1, 3);
  double ALPH = (c.llgcbz0 || c.llcapchnk || c.llnormagam) ? t.ALPHAMIN : t.ALPHA;
  double X0 = 0.005;
  for (int J = 1; J <= 30; ++J) {
    double FF = std::exp(t.XKAPPA / (X0 + t.ZALP));
    double F = ALPH * X0 * X0 * FF - 1.0;
    if (F == 0.0) break;
    double x = X0 / (X0 + t.ZALP);
    double DF = ALPH * FF * (2.0 * X0 - t.XKAPPA * (x * x));
    X0 = X0 - F / DF;
  }
  t.X0TAUHF = X0;
  t.WTAUHF.alloc(1, t.JTOT_TAUHF);
  double CONST1 = t.BETAMAXOXKAPPA2 / 3.0;
  t.WTAUHF(1) = CONST1;
  for (int J = 2; J <= t.JTOT_TAUHF - 1; J += 2) { t.WTAUHF(J) = 4.0 * CONST1; t.WTAUHF(J + 1) = 2.0 * CONST1; }
  t.WTAUHF(t.JTOT_TAUHF) = CONST1;
}

// init_sdiss_ardh.F90:69-96
static void init_sdiss_ardh(const Config& c, Tables& t) {
  const int NANG = c.nang;
  int NANGD = NANG / 2;
  t.NSDSNTH = (int)std::min<long>(nint(t.ISDSDTH * t.RAD / (t.DELTH)), NANGD - 1);
  double DELTH_TRUNC = (t.TH(1) + t.ISDSDTH * t.RAD) - (t.TH(1 + t.NSDSNTH) - 0.5 * t.DELTH);
  DELTH_TRUNC = std::max(0.0, std::min(DELTH_TRUNC, t.DELTH));
  t.INDICESSAT.alloc(1, NANG, 1, t.NSDSNTH * 2 + 1);
  t.SATWEIGHTS.alloc(1, NANG, 1, t.NSDSNTH * 2 + 1);
  for (int K = 1; K <= NANG; ++K) {
    for (int I_INT = K - t.NSDSNTH; I_INT <= K + t.NSDSNTH; ++I_INT) {
      int J_INT = I_INT;
      if (I_INT < 1) J_INT = I_INT + NANG;
      if (I_INT > NANG) J_INT = I_INT - NANG;
      t.INDICESSAT(K, I_INT - (K - t.NSDSNTH) + 1) = J_INT;
      double DELTH_LOC = (I_INT == K - t.NSDSNTH || I_INT == K + t.NSDSNTH) ? DELTH_TRUNC : t.DELTH;
      double cs = std::cos(t.TH(K) - t.TH(J_INT));
      t.SATWEIGHTS(K, I_INT - (K - t.NSDSNTH) + 1) = DELTH_LOC * (cs * cs);  // **ISB, ISB=2
    }
  }
}

// jafu.F90
static int jafu(double CL, int J, int IAN) {
  int IDPH = (int)CL;  // Fortran real->integer assignment truncates toward zero
  int JA = J + IDPH;
  if (JA <= 0) JA = IAN + JA - 1;
  if (JA >= IAN) JA = JA - IAN + 1;
  return JA;
}

// nlweigt.F90:94-262
static void nlweigt(const Config& c, Tables& t) {
  const int NANG = c.nang, NFRE = c.nfre;
  const double ALAMD = 0.25, CON = 3000.0;
  double F1P1 = std::log10(t.FRATIO);
  int ISP = (int)(std::log10(1.0 + ALAMD) / F1P1 + .000001);
  int ISM = (int)std::floor(std::log10(1.0 - ALAMD) / F1P1 + .0000001);
  t.MFRSTLW = 1 + ISM;
  t.MLSTHG = NFRE - ISM;
  t.KFRH = -ISM + ISP + 2;
  ArrI JA1, JA2;
  JA1.alloc(1, NANG, 1, 2); JA2.alloc(1, NANG, 1, 2);
  ArrD FRLON; FRLON.alloc(t.MFRSTLW, NFRE + t.KFRH);
  t.IKP.alloc(t.MFRSTLW, t.MLSTHG); t.IKP1.alloc(t.MFRSTLW, t.MLSTHG);
  t.IKM.alloc(t.MFRSTLW, t.MLSTHG); t.IKM1.alloc(t.MFRSTLW, t.MLSTHG);
  t.K1W.alloc(1, NANG, 1, 2); t.K2W.alloc(1, NANG, 1, 2); t.K11W.alloc(1, NANG, 1, 2); t.K21W.alloc(1, NANG, 1, 2);
  t.AF11.alloc(t.MFRSTLW, t.MLSTHG); t.FKLAP.alloc(t.MFRSTLW, t.MLSTHG); t.FKLAP1.alloc(t.MFRSTLW, t.MLSTHG);
  t.FKLAM.alloc(t.MFRSTLW, t.MLSTHG); t.FKLAM1.alloc(t.MFRSTLW, t.MLSTHG);
  t.FRH.alloc(1, t.KFRH);

  double XF = std::pow((1.0 + ALAMD) / (1.0 - ALAMD), 4);
  double COSTH3 = (1.0 + 2.0 * ALAMD + 2.0 * ALAMD * ALAMD * ALAMD) / ((1.0 + ALAMD) * (1.0 + ALAMD));
  double DELPHI1 = -180.0 / t.PI * std::acos(COSTH3);
  double COSTH4 = std::sqrt(1.0 - XF + XF * COSTH3 * COSTH3);
  double DELPHI2 = 180.0 / t.PI * std::acos(COSTH4);
  double DELTHA = t.DELTH * t.DEG;
  double CL1 = DELPHI1 / DELTHA, CL2 = DELPHI2 / DELTHA;

  int KLP1 = NANG + 1;
  int IC = 1;
  for (int KH = 1; KH <= 2; ++KH) {
    int KLH = NANG;
    if (KH == 2) KLH = KLP1;
    for (int K = 1; K <= KLH; ++K) {
      int KS = K;
      if (KH > 1) KS = KLP1 - K + 1;
      if (KS > NANG) continue;
      double CH = IC * CL1;
      JA1(KS, KH) = jafu(CH, K, KLP1);
      CH = IC * CL2;
      JA2(KS, KH) = jafu(CH, K, KLP1);
    }
    IC = -1;
  }
  CL1 = CL1 - (int)CL1;
  CL2 = CL2 - (int)CL2;
  t.ACL1 = std::fabs(CL1); t.ACL2 = std::fabs(CL2);
  t.CL11 = 1.0 - t.ACL1; t.CL21 = 1.0 - t.ACL2;
  double AL11 = std::pow(1.0 + ALAMD, 4), AL12 = std::pow(1.0 - ALAMD, 4);
  t.DAL1 = 1.0 / AL11; t.DAL2 = 1.0 / AL12;

  int ISG = 1;
  for (int KH = 1; KH <= 2; ++KH) {
    double CL1H = ISG * CL1, CL2H = ISG * CL2;
    for (int K = 1; K <= NANG; ++K) {
      int KS = K;
      if (KH == 2) KS = NANG - K + 2;
      if (K == 1) KS = 1;
      int K1 = JA1(K, KH);
      t.K1W(KS, KH) = K1;
      int K11;
      if (CL1H < 0.0) { K11 = K1 - 1; if (K11 < 1) K11 = NANG; } else { K11 = K1 + 1; if (K11 > NANG) K11 = 1; }
      t.K11W(KS, KH) = K11;
      int K2 = JA2(K, KH);
      t.K2W(KS, KH) = K2;
      int K21;
      if (CL2H < 0) { K21 = K2 - 1; if (K21 < 1) K21 = NANG; } else { K21 = K2 + 1; if (K21 > NANG) K21 = 1; }
      t.K21W(KS, KH) = K21;
    }
    ISG = -1;
  }

  for (int M = 1; M <= NFRE; ++M) FRLON(M) = t.FR(M);
  for (int M = 0; M >= t.MFRSTLW; --M) FRLON(M) = FRLON(M + 1) / t.FRATIO;
  for (int M = NFRE + 1; M <= NFRE + t.KFRH; ++M) FRLON(M) = t.FRATIO * FRLON(M - 1);

  for (int M = t.MFRSTLW; M <= t.MLSTHG; ++M) {
    double FRG = FRLON(M);
    t.AF11(M) = CON * std::pow(FRG, 11);
    double FLP = FRG * (1.0 + ALAMD), FLM = FRG * (1.0 - ALAMD);
    int IKN = M + ISP;
    t.IKP(M) = IKN;
    double FKP = FRLON(t.IKP(M));
    t.IKP1(M) = t.IKP(M) + 1;
    t.FKLAP(M) = (FLP - FKP) / (FRLON(t.IKP1(M)) - FKP);
    t.FKLAP1(M) = 1.0 - t.FKLAP(M);
    IKN = M + ISM;
    if (IKN >= t.MFRSTLW) {
      t.IKM(M) = IKN;
      double FKM = FRLON(t.IKM(M));
      t.IKM1(M) = t.IKM(M) + 1;
      t.FKLAM(M) = (FLM - FKM) / (FRLON(t.IKM1(M)) - FKM);
      t.FKLAM1(M) = 1.0 - t.FKLAM(M);
    } else if (IKN + 1 == t.MFRSTLW) {
      t.IKM(M) = 1;
      t.IKM1(M) = t.MFRSTLW;
      double FKM = FRLON(t.IKM1(M)) / t.FRATIO;
      t.FKLAM(M) = (FLM - FKM) / (FRLON(t.IKM1(M)) - FKM);
      t.FKLAM1(M) = 0.0;
    } else {
      t.IKM(M) = 1; t.FKLAM(M) = 0.0; t.IKM1(M) = 1; t.FKLAM1(M) = 0.0;
    }
  }
  for (int I = 1; I <= t.KFRH; ++I) {
    int M = NFRE + I - 1;
    t.FRH(I) = std::pow(FRLON(NFRE) / FRLON(M), 5);
  }
}

// inisnonlin.F90:89-270
static void inisnonlin(const Config& c, Tables& t) {
  const int NFRE = c.nfre;
  nlweigt(c, t);
  auto EPMMA = [](double X) { return std::exp(-std::min(1.25 * std::pow(X, 4), 50.0)) * std::pow(X, 5); };
  t.FTRF.alloc(t.MFRSTLW, 1);
  double ALPH = 1.0 / EPMMA(1.0);
  double FRR = 1.0;
  for (int MC = 1; MC >= t.MFRSTLW; --MC) { t.FTRF(MC) = ALPH * EPMMA(FRR); FRR = FRR * t.FRATIO; }
  t.INLCOEF.alloc(1, 5, 1, t.MLSTHG);
  t.RNLCOEF.alloc(1, 25, 1, t.MLSTHG);
  for (int MC = 1; MC <= t.MLSTHG; ++MC) {
    int MP = t.IKP(MC), MP1 = t.IKP1(MC), MM = t.IKM(MC), MM1 = t.IKM1(MC);
    double FFACP = 1.0, FFACP1 = 1.0, FFACM = 1.0, FFACM1 = 1.0, FTAIL = 1.0;
    int IC = MC;
    if (IC < 1) IC = 1;
    int IP = MP, IP1 = MP1, IM = MM, IM1 = MM1;
    if (IP < 1) { FFACP = t.FTRF(IP); IP = 1; }
    if (IP1 < 1) { FFACP1 = t.FTRF(IP1); IP1 = 1; }
    if (IM < t.MFRSTLW) { FFACM = 0.0; IM = 1; } else if (IM < 1) { FFACM = t.FTRF(IM); IM = 1; }
    if (IM1 < t.MFRSTLW) { FFACM1 = 0.0; IM1 = 1; } else if (IM1 < 1) { FFACM1 = t.FTRF(IM1); IM1 = 1; }
    if (IP1 > NFRE) {
      int ITEMP = IP1 - NFRE + 1;
      if (ITEMP > t.KFRH) ITEMP = t.KFRH;
      FFACP1 = t.FRH(ITEMP);
      IP1 = NFRE;
      if (IP > NFRE) {
        FFACP = t.FRH(IP - NFRE + 1);
        IP = NFRE;
        if (IC > NFRE) {
          FTAIL = t.FRH(IC - NFRE + 1);
          IC = NFRE;
          if (IM1 > NFRE) { FFACM1 = t.FRH(IM1 - NFRE + 1); IM1 = NFRE; }
        }
      }
    }
    t.INLCOEF(1, MC) = IC; t.INLCOEF(2, MC) = IP; t.INLCOEF(3, MC) = IP1; t.INLCOEF(4, MC) = IM;
    t.INLCOEF(5, MC) = IM1;
    double FKLAMP = t.FKLAP(MC), FKLAMP1 = t.FKLAP1(MC);
    double GW2 = FKLAMP1 * FFACP * t.DAL1;
    double GW1 = GW2 * t.CL11;
    GW2 = GW2 * t.ACL1;
    double GW4 = FKLAMP * FFACP1 * t.DAL1;
    double GW3 = GW4 * t.CL11;
    GW4 = GW4 * t.ACL1;
    double FKLAMPA = FKLAMP * t.CL11, FKLAMPB = FKLAMP * t.ACL1;
    double FKLAMP2 = FKLAMP1 * t.ACL1;
    FKLAMP1 = FKLAMP1 * t.CL11;
    double FKLAPA2 = FKLAMPA * FKLAMPA, FKLAPB2 = FKLAMPB * FKLAMPB, FKLAP12 = FKLAMP1 * FKLAMP1,
           FKLAP22 = FKLAMP2 * FKLAMP2;
    int r = 0;
    t.RNLCOEF(++r, MC) = FTAIL; t.RNLCOEF(++r, MC) = GW1; t.RNLCOEF(++r, MC) = GW2; t.RNLCOEF(++r, MC) = GW3;
    t.RNLCOEF(++r, MC) = GW4; t.RNLCOEF(++r, MC) = FKLAMPA; t.RNLCOEF(++r, MC) = FKLAMPB;
    t.RNLCOEF(++r, MC) = FKLAMP2; t.RNLCOEF(++r, MC) = FKLAMP1; t.RNLCOEF(++r, MC) = FKLAPA2;
    t.RNLCOEF(++r, MC) = FKLAPB2; t.RNLCOEF(++r, MC) = FKLAP12; t.RNLCOEF(++r, MC) = FKLAP22;
    double FKLAMM = t.FKLAM(MC), FKLAMM1 = t.FKLAM1(MC);
    double GW6 = FKLAMM1 * FFACM * t.DAL2;
    double GW5 = GW6 * t.CL21;
    GW6 = GW6 * t.ACL2;
    double GW8 = FKLAMM * FFACM1 * t.DAL2;
    double GW7 = GW8 * t.CL21;
    GW8 = GW8 * t.ACL2;
    double FKLAMMA = FKLAMM * t.CL21, FKLAMMB = FKLAMM * t.ACL2;
    double FKLAMM2 = FKLAMM1 * t.ACL2;
    FKLAMM1 = FKLAMM1 * t.CL21;
    double FKLAMA2 = FKLAMMA * FKLAMMA, FKLAMB2 = FKLAMMB * FKLAMMB, FKLAM12 = FKLAMM1 * FKLAMM1,
           FKLAM22 = FKLAMM2 * FKLAMM2;
    t.RNLCOEF(++r, MC) = GW5; t.RNLCOEF(++r, MC) = GW6; t.RNLCOEF(++r, MC) = GW7; t.RNLCOEF(++r, MC) = GW8;
    t.RNLCOEF(++r, MC) = FKLAMMA; t.RNLCOEF(++r, MC) = FKLAMMB; t.RNLCOEF(++r, MC) = FKLAMM2;
    t.RNLCOEF(++r, MC) = FKLAMM1; t.RNLCOEF(++r, MC) = FKLAMA2; t.RNLCOEF(++r, MC) = FKLAMB2;
    t.RNLCOEF(++r, MC) = FKLAM12; t.RNLCOEF(++r, MC) = FKLAM22;
  }
}

// cigetdeac.F90:60-75, :77-82 (assumed 1 s column), :end-10 (linear extrapolation of the 2..5 s columns)
static void cigetdeac(Tables& t) {
  static const double KM[36][11] = {
#include "../ecwam_b200/csrc/kohout_meylan_fig6.inc"
  };
  t.NICH = 36; t.DHIC = 0.1;
  t.NICT = 16; t.TICMIN = 1.0; t.DTIC = 1.0;
  t.CIDEAC.alloc(1, t.NICT, 1, t.NICH);
  t.CIDEAC(1, 1) = -2.00;
  t.CIDEAC(1, t.NICH) = -1.00;
  const double DHI = t.CIDEAC(1, t.NICH) - t.CIDEAC(1, 1);
  for (int IH = 2; IH <= t.NICH - 1; ++IH) t.CIDEAC(1, IH) = t.CIDEAC(1, 1) + (IH - 1) * DHI / (t.NICH - 1);
  for (int IH = 1; IH <= t.NICH; ++IH)
    for (int IT = 6; IT <= 16; ++IT) t.CIDEAC(IT, IH) = KM[IH - 1][IT - 6];
  for (int IH = 1; IH <= t.NICH; ++IH) {
    const double DCI = t.CIDEAC(6, IH) - t.CIDEAC(1, IH);
    for (int IT = 2; IT <= 5; ++IT) t.CIDEAC(IT, IH) = t.CIDEAC(1, IH) + DCI * (IT - 1) * t.DTIC / (5 * t.DTIC);
  }
}

void init_tables(const Config& c, Tables& t) {
  iniwcst(t);
  cigetdeac(t);
  setwavphys(c, t);
  mfredir(c, t);
  initmdl_freq(c, t);
  tabu_swellft(t);
  init_x0tauhf(c, t);
  initgc(t);
  if (c.iphys == 1) init_sdiss_ardh(c, t);
  inisnonlin(c, t);
}

}  // namespace orc
