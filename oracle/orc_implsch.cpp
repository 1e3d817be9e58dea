// ORACLE (test infrastructure) — IMPLSCH and its call tree, restated loop by loop from the reference.
// Branches exercised by the BASELINE configs only: LLGCBZ0=F, LLNORMAGAM=F, ISNONLIN=0,
// LWNEMOCOU*=F, ICODE_WND=3; IPHYS=0 and 1; LCIWA1-3 (SDICE).  Other off-by-default branches throw.
#include "oracle.h"
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

namespace {
// 1-based views on column-major storage
struct V1 { double* p; inline double& operator()(int i) const { return p[i - 1]; } };
struct V2 { double* p; int n0; inline double& operator()(int i, int j) const { return p[(i - 1) + (size_t)n0 * (j - 1)]; } };
struct V3 { double* p; int n0, n1; inline double& operator()(int i, int j, int k) const { return p[(i - 1) + (size_t)n0 * ((j - 1) + (size_t)n1 * (k - 1))]; } };
struct I1 { int* p; inline int& operator()(int i) const { return p[i - 1]; } };
struct L1 { std::vector<double> v; V1 view() { return V1{v.data()}; } L1(int n) : v(n, 0.0) {} };
inline double sq(double x) { return x * x; }
inline double p4(double x) { double y = x * x; return y * y; }

struct Ctx {
  const Config& c; const Tables& t; int KIJS, KIJL, NANG, NFRE;
};

// semean.F90:60-124
void SEMEAN(const Ctx& x, V3 FL1, V1 EM, bool LLEPSMIN) {
  const Tables& t = x.t;
  std::vector<double> TEMP(x.KIJL + 1);
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) EM(IJ) = LLEPSMIN ? t.EPSMIN : 0.0;
  for (int M = 1; M <= x.NFRE; ++M) {
    for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) TEMP[IJ] = FL1(IJ, 1, M);
    for (int K = 2; K <= x.NANG; ++K) for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) TEMP[IJ] = TEMP[IJ] + FL1(IJ, K, M);
    for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) EM(IJ) = EM(IJ) + t.DFIM(M) * TEMP[IJ];
  }
  double DELT25 = t.WETAIL * t.FR(x.NFRE) * t.DELTH;
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) EM(IJ) = EM(IJ) + DELT25 * TEMP[IJ];
}

// sdepthlim.F90:50-82
void SDEPTHLIM(const Ctx& x, V1 EMAXDPT, V3 FL1) {
  L1 EMs(x.KIJL); V1 EM = EMs.view();
  SEMEAN(x, FL1, EM, true);
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) EM(IJ) = std::min(EMAXDPT(IJ) / EM(IJ), 1.0);
  for (int M = 1; M <= x.NFRE; ++M) for (int K = 1; K <= x.NANG; ++K) for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ)
    FL1(IJ, K, M) = std::max(FL1(IJ, K, M) * EM(IJ), x.t.EPSMIN);
}

// fkmean.F90:60-154
void FKMEAN(const Ctx& x, V3 FL1, V2 WAVNUM, V1 EM, V1 FM1, V1 F1, V1 AK, V1 XK) {
  const Tables& t = x.t;
  std::vector<double> TEMPA(x.KIJL + 1), TEMPX(x.KIJL + 1), TEMP2(x.KIJL + 1);
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) { EM(IJ) = t.EPSMIN; FM1(IJ) = t.EPSMIN; F1(IJ) = t.EPSMIN; AK(IJ) = t.EPSMIN; XK(IJ) = t.EPSMIN; }
  double DELT25 = t.WETAIL * t.FR(x.NFRE) * t.DELTH;
  double COEFM1 = t.FRTAIL * t.DELTH;
  double COEF1 = t.WP1TAIL * t.DELTH * sq(t.FR(x.NFRE));
  double COEFA = COEFM1 * std::sqrt(t.G) / t.ZPI;
  double COEFX = COEF1 * (t.ZPI / std::sqrt(t.G));
  for (int M = 1; M <= x.NFRE; ++M) {
    for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
      double SQRTK = std::sqrt(WAVNUM(IJ, M));
      TEMPA[IJ] = t.DFIM(M) / SQRTK;
      TEMPX[IJ] = SQRTK * t.DFIM(M);
    }
    for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) TEMP2[IJ] = FL1(IJ, 1, M);
    for (int K = 2; K <= x.NANG; ++K) for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) TEMP2[IJ] = TEMP2[IJ] + FL1(IJ, K, M);
    for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
      EM(IJ) = EM(IJ) + t.DFIM(M) * TEMP2[IJ];
      FM1(IJ) = FM1(IJ) + t.DFIMOFR(M) * TEMP2[IJ];
      F1(IJ) = F1(IJ) + t.DFIMFR(M) * TEMP2[IJ];
      AK(IJ) = AK(IJ) + TEMPA[IJ] * TEMP2[IJ];
      XK(IJ) = XK(IJ) + TEMPX[IJ] * TEMP2[IJ];
    }
  }
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
    EM(IJ) = EM(IJ) + DELT25 * TEMP2[IJ];
    FM1(IJ) = FM1(IJ) + COEFM1 * TEMP2[IJ];
    FM1(IJ) = EM(IJ) / FM1(IJ);
    F1(IJ) = F1(IJ) + COEF1 * TEMP2[IJ];
    F1(IJ) = F1(IJ) / EM(IJ);
    AK(IJ) = AK(IJ) + COEFA * TEMP2[IJ];
    AK(IJ) = sq(EM(IJ) / AK(IJ));
    XK(IJ) = XK(IJ) + COEFX * TEMP2[IJ];
    XK(IJ) = sq(XK(IJ) / EM(IJ));
  }
}

// femeanws.F90:50-127
void FEMEANWS(const Ctx& x, V3 FL1, V3 XLLWS, V1 FM, double* EMout) {
  const Tables& t = x.t;
  std::vector<double> TEMP2(x.KIJL + 1), EM_LOC(x.KIJL + 1);
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) { EM_LOC[IJ] = t.EPSMIN; FM(IJ) = t.EPSMIN; }
  double DELT25 = t.WETAIL * t.FR(x.NFRE) * t.DELTH;
  double DELT2 = t.FRTAIL * t.DELTH;
  for (int M = 1; M <= x.NFRE; ++M) {
    for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) TEMP2[IJ] = 0.0;
    for (int K = 1; K <= x.NANG; ++K) for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) TEMP2[IJ] = TEMP2[IJ] + XLLWS(IJ, K, M) * FL1(IJ, K, M);
    for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) { EM_LOC[IJ] = EM_LOC[IJ] + t.DFIM(M) * TEMP2[IJ]; FM(IJ) = FM(IJ) + t.DFIMOFR(M) * TEMP2[IJ]; }
  }
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
    EM_LOC[IJ] = EM_LOC[IJ] + DELT25 * TEMP2[IJ];
    FM(IJ) = FM(IJ) + DELT2 * TEMP2[IJ];
    FM(IJ) = EM_LOC[IJ] / FM(IJ);
  }
  if (EMout) for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) EMout[IJ - 1] = EM_LOC[IJ];
}

// frcutindex.F90:84-108
void FRCUTINDEX(const Ctx& x, V1 FM, V1 FMWS, V1 UFRIC, V1 CICOVER, I1 MIJ, V2 RHOWGDFTH) {
  const Tables& t = x.t;
  const double FRIC = 28.0;  // yowfred.F90
  double FPMH = t.TAILFACTOR / t.FR(1);
  double FPPM = t.TAILFACTOR_PM * t.G / (FRIC * t.ZPIFR(1));
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
    if (CICOVER(IJ) <= x.c.cithrsh_tail) {
      double FM2 = std::max(FMWS(IJ), FM(IJ)) * FPMH;
      double FPM = FPPM / std::max(UFRIC(IJ), t.EPSMIN);
      double FPM4 = std::max(FM2, FPM);
      MIJ(IJ) = (int)nint(std::log10(FPM4) * t.FLOGSPRDM1) + 1;
      MIJ(IJ) = std::min(std::max(1, MIJ(IJ)), x.NFRE);
    } else {
      MIJ(IJ) = x.NFRE;
    }
  }
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
    for (int M = 1; M <= MIJ(IJ); ++M) RHOWGDFTH(IJ, M) = t.RHOWG_DFIM(M);
    if (MIJ(IJ) != x.NFRE) RHOWGDFTH(IJ, MIJ(IJ)) = 0.5 * RHOWGDFTH(IJ, MIJ(IJ));
    for (int M = MIJ(IJ) + 1; M <= x.NFRE; ++M) RHOWGDFTH(IJ, M) = 0.0;
  }
}

// chnkmin.F90
double CHNKMIN(const Tables& t, double U10) { return t.ALPHAMIN + (t.ALPHA - t.ALPHAMIN) * 0.5 * (1.0 - std::tanh(U10 - t.CHNKMIN_U)); }

// ns_gc.F90:44-48
static int NS_GC(const Tables& t, double USTAR) {
  const double XLOGKRATIOM1_GC = 1.0 / std::log(1.2);   // yowfred.F90:62-63
  const double XKS = t.SQRTGOSURFT / (1.48 + 2.05 * USTAR);
  return std::min((int)(std::log(std::max(XKS * t.XKM_GC(1), 1.0)) * XLOGKRATIOM1_GC) + 1, t.NWAV_GC - 1);
}
// stress_gc.F90:70-130: wave-induced stress of the gravity-capillary waves (part of the unresolved spectrum)
static double STRESS_GC(const Ctx& x, double ANG_GC, double USTAR, double Z0, double Z0MIN, double HALP, double RNFAC) {
  const Tables& t = x.t;
  const double XLAMA = 0.25, XLAMB = 4.0;
  const int NS = NS_GC(t, USTAR);
  const double TAUWCG_MIN = sq(USTAR * (Z0MIN / Z0));
  const double XLAMBDA = 1.0 + XLAMA * std::tanh(XLAMB * p4(USTAR));
  const double ZABHRC = ANG_GC * t.BETAMAXOXKAPPA2 * HALP * t.C2OSQRTVG_GC(NS);
  const double CONST = x.c.llnormagam ? RNFAC * t.BMAXOKAP * HALP * t.C2OSQRTVG_GC(NS) / std::max(USTAR, t.EPSUS) : 0.0;
  double TAUWCG = 0.0;
  for (int I = NS; I <= t.NWAV_GC; ++I) {
    const double X = USTAR * t.CM_GC(I);
    const double XLOG = std::log(t.XK_GC(I) * Z0) + t.XKAPPA / (X + t.ZALP);
    double ZLOG = XLOG - std::log(XLAMBDA);
    ZLOG = std::min(ZLOG, 0.0);
    const double ZLOG2X = ZLOG * ZLOG * X;
    const double GAM_W = ZLOG2X * ZLOG2X * std::exp(XLOG) * t.OM3GMKM_GC(I);
    const double ZN = CONST * t.XKMSQRTVGOC2_GC(I) * GAM_W;
    const double GAMNORMA = (1.0 + t.RN1_RN * ZN) / (1.0 + ZN);
    if (I == NS) TAUWCG = GAM_W * t.DELKCC_GC_NS(NS) * t.OMXKM3_GC(NS) * GAMNORMA;
    else TAUWCG = TAUWCG + GAM_W * t.DELKCC_OMXKM3_GC(I) * GAMNORMA;
  }
  return std::max(ZABHRC * TAUWCG, TAUWCG_MIN);
}
// cdm.func.h
static double CDM(double U) { return std::max(std::min(0.0006 + 0.00008 * U, 0.001 + 0.0018 * std::exp(-0.05 * (U - 33.))), 0.001); }

// taut_z0.F90:120-341 via airsea.F90 ICODE_WND == 3: the gravity-capillary model (LLGCBZ0, :148-279) or the Charnock/Janssen
// relation (:281-341)
void TAUT_Z0(const Ctx& x, int IUSFG, V1 HALP, V1 UTOP, V1 UDIR, V1 TAUW, V1 TAUWDIR, V1 RNFAC, V1 USTAR, V1 Z0, V1 Z0B, V1 CHRNCK) {
  if (x.c.llgcbz0) {
    const Tables& t = x.t;
    const Config& c = x.c;
    const int NITER = 18;
    const double PMAX = 0.99, Z0MIN = 0.000001;   // taut_z0.F90:93-99
    const double US2TOTAUW = 1.0 + t.EPS1;
    const double RNUEFF = 0.04 * c.rnu, RNUKAPPAM1 = RNUEFF / t.XKAPPA;
    const double PCE_GC = 0.001 * IUSFG + (1 - IUSFG) * 0.005;
    const double ACDLIN = t.ACDLIN, BCDLIN = t.BCDLIN;
    for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
      const double COSDIFF = std::cos(UDIR(IJ) - TAUWDIR(IJ));
      const double TAUWACT = std::max(TAUW(IJ) * COSDIFF, t.EPSMIN);
      const bool LLCOSDIFF = COSDIFF > 0.9;
      const double ALPHAOG = c.llcapchnk ? CHNKMIN(t, UTOP(IJ)) * t.GM1 : 0.0;
      const double u = UTOP(IJ);
      const double USMAX = std::max(-0.21339 + 0.093698 * u - 0.0020944 * (u * u) + 5.5091E-5 * (u * u * u), 0.03);
      const double TAUWEFF = std::min(TAUWACT * US2TOTAUW, USMAX * USMAX);
      if (IUSFG == 0) {
        const double ALPHAGM1 = t.ALPHA * t.GM1;
        double CDFG;
        if (u < 1.0) CDFG = 0.002;
        else if (LLCOSDIFF) {
          const double X = std::min(TAUWACT / sq(std::max(USTAR(IJ), t.EPSUS)), PMAX);
          double ZCHAR = std::min(ALPHAGM1 * sq(USTAR(IJ)) / std::sqrt(1.0 - X), 0.05 * std::exp(-0.05 * (u - 35.)));
          ZCHAR = std::min(ZCHAR, t.ALPHAMAX);
          CDFG = ACDLIN + BCDLIN * std::sqrt(ZCHAR) * u;
        } else CDFG = CDM(u);
        USTAR(IJ) = u * std::sqrt(CDFG);
      }
      const double W1 = 0.85 - 0.05 * (std::tanh(10.0 * (u - 5.0)) + 1.0);
      const double XKUTOP = t.XKAPPA * u;
      double USTOLD = USTAR(IJ), TAUOLD = USTOLD * USTOLD, TAUUNR = 0.0, X;
      int ITER;
      for (ITER = 1; ITER <= NITER; ++ITER) {
        Z0(IJ) = std::max(t.XNLEV / (std::exp(std::min(XKUTOP / USTOLD, 50.0)) - 1.0), Z0MIN);
        const double TAUV = RNUKAPPAM1 * USTOLD / Z0(IJ);
        const double ANG_GC = t.ANG_GC_A + t.ANG_GC_B * std::tanh(t.ANG_GC_C * TAUOLD);
        TAUUNR = STRESS_GC(x, ANG_GC, USTAR(IJ), Z0(IJ), Z0MIN, HALP(IJ), RNFAC(IJ));
        const double TAUNEW = TAUWEFF + TAUV + TAUUNR;
        const double USTNEW = std::sqrt(TAUNEW);
        USTAR(IJ) = W1 * USTOLD + (1.0 - W1) * USTNEW;
        const double DEL = USTAR(IJ) - USTOLD;
        if (std::fabs(DEL) < PCE_GC * USTAR(IJ)) break;
        TAUOLD = sq(USTAR(IJ));
        USTOLD = USTAR(IJ);
      }
      X = TAUWEFF / TAUOLD;
      if (ITER > NITER && X >= PMAX) {   // protection just in case there is no convergence
        const double CDFG = CDM(u);
        USTAR(IJ) = u * std::sqrt(CDFG);
        const double Z0MINRST = sq(USTAR(IJ)) * t.ALPHA * t.GM1;
        Z0(IJ) = std::max(t.XNLEV / (std::exp(XKUTOP / USTAR(IJ)) - 1.0), Z0MINRST);
        Z0B(IJ) = Z0MINRST;
      } else {
        Z0(IJ) = std::max(t.XNLEV / (std::exp(XKUTOP / USTAR(IJ)) - 1.0), Z0MIN);
        Z0B(IJ) = Z0(IJ) * std::sqrt(TAUUNR / TAUOLD);
      }
      if (X < PMAX) {                    // refine the solution (taut_z0.F90:230-276)
        const double USNRF = USTAR(IJ), Z0NRF = Z0(IJ), Z0BNRF = Z0B(IJ);
        USTOLD = USTAR(IJ);
        TAUOLD = std::max(USTOLD * USTOLD, TAUWEFF);
        const double ALPOG = std::max(std::min(Z0B(IJ) / TAUOLD, t.ALPHAMAX), ALPHAOG);
        double USTM1 = 0.0, Z0VIS = 0.0;
        for (ITER = 1; ITER <= NITER; ++ITER) {
          X = std::min(TAUWEFF / TAUOLD, PMAX);
          USTM1 = 1.0 / std::max(USTOLD, t.EPSUS);
          Z0VIS = c.rnum * USTM1;
          const double HZ0VISO1MX = 0.5 * Z0VIS / (1.0 - X);
          Z0B(IJ) = ALPOG * TAUOLD;
          Z0(IJ) = HZ0VISO1MX + std::sqrt(sq(HZ0VISO1MX) + sq(Z0B(IJ)) / (1.0 - X));
          const double XOLOGZ0 = 1.0 / std::log(t.XNLEV / Z0(IJ) + 1.0);
          const double F = USTOLD - XKUTOP * XOLOGZ0;
          const double ZZ = 2.0 * USTM1 * (3.0 * sq(Z0B(IJ)) + 0.5 * Z0VIS * Z0(IJ) - sq(Z0(IJ))) /
                            (2.0 * sq(Z0(IJ)) * (1.0 - X) - Z0VIS * Z0(IJ));
          const double DELF = 1.0 - XKUTOP * sq(XOLOGZ0) * ZZ;
          if (DELF != 0.0) USTAR(IJ) = USTOLD - F / DELF;
          const double TAUNEW = std::max(sq(USTAR(IJ)), TAUWEFF);
          USTAR(IJ) = std::sqrt(TAUNEW);
          const double DEL = TAUNEW - TAUOLD;
          if (std::fabs(DEL) < PCE_GC * TAUOLD) break;
          TAUOLD = TAUNEW;
          USTOLD = USTAR(IJ);
        }
        if (ITER > NITER) {
          USTAR(IJ) = USNRF; Z0(IJ) = Z0NRF; Z0B(IJ) = Z0BNRF;
          USTM1 = 1.0 / std::max(USTAR(IJ), t.EPSUS);
          Z0VIS = c.rnum * USTM1;
          CHRNCK(IJ) = std::max(t.G * (Z0(IJ) - Z0VIS) * sq(USTM1), t.ALPHAMIN);
        } else {
          CHRNCK(IJ) = std::max(t.G * (Z0B(IJ) / std::sqrt(1.0 - X)) / sq(std::max(USTAR(IJ), t.EPSUS)), t.ALPHAMIN);
        }
      } else {
        const double USTM1 = 1.0 / std::max(USTAR(IJ), t.EPSUS);
        const double Z0VIS = c.rnum * USTM1;
        CHRNCK(IJ) = std::max(t.G * (Z0(IJ) - Z0VIS) * sq(USTM1), t.ALPHAMIN);
      }
    }
    return;
  }

  const Tables& t = x.t;
  const Config& c = x.c;
  const int NITER = 18;
  const double TWOXMP1 = 3.0;
  double XLOGXL = std::log(t.XNLEV);
  double US2TOTAUW = 1.0 + t.EPS1;
  std::vector<double> TAUWACT(x.KIJL + 1), TAUWEFF(x.KIJL + 1), XMIN(x.KIJL + 1), ALPHAOG(x.KIJL + 1);
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
    double COSDIFF = std::cos(UDIR(IJ) - TAUWDIR(IJ));
    TAUWACT[IJ] = std::max(TAUW(IJ) * COSDIFF, t.EPSMIN);
  }
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) TAUWEFF[IJ] = TAUWACT[IJ] * US2TOTAUW;
  if (c.llcapchnk) {
    for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
      double CHARNOCK_MIN = CHNKMIN(t, UTOP(IJ));
      XMIN[IJ] = 0.15 * (t.ALPHA - CHARNOCK_MIN);
      ALPHAOG[IJ] = CHARNOCK_MIN * t.GM1;
    }
  } else {
    for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) { XMIN[IJ] = 0.0; ALPHAOG[IJ] = t.ALPHA * t.GM1; }
  }
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
    double XKUTOP = t.XKAPPA * UTOP(IJ);
    double USTOLD = (1 - IUSFG) * UTOP(IJ) * std::sqrt(std::min(t.ACD + t.BCD * UTOP(IJ), t.CDMAX)) + IUSFG * USTAR(IJ);
    double TAUOLD = std::max(sq(USTOLD), TAUWEFF[IJ]);
    USTAR(IJ) = std::sqrt(TAUOLD);
    double USTM1 = 1.0 / std::max(USTAR(IJ), t.EPSUS);
    double Z0CH = 0.0, Z0VIS, Z0TOT, X, XOLOGZ0, F, ZZ, DELF, TAUNEW;
    for (int ITER = 1; ITER <= NITER; ++ITER) {
      X = std::max(TAUWACT[IJ] / TAUOLD, XMIN[IJ]);
      Z0CH = ALPHAOG[IJ] * TAUOLD / std::sqrt(1.0 - X);
      Z0VIS = c.rnum * USTM1;
      Z0TOT = Z0CH + Z0VIS;
      XOLOGZ0 = 1.0 / (XLOGXL - std::log(Z0TOT));
      F = USTAR(IJ) - XKUTOP * XOLOGZ0;
      ZZ = USTM1 * (Z0CH * (2.0 - TWOXMP1 * X) / (1.0 - X) - Z0VIS) / Z0TOT;
      DELF = 1.0 - XKUTOP * sq(XOLOGZ0) * ZZ;
      if (DELF != 0.0) USTAR(IJ) = USTAR(IJ) - F / DELF;
      TAUNEW = std::max(sq(USTAR(IJ)), TAUWEFF[IJ]);
      USTAR(IJ) = std::sqrt(TAUNEW);
      if (TAUNEW == TAUOLD) break;
      USTM1 = 1.0 / std::max(USTAR(IJ), t.EPSUS);
      TAUOLD = TAUNEW;
    }
    Z0(IJ) = Z0CH;
    Z0B(IJ) = ALPHAOG[IJ] * TAUOLD;
    CHRNCK(IJ) = std::max(t.G * Z0(IJ) * sq(USTM1), t.ALPHAMIN);
  }
}

// wsigstar.F90:105-129 (LLGCBZ0=LLNORMAGAM=F branch)
void WSIGSTAR(const Ctx& x, V1 WSWAVE, V1 UFRIC, V1 Z0M, V1 WSTAR, V1 SIG_N) {
  const Tables& t = x.t;
  const double BG_GUST = 0.0, ONETHIRD = 1.0 / 3.0, SIG_NMAX = 0.9, C1 = 1.03e-3, C2 = 0.04e-3, P1 = 1.48, P2 = -0.21;
  if (x.c.llgcbz0 || x.c.llnormagam) {   // wsigstar.F90:87-103
    const double ZN = x.c.rnum;
    for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
      const double U10M1 = 1.0 / std::max(WSWAVE(IJ), x.c.wspmin);
      const double Z0VIS = ZN / std::max(UFRIC(IJ), t.EPSUS);
      double ZCHAR = t.G * (Z0M(IJ) - Z0VIS) / std::max(sq(UFRIC(IJ)), t.EPSUS);
      ZCHAR = std::max(std::min(ZCHAR, t.ALPHAMAX), t.ALPHAMIN);
      const double BCD_LOC = t.BCDLIN * std::sqrt(ZCHAR);
      const double C_D = t.ACDLIN + BCD_LOC * WSWAVE(IJ);
      const double DC_DDU = BCD_LOC;
      const double SIG_CONV = 1.0 + 0.5 * WSWAVE(IJ) / C_D * DC_DDU;
      SIG_N(IJ) = std::min(SIG_NMAX, SIG_CONV * U10M1 * std::pow(BG_GUST * UFRIC(IJ) * UFRIC(IJ) * UFRIC(IJ) + 0.5 * t.XKAPPA * WSTAR(IJ) * WSTAR(IJ) * WSTAR(IJ), ONETHIRD));
    }
    return;
  }
  double XKAPPAD = 1.0 / t.XKAPPA;
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
    double U10 = UFRIC(IJ) * XKAPPAD * (std::log(10.0) - std::log(Z0M(IJ)));
    U10 = std::max(U10, x.c.wspmin);
    double U10M1 = 1.0 / U10;
    double C2U10P1 = C2 * std::pow(U10, P1);
    double U10P2 = std::pow(U10, P2);
    double C_D = (C1 + C2U10P1) * U10P2;
    double DC_DDU = (P2 * C1 + (P1 + P2) * C2U10P1) * U10P2 * U10M1;
    double SIG_CONV = 1.0 + 0.5 * U10 / C_D * DC_DDU;
    SIG_N(IJ) = std::min(SIG_NMAX, SIG_CONV * U10M1 * std::pow(BG_GUST * UFRIC(IJ) * UFRIC(IJ) * UFRIC(IJ) + 0.5 * t.XKAPPA * WSTAR(IJ) * WSTAR(IJ) * WSTAR(IJ), ONETHIRD));
  }
}

// sinput_ard.F90:149-524
void SINPUT_ARD(const Ctx& x, int NGST, bool LLSNEG, V3 FL1, V2 WAVNUM, V2 CINV, V2 XK2CG, V1 WDWAVE, V1 WSWAVE,
                V1 UFRIC, V1 Z0M, V2 COSWDIF, V2 SINWDIF2, V1 RAORW, V1 WSTAR, V1 RNFAC, V3 FLD, V3 SL, V3 SPOS, V3 XLLWS) {
  const Tables& t = x.t;
  const Config& c = x.c;
  const int KIJS = x.KIJS, KIJL = x.KIJL, NANG = x.NANG, NFRE = x.NFRE;
  const int n = KIJL + 1;
  std::vector<double> CONSTF(n), Z0VIS(n), Z0NOZ(n), FWW(n), PVISC(n), PTURB(n), ZCN(n), SIG_Ns(n), UORBT(n), AORB(n),
      TEMP(n), RE(n), RE_C(n), ZORB(n), CNSN(n), FLP_AVG(n), SLP_AVG(n), ROGOROAIR(n), AIRD_PVISC(n), USG2(n), FLP(n),
      SLP(n), DSTAB1(n), TEMP1(n), TEMP2(n), CONST11(n), CONST22(n);
  std::vector<double> XSTRESS(2 * n), YSTRESS(2 * n), TAUX(2 * n), TAUY(2 * n), USTP(2 * n), USTPM1(2 * n), USDIRP(2 * n),
      UCN(2 * n), UCNZALPD(2 * n), GAMNORMA(2 * n, 1.0);
  auto i2 = [n](int ij, int ig) { return ij + n * (ig - 1); };
  std::vector<double> GAM0((size_t)n * NANG * 2), DSTAB((size_t)n * NANG * 2, 0.0), COSLP((size_t)n * NANG);
  auto i3 = [n, NANG](int ij, int k, int ig) { return ij + (size_t)n * ((k - 1) + (size_t)NANG * (ig - 1)); };
  auto ik = [n](int ij, int k) { return ij + (size_t)n * (k - 1); };

  double AVG_GST = 1.0 / NGST;
  double CONST1 = t.BETAMAXOXKAPPA2;
  const double CONSTN = t.DELTH / (t.XKAPPA * t.ZPI);
  std::vector<double> CSTRNFAC(n, 0.0), XNGAMCONST(n, 0.0);
  if (c.llnormagam) for (int IJ = KIJS; IJ <= KIJL; ++IJ) CSTRNFAC[IJ] = CONSTN * RNFAC(IJ) / RAORW(IJ);   // sinput_ard.F90:172-176
  double ABS_TAUWSHELTER = std::fabs(t.TAUWSHELTER);
  bool LTAUWSHELTER = ABS_TAUWSHELTER != 0.0;
  if (NGST > 1) WSIGSTAR(x, WSWAVE, UFRIC, Z0M, WSTAR, V1{SIG_Ns.data() + 1});
  double NU_AIR = 0, FACM1_NU_AIR = 0, FAC_NU_AIR = 0, FU = 0, FUD = 0, DELABM1 = 0;
  if (LLSNEG) {
    NU_AIR = c.rnu;
    FACM1_NU_AIR = 4.0 / NU_AIR;
    FAC_NU_AIR = c.rnum;
    FU = std::fabs(t.SWELLF3);
    FUD = t.SWELLF2;
    DELABM1 = (double)t.IAB / (t.ABMAX - t.ABMIN);
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) { UORBT[IJ] = t.EPSMIN; AORB[IJ] = t.EPSMIN; }
    for (int M = 1; M <= NFRE; ++M) {
      double SIG = t.ZPIFR(M), SIG2 = SIG * SIG, DFIM_SIG2 = t.DFIM(M) * SIG2;
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) TEMP[IJ] = FL1(IJ, 1, M);
      for (int K = 2; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) TEMP[IJ] = TEMP[IJ] + FL1(IJ, K, M);
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) { UORBT[IJ] = UORBT[IJ] + DFIM_SIG2 * TEMP[IJ]; AORB[IJ] = AORB[IJ] + t.DFIM(M) * TEMP[IJ]; }
    }
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      UORBT[IJ] = 2.0 * std::sqrt(UORBT[IJ]);
      AORB[IJ] = 2.0 * std::sqrt(AORB[IJ]);
      RE[IJ] = FACM1_NU_AIR * UORBT[IJ] * AORB[IJ];
      Z0VIS[IJ] = FAC_NU_AIR / std::max(UFRIC(IJ), 0.0001);
      double Z0TUB = t.Z0RAT * std::min(t.Z0TUBMAX, Z0M(IJ));
      Z0NOZ[IJ] = std::max(Z0VIS[IJ], Z0TUB);
      ZORB[IJ] = AORB[IJ] / Z0NOZ[IJ];
      double XI = (std::log10(std::max(ZORB[IJ], 3.0)) - t.ABMIN) * DELABM1;
      int IND = std::min(t.IAB - 1, (int)XI);
      double DELI1 = std::min(1.0, XI - (double)IND);
      double DELI2 = 1.0 - DELI1;
      FWW[IJ] = t.SWELLFT(IND) * DELI2 + t.SWELLFT(IND + 1) * DELI1;
      TEMP2[IJ] = FWW[IJ] * UORBT[IJ];
    }
    if (t.SWELLF6 == 1.0) { for (int IJ = KIJS; IJ <= KIJL; ++IJ) RE_C[IJ] = t.SWELLF4; }
    else { double H = 1.0 - t.SWELLF6; for (int IJ = KIJS; IJ <= KIJL; ++IJ) RE_C[IJ] = t.SWELLF4 * std::pow(2.0 / AORB[IJ], H); }
    if (t.SWELLF7 > 0.0) {
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
        double SMOOTH = 0.5 * std::tanh((RE[IJ] - RE_C[IJ]) * t.SWELLF7M1);
        PTURB[IJ] = 0.5 + SMOOTH; PVISC[IJ] = 0.5 - SMOOTH;
      }
    } else {
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
        if (RE[IJ] <= RE_C[IJ]) { PTURB[IJ] = 0.0; PVISC[IJ] = 0.5; } else { PTURB[IJ] = 0.5; PVISC[IJ] = 0.0; }
      }
    }
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) AIRD_PVISC[IJ] = PVISC[IJ] * RAORW(IJ);
  }
  if (NGST == 1) { for (int IJ = KIJS; IJ <= KIJL; ++IJ) USTP[i2(IJ, 1)] = UFRIC(IJ); }
  else if (NGST == 2) {
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) { USTP[i2(IJ, 1)] = UFRIC(IJ) * (1.0 + SIG_Ns[IJ]); USTP[i2(IJ, 2)] = UFRIC(IJ) * (1.0 - SIG_Ns[IJ]); }
  } else throw std::runtime_error("SINPUT_ARD: NGST > 2");
  for (int IGST = 1; IGST <= NGST; ++IGST) for (int IJ = KIJS; IJ <= KIJL; ++IJ) USTPM1[i2(IJ, IGST)] = 1.0 / std::max(USTP[i2(IJ, IGST)], t.EPSUS);
  if (LTAUWSHELTER) {
    for (int IGST = 1; IGST <= NGST; ++IGST)
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
        XSTRESS[i2(IJ, IGST)] = 0.0; YSTRESS[i2(IJ, IGST)] = 0.0;
        USG2[IJ] = sq(USTP[i2(IJ, IGST)]);
        TAUX[i2(IJ, IGST)] = USG2[IJ] * std::sin(WDWAVE(IJ));
        TAUY[i2(IJ, IGST)] = USG2[IJ] * std::cos(WDWAVE(IJ));
      }
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) ROGOROAIR[IJ] = t.G / RAORW(IJ);
  } else {
    for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) COSLP[ik(IJ, K)] = COSWDIF(IJ, K);
  }
  double COEF = 0, COEF5 = 0;
  for (int M = 1; M <= NFRE; ++M) {
    double SIG = t.ZPIFR(M), SIG2 = SIG * SIG, CONST = SIG * CONST1;
    if (LLSNEG) { COEF = -t.SWELLF * 16. * SIG2 / t.G; COEF5 = -t.SWELLF5 * 2. * std::sqrt(2. * NU_AIR * SIG); }
    if (LTAUWSHELTER) {
      for (int IGST = 1; IGST <= NGST; ++IGST)
        for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
          double TAUPX = TAUX[i2(IJ, IGST)] - ABS_TAUWSHELTER * XSTRESS[i2(IJ, IGST)];
          double TAUPY = TAUY[i2(IJ, IGST)] - ABS_TAUWSHELTER * YSTRESS[i2(IJ, IGST)];
          USDIRP[i2(IJ, IGST)] = std::atan2(TAUPX, TAUPY);
          USTP[i2(IJ, IGST)] = std::pow(TAUPX * TAUPX + TAUPY * TAUPY, 0.25);
          USTPM1[i2(IJ, IGST)] = 1.0 / std::max(USTP[i2(IJ, IGST)], t.EPSUS);
        }
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) CONSTF[IJ] = ROGOROAIR[IJ] * CINV(IJ, M) * t.DFIM(M);
    }
    for (int IGST = 1; IGST <= NGST; ++IGST)
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
        UCN[i2(IJ, IGST)] = USTP[i2(IJ, IGST)] * CINV(IJ, M);
        UCNZALPD[i2(IJ, IGST)] = t.XKAPPA / (UCN[i2(IJ, IGST)] + t.ZALP);
      }
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) { ZCN[IJ] = std::log(WAVNUM(IJ, M) * Z0M(IJ)); CNSN[IJ] = CONST * RAORW(IJ); }
    for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) XLLWS(IJ, K, M) = 0.0;
    if (c.llnormagam) for (int IJ = KIJS; IJ <= KIJL; ++IJ) XNGAMCONST[IJ] = CSTRNFAC[IJ] * XK2CG(IJ, M);
    if (LLSNEG) for (int IJ = KIJS; IJ <= KIJL; ++IJ) { DSTAB1[IJ] = COEF5 * AIRD_PVISC[IJ] * WAVNUM(IJ, M); TEMP1[IJ] = COEF * RAORW(IJ); }
    for (int IGST = 1; IGST <= NGST; ++IGST) {
      for (int K = 1; K <= NANG; ++K) {
        if (LTAUWSHELTER) for (int IJ = KIJS; IJ <= KIJL; ++IJ) COSLP[ik(IJ, K)] = std::cos(t.TH(K) - USDIRP[i2(IJ, IGST)]);
        for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
          GAM0[i3(IJ, K, IGST)] = 0.0;
          if (COSLP[ik(IJ, K)] > 0.01) {
            double X = COSLP[ik(IJ, K)] * UCN[i2(IJ, IGST)];
            double ZLOG = ZCN[IJ] + UCNZALPD[i2(IJ, IGST)] / COSLP[ik(IJ, K)];
            if (ZLOG < 0.0) {
              double ZLOG2X = ZLOG * ZLOG * X;
              GAM0[i3(IJ, K, IGST)] = std::exp(ZLOG) * ZLOG2X * ZLOG2X * CNSN[IJ];
              XLLWS(IJ, K, M) = 1.0;
            }
          }
        }
      }
      if (c.llnormagam) {   // sinput_ard.F90:436-452
        for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
          double SUMF = 0.0, SUMFSIN2 = 0.0;
          for (int K = 1; K <= NANG; ++K) {
            SUMF = SUMF + GAM0[i3(IJ, K, IGST)] * FL1(IJ, K, M);
            SUMFSIN2 = SUMFSIN2 + GAM0[i3(IJ, K, IGST)] * FL1(IJ, K, M) * SINWDIF2(IJ, K);
          }
          const double ZNZ = XNGAMCONST[IJ] * USTPM1[i2(IJ, IGST)];
          GAMNORMA[i2(IJ, IGST)] = (1.0 + ZNZ * SUMFSIN2) / (1.0 + ZNZ * SUMF);
        }
      }
      if (LLSNEG)
        for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
          double DSTAB2 = TEMP1[IJ] * (TEMP2[IJ] + (FU + FUD * COSLP[ik(IJ, K)]) * USTP[i2(IJ, IGST)]);
          DSTAB[i3(IJ, K, IGST)] = DSTAB1[IJ] + PTURB[IJ] * DSTAB2;
        }
    }
    for (int K = 1; K <= NANG; ++K) {
      for (int IGST = 1; IGST <= NGST; ++IGST) {
        for (int IJ = KIJS; IJ <= KIJL; ++IJ) { SLP[IJ] = GAM0[i3(IJ, K, IGST)] * GAMNORMA[i2(IJ, IGST)]; FLP[IJ] = SLP[IJ] + DSTAB[i3(IJ, K, IGST)]; }
        for (int IJ = KIJS; IJ <= KIJL; ++IJ) SLP[IJ] = SLP[IJ] * FL1(IJ, K, M);
        if (LTAUWSHELTER) {
          for (int IJ = KIJS; IJ <= KIJL; ++IJ) { CONST11[IJ] = CONSTF[IJ] * t.SINTH(K); CONST22[IJ] = CONSTF[IJ] * t.COSTH(K); }
          for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
            XSTRESS[i2(IJ, IGST)] = XSTRESS[i2(IJ, IGST)] + SLP[IJ] * CONST11[IJ];
            YSTRESS[i2(IJ, IGST)] = YSTRESS[i2(IJ, IGST)] + SLP[IJ] * CONST22[IJ];
          }
        }
        if (IGST == 1) { for (int IJ = KIJS; IJ <= KIJL; ++IJ) { SLP_AVG[IJ] = SLP[IJ]; FLP_AVG[IJ] = FLP[IJ]; } }
        else { for (int IJ = KIJS; IJ <= KIJL; ++IJ) { SLP_AVG[IJ] = SLP_AVG[IJ] + SLP[IJ]; FLP_AVG[IJ] = FLP_AVG[IJ] + FLP[IJ]; } }
      }
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) SPOS(IJ, K, M) = AVG_GST * SLP_AVG[IJ];
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) { FLD(IJ, K, M) = AVG_GST * FLP_AVG[IJ]; SL(IJ, K, M) = FLD(IJ, K, M) * FL1(IJ, K, M); }
    }
  }
}

// sinput_jan.F90:150-400
void SINPUT_JAN(const Ctx& x, int NGST, bool LLSNEG, V3 FL1, V2 WAVNUM, V2 CINV, V2 XK2CG, V1 WSWAVE, V1 UFRIC, V1 Z0M,
                V2 COSWDIF, V2 SINWDIF2, V1 RAORW, V1 WSTAR, V1 RNFAC, V3 FLD, V3 SL, V3 SPOS, V3 XLLWS) {
  const Tables& t = x.t;
  const Config& c = x.c;
  const int KIJS = x.KIJS, KIJL = x.KIJL, NANG = x.NANG, NFRE = x.NFRE;
  const int n = KIJL + 1;
  auto i2 = [n](int ij, int ig) { return ij + n * (ig - 1); };
  auto i3 = [n, NANG](int ij, int k, int ig) { return ij + (size_t)n * ((k - 1) + (size_t)NANG * (ig - 1)); };
  auto ik = [n](int ij, int k) { return ij + (size_t)n * (k - 1); };
  std::vector<double> ZTANHKD(n), SIG_Ns(n), CNSN(n), UFAC1(n), UFAC2(n);
  std::vector<double> SIGDEV(2 * n), US(2 * n), Z0(2 * n), UCN(2 * n), ZCN(2 * n), USTPM1(2 * n), XVD(2 * n), UCND(2 * n),
      CONST3_UCN2(2 * n), GAMNORMA(2 * n, 1.0);
  std::vector<double> GAM0((size_t)n * NANG * 2);
  std::vector<char> LZ((size_t)n * NANG);
  double WSIN[3] = {0, 0, 0};
  double CONST1 = t.BETAMAXOXKAPPA2;
  double CONST3 = 2.0 * t.XKAPPA / CONST1;
  double XKAPPAD = 1.E0 / t.XKAPPA;
  CONST3 = c.idamping * CONST3;
  if (NGST > 1) WSIGSTAR(x, WSWAVE, UFRIC, Z0M, WSTAR, V1{SIG_Ns.data() + 1});
  for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) LZ[ik(IJ, K)] = COSWDIF(IJ, K) > 0.01;
  if (NGST == 1) { WSIN[1] = 1.0; for (int IJ = KIJS; IJ <= KIJL; ++IJ) SIGDEV[i2(IJ, 1)] = 1.0; }
  else if (NGST == 2) {
    WSIN[1] = 0.5; WSIN[2] = 0.5;
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) { SIGDEV[i2(IJ, 1)] = 1.0 - SIG_Ns[IJ]; SIGDEV[i2(IJ, 2)] = 1.0 + SIG_Ns[IJ]; }
  } else throw std::runtime_error("SINPUT_JAN: NGST > 2");
  if (NGST == 1) { for (int IJ = KIJS; IJ <= KIJL; ++IJ) { US[i2(IJ, 1)] = UFRIC(IJ); Z0[i2(IJ, 1)] = Z0M(IJ); } }
  else for (int IGST = 1; IGST <= NGST; ++IGST) for (int IJ = KIJS; IJ <= KIJL; ++IJ) { US[i2(IJ, IGST)] = UFRIC(IJ) * SIGDEV[i2(IJ, IGST)]; Z0[i2(IJ, IGST)] = Z0M(IJ); }
  for (int IGST = 1; IGST <= NGST; ++IGST) for (int IJ = KIJS; IJ <= KIJL; ++IJ) USTPM1[i2(IJ, IGST)] = 1.0 / std::max(US[i2(IJ, IGST)], t.EPSUS);
  for (int M = 1; M <= NFRE; ++M) {
    double CONST = t.ZPIFR(M) * CONST1;
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) ZTANHKD[IJ] = sq(t.ZPIFR(M)) / (t.G * WAVNUM(IJ, M));
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) CNSN[IJ] = CONST * ZTANHKD[IJ] * RAORW(IJ);
    for (int IGST = 1; IGST <= NGST; ++IGST)
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
        UCN[i2(IJ, IGST)] = US[i2(IJ, IGST)] * CINV(IJ, M) + t.ZALP;
        CONST3_UCN2[i2(IJ, IGST)] = CONST3 * sq(UCN[i2(IJ, IGST)]);
        UCND[i2(IJ, IGST)] = 1.0 / UCN[i2(IJ, IGST)];
        ZCN[i2(IJ, IGST)] = std::log(WAVNUM(IJ, M) * Z0[i2(IJ, IGST)]);
        XVD[i2(IJ, IGST)] = 1.0 / (-US[i2(IJ, IGST)] * XKAPPAD * ZCN[i2(IJ, IGST)] * CINV(IJ, M));
      }
    for (int K = 1; K <= NANG; ++K) {
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) XLLWS(IJ, K, M) = 0.0;
      for (int IGST = 1; IGST <= NGST; ++IGST)
        for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
          if (LZ[ik(IJ, K)]) {
            double ZLOG = ZCN[i2(IJ, IGST)] + t.XKAPPA / COSWDIF(IJ, K) * UCND[i2(IJ, IGST)];
            if (ZLOG < 0.0) {
              double X = COSWDIF(IJ, K) * UCN[i2(IJ, IGST)];
              double ZLOG2X = ZLOG * ZLOG * X;
              GAM0[i3(IJ, K, IGST)] = ZLOG2X * ZLOG2X * std::exp(ZLOG) * CNSN[IJ];
              XLLWS(IJ, K, M) = 1.0;
            } else GAM0[i3(IJ, K, IGST)] = 0.0;
          } else GAM0[i3(IJ, K, IGST)] = 0.0;
        }
    }
    if (c.llnormagam) {   // sinput_jan.F90:329-348
      const double CONSTN = t.DELTH / (t.XKAPPA * t.ZPI);
      for (int IGST = 1; IGST <= NGST; ++IGST)
        for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
          const double XNGAMCONST = (CONSTN * RNFAC(IJ) / RAORW(IJ)) * XK2CG(IJ, M);
          double SUMF = 0.0, SUMFSIN2 = 0.0;
          for (int K = 1; K <= NANG; ++K) {
            SUMF = SUMF + GAM0[i3(IJ, K, IGST)] * FL1(IJ, K, M);
            SUMFSIN2 = SUMFSIN2 + GAM0[i3(IJ, K, IGST)] * FL1(IJ, K, M) * SINWDIF2(IJ, K);
          }
          const double ZNZ = XNGAMCONST * USTPM1[i2(IJ, IGST)];
          GAMNORMA[i2(IJ, IGST)] = (1.0 + ZNZ * SUMFSIN2) / (1.0 + ZNZ * SUMF);
        }
    }
    for (int K = 1; K <= NANG; ++K) {
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) UFAC1[IJ] = WSIN[1] * GAM0[i3(IJ, K, 1)] * GAMNORMA[i2(IJ, 1)];
      if (NGST == 2) for (int IJ = KIJS; IJ <= KIJL; ++IJ) UFAC1[IJ] = UFAC1[IJ] + WSIN[2] * GAM0[i3(IJ, K, 2)] * GAMNORMA[i2(IJ, 2)];
      if (LLSNEG) {
        for (int IJ = KIJS; IJ <= KIJL; ++IJ) { double ZBETA = CONST3_UCN2[i2(IJ, 1)] * (COSWDIF(IJ, K) - XVD[i2(IJ, 1)]); UFAC2[IJ] = WSIN[1] * ZBETA; }
        if (NGST == 2) for (int IJ = KIJS; IJ <= KIJL; ++IJ) { double ZBETA = CONST3_UCN2[i2(IJ, 2)] * (COSWDIF(IJ, K) - XVD[i2(IJ, 2)]); UFAC2[IJ] = UFAC2[IJ] + WSIN[2] * ZBETA; }
      } else for (int IJ = KIJS; IJ <= KIJL; ++IJ) UFAC2[IJ] = 0.0;
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
        FLD(IJ, K, M) = UFAC1[IJ] + UFAC2[IJ] * CNSN[IJ];
        SPOS(IJ, K, M) = UFAC1[IJ] * FL1(IJ, K, M);
        SL(IJ, K, M) = FLD(IJ, K, M) * FL1(IJ, K, M);
      }
    }
  }
}

// tau_phi_hf.F90:111-305
void TAU_PHI_HF(const Ctx& x, I1 MIJ, bool LTAUWSHELTER, V1 UFRIC, V1 Z0M, V3 FL1, V1 AIRD, V1 RNFAC, V2 COSWDIF, V2 SINWDIF2,
                V1 UST, V1 TAUHF, V1 PHIHF, bool LLPHIHF) {
  const Tables& t = x.t;
  const int KIJS = x.KIJS, KIJL = x.KIJL, NANG = x.NANG;
  const int n = KIJL + 1;
  const double ZSUPMAX = 0.0;
  std::vector<double> SQRTZ0OG(n), ZSUP(n), ZINF(n), DELZ(n), TAUL(n), XLOGGZ0(n), SQRTGZ0(n), USTPH(n), CONST1(n, 0.0),
      CONST2(n, 0.0), CONSTTAU(n), CONSTPHI(n), F1DCOS2(n), F1DCOS3(n), F1D(n), F1DSIN2(n);
  double X0G = t.X0TAUHF * t.G;
  if (LLPHIHF) for (int IJ = KIJS; IJ <= KIJL; ++IJ) USTPH[IJ] = UST(IJ);
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
    XLOGGZ0[IJ] = std::log(t.G * Z0M(IJ));
    double OMEGACC = std::max(t.ZPIFR(MIJ(IJ)), X0G / UST(IJ));
    SQRTZ0OG[IJ] = std::sqrt(Z0M(IJ) * t.GM1);
    SQRTGZ0[IJ] = 1.0 / SQRTZ0OG[IJ];
    double YC = OMEGACC * SQRTZ0OG[IJ];
    ZINF[IJ] = std::log(YC);
  }
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) CONSTTAU[IJ] = t.ZPI4GM2 * t.FR5(MIJ(IJ));
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
    double COSW = std::max(COSWDIF(IJ, 1), 0.0);
    double FCOSW2 = FL1(IJ, 1, MIJ(IJ)) * COSW * COSW;
    F1DCOS3[IJ] = FCOSW2 * COSW; F1DCOS2[IJ] = FCOSW2;
    F1DSIN2[IJ] = FL1(IJ, 1, MIJ(IJ)) * SINWDIF2(IJ, 1);
    F1D[IJ] = FL1(IJ, 1, MIJ(IJ));
  }
  for (int K = 2; K <= NANG; ++K)
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      double COSW = std::max(COSWDIF(IJ, K), 0.0);
      double FCOSW2 = FL1(IJ, K, MIJ(IJ)) * COSW * COSW;
      F1DCOS3[IJ] = F1DCOS3[IJ] + FCOSW2 * COSW;
      F1DCOS2[IJ] = F1DCOS2[IJ] + FCOSW2;
      F1DSIN2[IJ] = F1DSIN2[IJ] + FL1(IJ, K, MIJ(IJ)) * SINWDIF2(IJ, K);
      F1D[IJ] = F1D[IJ] + FL1(IJ, K, MIJ(IJ));
    }
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) { F1DCOS3[IJ] = t.DELTH * F1DCOS3[IJ]; F1DCOS2[IJ] = t.DELTH * F1DCOS2[IJ]; F1DSIN2[IJ] = t.DELTH * F1DSIN2[IJ]; F1D[IJ] = t.DELTH * F1D[IJ]; }
  if (x.c.llnormagam)   // tau_phi_hf.F90:177-182
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      const double CONFG = t.GAMNCONST * t.FR5(MIJ(IJ)) * RNFAC(IJ) * SQRTGZ0[IJ];
      CONST1[IJ] = CONFG * F1DSIN2[IJ];
      CONST2[IJ] = CONFG * F1D[IJ];
    }
  if (x.c.llgcbz0)      // omegagc.F90:51-55, tau_phi_hf.F90:190-193
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) ZSUP[IJ] = std::min(std::log(t.OMEGA_GC(NS_GC(t, UFRIC(IJ))) * SQRTZ0OG[IJ]), ZSUPMAX);
  else
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) ZSUP[IJ] = ZSUPMAX;
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
    TAUL[IJ] = sq(UST(IJ));
    DELZ[IJ] = std::max((ZSUP[IJ] - ZINF[IJ]) / (double)(t.JTOT_TAUHF - 1), 0.0);
    TAUHF(IJ) = 0.0;
  }
  if (LTAUWSHELTER) {
    for (int IJ = KIJS; IJ <= KIJL; ++IJ)
      for (int J = 1; J <= t.JTOT_TAUHF; ++J) {
        double Y = std::exp(ZINF[IJ] + (double)(J - 1) * DELZ[IJ]);
        double OMEGA = Y * SQRTGZ0[IJ];
        double CM1 = OMEGA * t.GM1;
        double ZX = UST(IJ) * CM1 + t.ZALP;
        double ZARG = t.XKAPPA / ZX;
        double ZLOG = XLOGGZ0[IJ] + 2.0 * std::log(CM1) + ZARG;
        ZLOG = std::min(ZLOG, 0.0);
        double ZBETA = p4(ZLOG) * std::exp(ZLOG);
        double ZNZ = ZBETA * UST(IJ) * Y;
        double GAMNORMA = (1.0 + CONST1[IJ] * ZNZ) / (1.0 + CONST2[IJ] * ZNZ);
        double FNC2 = F1DCOS3[IJ] * CONSTTAU[IJ] * ZBETA * TAUL[IJ] * t.WTAUHF(J) * DELZ[IJ] * GAMNORMA;
        TAUL[IJ] = std::max(TAUL[IJ] - t.TAUWSHELTER * FNC2, 0.0);
        UST(IJ) = std::sqrt(TAUL[IJ]);
        TAUHF(IJ) = TAUHF(IJ) + FNC2;
      }
  } else {
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      for (int J = 1; J <= t.JTOT_TAUHF; ++J) {
        double Y = std::exp(ZINF[IJ] + (double)(J - 1) * DELZ[IJ]);
        double OMEGA = Y * SQRTGZ0[IJ];
        double CM1 = OMEGA * t.GM1;
        double ZX = UST(IJ) * CM1 + t.ZALP;
        double ZARG = t.XKAPPA / ZX;
        double ZLOG = XLOGGZ0[IJ] + 2.0 * std::log(CM1) + ZARG;
        ZLOG = std::min(ZLOG, 0.0);
        double ZBETA = p4(ZLOG) * std::exp(ZLOG);
        double FNC2 = ZBETA * t.WTAUHF(J);
        double ZNZ = ZBETA * UST(IJ) * Y;
        double GAMNORMA = (1.0 + CONST1[IJ] * ZNZ) / (1.0 + CONST2[IJ] * ZNZ);
        TAUHF(IJ) = TAUHF(IJ) + FNC2 * GAMNORMA;
      }
      TAUHF(IJ) = F1DCOS3[IJ] * CONSTTAU[IJ] * TAUL[IJ] * TAUHF(IJ) * DELZ[IJ];
    }
  }
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) PHIHF(IJ) = 0.0;
  if (LLPHIHF) {
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      TAUL[IJ] = sq(USTPH[IJ]);
      ZSUP[IJ] = ZSUPMAX;
      DELZ[IJ] = std::max((ZSUP[IJ] - ZINF[IJ]) / (double)(t.JTOT_TAUHF - 1), 0.0);
    }
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) CONSTPHI[IJ] = AIRD(IJ) * t.ZPI4GM1 * t.FR5(MIJ(IJ));
    if (LTAUWSHELTER) {
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
        for (int J = 1; J <= t.JTOT_TAUHF; ++J) {
          double Y = std::exp(ZINF[IJ] + (double)(J - 1) * DELZ[IJ]);
          double OMEGA = Y * SQRTGZ0[IJ];
          double CM1 = OMEGA * t.GM1;
          double ZX = USTPH[IJ] * CM1 + t.ZALP;
          double ZARG = t.XKAPPA / ZX;
          double ZLOG = XLOGGZ0[IJ] + 2.0 * std::log(CM1) + ZARG;
          ZLOG = std::min(ZLOG, 0.0);
          double ZBETA = p4(ZLOG) * std::exp(ZLOG);
          double ZNZ = ZBETA * UST(IJ) * Y;
          double GAMNORMA = (1.0 + CONST1[IJ] * ZNZ) / (1.0 + CONST2[IJ] * ZNZ);
          double FNC2 = ZBETA * TAUL[IJ] * t.WTAUHF(J) * DELZ[IJ] * GAMNORMA;
          TAUL[IJ] = std::max(TAUL[IJ] - t.TAUWSHELTER * F1DCOS3[IJ] * CONSTTAU[IJ] * FNC2, 0.0);
          USTPH[IJ] = std::sqrt(TAUL[IJ]);
          PHIHF(IJ) = PHIHF(IJ) + FNC2 / Y;
        }
        PHIHF(IJ) = F1DCOS2[IJ] * CONSTPHI[IJ] * SQRTZ0OG[IJ] * PHIHF(IJ);
      }
    } else {
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
        for (int J = 1; J <= t.JTOT_TAUHF; ++J) {
          double Y = std::exp(ZINF[IJ] + (double)(J - 1) * DELZ[IJ]);
          double OMEGA = Y * SQRTGZ0[IJ];
          double CM1 = OMEGA * t.GM1;
          double ZX = USTPH[IJ] * CM1 + t.ZALP;
          double ZARG = t.XKAPPA / ZX;
          double ZLOG = XLOGGZ0[IJ] + 2.0 * std::log(CM1) + ZARG;
          ZLOG = std::min(ZLOG, 0.0);
          double ZBETA = p4(ZLOG) * std::exp(ZLOG);
          double ZNZ = ZBETA * UST(IJ) * Y;
          double GAMNORMA = (1.0 + CONST1[IJ] * ZNZ) / (1.0 + CONST2[IJ] * ZNZ);
          double FNC2 = ZBETA * t.WTAUHF(J) * GAMNORMA;
          PHIHF(IJ) = PHIHF(IJ) + FNC2 / Y;
        }
        PHIHF(IJ) = F1DCOS2[IJ] * CONSTPHI[IJ] * SQRTZ0OG[IJ] * TAUL[IJ] * PHIHF(IJ) * DELZ[IJ];
      }
    }
  }
}

// stresso.F90:120-233
void STRESSO(const Ctx& x, I1 MIJ, V2 RHOWGDFTH, V3 FL1, V3 SL, V3 SPOS, V2 CINV, V1 WDWAVE, V1 UFRIC, V1 Z0M, V1 AIRD, V1 RNFAC,
             V2 COSWDIF, V2 SINWDIF2, V1 TAUW, V1 TAUWDIR, V1 PHIWA, bool LLPHIWA) {
  const Tables& t = x.t;
  const int KIJS = x.KIJS, KIJL = x.KIJL, NANG = x.NANG, NFRE = x.NFRE;
  const int n = KIJL + 1;
  std::vector<double> XSTRESS(n), YSTRESS(n), TAUHF(n), PHIHF(n), USDIRP(n), UST(n), SUMT(n), SUMX(n), SUMY(n);
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) { PHIWA(IJ) = 0.0; XSTRESS[IJ] = 0.0; YSTRESS[IJ] = 0.0; }
  if (LLPHIWA)
    for (int M = 1; M <= NFRE; ++M) for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ)
      PHIWA(IJ) = PHIWA(IJ) + (SL(IJ, K, M) - SPOS(IJ, K, M)) * t.RHOWG_DFIM(M);
  for (int M = 1; M <= NFRE; ++M) {
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) { SUMX[IJ] = SPOS(IJ, 1, M) * t.SINTH(1); SUMY[IJ] = SPOS(IJ, 1, M) * t.COSTH(1); SUMT[IJ] = SPOS(IJ, 1, M); }
    for (int K = 2; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      SUMX[IJ] = SUMX[IJ] + SPOS(IJ, K, M) * t.SINTH(K);
      SUMY[IJ] = SUMY[IJ] + SPOS(IJ, K, M) * t.COSTH(K);
      SUMT[IJ] = SUMT[IJ] + SPOS(IJ, K, M);
    }
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      double CMRHOWGDFTH = RHOWGDFTH(IJ, M) * CINV(IJ, M);
      XSTRESS[IJ] = XSTRESS[IJ] + CMRHOWGDFTH * SUMX[IJ];
      YSTRESS[IJ] = YSTRESS[IJ] + CMRHOWGDFTH * SUMY[IJ];
    }
    if (LLPHIWA) for (int IJ = KIJS; IJ <= KIJL; ++IJ) PHIWA(IJ) = PHIWA(IJ) + RHOWGDFTH(IJ, M) * SUMT[IJ];
  }
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) { XSTRESS[IJ] = XSTRESS[IJ] / std::max(AIRD(IJ), 1.0); YSTRESS[IJ] = YSTRESS[IJ] / std::max(AIRD(IJ), 1.0); }
  bool LTAUWSHELTER;
  if (x.c.iphys == 0 || t.TAUWSHELTER == 0.0) {
    LTAUWSHELTER = false;
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) { USDIRP[IJ] = WDWAVE(IJ); UST[IJ] = UFRIC(IJ); }
  } else {
    LTAUWSHELTER = true;
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      double TAUX = sq(UFRIC(IJ)) * std::sin(WDWAVE(IJ));
      double TAUY = sq(UFRIC(IJ)) * std::cos(WDWAVE(IJ));
      double TAUPX = TAUX - t.TAUWSHELTER * XSTRESS[IJ];
      double TAUPY = TAUY - t.TAUWSHELTER * YSTRESS[IJ];
      USDIRP[IJ] = std::atan2(TAUPX, TAUPY);
      UST[IJ] = std::pow(TAUPX * TAUPX + TAUPY * TAUPY, 0.25);
    }
  }
  TAU_PHI_HF(x, MIJ, LTAUWSHELTER, UFRIC, Z0M, FL1, AIRD, RNFAC, COSWDIF, SINWDIF2, V1{UST.data() + 1}, V1{TAUHF.data() + 1},
             V1{PHIHF.data() + 1}, LLPHIWA);
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
    XSTRESS[IJ] = XSTRESS[IJ] + TAUHF[IJ] * std::sin(USDIRP[IJ]);
    YSTRESS[IJ] = YSTRESS[IJ] + TAUHF[IJ] * std::cos(USDIRP[IJ]);
    TAUW(IJ) = std::sqrt(sq(XSTRESS[IJ]) + sq(YSTRESS[IJ]));
    TAUW(IJ) = std::max(TAUW(IJ), 0.0);
    TAUWDIR(IJ) = std::atan2(XSTRESS[IJ], YSTRESS[IJ]);
  }
  if (!x.c.llgcbz0) {
    double TAUTOUS2 = 1.0 / (1.0 + t.EPS1);
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) TAUW(IJ) = std::min(TAUW(IJ), sq(UFRIC(IJ)) * TAUTOUS2);
  }
  if (LLPHIWA) for (int IJ = KIJS; IJ <= KIJL; ++IJ) PHIWA(IJ) = PHIWA(IJ) + PHIHF[IJ];
}

// sdissip_ard.F90:131-318 (SSDSC3 = 0 and SSDSC5 = 0: cumulative and turbulence terms are dead, setwavphys.F90:152,177)
void SDISSIP_ARD(const Ctx& x, V3 FL1, V3 FLD, V3 SL, V2 WAVNUM, V2 CGROUP, V2 XK2CG, V1 UFRIC, V2 COSWDIF, V1 RAORW) {
  (void)CGROUP; (void)UFRIC; (void)COSWDIF; (void)RAORW;
  const Tables& t = x.t;
  const int KIJS = x.KIJS, KIJL = x.KIJL, NANG = x.NANG, NFRE = x.NFRE;
  if (t.SSDSC3 != 0.0 || t.SSDSC5 != 0.0) throw std::runtime_error("SDISSIP_ARD: SSDSC3/SSDSC5 terms not restated");
  const int n = KIJL;
  std::vector<double> FACSAT((size_t)n * NFRE), BTH0((size_t)n * NFRE, 0.0), BTH((size_t)n * NANG * NFRE, 0.0), D((size_t)n * NANG * NFRE);
  V2 vFACSAT{FACSAT.data(), n}, vBTH0{BTH0.data(), n};
  V3 vBTH{BTH.data(), n, NANG}, vD{D.data(), n, NANG};
  double TPIINV = 1.0 / t.ZPI;
  double TMP03 = 1.0 / (t.SDSBR * t.MICHE);
  double SSDSC6M1 = 1. - t.SSDSC6;
  for (int M = 1; M <= NFRE; ++M) for (int IJ = KIJS; IJ <= KIJL; ++IJ) vFACSAT(IJ, M) = WAVNUM(IJ, M) * TPIINV * XK2CG(IJ, M);
  for (int M = 1; M <= NFRE; ++M)
    for (int K = 1; K <= NANG; ++K) {
      for (int K2 = 1; K2 <= t.NSDSNTH * 2 + 1; ++K2) {
        int KK = t.INDICESSAT(K, K2);
        for (int IJ = KIJS; IJ <= KIJL; ++IJ) vBTH(IJ, K, M) = vBTH(IJ, K, M) + t.SATWEIGHTS(K, K2) * FL1(IJ, KK, M);
      }
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
        vBTH(IJ, K, M) = vBTH(IJ, K, M) * vFACSAT(IJ, M);
        vBTH0(IJ, M) = std::max(vBTH0(IJ, M), vBTH(IJ, K, M));
      }
    }
  for (int M = 1; M <= NFRE; ++M) {
    double SSDSC2_SIG = t.SSDSC2 * t.ZPIFR(M);
    double ZCOEF = SSDSC2_SIG * t.SSDSC6;
    double ZCOEFM1 = SSDSC2_SIG * SSDSC6M1;
    for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ)
      vD(IJ, K, M) = ZCOEF * sq(std::max(0., vBTH0(IJ, M) * TMP03 - t.SSDSC4)) + ZCOEFM1 * sq(std::max(0., vBTH(IJ, K, M) * TMP03 - t.SSDSC4));  // **IPSAT, IPSAT=2
  }
  for (int M = 1; M <= NFRE; ++M) for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
    SL(IJ, K, M) = SL(IJ, K, M) + vD(IJ, K, M) * FL1(IJ, K, M);
    FLD(IJ, K, M) = FLD(IJ, K, M) + vD(IJ, K, M);
  }
}

// sdissip_jan.F90:96-132
void SDISSIP_JAN(const Ctx& x, V3 FL1, V3 FLD, V3 SL, V2 WAVNUM, V1 EMEAN, V1 F1MEAN, V1 XKMEAN) {
  const Tables& t = x.t;
  const int KIJS = x.KIJS, KIJL = x.KIJL, NANG = x.NANG, NFRE = x.NFRE;
  const int n = KIJL + 1;
  std::vector<double> TEMP1(n), SDS(n), X(n), XK2(n);
  double DELTA_SDISM1 = 1.0 - t.DELTA_SDIS;
  double CONSS = t.CDIS * t.ZPI;
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) SDS[IJ] = CONSS * F1MEAN(IJ) * sq(EMEAN(IJ)) * p4(XKMEAN(IJ));
  for (int M = 1; M <= NFRE; ++M) {
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) { X[IJ] = WAVNUM(IJ, M) / XKMEAN(IJ); XK2[IJ] = sq(WAVNUM(IJ, M)); }
    double CVIS = x.c.rnu * t.CDISVIS;
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) TEMP1[IJ] = SDS[IJ] * X[IJ] * (DELTA_SDISM1 + t.DELTA_SDIS * X[IJ]) + CVIS * XK2[IJ];
    for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      FLD(IJ, K, M) = FLD(IJ, K, M) + TEMP1[IJ];
      SL(IJ, K, M) = SL(IJ, K, M) + TEMP1[IJ] * FL1(IJ, K, M);
    }
  }
}

// transf.F90: ratio of the shallow- to the deep-water narrow-band interaction coefficient (Janssen & Onorato 2007)
static double TRANSF(const Ctx& x, double XK, double D) {
  const Tables& t = x.t;
  const double EPS = 0.0001;
  if (D < x.c.bathymax && D > 0.0) {
    const double X = XK * D;
    if (X > t.DKMAX) return 1.0;
    const double T_0 = std::tanh(X), OM = std::sqrt(t.G * XK * T_0), C_0 = OM / XK;
    const double V_G = X < EPS ? C_0 : 0.5 * C_0 * (1.0 + 2.0 * X / std::sinh(2.0 * X));
    const double T2 = T_0 * T_0;
    const double a = T_0 - X * (1.0 - T2);
    const double DV_G = a * a + 4.0 * (X * X) * T2 * (1.0 - T2);
    const double XNL_1 = (9.0 * (T2 * T2) - 10.0 * T2 + 9.0) / (8.0 * (T2 * T_0));
    const double b = 2.0 * V_G - 0.5 * C_0;
    const double XNL_2 = (b * b / (t.G * D - V_G * V_G) + 1.0) / X;
    const double XNL = XNL_1 - XNL_2;
    return XNL * XNL / (DV_G * ((T2 * T2) * (T2 * T2)));
  }
  return 1.0;
}
// transf_snl.F90:57-100: the same with the directional-width correction of the mean-flow term
static double TRANSF_SNL(const Ctx& x, double XK0, double D, double XNU, double SIG_TH) {
  const Tables& t = x.t;
  const double EPS = 0.0001, XKDMIN = 0.75, TMIN = 0.1, TMAX = 10.0;
  if (D < x.c.bathymax && D > 0.0) {
    double X = XK0 * D;
    if (X > t.DKMAX) return 1.0;
    const double XK = std::max(XK0, XKDMIN / D);
    X = XK * D;
    const double T_0 = std::tanh(X), T_0_SQ = T_0 * T_0, OM = std::sqrt(t.G * XK * T_0), C_0 = OM / XK, C_S_SQ = t.G * D;
    const double V_G = X < EPS ? C_0 : 0.5 * C_0 * (1.0 + 2.0 * X / std::sinh(2.0 * X));
    const double V_G_SQ = V_G * V_G;
    const double a = T_0 - X * (1. - T_0_SQ);
    const double DV_G = a * a + 4.0 * (X * X) * T_0_SQ * (1.0 - T_0_SQ);
    const double XNL_1 = (9.0 * (T_0_SQ * T_0_SQ) - 10.0 * T_0_SQ + 9.0) / (8.0 * T_0_SQ * T_0);
    const double b = 2.0 * V_G - 0.5 * C_0;
    const double XNL_2 = (b * b / (t.G * D - V_G_SQ) + 1.0) / X;
    const double e = 2.0 * C_0 + V_G * (1.0 - T_0_SQ);
    const double XNL_4 = 1. / (4.0 * T_0) * (e * e) / (C_S_SQ - V_G_SQ);
    const double ALP = (1. - V_G_SQ / C_S_SQ) * (C_0 * C_0) / V_G_SQ;
    const double ZFAC = (SIG_TH * SIG_TH) / (SIG_TH * SIG_TH + ALP * (XNU * XNU));
    const double XNL = XNL_1 - XNL_2 + ZFAC * XNL_4;
    const double T4 = (T_0_SQ * T_0_SQ) * (T_0_SQ * T_0_SQ);
    return std::max(std::min(TMAX, XNL * XNL / (DV_G * T4)), TMIN);
  }
  return 1.0;
}
// peak_ang.F90:72-175 (same statements as in orc_output.cpp's kurtosis, on the chunk views of this file)
static void PEAK_ANG(const Ctx& x, V3 FL1, std::vector<double>& XNU, std::vector<double>& SIG_TH) {
  const Tables& t = x.t;
  const int NANG = x.NANG, NFRE = x.NFRE;
  const double ZEPSILON = 10.0 * 2.220446049250313e-16;
  const int NSH = 1 + (int)(std::log(1.5) / std::log(t.FRATIO));
  const double DELT25 = t.WETAIL * t.FR(NFRE) * t.DELTH, COEF_FR = t.WP1TAIL * t.DELTH * (t.FR(NFRE) * t.FR(NFRE));
  const double COEF_FR2 = t.WP2TAIL * t.DELTH * (t.FR(NFRE) * t.FR(NFRE) * t.FR(NFRE));
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
    double SUM0 = ZEPSILON, SUM1 = 0.0, SUM2 = 0.0, TEMP = 0.0;
    for (int M = 1; M <= NFRE; ++M) {
      TEMP = FL1(IJ, 1, M);
      for (int K = 2; K <= NANG; ++K) TEMP = TEMP + FL1(IJ, K, M);
      SUM0 = SUM0 + TEMP * t.DFIM(M); SUM1 = SUM1 + TEMP * t.DFIMFR(M); SUM2 = SUM2 + TEMP * (t.DFIM(M) * (t.FR(M) * t.FR(M)));
    }
    SUM0 = SUM0 + DELT25 * TEMP; SUM1 = SUM1 + COEF_FR * TEMP; SUM2 = SUM2 + COEF_FR2 * TEMP;
    XNU[IJ] = SUM0 > ZEPSILON ? std::sqrt(std::max(ZEPSILON, SUM2 * SUM0 / (SUM1 * SUM1) - 1.0)) : ZEPSILON;
    double XMAX = 0.0;
    int MMAX = 2;
    for (int M = 2; M <= NFRE - 1; ++M) for (int K = 1; K <= NANG; ++K) if (FL1(IJ, K, M) > XMAX) { MMAX = M; XMAX = FL1(IJ, K, M); }
    double S1 = ZEPSILON, S2 = 0.0, SUM_S = 0.0, SUM_C = ZEPSILON, THMEAN = 0.0;
    for (int M = std::max(1, MMAX - NSH); M <= std::min(NFRE, MMAX + NSH); ++M) {
      for (int K = 1; K <= NANG; ++K) { SUM_S = SUM_S + t.SINTH(K) * FL1(IJ, K, M); SUM_C = SUM_C + t.COSTH(K) * FL1(IJ, K, M); }
      THMEAN = std::atan2(SUM_S, SUM_C);
      for (int K = 1; K <= NANG; ++K) { S1 = S1 + FL1(IJ, K, M) * t.DFIM(M); S2 = S2 + std::cos(t.TH(K) - THMEAN) * FL1(IJ, K, M) * t.DFIM(M); }
    }
    SIG_TH[IJ] = S1 > ZEPSILON ? std::sqrt(2.0 * (1.0 - S2 / S1)) : 0.0;
  }
}

// snonlin.F90:116-498
void SNONLIN(const Ctx& x, V3 FL1, V3 FLD, V3 SL, V2 WAVNUM, V1 DEPTH, V1 AKMEAN) {
  const Tables& t = x.t;
  const int KIJS = x.KIJS, KIJL = x.KIJL, NANG = x.NANG, NFRE = x.NFRE;
  const int n = KIJL + 1;
  std::vector<double> FTEMP(n), AD(n), DELAD(n), DELAP(n), DELAM(n), ENHFR(n);
  std::vector<double> ENH((size_t)n * t.MLSTHG);
  auto enh = [&](int ij, int mc) -> double& { return ENH[ij + (size_t)n * (mc - 1)]; };
  const double ENH_MAX = 10.0, ENH_MIN = 0.1;   // snonlin.F90:101-102
  auto xk_above = [&](int MC) {   // XK = GM1*(ZPIFR(NFRE)*FRATIO**(MC-NFRE))**2 (:145, :157), integer power by repeated multiplication
    double pw = 1.0;
    for (int i = 0; i < MC - NFRE; ++i) pw = pw * t.FRATIO;
    const double w = t.ZPIFR(NFRE) * pw;
    return t.GM1 * (w * w);
  };
  if (x.c.isnonlin == 0) {
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      ENHFR[IJ] = std::max(0.75 * DEPTH(IJ) * AKMEAN(IJ), 0.5);
      ENHFR[IJ] = 1.0 + (5.5 / ENHFR[IJ]) * (1.0 - .833 * ENHFR[IJ]) * std::exp(-1.25 * ENHFR[IJ]);
    }
    for (int MC = 1; MC <= t.MLSTHG; ++MC) for (int IJ = KIJS; IJ <= KIJL; ++IJ) enh(IJ, MC) = ENHFR[IJ];
  } else if (x.c.isnonlin == 1) {   // snonlin.F90:138-150 (oracle only so far: the product rejects ISNONLIN /= 0)
    for (int MC = 1; MC <= NFRE; ++MC) for (int IJ = KIJS; IJ <= KIJL; ++IJ)
      enh(IJ, MC) = std::max(std::min(ENH_MAX, TRANSF(x, WAVNUM(IJ, MC), DEPTH(IJ))), ENH_MIN);
    for (int MC = NFRE + 1; MC <= t.MLSTHG; ++MC) for (int IJ = KIJS; IJ <= KIJL; ++IJ)
      enh(IJ, MC) = std::max(std::min(ENH_MAX, TRANSF(x, xk_above(MC), DEPTH(IJ))), ENH_MIN);
  } else if (x.c.isnonlin == 2) {   // snonlin.F90:151-163
    std::vector<double> XNU(n), SIG_TH(n);
    PEAK_ANG(x, FL1, XNU, SIG_TH);
    for (int MC = 1; MC <= NFRE; ++MC) for (int IJ = KIJS; IJ <= KIJL; ++IJ) enh(IJ, MC) = TRANSF_SNL(x, WAVNUM(IJ, MC), DEPTH(IJ), XNU[IJ], SIG_TH[IJ]);
    for (int MC = NFRE + 1; MC <= t.MLSTHG; ++MC) for (int IJ = KIJS; IJ <= KIJL; ++IJ)
      enh(IJ, MC) = TRANSF_SNL(x, xk_above(MC), DEPTH(IJ), XNU[IJ], SIG_TH[IJ]);
  } else throw std::runtime_error("SNONLIN: ISNONLIN must be 0, 1 or 2");
  int MFR1STFR = -t.MFRSTLW + 1;
  int MFRLSTFR = NFRE - t.KFRH + MFR1STFR;
  for (int MC = 1; MC <= t.MLSTHG; ++MC) {
    int MP = t.IKP(MC), MP1 = t.IKP1(MC), MM = t.IKM(MC), MM1 = t.IKM1(MC);
    int IC = t.INLCOEF(1, MC), IP = t.INLCOEF(2, MC), IP1 = t.INLCOEF(3, MC), IM = t.INLCOEF(4, MC), IM1 = t.INLCOEF(5, MC);
    double FTAIL = t.RNLCOEF(1, MC);
    double GW1 = t.RNLCOEF(2, MC), GW2 = t.RNLCOEF(3, MC), GW3 = t.RNLCOEF(4, MC), GW4 = t.RNLCOEF(5, MC);
    double FKLAMPA = t.RNLCOEF(6, MC), FKLAMPB = t.RNLCOEF(7, MC), FKLAMP2 = t.RNLCOEF(8, MC), FKLAMP1 = t.RNLCOEF(9, MC);
    double FKLAPA2 = t.RNLCOEF(10, MC), FKLAPB2 = t.RNLCOEF(11, MC), FKLAP12 = t.RNLCOEF(12, MC), FKLAP22 = t.RNLCOEF(13, MC);
    double GW5 = t.RNLCOEF(14, MC), GW6 = t.RNLCOEF(15, MC), GW7 = t.RNLCOEF(16, MC), GW8 = t.RNLCOEF(17, MC);
    double FKLAMMA = t.RNLCOEF(18, MC), FKLAMMB = t.RNLCOEF(19, MC), FKLAMM2 = t.RNLCOEF(20, MC), FKLAMM1 = t.RNLCOEF(21, MC);
    double FKLAMA2 = t.RNLCOEF(22, MC), FKLAMB2 = t.RNLCOEF(23, MC), FKLAM12 = t.RNLCOEF(24, MC), FKLAM22 = t.RNLCOEF(25, MC);
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) FTEMP[IJ] = t.AF11(MC) * enh(IJ, MC);
    const int branch = (MC > MFR1STFR && MC < MFRLSTFR) ? 0 : (MC >= MFRLSTFR ? 1 : 2);
    for (int KH = 1; KH <= 2; ++KH) {
      for (int K = 1; K <= NANG; ++K) {
        int K1 = t.K1W(K, KH), K2 = t.K2W(K, KH), K11 = t.K11W(K, KH), K21 = t.K21W(K, KH);
        for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
          double SAP = GW1 * FL1(IJ, K1, IP) + GW2 * FL1(IJ, K11, IP) + GW3 * FL1(IJ, K1, IP1) + GW4 * FL1(IJ, K11, IP1);
          double SAM = GW5 * FL1(IJ, K2, IM) + GW6 * FL1(IJ, K21, IM) + GW7 * FL1(IJ, K2, IM1) + GW8 * FL1(IJ, K21, IM1);
          double FIJ = (branch == 0) ? FL1(IJ, K, IC) : FL1(IJ, K, IC) * FTAIL;
          double FAD1 = FIJ * (SAP + SAM);
          double FAD2 = FAD1 - 2.0 * SAP * SAM;
          FAD1 = FAD1 + FAD2;
          double FCEN = FTEMP[IJ] * FIJ;
          AD[IJ] = FAD2 * FCEN;
          DELAD[IJ] = FAD1 * FTEMP[IJ];
          DELAP[IJ] = (FIJ - 2.0 * SAM) * t.DAL1 * FCEN;
          DELAM[IJ] = (FIJ - 2.0 * SAP) * t.DAL2 * FCEN;
        }
        auto centre = [&]() { for (int IJ = KIJS; IJ <= KIJL; ++IJ) { SL(IJ, K, MC) = SL(IJ, K, MC) - 2.0 * AD[IJ]; FLD(IJ, K, MC) = FLD(IJ, K, MC) - 2.0 * DELAD[IJ]; } };
        auto mm = [&]() { for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
          SL(IJ, K2, MM) = SL(IJ, K2, MM) + AD[IJ] * FKLAMM1; FLD(IJ, K2, MM) = FLD(IJ, K2, MM) + DELAM[IJ] * FKLAM12;
          SL(IJ, K21, MM) = SL(IJ, K21, MM) + AD[IJ] * FKLAMM2; FLD(IJ, K21, MM) = FLD(IJ, K21, MM) + DELAM[IJ] * FKLAM22; } };
        auto mm1 = [&]() { for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
          SL(IJ, K2, MM1) = SL(IJ, K2, MM1) + AD[IJ] * FKLAMMA; FLD(IJ, K2, MM1) = FLD(IJ, K2, MM1) + DELAM[IJ] * FKLAMA2;
          SL(IJ, K21, MM1) = SL(IJ, K21, MM1) + AD[IJ] * FKLAMMB; FLD(IJ, K21, MM1) = FLD(IJ, K21, MM1) + DELAM[IJ] * FKLAMB2; } };
        auto mp = [&]() { for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
          SL(IJ, K1, MP) = SL(IJ, K1, MP) + AD[IJ] * FKLAMP1; FLD(IJ, K1, MP) = FLD(IJ, K1, MP) + DELAP[IJ] * FKLAP12;
          SL(IJ, K11, MP) = SL(IJ, K11, MP) + AD[IJ] * FKLAMP2; FLD(IJ, K11, MP) = FLD(IJ, K11, MP) + DELAP[IJ] * FKLAP22; } };
        auto mp1 = [&]() { for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
          SL(IJ, K1, MP1) = SL(IJ, K1, MP1) + AD[IJ] * FKLAMPA; FLD(IJ, K1, MP1) = FLD(IJ, K1, MP1) + DELAP[IJ] * FKLAPA2;
          SL(IJ, K11, MP1) = SL(IJ, K11, MP1) + AD[IJ] * FKLAMPB; FLD(IJ, K11, MP1) = FLD(IJ, K11, MP1) + DELAP[IJ] * FKLAPB2; } };
        if (branch == 0) {          // :253-308
          centre(); mm(); mm1(); mp(); mp1();
        } else if (branch == 1) {   // :312-410
          mm();
          if (MM1 <= NFRE) {
            mm1();
            if (MC <= NFRE) {
              centre();
              if (MP <= NFRE) { mp(); if (MP1 <= NFRE) mp1(); }
            }
          }
        } else {                    // :414-490
          if (MM1 >= 1) mm1();
          centre(); mp(); mp1();
        }
      }
    }
  }
}

// sdiwbk.F90:69-117
void SDIWBK(const Ctx& x, V3 FL1, V3 FLD, V3 SL, V1 DEPTH, V1 EMAXDPT, V1 EMEAN, V1 F1MEAN) {
  if (!x.c.lbiwbk) return;
  const int KIJS = x.KIJS, KIJL = x.KIJL;
  const double ALPH_B_J = 1.0, COEF_B_J = 2 * ALPH_B_J, DEPTHTRS = 50.0;
  std::vector<double> SDS(KIJL + 1, 0.0);
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
    if (DEPTH(IJ) < DEPTHTRS) {
      double ALPH = 2.0 * EMAXDPT(IJ) / EMEAN(IJ);
      double ARG = std::min(ALPH, 50.0);
      double Q_OLD = std::exp(-ARG), Q = Q_OLD;
      for (int IC = 1; IC <= 15; ++IC) {
        double EXPQ = std::exp(-ARG * (1.0 - Q_OLD));
        Q = Q_OLD - (EXPQ - Q_OLD) / (ARG * EXPQ - 1.0);
        double REL_ERR = std::fabs(Q - Q_OLD) / Q_OLD;
        if (REL_ERR < 0.00001) break;
        Q_OLD = Q;
      }
      Q = std::min(Q, 1.0);
      SDS[IJ] = COEF_B_J * ALPH * Q * F1MEAN(IJ);
    }
  }
  for (int M = 1; M <= x.c.nfre_red; ++M) for (int K = 1; K <= x.NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ)
    if (DEPTH(IJ) < DEPTHTRS) { SL(IJ, K, M) = SL(IJ, K, M) - SDS[IJ] * FL1(IJ, K, M); FLD(IJ, K, M) = FLD(IJ, K, M) - SDS[IJ]; }
}

// sbottom.F90:76-97
void SBOTTOM(const Ctx& x, V3 FL1, V3 FLD, V3 SL, V2 WAVNUM, V1 DEPTH) {
  const int KIJS = x.KIJS, KIJL = x.KIJL;
  double CONST = -2.0 * 0.038 * x.t.GM1;
  std::vector<double> SBO(KIJL + 1);
  for (int M = 1; M <= x.c.nfre_red; ++M) {
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      if (DEPTH(IJ) < x.c.bathymax) {
        double ARG = 2.0 * DEPTH(IJ) * WAVNUM(IJ, M);
        ARG = std::min(ARG, 50.0);
        SBO[IJ] = CONST * WAVNUM(IJ, M) / std::sinh(ARG);
      } else SBO[IJ] = 0.0;
    }
    for (int K = 1; K <= x.NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) { SL(IJ, K, M) = SL(IJ, K, M) + SBO[IJ] * FL1(IJ, K, M); FLD(IJ, K, M) = FLD(IJ, K, M) + SBO[IJ]; }
  }
}

// wnfluxes.F90:163-331 (LWNEMOCOU=F, LWNEMOCOUWRS=F, LCIWA*=F)
// the NEMO fields WNFLUXES updates (wnfluxes.F90:304-330)
struct NemoFlux { V1 NPHIEPS, NTAUOC, NSWH, NMWP, NEMOTAUX, NEMOTAUY, NEMOTAUICX, NEMOTAUICY, NEMOWSWAVE, NEMOPHIF; };
void WNFLUXES(const Ctx& x, I1 MIJ, V2 RHOWGDFTH, V2 CINV, V3 SSURF, V1 CICOVER, V1 PHIWA, V1 EM, V1 F1, V1 WSWAVE,
              V1 WDWAVE, V1 USTRA, V1 VSTRA, V1 UFRIC, V1 AIRD, V1 TAUXD, V1 TAUYD, V1 TAUOCXD, V1 TAUOCYD, V1 TAUOC,
              V1 TAUICX, V1 TAUICY, V1 PHIOCD, V1 PHIEPS, V1 PHIAW, const NemoFlux* NE = nullptr, const V3* SLICE = nullptr) {
  (void)MIJ;
  const Tables& t = x.t;
  const Config& c = x.c;
  const int KIJS = x.KIJS, KIJL = x.KIJL, NANG = x.NANG, NFRE = x.NFRE;
  const int n = KIJL + 1;
  const double PHIOC_ICE = -3.75, PHIAW_ICE = 3.75, C1 = 1.03e-3, C2 = 0.04e-3, P1 = 1.48, P2 = -0.21, CDMAX_LOC = 0.003,
               EFD_MIN = 0.0625, EFD_MAX = 6.25;
  std::vector<double> XSTRESS(n, 0.0), YSTRESS(n, 0.0), USTAR(n), PHILF(n, 0.0), OOVAL(n), EM_OC(n), F1_OC(n), SUMT(n), SUMX(n), SUMY(n);
  double EPSUS3 = t.EPSUS * std::sqrt(t.EPSUS);
  double ZCITHRS = c.ciblock, CITHRSH_INV = 1. / std::max(c.cithrsh, 0.01), ZMAXEXP = 10.;
  if (c.lciwa1 || c.lciwa2 || c.lciwa3) { ZCITHRS = 0.; CITHRSH_INV = 50.; ZMAXEXP = 20.; }   // wnfluxes.F90:150-158
  double EFD_FAC = 4.0 * t.EGRCRV / (t.G * t.G);
  double FFD_FAC = std::pow(t.EGRCRV / t.AFCRV, 1.0 / t.BFCRV) * t.G;
  for (int M = 1; M <= NFRE; ++M) {
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) { SUMT[IJ] = SSURF(IJ, 1, M); SUMX[IJ] = t.SINTH(1) * SSURF(IJ, 1, M); SUMY[IJ] = t.COSTH(1) * SSURF(IJ, 1, M); }
    for (int K = 2; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      SUMT[IJ] = SUMT[IJ] + SSURF(IJ, K, M);
      SUMX[IJ] = SUMX[IJ] + t.SINTH(K) * SSURF(IJ, K, M);
      SUMY[IJ] = SUMY[IJ] + t.COSTH(K) * SSURF(IJ, K, M);
    }
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      PHILF[IJ] = PHILF[IJ] + SUMT[IJ] * RHOWGDFTH(IJ, M);
      double CMRHOWGDFTH = CINV(IJ, M) * RHOWGDFTH(IJ, M);
      XSTRESS[IJ] = XSTRESS[IJ] + SUMX[IJ] * CMRHOWGDFTH;
      YSTRESS[IJ] = YSTRESS[IJ] + SUMY[IJ] * CMRHOWGDFTH;
    }
  }
  if (c.licerun && c.lwamrsetci) {
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      if (CICOVER(IJ) > ZCITHRS) {
        OOVAL[IJ] = std::exp(-std::min(p4(CICOVER(IJ) * CITHRSH_INV), ZMAXEXP));
        double U10P = std::max(WSWAVE(IJ), t.EPSU10);
        double CD_BULK = std::min((C1 + C2 * std::pow(U10P, P1)) * std::pow(U10P, P2), CDMAX_LOC);
        double CD_WAVE = sq(UFRIC(IJ) / U10P);
        double CD_ICE = OOVAL[IJ] * CD_WAVE + (1.0 - OOVAL[IJ]) * CD_BULK;
        USTAR[IJ] = std::max(std::sqrt(CD_ICE) * U10P, t.EPSUS);
        double EFD = std::min(EFD_FAC * p4(USTAR[IJ]), EFD_MAX);
        EM_OC[IJ] = std::max(OOVAL[IJ] * EM(IJ) + (1.0 - OOVAL[IJ]) * EFD, EFD_MIN);
        double FFD = FFD_FAC / USTAR[IJ];
        F1_OC[IJ] = OOVAL[IJ] * F1(IJ) + (1.0 - OOVAL[IJ]) * FFD;
        F1_OC[IJ] = std::min(std::max(F1_OC[IJ], t.FR(2)), t.FR(NFRE));
      } else { OOVAL[IJ] = 1.0; USTAR[IJ] = UFRIC(IJ); EM_OC[IJ] = EM(IJ); F1_OC[IJ] = F1(IJ); }
    }
  } else {
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) { OOVAL[IJ] = 1.0; USTAR[IJ] = UFRIC(IJ); EM_OC[IJ] = EM(IJ); F1_OC[IJ] = F1(IJ); }
  }
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
    double TAU = AIRD(IJ) * std::max(sq(USTAR[IJ]), t.EPSUS);
    TAUXD(IJ) = TAU * std::sin(WDWAVE(IJ));
    TAUYD(IJ) = TAU * std::cos(WDWAVE(IJ));
    TAUOCXD(IJ) = TAUXD(IJ) - OOVAL[IJ] * XSTRESS[IJ];
    TAUOCYD(IJ) = TAUYD(IJ) - OOVAL[IJ] * YSTRESS[IJ];
    double TAUO = std::sqrt(sq(TAUOCXD(IJ)) + sq(TAUOCYD(IJ)));
    TAUOC(IJ) = std::min(std::max(TAUO / TAU, t.TAUOCMIN), t.TAUOCMAX);
  }
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) { TAUICX(IJ) = 0.0; TAUICY(IJ) = 0.0; }
  if (c.lwnemocouwrs && SLICE) {   // wave radiative stress on the sea ice (wnfluxes.F90:178-196, 266-271): integrated up to FR(NFRE)
    const double EPSMIN1000 = t.EPSMIN * 1000.0;
    std::vector<double> XSI(n, 0.0), YSI(n, 0.0), SX(n), SY(n);
    for (int M = 1; M <= NFRE; ++M) {
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) { SX[IJ] = t.SINTH(1) * std::min((*SLICE)(IJ, 1, M), -EPSMIN1000); SY[IJ] = t.COSTH(1) * std::min((*SLICE)(IJ, 1, M), -EPSMIN1000); }
      for (int K = 2; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
        SX[IJ] = SX[IJ] + t.SINTH(K) * std::min((*SLICE)(IJ, K, M), -EPSMIN1000);
        SY[IJ] = SY[IJ] + t.COSTH(K) * std::min((*SLICE)(IJ, K, M), -EPSMIN1000);
      }
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
        XSI[IJ] = XSI[IJ] + c.zalpwrs * SX[IJ] * CINV(IJ, M) * t.RHOWG_DFIM(M);
        YSI[IJ] = YSI[IJ] + c.zalpwrs * SY[IJ] * CINV(IJ, M) * t.RHOWG_DFIM(M);
      }
    }
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) { TAUICX(IJ) = -XSI[IJ]; TAUICY(IJ) = -YSI[IJ]; }
  }
  if (c.lwcouast)
    for (int IJ = KIJS; IJ <= KIJL; ++IJ)
      if (USTRA(IJ) != 0.0 || VSTRA(IJ) != 0.0) {
        TAUXD(IJ) = USTRA(IJ); TAUOCXD(IJ) = USTRA(IJ) * TAUOC(IJ); TAUYD(IJ) = VSTRA(IJ); TAUOCYD(IJ) = VSTRA(IJ) * TAUOC(IJ);
      }
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
    double XN = AIRD(IJ) * std::max(USTAR[IJ] * USTAR[IJ] * USTAR[IJ], EPSUS3);
    PHIOCD(IJ) = OOVAL[IJ] * (PHILF[IJ] - PHIWA(IJ)) + (1.0 - OOVAL[IJ]) * PHIOC_ICE * XN;
    PHIEPS(IJ) = PHIOCD(IJ) / XN;
    PHIEPS(IJ) = std::min(std::max(PHIEPS(IJ), t.PHIEPSMIN), t.PHIEPSMAX);
    PHIOCD(IJ) = PHIEPS(IJ) * XN;
    PHIAW(IJ) = PHIWA(IJ) / XN;
    PHIAW(IJ) = OOVAL[IJ] * PHIWA(IJ) / XN + (1.0 - OOVAL[IJ]) * PHIAW_ICE;
  }
  if (c.lwnemocou && NE) {   // LNUPD = T (implsch.F90:413); wnfluxes.F90:304-330
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      NE->NPHIEPS(IJ) = PHIEPS(IJ);
      NE->NTAUOC(IJ) = TAUOC(IJ);
      NE->NSWH(IJ) = EM_OC[IJ] != 0.0 ? 4.0 * std::sqrt(EM_OC[IJ]) : 0.0;
      NE->NMWP(IJ) = F1_OC[IJ] != 0.0 ? 1.0 / F1_OC[IJ] : 0.0;
      if (c.lwnemotauoc) { NE->NEMOTAUX(IJ) = NE->NEMOTAUX(IJ) + TAUOCXD(IJ); NE->NEMOTAUY(IJ) = NE->NEMOTAUY(IJ) + TAUOCYD(IJ); }
      else { NE->NEMOTAUX(IJ) = NE->NEMOTAUX(IJ) + TAUXD(IJ); NE->NEMOTAUY(IJ) = NE->NEMOTAUY(IJ) + TAUYD(IJ); }
      NE->NEMOWSWAVE(IJ) = NE->NEMOWSWAVE(IJ) + WSWAVE(IJ);
      NE->NEMOPHIF(IJ) = NE->NEMOPHIF(IJ) + PHIOCD(IJ);
      NE->NEMOTAUICX(IJ) = NE->NEMOTAUICX(IJ) + TAUICX(IJ);
      NE->NEMOTAUICY(IJ) = NE->NEMOTAUICY(IJ) + TAUICY(IJ);
    }
  }
}

// aki_ice.F90:66-112: wavenumber of a flexural-gravity wave under an elastic ice sheet (Fox & Squire 1991), Newton's method
static double AKI_ICE(double G, double XK, double DEPTH, double RHOW, double CITH) {
  const double YMICE = 5.5E+9, RMUICE = 0.3, RHOI = 922.5, EBS = 0.000001, AKI_MAX = 20.0;
  double AKI;
  if (CITH <= 0.0) AKI = XK;
  else {
    const double FICSTF = (YMICE * CITH * CITH * CITH / (12 * (1 - RMUICE * RMUICE))) / RHOW;
    const double RDH = (RHOI / RHOW) * CITH;
    const double OM2 = G * XK * std::tanh(XK * DEPTH);
    double AKIOLD = 0.0;
    AKI = std::min(XK, std::pow(OM2 / std::max(FICSTF, 1.0), 0.2));
    while (std::fabs(AKI - AKIOLD) > EBS * AKIOLD && AKI < AKI_MAX) {
      AKIOLD = AKI;
      const double AKID = std::min(DEPTH * AKI, 50.0);
      const double F = FICSTF * std::pow(AKI, 5) + G * AKI - OM2 * (RDH * AKI + 1. / std::tanh(AKID));
      const double FPRIME = 5. * FICSTF * p4(AKI) + G - OM2 * (RDH - DEPTH / sq(std::sinh(AKID)));
      AKI = AKI - F / FPRIME;
      if (AKI <= 0.0) AKI = AKI_MAX;
    }
  }
  return AKI;
}

// cimsstrn.F90:83-119: mean square wave strain in the sea ice
void CIMSSTRN(const Ctx& x, V3 FL1, V2 WAVNUM, V1 DEPTH, V1 CITHICK, V1 STRN) {
  const Tables& t = x.t;
  const int KIJS = x.KIJS, KIJL = x.KIJL, NANG = x.NANG, NFRE = x.NFRE;
  const double F1LIM = x.c.flmin / t.DELTH;
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) STRN(IJ) = 0.0;
  for (int M = 1; M <= NFRE; ++M)
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      const double XKI = AKI_ICE(t.G, WAVNUM(IJ, M), DEPTH(IJ), t.ROWATER, CITHICK(IJ));
      const double E = 0.5 * CITHICK(IJ) * XKI * XKI * XKI / WAVNUM(IJ, M);
      double SUME = 0.0;
      for (int K = 1; K <= NANG; ++K) SUME = SUME + FL1(IJ, K, M);
      if (SUME > F1LIM) STRN(IJ) = STRN(IJ) + E * E * SUME * t.DFIM(M);
    }
}

// imphftail.F90:71-87
void IMPHFTAIL(const Ctx& x, I1 MIJ, V2 FLM, V2 WAVNUM, V2 XK2CG, V3 FL1) {
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
    double TEMP1 = 1.0 / XK2CG(IJ, MIJ(IJ)) / WAVNUM(IJ, MIJ(IJ));
    for (int M = MIJ(IJ) + 1; M <= x.NFRE; ++M) {
      double TEMP2 = 1.0 / XK2CG(IJ, M) / WAVNUM(IJ, M);
      TEMP2 = TEMP2 / TEMP1;
      for (int K = 1; K <= x.NANG; ++K) { double TFAC = FL1(IJ, K, MIJ(IJ)); FL1(IJ, K, M) = std::max(TEMP2 * TFAC, FLM(IJ, K)); }
    }
  }
}

// setice.F90:64-86
void SETICE(const Ctx& x, V3 FL1, V1 CICOVER, V2 COSWDIF) {
  const int n = x.KIJL + 1;
  std::vector<double> CIREDUC(n), TEMP(n), ICEFREE(n);
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
    if (CICOVER(IJ) > x.c.cithrsh) { CIREDUC[IJ] = std::max(x.t.EPSMIN, (1.0 - CICOVER(IJ))); ICEFREE[IJ] = 0.0; }
    else { CIREDUC[IJ] = 0.0; ICEFREE[IJ] = 1.0; }
  }
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) TEMP[IJ] = CIREDUC[IJ] * x.c.flmin;
  for (int M = 1; M <= x.NFRE; ++M) for (int K = 1; K <= x.NANG; ++K) for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ)
    FL1(IJ, K, M) = FL1(IJ, K, M) * ICEFREE[IJ] + TEMP[IJ] * sq(std::max(0.0, COSWDIF(IJ, K)));
}

// stokesdrift.F90:84-142
void STOKESDRIFT(const Ctx& x, V3 FL1, V2 STOKFAC, V1 WSWAVE, V1 WDWAVE, V1 CICOVER, V1 USTOKES, V1 VSTOKES) {
  const Tables& t = x.t;
  const double STMAX = 1.5;
  std::vector<double> STFAC(x.KIJL + 1);
  double CONST = 2.0 * t.DELTH * t.ZPI * t.ZPI * t.ZPI / t.G * p4(t.FR(t.NFRE_ODD));
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) { USTOKES(IJ) = 0.0; VSTOKES(IJ) = 0.0; }
  for (int M = 1; M <= t.NFRE_ODD; ++M) {
    for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) STFAC[IJ] = STOKFAC(IJ, M) * t.DFIM_SIM(M);
    for (int K = 1; K <= x.NANG; ++K) for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
      double FAC3 = STFAC[IJ] * FL1(IJ, K, M);
      USTOKES(IJ) = USTOKES(IJ) + FAC3 * t.SINTH(K);
      VSTOKES(IJ) = VSTOKES(IJ) + FAC3 * t.COSTH(K);
    }
  }
  for (int K = 1; K <= x.NANG; ++K) {
    double FAC1 = CONST * t.SINTH(K), FAC2 = CONST * t.COSTH(K);
    for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) { USTOKES(IJ) = USTOKES(IJ) + FAC1 * FL1(IJ, K, t.NFRE_ODD); VSTOKES(IJ) = VSTOKES(IJ) + FAC2 * FL1(IJ, K, t.NFRE_ODD); }
  }
  if (x.c.licerun && x.c.lwamrsetci)
    for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ)
      if (CICOVER(IJ) > x.c.cithrsh) {
        USTOKES(IJ) = 0.016 * WSWAVE(IJ) * std::sin(WDWAVE(IJ)) * (1.0 - CICOVER(IJ));
        VSTOKES(IJ) = 0.016 * WSWAVE(IJ) * std::cos(WDWAVE(IJ)) * (1.0 - CICOVER(IJ));
      }
  for (int IJ = x.KIJS; IJ <= x.KIJL; ++IJ) {
    USTOKES(IJ) = std::min(std::max(USTOKES(IJ), -STMAX), STMAX);
    VSTOKES(IJ) = std::min(std::max(VSTOKES(IJ), -STMAX), STMAX);
  }
}

}  // namespace

// femean.F90 (outblock.F90:223-244 uses it for Hs / mean period)
void femean(const Tables& t, const Config& c, int KIJL, const double* Fp, double* EM, double* FM) {
  V3 F{const_cast<double*>(Fp), KIJL, c.nang};
  std::vector<double> TEMP2(KIJL + 1);
  for (int IJ = 1; IJ <= KIJL; ++IJ) { EM[IJ - 1] = 0.0; FM[IJ - 1] = 0.0; }   // femean.F90:84-87
  double DELT25 = t.WETAIL * t.FR(c.nfre) * t.DELTH;
  double DELT2 = t.FRTAIL * t.DELTH;
  for (int M = 1; M <= c.nfre; ++M) {
    for (int IJ = 1; IJ <= KIJL; ++IJ) TEMP2[IJ] = std::max(F(IJ, 1, M), t.EPSMIN);
    for (int K = 2; K <= c.nang; ++K) for (int IJ = 1; IJ <= KIJL; ++IJ) TEMP2[IJ] = TEMP2[IJ] + std::max(F(IJ, K, M), t.EPSMIN);
    for (int IJ = 1; IJ <= KIJL; ++IJ) { EM[IJ - 1] = EM[IJ - 1] + TEMP2[IJ] * t.DFIM(M); FM[IJ - 1] = FM[IJ - 1] + t.DFIMOFR(M) * TEMP2[IJ]; }
  }
  for (int IJ = 1; IJ <= KIJL; ++IJ) {
    EM[IJ - 1] = EM[IJ - 1] + DELT25 * TEMP2[IJ];
    FM[IJ - 1] = FM[IJ - 1] + DELT2 * TEMP2[IJ];
    FM[IJ - 1] = EM[IJ - 1] / FM[IJ - 1];
    FM[IJ - 1] = std::max(FM[IJ - 1], t.FR(1));
  }
}

// halphap.F90:68-115 with meansqs_lf.F90:80-100 (NFRE_EFF = NFRE) and FEMEAN on the spectrum in the wind direction
void HALPHAP(const Ctx& x, V2 WAVNUM, V2 COSWDIF, V3 FL1, V1 HALP) {
  const Tables& t = x.t;
  const int KIJS = x.KIJS, KIJL = x.KIJL, NANG = x.NANG, NFRE = x.NFRE;
  const double ZLNFRNFRE = std::log(t.FR(NFRE));
  std::vector<double> sFLWD((size_t)KIJL * NANG * NFRE), XMSS(KIJL + 1, 0.0), EM(KIJL), FM(KIJL);
  V3 FLWD{sFLWD.data(), KIJL, NANG};
  for (int M = 1; M <= NFRE; ++M) for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
    const double WD = 0.5 + 0.5 * std::copysign(1.0, COSWDIF(IJ, K));
    FLWD(IJ, K, M) = FL1(IJ, K, M) * WD;
  }
  for (int M = 1; M <= NFRE; ++M)
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      const double TEMP1 = t.DFIM(M) * sq(WAVNUM(IJ, M));
      double TEMP2 = 0.0;
      for (int K = 1; K <= NANG; ++K) TEMP2 = TEMP2 + FLWD(IJ, K, M);
      XMSS[IJ] = XMSS[IJ] + TEMP1 * TEMP2;
    }
  femean(t, x.c, KIJL, sFLWD.data(), EM.data(), FM.data());
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
    double ALPHAP;
    bool tail = true;
    if (EM[IJ - 1] > 0.0 && FM[IJ - 1] < t.FR(NFRE - 2)) {
      ALPHAP = XMSS[IJ] / (ZLNFRNFRE - std::log(FM[IJ - 1]));
      tail = ALPHAP > t.ALPHAPMAX;
    }
    if (tail) {
      double F1D = 0.0;
      for (int K = 1; K <= NANG; ++K) F1D = F1D + FLWD(IJ, K, NFRE) * t.DELTH;
      ALPHAP = t.ZPI4GM2 * t.FR5(NFRE) * F1D;
    }
    HALP(IJ) = 0.5 * std::min(ALPHAP, t.ALPHAPMAX);
  }
}

// sdice.F90:99-112 with sdice1.F90:102-187, sdice2.F90:97-121, sdice3.F90:107-147
void SDICE(const Ctx& x, V3 FL1, V3 FLD, V3 SL, V2 WAVNUM, V2 CGROUP, V1 CICOVER, V1 CITHICK, const double* ALPFAC /*1-based, null: ZALPFACX*/,
           V3* SLICEO /*null: not kept*/) {
  const double DELT5 = x.c.ximp * x.c.idelt;
  const Tables& t = x.t;
  const Config& c = x.c;
  const int KIJS = x.KIJS, KIJL = x.KIJL, NANG = x.NANG, NFRE = x.NFRE;
  // SDICE (sdice.F90:99-112): type 1 scattering, type 2 under-ice friction, type 3 viscous friction, in this order.  SLICE (the
  // attenuation's own source function, INTENT(OUT) of every term: the last one that is on wins) feeds the LWNEMOCOUWRS radiative
  // stress of WNFLUXES (wnfluxes.F90:178-196).
  if (c.lciwa1) {    // SDICE1 (sdice1.F90:102-187)
    const double CIFRGL = 0.955, CIDMIN = 20.0, CIFRGMT = 2.0, A = 200.0, Cc = 300.0;
    const int MAXICM = (int)(std::log(A / CIDMIN) / std::log(CIFRGMT));
    std::vector<double> DINV(KIJL + 1), ALP((size_t)(KIJL + 1) * (NFRE + 1));
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      if (CITHICK(IJ) > 0.0) {
        const double CIDMAX = A + Cc * CICOVER(IJ);
        const int ICM = std::min((int)(std::log(CIDMAX / CIDMIN) / std::log(CIFRGMT)), MAXICM);
        double SN = 0.0, SD = 0.0;
        for (int I = 0; I <= ICM; ++I) {
          const double X = std::pow(CIFRGMT * CIFRGMT * CIFRGL, I);
          SN = SN + X * CIDMAX / std::pow(CIFRGMT, I);
          SD = SD + X;
        }
        const double CIDMEAN = SN / SD;
        DINV[IJ] = 1.0 / CIDMEAN;
      } else DINV[IJ] = CIDMIN;
    }
    for (int M = 1; M <= NFRE; ++M) {
      const double TW = 1.0 / t.FR(M);
      int IT = (int)std::floor((TW - t.TICMIN) / t.DTIC + 1);
      IT = std::max(1, std::min(IT, t.NICT));
      int IT1 = IT + 1;
      IT1 = std::max(1, std::min(IT1, t.NICT));
      const double WT1 = std::max(std::min(1.0, (TW - (t.TICMIN + (IT - 1) * t.DTIC)) / t.DTIC), 0.0);
      const double WT = 1.0 - WT1;
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
        if (CITHICK(IJ) > 0.0) {
          int IH = (int)std::floor((CITHICK(IJ) - t.HICMIN) / t.DHIC + 1);
          IH = std::max(1, std::min(IH, t.NICH));
          int IH1 = IH + 1;
          IH1 = std::max(1, std::min(IH1, t.NICH));
          const double WH1 = std::max(std::min(1., (CITHICK(IJ) - (t.HICMIN + (IH - 1) * t.DHIC)) / t.DHIC), 0.0);
          const double WH = 1.0 - WH1;
          const double CIDEAC_INT = WT * (WH * t.CIDEAC(IT, IH) + WH1 * t.CIDEAC(IT, IH1)) +
                                    WT1 * (WH * t.CIDEAC(IT1, IH) + WH1 * t.CIDEAC(IT1, IH1));
          ALP[(size_t)M * (KIJL + 1) + IJ] = std::exp(CIDEAC_INT) * DINV[IJ] * c.zalpfacb;
        } else ALP[(size_t)M * (KIJL + 1) + IJ] = 0.0;
      }
    }
    for (int M = 1; M <= NFRE; ++M) for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      const double FLDICE = -ALP[(size_t)M * (KIJL + 1) + IJ] * CGROUP(IJ, M);
      const double SLICE = FL1(IJ, K, M) * FLDICE;
      SL(IJ, K, M) = SL(IJ, K, M) + CICOVER(IJ) * SLICE;
      FLD(IJ, K, M) = FLD(IJ, K, M) + CICOVER(IJ) * FLDICE;
      if (SLICEO) (*SLICEO)(IJ, K, M) = SLICE / std::max((1.0 - DELT5 * FLDICE), 1.0);
    }
  }
  if (c.lciwa2) {    // SDICE2 (sdice2.F90:97-121)
    for (int M = 1; M <= NFRE; ++M) for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      const double EWH = 4.0 * std::sqrt(std::max(t.EPSMIN, FL1(IJ, K, M) * t.DFIM(M)));
      const double XK2 = WAVNUM(IJ, M) * WAVNUM(IJ, M);
      const double ALP = c.cdicwa * XK2 * EWH * c.zalpfacb;
      const double FLDICE = -ALP * CGROUP(IJ, M);
      const double SLICE = FL1(IJ, K, M) * FLDICE;
      SL(IJ, K, M) = SL(IJ, K, M) + CICOVER(IJ) * SLICE;
      FLD(IJ, K, M) = FLD(IJ, K, M) + CICOVER(IJ) * FLDICE;
      if (SLICEO) (*SLICEO)(IJ, K, M) = SLICE / std::max((1.0 - DELT5 * FLDICE), 1.0);
    }
  }
  if (c.lciwa3) {    // SDICE3 (sdice3.F90:107-147), IMODEL = 2 (Jie Yu 2022), ALPFAC = ZALPFACX (no ice-breakup coupling)
        const double CDICE = 0.1274 * std::pow(t.ZPI / std::sqrt(t.G), 4.5);
    for (int M = 1; M <= NFRE; ++M) for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      const double ALP = (2. * CDICE * std::pow(CITHICK(IJ), 1.25) * std::pow(t.FR(M), 4.5)) * (ALPFAC ? ALPFAC[IJ] : c.zalpfacx);
      const double TEMP = -CICOVER(IJ) * ALP * CGROUP(IJ, M);
      SL(IJ, K, M) = SL(IJ, K, M) + FL1(IJ, K, M) * TEMP;
      FLD(IJ, K, M) = FLD(IJ, K, M) + TEMP;
      if (SLICEO) { const double FLDICE = -ALP * CGROUP(IJ, M); (*SLICEO)(IJ, K, M) = (FL1(IJ, K, M) * FLDICE) / std::max((1.0 - DELT5 * FLDICE), 1.0); }
    }
  }
}

// implsch.F90:177-465 for one NPROMA chunk, with sinflx.F90:101-185 inlined as a lambda
void implsch_chunk(const Config& c, const Tables& t, Fields& f, int KIJL, int ICHNK) {
  const int NANG = c.nang, NFRE = c.nfre, KIJS = 1;
  Ctx x{c, t, KIJS, KIJL, NANG, NFRE};
  const int P = KIJL;
  auto s3 = [&](ArrD& a) { return V3{&a(1, 1, 1, ICHNK), P, NANG}; };
  auto s2 = [&](ArrD& a) { return V2{&a(1, 1, ICHNK), P}; };
  auto s1 = [&](ArrD& a) { return V1{&a(1, ICHNK)}; };
  V3 FL1 = s3(f.FL1), XLLWS = s3(f.XLLWS);
  V2 WAVNUM = s2(f.WAVNUM), CGROUP = s2(f.CGROUP), CINV = s2(f.CINV), XK2CG = s2(f.XK2CG), STOKFAC = s2(f.STOKFAC);
  V1 EMAXDPT = s1(f.EMAXDPT), DEPTH = s1(f.DEPTH), AIRD = s1(f.AIRD), WDWAVE = s1(f.WDWAVE), CICOVER = s1(f.CICOVER),
     WSWAVE = s1(f.WSWAVE), WSTAR = s1(f.WSTAR), USTRA = s1(f.USTRA), VSTRA = s1(f.VSTRA), UFRIC = s1(f.UFRIC),
     TAUW = s1(f.TAUW), TAUWDIR = s1(f.TAUWDIR), Z0M = s1(f.Z0M), Z0B = s1(f.Z0B), CHRNCK = s1(f.CHRNCK),
     WSEMEAN = s1(f.WSEMEAN), WSFMEAN = s1(f.WSFMEAN), USTOKES = s1(f.USTOKES), VSTOKES = s1(f.VSTOKES),
     TAUXD = s1(f.TAUXD), TAUYD = s1(f.TAUYD), TAUOCXD = s1(f.TAUOCXD), TAUOCYD = s1(f.TAUOCYD), TAUOC = s1(f.TAUOC),
     TAUICX = s1(f.TAUICX), TAUICY = s1(f.TAUICY), PHIOCD = s1(f.PHIOCD), PHIEPS = s1(f.PHIEPS), PHIAW = s1(f.PHIAW);
  I1 MIJ{&f.MIJ(1, ICHNK)};

  const size_t n3 = (size_t)P * NANG * NFRE;
  std::vector<double> sFLD(n3), sSL(n3), sSPOS(n3), sSSOURCE(n3, 0.0);
  V3 FLD{sFLD.data(), P, NANG}, SL{sSL.data(), P, NANG}, SPOS{sSPOS.data(), P, NANG}, SSOURCE{sSSOURCE.data(), P, NANG};
  std::vector<double> sSLICE(c.lwnemocouwrs ? n3 : 0, 0.0);     // implsch.F90:205-213: zero unless an SDICE term fills it
  V3 SLICE{sSLICE.data(), P, NANG};
  std::vector<double> sFLM((size_t)P * NANG), sCOSWDIF((size_t)P * NANG), sSINWDIF2((size_t)P * NANG), sTEMP((size_t)P * NFRE), sRHOWGDFTH((size_t)P * NFRE);
  V2 FLM{sFLM.data(), P}, COSWDIF{sCOSWDIF.data(), P}, SINWDIF2{sSINWDIF2.data(), P}, TEMP{sTEMP.data(), P}, RHOWGDFTH{sRHOWGDFTH.data(), P};
  L1 lRAORW(P), lEMEAN(P), lFMEAN(P), lHALP(P), lEMEANWS(P), lFMEANWS(P), lUSFM(P), lF1MEAN(P), lAKMEAN(P), lXKMEAN(P), lPHIWA(P), lRNFAC(P);
  V1 RAORW = lRAORW.view(), EMEAN = lEMEAN.view(), FMEAN = lFMEAN.view(), HALP = lHALP.view(), FMEANWS = lFMEANWS.view(),
     USFM = lUSFM.view(), F1MEAN = lF1MEAN.view(), AKMEAN = lAKMEAN.view(), XKMEAN = lXKMEAN.view(), PHIWA = lPHIWA.view(), RNFAC = lRNFAC.view();
  std::vector<double> DELFL(NFRE + 1);

  double DELT = c.idelt, DELTM = 1.0 / DELT, DELT5 = c.ximp * DELT;
  bool LCFLX = c.lwflux || c.lwfluxout || c.lwnemocou;
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) RAORW(IJ) = std::max(AIRD(IJ), 1.0) * t.ROWATERM1;
  for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
    COSWDIF(IJ, K) = std::cos(t.TH(K) - WDWAVE(IJ));
    SINWDIF2(IJ, K) = sq(std::sin(t.TH(K) - WDWAVE(IJ)));
  }
  if (c.lbiwbk) SDEPTHLIM(x, EMAXDPT, FL1);
  FKMEAN(x, FL1, WAVNUM, EMEAN, FMEAN, F1MEAN, AKMEAN, XKMEAN);
  for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ)
    FLM(IJ, K) = (1. - 0.9 * std::min(CICOVER(IJ), 0.99)) * c.flmin * sq(std::max(0.0, COSWDIF(IJ, K)));

  // ---- SINFLX x2 (sinflx.F90)
  const int NCALL = 2;
  for (int ICALL = 1; ICALL <= NCALL; ++ICALL) {
    int IUSFG, ICODE_WND;
    if (ICALL == 1) { IUSFG = 0; ICODE_WND = c.icode; } else { IUSFG = 1; ICODE_WND = 3; }
    if (ICODE_WND != 3 && ICODE_WND != 1 && ICODE_WND != 2) throw std::runtime_error("AIRSEA: INVALID VALUE OF ICODE_WND");
    if (ICALL == 1) {
      for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) FL1(IJ, K, NFRE) = std::max(FL1(IJ, K, NFRE), FLM(IJ, K));
      if (c.llgcbz0) HALPHAP(x, WAVNUM, COSWDIF, FL1, HALP);
      else for (int IJ = KIJS; IJ <= KIJL; ++IJ) HALP(IJ) = 0.0;
    }
    for (int IJ = KIJS; IJ <= KIJL; ++IJ)   // sinflx.F90:117-121
      RNFAC(IJ) = (c.llnormagam && c.llcapchnk) ? 1.0 + t.DTHRN_A * (1.0 + std::tanh(WSWAVE(IJ) - t.DTHRN_U)) : 1.0;
    if (ICODE_WND == 3) TAUT_Z0(x, IUSFG, HALP, WSWAVE, WDWAVE, TAUW, TAUWDIR, RNFAC, UFRIC, Z0M, Z0B, CHRNCK);
    else {   // the friction velocity is the forcing (airsea.F90:102-120): Z0WAVE (z0wave.F90:68-93), then U10 from the logarithmic profile
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
        const double ALPHAOG = (c.llcapchnk ? CHNKMIN(t, WSWAVE(IJ)) : t.ALPHA) * t.GM1;
        const double UST2 = UFRIC(IJ) * UFRIC(IJ), UST3 = UFRIC(IJ) * UFRIC(IJ) * UFRIC(IJ);
        const double ARG = std::max(UST2 - TAUW(IJ), t.EPS1);
        Z0M(IJ) = ALPHAOG * UST3 / std::sqrt(ARG);
        Z0B(IJ) = ALPHAOG * UST2;
        CHRNCK(IJ) = t.G * Z0M(IJ) / UST2;
      }
      const double XKAPPAD = 1.0 / t.XKAPPA, XLOGLEV = std::log(t.XNLEV);
      for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
        WSWAVE(IJ) = XKAPPAD * UFRIC(IJ) * (XLOGLEV - std::log(Z0M(IJ)));
        WSWAVE(IJ) = std::max(WSWAVE(IJ), c.wspmin);
      }
    }
    int NGST; bool LLPHIWA, LLSNEG;
    if (ICALL < NCALL) { NGST = 1; LLPHIWA = false; LLSNEG = false; } else { NGST = 2; LLPHIWA = true; LLSNEG = true; }
    if (c.iphys == 0) SINPUT_JAN(x, NGST, LLSNEG, FL1, WAVNUM, CINV, XK2CG, WSWAVE, UFRIC, Z0M, COSWDIF, SINWDIF2, RAORW, WSTAR, RNFAC, FLD, SL, SPOS, XLLWS);
    else SINPUT_ARD(x, NGST, LLSNEG, FL1, WAVNUM, CINV, XK2CG, WDWAVE, WSWAVE, UFRIC, Z0M, COSWDIF, SINWDIF2, RAORW, WSTAR, RNFAC, FLD, SL, SPOS, XLLWS);
    FEMEANWS(x, FL1, XLLWS, FMEANWS, nullptr);
    FRCUTINDEX(x, FMEAN, FMEANWS, UFRIC, CICOVER, MIJ, RHOWGDFTH);
    STRESSO(x, MIJ, RHOWGDFTH, FL1, SL, SPOS, CINV, WDWAVE, UFRIC, Z0M, AIRD, RNFAC, COSWDIF, SINWDIF2, TAUW, TAUWDIR, PHIWA, LLPHIWA);
  }
  if (c.iphys == 0) SDISSIP_JAN(x, FL1, FLD, SL, WAVNUM, EMEAN, F1MEAN, XKMEAN);
  else SDISSIP_ARD(x, FL1, FLD, SL, WAVNUM, CGROUP, XK2CG, UFRIC, COSWDIF, RAORW);
  if (LCFLX && !c.lwvflx_snl) for (size_t i = 0; i < n3; ++i) sSSOURCE[i] = sSL[i];
  SNONLIN(x, FL1, FLD, SL, WAVNUM, DEPTH, AKMEAN);
  if (LCFLX && c.lwvflx_snl)
    for (int M = 1; M <= NFRE; ++M) for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      double GTEMP1 = std::max((1.0 - DELT5 * FLD(IJ, K, M)), 1.0);
      SSOURCE(IJ, K, M) = SL(IJ, K, M) / GTEMP1;
    }
  SDIWBK(x, FL1, FLD, SL, DEPTH, EMAXDPT, EMEAN, F1MEAN);
  if (c.licerun) {   // implsch.F90:312-339
    if (c.lciscal)
      for (int M = 1; M <= NFRE; ++M) for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
        const double BETA = 1. - CICOVER(IJ);
        SL(IJ, K, M) = BETA * SL(IJ, K, M);
        FLD(IJ, K, M) = BETA * FLD(IJ, K, M);
      }
    // ICEBREAK_MODIFY_ATTENUATION (icebreak_modify_attenuation.F90:82-94, LWNEMOCOUIBR): broken ice attenuates 1/ZALPFACX instead of ZALPFACX
    std::vector<double> ALPFAC(KIJL + 1, c.zalpfacx);
    if (c.lwnemocouibr) { V1 IBRMEM = s1(f.IBRMEM); for (int IJ = KIJS; IJ <= KIJL; ++IJ) if (IBRMEM(IJ) <= c.zibrw_thrsh) ALPFAC[IJ] = 1.0 / c.zalpfacx; }
    if (c.lciwa1 || c.lciwa2 || c.lciwa3) SDICE(x, FL1, FLD, SL, WAVNUM, CGROUP, CICOVER, s1(f.CITHICK), ALPFAC.data(), c.lwnemocouwrs ? &SLICE : nullptr);
  }
  SBOTTOM(x, FL1, FLD, SL, WAVNUM, DEPTH);
  // ---- 2.4 new spectra (implsch.F90:352-395)
  for (int M = 1; M <= NFRE; ++M) DELFL[M] = t.COFRM4(M) * DELT;
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) USFM(IJ) = UFRIC(IJ) * std::max(FMEANWS(IJ), FMEAN(IJ));
  for (int M = 1; M <= NFRE; ++M) for (int IJ = KIJS; IJ <= KIJL; ++IJ) TEMP(IJ, M) = USFM(IJ) * DELFL[M];
  for (int K = 1; K <= NANG; ++K) for (int M = 1; M <= NFRE; ++M) for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
    double GTEMP1 = std::max((1.0 - DELT5 * FLD(IJ, K, M)), 1.0);
    double GTEMP2 = DELT * SL(IJ, K, M) / GTEMP1;
    double FLHAB = std::fabs(GTEMP2);
    FLHAB = std::min(FLHAB, TEMP(IJ, M));
    FL1(IJ, K, M) = FL1(IJ, K, M) + sign(FLHAB, GTEMP2);
    FL1(IJ, K, M) = std::max(FL1(IJ, K, M), FLM(IJ, K));
    SSOURCE(IJ, K, M) = SSOURCE(IJ, K, M) + DELTM * std::min(t.FLMAX(M) - FL1(IJ, K, M), 0.0);
    FL1(IJ, K, M) = std::min(FL1(IJ, K, M), t.FLMAX(M));
  }
  if (f.capture) {
    for (int M = 1; M <= NFRE; ++M) for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) f.DBG_SSOURCE(IJ, K, M, ICHNK) = SSOURCE(IJ, K, M);
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) { f.DBG_EM(IJ, ICHNK) = EMEAN(IJ); f.DBG_F1(IJ, ICHNK) = F1MEAN(IJ); f.DBG_PHIWA(IJ, ICHNK) = PHIWA(IJ); }
  }
  NemoFlux NE{s1(f.NPHIEPS), s1(f.NTAUOC), s1(f.NSWH), s1(f.NMWP), s1(f.NEMOTAUX), s1(f.NEMOTAUY), s1(f.NEMOTAUICX), s1(f.NEMOTAUICY),
              s1(f.NEMOWSWAVE), s1(f.NEMOPHIF)};
  if (LCFLX)
    WNFLUXES(x, MIJ, RHOWGDFTH, CINV, SSOURCE, CICOVER, PHIWA, EMEAN, F1MEAN, WSWAVE, WDWAVE, USTRA, VSTRA, UFRIC, AIRD,
             TAUXD, TAUYD, TAUOCXD, TAUOCYD, TAUOC, TAUICX, TAUICY, PHIOCD, PHIEPS, PHIAW, &NE, &SLICE);
  // ---- 2.5 tail
  FKMEAN(x, FL1, WAVNUM, EMEAN, FMEAN, F1MEAN, AKMEAN, XKMEAN);
  FEMEANWS(x, FL1, XLLWS, FMEANWS, lEMEANWS.v.data());
  IMPHFTAIL(x, MIJ, FLM, WAVNUM, XK2CG, FL1);
  if (c.lwflux)
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      if (lEMEANWS.v[IJ - 1] < t.WSEMEAN_MIN) { WSEMEAN(IJ) = t.WSEMEAN_MIN; WSFMEAN(IJ) = 2. * t.FR(NFRE); }
      else { WSEMEAN(IJ) = lEMEANWS.v[IJ - 1]; WSFMEAN(IJ) = FMEANWS(IJ); }
    }
  if (c.licerun && c.lmaskice) SETICE(x, FL1, CICOVER, COSWDIF);
  // STOKESTRN (stokestrn.F90:64-88)
  STOKESDRIFT(x, FL1, STOKFAC, WSWAVE, WDWAVE, CICOVER, USTOKES, VSTOKES);
  V1 STRNMS = s1(f.STRNMS);
  if (c.lwnemocoustrn) CIMSSTRN(x, FL1, WAVNUM, DEPTH, s1(f.CITHICK), STRNMS);
  if (c.lwnemocou && ((c.lwnemocousend && c.lwcou) || !c.lwcou)) {
    V1 NUS = s1(f.NEMOUSTOKES), NVS = s1(f.NEMOVSTOKES), NST = s1(f.NEMOSTRN);
    for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
      if (c.lwnemocoustk) { NUS(IJ) = USTOKES(IJ); NVS(IJ) = VSTOKES(IJ); }
      else { NUS(IJ) = 0.0; NVS(IJ) = 0.0; }
      if (c.lwnemocoustrn) NST(IJ) = STRNMS(IJ);
    }
  }
}

// SNONLIN (snonlin.F90) on its own with AKMEAN from FKMEAN, as IMPLSCH calls it (implsch.F90:288)
void snonlin_chunk(const Config& c, const Tables& t, Fields& f, int KIJL, int ICHNK, double* SLp, double* FLDp) {
  const int NANG = c.nang, NFRE = c.nfre, KIJS = 1;
  Ctx x{c, t, KIJS, KIJL, NANG, NFRE};
  const int P = KIJL;
  V3 FL1{&f.FL1(1, 1, 1, ICHNK), P, NANG};
  V2 WAVNUM{&f.WAVNUM(1, 1, ICHNK), P};
  V1 DEPTH{&f.DEPTH(1, ICHNK)};
  L1 lEM(P), lFM(P), lF1(P), lAK(P), lXK(P);
  FKMEAN(x, FL1, WAVNUM, lEM.view(), lFM.view(), lF1.view(), lAK.view(), lXK.view());
  const size_t n3 = (size_t)P * NANG * NFRE;
  for (size_t i = 0; i < n3; ++i) { SLp[i] = 0.0; FLDp[i] = 0.0; }
  SNONLIN(x, FL1, V3{FLDp, P, NANG}, V3{SLp, P, NANG}, WAVNUM, DEPTH, lAK.view());
}

// One source term alone for a chunk, from the fields as they are stored (test infrastructure for the per-term cross-checks):
// which = 1: SINPUT with NGST = 1, LLSNEG = F (the first SINFLX call, sinflx.F90:156-167) using the stored UFRIC, Z0M;
// which = 2: SDISSIP (sdissip.F90); which = 3: SBOTTOM; which = 4: SDIWBK (with FKMEAN's EMEAN, F1MEAN); which = 5: SDICE (the LCIWA1-3 terms that are switched on).  SL, FLD: (KIJL, NANG, NFRE).
void term_chunk(const Config& c, const Tables& t, Fields& f, int KIJL, int ICHNK, int which, double* SLp, double* FLDp) {
  const int NANG = c.nang, NFRE = c.nfre, KIJS = 1;
  Ctx x{c, t, KIJS, KIJL, NANG, NFRE};
  const int P = KIJL;
  auto s1 = [&](ArrD& a) { return V1{&a(1, ICHNK)}; };
  V3 FL1{&f.FL1(1, 1, 1, ICHNK), P, NANG};
  V2 WAVNUM{&f.WAVNUM(1, 1, ICHNK), P}, CINV{&f.CINV(1, 1, ICHNK), P}, XK2CG{&f.XK2CG(1, 1, ICHNK), P}, CGROUP{&f.CGROUP(1, 1, ICHNK), P};
  V1 WSWAVE = s1(f.WSWAVE), WDWAVE = s1(f.WDWAVE), UFRIC = s1(f.UFRIC), Z0M = s1(f.Z0M), AIRD = s1(f.AIRD), WSTAR = s1(f.WSTAR);
  const size_t n3 = (size_t)P * NANG * NFRE;
  for (size_t i = 0; i < n3; ++i) { SLp[i] = 0.0; FLDp[i] = 0.0; }
  V3 SL{SLp, P, NANG}, FLD{FLDp, P, NANG};
  std::vector<double> sCOS((size_t)P * NANG), sSIN2((size_t)P * NANG);
  V2 COSWDIF{sCOS.data(), P}, SINWDIF2{sSIN2.data(), P};
  L1 lRAORW(P), lRNFAC(P);
  V1 RAORW = lRAORW.view(), RNFAC = lRNFAC.view();
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) { RAORW(IJ) = std::max(AIRD(IJ), 1.0) * t.ROWATERM1; RNFAC(IJ) = 1.0; }
  for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
    COSWDIF(IJ, K) = std::cos(t.TH(K) - WDWAVE(IJ));
    SINWDIF2(IJ, K) = sq(std::sin(t.TH(K) - WDWAVE(IJ)));
  }
  if (which == 1) {
    std::vector<double> sSPOS(n3), sX(n3);
    V3 SPOS{sSPOS.data(), P, NANG}, XLLWS{sX.data(), P, NANG};
    if (c.iphys == 0) SINPUT_JAN(x, 1, false, FL1, WAVNUM, CINV, XK2CG, WSWAVE, UFRIC, Z0M, COSWDIF, SINWDIF2, RAORW, WSTAR, RNFAC, FLD, SL, SPOS, XLLWS);
    else SINPUT_ARD(x, 1, false, FL1, WAVNUM, CINV, XK2CG, WDWAVE, WSWAVE, UFRIC, Z0M, COSWDIF, SINWDIF2, RAORW, WSTAR, RNFAC, FLD, SL, SPOS, XLLWS);
  } else if (which == 2) {
    L1 lEM(P), lFM(P), lF1(P), lAK(P), lXK(P);
    FKMEAN(x, FL1, WAVNUM, lEM.view(), lFM.view(), lF1.view(), lAK.view(), lXK.view());
    if (c.iphys == 0) SDISSIP_JAN(x, FL1, FLD, SL, WAVNUM, lEM.view(), lF1.view(), lXK.view());
    else SDISSIP_ARD(x, FL1, FLD, SL, WAVNUM, CGROUP, XK2CG, UFRIC, COSWDIF, RAORW);
  } else if (which == 3) {
    V1 DEPTH = s1(f.DEPTH);
    SBOTTOM(x, FL1, FLD, SL, WAVNUM, DEPTH);
  } else if (which == 4) {
    V1 DEPTH = s1(f.DEPTH), EMAXDPT = s1(f.EMAXDPT);
    L1 lEM(P), lFM(P), lF1(P), lAK(P), lXK(P);
    FKMEAN(x, FL1, WAVNUM, lEM.view(), lFM.view(), lF1.view(), lAK.view(), lXK.view());
    SDIWBK(x, FL1, FLD, SL, DEPTH, EMAXDPT, lEM.view(), lF1.view());
  } else if (which == 5) {
    SDICE(x, FL1, FLD, SL, WAVNUM, CGROUP, s1(f.CICOVER), s1(f.CITHICK), nullptr, nullptr);
  } else throw std::runtime_error("term_chunk: unknown term");
}

// STRESSO + TAU_PHI_HF alone for a chunk (test infrastructure for the per-term cross-checks): the wind input of the second SINFLX
// call (NGST = 2, LLSNEG = T; sinflx.F90:156-167) from the stored UFRIC / Z0M, then STRESSO with LLPHIWA = T on the stored MIJ
// (stresso.F90:120-233).  SL, SPOS: (KIJL, NANG, NFRE); OUT: (KIJL, 3) = TAUW, TAUWDIR, PHIWA.
void stresso_chunk(const Config& c, const Tables& t, Fields& f, int KIJL, int ICHNK, double* SLp, double* SPOSp, double* OUTp) {
  const int NANG = c.nang, NFRE = c.nfre, KIJS = 1;
  Ctx x{c, t, KIJS, KIJL, NANG, NFRE};
  const int P = KIJL;
  auto s1 = [&](ArrD& a) { return V1{&a(1, ICHNK)}; };
  V3 FL1{&f.FL1(1, 1, 1, ICHNK), P, NANG};
  V2 WAVNUM{&f.WAVNUM(1, 1, ICHNK), P}, CINV{&f.CINV(1, 1, ICHNK), P}, XK2CG{&f.XK2CG(1, 1, ICHNK), P};
  V1 WSWAVE = s1(f.WSWAVE), WDWAVE = s1(f.WDWAVE), UFRIC = s1(f.UFRIC), Z0M = s1(f.Z0M), AIRD = s1(f.AIRD), WSTAR = s1(f.WSTAR);
  const size_t n3 = (size_t)P * NANG * NFRE;
  std::vector<double> sFLD(n3, 0.0), sX(n3), sCOS((size_t)P * NANG), sSIN2((size_t)P * NANG), sRH((size_t)P * NFRE);
  for (size_t i = 0; i < n3; ++i) { SLp[i] = 0.0; SPOSp[i] = 0.0; }
  V3 SL{SLp, P, NANG}, SPOS{SPOSp, P, NANG}, FLD{sFLD.data(), P, NANG}, XLLWS{sX.data(), P, NANG};
  V2 COSWDIF{sCOS.data(), P}, SINWDIF2{sSIN2.data(), P}, RHOWGDFTH{sRH.data(), P};
  L1 lRAORW(P), lRNFAC(P);
  V1 RAORW = lRAORW.view(), RNFAC = lRNFAC.view();
  std::vector<int> sMIJ(P);
  I1 MIJ{sMIJ.data()};
  for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
    RAORW(IJ) = std::max(AIRD(IJ), 1.0) * t.ROWATERM1; RNFAC(IJ) = 1.0;
    MIJ(IJ) = f.MIJ(IJ, ICHNK);
    for (int M = 1; M <= NFRE; ++M) RHOWGDFTH(IJ, M) = M <= MIJ(IJ) ? t.RHOWG_DFIM(M) : 0.0;     // frcutindex.F90:99-108
    if (MIJ(IJ) != NFRE) RHOWGDFTH(IJ, MIJ(IJ)) = 0.5 * RHOWGDFTH(IJ, MIJ(IJ));
  }
  for (int K = 1; K <= NANG; ++K) for (int IJ = KIJS; IJ <= KIJL; ++IJ) {
    COSWDIF(IJ, K) = std::cos(t.TH(K) - WDWAVE(IJ));
    SINWDIF2(IJ, K) = sq(std::sin(t.TH(K) - WDWAVE(IJ)));
  }
  if (c.iphys == 0) SINPUT_JAN(x, 2, true, FL1, WAVNUM, CINV, XK2CG, WSWAVE, UFRIC, Z0M, COSWDIF, SINWDIF2, RAORW, WSTAR, RNFAC, FLD, SL, SPOS, XLLWS);
  else SINPUT_ARD(x, 2, true, FL1, WAVNUM, CINV, XK2CG, WDWAVE, WSWAVE, UFRIC, Z0M, COSWDIF, SINWDIF2, RAORW, WSTAR, RNFAC, FLD, SL, SPOS, XLLWS);
  STRESSO(x, MIJ, RHOWGDFTH, FL1, SL, SPOS, CINV, WDWAVE, UFRIC, Z0M, AIRD, RNFAC, COSWDIF, SINWDIF2, V1{OUTp}, V1{OUTp + P}, V1{OUTp + 2 * P}, true);
}

// wamintgr.F90:117-146
void implsch_all(Model& m) {
  for (int ir = 0; ir < m.cfg.npr; ++ir) {
    RankDecomp& r = m.ranks[ir];
    Fields& f = m.fld[ir];
#pragma omp parallel for schedule(dynamic, 1)
    for (int ICHNK = 1; ICHNK <= r.NCHNK; ++ICHNK) implsch_chunk(m.cfg, m.tab, f, r.NPROMA, ICHNK);
  }
}

}  // namespace orc
