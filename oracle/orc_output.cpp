// ORACLE (test infrastructure) — the steps either side of the hot path each time step / output step:
// NEWWIND (newwind.F90:105-167), OUTBS/OUTBLOCK's core integrated parameters (outbs.F90:97-126, outblock.F90:150-610 with
// FEMEAN, STHQ, DOMINANT_PERIOD, SEPWISW (LLPARTITION=F), MWP1, MWP2, WDIRSPREAD/PEAKFRI/SCOSFL, OUTBETA, WEFLUX,
// OUTSETWMASK) and the WAMNORM statistics (outwnorm.F90:83, mpminmaxavg.F90:68-195).  Loop by loop, in the reference's
// operation order.  Paths relative to /root/reference/src/ecwam.
#include "oracle.h"

namespace orc {

namespace {
struct S3 {   // (KIJL, NANG, NFRE) view, 1-based
  const double* p; long P, A;
  inline double operator()(int ij, int k, int m) const { return p[(ij - 1) + P * ((k - 1) + A * (long)(m - 1))]; }
};
struct W3 {
  std::vector<double> d; long P, A;
  W3(long P_, long A_, long F_) : d((size_t)P_ * A_ * F_, 0.0), P(P_), A(A_) {}
  inline double& operator()(int ij, int k, int m) { return d[(ij - 1) + P * ((k - 1) + A * (long)(m - 1))]; }
  S3 view() const { return S3{d.data(), P, A}; }
};
typedef std::vector<double> V;

// femean.F90:84-121
void femean_out(const Tables& t, int KIJL, int NANG, int NFRE, const S3& F, V& EM, V& FM) {
  V TEMP2(KIJL + 1);
  for (int IJ = 1; IJ <= KIJL; ++IJ) { EM[IJ] = 0.0; FM[IJ] = 0.0; }
  const double DELT25 = t.WETAIL * t.FR(NFRE) * t.DELTH;
  const double DELT2 = t.FRTAIL * t.DELTH;
  for (int M = 1; M <= NFRE; ++M) {
    for (int IJ = 1; IJ <= KIJL; ++IJ) TEMP2[IJ] = std::max(F(IJ, 1, M), t.EPSMIN);
    for (int K = 2; K <= NANG; ++K)
      for (int IJ = 1; IJ <= KIJL; ++IJ) TEMP2[IJ] = TEMP2[IJ] + std::max(F(IJ, K, M), t.EPSMIN);
    for (int IJ = 1; IJ <= KIJL; ++IJ) { EM[IJ] = EM[IJ] + TEMP2[IJ] * t.DFIM(M); FM[IJ] = FM[IJ] + t.DFIMOFR(M) * TEMP2[IJ]; }
  }
  for (int IJ = 1; IJ <= KIJL; ++IJ) {
    EM[IJ] = EM[IJ] + DELT25 * TEMP2[IJ];
    FM[IJ] = FM[IJ] + DELT2 * TEMP2[IJ];
    FM[IJ] = EM[IJ] / FM[IJ];
    FM[IJ] = std::max(FM[IJ], t.FR(1));
  }
}

// sthq.F90:69-115
void sthq(const Tables& t, int KIJL, int NANG, int NFRE, const S3& F, V& THQ) {
  V TEMP(KIJL + 1), SI(KIJL + 1, 0.0), CI(KIJL + 1, 0.0);
  for (int K = 1; K <= NANG; ++K) {
    for (int IJ = 1; IJ <= KIJL; ++IJ) TEMP[IJ] = 0.0;
    for (int M = 1; M <= NFRE; ++M)
      for (int IJ = 1; IJ <= KIJL; ++IJ) TEMP[IJ] = TEMP[IJ] + F(IJ, K, M) * t.DFIM(M);
    for (int IJ = 1; IJ <= KIJL; ++IJ) { SI[IJ] = SI[IJ] + t.SINTH(K) * TEMP[IJ]; CI[IJ] = CI[IJ] + t.COSTH(K) * TEMP[IJ]; }
  }
  for (int IJ = 1; IJ <= KIJL; ++IJ) if (CI[IJ] == 0.0) CI[IJ] = t.EPSMIN;
  for (int IJ = 1; IJ <= KIJL; ++IJ) THQ[IJ] = std::atan2(SI[IJ], CI[IJ]);
  for (int IJ = 1; IJ <= KIJL; ++IJ) if (THQ[IJ] < 0.0) THQ[IJ] = THQ[IJ] + t.ZPI;
}

// mwp1.F90:75-112 (IP=1) and mwp2.F90 (IP=2): same loops with DFIMFR_SIM / DFIMFR2_SIM (initmdl.F90:497-502)
void mwp12(const Tables& t, int KIJL, int NANG, int IP, const S3& F, V& MEANWP) {
  V TEMP(KIJL + 1), EM(KIJL + 1, 0.0);
  const int NO = t.NFRE_ODD;
  for (int IJ = 1; IJ <= KIJL; ++IJ) MEANWP[IJ] = 0.0;
  for (int M = 1; M <= NO; ++M) {
    for (int IJ = 1; IJ <= KIJL; ++IJ) TEMP[IJ] = 0.0;
    for (int K = 1; K <= NANG; ++K)
      for (int IJ = 1; IJ <= KIJL; ++IJ) TEMP[IJ] = TEMP[IJ] + F(IJ, K, M);
    const double w = IP == 1 ? t.DFIM_SIM(M) * t.FR(M) : t.DFIM_SIM(M) * (t.FR(M) * t.FR(M));
    for (int IJ = 1; IJ <= KIJL; ++IJ) { EM[IJ] = EM[IJ] + t.DFIM_SIM(M) * TEMP[IJ]; MEANWP[IJ] = MEANWP[IJ] + w * TEMP[IJ]; }
  }
  const double FR1M1 = 1.0 / t.FR(1);
  const double DELT25 = t.WETAIL * t.FR(NO) * t.DELTH;
  const double COEF_FR = IP == 1 ? t.WP1TAIL * t.DELTH * (t.FR(NO) * t.FR(NO)) : t.WP2TAIL * t.DELTH * (t.FR(NO) * t.FR(NO) * t.FR(NO));
  for (int IJ = 1; IJ <= KIJL; ++IJ) {
    EM[IJ] = EM[IJ] + DELT25 * TEMP[IJ];
    MEANWP[IJ] = MEANWP[IJ] + COEF_FR * TEMP[IJ];
    if (EM[IJ] > 0.0 && MEANWP[IJ] > t.EPSMIN) {
      MEANWP[IJ] = IP == 1 ? EM[IJ] / MEANWP[IJ] : std::sqrt(EM[IJ] / MEANWP[IJ]);
      MEANWP[IJ] = std::min(MEANWP[IJ], FR1M1);
    } else MEANWP[IJ] = 0.0;
  }
}

// scosfl.F90:74-104
void scosfl(const Tables& t, int KIJL, int NANG, const S3& F, const std::vector<int>& MM, V& MEANCOSFL) {
  V SI(KIJL + 1, 0.0), CI(KIJL + 1, 0.0), MEANDIR(KIJL + 1);
  for (int IJ = 1; IJ <= KIJL; ++IJ) MEANCOSFL[IJ] = 0.0;
  for (int K = 1; K <= NANG; ++K)
    for (int IJ = 1; IJ <= KIJL; ++IJ) { SI[IJ] = SI[IJ] + t.SINTH(K) * F(IJ, K, MM[IJ]); CI[IJ] = CI[IJ] + t.COSTH(K) * F(IJ, K, MM[IJ]); }
  for (int IJ = 1; IJ <= KIJL; ++IJ) MEANDIR[IJ] = (CI[IJ] == 0.0 && SI[IJ] == 0.0) ? 0.0 : std::atan2(SI[IJ], CI[IJ]);
  for (int K = 1; K <= NANG; ++K)
    for (int IJ = 1; IJ <= KIJL; ++IJ) MEANCOSFL[IJ] = MEANCOSFL[IJ] + std::cos(t.TH(K) - MEANDIR[IJ]) * F(IJ, K, MM[IJ]);
  for (int IJ = 1; IJ <= KIJL; ++IJ) MEANCOSFL[IJ] = t.DELTH * MEANCOSFL[IJ];
}

// wdirspread.F90:92-140 with peakfri.F90:66-88
void wdirspread(const Tables& t, int KIJL, int NANG, int NFRE, const S3& F, const V& EMEAN, bool LLPEAKF, V& WDIRSPRD) {
  const double ONE = 1.0, COEF_FR = t.WETAIL * t.FR(NFRE);
  V TEMP(KIJL + 1, 0.0);
  std::vector<int> IFRINDEX(KIJL + 1, NFRE);
  for (int IJ = 1; IJ <= KIJL; ++IJ) WDIRSPRD[IJ] = 0.0;
  if (LLPEAKF) {
    V F1D(KIJL + 1);
    for (int IJ = 1; IJ <= KIJL; ++IJ) { TEMP[IJ] = 0.0; IFRINDEX[IJ] = NFRE; }
    for (int M = 1; M <= NFRE; ++M) {
      for (int IJ = 1; IJ <= KIJL; ++IJ) F1D[IJ] = 0.0;
      for (int K = 1; K <= NANG; ++K)
        for (int IJ = 1; IJ <= KIJL; ++IJ) F1D[IJ] = F1D[IJ] + F(IJ, K, M) * t.DELTH;
      for (int IJ = 1; IJ <= KIJL; ++IJ) if (TEMP[IJ] < F1D[IJ]) { TEMP[IJ] = F1D[IJ]; IFRINDEX[IJ] = M; }
    }
    scosfl(t, KIJL, NANG, F, IFRINDEX, WDIRSPRD);
    for (int IJ = 1; IJ <= KIJL; ++IJ) WDIRSPRD[IJ] = TEMP[IJ] > 0.0 ? std::min(WDIRSPRD[IJ] / TEMP[IJ], ONE) : ONE;
  } else {
    for (int M = 1; M <= NFRE; ++M) {
      for (int IJ = 1; IJ <= KIJL; ++IJ) IFRINDEX[IJ] = M;
      scosfl(t, KIJL, NANG, F, IFRINDEX, TEMP);
      for (int IJ = 1; IJ <= KIJL; ++IJ) WDIRSPRD[IJ] = WDIRSPRD[IJ] + TEMP[IJ] * t.DFIM(M);
    }
    for (int IJ = 1; IJ <= KIJL; ++IJ) WDIRSPRD[IJ] = WDIRSPRD[IJ] / t.DELTH + TEMP[IJ] * COEF_FR;
    for (int IJ = 1; IJ <= KIJL; ++IJ) WDIRSPRD[IJ] = EMEAN[IJ] > t.EPSMIN ? std::min(WDIRSPRD[IJ] / EMEAN[IJ], ONE) : ONE;
  }
  for (int IJ = 1; IJ <= KIJL; ++IJ) WDIRSPRD[IJ] = std::sqrt(2.0 * (ONE - WDIRSPRD[IJ]));
}

// dominant_period.F90:63-123
void dominant_period(const Tables& t, int KIJL, int NANG, int NFRE, const S3& F, V& DP) {
  const double FLTHRS = 0.1;
  V EM(KIJL + 1, 0.0), FCROP(KIJL + 1, 0.0), F1D4((size_t)(KIJL + 1) * (NFRE + 1), 0.0);
  for (int IJ = 1; IJ <= KIJL; ++IJ) DP[IJ] = 0.0;
  for (int M = 1; M <= NFRE; ++M)
    for (int K = 1; K <= NANG; ++K)
      for (int IJ = 1; IJ <= KIJL; ++IJ) if (F(IJ, K, M) > FCROP[IJ]) FCROP[IJ] = F(IJ, K, M);
  for (int IJ = 1; IJ <= KIJL; ++IJ) FCROP[IJ] = FLTHRS * FCROP[IJ];
  auto f1 = [&](int IJ, int M) -> double& { return F1D4[(size_t)M * (KIJL + 1) + IJ]; };
  for (int M = 1; M <= NFRE; ++M)
    for (int K = 1; K <= NANG; ++K)
      for (int IJ = 1; IJ <= KIJL; ++IJ) if (F(IJ, K, M) > FCROP[IJ]) f1(IJ, M) = f1(IJ, M) + F(IJ, K, M) * t.DELTH;
  for (int M = 1; M <= NFRE; ++M)
    for (int IJ = 1; IJ <= KIJL; ++IJ) {
      const double x = f1(IJ, M);
      f1(IJ, M) = (x * x) * (x * x);   // F1D4**4
      EM[IJ] = EM[IJ] + t.DFIM(M) * f1(IJ, M);
      DP[IJ] = DP[IJ] + t.DFIMFR(M) * f1(IJ, M);
    }
  for (int IJ = 1; IJ <= KIJL; ++IJ) DP[IJ] = (EM[IJ] > 0.0 && DP[IJ] > t.EPSMIN) ? EM[IJ] / DP[IJ] : 0.0;
}

struct SepOut { V ESWELL, FSWELL, THSWELL, P1SWELL, P2SWELL, SPRDSWELL, ESEA, FSEA, THWISEA, P1SEA, P2SEA, SPRDSEA; };

// sepwisw.F90:128-300, LLPARTITION = .FALSE. (mpcrtbl.F90:535), CLDOMAIN /= 's'
void sepwisw(const Tables& t, int KIJL, int NANG, int NFRE, const S3& FL1, const S3& XLLWS, const double* CINV /*(KIJL,NFRE)*/,
             const double* UFRIC, const double* WDWAVE, const V& COSWDIF /*(KIJL,NANG) 1-based*/, SepOut& o) {
  const double FRIC = 28.0, OLDWSFC = 1.2;   // yowfred.F90:81-82
  const double COEF = OLDWSFC * FRIC;
  auto cwd = [&](int IJ, int K) { return COSWDIF[(size_t)(K - 1) * KIJL + IJ]; };
  V R(KIJL + 1), XINVWVAGE((size_t)KIJL * NFRE + 1), DIRCOEF((size_t)KIJL * NANG + 1);
  auto xin = [&](int IJ, int M) -> double& { return XINVWVAGE[(size_t)(M - 1) * KIJL + IJ]; };
  auto dco = [&](int IJ, int K) -> double& { return DIRCOEF[(size_t)(K - 1) * KIJL + IJ]; };
  W3 SWM(KIJL, NANG, NFRE), F1(KIJL, NANG, NFRE);
  for (V* v : {&o.ESWELL, &o.FSWELL, &o.THSWELL, &o.P1SWELL, &o.P2SWELL, &o.SPRDSWELL, &o.ESEA, &o.FSEA, &o.THWISEA, &o.P1SEA,
               &o.P2SEA, &o.SPRDSEA}) v->assign(KIJL + 1, 0.0);
  for (int M = 1; M <= NFRE; ++M) for (int IJ = 1; IJ <= KIJL; ++IJ) xin(IJ, M) = UFRIC[IJ - 1] * CINV[(size_t)(M - 1) * KIJL + IJ - 1];
  for (int K = 1; K <= NANG; ++K) for (int IJ = 1; IJ <= KIJL; ++IJ) dco(IJ, K) = COEF * cwd(IJ, K);
  for (int M = 1; M <= NFRE; ++M)
    for (int K = 1; K <= NANG; ++K)
      for (int IJ = 1; IJ <= KIJL; ++IJ) {
        if (XLLWS(IJ, K, M) != 0.0) SWM(IJ, K, M) = 0.0;
        else {
          const double CHECKTA = xin(IJ, M) * dco(IJ, K);
          SWM(IJ, K, M) = CHECKTA >= 1.0 ? 0.0 : 1.0;
        }
      }
  // swell / windsea mean frequencies -> reset of the wind sector (sepwisw.F90:166-212)
  for (int M = 1; M <= NFRE; ++M) for (int K = 1; K <= NANG; ++K) for (int IJ = 1; IJ <= KIJL; ++IJ) F1(IJ, K, M) = FL1(IJ, K, M) * SWM(IJ, K, M);
  femean_out(t, KIJL, NANG, NFRE, F1.view(), o.ESWELL, o.FSWELL);
  for (int M = 1; M <= NFRE; ++M) for (int K = 1; K <= NANG; ++K) for (int IJ = 1; IJ <= KIJL; ++IJ) F1(IJ, K, M) = std::max(FL1(IJ, K, M) - F1(IJ, K, M), 0.0);
  femean_out(t, KIJL, NANG, NFRE, F1.view(), o.ESEA, o.FSEA);
  for (int IJ = 1; IJ <= KIJL; ++IJ) R[IJ] = o.FSWELL[IJ] > 0.96 * o.FSEA[IJ] ? 1.0 : 0.0;
  for (int K = 1; K <= NANG; ++K) for (int IJ = 1; IJ <= KIJL; ++IJ) dco(IJ, K) = R[IJ] * COEF * sign(1.0, 0.4 + cwd(IJ, K));
  for (int M = 1; M <= NFRE; ++M)
    for (int K = 1; K <= NANG; ++K)
      for (int IJ = 1; IJ <= KIJL; ++IJ) { const double CHECKTA = xin(IJ, M) * dco(IJ, K); if (CHECKTA >= 1.0) SWM(IJ, K, M) = 0.0; }
  // connect the low-frequency boundary of the windsea area (sepwisw.F90:216-226)
  for (int IJ = 1; IJ <= KIJL; ++IJ)
    for (int K = 1; K <= NANG; ++K)
      for (int M = NFRE; M >= 2; --M) {
        if (SWM(IJ, K, M) == 1.0 && SWM(IJ, K, M - 1) == 1.0) break;
        else if (SWM(IJ, K, M) == 0.0 && SWM(IJ, K, M - 1) == 1.0) { if (FL1(IJ, K, M) >= FL1(IJ, K, M - 1)) SWM(IJ, K, M - 1) = 0.0; }
      }
  for (int M = 1; M <= NFRE; ++M) for (int K = 1; K <= NANG; ++K) for (int IJ = 1; IJ <= KIJL; ++IJ) F1(IJ, K, M) = std::max(FL1(IJ, K, M), t.EPSMIN) * SWM(IJ, K, M);
  // total swell (sepwisw.F90:254-265)
  femean_out(t, KIJL, NANG, NFRE, F1.view(), o.ESWELL, o.FSWELL);
  sthq(t, KIJL, NANG, NFRE, F1.view(), o.THSWELL);
  mwp12(t, KIJL, NANG, 1, F1.view(), o.P1SWELL);
  mwp12(t, KIJL, NANG, 2, F1.view(), o.P2SWELL);
  wdirspread(t, KIJL, NANG, NFRE, F1.view(), o.ESWELL, true, o.SPRDSWELL);
  // wind sea (sepwisw.F90:271-298)
  for (int M = 1; M <= NFRE; ++M)
    for (int K = 1; K <= NANG; ++K)
      for (int IJ = 1; IJ <= KIJL; ++IJ) {
        if (cwd(IJ, K) > 0.8 && M >= NFRE / 2) {
          const double c2 = cwd(IJ, K) * cwd(IJ, K);
          F1(IJ, K, M) = std::max(FL1(IJ, K, M) - F1(IJ, K, M) + t.EPSMIN * (c2 * c2), 0.0);
        } else F1(IJ, K, M) = std::max(FL1(IJ, K, M) - F1(IJ, K, M), 0.0);
      }
  femean_out(t, KIJL, NANG, NFRE, F1.view(), o.ESEA, o.FSEA);
  sthq(t, KIJL, NANG, NFRE, F1.view(), o.THWISEA);
  for (int IJ = 1; IJ <= KIJL; ++IJ) if (o.ESEA[IJ] <= 1.0e-9) o.THWISEA[IJ] = WDWAVE[IJ - 1];
  mwp12(t, KIJL, NANG, 1, F1.view(), o.P1SEA);
  mwp12(t, KIJL, NANG, 2, F1.view(), o.P2SEA);
  wdirspread(t, KIJL, NANG, NFRE, F1.view(), o.ESEA, true, o.SPRDSEA);
}

// weflux.F90:70-140
void weflux(const Tables& t, int KIJL, int NANG, int NFRE, const S3& FL1, const double* CGROUP, V& WEFMAG, V& WEFDIR) {
  V TEMP(KIJL + 1), TEMPX(KIJL + 1), TEMPY(KIJL + 1), WEFX(KIJL + 1, 0.0), WEFY(KIJL + 1, 0.0);
  for (int IJ = 1; IJ <= KIJL; ++IJ) WEFMAG[IJ] = 0.0;
  const double ROG = t.ROWATER * t.G;
  const double DELT = t.FRTAIL * t.DELTH * t.G / (2.0 * t.ZPI);
  for (int M = 1; M <= NFRE + 1; ++M) {   // M = NFRE+1: the tail block (weflux.F90:108-130) without CGROUP
    const bool tail = M == NFRE + 1;
    const int MM = tail ? NFRE : M;
    for (int IJ = 1; IJ <= KIJL; ++IJ) {
      const double FCG = tail ? FL1(IJ, 1, MM) : FL1(IJ, 1, MM) * CGROUP[(size_t)(MM - 1) * KIJL + IJ - 1];
      TEMP[IJ] = FCG; TEMPX[IJ] = FCG * t.SINTH(1); TEMPY[IJ] = FCG * t.COSTH(1);
    }
    for (int K = 2; K <= NANG; ++K)
      for (int IJ = 1; IJ <= KIJL; ++IJ) {
        const double FCG = tail ? FL1(IJ, K, MM) : FL1(IJ, K, MM) * CGROUP[(size_t)(MM - 1) * KIJL + IJ - 1];
        TEMP[IJ] = TEMP[IJ] + FCG; TEMPX[IJ] = TEMPX[IJ] + FCG * t.SINTH(K); TEMPY[IJ] = TEMPY[IJ] + FCG * t.COSTH(K);
      }
    const double w = tail ? DELT : t.DFIM(MM);
    for (int IJ = 1; IJ <= KIJL; ++IJ) { WEFMAG[IJ] = WEFMAG[IJ] + w * TEMP[IJ]; WEFX[IJ] = WEFX[IJ] + w * TEMPX[IJ]; WEFY[IJ] = WEFY[IJ] + w * TEMPY[IJ]; }
  }
  for (int IJ = 1; IJ <= KIJL; ++IJ) WEFMAG[IJ] = ROG * WEFMAG[IJ];
  for (int IJ = 1; IJ <= KIJL; ++IJ) if (WEFY[IJ] == 0.0) WEFY[IJ] = t.EPSMIN;
  for (int IJ = 1; IJ <= KIJL; ++IJ) WEFDIR[IJ] = std::atan2(WEFX[IJ], WEFY[IJ]);
  for (int IJ = 1; IJ <= KIJL; ++IJ) if (WEFDIR[IJ] < 0.0) WEFDIR[IJ] = WEFDIR[IJ] + t.ZPI;
}
// sebtmean.F90:66-198: energy between the periods TB and TT (trapezoid over the frequencies inside the band, linear
// interpolation at its two ends, f**-5 tail above FR(NFRE))
void sebtmean(const Tables& t, int KIJL, int NANG, int NFRE, const S3& FL1, double TB, double TT, V& EBT) {
  std::vector<double> FRLOC(NFRE + 2, 0.0), F1D((size_t)(KIJL + 1) * (NFRE + 2), 0.0);
  auto f1 = [&](int IJ, int M) -> double& { return F1D[(size_t)M * (KIJL + 1) + IJ]; };
  double FBOT = 1.0 / std::max(TT, t.EPSMIN);
  const double FCUTB_FT = std::min(FBOT, t.FR(NFRE));
  const double FCUTB = std::max(t.FR(1), FCUTB_FT);
  FBOT = std::max(FBOT, t.FR(NFRE));
  int MCUTB = 1;
  while (t.FR(MCUTB) < FCUTB && MCUTB < NFRE) MCUTB = MCUTB + 1;
  double FTOP = 1.0 / std::max(TB, t.EPSMIN);
  const double FCUTT = std::max(t.FR(1), std::min(FTOP, t.FR(NFRE)));
  FTOP = std::max(FTOP, t.FR(NFRE));
  int MCUTT = NFRE;
  while (t.FR(MCUTT) > FCUTT && MCUTT > 1) MCUTT = MCUTT - 1;
  if (FCUTB == FCUTT) MCUTT = MCUTB - 1;
  for (int IJ = 1; IJ <= KIJL; ++IJ) EBT[IJ] = t.EPSMIN;
  if (MCUTB > 1) {
    FRLOC[MCUTB - 1] = FCUTB;
    const double WL = (t.FR(MCUTB) - FCUTB) / (t.FR(MCUTB) - t.FR(MCUTB - 1)), WR = 1.0 - WL;
    for (int IJ = 1; IJ <= KIJL; ++IJ) f1(IJ, MCUTB - 1) = (WL * FL1(IJ, 1, MCUTB - 1) + WR * FL1(IJ, 1, MCUTB)) * t.DELTH;
    for (int K = 2; K <= NANG; ++K)
      for (int IJ = 1; IJ <= KIJL; ++IJ) f1(IJ, MCUTB - 1) = f1(IJ, MCUTB - 1) + (WL * FL1(IJ, K, MCUTB - 1) + WR * FL1(IJ, K, MCUTB)) * t.DELTH;
  }
  for (int M = MCUTB; M <= MCUTT; ++M) {
    FRLOC[M] = t.FR(M);
    for (int IJ = 1; IJ <= KIJL; ++IJ) f1(IJ, M) = FL1(IJ, 1, M) * t.DELTH;
    for (int K = 2; K <= NANG; ++K) for (int IJ = 1; IJ <= KIJL; ++IJ) f1(IJ, M) = f1(IJ, M) + FL1(IJ, K, M) * t.DELTH;
  }
  if (MCUTT < NFRE) {
    FRLOC[MCUTT + 1] = FCUTT;
    // MCUTT = 0 (band entirely below FR(1), e.g. 25-30 s on the 25-frequency grid): the reference reads FR(0) and FL1(:,:,0)
    // out of bounds with WL = (FR(1)-FCUTT)/(FR(1)-FR(0)) = 0/x; the restatement takes that limit (WL = 0, F1D(:,1) = row 1)
    const double WL = MCUTT >= 1 ? (t.FR(MCUTT + 1) - FCUTT) / (t.FR(MCUTT + 1) - t.FR(MCUTT)) : 0.0, WR = 1.0 - WL;
    auto lo = [&](int IJ, int K) { return MCUTT >= 1 ? WL * FL1(IJ, K, MCUTT) : 0.0; };
    for (int IJ = 1; IJ <= KIJL; ++IJ) f1(IJ, MCUTT + 1) = (lo(IJ, 1) + WR * FL1(IJ, 1, MCUTT + 1)) * t.DELTH;
    for (int K = 2; K <= NANG; ++K)
      for (int IJ = 1; IJ <= KIJL; ++IJ) f1(IJ, MCUTT + 1) = f1(IJ, MCUTT + 1) + (lo(IJ, K) + WR * FL1(IJ, K, MCUTT + 1)) * t.DELTH;
  }
  for (int M = std::max(MCUTB - 1, 1); M <= std::min(MCUTT, NFRE - 1); ++M) {
    const double DF = 0.5 * (FRLOC[M + 1] - FRLOC[M]);
    for (int IJ = 1; IJ <= KIJL; ++IJ) EBT[IJ] = EBT[IJ] + DF * (f1(IJ, M + 1) + f1(IJ, M));
  }
  if (FCUTB_FT < FCUTB && FCUTB == t.FR(1)) {
    const double WL = (t.FR(1) - FCUTB_FT) / t.FR(1), WR = 1.0 - WL;
    const double DF = 0.5 * (t.FR(1) - FCUTB_FT) * (1.0 + WR);
    for (int IJ = 1; IJ <= KIJL; ++IJ) EBT[IJ] = EBT[IJ] + DF * f1(IJ, 1);
  }
  if (FBOT < FTOP) {
    const double ZW = 0.25 * t.FR5(NFRE) * (1.0 / (FBOT * FBOT * FBOT * FBOT) - 1.0 / (FTOP * FTOP * FTOP * FTOP));
    for (int IJ = 1; IJ <= KIJL; ++IJ) f1(IJ, NFRE) = FL1(IJ, 1, NFRE) * t.DELTH;
    for (int K = 2; K <= NANG; ++K) for (int IJ = 1; IJ <= KIJL; ++IJ) f1(IJ, NFRE) = f1(IJ, NFRE) + FL1(IJ, K, NFRE) * t.DELTH;
    for (int IJ = 1; IJ <= KIJL; ++IJ) EBT[IJ] = EBT[IJ] + ZW * f1(IJ, NFRE);
  }
}
}  // namespace

// newwind.F90:105-167 for ICODE_WND = 3: FF_NOW <- FF_NEXT with the low-wind cap on the first-guess wave stress
void newwind(Model& m, int ir, const Fields& nx) {
  const Tables& t = m.tab;
  Fields& f = m.fld[ir];
  const double WSPMIN_RESET_TAUW = 4.0;   // yowwind.F90:19
  const double WGHT = 1.0 / std::max(WSPMIN_RESET_TAUW, t.EPSMIN);
  const RankDecomp& r = m.ranks[ir];
  const double USTMIN_RESET_TAUW = 0.08;  // yowwind.F90:20
  for (int ICHNK = 1; ICHNK <= r.NCHNK; ++ICHNK) {
    if (m.cfg.icode != 3) {   // newwind.F90:141-150: the forcing is the friction velocity (handed over in nx.WSWAVE's place)
      for (int IJ = 1; IJ <= r.NPROMA; ++IJ) {
        f.UFRIC(IJ, ICHNK) = nx.WSWAVE.d[(IJ - 1) + (size_t)r.NPROMA * (ICHNK - 1)];
        const double q = t.ALPHA / f.CHRNCK(IJ, ICHNK);
        f.TAUW(IJ, ICHNK) = f.UFRIC(IJ, ICHNK) * f.UFRIC(IJ, ICHNK) * (1.0 - q * q);
        if (f.UFRIC(IJ, ICHNK) < USTMIN_RESET_TAUW) f.TAUW(IJ, ICHNK) = 0.0;
      }
    } else
    for (int IJ = 1; IJ <= r.NPROMA; ++IJ) {
      f.WSWAVE(IJ, ICHNK) = nx.WSWAVE.d[(IJ - 1) + (size_t)r.NPROMA * (ICHNK - 1)];
      if (f.WSWAVE(IJ, ICHNK) < WSPMIN_RESET_TAUW) {
        const double w = f.WSWAVE(IJ, ICHNK);
        const double TLWMAX = WGHT * (t.ACD + t.BCD * w) * (w * w * w);
        f.TAUW(IJ, ICHNK) = std::min(f.TAUW(IJ, ICHNK), TLWMAX);
      }
    }
    for (int IJ = 1; IJ <= r.NPROMA; ++IJ) {
      const size_t o = (IJ - 1) + (size_t)r.NPROMA * (ICHNK - 1);
      f.WDWAVE(IJ, ICHNK) = nx.WDWAVE.d[o]; f.AIRD(IJ, ICHNK) = nx.AIRD.d[o]; f.WSTAR(IJ, ICHNK) = nx.WSTAR.d[o];
      f.CICOVER(IJ, ICHNK) = nx.CICOVER.d[o]; f.CITHICK(IJ, ICHNK) = nx.CITHICK.d[o];
      f.USTRA(IJ, ICHNK) = nx.USTRA.d[o]; f.VSTRA(IJ, ICHNK) = nx.VSTRA.d[o];
    }
  }
}

// meansqs.F90:80-100 with halphap.F90:68-115, meansqs_gc.F90:60-82 (OMEGAGC / NS_GC) and meansqs_lf.F90:80-100
void meansqs(const Config& c, const Tables& t, double XKMSS, int KIJL, int NANG, int NFRE, const S3& F, const double* WAVNUM /*(KIJL,NFRE)*/,
             const double* USTAR, const V& COSWDIF /*(K-1)*KIJL+IJ*/, double* XMSS /*1-based*/) {
  auto WN = [&](int IJ, int M) { return WAVNUM[(IJ - 1) + (size_t)KIJL * (M - 1)]; };
  // HALPHAP
  V HALP(KIJL + 1), XM(KIJL + 1, 0.0), EM(KIJL + 1), FM(KIJL + 1);
  W3 FLWD(KIJL, NANG, NFRE);
  for (int M = 1; M <= NFRE; ++M) for (int K = 1; K <= NANG; ++K) for (int IJ = 1; IJ <= KIJL; ++IJ) {
    const double WD = 0.5 + 0.5 * std::copysign(1.0, COSWDIF[(size_t)(K - 1) * KIJL + IJ]);
    FLWD(IJ, K, M) = F(IJ, K, M) * WD;
  }
  auto lf = [&](int NFRE_EFF, const S3& G, V& OUT) {       // MEANSQS_LF
    const int KFRE = std::min(NFRE_EFF, NFRE);
    for (int IJ = 1; IJ <= KIJL; ++IJ) OUT[IJ] = 0.0;
    for (int M = 1; M <= KFRE; ++M)
      for (int IJ = 1; IJ <= KIJL; ++IJ) {
        const double TEMP1 = t.DFIM(M) * (WN(IJ, M) * WN(IJ, M));
        double TEMP2 = 0.0;
        for (int K = 1; K <= NANG; ++K) TEMP2 = TEMP2 + G(IJ, K, M);
        OUT[IJ] = OUT[IJ] + TEMP1 * TEMP2;
      }
  };
  lf(NFRE, FLWD.view(), XM);
  femean_out(t, KIJL, NANG, NFRE, FLWD.view(), EM, FM);
  const double ZLNFRNFRE = std::log(t.FR(NFRE));
  for (int IJ = 1; IJ <= KIJL; ++IJ) {
    double ALPHAP = 0.0;
    bool tail = true;
    if (EM[IJ] > 0.0 && FM[IJ] < t.FR(NFRE - 2)) { ALPHAP = XM[IJ] / (ZLNFRNFRE - std::log(FM[IJ])); tail = ALPHAP > t.ALPHAPMAX; }
    if (tail) {
      double F1D = 0.0;
      for (int K = 1; K <= NANG; ++K) F1D = F1D + FLWD(IJ, K, NFRE) * t.DELTH;
      ALPHAP = t.ZPI4GM2 * t.FR5(NFRE) * F1D;
    }
    HALP[IJ] = 0.5 * std::min(ALPHAP, t.ALPHAPMAX);
  }
  // MEANSQS_GC
  const double XLOGKRATIOM1_GC = 1.0 / std::log(1.2);
  const int NE = std::min(std::max((int)nint(std::log(XKMSS * t.XKM_GC(1)) * XLOGKRATIOM1_GC), 1), t.NWAV_GC);
  V FRGC(KIJL + 1);
  for (int IJ = 1; IJ <= KIJL; ++IJ) {
    const double XKS0 = t.SQRTGOSURFT / (1.48 + 2.05 * USTAR[IJ - 1]);                                  // ns_gc.F90:44-48
    int NS = std::min((int)(std::log(std::max(XKS0 * t.XKM_GC(1), 1.0)) * XLOGKRATIOM1_GC) + 1, t.NWAV_GC - 1);
    const double XKS = t.XK_GC(NS), OMS = t.OMEGA_GC(NS);                                                // omegagc.F90:51-55
    FRGC[IJ] = OMS / t.ZPI;
    double XMSSCG;
    if (XKS > XKMSS) { NS = NE; XMSSCG = 0.0; }
    else XMSSCG = t.DELKCC_GC_NS(NS) * t.XKM_GC(NS);
    for (int I = NS + 1; I <= NE; ++I) XMSSCG = XMSSCG + t.DELKCC_GC(I) * t.XKM_GC(I);
    const double COEF = t.C2OSQRTVG_GC(NS) * HALP[IJ];
    XMSS[IJ] = XMSSCG * COEF;
  }
  const double FCUT = std::sqrt(t.G * XKMSS) / t.ZPI;
  const int NFRE_MSS = (int)(std::log(FCUT / t.FR(1)) / std::log(t.FRATIO)) + 1;
  const int NFRE_EFF = std::min(NFRE, NFRE_MSS);
  V XMSSLF(KIJL + 1);
  lf(NFRE_EFF, F, XMSSLF);
  const double XLOGFS = std::log(t.FR(NFRE_EFF));
  for (int IJ = 1; IJ <= KIJL; ++IJ) {
    XMSS[IJ] = XMSS[IJ] + XMSSLF[IJ];
    const double XMSS_TAIL = 2.0 * HALP[IJ] * std::max(std::log(std::min(FRGC[IJ], FCUT)) - XLOGFS, 0.0);
    XMSS[IJ] = XMSS[IJ] + XMSS_TAIL;
  }
}

// ---- the KURTOSIS family of OUTBLOCK (outblock.F90:208-211): skewness, kurtosis, Benjamin-Feir index, Goda's peakedness,
//      expected maximum wave height and its period.  Oracle only so far: the product rejects these parameters.
// transf_r.F90:57-82
static double transf_r(const Tables& t, const Config& c, double XK0, double D) {
  const double EPS = 0.0001, XKDMIN = 0.75;   // yowshal.F90:23
  if (D < c.bathymax && D > 0.0 && XK0 > 0.0) {
    double X = XK0 * D;
    if (X > t.DKMAX) return 0.5;
    const double XK = std::max(XK0, XKDMIN / D);
    X = XK * D;
    const double T_0 = std::tanh(X), T_0_SQ = T_0 * T_0;
    const double OM = std::sqrt(t.G * XK * T_0), C_0 = OM / XK;
    const double V_G = X < EPS ? C_0 : 0.5 * C_0 * (1.0 + 2.0 * X / std::sinh(2.0 * X));
    const double a = T_0 - X * (1.0 - T_0_SQ);
    const double D2OM = a * a + 4.0 * (X * X) * T_0_SQ * (1.0 - T_0_SQ);
    const double q = V_G / C_0;
    return 4.0 * (q * q * q) * T_0_SQ / D2OM;
  }
  return 0.5;
}
// transf_bfi.F90:62-105
static double transf_bfi(const Tables& t, const Config& c, double XK0, double D, double XNU, double SIG_TH) {
  const double EPS = 0.0001, XKDMIN = 0.75, TMIN = -4.0, TMAX = 4.0;
  if (D < c.bathymax && D > 0.0) {
    double X = XK0 * D;
    if (X > t.DKMAX) return 1.0;
    const double XK = std::max(XK0, XKDMIN / D);
    X = XK * D;
    const double T_0 = std::tanh(X), T_0_SQ = T_0 * T_0;
    const double OM = std::sqrt(t.G * XK * T_0), C_0 = OM / XK, C_S_SQ = t.G * D;
    const double V_G = X < EPS ? C_0 : 0.5 * C_0 * (1.0 + 2.0 * X / std::sinh(2.0 * X));
    const double V_G_SQ = V_G * V_G;
    const double a = T_0 - X * (1.0 - T_0_SQ);
    const double D2OM = a * a + 4.0 * (X * X) * T_0_SQ * (1.0 - T_0_SQ);
    const double XNL_1 = (9.0 * (T_0_SQ * T_0_SQ) - 10.0 * T_0_SQ + 9.0) / (8.0 * T_0_SQ * T_0);
    const double b = 2.0 * V_G - 0.5 * C_0;
    const double XNL_2 = (b * b / (t.G * D - V_G_SQ) + 1.0) / X;
    const double e = 2.0 * C_0 + V_G * (1.0 - T_0_SQ);
    const double XNL_4 = 1. / (4.0 * T_0) * (e * e) / (C_S_SQ - V_G_SQ);
    const double ALP = (1.0 - V_G_SQ / C_S_SQ) * (C_0 * C_0) / V_G_SQ;
    const double ZFAC = (SIG_TH * SIG_TH) / (SIG_TH * SIG_TH + ALP * (XNU * XNU));
    const double XNL_3 = ZFAC * XNL_4;
    const double T_NL = XNL_1 - XNL_2 + XNL_3;
    const double q = V_G / C_0;
    return std::max(std::min(TMAX, 4.0 * (q * q) * T_NL * T_0 / D2OM), TMIN);
  }
  return 1.0;
}
struct KurtOut { V C3, C4, BF2, QP, HMAX, TMAX, ETA_M, R, XNSLC, SIG_TH, EPS, XNU; };
void kurtosis(const Config& c, const Tables& t, int KIJL, int NANG, int NFRE, const S3& FL1, const double* DEPTH, KurtOut& o) {
  for (V* v : {&o.C3, &o.C4, &o.BF2, &o.QP, &o.HMAX, &o.TMAX, &o.ETA_M, &o.R, &o.XNSLC, &o.SIG_TH, &o.EPS, &o.XNU}) v->assign(KIJL + 1, 0.0);
  const double ZEPSILON = 10.0 * 2.220446049250313e-16, ZSQREPSILON = std::sqrt(ZEPSILON);
  const double DELT25 = t.WETAIL * t.FR(NFRE) * t.DELTH, COEF_FR1 = t.WP1TAIL * t.DELTH * (t.FR(NFRE) * t.FR(NFRE));
  const double COEF_FR2 = t.WP2TAIL * t.DELTH * (t.FR(NFRE) * t.FR(NFRE) * t.FR(NFRE)), DELT2 = t.FRTAIL * t.DELTH;
  auto DFIMFR2 = [&](int M) { return t.DFIM(M) * (t.FR(M) * t.FR(M)); };   // initmdl.F90:447
  // ---- PEAK_ANG (peak_ang.F90:72-175): spectral width XNU and the angular width SIG_TH around the peak
  {
    const int NSH = 1 + (int)(std::log(1.5) / std::log(t.FRATIO));
    for (int IJ = 1; IJ <= KIJL; ++IJ) {
      double SUM0 = ZEPSILON, SUM1 = 0.0, SUM2 = 0.0, TEMP = 0.0;
      for (int M = 1; M <= NFRE; ++M) {
        TEMP = FL1(IJ, 1, M);
        for (int K = 2; K <= NANG; ++K) TEMP = TEMP + FL1(IJ, K, M);
        SUM0 = SUM0 + TEMP * t.DFIM(M); SUM1 = SUM1 + TEMP * t.DFIMFR(M); SUM2 = SUM2 + TEMP * DFIMFR2(M);
      }
      SUM0 = SUM0 + DELT25 * TEMP; SUM1 = SUM1 + COEF_FR1 * TEMP; SUM2 = SUM2 + COEF_FR2 * TEMP;
      o.XNU[IJ] = SUM0 > ZEPSILON ? std::sqrt(std::max(ZEPSILON, SUM2 * SUM0 / (SUM1 * SUM1) - 1.0)) : ZEPSILON;
      double XMAX = 0.0;
      int MMAX = 2;
      for (int M = 2; M <= NFRE - 1; ++M) for (int K = 1; K <= NANG; ++K) if (FL1(IJ, K, M) > XMAX) { MMAX = M; XMAX = FL1(IJ, K, M); }
      double S1 = ZEPSILON, S2 = 0.0, SUM_S = 0.0, SUM_C = ZEPSILON, THMEAN = 0.0;
      for (int M = std::max(1, MMAX - NSH); M <= std::min(NFRE, MMAX + NSH); ++M) {
        for (int K = 1; K <= NANG; ++K) { SUM_S = SUM_S + t.SINTH(K) * FL1(IJ, K, M); SUM_C = SUM_C + t.COSTH(K) * FL1(IJ, K, M); }
        THMEAN = std::atan2(SUM_S, SUM_C);
        for (int K = 1; K <= NANG; ++K) {
          S1 = S1 + FL1(IJ, K, M) * t.DFIM(M);
          S2 = S2 + std::cos(t.TH(K) - THMEAN) * FL1(IJ, K, M) * t.DFIM(M);
        }
      }
      o.SIG_TH[IJ] = S1 > ZEPSILON ? std::sqrt(2.0 * (1.0 - S2 / S1)) : 0.0;
    }
  }
  // ---- KURTOSIS (kurtosis.F90:239-350)
  const double CONST_SIG_SQRTPIM1 = 1.0 / std::sqrt(t.PI), CONST_OM_ZPI = 0.89 * t.ZPI, FRMAX = t.FR(NFRE), FRMIN = t.FR(1);
  const double QPMIN = 0.5, QPMAX = 15.0, BF2MIN = -5.0, BF2MAX = 5.0, FLTHRS = 0.4;
  V SUM0(KIJL + 1), SUM1(KIJL + 1), F_M(KIJL + 1), XKP(KIJL + 1);
  std::vector<double> FF(NFRE + 1);
  for (int IJ = 1; IJ <= KIJL; ++IJ) {
    for (int M = 1; M <= NFRE; ++M) { FF[M] = FL1(IJ, 1, M); for (int K = 2; K <= NANG; ++K) FF[M] = FF[M] + FL1(IJ, K, M); }
    double FFMAX = FF[1];
    for (int M = 2; M <= NFRE; ++M) FFMAX = std::max(FFMAX, FF[M]);
    double S0 = ZEPSILON, S1 = 0.0, S2 = 0.0, S6 = 0.0;
    for (int M = 1; M <= NFRE; ++M) { S0 = S0 + FF[M] * t.DFIM(M); S1 = S1 + FF[M] * t.DFIMFR(M); S2 = S2 + FF[M] * DFIMFR2(M); S6 = S6 + FF[M] * t.DFIMOFR(M); }
    S0 = S0 + DELT25 * FF[NFRE]; S1 = S1 + COEF_FR1 * FF[NFRE]; S2 = S2 + COEF_FR2 * FF[NFRE]; S6 = S6 + DELT2 * FF[NFRE];
    (void)S2;
    double S40 = ZSQREPSILON, S4 = 0.0;
    FFMAX = FLTHRS * FFMAX;
    for (int M = 1; M <= NFRE; ++M) if (FF[M] > FFMAX) { S40 = S40 + FF[M] * t.DFIM(M); S4 = S4 + (FF[M] * FF[M]) * (2.0 * t.DELTH * t.DFIMFR(M)); }
    SUM0[IJ] = S0; SUM1[IJ] = S1;
    if (S1 > ZSQREPSILON && S0 > ZEPSILON) {
      F_M[IJ] = std::max(std::min(S1 / S0, FRMAX), FRMIN);
      o.QP[IJ] = std::max(std::min(S4 / (S40 * S40), QPMAX), QPMIN);
      const double SIG_OM = CONST_SIG_SQRTPIM1 / o.QP[IJ];
      const double OM_MEAN = CONST_OM_ZPI * std::max(std::min(S0 / S6, FRMAX), FRMIN);
      XKP[IJ] = aki(t, OM_MEAN, DEPTH[IJ - 1]);
      o.EPS[IJ] = XKP[IJ] * std::sqrt(S0);
      const double TRANS = transf_bfi(t, c, XKP[IJ], DEPTH[IJ - 1], o.XNU[IJ], o.SIG_TH[IJ]);
      const double r = o.EPS[IJ] / std::max(SIG_OM, ZEPSILON);
      o.BF2[IJ] = std::max(std::min(2.0 * TRANS * (r * r), BF2MAX), BF2MIN);
    } else {
      F_M[IJ] = 0.0; o.QP[IJ] = 0.0;
      const double OM_MEAN = CONST_OM_ZPI * FRMAX;
      XKP[IJ] = OM_MEAN * OM_MEAN / t.G; o.EPS[IJ] = 0.0; o.BF2[IJ] = 0.0;
    }
  }
  // ---- STAT_NL (stat_nl.F90:74-170)
  {
    const double EPS = 0.0001, XKDMIN = 0.75, RMIN = 0.0, RMAX = 16.0, C3MIN = 0.0, C3MAX = 0.25, C4MIN = -0.25, C4MAX = 0.25;
    const double CONST_C3 = 1.12 * 2.0, CONST_C4 = 0.93 * 8.0, C3DELTA_ADJ = 0.9, SQRT3 = std::sqrt(3.0);
    const double C4_CONST = 0.9 * t.PI / (3.0 * SQRT3), ZC1 = 4.0 * SQRT3 / t.PI, ZC2 = (1.0 / 3.0 + 2.0 * SQRT3 / t.PI), ZC3 = 2.0 * SQRT3 / t.PI - 4.0 / 3.0;
    for (int IJ = 1; IJ <= KIJL; ++IJ) {
      const double TRANSF = transf_r(t, c, XKP[IJ], DEPTH[IJ - 1]);
      const double D = DEPTH[IJ - 1], XM0 = SUM0[IJ];
      if (XM0 > ZEPSILON && D > 0.0 && XKP[IJ] > 0.0) {
        const double XK = std::max(XKP[IJ], XKDMIN / D), X = XK * D, T0 = std::tanh(X), OM = std::sqrt(t.G * XK * T0), T0_SQ = T0 * T0;
        const double ALPH = XK / (4.0 * T0_SQ * T0) * (3.0 - T0_SQ), GAM = -0.5 * (ALPH * ALPH);
        const double C_0 = OM / XK, C_S_SQ = t.G * D;
        double V_G;
        if (X > t.DKMAX) V_G = 0.5 * C_0; else if (X < EPS) V_G = C_0; else V_G = 0.5 * C_0 * (1.0 + 2.0 * X / std::sinh(2.0 * X));
        const double V_G_SQ = V_G * V_G;
        const double ZFAC = -0.25 * XK * C_S_SQ / (C_S_SQ - V_G_SQ);
        const double DELTA_1D = ZFAC * (2.0 * (1.0 - T0_SQ) / T0 + 1.0 / X);
        const double ZFAC1 = 0.5 * C_0 * C_S_SQ * V_G / T0;
        const double XKAPPA1 = ZFAC1 * (2.0 * C_0 + V_G * (1.0 - T0_SQ)) / (C_S_SQ - V_G_SQ);
        const double ALPHA = (1.0 - V_G_SQ / C_S_SQ) * (C_0 * C_0) / V_G_SQ;
        const double st2 = o.SIG_TH[IJ] * o.SIG_TH[IJ];
        const double ZFAC2 = st2 / (st2 + ALPHA * (o.XNU[IJ] * o.XNU[IJ]));
        const double DELTA_2D = 0.5 * (XK * XK) * XKAPPA1 / (OM * C_S_SQ) * ZFAC2;
        const double DELTA = DELTA_1D + DELTA_2D;
        o.ETA_M[IJ] = 2.0 * XM0 * DELTA;
        o.C3[IJ] = std::max(std::min(C3MAX, CONST_C3 * std::sqrt(XM0) * (ALPH + C3DELTA_ADJ * DELTA)), C3MIN);
        const double C4_B = CONST_C4 * XM0 * (GAM + ALPH * ALPH + (ALPH + DELTA) * (ALPH + DELTA));
        const double q = o.SIG_TH[IJ] / o.XNU[IJ];
        o.R[IJ] = std::max(std::min(TRANSF * (q * q), RMAX), RMIN);
        const double ZR = o.R[IJ];
        const double XJ = ZR > 1.0 ? -C4_CONST / ZR * (1.0 - ZC1 / std::sqrt(ZR) + ZC2 / ZR + ZC3 / (ZR * ZR))
                                   : C4_CONST * (1.0 - ZC1 * std::sqrt(ZR) + ZC2 * ZR + ZC3 * (ZR * ZR));
        o.C4[IJ] = std::max(std::min(C4MAX, XJ * o.BF2[IJ] + C4_B), C4MIN);
      }
    }
  }
  // ---- number of waves, H_MAX (h_max.F90:60-110), TMAX, HMAX (kurtosis.F90:360-403)
  const double ZFACN = 2.0 * t.ZPI / std::sqrt(t.ZPI), DUR = 1200.0;
  const double GAMC = 0.5772, EB = 10.0, FLOGMIN = 0.1, AA_MAX = 1000.0, H_C = 2.0, H_C_MIN = 1.0, H_C_MAX = 4.0;
  const double TWOG1 = -2.0 * GAMC, G2 = GAMC * GAMC + t.PI * t.PI / 6.0, AE = 0.5 * EB * (EB - 2.0), BE = 0.5 * EB * (EB * EB - 6.0 * EB + 6.0);
  const double EMIN = 2.0 * H_C_MIN * H_C_MIN, EMAX = 2.0 * H_C_MAX * H_C_MAX, EVAL = 2.0 * H_C * H_C;
  for (int IJ = 1; IJ <= KIJL; ++IJ) {
    o.XNSLC[IJ] = F_M[IJ] > 0.0 ? (double)nint(DUR * (ZFACN * o.XNU[IJ] * F_M[IJ])) : 0.0;
    double HMAXN, E = EVAL;
    const double DFNORMA = o.C4[IJ] * AE + (o.C3[IJ] * o.C3[IJ]) * BE;
    if (o.XNSLC[IJ] > 0.0 && std::fabs(DFNORMA) > ZEPSILON) {
      const double F = std::log(std::max(1.0 + DFNORMA, FLOGMIN));
      const double AA = std::min(((EB - F) * (EB - F) - 2.0 * EB) / (2.0 * F), AA_MAX), BB = 2.0 * (1.0 + AA);
      const double BBM1 = 1.0 / (BB + ZEPSILON * std::copysign(1.0, BB));
      for (int I = 1; I <= 5; ++I) {
        const double Z0 = std::log(o.XNSLC[IJ] * std::sqrt(0.5 * E));
        E = (G2 - TWOG1 * (AA + Z0) + (2.0 * AA + Z0) * Z0) * BBM1;
        E = std::min(std::max(E, EMIN), EMAX);
      }
      HMAXN = std::sqrt(0.5 * E);
    } else HMAXN = H_C_MIN;
    if (SUM1[IJ] > ZEPSILON && HMAXN > ZEPSILON) {
      const double ZEPS = o.XNU[IJ] / (std::sqrt(2.0) * HMAXN), z2 = ZEPS * ZEPS;
      o.TMAX[IJ] = (SUM0[IJ] / SUM1[IJ]) * (1.0 + 0.5 * z2 + 0.75 * (z2 * z2));
    } else o.TMAX[IJ] = 0.0;
    o.HMAX[IJ] = SUM0[IJ] > 0.0 ? HMAXN * (4.0 * std::sqrt(SUM0[IJ])) : 0.0;
    if (SUM0[IJ] <= 0.0) o.XNU[IJ] = 0.0;
  }
}

bool outparam_supported(int itg) {
  static const int ok[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 20, 21, 22, 23, 24, 25, 26, 27, 28, 32, 35, 36, 37, 38,
                           39, 40, 41, 52, 53, 54, 55, 56, 62, 63, 64, 65, 66, 67, 68, 69, 73, 74, 75, 76, 77,
                           29, 30, 31, 33, 34, 57, 70, 71, 72};   // the KURTOSIS family: oracle only so far
  for (int v : ok) if (v == itg) return true;
  return false;
}

// intpol.F90:96-271: the spectrum on the absolute (ground-based) frequency axis from the one relative to the current (IRA = 1):
// every bin moves to F + k.U / 2 pi and is shared between the two neighbouring frequency bins, energy-conserving on the
// FR(M)*CDF intervals, with an f**-5 extension above FR(NFRE) up to what CURRENT_MAX can shift back into the grid.
static void intpol(const Tables& t, int KIJL, int NANG, int NFRE, const S3& FLR, W3& FLA, const double* WAVNUM /*(KIJL,NFRE)*/,
                   const double* UCUR, const double* VCUR, int IRA) {
  const double CURRENT_MAX = 1.5;     // yowcurr.F90:18
  const double FRE0 = t.FRATIO - 1.0, ZPI2GM = t.ZPI * t.ZPI / t.G, COEF = IRA / t.ZPI;
  const double FMAX = t.FR(NFRE) + (t.ZPI / t.G) * (t.FR(NFRE) * t.FR(NFRE)) * CURRENT_MAX;
  const int NFRE_MAX = (int)std::floor(std::log10(FMAX / t.FR(1)) * t.FLOGSPRDM1) + 1;
  const double CDF = 0.5 * (t.FRATIO - 1.0 / t.FRATIO) * t.DELTH;
  V DFTH(NFRE + 1);
  for (int M = 1; M <= NFRE; ++M) DFTH[M] = t.FR(M) * CDF;
  std::vector<char> LICE2SEA(KIJL + 1, 1);
  for (int K = 1; K <= NANG; ++K) for (int M = 1; M <= NFRE; ++M) for (int IJ = 1; IJ <= KIJL; ++IJ) if (FLR(IJ, K, M) > t.EPSMIN) LICE2SEA[IJ] = 0;
  for (int M = 1; M <= NFRE; ++M) for (int K = 1; K <= NANG; ++K) for (int IJ = 1; IJ <= KIJL; ++IJ) FLA(IJ, K, M) = 0.0;
  std::vector<int> NEWF(KIJL + 1), NEWFLA(KIJL + 1), KNEW(KIJL + 1);
  V WAVN(KIJL + 1), FNEF(KIJL + 1), OLDFL(KIJL + 1), GWP(KIJL + 1), GWM(KIJL + 1);
  for (int M = 1; M <= NFRE_MAX; ++M) {
    double FREQ, DFREQTH;
    if (M <= NFRE) {
      FREQ = t.FR(M); DFREQTH = DFTH[M];
      for (int IJ = 1; IJ <= KIJL; ++IJ) WAVN[IJ] = WAVNUM[(IJ - 1) + (size_t)KIJL * (M - 1)];
    } else {
      FREQ = t.FR(NFRE) * std::pow(t.FRATIO, M - NFRE); DFREQTH = FREQ * CDF;
      for (int IJ = 1; IJ <= KIJL; ++IJ) WAVN[IJ] = ZPI2GM * (FREQ * FREQ);
    }
    const double FR5OFREQ5 = t.FR5(NFRE) / std::pow(FREQ, 5);
    for (int K = 1; K <= NANG; ++K) {
      for (int IJ = 1; IJ <= KIJL; ++IJ) {
        FNEF[IJ] = FREQ + COEF * WAVN[IJ] * (t.COSTH(K) * VCUR[IJ - 1] + t.SINTH(K) * UCUR[IJ - 1]);
        if (FNEF[IJ] > 0.0) KNEW[IJ] = K;
        else { KNEW[IJ] = (K + NANG / 2 - 1) % NANG + 1; FNEF[IJ] = -FNEF[IJ]; }
      }
      for (int IJ = 1; IJ <= KIJL; ++IJ) {
        if (FNEF[IJ] <= t.FR(1) / t.FRATIO) NEWF[IJ] = -1;
        else NEWF[IJ] = (int)std::floor(std::log10(FNEF[IJ] / t.FR(1)) * t.FLOGSPRDM1) + 1;
      }
      for (int IJ = 1; IJ <= KIJL; ++IJ) {
        if (LICE2SEA[IJ]) OLDFL[IJ] = 0.0;
        else OLDFL[IJ] = M <= NFRE ? FLR(IJ, K, M) : FLR(IJ, K, NFRE) * FR5OFREQ5;
      }
      for (int IJ = 1; IJ <= KIJL; ++IJ) {
        const double FNEW = FNEF[IJ];
        const int NEWM = NEWF[IJ];
        if (NEWM < NFRE && NEWM >= 1) {
          const int NEWM1 = NEWM + 1;
          const double GWH = DFREQTH / (t.FR(NEWM1) - t.FR(NEWM)) * OLDFL[IJ];
          GWM[IJ] = GWH * (t.FR(NEWM1) - FNEW) / DFTH[NEWM];
          GWP[IJ] = GWH * (FNEW - t.FR(NEWM)) / DFTH[NEWM1];
          NEWFLA[IJ] = NEWM1;
        } else if (NEWM == 0) {
          const double GWH = t.FRATIO * DFREQTH / (FRE0 * t.FR(1)) * OLDFL[IJ];
          GWP[IJ] = GWH * (FNEW - t.FR(1) / t.FRATIO) / DFTH[1];
          NEWF[IJ] = -1; NEWFLA[IJ] = 1;
        } else if (NEWM == NFRE) {
          const double GWH = DFREQTH / (FRE0 * t.FR(NFRE)) * OLDFL[IJ];
          GWM[IJ] = GWH * (t.FRATIO * t.FR(NFRE) - FNEW) / DFTH[NFRE];
          NEWFLA[IJ] = -1;
        } else { NEWF[IJ] = -1; NEWFLA[IJ] = -1; }
      }
      for (int IJ = 1; IJ <= KIJL; ++IJ) {
        const int NEWM = NEWF[IJ], NEWM1 = NEWFLA[IJ], KH = KNEW[IJ];
        if (NEWM != -1) FLA(IJ, KH, NEWM) = FLA(IJ, KH, NEWM) + GWM[IJ];
        if (NEWM1 != -1) FLA(IJ, KH, NEWM1) = FLA(IJ, KH, NEWM1) + GWP[IJ];
      }
    }
  }
  for (int M = 1; M <= NFRE; ++M) for (int K = 1; K <= NANG; ++K) for (int IJ = 1; IJ <= KIJL; ++IJ) FLA(IJ, K, M) = std::max(FLA(IJ, K, M), t.EPSMIN);
}

// outblock.F90:150-610 for one chunk; BOUT (KIJL, NIPRMOUT).  sel.itg[c] = reference parameter number of column c+1.
void outblock(const Config& c, const Tables& t, Fields& f, int KIJL, int ICHNK, const OutSel& sel, double* BOUT) {
  const int NANG = c.nang, NFRE = c.nfre, NTRAIN = 3, NTEWH = 6;   // yowcout.F90:19, mpcrtbl.F90:373-399
  const S3 FL1{&f.FL1(1, 1, 1, ICHNK), KIJL, NANG}, XLLWS{&f.XLLWS(1, 1, 1, ICHNK), KIJL, NANG};
  auto p1 = [&](ArrD& a) { return &a(1, ICHNK); };
  const double *CINV = &f.CINV(1, 1, ICHNK), *CGROUP = &f.CGROUP(1, 1, ICHNK);
  const double *UFRIC = p1(f.UFRIC), *WSWAVE = p1(f.WSWAVE), *WDWAVE = p1(f.WDWAVE), *CICOVER = p1(f.CICOVER);
  auto col = [&](int itg) -> double* {   // ITOBOUT
    for (int i = 0; i < sel.n; ++i) if (sel.itg[i] == itg) return BOUT + (size_t)i * KIJL - 1;   // 1-based IJ
    return nullptr;
  };
  for (size_t i = 0; i < (size_t)KIJL * sel.n; ++i) BOUT[i] = 0.0;
  // output spectrum (outblock.F90:168-194): INTPOL with currents, LSECONDORDER = F; noise-level restructuring under sea ice
  W3 FL2ND(KIJL, NANG, NFRE);
  if (c.irefra == 2 || c.irefra == 3) intpol(t, KIJL, NANG, NFRE, FL1, FL2ND, &f.WAVNUM(1, 1, ICHNK), p1(f.UCUR), p1(f.VCUR), 1);   // outblock.F90:168-169
  else for (int M = 1; M <= NFRE; ++M) for (int K = 1; K <= NANG; ++K) for (int IJ = 1; IJ <= KIJL; ++IJ) FL2ND(IJ, K, M) = FL1(IJ, K, M);
  if (c.licerun && !c.lmaskice) {
    V ZTHRS(KIJL + 1), ZRDUC(KIJL + 1);
    for (int IJ = 1; IJ <= KIJL; ++IJ) ZTHRS[IJ] = (1.0 - 0.9 * std::min(CICOVER[IJ - 1], 0.99)) * c.flmin;
    for (int M = 1; M <= NFRE; ++M) {
      for (int IJ = 1; IJ <= KIJL; ++IJ) ZRDUC[IJ] = std::exp(-10.0 * (t.FR(M) * t.FR(M)) / std::sqrt(std::max(WSWAVE[IJ - 1], 1.0)));
      for (int K = 1; K <= NANG; ++K)
        for (int IJ = 1; IJ <= KIJL; ++IJ)
          if (FL2ND(IJ, K, M) <= ZTHRS[IJ]) FL2ND(IJ, K, M) = std::max(ZRDUC[IJ] * FL2ND(IJ, K, M), ZTHRS[IJ] * (ZRDUC[IJ] * ZRDUC[IJ]));
    }
  }
  V COSWDIF((size_t)KIJL * NANG + 1);
  for (int K = 1; K <= NANG; ++K) for (int IJ = 1; IJ <= KIJL; ++IJ) COSWDIF[(size_t)(K - 1) * KIJL + IJ] = std::cos(t.TH(K) - WDWAVE[IJ - 1]);
  V EM(KIJL + 1), FM(KIJL + 1), DP(KIJL + 1), TMP(KIJL + 1), TMP2(KIJL + 1);
  femean_out(t, KIJL, NANG, NFRE, FL2ND.view(), EM, FM);
  dominant_period(t, KIJL, NANG, NFRE, FL2ND.view(), DP);
  SepOut so;
  sepwisw(t, KIJL, NANG, NFRE, FL1, XLLWS, CINV, UFRIC, WDWAVE, COSWDIF, so);
  auto todeg = [&](double th) { return std::fmod(t.DEG * th + 180.0, 360.0); };
  auto invf = [&](double fq) { return fq > 0.0 ? 1.0 / fq : sel.zmiss; };
  double* b;
  if ((b = col(1))) for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = 4.0 * std::sqrt(std::max(EM[IJ], 0.0));
  if ((b = col(2))) { sthq(t, KIJL, NANG, NFRE, FL2ND.view(), TMP); for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = todeg(TMP[IJ]); }
  if ((b = col(3))) for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = invf(FM[IJ]);
  if ((b = col(4))) for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = UFRIC[IJ - 1];
  if ((b = col(5))) for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = todeg(WDWAVE[IJ - 1]);
  if ((b = col(6))) for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = DP[IJ] > 0.0 ? DP[IJ] : sel.zmiss;
  if ((b = col(7))) {   // outbeta.F90:66-91, 113-117
    const double AMAX = 0.02, BMAX = 0.01;
    const double *Z0M = p1(f.Z0M), *Z0B = p1(f.Z0B), *CHRNCK = p1(f.CHRNCK);
    (void)Z0M; (void)Z0B;
    for (int IJ = 1; IJ <= KIJL; ++IJ) {
      const double ALPHAMAXU10 = c.llgcbz0 ? t.ALPHAMAX : std::min(t.ALPHAMAX, AMAX + BMAX * WSWAVE[IJ - 1]);
      const double USM = 1.0 / std::max(UFRIC[IJ - 1], t.EPSUS);
      const double BETAM = std::max(std::min(CHRNCK[IJ - 1], ALPHAMAXU10), t.ALPHAMIN);
      const double Z0ATM = c.rnum * USM + t.GM1 * BETAM * (UFRIC[IJ - 1] * UFRIC[IJ - 1]);
      const double q = t.XKAPPA / std::log(1.0 + t.XNLEV / Z0ATM);
      b[IJ] = std::min(q * q, 0.01);
    }
  }
  if ((b = col(8))) { const double* TAUW = p1(f.TAUW); for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = TAUW[IJ - 1] / std::max(UFRIC[IJ - 1] * UFRIC[IJ - 1], t.EPSUS); }
  if ((b = col(9))) meansqs(c, t, t.XK_GC(t.NWAV_GC), KIJL, NANG, NFRE, FL1, &f.WAVNUM(1, 1, ICHNK), UFRIC, COSWDIF, b);   // outblock.F90:285-287, userin.F90:1213-1215
  if (col(29) || col(30) || col(31) || col(33) || col(34) || col(57) || col(70) || col(71) || col(72)) {   // outblock.F90:208-211, 385-404, 481, 526-535
    KurtOut ko;
    kurtosis(c, t, KIJL, NANG, NFRE, FL1, p1(f.DEPTH), ko);
    const std::pair<int, V*> cols[] = {{29, &ko.C4}, {30, &ko.BF2}, {31, &ko.QP}, {33, &ko.HMAX}, {34, &ko.TMAX}, {57, &ko.C3}, {70, &ko.ETA_M},
                                       {71, &ko.R}, {72, &ko.XNSLC}};
    for (auto& pr : cols) if ((b = col(pr.first))) for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = (*pr.second)[IJ];
  }
  if ((b = col(10))) for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = WSWAVE[IJ - 1];
  if ((b = col(11))) for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = 4.0 * std::sqrt(std::max(so.ESEA[IJ], 0.0));
  if ((b = col(12))) for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = 4.0 * std::sqrt(std::max(so.ESWELL[IJ], 0.0));
  if ((b = col(13))) for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = todeg(so.THWISEA[IJ]);
  if ((b = col(14))) for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = todeg(so.THSWELL[IJ]);
  if ((b = col(15))) for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = invf(so.FSEA[IJ]);
  if ((b = col(16))) for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = invf(so.FSWELL[IJ]);
  if ((b = col(20))) { mwp12(t, KIJL, NANG, 1, FL2ND.view(), TMP); for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = TMP[IJ]; }
  if ((b = col(21))) { mwp12(t, KIJL, NANG, 2, FL2ND.view(), TMP); for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = TMP[IJ]; }
  if ((b = col(22))) { wdirspread(t, KIJL, NANG, NFRE, FL2ND.view(), EM, false, TMP); for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = TMP[IJ]; }
  struct { int itg; V* v; } sw[] = {{23, &so.P1SEA}, {24, &so.P1SWELL}, {25, &so.P2SEA}, {26, &so.P2SWELL}, {27, &so.SPRDSEA}, {28, &so.SPRDSWELL}};
  for (auto& e : sw) if ((b = col(e.itg))) for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = (*e.v)[IJ];
  struct { int itg; ArrD* a; } cp[] = {{32, &f.DEPTH}, {35, &f.USTOKES}, {36, &f.VSTOKES}, {37, &f.UCUR}, {38, &f.VCUR}, {39, &f.PHIEPS},
                                       {40, &f.PHIAW}, {41, &f.TAUOC}, {44 + 3 * NTRAIN, &f.AIRD}, {45 + 3 * NTRAIN, &f.WSTAR},
                                       {46 + 3 * NTRAIN, &f.CICOVER}, {47 + 3 * NTRAIN, &f.CITHICK},
                                       {58 + 3 * NTRAIN + NTEWH, &f.TAUXD}, {59 + 3 * NTRAIN + NTEWH, &f.TAUYD},
                                       {60 + 3 * NTRAIN + NTEWH, &f.TAUOCXD}, {61 + 3 * NTRAIN + NTEWH, &f.TAUOCYD}};
  for (auto& e : cp) if ((b = col(e.itg))) { const double* s = p1(*e.a); for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = s[IJ - 1]; }
  if (col(53 + 3 * NTRAIN) || col(54 + 3 * NTRAIN)) {
    weflux(t, KIJL, NANG, NFRE, FL1, CGROUP, TMP, TMP2);
    if ((b = col(53 + 3 * NTRAIN))) for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = TMP[IJ];
    if ((b = col(54 + 3 * NTRAIN))) for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = todeg(TMP2[IJ]);
  }
  // SE10MEAN (se10mean.F90:60-66) and the period bands of mpcrtbl.F90:373-399 (IPRMINFO(:,4:5)), outblock.F90:443-446, 525-533
  static const double BANDS[6][2] = {{10, 12}, {12, 14}, {14, 17}, {17, 21}, {21, 25}, {25, 30}};
  if ((b = col(43 + 3 * NTRAIN))) {
    sebtmean(t, KIJL, NANG, NFRE, FL2ND.view(), 10.0, 1.0 / t.FR(1), TMP);
    for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = 4.0 * std::sqrt(std::max(TMP[IJ], 0.0));
  }
  for (int IH = 1; IH <= NTEWH; ++IH)
    if ((b = col(54 + 3 * NTRAIN + IH))) {
      sebtmean(t, KIJL, NANG, NFRE, FL2ND.view(), BANDS[IH - 1][0], BANDS[IH - 1][1], TMP);
      for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = 4.0 * std::sqrt(std::max(TMP[IJ], 0.0));
    }
  if ((b = col(62 + 3 * NTRAIN + NTEWH))) { const double* s = p1(f.PHIOCD); for (int IJ = 1; IJ <= KIJL; ++IJ) b[IJ] = std::max(-s[IJ - 1], 0.0); }
  // outsetwmask.F90:62-78
  for (int i = 0; i < sel.n; ++i) {
    double* bc = BOUT + (size_t)i * KIJL - 1;
    if (c.licerun && sel.llsource && sel.icemask[i] == 1)
      for (int IJ = 1; IJ <= KIJL; ++IJ) if (CICOVER[IJ - 1] > c.cithrsh) bc[IJ] = sel.zmiss;
    if (sel.seamask[i] == 1)
      for (int IJ = 1; IJ <= KIJL; ++IJ) { const int io = f.IODP(IJ, ICHNK); bc[IJ] = bc[IJ] * io + (1 - io) * sel.zmiss; }
  }
}

// mpminmaxavg.F90:68-195 over all emulated ranks.  global = LLNORMWAMOUT_GLOBAL: sums in the ORIGINAL global point order
// on one rank (reproducible for any NPROC); otherwise per-rank partial sums combined in rank order (MPL_ALLREDUCE).
void mpminmaxavg(Model& m, const OutSel& sel, const std::vector<std::vector<double>>& bout, bool global, double* WNORM /*(4,NIPRMOUT)*/) {
  const int NI = sel.n;
  const double HUGE_ = std::numeric_limits<double>::max();
  if (global) {
    std::vector<double> ZGLOBAL(m.grid.NIBLO + 1);
    for (int it = 0; it < NI; ++it) {
      for (int ir = 0; ir < m.cfg.npr; ++ir) {
        const RankDecomp& r = m.ranks[ir];
        for (int ICHNK = 1; ICHNK <= r.NCHNK; ++ICHNK)
          for (int IJ = 1; IJ <= r.KIJL4CHNK(ICHNK); ++IJ)
            ZGLOBAL[r.IJFROMCHNK(IJ, ICHNK)] = bout[ir][(IJ - 1) + (size_t)r.NPROMA * (it + (size_t)NI * (ICHNK - 1))];
      }
      double ZSUM = 0.0, ZMIN = HUGE_, ZMAX = -HUGE_;
      int ICOUNT = 0;
      for (int IJOLD = 1; IJOLD <= m.grid.NIBLO; ++IJOLD) {
        const int IJ = (m.cfg.ll1d || m.cfg.npr == 1) ? IJOLD : m.grid.IJ2NEWIJ(IJOLD);
        if (ZGLOBAL[IJ] != sel.zmiss) { ICOUNT = ICOUNT + 1; ZSUM = ZSUM + ZGLOBAL[IJ]; ZMIN = std::min(ZMIN, ZGLOBAL[IJ]); ZMAX = std::max(ZMAX, ZGLOBAL[IJ]); }
      }
      WNORM[4 * it + 0] = ZSUM / std::max(ICOUNT, 1); WNORM[4 * it + 1] = ZMIN; WNORM[4 * it + 2] = ZMAX; WNORM[4 * it + 3] = ICOUNT;
    }
    return;
  }
  for (int it = 0; it < NI; ++it) {
    double ZSUMT = 0.0, ZCNT = 0.0, ZMIN = HUGE_, ZMAX = -HUGE_;
    for (int ir = 0; ir < m.cfg.npr; ++ir) {
      const RankDecomp& r = m.ranks[ir];
      double ZSUM = 0.0, ZC = 0.0;
      for (int ICHNK = 1; ICHNK <= r.NCHNK; ++ICHNK)
        for (int IPRM = 1; IPRM <= r.KIJL4CHNK(ICHNK); ++IPRM) {
          const double v = bout[ir][(IPRM - 1) + (size_t)r.NPROMA * (it + (size_t)NI * (ICHNK - 1))];
          if (v != sel.zmiss) { ZSUM = ZSUM + v; ZC = ZC + 1.0; ZMIN = std::min(ZMIN, v); ZMAX = std::max(ZMAX, v); }
        }
      ZSUMT += ZSUM; ZCNT += ZC;
    }
    WNORM[4 * it + 1] = ZMIN; WNORM[4 * it + 2] = ZMAX; WNORM[4 * it + 3] = ZCNT;
    WNORM[4 * it + 0] = ZCNT < 1.0 ? -HUGE_ : ZSUMT / ZCNT;
  }
}

// getwnd.F90:196-212: WAMWND (wamwnd.F90:120-300, ICODE_WND = 3) followed by MICEP (micep.F90:84-240, LWCOU = F, no NEMO
// fields, CLDOMAIN /= 's') for N points.  FIELDG arrays (NX, NY) first index fastest, IFROMIJ / JFROMIJ relative to NXS / NYS.
void getwnd_points(long N, const int* IFROMIJ, const int* JFROMIJ, int NXS, int NYS, int NX, const double* UWND, const double* VWND,
                   const double* AIRD, const double* WSTARG, const double* CICOVERG, const double* CITHICKG, const double* USTRAG,
                   const double* VSTRAG, const double* WSWAVEG, const double* WDWAVEG, const double* UCUR, const double* VCUR,
                   int LLWSWAVE, int LLWDWAVE, int LCORREL, int IPARAM, int LICETH, int LICERUN, int LMASKICE, double WSPMIN, double ZMISS,
                   double ZPI, double* U10, double* THW, double* ADS, double* WSTAR, double* CICVR, double* CITH, double* USTRA, double* VSTRA) {
  const double RWFAC = 0.5, C1 = 0.2, C2 = 0.4, HICMIN = 0.2;   // yowwind.F90:21, micep.F90:81-82, yowice.F90:23
  auto G = [&](const double* a, long IJ) { return a[(size_t)(IFROMIJ[IJ] - NXS) + (size_t)NX * (size_t)(JFROMIJ[IJ] - NYS)]; };
  std::vector<double> UU(N), VV(N), WSPEED(N);
  for (long IJ = 0; IJ < N; ++IJ) {
    UU[IJ] = G(UWND, IJ); VV[IJ] = G(VWND, IJ); ADS[IJ] = G(AIRD, IJ); WSTAR[IJ] = G(WSTARG, IJ); CITH[IJ] = G(CITHICKG, IJ);
    USTRA[IJ] = G(USTRAG, IJ); VSTRA[IJ] = G(VSTRAG, IJ);
  }
  if (LLWSWAVE && LLWDWAVE) {
    for (long IJ = 0; IJ < N; ++IJ) { U10[IJ] = G(WSWAVEG, IJ); THW[IJ] = G(WDWAVEG, IJ); }
    for (long IJ = 0; IJ < N; ++IJ)
      if (U10[IJ] <= 0.0) {
        WSPEED[IJ] = std::sqrt(UU[IJ] * UU[IJ] + VV[IJ] * VV[IJ]);
        if (WSPEED[IJ] > 0.0) { U10[IJ] = WSPEED[IJ]; THW[IJ] = std::atan2(UU[IJ], VV[IJ]); } else { U10[IJ] = 0.0; THW[IJ] = 0.0; }
      }
    if (LCORREL)
      for (long IJ = 0; IJ < N; ++IJ) {
        UU[IJ] = U10[IJ] * std::sin(THW[IJ]); VV[IJ] = U10[IJ] * std::cos(THW[IJ]);
        UU[IJ] = UU[IJ] - RWFAC * UCUR[IJ]; VV[IJ] = VV[IJ] - RWFAC * VCUR[IJ];
        WSPEED[IJ] = std::sqrt(UU[IJ] * UU[IJ] + VV[IJ] * VV[IJ]);
        if (WSPEED[IJ] > 0.0) { U10[IJ] = WSPEED[IJ]; THW[IJ] = std::atan2(UU[IJ], VV[IJ]); } else { U10[IJ] = 0.0; THW[IJ] = 0.0; }
      }
  } else {
    if (LLWSWAVE)
      for (long IJ = 0; IJ < N; ++IJ) {
        const double WS = G(WSWAVEG, IJ);
        if (WS != ZMISS && WS > 0.0) {
          WSPEED[IJ] = std::sqrt(UU[IJ] * UU[IJ] + VV[IJ] * VV[IJ]);
          if (WSPEED[IJ] > 0.0) { const double RESCALE = WS / WSPEED[IJ]; UU[IJ] = UU[IJ] * RESCALE; VV[IJ] = VV[IJ] * RESCALE; }
        }
      }
    if (LCORREL) for (long IJ = 0; IJ < N; ++IJ) { UU[IJ] = UU[IJ] + RWFAC * UCUR[IJ]; VV[IJ] = VV[IJ] + RWFAC * VCUR[IJ]; }
    for (long IJ = 0; IJ < N; ++IJ) {
      U10[IJ] = std::sqrt(UU[IJ] * UU[IJ] + VV[IJ] * VV[IJ]);
      THW[IJ] = U10[IJ] != 0.0 ? std::atan2(UU[IJ], VV[IJ]) : 0.0;
    }
  }
  for (long IJ = 0; IJ < N; ++IJ) U10[IJ] = std::max(U10[IJ], WSPMIN);
  for (long IJ = 0; IJ < N; ++IJ) if (THW[IJ] < 0.0) THW[IJ] = THW[IJ] + ZPI;
  // MICEP
  if (IPARAM == 31) {
    for (long IJ = 0; IJ < N; ++IJ) {
      const double CI = G(CICOVERG, IJ);
      if (CI == ZMISS || CI < 0.01 || CI > 1.01) CICVR[IJ] = 0.0;
      else if (CI > 0.95) CICVR[IJ] = 1.0;
      else CICVR[IJ] = CI;
    }
  } else if (IPARAM == 139) {
    for (long IJ = 0; IJ < N; ++IJ) CICVR[IJ] = G(CICOVERG, IJ) < 271.5 ? 1.0 : 0.0;
  }
  if (!LICERUN || LMASKICE) {
    for (long IJ = 0; IJ < N; ++IJ) CITH[IJ] = 0.0;
  } else if (!LICETH) {
    for (long IJ = 0; IJ < N; ++IJ) CITH[IJ] = CICVR[IJ] > 0.0 ? std::max(C1 + C2 * CICVR[IJ], 0.0) : 0.0;
  } else {
    for (long IJ = 0; IJ < N; ++IJ) CITH[IJ] = CICVR[IJ] * CITH[IJ];
    for (long IJ = 0; IJ < N; ++IJ) if (CICVR[IJ] > 0.0 && CITH[IJ] < 0.5 * HICMIN) { CICVR[IJ] = 0.0; CITH[IJ] = 0.0; }
  }
}

}  // namespace orc

extern "C" void orc_getwnd_points(long N, const int* IFROMIJ, const int* JFROMIJ, int NXS, int NYS, int NX, const double* const* FIELDG /*[10]*/,
                                  const double* UCUR, const double* VCUR, const int* OPT /*[7]*/, const double* ROPT /*[3]*/, double* const* OUT /*[8]*/) {
  orc::getwnd_points(N, IFROMIJ, JFROMIJ, NXS, NYS, NX, FIELDG[0], FIELDG[1], FIELDG[2], FIELDG[3], FIELDG[4], FIELDG[5], FIELDG[6], FIELDG[7],
                     FIELDG[8], FIELDG[9], UCUR, VCUR, OPT[0], OPT[1], OPT[2], OPT[3], OPT[4], OPT[5], OPT[6], ROPT[0], ROPT[1], ROPT[2],
                     OUT[0], OUT[1], OUT[2], OUT[3], OUT[4], OUT[5], OUT[6], OUT[7]);
}
