// The steps either side of the WAMINTGR hot path, kept on the device:
//   k_newwind   NEWWIND (newwind.F90:105-167, ICODE_WND = 3)
//   k_outblock  OUTBS/OUTBLOCK core integrated parameters (outbs.F90:97-122, outblock.F90:150-610): FEMEAN, STHQ,
//               DOMINANT_PERIOD, SEPWISW (LLPARTITION = F), MWP1, MWP2, WDIRSPREAD/PEAKFRI/SCOSFL, OUTBETA, WEFLUX,
//               pass-through fields, OUTSETWMASK
//   k_norm_*    OUTWNORM/MPMINMAXAVG statistics (mpminmaxavg.F90:68-195), both the per-rank-partial and the
//               reproducible global-order flavour
// Lane = grid point (32 consecutive points of one (k,m) bin are one 256-byte row of the NPROMA-chunked arrays), the
// per-point accumulators live in registers, cos(TH(k)-WDWAVE) in a thread-private shared-memory column.
#include "internal.h"

namespace ew {

__constant__ OutConst c_oc;

int upload_out_const(const OutConst& h, cudaStream_t st) {
  EW_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_oc, &h, sizeof(OutConst), 0, cudaMemcpyHostToDevice, st));
  return 0;
}

namespace {
__device__ __forceinline__ double omax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double omin(double a, double b) { return a < b ? a : b; }

// ---------------------------------------------------------------------------------------------------------
__global__ void k_newwind(long long n, ecwam_b200_fields f, ecwam_b200_forcing_next nx, double acd, double bcd, double epsmin) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double WSPMIN_RESET_TAUW = 4.0;   // yowwind.F90:19
  const double wght = 1.0 / omax(WSPMIN_RESET_TAUW, epsmin);
  const double w = nx.wswave[p];
  f.wswave[p] = w;
  if (w < WSPMIN_RESET_TAUW) {            // low-wind cap on the first-guess wave stress (newwind.F90:133-139)
    const double tlwmax = wght * (acd + bcd * w) * (w * w * w);
    f.tauw[p] = omin(f.tauw[p], tlwmax);
  }
  f.wdwave[p] = nx.wdwave[p]; f.aird[p] = nx.aird[p]; f.wstar[p] = nx.wstar[p]; f.cicover[p] = nx.cicover[p];
  f.cithick[p] = nx.cithick[p]; f.ustra[p] = nx.ustra[p]; f.vstra[p] = nx.vstra[p];
}

// NEWWIND with the friction velocity as forcing (newwind.F90:141-161, ICODE_WND = 1, 2)
__global__ void k_newwind_ustar(long long n, ecwam_b200_fields f, ecwam_b200_forcing_next nx, const double* __restrict__ ufric_next, double alpha) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double USTMIN_RESET_TAUW = 0.08;   // yowwind.F90:20
  const double us = ufric_next[p];
  f.ufric[p] = us;
  const double q = alpha / f.chrnck[p];
  f.tauw[p] = us < USTMIN_RESET_TAUW ? 0.0 : us * us * (1.0 - q * q);
  f.wdwave[p] = nx.wdwave[p]; f.aird[p] = nx.aird[p]; f.wstar[p] = nx.wstar[p]; f.cicover[p] = nx.cicover[p];
  f.cithick[p] = nx.cithick[p]; f.ustra[p] = nx.ustra[p]; f.vstra[p] = nx.vstra[p];
}

// GETWND's blocking step: WAMWND (wamwnd.F90:120-300, ICODE_WND = 3) + MICEP (micep.F90:84-240, uncoupled) per grid point
__global__ void k_getwnd(GetwndArgs a, ecwam_b200_fieldg g, ecwam_b200_getwnd_opts o, const int* __restrict__ ifromij,
                         const int* __restrict__ jfromij, ecwam_b200_forcing_next nx) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.npts) return;
  const size_t q = (size_t)(ifromij[p] - o.nxs) + (size_t)a.nx * (size_t)(jfromij[p] - o.nys);
  double uu = g.uwnd[q], vv = g.vwnd[q];
  double u10, thw;
  auto polar = [&](bool strict) {           // speed and direction of (UU, VV); :163-171 tests > 0, :214-219 tests /= 0
    u10 = sqrt(uu * uu + vv * vv);
    if (strict ? (u10 > 0.0) : (u10 != 0.0)) thw = atan2(uu, vv);
    else { u10 = strict ? 0.0 : u10; thw = 0.0; }
  };
  const double RWFAC = 0.5;                 // yowwind.F90:21
  if (o.llwswave && o.llwdwave) {           // wamwnd.F90:152-190
    u10 = g.wswave[q]; thw = g.wdwave[q];
    if (u10 <= 0.0) polar(true);
    if (a.lcorrel) {
      uu = u10 * sin(thw); vv = u10 * cos(thw);
      uu = uu - RWFAC * a.ucur[p]; vv = vv - RWFAC * a.vcur[p];
      polar(true);
    }
  } else {
    if (o.llwswave) {                       // :192-205 rescale the components to the wave model's wind speed
      const double ws = g.wswave[q];
      if (ws != o.zmiss && ws > 0.0) {
        const double wspeed = sqrt(uu * uu + vv * vv);
        if (wspeed > 0.0) { const double rescale = ws / wspeed; uu = uu * rescale; vv = vv * rescale; }
      }
    }
    if (a.lcorrel) { uu = uu + RWFAC * a.ucur[p]; vv = vv + RWFAC * a.vcur[p]; }   // :206-211, :223-228
    polar(false);
  }
  u10 = omax(u10, a.wspmin);                // :240-242
  if (thw < 0.0) thw = thw + a.zpi;         // :290-292
  // MICEP
  const double ci = g.cicover[q];
  double cicvr = 0.0;
  if (o.iparamci == 31) {                   // micep.F90:113-125
    if (ci == o.zmiss || ci < 0.01 || ci > 1.01) cicvr = 0.0;
    else if (ci > 0.95) cicvr = 1.0;
    else cicvr = ci;
  } else cicvr = ci < 271.5 ? 1.0 : 0.0;    // 139: sea-surface temperature (:126-135)
  double cith = g.cithick[q];               // WAMWND's CITH (:136)
  if (!a.licerun || a.lmaskice) cith = 0.0;                                             // micep.F90:139-142
  else if (!o.liceth) cith = cicvr > 0.0 ? omax(0.2 + 0.4 * cicvr, 0.0) : 0.0;         // :181-188 (C1 = 0.2, C2 = 0.4)
  else {
    cith = cicvr * cith;                                                               // :231-233
    if (cicvr > 0.0 && cith < 0.5 * 0.2) { cicvr = 0.0; cith = 0.0; }                  // HICMIN = 0.2 (:234-239)
  }
  const_cast<double*>(nx.wswave)[p] = u10; const_cast<double*>(nx.wdwave)[p] = thw;
  const_cast<double*>(nx.aird)[p] = g.aird[q]; const_cast<double*>(nx.wstar)[p] = g.wstar[q];
  const_cast<double*>(nx.cicover)[p] = cicvr; const_cast<double*>(nx.cithick)[p] = cith;
  const_cast<double*>(nx.ustra)[p] = g.ustra[q]; const_cast<double*>(nx.vstra)[p] = g.vstra[q];
}

// WAMINTGR without a source-term update (wamintgr.F90:163-171 when LLSOURCE = F, :188-195 when it is not yet time to
// integrate): MIJ = NFRE, XLLWS = 0 and, for LLSOURCE = F, FL1 = MAX(FL1, EPSMIN)
__global__ void k_no_source(long long n4, long long n2, double* __restrict__ fl1, double* __restrict__ xllws, int* __restrict__ mij,
                            int nfre, int clip, double epsmin) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n2) mij[i] = nfre;
  if (i >= n4) return;
  xllws[i] = 0.0;
  if (clip) fl1[i] = omax(fl1[i], epsmin);
}

// ---------------------------------------------------------------------------------------------------------
// per-spectrum accumulators of the frequency sweep (FEMEAN, STHQ, MWP1/MWP2, PEAKFRI + SCOSFL)
struct SpecAcc {
  double e, fo, t2last;          // femean.F90:95-119
  double si, ci;                 // sthq.F90:78-96
  double es, p1, p2, t0odd;      // mwp1.F90:84-108, mwp2.F90
  double epk, spk, cpk;          // peakfri.F90:66-88 (peak of the 1-D spectrum) and the SI/CI of that row (scosfl.F90:81-86)
  __device__ void zero() { e = fo = t2last = si = ci = es = p1 = p2 = t0odd = epk = spk = cpk = 0.0; }
  // rows are visited from NFRE down to 1: ">=" keeps the LOWEST row among equal maxima, as the ascending "<" of PEAKFRI
  __device__ __forceinline__ void row(int m, double t2, double t0, double s, double c, double f1d) {
    const int F = c_oc.F;
    e += t2 * c_oc.DFIM[m]; fo += c_oc.DFIMOFR[m] * t2;
    if (m == F - 1) t2last = t2;
    si += c_oc.DFIM[m] * s; ci += c_oc.DFIM[m] * c;
    if (m < c_oc.NFRE_ODD) {
      const double ds = c_oc.DFIM_SIM[m], fr = c_oc.FR[m];
      es += ds * t0; p1 += (ds * fr) * t0; p2 += (ds * (fr * fr)) * t0;
      if (m == c_oc.NFRE_ODD - 1) t0odd = t0;
    }
    if (f1d >= epk && f1d > 0.0) { epk = f1d; spk = s; cpk = c; }
  }
  __device__ void femean(double& em, double& fm) const {
    const int F = c_oc.F;
    em = e + (c_oc.WETAIL * c_oc.FR[F - 1] * c_oc.DELTH) * t2last;
    fm = fo + (c_oc.FRTAIL * c_oc.DELTH) * t2last;
    fm = omax(em / fm, c_oc.FR[0]);
  }
  __device__ double thq() const {
    const double c = ci == 0.0 ? c_oc.EPSMIN : ci;
    double th = atan2(si, c);
    if (th < 0.0) th += c_oc.ZPI;
    return th;
  }
  __device__ double mwp(int ip) const {
    const double fro = c_oc.FR[c_oc.NFRE_ODD - 1];
    const double em = es + (c_oc.WETAIL * fro * c_oc.DELTH) * t0odd;
    double mw = ip == 1 ? p1 + (c_oc.WP1TAIL * c_oc.DELTH * (fro * fro)) * t0odd : p2 + (c_oc.WP2TAIL * c_oc.DELTH * (fro * fro * fro)) * t0odd;
    if (em > 0.0 && mw > c_oc.EPSMIN) {
      mw = ip == 1 ? em / mw : sqrt(em / mw);
      return omin(mw, 1.0 / c_oc.FR[0]);
    }
    return 0.0;
  }
  // wdirspread.F90:104-118,135 with LLPEAKF: DELTH*sum_k cos(TH(k)-MEANDIR) F = DELTH*sqrt(SI^2+CI^2)
  __device__ double spread_peak() const {
    double x = 1.0;
    if (epk > 0.0) x = omin(c_oc.DELTH * sqrt(spk * spk + cpk * cpk) / epk, 1.0);
    return sqrt(2.0 * (1.0 - x));
  }
};

#define OB_NTH 128
#ifndef OB_MINB
#define OB_MINB 4   // CTAs per SM (register cap): the kernel is bound by the latency of its row loads
#endif
#ifndef OB_UNR
#define OB_UNR 4    // directions in flight per thread
#endif
#define OB_STR2(x) #x
#define OB_STR(x) OB_STR2(x)
#define OB_UNROLL _Pragma(OB_STR(unroll OB_UNR))

// CUR: IREFRA = 2, 3 -- the output spectrum FL2ND is INTPOL's (k_intpol, d.fl2) instead of FL1 itself (outblock.F90:168-172)
template <bool CUR>
__global__ void __launch_bounds__(OB_NTH, OB_MINB) k_outblock(OutDev d) {
  extern __shared__ double cwd_s[];   // [A][OB_NTH]: COSWDIF(IJ,K) = cos(TH(K) - WDWAVE(IJ)) (outblock.F90:198-202)
  const int A = c_oc.A, F = c_oc.F, P = d.P;
  const int tid = threadIdx.x;
  const long long p = (long long)blockIdx.x * OB_NTH + tid;
  if (p >= d.npts) return;
  const long long pc = p / P;
  const int pi = (int)(p - pc * P);
  const size_t rs = (size_t)P * A;    // frequency stride of a (P,A,F,C) array
  const double* fl = d.f.fl1 + (size_t)pi + rs * F * (size_t)pc;
  const double* xl = d.f.xllws + (size_t)pi + rs * F * (size_t)pc;
  const double* fl2 = CUR ? d.fl2 + (size_t)pi + rs * F * (size_t)pc : fl;
  const size_t b3 = (size_t)pi + (size_t)P * F * (size_t)pc;
  const double* cinv = d.f.cinv + b3;
  const double* cgr = d.f.cgroup + b3;
  const double wd = d.f.wdwave[p], ufric = d.f.ufric[p], wsw = d.f.wswave[p], cic = d.f.cicover[p];
  double* cw = cwd_s + tid;
  for (int k = 0; k < A; ++k) cw[k * OB_NTH] = cos(c_oc.TH[k] - wd);
  const double COEF = 1.2 * 28.0;     // OLDWSFC*FRIC (yowfred.F90:81-82, sepwisw.F90:128)
  const double EPSMIN = c_oc.EPSMIN, DELTH = c_oc.DELTH;
  // noise-level restructuring of the output spectrum under sea ice (outblock.F90:176-194)
  const bool icen = c_oc.licerun && !c_oc.lmaskice;
  const double zthrs = (1.0 - 0.9 * omin(cic, 0.99)) * c_oc.flmin;
  const double wsq = sqrt(omax(wsw, 1.0));

  // ---------------- pass 1: first-guess swell mask -> FSWELL, FSEA -> R (sepwisw.F90:138-196); maximum for FCROP;
  //                  with parameter 9 also the moments of the spectrum in the wind direction (HALPHAP, MEANSQS_LF)
  double fmax = 0.0, Rf, xmss9 = 0.0;
  {
    double es = 0, fs = 0, ew = 0, fw = 0, t2s = 0, t2w = 0;
    double h_xm = 0, h_em = 0, h_fm = 0, h_t2 = 0, h_f1d = 0, lf = 0;    // HALPHAP's XMSS / EM / FM / F1D, MEANSQS_LF of FL1
    const bool mss = c_oc.want_mss != 0;
    for (int m = 0; m < F; ++m) {
      const double xinv = ufric * __ldg(cinv + (size_t)m * P);
      const double zr = icen ? exp(-10.0 * (c_oc.FR[m] * c_oc.FR[m]) / wsq) : 1.0;
      t2s = 0.0; t2w = 0.0;
      double h_s = 0.0, r_s = 0.0;
      h_t2 = 0.0;
      OB_UNROLL
      for (int k = 0; k < A; ++k) {
        const size_t o = (size_t)m * rs + (size_t)k * P;
        const double f = __ldg(fl + o), x = __ldg(xl + o), c = cw[k * OB_NTH];
        const double swm = (x != 0.0) ? 0.0 : ((xinv * (COEF * c) >= 1.0) ? 0.0 : 1.0);
        const double f1 = f * swm;
        t2s += omax(f1, EPSMIN);
        t2w += omax(omax(f - f1, 0.0), EPSMIN);
        const double g = CUR ? __ldg(fl2 + o) : f;
        const double f2 = (icen && g <= zthrs) ? omax(zr * g, zthrs * (zr * zr)) : g;
        fmax = omax(fmax, f2);
        if (mss) {
          const double flwd = signbit(c) ? f * 0.0 : f;      // FL1 * (0.5 + 0.5*SIGN(1,COSWDIF)) (halphap.F90:72-84)
          h_s += flwd; h_t2 += omax(flwd, EPSMIN); r_s += f;
          if (m == F - 1) h_f1d += flwd * DELTH;
        }
      }
      es += t2s * c_oc.DFIM[m]; fs += c_oc.DFIMOFR[m] * t2s;
      ew += t2w * c_oc.DFIM[m]; fw += c_oc.DFIMOFR[m] * t2w;
      if (mss) {
        const double wn = __ldg(d.f.wavnum + b3 + (size_t)m * P);
        const double t1 = c_oc.DFIM[m] * (wn * wn);           // meansqs_lf.F90:87-100
        h_xm += t1 * h_s;
        if (m < c_oc.NFRE_EFF) lf += t1 * r_s;
        h_em += h_t2 * c_oc.DFIM[m]; h_fm += c_oc.DFIMOFR[m] * h_t2;
      }
    }
    if (mss) {
      // HALPHAP (halphap.F90:92-115)
      const double em = h_em + (c_oc.WETAIL * c_oc.FR[F - 1] * DELTH) * h_t2;
      const double fm = omax(em / (h_fm + (c_oc.FRTAIL * DELTH) * h_t2), c_oc.FR[0]);
      double alphap = 0.0;
      bool tail = true;
      if (em > 0.0 && fm < c_oc.FR[F - 3]) { alphap = h_xm / (log(c_oc.FR[F - 1]) - log(fm)); tail = alphap > c_oc.ALPHAPMAX; }
      if (tail) alphap = c_oc.ZPI4GM2_FR5N * h_f1d;
      const double halp = 0.5 * omin(alphap, c_oc.ALPHAPMAX);
      // MEANSQS_GC (meansqs_gc.F90:60-82) with OMEGAGC / NS_GC
      const int N = c_oc.NWAV_GC;
      const double xks0 = c_oc.SQRTGOSURFT / (1.48 + 2.05 * ufric);
      int ns = min((int)(log(omax(xks0 * c_oc.XKM1_GC, 1.0)) * c_oc.XLOGKRATIOM1_GC) + 1, N - 1);
      const double xks = __ldg(d.gc + GC_XK * N + ns - 1), oms = __ldg(d.gc + GC_OMEGA * N + ns - 1);
      const double frgc = oms / c_oc.ZPI;
      double cg;
      if (xks > c_oc.XKMSS) { ns = c_oc.NE_MSS; cg = 0.0; }
      else cg = __ldg(d.gc + GC_DELKCC_NS * N + ns - 1) * (1.0 / __ldg(d.gc + GC_XK * N + ns - 1));
      for (int i = ns + 1; i <= c_oc.NE_MSS; ++i) cg = cg + __ldg(d.gc + GC_DELKCC * N + i - 1) * (1.0 / __ldg(d.gc + GC_XK * N + i - 1));
      cg = cg * (__ldg(d.gc + GC_C2OSQRTVG * N + ns - 1) * halp);
      // meansqs.F90:90-99
      const double tl = 2.0 * halp * omax(log(omin(frgc, c_oc.FCUT_MSS)) - log(c_oc.FR[c_oc.NFRE_EFF - 1]), 0.0);
      xmss9 = (cg + lf) + tl;
    }
    const double d25 = c_oc.WETAIL * c_oc.FR[F - 1] * DELTH, d2 = c_oc.FRTAIL * DELTH;
    const double fswell = omax((es + d25 * t2s) / (fs + d2 * t2s), c_oc.FR[0]);
    const double fsea = omax((ew + d25 * t2w) / (fw + d2 * t2w), c_oc.FR[0]);
    Rf = fswell > 0.96 * fsea ? 1.0 : 0.0;
  }
  const double fcrop = 0.1 * fmax;    // dominant_period.F90:63,89

  // ---------------- pass 2: frequency sweep NFRE -> 1 with the final swell mask
  // swell mask of row m (bit k): first guess + extended wind sector (sepwisw.F90:138-212)
  auto base_mask = [&](int m, int k, double xinv, double c) -> bool {
    const double x = __ldg(xl + (size_t)m * rs + (size_t)k * P);
    bool sw = (x != 0.0) ? false : !(xinv * (COEF * c) >= 1.0);
    if (xinv * ((Rf * COEF) * copysign(1.0, 0.4 + c)) >= 1.0) sw = false;
    return sw;
  };
  unsigned long long cur = 0ull, done = 0ull;
  {
    const double xinv = ufric * __ldg(cinv + (size_t)(F - 1) * P);
    for (int k = 0; k < A; ++k) if (base_mask(F - 1, k, xinv, cw[k * OB_NTH])) cur |= 1ull << k;
  }
  SpecAcc S, W, T;
  S.zero(); W.zero(); T.zero();
  double wds = 0.0, rlast = 0.0, em4 = 0.0, dp4 = 0.0;          // WDIRSPREAD (LLPEAKF=F), DOMINANT_PERIOD
  double wfm = 0.0, wfx = 0.0, wfy = 0.0, wt0 = 0.0, wts = 0.0, wtc = 0.0;   // WEFLUX
  double ebt[7];                                                               // SE10MEAN + the six SEBTMEAN period bands
#pragma unroll
  for (int b = 0; b < 7; ++b) ebt[b] = EPSMIN;
  for (int m = F - 1; m >= 0; --m) {
    const double xinv1 = m > 0 ? ufric * __ldg(cinv + (size_t)(m - 1) * P) : 0.0;
    const double zr = icen ? exp(-10.0 * (c_oc.FR[m] * c_oc.FR[m]) / wsq) : 1.0;
    const bool noise_rows = (m + 1) >= F / 2;
    unsigned long long nxt = 0ull;
    double s_t2 = 0, s_t0 = 0, s_s = 0, s_c = 0, s_f1 = 0;
    double w_t2 = 0, w_t0 = 0, w_s = 0, w_c = 0, w_f1 = 0;
    double t_t2 = 0, t_t0 = 0, t_s = 0, t_c = 0, t_dp = 0;
    double r_t0 = 0, r_s = 0, r_c = 0;
    OB_UNROLL
    for (int k = 0; k < A; ++k) {
      const size_t o = (size_t)m * rs + (size_t)k * P;
      const double f = __ldg(fl + o), c = cw[k * OB_NTH];
      const unsigned long long bit = 1ull << k;
      const bool curk = (cur & bit) != 0ull;
      if (m > 0) {
        // connect the low-frequency boundary of the wind-sea area (sepwisw.F90:216-226)
        bool nb = base_mask(m - 1, k, xinv1, c);
        if (!(done & bit)) {
          if (curk && nb) done |= bit;
          else if (!curk && nb && f >= __ldg(fl + o - rs)) nb = false;
        }
        if (nb) nxt |= bit;
      }
      const double sth = c_oc.SINTH[k], cth = c_oc.COSTH[k];
      const double fs_ = curk ? omax(f, EPSMIN) : 0.0;                                   // swell spectrum (sepwisw.F90:231-237)
      double fw_ = f - fs_;                                                              // wind sea (sepwisw.F90:271-282)
      if (c > 0.8 && noise_rows) { const double c2 = c * c; fw_ = fw_ + EPSMIN * (c2 * c2); }
      fw_ = omax(fw_, 0.0);
      const double g = CUR ? __ldg(fl2 + o) : f;
      const double f2 = (icen && g <= zthrs) ? omax(zr * g, zthrs * (zr * zr)) : g;     // output spectrum
      s_t2 += omax(fs_, EPSMIN); s_t0 += fs_; s_s += sth * fs_; s_c += cth * fs_; s_f1 = __dadd_rn(s_f1, __dmul_rn(fs_, DELTH));
      w_t2 += omax(fw_, EPSMIN); w_t0 += fw_; w_s += sth * fw_; w_c += cth * fw_; w_f1 = __dadd_rn(w_f1, __dmul_rn(fw_, DELTH));
      t_t2 += omax(f2, EPSMIN); t_t0 += f2; t_s += sth * f2; t_c += cth * f2;
      if (f2 > fcrop) t_dp += f2 * DELTH;
      r_t0 += f; r_s += sth * f; r_c += cth * f;
    }
    S.row(m, s_t2, s_t0, s_s, s_c, s_f1);
    W.row(m, w_t2, w_t0, w_s, w_c, w_f1);
    T.row(m, t_t2, t_t0, t_s, t_c, 0.0);
    const double rr = DELTH * sqrt(t_s * t_s + t_c * t_c);      // scosfl.F90:88-104 for row m
    wds += rr * c_oc.DFIM[m];
    if (m == F - 1) { rlast = rr; wt0 = r_t0; wts = r_s; wtc = r_c; }
    const double q2 = t_dp * t_dp, q4 = q2 * q2;
    em4 += c_oc.DFIM[m] * q4; dp4 += c_oc.DFIMFR[m] * q4;
#pragma unroll
    for (int b = 0; b < 7; ++b) ebt[b] += c_oc.SEBT[b][m] * t_t0;
    const double cg = __ldg(cgr + (size_t)m * P);
    wfm += c_oc.DFIM[m] * (cg * r_t0); wfx += c_oc.DFIM[m] * (cg * r_s); wfy += c_oc.DFIM[m] * (cg * r_c);
    cur = nxt;
  }

  // ---------------- closures and the output buffer (outblock.F90:216-600)
  double EM, FM, ESW, FSW, ESE, FSE;
  T.femean(EM, FM); S.femean(ESW, FSW); W.femean(ESE, FSE);
  const double DEG = c_oc.DEG, ZMISS = c_oc.zmiss;
  const int NT = 3, NW = 6;   // NTRAIN (yowcout.F90:19), NTEWH (mpcrtbl.F90:373-399)
  const int iodp = d.iodp ? d.iodp[p] : 1;
  double wefdir = 0.0, wefmag = 0.0;
  {
    const double dl = c_oc.FRTAIL * DELTH * c_oc.G / (2.0 * c_oc.ZPI);   // weflux.F90:84,108-130
    wefmag = c_oc.ROWATER * c_oc.G * (wfm + dl * wt0);
    const double x = wfx + dl * wts;
    double y = wfy + dl * wtc;
    if (y == 0.0) y = EPSMIN;
    wefdir = atan2(x, y);
    if (wefdir < 0.0) wefdir += c_oc.ZPI;
  }
  for (int i = 0; i < c_oc.ncol; ++i) {
    const int itg = c_oc.itg[i];
    double v = 0.0;
    switch (itg) {
      case 1: v = 4.0 * sqrt(omax(EM, 0.0)); break;
      case 2: v = fmod(DEG * T.thq() + 180.0, 360.0); break;
      case 3: v = FM > 0.0 ? 1.0 / FM : ZMISS; break;
      case 4: v = ufric; break;
      case 5: v = fmod(DEG * wd + 180.0, 360.0); break;
      case 6: { const double dp = (em4 > 0.0 && dp4 > EPSMIN) ? em4 / dp4 : 0.0; v = dp > 0.0 ? dp : ZMISS; } break;
      case 7: {   // outbeta.F90:66-91, LLGCBZ0 = F
        const double amx = c_oc.llgcbz0 ? c_oc.ALPHAMAX : omin(c_oc.ALPHAMAX, 0.02 + 0.01 * wsw);   // outbeta.F90:113-117
        const double usm = 1.0 / omax(ufric, c_oc.EPSUS);
        const double betam = omax(omin(d.f.chrnck[p], amx), c_oc.ALPHAMIN);
        const double z0atm = c_oc.rnum * usm + c_oc.GM1 * betam * (ufric * ufric);
        const double q = c_oc.XKAPPA / log(1.0 + c_oc.XNLEV / z0atm);
        v = omin(q * q, 0.01);
      } break;
      case 8: v = d.f.tauw[p] / omax(ufric * ufric, c_oc.EPSUS); break;
      case 9: v = xmss9; break;
      case 10: v = wsw; break;
      case 11: v = 4.0 * sqrt(omax(ESE, 0.0)); break;
      case 12: v = 4.0 * sqrt(omax(ESW, 0.0)); break;
      case 13: v = fmod(DEG * (ESE <= 1.0e-9 ? wd : W.thq()) + 180.0, 360.0); break;
      case 14: v = fmod(DEG * S.thq() + 180.0, 360.0); break;
      case 15: v = FSE > 0.0 ? 1.0 / FSE : ZMISS; break;
      case 16: v = FSW > 0.0 ? 1.0 / FSW : ZMISS; break;
      case 20: v = T.mwp(1); break;
      case 21: v = T.mwp(2); break;
      case 22: {   // wdirspread.F90:120-135, LLPEAKF = F
        double x = wds / DELTH + rlast * (c_oc.WETAIL * c_oc.FR[F - 1]);
        x = EM > EPSMIN ? omin(x / EM, 1.0) : 1.0;
        v = sqrt(2.0 * (1.0 - x));
      } break;
      case 23: v = W.mwp(1); break;
      case 24: v = S.mwp(1); break;
      case 25: v = W.mwp(2); break;
      case 26: v = S.mwp(2); break;
      case 27: v = W.spread_peak(); break;
      case 28: v = S.spread_peak(); break;
      case 32: v = d.f.depth[p]; break;
      case 35: v = d.f.ustokes[p]; break;
      case 36: v = d.f.vstokes[p]; break;
      case 37: v = d.f.ucur ? d.f.ucur[p] : 0.0; break;
      case 38: v = d.f.vcur ? d.f.vcur[p] : 0.0; break;
      case 39: v = d.f.phieps[p]; break;
      case 40: v = d.f.phiaw[p]; break;
      case 41: v = d.f.tauoc[p]; break;
      case 44 + 3 * NT: v = d.f.aird[p]; break;
      case 45 + 3 * NT: v = d.f.wstar[p]; break;
      case 46 + 3 * NT: v = cic; break;
      case 47 + 3 * NT: v = d.f.cithick[p]; break;
      case 43 + 3 * NT: v = 4.0 * sqrt(omax(ebt[0], 0.0)); break;
      case 55 + 3 * NT: case 56 + 3 * NT: case 57 + 3 * NT: case 58 + 3 * NT: case 59 + 3 * NT: case 60 + 3 * NT:
        v = 4.0 * sqrt(omax(ebt[itg - (54 + 3 * NT)], 0.0)); break;
      case 53 + 3 * NT: v = wefmag; break;
      case 54 + 3 * NT: v = fmod(DEG * wefdir + 180.0, 360.0); break;
      case 58 + 3 * NT + NW: v = d.f.tauxd[p]; break;
      case 59 + 3 * NT + NW: v = d.f.tauyd[p]; break;
      case 60 + 3 * NT + NW: v = d.f.tauocxd[p]; break;
      case 61 + 3 * NT + NW: v = d.f.tauocyd[p]; break;
      case 62 + 3 * NT + NW: v = omax(-d.f.phiocd[p], 0.0); break;
      default: break;
    }
    // outsetwmask.F90:62-78
    if (c_oc.licerun && c_oc.llsource && c_oc.icemask[i] == 1 && cic > c_oc.cithrsh) v = ZMISS;
    if (c_oc.seamask[i] == 1) v = v * iodp + (1 - iodp) * ZMISS;
    d.bout[(size_t)pi + (size_t)P * ((size_t)i + (size_t)c_oc.ncol * (size_t)pc)] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// MPMINMAXAVG, LLGLOBAL = F (mpminmaxavg.F90:160-176): per-rank sum / count / min / max over the own points that are not
// ZMISS.  Fixed partition and fixed tree -> the same bits on every run.
#define NM_NB 64
#define NM_NTH 256
__global__ void __launch_bounds__(NM_NTH) k_norm_partial(const double* bout, int P, int ncol, long long nloc, double zmiss, double* part /*[ncol][NM_NB][4]*/) {
  __shared__ double sh[4][NM_NTH];
  const int i = blockIdx.y, b = blockIdx.x, t = threadIdx.x;
  const long long len = (nloc + NM_NB - 1) / NM_NB;
  const long long l0 = (long long)b * len, l1 = l0 + len < nloc ? l0 + len : nloc;
  double s = 0.0, c = 0.0, mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;
  for (long long l = l0 + t; l < l1; l += NM_NTH) {
    const long long ic = l / P;
    const double v = bout[(size_t)(l - ic * P) + (size_t)P * ((size_t)i + (size_t)ncol * (size_t)ic)];
    if (v != zmiss) { s += v; c += 1.0; mn = omin(mn, v); mx = omax(mx, v); }
  }
  sh[0][t] = s; sh[1][t] = c; sh[2][t] = mn; sh[3][t] = mx;
  __syncthreads();
  for (int o = NM_NTH / 2; o > 0; o >>= 1) {
    if (t < o) {
      sh[0][t] += sh[0][t + o]; sh[1][t] += sh[1][t + o];
      sh[2][t] = omin(sh[2][t], sh[2][t + o]); sh[3][t] = omax(sh[3][t], sh[3][t + o]);
    }
    __syncthreads();
  }
  if (t < 4) part[((size_t)i * NM_NB + b) * 4 + t] = sh[t][0];
}
__global__ void k_norm_final(const double* part, int ncol, double* out /*[ncol][4]: sum, count, min, max*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncol) return;
  double s = 0.0, c = 0.0, mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;
  for (int b = 0; b < NM_NB; ++b) {
    const double* q = part + ((size_t)i * NM_NB + b) * 4;
    s += q[0]; c += q[1]; mn = omin(mn, q[2]); mx = omax(mx, q[3]);
  }
  out[i * 4 + 0] = s; out[i * 4 + 1] = c; out[i * 4 + 2] = mn; out[i * 4 + 3] = mx;
}

// own points of every column, contiguous: out[i][l] (the rank's segment of the relabelled global vector ZGLOBAL)
__global__ void k_pack_cols(const double* bout, int P, int ncol, long long nloc, double* out, long long ostride) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (l >= nloc) return;
  const long long ic = l / P;
  out[(size_t)i * ostride + l] = bout[(size_t)(l - ic * P) + (size_t)P * ((size_t)i + (size_t)ncol * (size_t)ic)];
}
// MPMINMAXAVG, LLGLOBAL = T (mpminmaxavg.F90:121-153): one strictly sequential sum per column over the ORIGINAL global
// point order (IJ = IJ2NEWIJ(IJOLD)), reproducible for any number of ranks.  One warp per column.  The sum is the reference's
// left-to-right chain, so its cost is one dependent DADD per element and nothing else may sit on that chain: the 32 lanes fetch 32
// consecutive elements one block AHEAD (the gathers overlap the chain of the block before), missing values enter the chain as
// +0.0, every lane replays the same 32 unrolled additions (the shuffles pipeline in front of them), and count / minimum / maximum
// are order-independent: per-lane partials, reduced once at the end.  (First version: shuffle, test and four updates per element
// inside a rolled loop, ~55 cycles per element = 31 ms per call at O640; now ~18 = the latency of the dependent FP64 add, 11 ms;
// feeding the chain from shared memory instead of shuffles changes nothing.)
__global__ void __launch_bounds__(32) k_norm_seq(const double* zg /*[ncol][niblo] relabelled order*/, const int* ij2new /*(0:NIBLO) or null*/,
                                                 long long niblo, double zmiss, double* out /*[ncol][4]: avg, min, max, count*/) {
  const int i = blockIdx.x, lane = threadIdx.x;
  const double* z = zg + (size_t)i * niblo;
  double s = 0.0, mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;
  long long cnt = 0;
  auto fetch = [&](long long j) { return j < niblo ? z[ij2new ? (long long)ij2new[j + 1] - 1 : j] : zmiss; };
  double v = fetch(lane);
  for (long long j0 = 0; j0 < niblo; j0 += 32) {
    const double vnext = fetch(j0 + 32 + lane);
    const bool ok = v != zmiss;
    if (ok) { cnt += 1; mn = omin(mn, v); mx = omax(mx, v); }
    const double vv = ok ? v : 0.0;
#pragma unroll
    for (int q = 0; q < 32; ++q) s = s + __shfl_sync(0xffffffffu, vv, q);
    v = vnext;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    mn = omin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = omax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) {
    out[i * 4 + 0] = s / (double)(cnt > 1 ? cnt : 1); out[i * 4 + 1] = mn; out[i * 4 + 2] = mx; out[i * 4 + 3] = (double)cnt;
  }
}
}  // namespace

void launch_newwind_ustar(long long npts, const ecwam_b200_fields& f, const ecwam_b200_forcing_next& nx, const double* ufric_next, double alpha,
                          cudaStream_t st) {
  if (npts <= 0) return;
  k_newwind_ustar<<<(unsigned)((npts + 255) / 256), 256, 0, st>>>(npts, f, nx, ufric_next, alpha);
}
void launch_newwind(long long npts, const ecwam_b200_fields& f, const ecwam_b200_forcing_next& nx, double acd, double bcd, double epsmin,
                    cudaStream_t st) {
  if (npts <= 0) return;
  k_newwind<<<(unsigned)((npts + 255) / 256), 256, 0, st>>>(npts, f, nx, acd, bcd, epsmin);
}
void launch_getwnd(const GetwndArgs& a, const ecwam_b200_fieldg& g, const ecwam_b200_getwnd_opts& o, const int* ifromij, const int* jfromij,
                   const ecwam_b200_forcing_next& nx, cudaStream_t st) {
  if (a.npts <= 0) return;
  k_getwnd<<<(unsigned)((a.npts + 255) / 256), 256, 0, st>>>(a, g, o, ifromij, jfromij, nx);
}
void launch_no_source(long long n4, long long n2, double* fl1, double* xllws, int* mij, int nfre, int clip, double epsmin, cudaStream_t st) {
  if (n4 <= 0) return;
  k_no_source<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(n4, n2, fl1, xllws, mij, nfre, clip, epsmin);
}
int launch_outblock(const OutDev& d, cudaStream_t st) {
  if (d.npts <= 0) return 0;
  const size_t sm = (size_t)d.A * OB_NTH * sizeof(double);
  static bool attr = false;
  if (!attr) {
    EW_CUDA_CHECK(cudaFuncSetAttribute(k_outblock<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    EW_CUDA_CHECK(cudaFuncSetAttribute(k_outblock<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr = true;
  }
  if (d.A > 64 || sm > 64 * 1024) { ew_set_error("k_outblock: NANG %d too large", d.A); return ECWAM_B200_EINVAL; }
  if (d.fl2) k_outblock<true><<<(unsigned)((d.npts + OB_NTH - 1) / OB_NTH), OB_NTH, sm, st>>>(d);
  else k_outblock<false><<<(unsigned)((d.npts + OB_NTH - 1) / OB_NTH), OB_NTH, sm, st>>>(d);
  return 0;
}

// INTPOL (intpol.F90:96-271, IRA = 1): the spectrum on the absolute frequency axis from the one relative to the current.  One thread
// per grid point walks the (frequency, direction) bins in the reference's order and adds each bin's two shares to its own column of
// `fla` (same (P,A,F,C) layout as FL1; the column is private to the thread: no atomics, the reference's summation order).
__global__ void __launch_bounds__(128) k_intpol(OutDev d, double* __restrict__ fla, double fratio, double flogsprdm1, double fr5n) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= d.npts) return;
  const int A = c_oc.A, F = c_oc.F, P = d.P;
  const long long pc = p / P;
  const int pi = (int)(p - pc * P);
  const size_t rs = (size_t)P * A;
  const double* fl = d.f.fl1 + (size_t)pi + rs * F * (size_t)pc;
  double* fa = fla + (size_t)pi + rs * F * (size_t)pc;
  const double* wnp = d.f.wavnum + (size_t)pi + (size_t)P * F * (size_t)pc;
  const double u = d.f.ucur[p], v = d.f.vcur[p];
  const double CURRENT_MAX = 1.5;      // yowcurr.F90:18
  const double fre0 = fratio - 1.0, zpi2gm = c_oc.ZPI * c_oc.ZPI / c_oc.G, coef = 1.0 / c_oc.ZPI;
  const double fr1 = c_oc.FR[0], frn = c_oc.FR[F - 1];
  const double fmax = frn + (c_oc.ZPI / c_oc.G) * (frn * frn) * CURRENT_MAX;
  const int nfre_max = (int)floor(log10(fmax / fr1) * flogsprdm1) + 1;
  const double cdf = 0.5 * (fratio - 1.0 / fratio) * c_oc.DELTH;
  bool ice2sea = true;
  for (int m = 0; m < F; ++m)
    for (int k = 0; k < A; ++k) {
      const size_t o = (size_t)m * rs + (size_t)k * P;
      if (__ldg(fl + o) > c_oc.EPSMIN) ice2sea = false;
      fa[o] = 0.0;
    }
  double freq = frn;
  for (int m = 0; m < nfre_max; ++m) {
    double dfreqth, wavn;
    if (m < F) { freq = c_oc.FR[m]; dfreqth = freq * cdf; wavn = __ldg(wnp + (size_t)m * P); }
    else { freq = frn * pow(fratio, (double)(m + 1 - F)); dfreqth = freq * cdf; wavn = zpi2gm * (freq * freq); }
    const double f5 = fr5n / pow(freq, 5.0);
    for (int k = 0; k < A; ++k) {
      double fnew = freq + coef * wavn * (c_oc.COSTH[k] * v + c_oc.SINTH[k] * u);
      int kh = k;
      if (!(fnew > 0.0)) { kh = (k + A / 2) % A; fnew = -fnew; }
      int newm = -1;                                   // 1-based bin below FNEW, 0: between FR(1)/FRATIO and FR(1)
      if (!(fnew <= fr1 / fratio)) newm = (int)floor(log10(fnew / fr1) * flogsprdm1) + 1;
      double old = 0.0;
      if (!ice2sea) old = m < F ? __ldg(fl + (size_t)m * rs + (size_t)k * P) : __ldg(fl + (size_t)(F - 1) * rs + (size_t)k * P) * f5;
      if (newm < F && newm >= 1) {
        const double fa0 = c_oc.FR[newm - 1], fa1 = c_oc.FR[newm];
        const double gwh = dfreqth / (fa1 - fa0) * old;
        fa[(size_t)(newm - 1) * rs + (size_t)kh * P] += gwh * (fa1 - fnew) / (fa0 * cdf);
        fa[(size_t)newm * rs + (size_t)kh * P] += gwh * (fnew - fa0) / (fa1 * cdf);
      } else if (newm == 0) {
        const double gwh = fratio * dfreqth / (fre0 * fr1) * old;
        fa[(size_t)kh * P] += gwh * (fnew - fr1 / fratio) / (fr1 * cdf);
      } else if (newm == F) {
        const double gwh = dfreqth / (fre0 * frn) * old;
        fa[(size_t)(F - 1) * rs + (size_t)kh * P] += gwh * (fratio * frn - fnew) / (frn * cdf);
      }
    }
  }
  for (int m = 0; m < F; ++m)
    for (int k = 0; k < A; ++k) { const size_t o = (size_t)m * rs + (size_t)k * P; fa[o] = omax(fa[o], c_oc.EPSMIN); }
}
void launch_intpol(const OutDev& d, double* fla, double fratio, double flogsprdm1, double fr5n, cudaStream_t st) {
  if (d.npts <= 0) return;
  k_intpol<<<(unsigned)((d.npts + 127) / 128), 128, 0, st>>>(d, fla, fratio, flogsprdm1, fr5n);
}
size_t norm_scratch_doubles(int ncol) { return (size_t)ncol * NM_NB * 4 + (size_t)ncol * 4; }
void launch_norm_local(const double* bout, int P, int ncol, long long nloc, double zmiss, double* scratch, double* out4, cudaStream_t st) {
  k_norm_partial<<<dim3(NM_NB, ncol), NM_NTH, 0, st>>>(bout, P, ncol, nloc, zmiss, scratch);
  k_norm_final<<<(ncol + 63) / 64, 64, 0, st>>>(scratch, ncol, out4);
}
void launch_pack_cols(const double* bout, int P, int ncol, long long nloc, double* out, long long ostride, cudaStream_t st) {
  if (nloc <= 0) return;
  k_pack_cols<<<dim3((unsigned)((nloc + 255) / 256), ncol), 256, 0, st>>>(bout, P, ncol, nloc, out, ostride);
}
void launch_norm_seq(const double* zg, const int* ij2new, long long niblo, int ncol, double zmiss, double* out4, cudaStream_t st) {
  k_norm_seq<<<ncol, 32, 0, st>>>(zg, ij2new, niblo, zmiss, out4);
}

}  // namespace ew
