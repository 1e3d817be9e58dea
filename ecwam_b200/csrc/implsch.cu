// IMPLSCH (src/ecwam/implsch.F90:10-468) for sm_100a: the implicit source-term step of WAMINTGR.
//
// The reference runs the whole call tree per NPROMA chunk with every (IJ,K,M) temporary in memory
// (implsch.F90:152-170).  Here the step is two kernels:
//   k_point    lane = grid point: the parts that stream over the spectrum with per-point state (SDEPTHLIM, FKMEAN, both
//              SINFLX calls with AIRSEA/TAUT_Z0, SINPUT, FEMEANWS, FRCUTINDEX, STRESSO/TAU_PHI_HF, WSIGSTAR, SDIWBK's Q)
//   k_stencil  CTA = 8 grid points x NANG threads: the parts that couple spectral bins (SDISSIP saturation window, SNONLIN
//              DIA quadruplets) fused with SDIWBK, SBOTTOM, the implicit update, WNFLUXES, IMPHFTAIL, SETICE, STOKESDRIFT
// Nothing but FL1, XLLWS, the 1-D outputs and one scratch array (the wind-input linearisation) touches HBM.
#include "internal.h"
#include <cstdlib>
#include <cstring>

namespace ew {

static_assert(sizeof(DevConst) % 16 == 0, "keep sizeof(DevConst) a multiple of 16 (alignment of the constants behind it)");
__constant__ DevConst c_dc;
int upload_dev_const(const DevConst& h, cudaStream_t st) {
  EW_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_dc, &h, sizeof(DevConst), 0, cudaMemcpyHostToDevice, st));
  return 0;
}

// scalar scratch slots [slot][npts]
enum {
  S_EMEAN = 0, S_FMEAN, S_F1MEAN, S_AKMEAN, S_XKMEAN,   // FKMEAN (first call)
  S_XSTR, S_YSTR, S_F1DCOS3, S_F1DCOS2, S_PHIWA,         // S_XSTR / S_YSTR: XSTRESSICE / YSTRESSICE of LWNEMOCOUWRS (k_ice -> k_nemo); PHIWA
  S_UORBT, S_AORB, S_SIGN, S_TEMP2, S_PTURB, S_PVISC,    // SINPUT_ARD swell-dissipation scalars, WSIGSTAR
  S_SDS,                                                  // SDIWBK
  S_PHILF, S_XSTROC, S_YSTROC,                            // WNFLUXES sums
  S_MIJ, S_USTOLD, S_FAC, S_USFM,
  S_USTAR1, S_TAUW1, S_TWDIR1,                            // result of the first SINFLX call (k_point<.,1> -> k_point<.,2>)
  S_HALP,                                                 // HALPHAP of the first SINFLX call (LLGCBZ0)
  NSCR
};
size_t implsch_scratch_doubles(long long npts) { return (size_t)NSCR * (size_t)npts; }
// planes of ImplDev::tbg [TQ_N][F][npts]: per-(point, frequency) scalars that k_point derives once for k_stencil
enum { TQ_FACSAT = 0, TQ_SBO, TQ_CINV, TQ_TAIL, TQ_STF, TQ_JAN, TQ_N };
static_assert(TQ_N == EW_TQ_N, "tbg planes");

#define FULLMASK 0xffffffffu
// max / min without fmax's NaN bookkeeping (DSETP + 2 selects instead of ~8 instructions); no NaNs occur on this path
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }
// a / b for b in the normal range (no overflow / denormal / zero paths): reciprocal seed, 2 Newton steps, residual correction
__device__ __forceinline__ double div_norm(double a, double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}
// a / b as div_norm with one Newton step less: the seed has >= 20 bits, so r is good to 2^-40 and the residual step to 2^-80
__device__ __forceinline__ double div_fast(double a, double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  const double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}
// exp(x) for x <= 0 without libm's range checks: k = nint(x log2 e) by the magic-number add, r = x - k ln2 (two-part ln2),
// 13th-degree Taylor polynomial on |r| <= ln2 / 2 (remainder 4e-18), scaling by 2^k through the exponent field.  x < -708 is
// clamped (3e-308 instead of a denormal).  Relative error <= 2.3e-16 over [-708, 0] (checked against libm on 6e5 samples).
#ifndef KP_EXP
#define KP_EXP 1
#endif
#ifndef KP_ESTRIN
#define KP_ESTRIN 0   // 1: Estrin's scheme for the polynomial (5 dependent levels instead of 13): k_point 19.6 ms vs 17.3 -- slower
#endif
// (the coefficients live in constant memory: as literals every one of them costs two UMOV per use, 93 per bin of the second SINPUT pass)
__constant__ double c_exp[20] = {1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,
                                 2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03,
                                 8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5, 1.0, 1.0,
                                 1.4426950408889634074, 6755399441055744.0, 6.93147180369123816490e-01, 1.90821492927058770002e-10,
                                 -708.0, 0.0};
__device__ __forceinline__ double exp_le0(double x) {
#if KP_EXP
  x = dmax(x, c_exp[18]);
  const double t = fma(x, c_exp[14], c_exp[15]);
  const double kf = t - c_exp[15];
  double r = fma(-kf, c_exp[16], x);
  r = fma(-kf, c_exp[17], r);
#if KP_ESTRIN
  // Estrin's scheme: the same polynomial as five dependent levels instead of Horner's thirteen (a dependent FP64 instruction
  // issues ~18 cycles after its producer, and this chain is most of a bin's latency); c_exp[13 - i] is the coefficient of r^i
  const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  const double a0 = fma(c_exp[12], r, c_exp[13]), a1 = fma(c_exp[10], r, c_exp[11]), a2 = fma(c_exp[8], r, c_exp[9]),
               a3 = fma(c_exp[6], r, c_exp[7]), a4 = fma(c_exp[4], r, c_exp[5]), a5 = fma(c_exp[2], r, c_exp[3]),
               a6 = fma(c_exp[0], r, c_exp[1]);
  const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
  const double d0 = fma(b1, r4, b0), d1 = fma(a6, r4, b2);
  const double p = fma(d1, r8, d0);
#else
  double p = c_exp[0];                        // 1/13!, ..., 1/2!, 1, 1
#pragma unroll
  for (int i = 1; i < 14; ++i) p = fma(p, r, c_exp[i]);
#endif
  return p * __hiloint2double((1023 + __double2loint(t)) << 20, 0);
#else
  return exp(x);
#endif
}
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = dmax(v, __shfl_xor_sync(FULLMASK, v, o));
  return v;
}
__device__ __forceinline__ double sq(double x) { return x * x; }
__device__ __forceinline__ double p4(double x) { double y = x * x; return y * y; }

// (P,C) field element of point p
#define PT2(ptr, p) ((ptr)[(p)])
// (P,F,C) field element (m 0-based) of point p = i + P*c
__device__ __forceinline__ size_t idx3(const ImplDev& d, long long p, int m) {
  const long long c = p / d.P;
  const int i = (int)(p - c * d.P);
  return (size_t)i + (size_t)d.P * ((size_t)m + (size_t)d.F * (size_t)c);
}

// ---------------------------------------------------------------------------------------------------------
// chnkmin.F90 ; taut_z0.F90:281-341 (LLGCBZ0=F) through airsea.F90 (ICODE_WND=3)
__device__ __forceinline__ double chnkmin(double u10) {
  return c_dc.ALPHAMIN + (c_dc.ALPHA - c_dc.ALPHAMIN) * 0.5 * (1.0 - tanh(u10 - c_dc.CHNKMIN_U));
}
__device__ void taut_z0(int iusfg, double utop, double udir, double tauw, double tauwdir, double& ustar, double& z0,
                        double& z0b, double& chrnck) {
  const int NITER = 18;
  const double TWOXMP1 = 3.0;
  const double xlogxl = log(c_dc.XNLEV);
  const double us2totauw = 1.0 + c_dc.EPS1;
  const double cosdiff = cos(udir - tauwdir);
  const double tauwact = dmax(tauw * cosdiff, c_dc.EPSMIN);
  const double tauweff = tauwact * us2totauw;
  double xmin, alphaog;
  if (c_dc.llcapchnk) {
    const double cm = chnkmin(utop);
    xmin = 0.15 * (c_dc.ALPHA - cm);
    alphaog = cm * c_dc.GM1;
  } else {
    xmin = 0.0;
    alphaog = c_dc.ALPHA * c_dc.GM1;
  }
  const double xkutop = c_dc.XKAPPA * utop;
  const double ustold = (1 - iusfg) * utop * sqrt(dmin(c_dc.ACD + c_dc.BCD * utop, c_dc.CDMAX)) + iusfg * ustar;
  double tauold = dmax(sq(ustold), tauweff);
  ustar = sqrt(tauold);
  double ustm1 = 1.0 / dmax(ustar, c_dc.EPSUS);
  double z0ch = 0.0;
  for (int iter = 1; iter <= NITER; ++iter) {
    const double x = dmax(tauwact / tauold, xmin);
    z0ch = alphaog * tauold / sqrt(1.0 - x);
    const double z0vis = c_dc.rnum * ustm1;
    const double z0tot = z0ch + z0vis;
    const double xologz0 = 1.0 / (xlogxl - log(z0tot));
    const double f = ustar - xkutop * xologz0;
    const double zz = ustm1 * (z0ch * (2.0 - TWOXMP1 * x) / (1.0 - x) - z0vis) / z0tot;
    const double delf = 1.0 - xkutop * sq(xologz0) * zz;
    if (delf != 0.0) ustar = ustar - f / delf;
    const double taunew = dmax(sq(ustar), tauweff);
    ustar = sqrt(taunew);
    if (taunew == tauold) break;
    ustm1 = 1.0 / dmax(ustar, c_dc.EPSUS);
    tauold = taunew;
  }
  z0 = z0ch;
  z0b = alphaog * tauold;
  chrnck = dmax(c_dc.G * z0 * sq(ustm1), c_dc.ALPHAMIN);
}

// ---------------------------------------------------------------------------------------------------------
// LLGCBZ0, the gravity-capillary model of the background roughness: ns_gc.F90:44-48, stress_gc.F90:70-130, cdm.func.h and
// the first branch of taut_z0.F90 (:148-279).  gc = ImplDev::gc, [GC_NT][NWAV_GC].
__device__ __forceinline__ int ns_gc(double ustar) {   // 1-based index of the first gravity-capillary wavenumber
  const double xks = c_dc.SQRTGOSURFT / (1.48 + 2.05 * ustar);
  return min((int)(log(dmax(xks * c_dc.XKM1_GC, 1.0)) * c_dc.XLOGKRATIOM1_GC) + 1, c_dc.NWAV_GC - 1);
}
__device__ double stress_gc(const double* __restrict__ gc, double ang_gc, double ustar, double z0, double z0min, double halp,
                            double rnfac) {
  const int N = c_dc.NWAV_GC;
  const int ns = ns_gc(ustar);
  const double tauwcg_min = sq(ustar * (z0min / z0));
  const double xlambda = 1.0 + 0.25 * tanh(4.0 * p4(ustar));
  const double c2 = __ldg(gc + GC_C2OSQRTVG * N + ns - 1);
  const double zabhrc = ang_gc * c_dc.BETAMAXOXKAPPA2 * halp * c2;
  const double cst = c_dc.llnormagam ? rnfac * c_dc.BMAXOKAP * halp * c2 / dmax(ustar, c_dc.EPSUS) : 0.0;
  const double logl = log(xlambda);
  // LOG(XK_GC(I)*Z0) = log XK_GC(I) (a table row) + log Z0 (once per call), and exp_le0 (valid up to +709) for EXP(XLOG): the loop runs
  // ~26 wavenumbers x 18 iterations x 2 calls per point and grid point, libm's log + exp were most of the cy49r1 instance's extra cost
  const double lz0 = log(z0);
  double tauwcg = 0.0;
  for (int i = ns; i <= N; ++i) {
    const double x = ustar * __ldg(gc + GC_CM * N + i - 1);
    const double xlog = (__ldg(gc + GC_LXK * N + i - 1) + lz0) + div_fast(c_dc.XKAPPA, x + c_dc.ZALP);
    const double zlog = dmin(xlog - logl, 0.0);
    const double zlog2x = zlog * zlog * x;
    const double gam_w = zlog2x * zlog2x * exp_le0(xlog) * __ldg(gc + GC_OM3GMKM * N + i - 1);
    const double zn = cst * __ldg(gc + GC_XKMSQRTVGOC2 * N + i - 1) * gam_w;
    const double gamnorma = div_fast(1.0 + c_dc.RN1_RN * zn, 1.0 + zn);
    if (i == ns) tauwcg = gam_w * __ldg(gc + GC_DELKCC_NS * N + ns - 1) * __ldg(gc + GC_OMXKM3 * N + ns - 1) * gamnorma;
    else tauwcg = tauwcg + gam_w * __ldg(gc + GC_DELKCC_OMXKM3 * N + i - 1) * gamnorma;
  }
  return dmax(zabhrc * tauwcg, tauwcg_min);
}
__device__ __forceinline__ double cdm(double u) { return dmax(dmin(0.0006 + 0.00008 * u, 0.001 + 0.0018 * exp(-0.05 * (u - 33.))), 0.001); }
__device__ void taut_z0_gc(const double* __restrict__ gc, int iusfg, double halp, double utop, double udir, double tauw,
                           double tauwdir, double rnfac, double& ustar, double& z0, double& z0b, double& chrnck) {
  const int NITER = 18;
  const double PMAX = 0.99, Z0MIN = 0.000001;
  const double us2totauw = 1.0 + c_dc.EPS1;
  const double rnukappam1 = (0.04 * c_dc.rnu) / c_dc.XKAPPA;
  const double pce_gc = 0.001 * iusfg + (1 - iusfg) * 0.005;
  const double cosdiff = cos(udir - tauwdir);
  const double tauwact = dmax(tauw * cosdiff, c_dc.EPSMIN);
  const double alphaog = c_dc.llcapchnk ? chnkmin(utop) * c_dc.GM1 : 0.0;
  const double usmax = dmax(-0.21339 + 0.093698 * utop - 0.0020944 * (utop * utop) + 5.5091E-5 * (utop * utop * utop), 0.03);
  const double tauweff = dmin(tauwact * us2totauw, usmax * usmax);
  if (iusfg == 0) {
    double cdfg;
    if (utop < 1.0) cdfg = 0.002;
    else if (cosdiff > 0.9) {
      const double x = dmin(tauwact / sq(dmax(ustar, c_dc.EPSUS)), PMAX);
      double zchar = dmin((c_dc.ALPHA * c_dc.GM1) * sq(ustar) / sqrt(1.0 - x), 0.05 * exp(-0.05 * (utop - 35.)));
      zchar = dmin(zchar, c_dc.ALPHAMAX);
      cdfg = c_dc.ACDLIN + c_dc.BCDLIN * sqrt(zchar) * utop;
    } else cdfg = cdm(utop);
    ustar = utop * sqrt(cdfg);
  }
  const double w1 = 0.85 - 0.05 * (tanh(10.0 * (utop - 5.0)) + 1.0);
  const double xkutop = c_dc.XKAPPA * utop;
  double ustold = ustar, tauold = ustold * ustold, tauunr = 0.0, x;
  int iter;
  for (iter = 1; iter <= NITER; ++iter) {
    z0 = dmax(c_dc.XNLEV / (exp(dmin(xkutop / ustold, 50.0)) - 1.0), Z0MIN);
    const double tauv = rnukappam1 * ustold / z0;
    const double ang_gc = c_dc.ANG_GC_A + c_dc.ANG_GC_B * tanh(c_dc.ANG_GC_C * tauold);
    tauunr = stress_gc(gc, ang_gc, ustar, z0, Z0MIN, halp, rnfac);
    const double taunew = tauweff + tauv + tauunr;
    const double ustnew = sqrt(taunew);
    ustar = w1 * ustold + (1.0 - w1) * ustnew;
    const double del = ustar - ustold;
    if (fabs(del) < pce_gc * ustar) break;
    tauold = sq(ustar);
    ustold = ustar;
  }
  x = tauweff / tauold;
  if (iter > NITER && x >= PMAX) {   // protection just in case there is no convergence
    ustar = utop * sqrt(cdm(utop));
    const double z0minrst = sq(ustar) * c_dc.ALPHA * c_dc.GM1;
    z0 = dmax(c_dc.XNLEV / (exp(xkutop / ustar) - 1.0), z0minrst);
    z0b = z0minrst;
  } else {
    z0 = dmax(c_dc.XNLEV / (exp(xkutop / ustar) - 1.0), Z0MIN);
    z0b = z0 * sqrt(tauunr / tauold);
  }
  if (x < PMAX) {                    // refine the solution (taut_z0.F90:230-276)
    const double usnrf = ustar, z0nrf = z0, z0bnrf = z0b;
    ustold = ustar;
    tauold = dmax(ustold * ustold, tauweff);
    const double alpog = dmax(dmin(z0b / tauold, c_dc.ALPHAMAX), alphaog);
    for (iter = 1; iter <= NITER; ++iter) {
      x = dmin(tauweff / tauold, PMAX);
      const double ustm1 = 1.0 / dmax(ustold, c_dc.EPSUS);
      const double z0vis = c_dc.rnum * ustm1;
      const double hz0viso1mx = 0.5 * z0vis / (1.0 - x);
      z0b = alpog * tauold;
      z0 = hz0viso1mx + sqrt(sq(hz0viso1mx) + sq(z0b) / (1.0 - x));
      const double xologz0 = 1.0 / log(c_dc.XNLEV / z0 + 1.0);
      const double f = ustold - xkutop * xologz0;
      const double zz = 2.0 * ustm1 * (3.0 * sq(z0b) + 0.5 * z0vis * z0 - sq(z0)) / (2.0 * sq(z0) * (1.0 - x) - z0vis * z0);
      const double delf = 1.0 - xkutop * sq(xologz0) * zz;
      if (delf != 0.0) ustar = ustold - f / delf;
      const double taunew = dmax(sq(ustar), tauweff);
      ustar = sqrt(taunew);
      const double del = taunew - tauold;
      if (fabs(del) < pce_gc * tauold) break;
      tauold = taunew;
      ustold = ustar;
    }
    if (iter > NITER) {
      ustar = usnrf; z0 = z0nrf; z0b = z0bnrf;
      const double ustm1 = 1.0 / dmax(ustar, c_dc.EPSUS);
      chrnck = dmax(c_dc.G * (z0 - c_dc.rnum * ustm1) * sq(ustm1), c_dc.ALPHAMIN);
    } else {
      chrnck = dmax(c_dc.G * (z0b / sqrt(1.0 - x)) / sq(dmax(ustar, c_dc.EPSUS)), c_dc.ALPHAMIN);
    }
  } else {
    const double ustm1 = 1.0 / dmax(ustar, c_dc.EPSUS);
    chrnck = dmax(c_dc.G * (z0 - c_dc.rnum * ustm1) * sq(ustm1), c_dc.ALPHAMIN);
  }
}

// ---------------------------------------------------------------------------------------------------------
// tau_phi_hf.F90:111-305.  CY: the LLNORMAGAM renormalisation (CONST1, CONST2 of :177-182; both 0 otherwise, so GAMNORMA = 1)
// and LLGCBZ0's upper limit ZSUP of the TAUHF integral (:190-193); PHIHF always integrates to ZSUPMAX (:246-250).
#ifndef KP_TPH
#define KP_TPH 0     // TAU_PHI_HF nodes: 1 = one logarithm per point + exp_le0 (110 instead of 280 instructions per node, yet the kernel is 0.6 ms SLOWER: 17.87 vs 17.24 ms), 0 = libm exp / log per node
#endif
#if KP_TPH
#define TPH_EXP(x) exp_le0(x)
#define TPH_LOG2(zz, cm1) (2.0 * ((zz) + lcm1))
#else
#define TPH_EXP(x) exp(x)
#define TPH_LOG2(zz, cm1) (2.0 * log(cm1))
#endif
template <bool CY>
__device__ void tau_phi_hf(int mij, bool shelter, double z0m, double aird, double f1dcos3, double f1dcos2, double& ust,
                           double& tauhf, double& phihf, bool llphihf, double confg0 = 0.0, double f1dsin2 = 0.0, double f1d = 0.0,
                           double oms = 0.0) {   // confg0 = GAMNCONST*FR5(MIJ)*RNFAC (0 without LLNORMAGAM), oms = OMEGA_GC(NS_GC(UFRIC))
  const double ZSUPMAX = 0.0;
  const double x0g = c_dc.X0TAUHF * c_dc.G;
  double ustph = ust;
  const double xloggz0 = log(c_dc.G * z0m);
  const double omegacc = dmax(c_dc.ZPIFR[mij - 1], x0g / ust);
  const double sqrtz0og = sqrt(z0m * c_dc.GM1);
  const double sqrtgz0 = 1.0 / sqrtz0og;
  const double yc = omegacc * sqrtz0og;
  const double zinf = log(yc);
  const double consttau = c_dc.ZPI4GM2 * c_dc.FR5[mij - 1];
  double taul = sq(ust);
  double zsup = ZSUPMAX;
  double const1 = 0.0, const2 = 0.0;
  if (CY) {
    const double confg = confg0 * sqrtgz0;
    const1 = confg * f1dsin2; const2 = confg * f1d;
    if (c_dc.llgcbz0) zsup = dmin(log(oms * sqrtz0og), ZSUPMAX);
  }
  double delz = dmax((zsup - zinf) / (double)(c_dc.JTOT - 1), 0.0);
  // Y = exp(Z) on the Simpson nodes Z = ZINF + (J-1) DELZ, so LOG(CM1) = Z + log(SQRTGZ0 GM1): one logarithm per point instead of
  // one per node, and exp_le0 (straight-line, valid up to +709 as well) instead of libm's exp: 280 -> ~110 instructions per node
  const double lcm1 = log(sqrtgz0 * c_dc.GM1);
  tauhf = 0.0;
  if (shelter) {
    for (int j = 1; j <= c_dc.JTOT; ++j) {
      const double zz = zinf + (double)(j - 1) * delz;
      const double y = TPH_EXP(zz);
      const double omega = y * sqrtgz0;
      const double cm1 = omega * c_dc.GM1;
      const double zx = ust * cm1 + c_dc.ZALP;
      const double zarg = c_dc.XKAPPA / zx;
      double zlog = xloggz0 + TPH_LOG2(zz, cm1) + zarg;
      zlog = dmin(zlog, 0.0);
      const double zbeta = p4(zlog) * TPH_EXP(zlog);
      double fnc2 = f1dcos3 * consttau * zbeta * taul * c_dc.WTAUHF[j - 1] * delz;
      if (CY) { const double znz = zbeta * ust * y; fnc2 = fnc2 * ((1.0 + const1 * znz) / (1.0 + const2 * znz)); }
      taul = dmax(taul - c_dc.TAUWSHELTER * fnc2, 0.0);
      ust = sqrt(taul);
      tauhf = tauhf + fnc2;
    }
  } else {
    for (int j = 1; j <= c_dc.JTOT; ++j) {
      const double zz = zinf + (double)(j - 1) * delz;
      const double y = TPH_EXP(zz);
      const double omega = y * sqrtgz0;
      const double cm1 = omega * c_dc.GM1;
      const double zx = ust * cm1 + c_dc.ZALP;
      const double zarg = c_dc.XKAPPA / zx;
      double zlog = xloggz0 + TPH_LOG2(zz, cm1) + zarg;
      zlog = dmin(zlog, 0.0);
      const double zbeta = p4(zlog) * TPH_EXP(zlog);
      if (CY) { const double znz = zbeta * ust * y; tauhf = tauhf + (zbeta * c_dc.WTAUHF[j - 1]) * ((1.0 + const1 * znz) / (1.0 + const2 * znz)); }
      else tauhf = tauhf + zbeta * c_dc.WTAUHF[j - 1];
    }
    tauhf = f1dcos3 * consttau * taul * tauhf * delz;
  }
  phihf = 0.0;
  if (llphihf) {
    taul = sq(ustph);
    delz = dmax((ZSUPMAX - zinf) / (double)(c_dc.JTOT - 1), 0.0);
    const double constphi = aird * c_dc.ZPI4GM1 * c_dc.FR5[mij - 1];
    if (shelter) {
      for (int j = 1; j <= c_dc.JTOT; ++j) {
        const double zz = zinf + (double)(j - 1) * delz;
      const double y = TPH_EXP(zz);
        const double omega = y * sqrtgz0;
        const double cm1 = omega * c_dc.GM1;
        const double zx = ustph * cm1 + c_dc.ZALP;
        const double zarg = c_dc.XKAPPA / zx;
        double zlog = xloggz0 + TPH_LOG2(zz, cm1) + zarg;
        zlog = dmin(zlog, 0.0);
        const double zbeta = p4(zlog) * TPH_EXP(zlog);
        double fnc2 = zbeta * taul * c_dc.WTAUHF[j - 1] * delz;
        if (CY) { const double znz = zbeta * ust * y; fnc2 = fnc2 * ((1.0 + const1 * znz) / (1.0 + const2 * znz)); }
        taul = dmax(taul - c_dc.TAUWSHELTER * f1dcos3 * consttau * fnc2, 0.0);
        ustph = sqrt(taul);
        phihf = phihf + fnc2 / y;
      }
      phihf = f1dcos2 * constphi * sqrtz0og * phihf;
    } else {
      for (int j = 1; j <= c_dc.JTOT; ++j) {
        const double zz = zinf + (double)(j - 1) * delz;
      const double y = TPH_EXP(zz);
        const double omega = y * sqrtgz0;
        const double cm1 = omega * c_dc.GM1;
        const double zx = ustph * cm1 + c_dc.ZALP;
        const double zarg = c_dc.XKAPPA / zx;
        double zlog = xloggz0 + TPH_LOG2(zz, cm1) + zarg;
        zlog = dmin(zlog, 0.0);
        const double zbeta = p4(zlog) * TPH_EXP(zlog);
        if (CY) { const double znz = zbeta * ust * y; phihf = phihf + ((zbeta * c_dc.WTAUHF[j - 1]) * ((1.0 + const1 * znz) / (1.0 + const2 * znz))) / y; }
        else phihf = phihf + zbeta * c_dc.WTAUHF[j - 1] / y;
      }
      phihf = f1dcos2 * constphi * sqrtz0og * taul * phihf * delz;
    }
  }
}

// wsigstar.F90:105-129
__device__ double wsigstar(double ufric, double z0m, double wstar) {
  const double ONETHIRD = 1.0 / 3.0, SIG_NMAX = 0.9, C1 = 1.03e-3, C2 = 0.04e-3, P1 = 1.48, P2 = -0.21;
  const double xkappad = 1.0 / c_dc.XKAPPA;
  double u10 = ufric * xkappad * (log(10.0) - log(z0m));
  u10 = dmax(u10, c_dc.wspmin);
  const double u10m1 = 1.0 / u10;
  const double c2u10p1 = C2 * pow(u10, P1);
  const double u10p2 = pow(u10, P2);
  const double c_d = (C1 + c2u10p1) * u10p2;
  const double dc_ddu = (P2 * C1 + (P1 + P2) * c2u10p1) * u10p2 * u10m1;
  const double sig_conv = 1.0 + 0.5 * u10 / c_d * dc_ddu;
  return dmin(SIG_NMAX, sig_conv * u10m1 * pow(0.0 * ufric * ufric * ufric + 0.5 * c_dc.XKAPPA * wstar * wstar * wstar, ONETHIRD));
}

// wsigstar.F90:87-103 (LLGCBZ0 or LLNORMAGAM): the drag law is linearised around the model's own Charnock parameter
__device__ double wsigstar_gc(double wswave, double ufric, double z0m, double wstar) {
  const double ONETHIRD = 1.0 / 3.0, SIG_NMAX = 0.9;
  const double u10m1 = 1.0 / dmax(wswave, c_dc.wspmin);
  const double z0vis = c_dc.rnum / dmax(ufric, c_dc.EPSUS);
  double zchar = c_dc.G * (z0m - z0vis) / dmax(sq(ufric), c_dc.EPSUS);
  zchar = dmax(dmin(zchar, c_dc.ALPHAMAX), c_dc.ALPHAMIN);
  const double bcd_loc = c_dc.BCDLIN * sqrt(zchar);
  const double c_d = c_dc.ACDLIN + bcd_loc * wswave;
  const double sig_conv = 1.0 + 0.5 * wswave / c_d * bcd_loc;
  return dmin(SIG_NMAX, sig_conv * u10m1 * pow(0.0 * ufric * ufric * ufric + 0.5 * c_dc.XKAPPA * wstar * wstar * wstar, ONETHIRD));
}

// RHOWGDFTH(IJ,M) of frcutindex.F90:99-108 (m 0-based)
__device__ __forceinline__ double rhowgdfth(int m, int mij) {
  if (m + 1 > mij) return 0.0;
  double r = c_dc.RHOWG_DFIM[m];
  if (m + 1 == mij && mij != c_dc.F) r = 0.5 * r;
  return r;
}

// =========================================================================================================
// k_point: lane = grid point.  Everything of IMPLSCH that is a stream over the spectrum with per-point state:
//   SDEPTHLIM/SEMEAN, FKMEAN, both SINFLX calls (AIRSEA/TAUT_Z0, SINPUT with its sheltering recurrence over
//   frequency, FEMEANWS, FRCUTINDEX, STRESSO + TAU_PHI_HF), WSIGSTAR, the swell-friction scalars and SDIWBK's Q.
// The 32 lanes of a warp are 32 consecutive grid points, so every FL1(ij,k,m) access is one coalesced 256-byte
// row of the NPROMA-chunked array, all per-point scalars are thread-private (no redundant work, no shuffles),
// and the direction tables come out of constant memory with a warp-uniform index.  Writes: the wind-input
// linearisation FLD (scratch, same layout as FL1), XLLWS (final), MIJ and the 1-D stress fields.
// =========================================================================================================
#ifndef KP_NTH
#define KP_NTH 128    // threads per block of k_point
#endif
#ifndef KP_MINB
#define KP_MINB 3     // resident CTAs per SM the register allocation aims at (the double-buffered row stage allows 3 at 128 threads)
#endif
#ifndef KP_UNROLL
#define KP_UNROLL 1   // the direction loop stays rolled: the frequency loop of the second SINFLX call must fit the 32 KB instruction cache
#endif
struct PointSrc {
  const double* lo;   // frequencies [0, mlo): propagation scratch or FL1 itself
  const double* hi;   // FL1
  int mlo;
  size_t kstr;        // P
  double* stage;      // shared memory: [2 stages][A][KP_NTH], this thread's column = + threadIdx.x
};
// Row pipeline: the NANG values of frequency m of this thread's grid point go HBM -> shared memory with cp.async
// (8 bytes per lane = one coalesced 256-byte row segment per warp and direction) one whole row ahead of the arithmetic,
// so that ~NANG loads per thread are in flight instead of one.
__device__ __forceinline__ void row_issue(const PointSrc& S, int m, int A) {
  const double* g = (m < S.mlo ? S.lo : S.hi) + (size_t)m * A * S.kstr;
  unsigned sa = (unsigned)__cvta_generic_to_shared(S.stage + (size_t)(m & 1) * A * KP_NTH);
  for (int k = 0; k < A; ++k) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(g));
    sa += KP_NTH * 8;
    g += S.kstr;
  }
  asm volatile("cp.async.commit_group;\n" ::);
}
template <int N>
__device__ __forceinline__ void row_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ const double* row_ptr(const PointSrc& S, int m, int A) { return S.stage + (size_t)(m & 1) * A * KP_NTH; }

template <bool ARD, int NGST, bool LLSNEG, bool STORE, bool CY>
__device__ __forceinline__ void sinput_point(const PointSrc& S, const ImplDev& d, long long p, double rnfac, double fac, double flmc,
                                             double snw, double csw, double ufric, double z0m, double raorw, double sig_n,
                                             double temp2_sw, double pturb, double aird_pvisc, double* __restrict__ fld_out,
                                             double* __restrict__ xl_out, double* sumx, double* sumy, double* sumt,
                                             double& ws_em, double& ws_fm, double& ws_last, double& phiwa_acc,
                                             double& uorbt_acc, double& aorb_acc, double* mom, bool dostore,
                                             double depth = 0.0, double jan_sds = 0.0, double xkmean = 1.0) {
  const int A = c_dc.A, F = c_dc.F;
  const double CONST1 = c_dc.BETAMAXOXKAPPA2;
  const size_t kstr = S.kstr;
  ws_em = 0.0; ws_fm = 0.0; ws_last = 0.0; phiwa_acc = 0.0; uorbt_acc = 0.0; aorb_acc = 0.0;
  constexpr bool ard = ARD;
  double cicover_pt = 0.0, cith125 = 0.0;   // SDICE3: CICV and CITH**1.25 of the point (only when LCIWA3)
  if (STORE && c_dc.licerun && c_dc.lciwa3) { cicover_pt = d.f.cicover[p]; cith125 = pow(d.f.cithick[p], 1.25); }
  const double abs_shelter = fabs(c_dc.TAUWSHELTER);
  const bool ltauwshelter = ard && abs_shelter != 0.0;
  double ustp[NGST], xstress[NGST], ystress[NGST], taux[NGST], tauy[NGST], wsin[NGST];
  if (ard) {
    if (NGST == 1) ustp[0] = ufric;
    else { ustp[0] = ufric * (1.0 + sig_n); ustp[NGST - 1] = ufric * (1.0 - sig_n); }
  } else {
    if (NGST == 1) ustp[0] = ufric;
    else { ustp[0] = ufric * (1.0 - sig_n); ustp[NGST - 1] = ufric * (1.0 + sig_n); }
  }
#pragma unroll
  for (int g = 0; g < NGST; ++g) {
    xstress[g] = 0.0; ystress[g] = 0.0;
    const double usg2 = sq(ustp[g]);
    taux[g] = usg2 * snw; tauy[g] = usg2 * csw;
    wsin[g] = 1.0 / NGST;
  }
  const double rogoroair = c_dc.G / raorw;
  const double FU = fabs(c_dc.SWELLF3), FUD = c_dc.SWELLF2;
  const double CONST3 = c_dc.idamping * (2.0 * c_dc.XKAPPA / CONST1);
  const double xkappad = 1.0 / c_dc.XKAPPA;
  const double avg = 1.0 / NGST;
  // LLNORMAGAM (sinput_ard.F90:172-176, 400-404, 436-452; sinput_jan.F90:203-207, 329-348)
  const bool normagam = CY && c_dc.llnormagam;
  const double cstrnfac = normagam ? (c_dc.DELTH / (c_dc.XKAPPA * c_dc.ZPI)) * rnfac / raorw : 0.0;
  double gamnorma[NGST];
#pragma unroll
  for (int g = 0; g < NGST; ++g) gamnorma[g] = 1.0;
  row_issue(S, 0, A);
  for (int m = 0; m < F; ++m) {
    if (m + 1 < F) { row_issue(S, m + 1, A); row_wait<1>(); } else row_wait<0>();
    const size_t o3 = idx3(d, p, m);
    const double wavnum = d.f.wavnum[o3], cinv = d.f.cinv[o3];
    const double sig = c_dc.ZPIFR[m], sig2 = sig * sig;
    const double zcn = log(wavnum * z0m);
    const double dfim = c_dc.DFIM[m], dfimofr = c_dc.DFIMOFR[m], rhowg = c_dc.RHOWG_DFIM[m];
    if (STORE) {
      if (dostore) {   // per-(point, frequency) scalars of the frequency sweep in k_stencil
        const size_t n = (size_t)d.npts, qs = (size_t)F * n;
        double* tg = d.tbg + (size_t)m * n + (size_t)p;
        const double xk = d.f.xk2cg[o3];
        tg[TQ_FACSAT * qs] = wavnum * (1.0 / c_dc.ZPI) * xk;                                        // sdissip_ard.F90:142-160
        double sbo = 0.0;
        if (m < c_dc.Fr && depth < c_dc.bathymax) sbo = (-2.0 * 0.038 * c_dc.GM1) * wavnum / sinh(dmin(2.0 * depth * wavnum, 50.0));   // sbottom.F90:76-97
        if (c_dc.licerun && c_dc.lciwa3)   // SDICE3 (sdice3.F90:131-141): TEMP = -CICV*ALP*CGROUP, added to SL/FLD like SBOTTOM's coefficient
          sbo += -cicover_pt * ((c_dc.fr45[m] * cith125) * c_dc.zalpfacx) * d.f.cgroup[o3];
        tg[TQ_SBO * qs] = sbo;
        tg[TQ_CINV * qs] = cinv;
        tg[TQ_TAIL * qs] = div_norm(div_norm(1.0, xk), wavnum);                                                       // imphftail.F90:73-81
        tg[TQ_STF * qs] = (m < c_dc.NFRE_ODD) ? d.f.stokfac[o3] * c_dc.DFIM_SIM[m] : 0.0;          // stokesdrift.F90:100-116
        double tj = 0.0;
        if (!ard) {   // SDISSIP_JAN (sdissip_jan.F90:96-132)
          const double xx = wavnum / xkmean;
          tj = jan_sds * xx * ((1.0 - c_dc.DELTA_SDIS) + c_dc.DELTA_SDIS * xx) + c_dc.rnu * c_dc.CDISVIS * sq(wavnum);
        }
        tg[TQ_JAN * qs] = tj;
      }
    }
    const double* fsrc = row_ptr(S, m, A);
    double* fo = STORE ? fld_out + (size_t)m * A * kstr : nullptr;
    double* xo = STORE ? xl_out + (size_t)m * A * kstr : nullptr;
    const bool lastm = (m == F - 1);
    double cnsn, constf = 0.0, dstab1 = 0.0, temp1 = 0.0;
    double cosu[NGST], sinu[NGST], ucn[NGST], ucnzalpd[NGST], const3_ucn2[NGST], xvd[NGST], dsd[NGST], dsc[NGST];
    if (ard) {
      cnsn = sig * CONST1 * raorw;
      constf = rogoroair * cinv * dfim;
      if (LLSNEG) {
        const double coef = -c_dc.SWELLF * 16. * sig2 / c_dc.G;
        const double coef5 = -c_dc.SWELLF5 * 2. * sqrt(2. * c_dc.rnu * sig);
        dstab1 = coef5 * aird_pvisc * wavnum;
        temp1 = coef * raorw;
      }
#pragma unroll
      for (int g = 0; g < NGST; ++g) {
        if (ltauwshelter) {
          const double taupx = taux[g] - abs_shelter * xstress[g];
          const double taupy = tauy[g] - abs_shelter * ystress[g];
          // USDIRP = ATAN2(TAUPX,TAUPY) is only used as cos(TH-USDIRP): keep the unit vector instead of the angle
          const double t2 = taupx * taupx + taupy * taupy;
          const double rt = sqrt(t2);
          ustp[g] = sqrt(rt);
          if (rt > 1e-300) { const double ri = div_norm(1.0, rt); cosu[g] = taupy * ri; sinu[g] = taupx * ri; }
          else { cosu[g] = 1.0; sinu[g] = 0.0; }
        } else { cosu[g] = csw; sinu[g] = snw; }
        ucn[g] = ustp[g] * cinv;
        ucnzalpd[g] = div_norm(c_dc.XKAPPA, ucn[g] + c_dc.ZALP);
        if (LLSNEG) {   // DSTAB = DSTAB1 + PTURB*TEMP1*(TEMP2 + (FU + FUD*COSLP)*USTP) (sinput_ard.F90:423-428) as dsd + dsc*COSLP
          const double b = pturb * temp1 * ustp[g];
          dsd[g] = fma(b, FU, fma(pturb * temp1, temp2_sw, dstab1));
          dsc[g] = b * FUD;
        }
      }
    } else {
      const double ztanhkd = sig2 / (c_dc.G * wavnum);
      cnsn = sig * CONST1 * ztanhkd * raorw;
#pragma unroll
      for (int g = 0; g < NGST; ++g) {
        ucn[g] = ustp[g] * cinv + c_dc.ZALP;
        const3_ucn2[g] = CONST3 * sq(ucn[g]);
        ucnzalpd[g] = 1.0 / ucn[g];                    // UCND
        xvd[g] = 1.0 / (-ustp[g] * xkappad * zcn * cinv);
        cosu[g] = csw; sinu[g] = snw;
      }
    }
    if (CY) {
      if (normagam) {   // GAMNORMA(IGST) of this frequency: the growth rates once more, summed over direction
        const double xngamconst = cstrnfac * d.f.xk2cg[o3];
        double sumf[NGST], sumfsin2[NGST];
#pragma unroll
        for (int g = 0; g < NGST; ++g) { sumf[g] = 0.0; sumfsin2[g] = 0.0; }
        for (int k = 0; k < A; ++k) {
          double f = dmax(fsrc[k * KP_NTH] * fac, c_dc.EPSMIN);
          const double snk = c_dc.SINTH[k], csk = c_dc.COSTH[k];
          const double cwd = csk * csw + snk * snw;
          if (lastm) f = dmax(f, flmc * sq(dmax(0.0, cwd)));
          const double sinwdif2 = sq(snk * csw - csk * snw);
#pragma unroll
          for (int g = 0; g < NGST; ++g) {
            double gam0 = 0.0;
            if (ard) {
              const double coslp = ltauwshelter ? (csk * cosu[g] + snk * sinu[g]) : cwd;
              if (coslp > 0.01) {
                const double x = coslp * ucn[g];
                const double zlog = zcn + div_norm(ucnzalpd[g], coslp);
                if (zlog < 0.0) { const double zlog2x = zlog * zlog * x; gam0 = exp_le0(zlog) * zlog2x * zlog2x * cnsn; }
              }
            } else if (cwd > 0.01) {
              const double zlog = zcn + div_norm(c_dc.XKAPPA, cwd) * ucnzalpd[g];
              if (zlog < 0.0) { const double x = cwd * ucn[g]; const double zlog2x = zlog * zlog * x; gam0 = zlog2x * zlog2x * exp_le0(zlog) * cnsn; }
            }
            sumf[g] = sumf[g] + gam0 * f;
            sumfsin2[g] = sumfsin2[g] + gam0 * f * sinwdif2;
          }
        }
#pragma unroll
        for (int g = 0; g < NGST; ++g) {
          const double znz = xngamconst * (1.0 / dmax(ustp[g], c_dc.EPSUS));
          gamnorma[g] = (1.0 + znz * sumfsin2[g]) / (1.0 + znz * sumf[g]);
        }
      }
    }
    double sx[NGST], sy[NGST], st = 0.0, tsum = 0.0, traw = 0.0, rawt = 0.0;
#pragma unroll
    for (int g = 0; g < NGST; ++g) { sx[g] = 0.0; sy[g] = 0.0; }
    constexpr int kUnroll = KP_UNROLL;
#pragma unroll kUnroll
    for (int k = 0; k < A; ++k) {
      const double fraw = fsrc[k * KP_NTH];
      if (!STORE) rawt += fraw;                                                  // SEMEAN of the incoming spectrum (sdepthlim.F90:60-70)
      double f = dmax(fraw * fac, c_dc.EPSMIN);                                  // SDEPTHLIM applied on the fly
      const double snk = c_dc.SINTH[k], csk = c_dc.COSTH[k];
      const double cwd = csk * csw + snk * snw;                                  // COSWDIF(K)
      traw += f;                                                                 // FKMEAN sees the spectrum before the floor
      if (lastm) f = dmax(f, flmc * sq(dmax(0.0, cwd)));                         // sinflx.F90:126-129
      tsum += f;
      double slp_avg = 0.0, flp_avg = 0.0, ufac2 = 0.0;
      bool xll = false;
#pragma unroll
      for (int g = 0; g < NGST; ++g) {
        double gam0 = 0.0;
        if (ard) {
          const double coslp = ltauwshelter ? (csk * cosu[g] + snk * sinu[g]) : cwd;
          if (coslp > 0.01) {
            const double x = coslp * ucn[g];
            const double zlog = zcn + div_fast(ucnzalpd[g], coslp);
            if (zlog < 0.0) {
              const double zlog2x = zlog * zlog * x;
              gam0 = exp_le0(zlog) * zlog2x * zlog2x * cnsn;
              xll = true;
            }
          }
          double dstab = 0.0;
          if (LLSNEG) dstab = fma(dsc[g], coslp, dsd[g]);
          if (CY) gam0 = gam0 * gamnorma[g];
          const double slp = gam0 * f;
          sx[g] += slp * snk; sy[g] += slp * csk;
          slp_avg += slp; flp_avg += gam0 + dstab;
        } else {
          if (cwd > 0.01) {
            const double zlog = zcn + div_norm(c_dc.XKAPPA, cwd) * ucnzalpd[g];
            if (zlog < 0.0) {
              const double x = cwd * ucn[g];
              const double zlog2x = zlog * zlog * x;
              gam0 = zlog2x * zlog2x * exp_le0(zlog) * cnsn;
              xll = true;
            }
          }
          if (CY) slp_avg += wsin[g] * gam0 * gamnorma[g];
          else slp_avg += wsin[g] * gam0;              // UFAC1
          if (LLSNEG) ufac2 += wsin[g] * (const3_ucn2[g] * (cwd - xvd[g]));
        }
      }
      double spos, fldv;
      if (ard) { spos = avg * slp_avg; fldv = avg * flp_avg; }
      else { fldv = slp_avg + ufac2 * cnsn; spos = slp_avg * f; sx[0] += spos * snk; sy[0] += spos * csk; }
      const double slv = fldv * f;
      st += spos;
      if (STORE) {
        if (dostore) {
          *fo = fldv;
          *xo = xll ? 1.0 : 0.0;
        }
        fo += kstr; xo += kstr;
        phiwa_acc += (slv - spos) * rhowg;
      }
      if (xll) { ws_em += dfim * f; ws_fm += dfimofr * f; if (lastm) ws_last += f; }
    }
    // moments of the (depth-limited, floored) spectrum
    if (!STORE) {
      const double sqk = sqrt(wavnum);
      mom[0] += dfim * traw; mom[1] += dfimofr * traw; mom[2] += c_dc.DFIMFR[m] * traw;
      mom[3] += (dfim / sqk) * traw; mom[4] += (sqk * dfim) * traw; mom[5] = traw;
      mom[6] += dfim * rawt; mom[7] = rawt;
      uorbt_acc += dfim * sig2 * tsum; aorb_acc += dfim * tsum;
    }
    double sxa = 0.0, sya = 0.0;
    if (ard) {
#pragma unroll
      for (int g = 0; g < NGST; ++g) {
        xstress[g] += constf * sx[g];
        ystress[g] += constf * sy[g];
        sxa += sx[g]; sya += sy[g];
      }
      sxa *= avg; sya *= avg;
    } else { sxa = sx[0]; sya = sy[0]; }
    sumx[m] = sxa; sumy[m] = sya; sumt[m] = st;
  }
}

// PH = 1: SDEPTHLIM, first SINFLX call (AIRSEA, SINPUT with NGST=1, FKMEAN, FEMEANWS, FRCUTINDEX, STRESSO), SDIWBK's Q
// PH = 2: second SINFLX call (AIRSEA, WSIGSTAR, SINPUT with NGST=2 and swell damping -> FLD, XLLWS; FRCUTINDEX; STRESSO with PHIWA)
// Two kernels instead of one: each frequency loop then fits the 32 KB instruction cache (no_instruction stalls of the fused
// kernel were as frequent as its dependency stalls), and the few scalars in between travel through the scratch slots.
// CY: the cy49r1 physics instance (LLGCBZ0 and/or LLNORMAGAM, read at run time inside it); the CY = false instances do not
// contain any of it.
#ifdef KP_MAXREG2   // experiment: register cap of the second-call instance
#define KP_BOUNDS(PH) __maxnreg__((PH) == 2 ? KP_MAXREG2 : 168)
#else
#define KP_BOUNDS(PH) __launch_bounds__(KP_NTH, KP_MINB)
#endif
// Z0WAVE + the logarithmic profile (z0wave.F90:68-93, airsea.F90:102-120): AIRSEA of the first SINFLX call when the forcing is the
// friction velocity (ICODE_WND = 1, 2); US stays as it is, U10 (WSWAVE) is derived and stored for the second call's TAUT_Z0.
__device__ __forceinline__ double z0wave_u10(double us, double tauw, double utop, double& z0, double& z0b, double& chrnck) {
  const double alphaog = (c_dc.llcapchnk ? chnkmin(utop) : c_dc.ALPHA) * c_dc.GM1;
  const double ust2 = us * us, ust3 = us * us * us;
  const double arg = dmax(ust2 - tauw, c_dc.EPS1);
  z0 = alphaog * ust3 / sqrt(arg);
  z0b = alphaog * ust2;
  chrnck = c_dc.G * z0 / ust2;
  return dmax((1.0 / c_dc.XKAPPA) * us * (log(c_dc.XNLEV) - log(z0)), c_dc.wspmin);
}

// USF (only with PH = 1): the ICODE_WND = 1, 2 instance -- AIRSEA of the first SINFLX call is Z0WAVE instead of TAUT_Z0
template <bool ARD, int PH, bool CY, bool USF = false>
__global__ void KP_BOUNDS(PH) k_point(ImplDev d, long long p0, long long np) {
  extern __shared__ double smem[];
  long long p = p0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = p < p0 + np;
  if (!valid) p = p0 + np - 1;      // tail threads shadow the last point (they keep the cp.async pipeline uniform) and never store
  const int A = c_dc.A, F = c_dc.F;
  const long long n = d.npts;
  double* s = d.scr;
  const long long c = p / d.P;
  const int i = (int)(p - c * d.P);
  PointSrc S;
  S.stage = smem + threadIdx.x;
  S.kstr = (size_t)d.P;
  S.hi = d.f.fl1 + (size_t)i + (size_t)d.P * A * F * (size_t)c;
  if (d.lo_on) {
    // propagation scratch (P,A,Fr,C): padded lanes of the last chunk take the chunk's first point (propag_wam.F90:388-398)
    const int il = (p < d.nloc) ? i : 0;
    S.lo = d.fl_lo + (size_t)il + (size_t)d.P * A * d.lo_F * (size_t)c;
    S.mlo = d.Fr;
  } else { S.lo = S.hi; S.mlo = 0; }
  const double aird = d.f.aird[p], wdwave = d.f.wdwave[p], cicover = d.f.cicover[p], wswave = d.f.wswave[p];
  const double raorw = dmax(aird, 1.0) * c_dc.ROWATERM1;
  double snw, csw;
  sincos(wdwave, &snw, &csw);
  const double flmc = (1. - 0.9 * dmin(cicover, 0.99)) * c_dc.flmin;      // FLM(K) = flmc*max(0,COSWDIF)**2
  const double DELT25 = c_dc.WETAIL * c_dc.FR[F - 1] * c_dc.DELTH;
  double fac = 1.0;
  double ustar = d.f.ufric[p], z0 = 0.0, z0b = 0.0, ch = 0.0;
  double tauw, tauwdir;
  double sumx[EW_MAXF], sumy[EW_MAXF], sumt[EW_MAXF];
  double ws_em, ws_fm, ws_last, phiwa_acc, uorbt_acc, aorb_acc;
  double emean, fmean, f1mean, xkmean;
  // RNFAC (sinflx.F90:117-121) and 1/2 ALPHAP of the Phillips tail (HALPHAP)
  double rnfac = 1.0, halp = 0.0;
  if (CY) { if (c_dc.llnormagam && c_dc.llcapchnk) rnfac = 1.0 + c_dc.DTHRN_A * (1.0 + tanh(wswave - c_dc.DTHRN_U)); }
  bool fac_known = false;
  if (PH == 1) {
    // ---- SINFLX call 1: AIRSEA (IUSFG=0), SINPUT (NGST=1), FEMEANWS, FRCUTINDEX, STRESSO
    tauw = d.f.tauw[p]; tauwdir = d.f.tauwdir[p];
    if (CY) {
      if (c_dc.llgcbz0) {
        // HALPHAP (halphap.F90:68-115 with MEANSQS_LF and FEMEAN on the spectrum in the wind direction) needs the depth-limited,
        // floored spectrum before AIRSEA: its own pass over FL1, which also settles SDEPTHLIM's factor (same optimistic scheme
        // as below: factor 1 first, depth-limited lanes repeat the pass).
        const double zlnfrnfre = log(c_dc.FR[F - 1]);
        for (int attempt = 0; attempt < 2; ++attempt) {
          double xmss = 0.0, em = 0.0, fm = 0.0, temp2 = 0.0, f1d = 0.0, m6 = 0.0, rawt = 0.0;
          row_issue(S, 0, A);
          for (int m = 0; m < F; ++m) {
            if (m + 1 < F) { row_issue(S, m + 1, A); row_wait<1>(); } else row_wait<0>();
            const double* fsrc = row_ptr(S, m, A);
            const double wavnum = d.f.wavnum[idx3(d, p, m)];
            double sflwd = 0.0;
            temp2 = 0.0; rawt = 0.0;
            for (int k = 0; k < A; ++k) {
              const double fraw = fsrc[k * KP_NTH];
              rawt += fraw;
              double f = dmax(fraw * fac, c_dc.EPSMIN);
              const double cwd = c_dc.COSTH[k] * csw + c_dc.SINTH[k] * snw;
              if (m == F - 1) f = dmax(f, flmc * sq(dmax(0.0, cwd)));
              const double flwd = signbit(cwd) ? f * 0.0 : f;          // FL1 * (0.5 + 0.5*SIGN(1,COSWDIF))
              sflwd = sflwd + flwd;
              temp2 = temp2 + dmax(flwd, c_dc.EPSMIN);
              if (m == F - 1) f1d = f1d + flwd * c_dc.DELTH;
            }
            xmss = xmss + (c_dc.DFIM[m] * sq(wavnum)) * sflwd;
            em = em + temp2 * c_dc.DFIM[m];
            fm = fm + c_dc.DFIMOFR[m] * temp2;
            m6 += c_dc.DFIM[m] * rawt;
          }
          em = em + DELT25 * temp2;
          fm = fm + (c_dc.FRTAIL * c_dc.DELTH) * temp2;
          fm = dmax(em / fm, c_dc.FR[0]);
          double alphap;
          bool tail = true;
          if (em > 0.0 && fm < c_dc.FR[F - 3]) {
            alphap = xmss / (zlnfrnfre - log(fm));
            tail = alphap > c_dc.ALPHAPMAX;
          }
          if (tail) alphap = c_dc.ZPI4GM2 * c_dc.FR5[F - 1] * f1d;
          halp = 0.5 * dmin(alphap, c_dc.ALPHAPMAX);
          if (attempt == 1 || !c_dc.lbiwbk) break;
          const double emr = (c_dc.EPSMIN + m6) + DELT25 * rawt;
          fac = dmin(d.f.emaxdpt[p] / emr, 1.0);
          if (fac == 1.0) break;
        }
        fac_known = true;
        if (valid) s[S_HALP * n + p] = halp;
        if (USF) { const double u10 = z0wave_u10(ustar, tauw, wswave, z0, z0b, ch); if (valid) d.f.wswave[p] = u10; }
        else taut_z0_gc(d.gc, 0, halp, wswave, wdwave, tauw, tauwdir, rnfac, ustar, z0, z0b, ch);
      } else if (USF) { const double u10 = z0wave_u10(ustar, tauw, wswave, z0, z0b, ch); if (valid) d.f.wswave[p] = u10; }
      else taut_z0(0, wswave, wdwave, tauw, tauwdir, ustar, z0, z0b, ch);
    } else if (USF) { const double u10 = z0wave_u10(ustar, tauw, wswave, z0, z0b, ch); if (valid) d.f.wswave[p] = u10; }
    else taut_z0(0, wswave, wdwave, tauw, tauwdir, ustar, z0, z0b, ch);
    // SDEPTHLIM (sdepthlim.F90:50-82) needs the total energy of the incoming spectrum before anything else can be formed.
    // Instead of a separate pass over FL1, the first SINPUT pass is run with the limiting factor 1 (exact wherever the
    // spectrum is not depth-limited, i.e. almost everywhere) and sums that energy on the side; only the lanes that turn
    // out to be depth-limited repeat the pass with their factor.
    double mom[8];
    for (int attempt = 0; attempt < 2; ++attempt) {
#pragma unroll
      for (int x = 0; x < 8; ++x) mom[x] = 0.0;
      sinput_point<ARD, 1, false, false, CY>(S, d, p, rnfac, fac, flmc, snw, csw, ustar, z0, raorw, 0.0, 0.0, 0.0, 0.0, nullptr, nullptr,
                                             sumx, sumy, sumt, ws_em, ws_fm, ws_last, phiwa_acc, uorbt_acc, aorb_acc, mom, false);
      if (CY) { if (fac_known) break; }
      if (attempt == 1 || !c_dc.lbiwbk) break;
      const double em = (c_dc.EPSMIN + mom[6]) + DELT25 * mom[7];
      fac = dmin(d.f.emaxdpt[p] / em, 1.0);
      if (fac == 1.0) break;
    }
    // FKMEAN (fkmean.F90:60-154)
    const double COEFM1 = c_dc.FRTAIL * c_dc.DELTH;
    const double COEF1 = c_dc.WP1TAIL * c_dc.DELTH * sq(c_dc.FR[F - 1]);
    const double COEFA = COEFM1 * sqrt(c_dc.G) / c_dc.ZPI;
    const double COEFX = COEF1 * (c_dc.ZPI / sqrt(c_dc.G));
    emean = c_dc.EPSMIN + mom[0] + DELT25 * mom[5];
    fmean = emean / (c_dc.EPSMIN + mom[1] + COEFM1 * mom[5]);
    f1mean = (c_dc.EPSMIN + mom[2] + COEF1 * mom[5]) / emean;
    const double akmean = sq(emean / (c_dc.EPSMIN + mom[3] + COEFA * mom[5]));
    xkmean = sq((c_dc.EPSMIN + mom[4] + COEFX * mom[5]) / emean);
    if (valid) {
      s[S_EMEAN * n + p] = emean; s[S_FMEAN * n + p] = fmean; s[S_F1MEAN * n + p] = f1mean;
      s[S_AKMEAN * n + p] = akmean; s[S_XKMEAN * n + p] = xkmean;
      s[S_FAC * n + p] = fac;
      s[S_UORBT * n + p] = uorbt_acc; s[S_AORB * n + p] = aorb_acc;
    }
  } else {
    fac = s[S_FAC * n + p];
    emean = s[S_EMEAN * n + p]; fmean = s[S_FMEAN * n + p]; f1mean = s[S_F1MEAN * n + p]; xkmean = s[S_XKMEAN * n + p];
    ustar = s[S_USTAR1 * n + p]; tauw = s[S_TAUW1 * n + p]; tauwdir = s[S_TWDIR1 * n + p];
    uorbt_acc = s[S_UORBT * n + p]; aorb_acc = s[S_AORB * n + p];
    if (CY) halp = s[S_HALP * n + p];
  }

  auto frcut = [&](double fmeanws, double ust) -> int {   // frcutindex.F90:84-97
    if (cicover > c_dc.cithrsh_tail) return F;
    const double fpmh = c_dc.TAILFACTOR / c_dc.FR[0];
    const double fppm = c_dc.TAILFACTOR_PM * c_dc.G / (28.0 * c_dc.ZPIFR[0]);
    const double fm2 = dmax(fmeanws, fmean) * fpmh;
    const double fpm = fppm / dmax(ust, c_dc.EPSMIN);
    const double fpm4 = dmax(fm2, fpm);
    int mij = (int)lround(log10(fpm4) * c_dc.FLOGSPRDM1) + 1;
    return min(max(1, mij), F);
  };
  // STRESSO (stresso.F90:120-233) from the per-frequency SPOS moments; returns TAUW, TAUWDIR, PHIWA
  auto stresso = [&](int mij, double ust_in, double z0m, bool llphiwa, double phiwa_lf, double& tw, double& twd, double& phiwa) {
    double xs = 0.0, ys = 0.0, pw = phiwa_lf;
    for (int m = 0; m < mij; ++m) {
      const double r = rhowgdfth(m, mij);
      const double cm = r * d.f.cinv[idx3(d, p, m)];
      xs += cm * sumx[m]; ys += cm * sumy[m]; pw += r * sumt[m];
    }
    const double am = dmax(aird, 1.0);
    xs = xs / am; ys = ys / am;
    // directional moments of the spectrum at the cut-off frequency (tau_phi_hf.F90:150-178)
    double f3 = 0.0, f2 = 0.0, f1d = 0.0, f1dsin2 = 0.0;
    {
      const int m = mij - 1;
      const double* fsrc = (m < S.mlo ? S.lo : S.hi) + (size_t)m * A * S.kstr;
      for (int k = 0; k < A; ++k) {
        double f = dmax(__ldg(fsrc + (size_t)k * S.kstr) * fac, c_dc.EPSMIN);
        const double cwd = c_dc.COSTH[k] * csw + c_dc.SINTH[k] * snw;
        if (m == F - 1) f = dmax(f, flmc * sq(dmax(0.0, cwd)));
        const double cw = dmax(cwd, 0.0);
        const double fc2 = f * cw * cw;
        f3 += fc2 * cw; f2 += fc2;
        if (CY) { f1d += f; f1dsin2 += f * sq(c_dc.SINTH[k] * csw - c_dc.COSTH[k] * snw); }
      }
      f3 *= c_dc.DELTH; f2 *= c_dc.DELTH;
      if (CY) { f1d *= c_dc.DELTH; f1dsin2 *= c_dc.DELTH; }
    }
    bool shelter;
    double usdirp_s, usdirp_c, ust;
    if (!ARD || c_dc.TAUWSHELTER == 0.0) { shelter = false; usdirp_s = snw; usdirp_c = csw; ust = ust_in; }
    else {
      shelter = true;
      const double taupx = sq(ust_in) * snw - c_dc.TAUWSHELTER * xs, taupy = sq(ust_in) * csw - c_dc.TAUWSHELTER * ys;
      const double rt = sqrt(taupx * taupx + taupy * taupy);
      ust = sqrt(rt);
      if (rt > 0.0) { usdirp_s = taupx / rt; usdirp_c = taupy / rt; } else { usdirp_s = 0.0; usdirp_c = 1.0; }
    }
    double tauhf, phihf;
    if (CY) {
      const double confg0 = c_dc.llnormagam ? c_dc.GAMNCONST * c_dc.FR5[mij - 1] * rnfac : 0.0;
      const double oms = c_dc.llgcbz0 ? __ldg(d.gc + GC_OMEGA * c_dc.NWAV_GC + ns_gc(ust_in) - 1) : 0.0;   // omegagc.F90:51-55
      tau_phi_hf<true>(mij, shelter, z0m, aird, f3, f2, ust, tauhf, phihf, llphiwa, confg0, f1dsin2, f1d, oms);
    } else tau_phi_hf<false>(mij, shelter, z0m, aird, f3, f2, ust, tauhf, phihf, llphiwa);
    xs = xs + tauhf * usdirp_s;
    ys = ys + tauhf * usdirp_c;
    tw = dmax(sqrt(sq(xs) + sq(ys)), 0.0);
    twd = atan2(xs, ys);
    if (CY) { if (!c_dc.llgcbz0) tw = dmin(tw, sq(ust_in) * (1.0 / (1.0 + c_dc.EPS1))); }   // stresso.F90:218-223
    else tw = dmin(tw, sq(ust_in) * (1.0 / (1.0 + c_dc.EPS1)));
    phiwa = llphiwa ? pw + phihf : 0.0;
  };
  if (PH == 1) {
    const double emeanws = c_dc.EPSMIN + ws_em + DELT25 * ws_last;
    const double fmeanws = emeanws / (c_dc.EPSMIN + ws_fm + c_dc.FRTAIL * c_dc.DELTH * ws_last);
    const int mij1 = frcut(fmeanws, ustar);
    double ph;
    stresso(mij1, ustar, z0, false, 0.0, tauw, tauwdir, ph);
    // ---- SDIWBK (sdiwbk.F90:69-104)
    double sds = 0.0;
    if (c_dc.lbiwbk && d.f.depth[p] < 50.0) {
      const double alph = 2.0 * d.f.emaxdpt[p] / emean;
      const double arg = dmin(alph, 50.0);
      double q_old = exp(-arg), q = q_old;
      for (int ic = 1; ic <= 15; ++ic) {
        const double expq = exp(-arg * (1.0 - q_old));
        q = q_old - (expq - q_old) / (arg * expq - 1.0);
        const double rel_err = fabs(q - q_old) / q_old;
        if (rel_err < 0.00001) break;
        q_old = q;
      }
      q = dmin(q, 1.0);
      sds = 2.0 * alph * q * f1mean;
    }
    if (valid) { s[S_USTAR1 * n + p] = ustar; s[S_TAUW1 * n + p] = tauw; s[S_TWDIR1 * n + p] = tauwdir; s[S_SDS * n + p] = sds; }
    return;
  }
  // ---- SINFLX call 2: AIRSEA (IUSFG=1), SINPUT (NGST=2, LLSNEG), FEMEANWS, FRCUTINDEX, STRESSO (LLPHIWA)
  if (CY) {
    if (c_dc.llgcbz0) taut_z0_gc(d.gc, 1, halp, wswave, wdwave, tauw, tauwdir, rnfac, ustar, z0, z0b, ch);
    else taut_z0(1, wswave, wdwave, tauw, tauwdir, ustar, z0, z0b, ch);
  } else taut_z0(1, wswave, wdwave, tauw, tauwdir, ustar, z0, z0b, ch);
  if (valid) { d.f.ufric[p] = ustar; d.f.z0m[p] = z0; d.f.z0b[p] = z0b; d.f.chrnck[p] = ch; }
  const double sig_n = CY ? wsigstar_gc(wswave, ustar, z0, d.f.wstar[p]) : wsigstar(ustar, z0, d.f.wstar[p]);
  double temp2_sw = 0.0, pturb = 0.0, aird_pvisc = 0.0;
  if (ARD) {   // sinput_ard.F90:179-271
    const double uorbt = 2.0 * sqrt(c_dc.EPSMIN + uorbt_acc), aorb = 2.0 * sqrt(c_dc.EPSMIN + aorb_acc);
    const double re = (4.0 / c_dc.rnu) * uorbt * aorb;
    const double z0vis = c_dc.rnum / dmax(ustar, 0.0001);
    const double z0tub = c_dc.Z0RAT * dmin(c_dc.Z0TUBMAX, z0);
    const double zorb = aorb / dmax(z0vis, z0tub);
    const double delabm1 = (double)c_dc.IAB / (c_dc.ABMAX - c_dc.ABMIN);
    const double xi = (log10(dmax(zorb, 3.0)) - c_dc.ABMIN) * delabm1;
    const int ind = min(c_dc.IAB - 1, (int)xi);
    const double deli1 = dmin(1.0, xi - (double)ind), deli2 = 1.0 - deli1;
    const double fww = __ldg(d.tab.swellft + ind - 1) * deli2 + __ldg(d.tab.swellft + ind) * deli1;
    temp2_sw = fww * uorbt;
    const double re_c = (c_dc.SWELLF6 == 1.0) ? c_dc.SWELLF4 : c_dc.SWELLF4 * pow(2.0 / aorb, 1.0 - c_dc.SWELLF6);
    double pvisc;
    if (c_dc.SWELLF7 > 0.0) { const double sm = 0.5 * tanh((re - re_c) * c_dc.SWELLF7M1); pturb = 0.5 + sm; pvisc = 0.5 - sm; }
    else if (re <= re_c) { pturb = 0.0; pvisc = 0.5; }
    else { pturb = 0.5; pvisc = 0.0; }
    aird_pvisc = pvisc * raorw;
  }
  double* fld_out = d.fldin + (size_t)i + (size_t)d.P * A * F * (size_t)c;
  double* xl_out = d.f.xllws + (size_t)i + (size_t)d.P * A * F * (size_t)c;
  double dum[8];
  sinput_point<ARD, 2, true, true, CY>(S, d, p, rnfac, fac, flmc, snw, csw, ustar, z0, raorw, sig_n, temp2_sw, pturb, aird_pvisc, fld_out, xl_out,
                              sumx, sumy, sumt, ws_em, ws_fm, ws_last, phiwa_acc, uorbt_acc, aorb_acc, dum, valid, d.f.depth[p],
                              c_dc.CDIS * c_dc.ZPI * f1mean * sq(emean) * p4(xkmean), xkmean);
  const double emeanws = c_dc.EPSMIN + ws_em + DELT25 * ws_last;
  const double fmeanws = emeanws / (c_dc.EPSMIN + ws_fm + c_dc.FRTAIL * c_dc.DELTH * ws_last);
  const int mij = frcut(fmeanws, ustar);
  double phiwa;
  stresso(mij, ustar, z0, true, phiwa_acc, tauw, tauwdir, phiwa);
  if (valid) {
    d.f.tauw[p] = tauw; d.f.tauwdir[p] = tauwdir; d.f.mij[p] = mij;
    s[S_PHIWA * n + p] = phiwa;
    s[S_MIJ * n + p] = (double)mij;
    s[S_USFM * n + p] = ustar * dmax(fmeanws, fmean);
  }
}


// =========================================================================================================
// k_stencil: CTA = 8 consecutive grid points x NANG directions; thread = (direction k, group of NP consecutive points).
// SDISSIP, SNONLIN, SDIWBK, SBOTTOM, the implicit update, WNFLUXES, IMPHFTAIL, SETICE and STOKESDRIFT as ONE sweep
// over frequency with ONE barrier per step.  The DIA quadruplets of "centre" frequency MC only touch the frequencies
// MC-4 .. MC+3 (nlweigt.F90: IKM=MC-4, IKM1=MC-3, IKP=MC+2, IKP1=MC+3 for FRATIO=1.1), so the sweep keeps
//   * a 9-row ring of the spectrum in shared memory, [row][direction + halo][point]: rows stream in from HBM one per
//     step (already depth-limited and floored); the cyclic direction halo turns every partner / window access into
//     "own address + uniform offset" (K1W..K21W are cyclic shifts, jafu.F90) -> no per-load address arithmetic,
//     and with NP=2 every shared-memory access is one 16-byte LDS/STS serving two grid points,
//   * the 8 pending rows of the SNONLIN sums SL/FLD of the thread's own bins in registers,
// and, at step MC, finishes row MC-4: it receives its last DIA contribution, gets the dissipation, breaking and
// bottom terms, is advanced in time and goes straight from registers to HBM (the thread that computes a bin also
// loads and stores it: 8 points x 8 B = 64-byte segments).  The reference's scatter of each quadruplet into 9 bins
// (snonlin.F90:253-308) is a gather through the inverse shifts: no atomics, fixed summation order.
// Hazards: row s+4 is written in phase A of step s into the slot of row s-5 (last read in step s-1, phase A);
// the interaction planes, the per-warp saturation maxima and BTH0 are double-buffered by the parity of s.
// =========================================================================================================
#define ST_NPT 8
#define ST_RING 9
#define ST_NSMAX 17   // 2*NSDSNTH+1 <= 17 (NANG <= 36, init_sdiss_ardh.F90:72)
enum { PC_FAC = 0, PC_ENH, PC_USFMDELT, PC_SDSBK, PC_RTAIL, PC_FLMC, PC_ICEADD, PC_ICEFREE, PC_SNW, PC_CSW, PC_BETA, PC_N };

template <int NP> struct Vd;
template <> struct Vd<1> { double v[1]; };
template <> struct __align__(16) Vd<2> { double v[2]; };
template <int NP> __device__ __forceinline__ Vd<NP> lds(const char* sm, unsigned off) { return *reinterpret_cast<const Vd<NP>*>(sm + off); }
template <int NP> __device__ __forceinline__ void sts(char* sm, unsigned off, const Vd<NP>& x) { *reinterpret_cast<Vd<NP>*>(sm + off) = x; }
template <int NP> __device__ __forceinline__ Vd<NP> ldg(const double* p) { return __ldg(reinterpret_cast<const Vd<NP>*>(p)); }
template <> __device__ __forceinline__ Vd<1> ldg<1>(const double* p) { Vd<1> r; r.v[0] = __ldg(p); return r; }
template <> __device__ __forceinline__ Vd<2> ldg<2>(const double* p) {
  const double2 t = __ldg(reinterpret_cast<const double2*>(p));
  Vd<2> r; r.v[0] = t.x; r.v[1] = t.y; return r;
}
template <int NP> __device__ __forceinline__ void stg(double* p, const Vd<NP>& x) { *reinterpret_cast<Vd<NP>*>(p) = x; }
__device__ __forceinline__ unsigned slot9(int r) { return (unsigned)c_dc.SLOT9[r]; }   // r % ST_RING, 0 <= r < EW_MAXF

struct StencilSmem {   // byte offsets into dynamic shared memory
  unsigned ring, cur, tbs, pc, part, bth0, stage, bsl, mbar, total, RSB, PSB, SSB;
};
__host__ __device__ constexpr StencilSmem stencil_smem(int A, int halo_r, int halo_c, int nth, int np, bool lwflux) {
  StencilSmem s{};
  const int nwarp = nth / 32;
  s.RSB = (unsigned)(A + 2 * halo_r) * ST_NPT * 8;
  s.PSB = (unsigned)(A + 2 * halo_c) * ST_NPT * 8;
  s.SSB = (unsigned)nth * np * 8;
  unsigned o = 0;
  s.ring = o; o += ST_RING * s.RSB;
  s.cur = o; o += 2 * 6 * s.PSB;
  s.tbs = o; o += 8 * TQ_N * ST_NPT * 8;
  s.pc = o; o += PC_N * ST_NPT * 8;
  s.part = o; o += 3u * (unsigned)nwarp * ST_NPT * 8;
  s.bth0 = o; o += 2 * ST_NPT * 8;
  s.stage = o; o += (lwflux ? 3 : 2) * s.SSB;      // thread-private landing slots of the cp.async row loads (FL1 row, wind-input row, XLLWS row)
  s.bsl = o; o += 3 * s.SSB;        // thread-private saturation values of the three rows between their window sum and their finish
  s.mbar = o; o += 16;
  s.total = o;
  return s;
}
template <int NB> __device__ __forceinline__ void cp_async(unsigned sa, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(sa), "l"(g), "n"(NB));
}
template <int N> struct IC { static constexpr int value = N; };
// split CTA barrier (arrive early, wait late) on a shared-memory mbarrier; cp.async completion is part of the phase
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_cp_async(unsigned bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile("{\n .reg .pred p;\n WAIT_LOOP:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra WAIT_DONE;\n bra WAIT_LOOP;\n WAIT_DONE:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
#ifndef SW_SLEEP
#define SW_SLEEP 64   // ns between polls of a role that waits for its partner
#endif
// the same with a back-off: a role that runs ahead of the other one must not spend the issue slots its partner needs
__device__ __forceinline__ void mbar_wait_backoff(unsigned bar, unsigned parity) {
  unsigned done;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  while (!done) {
#if SW_SLEEP > 0
    __nanosleep(SW_SLEEP);
#endif
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
// Direction geometry of the standard ecWAM grids (NANG = 12, 24, 36), known at compile time so that every shared-memory
// offset of k_stencil is an immediate: NSDSNTH = min(nint(80 deg / DELTH), NANG/2-1) (init_sdiss_ardh.F90:72) and the DIA
// partner shifts K1W, K11W, K2W, K21W(K,KH) - K of nlweigt.F90:108-206 (checked against the run-time tables on the host).
__host__ __device__ constexpr int geo_nsd(int A) { return A == 36 ? 8 : (A == 24 ? 5 : (A == 12 ? 3 : 0)); }
__host__ __device__ constexpr int geo_sh(int A, int kh, int q) {
  const int a = (A == 36) ? 1 : 0, b = (A == 36) ? 3 : (A == 24 ? 2 : 1);
  const int mag = q == 0 ? a : (q == 1 ? a + 1 : (q == 2 ? b : b + 1));
  const int sgn = ((q < 2) == (kh == 0)) ? -1 : 1;
  return (A == 36 || A == 24 || A == 12) ? sgn * mag : 0;
}
__host__ __device__ constexpr int geo_hc(int A) { return A == 36 ? 4 : (A == 24 ? 3 : (A == 12 ? 2 : 0)); }
__host__ __device__ constexpr int stencil_threads(int A, int np) { return ((A * (ST_NPT / np) + 31) / 32) * 32; }

#ifndef ST_MAXREG
#define ST_MAXREG 168   // register cap of the two-point instance: 2 CTAs of 5 warps per SM put 3 warps on two of the four
                        // sub-partitions (16384 registers each), so 3 x 32 x 168 is the most that still fits (176 halves the occupancy)
#endif
#define ST_BOUNDS(NP) __maxnreg__(NP == 2 ? ST_MAXREG : 112)

// scalar closures of one grid point after the sweep: STOKESDRIFT, the new wind-sea mean parameters, WNFLUXES
// qs: the eight direction sums of the point (stride ST_NPT), pcv: its PC_* constants (stride ST_NPT)
template <bool LWFLUX, int PST = ST_NPT>
__device__ __forceinline__ void stencil_closure(const ImplDev& d, const long long pp, const double* qs, const double* pcv) {
    const int F = c_dc.F;
    const long long n = d.npts;
    const double* s = d.scr;
    double q[8];
#pragma unroll
    for (int x = 0; x < 8; ++x) q[x] = qs[x * PST];
    const double snw = pcv[PC_SNW * PST], csw = pcv[PC_CSW * PST];
    const double cicover = d.f.cicover[pp];
    const double ufric = d.f.ufric[pp], aird = d.f.aird[pp], wsw = d.f.wswave[pp];
    // STOKESDRIFT closure (stokesdrift.F90:118-142)
    double us = q[3], vs = q[4];
    if (c_dc.licerun && c_dc.lwamrsetci && cicover > c_dc.cithrsh) { us = 0.016 * wsw * snw * (1.0 - cicover); vs = 0.016 * wsw * csw * (1.0 - cicover); }
    d.f.ustokes[pp] = dmin(dmax(us, -1.5), 1.5); d.f.vstokes[pp] = dmin(dmax(vs, -1.5), 1.5);
    if (LWFLUX) {   // implsch.F90:435-446
      const double DELT25 = c_dc.WETAIL * c_dc.FR[F - 1] * c_dc.DELTH;
      const double em2 = c_dc.EPSMIN + q[5] + DELT25 * q[7];
      const double fm2 = em2 / (c_dc.EPSMIN + q[6] + c_dc.FRTAIL * c_dc.DELTH * q[7]);
      if (em2 < c_dc.WSEMEAN_MIN) { d.f.wsemean[pp] = c_dc.WSEMEAN_MIN; d.f.wsfmean[pp] = 2. * c_dc.FR[F - 1]; }
      else { d.f.wsemean[pp] = em2; d.f.wsfmean[pp] = fm2; }
    }
    if (c_dc.lcflx) {    // WNFLUXES closure (wnfluxes.F90:222-331, LWNEMOCOU=F)
      const double PHIOC_ICE = -3.75, PHIAW_ICE = 3.75, C1 = 1.03e-3, C2 = 0.04e-3, P1 = 1.48, P2 = -0.21, CDMAX_LOC = 0.003;
      const double epsus3 = c_dc.EPSUS * sqrt(c_dc.EPSUS);
      // wnfluxes.F90:150-158: with a sea-ice attenuation scheme the open-water weight decays over CICOVER 0..0.02
      const double cithrsh_inv = c_dc.lciwa_any ? 50.0 : 1.0 / dmax(c_dc.cithrsh, 0.01);
      const double zcithrs = c_dc.lciwa_any ? 0.0 : c_dc.ciblock, zmaxexp = c_dc.lciwa_any ? 20.0 : 10.0;
      const double phiwa = s[S_PHIWA * n + pp];
      double ooval = 1.0, ustar = ufric;
      if (c_dc.licerun && c_dc.lwamrsetci && cicover > zcithrs) {
        ooval = exp(-dmin(p4(cicover * cithrsh_inv), zmaxexp));
        const double u10p = dmax(wsw, c_dc.EPSU10);
        const double cd_bulk = dmin((C1 + C2 * pow(u10p, P1)) * pow(u10p, P2), CDMAX_LOC);
        const double cd_wave = sq(ufric / u10p);
        ustar = dmax(sqrt(ooval * cd_wave + (1.0 - ooval) * cd_bulk) * u10p, c_dc.EPSUS);
      }
      const double tau = aird * dmax(sq(ustar), c_dc.EPSUS);
      double tauxd = tau * snw, tauyd = tau * csw;
      double tauocxd = tauxd - ooval * q[1], tauocyd = tauyd - ooval * q[2];
      const double tauoc = dmin(dmax(sqrt(sq(tauocxd) + sq(tauocyd)) / tau, c_dc.TAUOCMIN), c_dc.TAUOCMAX);
      if (c_dc.lwcouast) {
        const double ua = d.f.ustra[pp], va = d.f.vstra[pp];
        if (ua != 0.0 || va != 0.0) { tauxd = ua; tauocxd = ua * tauoc; tauyd = va; tauocyd = va * tauoc; }
      }
      d.f.tauxd[pp] = tauxd; d.f.tauyd[pp] = tauyd; d.f.tauocxd[pp] = tauocxd; d.f.tauocyd[pp] = tauocyd; d.f.tauoc[pp] = tauoc;
      d.f.tauicx[pp] = 0.0; d.f.tauicy[pp] = 0.0;
      const double xn = aird * dmax(ustar * ustar * ustar, epsus3);
      double phiocd = ooval * (q[0] - phiwa) + (1.0 - ooval) * PHIOC_ICE * xn;
      const double phieps = dmin(dmax(phiocd / xn, c_dc.PHIEPSMIN), c_dc.PHIEPSMAX);
      phiocd = phieps * xn;
      d.f.phiocd[pp] = phiocd; d.f.phieps[pp] = phieps; d.f.phiaw[pp] = ooval * phiwa / xn + (1.0 - ooval) * PHIAW_ICE;
    }
}

// TA = NANG when it is one of the standard grids (compile-time geometry), 0 = run-time geometry (any NANG <= 36)
template <int TA, int NP, bool LWFLUX>
__global__ void ST_BOUNDS(NP) k_stencil(ImplDev d, long long p0, long long np, StencilSmem Lrt) {
  extern __shared__ __align__(16) char sm[];
  typedef Vd<NP> V;
  const int A = TA > 0 ? TA : c_dc.A, F = c_dc.F;
  constexpr int NG = ST_NPT / NP;                    // point groups per CTA
  const int NSD = TA > 0 ? geo_nsd(TA) : c_dc.NSDSNTH, NS = 2 * NSD + 1;
  const bool ard = c_dc.iphys == 1;
  const int H = TA > 0 ? geo_nsd(TA) : d.halo_r, HC = TA > 0 ? geo_hc(TA) : d.halo_c;
  const int nwarp = TA > 0 ? stencil_threads(TA, NP) / 32 : (int)(blockDim.x >> 5);
  constexpr StencilSmem Lct = stencil_smem(TA, geo_nsd(TA), geo_hc(TA), stencil_threads(TA, NP), NP, LWFLUX);
  const StencilSmem L = TA > 0 ? Lct : Lrt;
  const unsigned RSB = L.RSB, PSB = L.PSB;
  int dsb[2][4];
#pragma unroll
  for (int kh = 0; kh < 2; ++kh)
#pragma unroll
    for (int q = 0; q < 4; ++q) dsb[kh][q] = TA > 0 ? geo_sh(TA, kh, q) * (ST_NPT * 8) : d.dsb[kh][q];
  const unsigned smb = (unsigned)__cvta_generic_to_shared(sm);
  const int t = threadIdx.x;
  int k = t / NG;
  const int j = t - k * NG;
  const bool act = k < A;                            // threads beyond NANG*NG shadow the last direction and never store
  if (!act) k = A - 1;
  const unsigned jo = (unsigned)(j * NP * 8);
  const unsigned me_r = (unsigned)(k + H) * (ST_NPT * 8) + jo;     // own bin inside a ring row
  const unsigned me_c = (unsigned)(k + HC) * (ST_NPT * 8) + jo;    // own bin inside an interaction plane
  const unsigned me_s = L.stage + (unsigned)t * (NP * 8);          // own landing slot
  // ---- points
  const long long pbase = p0 + (long long)blockIdx.x * ST_NPT;
  const long long plast = p0 + np - 1;
  long long pq = pbase + (long long)j * NP;          // first point of the group
  const bool pvalid = pq + (NP - 1) <= plast;
  if (!pvalid) pq = plast - (NP - 1);
  const bool dost = act && pvalid;
  const long long n = d.npts;
  const double* s = d.scr;
  const size_t P = (size_t)d.P;
  const size_t rstr = P * A;                         // row (frequency) stride of the chunked layout
  const long long pc_ = pq / d.P;
  const int pi_ = (int)(pq - pc_ * d.P);
  const size_t off_hi = (size_t)pi_ + P * A * F * (size_t)pc_ + P * (size_t)k;    // element (pi, k, 0, pc) of a (P,A,F,C) array
  const double* src_hi = d.f.fl1 + off_hi;
  const double* src_lo = src_hi;
  int mlo = 0;
  if (d.lo_on) {   // propagated frequencies come from the propagation scratch (P,A,Fr,C); its padded lanes are filled (launch_pad)
    src_lo = d.fl_lo + (size_t)pi_ + P * A * d.lo_F * (size_t)pc_ + P * (size_t)k;
    mlo = d.Fr;
  }
  double* dst = d.f.fl1 + off_hi;
  const double* src_in = d.fldin + off_hi;
  const double* src_xl = d.f.xllws + off_hi;

  // ---- prologue: per-point constants, barrier
  if (t == 0) mbar_init(smb + L.mbar, 2u * blockDim.x);   // per phase and thread: one arrive + one arrive on completion of its cp.async
  if (t < ST_NPT) {
    const long long qp = min(pbase + t, plast);
    double* pcv = reinterpret_cast<double*>(sm + L.pc) + t;
    const double wd = d.f.wdwave[qp], ci = d.f.cicover[qp], dep = d.f.depth[qp];
    double snw, csw;
    sincos(wd, &snw, &csw);
    double enhfr = dmax(0.75 * dep * s[S_AKMEAN * n + qp], 0.5);
    enhfr = 1.0 + (5.5 / enhfr) * (1.0 - .833 * enhfr) * exp(-1.25 * enhfr);
    const int mijq = (int)s[S_MIJ * n + qp];
    const bool seticeq = c_dc.licerun && c_dc.lmaskice && ci > c_dc.cithrsh;
    pcv[PC_FAC * ST_NPT] = s[S_FAC * n + qp];
    pcv[PC_ENH * ST_NPT] = enhfr;
    pcv[PC_USFMDELT * ST_NPT] = s[S_USFM * n + qp] * c_dc.delt;
    pcv[PC_SDSBK * ST_NPT] = (c_dc.lbiwbk && dep < 50.0) ? s[S_SDS * n + qp] : 0.0;
    pcv[PC_RTAIL * ST_NPT] = 1.0 / d.tbg[((size_t)TQ_TAIL * F + (mijq - 1)) * n + qp];
    pcv[PC_FLMC * ST_NPT] = (1. - 0.9 * dmin(ci, 0.99)) * c_dc.flmin;
    pcv[PC_ICEADD * ST_NPT] = seticeq ? dmax(c_dc.EPSMIN, 1.0 - ci) * c_dc.flmin : 0.0;
    pcv[PC_ICEFREE * ST_NPT] = seticeq ? 0.0 : 1.0;
    pcv[PC_SNW * ST_NPT] = snw;
    pcv[PC_CSW * ST_NPT] = csw;
    pcv[PC_BETA * ST_NPT] = (c_dc.licerun && c_dc.lciscal) ? 1.0 - ci : 1.0;
  }
  __syncthreads();
  const double sinth = c_dc.SINTH[k], costh = c_dc.COSTH[k];
  V cw2;      // max(0, cos(TH(k) - WDWAVE))**2: FLM(k) = FLMC*cw2 (implsch.F90:236-247), SETICE's noise floor likewise
  int mij[NP];
  {
    const V snw = lds<NP>(sm, L.pc + PC_SNW * 64 + jo), csw = lds<NP>(sm, L.pc + PC_CSW * 64 + jo);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      cw2.v[i] = sq(dmax(0.0, costh * csw.v[i] + sinth * snw.v[i]));
      mij[i] = (int)s[S_MIJ * n + pq + i];
    }
  }
  const bool setice = c_dc.licerun && c_dc.lmaskice;
  const double delt = c_dc.delt, deltm = 1.0 / delt, delt5 = c_dc.ximp * delt;
  const double tmp03 = 1.0 / (c_dc.SDSBR * c_dc.MICHE);
  const int MLSTHG = c_dc.MLSTHG;
  const bool lssource = c_dc.lcflx && c_dc.lwvflx_snl, lsspre = c_dc.lcflx && !c_dc.lwvflx_snl;

  // pending SNONLIN sums of rows st-4 .. st+2 (row st+3 receives its first contribution, MC -> MC+3, at this step)
  double asl[7][NP], afl[7][NP];
#pragma unroll
  for (int x = 0; x < 7; ++x)
#pragma unroll
    for (int i = 0; i < NP; ++i) { asl[x][i] = 0.0; afl[x][i] = 0.0; }
  V fmij, a_philf, a_ts, a_tu, a_e1, a_e2, a_el;
#pragma unroll
  for (int i = 0; i < NP; ++i) { fmij.v[i] = 0.0; a_philf.v[i] = 0.0; a_ts.v[i] = 0.0; a_tu.v[i] = 0.0; a_e1.v[i] = 0.0; a_e2.v[i] = 0.0; a_el.v[i] = 0.0; }
  const unsigned lane = (unsigned)t & 31u;

  // one step of the sweep
  auto step = [&](const int st) {
    // ================= phase A =================
    const int rnew = st + 5, rfin = st - 4, rb = st - 2, rtb = st - 1;
    const unsigned par = (unsigned)st & 1u;
    const bool dia = st >= 0 && st < MLSTHG;
    const bool fin = rfin >= 0 && rfin < F;
    const bool sat = ard && rb >= 0 && rb < F;
    const unsigned p3 = (unsigned)(st + 6) % 3u;   // triple buffers: per-warp maxima, thread-private saturation values
    // row loads of this step land in shared memory behind the barrier: FL1 row st+5 (enters the ring in phase B),
    // wind-input (and XLLWS) row st-4 (consumed by the finish in phase B)
    if (rnew >= 0 && rnew < F) cp_async<NP * 8>(smb + me_s, (rnew < mlo ? src_lo : src_hi) + (size_t)rnew * rstr);
    if (fin) {
      cp_async<NP * 8>(smb + me_s + L.SSB, src_in + (size_t)rfin * rstr);
      if (LWFLUX) cp_async<NP * 8>(smb + me_s + 2 * L.SSB, src_xl + (size_t)rfin * rstr);
    }
    if (rtb >= 0 && rtb < F && t < TQ_N * ST_NPT) {   // per-(point, frequency) scalars of row rtb (used from the next step on)
      const int q = t >> 3, pt = t & 7;
      cp_async<8>(smb + L.tbs + (unsigned)(((rtb & 7) * TQ_N + q) * 64 + pt * 8), d.tbg + ((size_t)q * F + rtb) * n + min(pbase + pt, plast));
    }
    mbar_arrive_cp_async(smb + L.mbar);
    if (dia) {
      // DIA interaction values of centre frequency MC = st+1 (snonlin.F90:225-250)
      const int MC0 = st, MC = st + 1;
      const double* R = c_dc.RNLCOEF[MC0];   // spectrum-edge cases are folded into the coefficients (fill_dev_const)
      const unsigned bIC = L.ring + (unsigned)c_dc.NLSLOT[MC0][0] * RSB + me_r;
      const unsigned bIP = L.ring + (unsigned)c_dc.NLSLOT[MC0][1] * RSB + me_r;
      const unsigned bIP1 = L.ring + (unsigned)c_dc.NLSLOT[MC0][2] * RSB + me_r;
      const unsigned bIM = L.ring + (unsigned)c_dc.NLSLOT[MC0][3] * RSB + me_r;
      const unsigned bIM1 = L.ring + (unsigned)c_dc.NLSLOT[MC0][4] * RSB + me_r;
      const unsigned cb = L.cur + par * 6u * PSB + me_c;
      const V fc = lds<NP>(sm, bIC);
      V enh = lds<NP>(sm, L.pc + PC_ENH * 64 + jo);
      if (d.enh) {   // ISNONLIN = 1, 2: ENH(IJ,MC) from k_enh's plane
#pragma unroll
        for (int i = 0; i < NP; ++i) enh.v[i] = __ldg(d.enh + (size_t)MC0 * n + pq + i);
      }
      const double af11 = c_dc.AF11[MC0];
      V fij, fcen, ftemp, fcd1, fcd2;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        ftemp.v[i] = af11 * enh.v[i];
        fij.v[i] = fc.v[i] * R[0];
        fcen.v[i] = ftemp.v[i] * fij.v[i];
        fcd1.v[i] = c_dc.DAL1 * fcen.v[i];
        fcd2.v[i] = c_dc.DAL2 * fcen.v[i];
      }
      V csl, cfl;
#pragma unroll
      for (int i = 0; i < NP; ++i) { csl.v[i] = 0.0; cfl.v[i] = 0.0; }
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) {
        const V p1 = lds<NP>(sm, bIP + dsb[kh][0]), p11 = lds<NP>(sm, bIP + dsb[kh][1]);
        const V q1 = lds<NP>(sm, bIP1 + dsb[kh][0]), q11 = lds<NP>(sm, bIP1 + dsb[kh][1]);
        const V m2 = lds<NP>(sm, bIM + dsb[kh][2]), m21 = lds<NP>(sm, bIM + dsb[kh][3]);
        const V n2 = lds<NP>(sm, bIM1 + dsb[kh][2]), n21 = lds<NP>(sm, bIM1 + dsb[kh][3]);
        V vad, vdp, vdm;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const double sap = R[1] * p1.v[i] + R[2] * p11.v[i] + R[3] * q1.v[i] + R[4] * q11.v[i];
          const double sam = R[13] * m2.v[i] + R[14] * m21.v[i] + R[15] * n2.v[i] + R[16] * n21.v[i];
          double fad1 = fij.v[i] * (sap + sam);
          const double sap2 = 2.0 * sap;
          const double fad2 = fma(-sap2, sam, fad1);              // FAD1 - 2 SAP SAM
          fad1 = fad1 + fad2;
          vad.v[i] = fad2 * fcen.v[i];
          csl.v[i] += vad.v[i];
          cfl.v[i] = fma(fad1, ftemp.v[i], cfl.v[i]);
          vdp.v[i] = fma(-2.0, sam, fij.v[i]) * fcd1.v[i];        // DELAP = (FIJ - 2 SAM) DAL1 FCEN
          vdm.v[i] = (fij.v[i] - sap2) * fcd2.v[i];               // DELAM = (FIJ - 2 SAP) DAL2 FCEN
        }
        if (act) {
          sts<NP>(sm, cb + (0 + kh) * PSB, vad);
          sts<NP>(sm, cb + (2 + kh) * PSB, vdp);
          sts<NP>(sm, cb + (4 + kh) * PSB, vdm);
          if (k < HC) {
            sts<NP>(sm, cb + (0 + kh) * PSB + (unsigned)A * 64u, vad);
            sts<NP>(sm, cb + (2 + kh) * PSB + (unsigned)A * 64u, vdp);
            sts<NP>(sm, cb + (4 + kh) * PSB + (unsigned)A * 64u, vdm);
          }
          if (k >= A - HC) {
            sts<NP>(sm, cb + (0 + kh) * PSB - (unsigned)A * 64u, vad);
            sts<NP>(sm, cb + (2 + kh) * PSB - (unsigned)A * 64u, vdp);
            sts<NP>(sm, cb + (4 + kh) * PSB - (unsigned)A * 64u, vdm);
          }
        }
      }
      const double c2 = c_dc.RNLC2[MC0];
#pragma unroll
      for (int i = 0; i < NP; ++i) { asl[4][i] = fma(-c2, csl.v[i], asl[4][i]); afl[4][i] = fma(-c2, cfl.v[i], afl[4][i]); }
    }
    mbar_arrive(smb + L.mbar);   // the interaction planes of this step are written; what follows does not touch them
    // saturation spectrum of row rb for SDISSIP_ARD (sdissip_ard.F90:142-160): cyclic window of 2*NSDSNTH+1 directions
    if (sat) {
      const unsigned wb = L.ring + slot9(rb) * RSB + me_r - (unsigned)NSD * 64u;
      V b;
#pragma unroll
      for (int i = 0; i < NP; ++i) b.v[i] = 0.0;
#pragma unroll
      for (int x = 0; x < ST_NSMAX; ++x) {
        if (x < NS) {
          const double w = c_dc.SATW1[x];   // SATWEIGHTS(K, x) of the middle direction: cos(TH(K)-TH(K+x-NSD))**2 does not depend on K
          const V f = lds<NP>(sm, wb + x * 64);
#pragma unroll
          for (int i = 0; i < NP; ++i) b.v[i] += w * f.v[i];
        }
      }
      const V fs = lds<NP>(sm, L.tbs + (unsigned)(((rb & 7) * TQ_N + TQ_FACSAT) * 64) + jo);
      V mx;
#pragma unroll
      for (int i = 0; i < NP; ++i) mx.v[i] = b.v[i] * fs.v[i];
      sts<NP>(sm, L.bsl + p3 * L.SSB + (unsigned)t * (NP * 8), mx);
      // BTH0 = max over direction: over the directions of this warp with shuffles, then one partial per warp
#pragma unroll
      for (int o = 16; o >= NG; o >>= 1)
#pragma unroll
        for (int i = 0; i < NP; ++i) mx.v[i] = dmax(mx.v[i], __shfl_xor_sync(FULLMASK, mx.v[i], o));
      if (lane < NG) sts<NP>(sm, L.part + (p3 * (unsigned)nwarp + ((unsigned)t >> 5)) * 64u + jo, mx);
    }
    mbar_wait(smb + L.mbar, (unsigned)(st + 5) & 1u);
    // ================= phase B =================
    V tot_sl, tot_fl;   // SNONLIN sums of the row that is finished at this step
    {
      // gather the quadruplet contributions of MC into the pending rows (snonlin.F90:253-308, :333-410, :446-490) and slide the
      // window: row st-3 -> slot 0, ..., row st+3 (first contribution) -> slot 6
      V sl_mm, fl_mm, sl_mm1, fl_mm1, sl_mp, fl_mp, sl_mp1, fl_mp1;
#pragma unroll
      for (int i = 0; i < NP; ++i) { sl_mm.v[i] = fl_mm.v[i] = sl_mm1.v[i] = fl_mm1.v[i] = sl_mp.v[i] = fl_mp.v[i] = sl_mp1.v[i] = fl_mp1.v[i] = 0.0; }
      if (dia) {
        const double* R = c_dc.RNLCOEF[st];
        const unsigned cb = L.cur + par * 6u * PSB + me_c;
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {
          const unsigned cA = cb + (0 + kh) * PSB, cP = cb + (2 + kh) * PSB, cM = cb + (4 + kh) * PSB;
          const V a2 = lds<NP>(sm, cA - dsb[kh][2]), a21 = lds<NP>(sm, cA - dsb[kh][3]);
          const V m2 = lds<NP>(sm, cM - dsb[kh][2]), m21 = lds<NP>(sm, cM - dsb[kh][3]);
          const V a1 = lds<NP>(sm, cA - dsb[kh][0]), a11 = lds<NP>(sm, cA - dsb[kh][1]);
          const V q1 = lds<NP>(sm, cP - dsb[kh][0]), q11 = lds<NP>(sm, cP - dsb[kh][1]);
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            sl_mm.v[i] = fma(a21.v[i], R[19], fma(a2.v[i], R[20], sl_mm.v[i]));   fl_mm.v[i] = fma(m21.v[i], R[24], fma(m2.v[i], R[23], fl_mm.v[i]));     // FKLAMM1, FKLAMM2 | FKLAM12, FKLAM22
            sl_mm1.v[i] = fma(a21.v[i], R[18], fma(a2.v[i], R[17], sl_mm1.v[i])); fl_mm1.v[i] = fma(m21.v[i], R[22], fma(m2.v[i], R[21], fl_mm1.v[i]));   // FKLAMMA, FKLAMMB | FKLAMA2, FKLAMB2
            sl_mp.v[i] = fma(a11.v[i], R[7], fma(a1.v[i], R[8], sl_mp.v[i]));     fl_mp.v[i] = fma(q11.v[i], R[12], fma(q1.v[i], R[11], fl_mp.v[i]));     // FKLAMP1, FKLAMP2 | FKLAP12, FKLAP22
            sl_mp1.v[i] = fma(a11.v[i], R[6], fma(a1.v[i], R[5], sl_mp1.v[i]));   fl_mp1.v[i] = fma(q11.v[i], R[10], fma(q1.v[i], R[9], fl_mp1.v[i]));    // FKLAMPA, FKLAMPB | FKLAPA2, FKLAPB2
          }
        }
      }
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        tot_sl.v[i] = asl[0][i] + sl_mm.v[i];   tot_fl.v[i] = afl[0][i] + fl_mm.v[i];
        asl[0][i] = asl[1][i] + sl_mm1.v[i];    afl[0][i] = afl[1][i] + fl_mm1.v[i];
        asl[1][i] = asl[2][i];                  afl[1][i] = afl[2][i];
        asl[2][i] = asl[3][i];                  afl[2][i] = afl[3][i];
        asl[3][i] = asl[4][i];                  afl[3][i] = afl[4][i];
        asl[4][i] = asl[5][i];                  afl[4][i] = afl[5][i];
        asl[5][i] = asl[6][i] + sl_mp.v[i];     afl[5][i] = afl[6][i] + fl_mp.v[i];
        asl[6][i] = sl_mp1.v[i];                afl[6][i] = fl_mp1.v[i];
      }
    }
    const V flmc = lds<NP>(sm, L.pc + PC_FLMC * 64 + jo);
    V flm;
#pragma unroll
    for (int i = 0; i < NP; ++i) flm.v[i] = flmc.v[i] * cw2.v[i];
    // finish row rfin (implsch.F90:276-395 for these bins)
    if (fin) {
      const int r = rfin;
      const unsigned tq = L.tbs + (unsigned)((r & 7) * TQ_N * 64) + jo;
      const V fold = lds<NP>(sm, L.ring + slot9(r) * RSB + me_r);
      const V xI = lds<NP>(sm, me_s + L.SSB);
      const V usfm = lds<NP>(sm, L.pc + PC_USFMDELT * 64 + jo), sdsbk = lds<NP>(sm, L.pc + PC_SDSBK * 64 + jo);
      const V tsbo = lds<NP>(sm, tq + TQ_SBO * 64), tcinv = lds<NP>(sm, tq + TQ_CINV * 64), ttail = lds<NP>(sm, tq + TQ_TAIL * 64);
      const V tstf = lds<NP>(sm, tq + TQ_STF * 64), rtail = lds<NP>(sm, L.pc + PC_RTAIL * 64 + jo);
      V beta;
      if (c_dc.lciscal) beta = lds<NP>(sm, L.pc + PC_BETA * 64 + jo);
      V dd;
      if (ard) {
        const V b0 = lds<NP>(sm, L.bth0 + (par ^ 1u) * 64u + jo);
        const V b_prev = lds<NP>(sm, L.bsl + ((unsigned)(r + 8) % 3u) * L.SSB + (unsigned)t * (NP * 8));
        const double ssdsc2_sig = c_dc.SSDSC2 * c_dc.ZPIFR[r];
#pragma unroll
        for (int i = 0; i < NP; ++i)
          dd.v[i] = ssdsc2_sig * c_dc.SSDSC6 * sq(dmax(0., b0.v[i] * tmp03 - c_dc.SSDSC4)) +
                    ssdsc2_sig * (1. - c_dc.SSDSC6) * sq(dmax(0., b_prev.v[i] * tmp03 - c_dc.SSDSC4));
      } else dd = lds<NP>(sm, tq + TQ_JAN * 64);
      const double cofrm4 = c_dc.COFRM4[r], flmax = c_dc.FLMAX[r];
      const double rhowg = c_dc.RHOWG_DFIM[r];
      V fnv;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const double f0 = fold.v[i];
        double fldv = xI.v[i];                             // wind input (SINPUT, second SINFLX call)
        double slv = fldv * f0;
        slv = slv + dd.v[i] * f0; fldv = fldv + dd.v[i];   // SDISSIP
        slv = slv + tot_sl.v[i]; fldv = fldv + tot_fl.v[i];    // SNONLIN
        double ssource = 0.0;
        if (lssource) ssource = div_norm(slv, dmax(1.0 - delt5 * fldv, 1.0));
        else if (lsspre) ssource = slv - tot_sl.v[i];    // LWVFLX_SNL = F: SL before SNONLIN, not modulated (implsch.F90:279-288)
        if (r < c_dc.Fr) { slv = slv - sdsbk.v[i] * f0; fldv = fldv - sdsbk.v[i]; }   // SDIWBK (0 where it does not apply)
        if (c_dc.lciscal) { slv = slv * beta.v[i]; fldv = fldv * beta.v[i]; }        // LCISCAL (implsch.F90:315-325)
        if (d.ice2) {   // SDICE2 (sdice2.F90:97-113): ALP = CDICWA*k^2*4*sqrt(max(EPSMIN, F*DFIM))*ZALPFACB, the per-(point, frequency) part from k_ice
          const double e = __ldg(d.ice2 + (size_t)r * n + pq + i) * sqrt(dmax(c_dc.EPSMIN, f0 * c_dc.DFIM[r]));
          slv = slv - e * f0; fldv = fldv - e;
        }
        slv = slv + tsbo.v[i] * f0; fldv = fldv + tsbo.v[i];                         // SDICE1 + SDICE3 + SBOTTOM (plane is 0 where none applies)
        const double gtemp1 = dmax(1.0 - delt5 * fldv, 1.0);
        const double gtemp2 = div_norm(delt * slv, gtemp1);
        const double flhab = dmin(fabs(gtemp2), usfm.v[i] * cofrm4);
        double fn = f0 + copysign(flhab, gtemp2);
        fn = dmax(fn, flm.v[i]);
        ssource = ssource + deltm * dmin(flmax - fn, 0.0);
        fn = dmin(fn, flmax);
        {   // WNFLUXES sums (wnfluxes.F90:200-220); RHOWGDFTH of frcutindex.F90:99-108
          double rr = (r + 1 > mij[i]) ? 0.0 : rhowg;
          if (r + 1 == mij[i] && mij[i] != F) rr = 0.5 * rr;
          a_philf.v[i] += ssource * rr;
          a_ts.v[i] += ssource * (tcinv.v[i] * rr);
        }
        if (LWFLUX) {   // FEMEANWS on the new spectrum (before the tail is imposed)
          const V xL = lds<NP>(sm, me_s + 2 * L.SSB);
          const double xf = (xL.v[i] != 0.0) ? fn : 0.0;
          a_e1.v[i] += c_dc.DFIM[r] * xf; a_e2.v[i] += c_dc.DFIMOFR[r] * xf;
          if (r == F - 1) a_el.v[i] += xf;
        }
        if (r == mij[i] - 1) fmij.v[i] = fn;                                                   // IMPHFTAIL reference row
        if (r > mij[i] - 1) fn = dmax((ttail.v[i] * rtail.v[i]) * fmij.v[i], flm.v[i]);
        fnv.v[i] = fn;
      }
      if (setice) {                                                                            // SETICE
        const V icefree = lds<NP>(sm, L.pc + PC_ICEFREE * 64 + jo), ia = lds<NP>(sm, L.pc + PC_ICEADD * 64 + jo);
#pragma unroll
        for (int i = 0; i < NP; ++i) fnv.v[i] = fnv.v[i] * icefree.v[i] + ia.v[i] * cw2.v[i];
      }
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        a_tu.v[i] += tstf.v[i] * fnv.v[i];                                                     // STOKESDRIFT
        if (r == c_dc.NFRE_ODD - 1) {
          const double cst = 2.0 * c_dc.DELTH * c_dc.ZPI * c_dc.ZPI * c_dc.ZPI / c_dc.G * p4(c_dc.FR[c_dc.NFRE_ODD - 1]);
          a_tu.v[i] += cst * fnv.v[i];
        }
      }
      if (dost) stg<NP>(dst + (size_t)r * rstr, fnv);
    }
    if (rnew >= 0 && rnew < F) {   // depth-limited (+ floored at NFRE) row st+5 -> ring (slot of row st-4, whose last reader is the line above)
      const V xF = lds<NP>(sm, me_s);
      const V fac = lds<NP>(sm, L.pc + PC_FAC * 64 + jo);
      V v;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        v.v[i] = dmax(xF.v[i] * fac.v[i], c_dc.EPSMIN);
        if (rnew == F - 1) v.v[i] = dmax(v.v[i], flm.v[i]);
      }
      if (act) {
        const unsigned rbse = L.ring + slot9(rnew) * RSB + me_r;
        sts<NP>(sm, rbse, v);
        if (k < H) sts<NP>(sm, rbse + (unsigned)A * 64u, v);
        if (k >= A - H) sts<NP>(sm, rbse - (unsigned)A * 64u, v);
      }
    }
    if (ard && t < ST_NPT && rb - 1 >= 0 && rb - 1 < F) {   // BTH0 of row st-3 from the per-warp partial maxima of the previous step
      const double* pp = reinterpret_cast<const double*>(sm + L.part + ((unsigned)(st + 5) % 3u) * (unsigned)nwarp * 64u) + t;
      double mx = 0.0;
      for (int w = 0; w < nwarp; ++w) mx = dmax(mx, pp[w * ST_NPT]);
      reinterpret_cast<double*>(sm + L.bth0 + par * 64u)[t] = mx;
    }
  };

#pragma unroll 1
  for (int st = -5; st < MLSTHG; ++st) step(st);
  // ---- per-point sums over direction, then the scalar closures (one thread per point)
  __syncthreads();
  {
    // red[q][k][pt]: 8 planes of A*8 doubles in the (now free) ring area
    const unsigned RP = (unsigned)A * 64u;
    const unsigned ro = (unsigned)k * 64u + jo;
    V a_xs, a_ys, a_us, a_vs;
#pragma unroll
    for (int i = 0; i < NP; ++i) { a_xs.v[i] = sinth * a_ts.v[i]; a_ys.v[i] = costh * a_ts.v[i]; a_us.v[i] = a_tu.v[i] * sinth; a_vs.v[i] = a_tu.v[i] * costh; }
    if (act) {
      sts<NP>(sm, L.ring + 0 * RP + ro, a_philf); sts<NP>(sm, L.ring + 1 * RP + ro, a_xs); sts<NP>(sm, L.ring + 2 * RP + ro, a_ys);
      sts<NP>(sm, L.ring + 3 * RP + ro, a_us); sts<NP>(sm, L.ring + 4 * RP + ro, a_vs);
      if (LWFLUX) { sts<NP>(sm, L.ring + 5 * RP + ro, a_e1); sts<NP>(sm, L.ring + 6 * RP + ro, a_e2); sts<NP>(sm, L.ring + 7 * RP + ro, a_el); }
    }
  }
  __syncthreads();
  double* red = reinterpret_cast<double*>(sm + L.ring);
  double* qs = reinterpret_cast<double*>(sm + L.cur);      // [8][8] reduced sums
  if (t < 8 * ST_NPT) {
    const int x = t >> 3, pt = t & 7;
    double v = 0.0;
    if (x < 5 || LWFLUX) for (int kk = 0; kk < A; ++kk) v += red[(x * A + kk) * ST_NPT + pt];
    qs[x * ST_NPT + pt] = v;
  }
  __syncthreads();
  if (t < ST_NPT && pbase + t <= plast)
    stencil_closure<LWFLUX>(d, pbase + t, qs + t, reinterpret_cast<const double*>(sm + L.pc) + t);
}


// =========================================================================================================
// k_stencil_dp: the same sweep with thread = (TWO ADJACENT DIRECTIONS, one grid point) for the standard grids (compile-time
// geometry, even NANG).  The shared-memory rows are laid out [row][point][direction + cyclic halo] (direction fastest), so
//   * the 2*NSDSNTH+1 window of SDISSIP_ARD for both bins of a thread is NSDSNTH+1 aligned 16-byte loads (18 values for
//     two 17-point windows) instead of 17 per bin pair,
//   * the two bilinear partners of a DIA quadruplet family (shifts s and s+-1) of both bins are three consecutive values:
//     one 16-byte + one 8-byte load instead of two 16-byte loads,
//   * everything that is per grid point (the PC_* constants, the per-(point, frequency) scalars) is an 8-byte load that
//     eight lanes of a quarter-warp share, instead of a 16-byte load per lane.
// About 30 % fewer shared-memory wavefronts per bin than the two-points-per-thread mapping, which is what bounds that kernel;
// any NPROMA (no 16-byte alignment of the global rows is needed: 8-byte cp.async / stores, 8 points x 8 B = 64-byte segments).
// The point stride of a row is padded to 8 (mod 16) doubles so that the two points of a quarter-warp use disjoint banks.
// =========================================================================================================
__host__ __device__ constexpr int dp_hr(int A) { return geo_nsd(A) + (geo_nsd(A) & 1); }            // ring halo (even, >= NSDSNTH)
__host__ __device__ constexpr int dp_hc(int A) { return geo_hc(A) + (geo_hc(A) & 1); }              // interaction-plane halo (even)
__host__ __device__ constexpr int dp_pad(int n) { return n + ((8 - (n % 16)) + 16) % 16; }          // -> 8 (mod 16)
__host__ __device__ constexpr int dp_pst(int A) { return dp_pad(A + 2 * dp_hr(A)); }                // point stride of a ring row [doubles]
__host__ __device__ constexpr int dp_pstc(int A) { return dp_pad(A + 2 * dp_hc(A)); }               // point stride of an interaction plane
__host__ __device__ constexpr int dp_threads(int A) { return (((A / 2) * ST_NPT + 31) / 32) * 32; }
#ifndef ST_DP_FIFO
#define ST_DP_FIFO 0   // experiment for the next round: the three pending SNONLIN rows that receive nothing live in a per-thread
#endif                 // shared-memory FIFO instead of sliding through 24 registers every step
struct DpSmem { unsigned ring, cur, tbs, pc, part, bth0, stage, bsl, fifo, mbar, total, RSB, PSB, SSB; };
__host__ __device__ constexpr DpSmem dp_smem(int A, bool lwflux) {
  DpSmem s{};
  const int nth = dp_threads(A), nwarp = nth / 32;
  s.RSB = (unsigned)dp_pst(A) * ST_NPT * 8;
  s.PSB = (unsigned)dp_pstc(A) * ST_NPT * 8;
  s.SSB = (unsigned)nth * 16;
  unsigned o = 0;
  s.ring = o; o += ST_RING * s.RSB;
  s.cur = o; o += 2 * 6 * s.PSB;
  s.tbs = o; o += 8 * TQ_N * ST_NPT * 8;
  s.pc = o; o += PC_N * ST_NPT * 8;
  s.part = o; o += 3u * (unsigned)nwarp * ST_NPT * 8;
  s.bth0 = o; o += 2 * ST_NPT * 8;
  s.stage = o; o += (lwflux ? 3 : 2) * s.SSB;
  s.bsl = o; o += 3 * s.SSB;
#if ST_DP_FIFO
  s.fifo = o; o += 6 * s.SSB;    // 3 quiet pending rows x (SL, FLD), one 16-byte pair per thread and slot
#endif
  s.mbar = o; o += 16;
  s.total = o;
  return s;
}
__device__ __forceinline__ double lds1(const char* sm, unsigned off) { return *reinterpret_cast<const double*>(sm + off); }
// three consecutive directions starting LO elements from the thread's (even) first direction: one aligned pair + one single
template <int LO>
__device__ __forceinline__ void load3(const char* sm, unsigned base, double (&d)[3]) {
  constexpr bool even = ((LO % 2) + 2) % 2 == 0;
  if (even) {
    const Vd<2> a = lds<2>(sm, base + (unsigned)(LO * 8));
    d[0] = a.v[0]; d[1] = a.v[1]; d[2] = lds1(sm, base + (unsigned)((LO + 2) * 8));
  } else {
    d[0] = lds1(sm, base + (unsigned)(LO * 8));
    const Vd<2> a = lds<2>(sm, base + (unsigned)((LO + 1) * 8));
    d[1] = a.v[0]; d[2] = a.v[1];
  }
}
__host__ __device__ constexpr int cmin(int a, int b) { return a < b ? a : b; }

template <int TA, bool LWFLUX>
__global__ void __maxnreg__(ST_MAXREG) k_stencil_dp(ImplDev d, long long p0, long long np) {
  extern __shared__ __align__(16) char sm[];
  typedef Vd<2> V;
  constexpr int A = TA, NSD = geo_nsd(TA), NS = 2 * NSD + 1, HR = dp_hr(TA), HC = dp_hc(TA);
  constexpr unsigned PSTB = dp_pst(TA) * 8, PSTCB = dp_pstc(TA) * 8;
  constexpr DpSmem L = dp_smem(TA, LWFLUX);
  constexpr unsigned RSB = L.RSB, PSB = L.PSB;
  constexpr int nwarp = dp_threads(TA) / 32;
  const int F = c_dc.F;
  const bool ard = c_dc.iphys == 1;
  const unsigned smb = (unsigned)__cvta_generic_to_shared(sm);
  const int t = threadIdx.x;
  const int lane = t & 31, wq = t >> 5;
  const int pt = lane >> 2;                          // grid point of the CTA's eight
  int pr = 4 * wq + (lane & 3);                      // direction pair
  const bool act = pr < A / 2;                       // lanes beyond NANG/2 pairs shadow the last pair and never store
  if (!act) pr = A / 2 - 1;
  const int k0 = 2 * pr;
  // thread-invariant shared-memory offsets: made opaque so that ptxas keeps them in registers instead of re-deriving them
  // from threadIdx inside the sweep (11 % of the executed instructions were such re-derivations; 45.3 -> 44.9 ms)
  unsigned po = (unsigned)pt * 8u;
  unsigned me_r = (unsigned)pt * PSTB + (unsigned)(k0 + HR) * 8u;     // own pair inside a ring row
  unsigned me_c = (unsigned)pt * PSTCB + (unsigned)(k0 + HC) * 8u;    // own pair inside an interaction plane
  unsigned me_s = L.stage + (unsigned)t * 16u;                         // own landing slot
  asm volatile("" : "+r"(po), "+r"(me_r), "+r"(me_c), "+r"(me_s));
  // ---- points
  const long long pbase = p0 + (long long)blockIdx.x * ST_NPT;
  const long long plast = p0 + np - 1;
  long long pq = pbase + pt;
  const bool pvalid = pq <= plast;
  if (!pvalid) pq = plast;
  const bool dost = act && pvalid;
  const long long n = d.npts;
  const double* s = d.scr;
  const size_t P = (size_t)d.P;
  const size_t rstr = P * A;
  const long long pc_ = pq / d.P;
  const int pi_ = (int)(pq - pc_ * d.P);
  const size_t off_hi = (size_t)pi_ + P * A * F * (size_t)pc_ + P * (size_t)k0;
  const double* src_hi = d.f.fl1 + off_hi;
  const double* src_lo = src_hi;
  int mlo = 0;
  if (d.lo_on) {
    src_lo = d.fl_lo + (size_t)pi_ + P * A * d.lo_F * (size_t)pc_ + P * (size_t)k0;
    mlo = d.Fr;
  }
  double* dst = d.f.fl1 + off_hi;
  const double* src_in = d.fldin + off_hi;
  const double* src_xl = d.f.xllws + off_hi;

  // ---- prologue: per-point constants, barrier
  if (t == 0) mbar_init(smb + L.mbar, 2u * blockDim.x);   // per phase and thread: one arrive + one arrive on completion of its cp.async
  if (t < ST_NPT) {
    const long long qp = min(pbase + t, plast);
    double* pcv = reinterpret_cast<double*>(sm + L.pc) + t;
    const double wd = d.f.wdwave[qp], ci = d.f.cicover[qp], dep = d.f.depth[qp];
    double snw, csw;
    sincos(wd, &snw, &csw);
    double enhfr = dmax(0.75 * dep * s[S_AKMEAN * n + qp], 0.5);
    enhfr = 1.0 + (5.5 / enhfr) * (1.0 - .833 * enhfr) * exp(-1.25 * enhfr);
    const int mijq = (int)s[S_MIJ * n + qp];
    const bool seticeq = c_dc.licerun && c_dc.lmaskice && ci > c_dc.cithrsh;
    pcv[PC_FAC * ST_NPT] = s[S_FAC * n + qp];
    pcv[PC_ENH * ST_NPT] = enhfr;
    pcv[PC_USFMDELT * ST_NPT] = s[S_USFM * n + qp] * c_dc.delt;
    pcv[PC_SDSBK * ST_NPT] = (c_dc.lbiwbk && dep < 50.0) ? s[S_SDS * n + qp] : 0.0;
    pcv[PC_RTAIL * ST_NPT] = 1.0 / d.tbg[((size_t)TQ_TAIL * F + (mijq - 1)) * n + qp];
    pcv[PC_FLMC * ST_NPT] = (1. - 0.9 * dmin(ci, 0.99)) * c_dc.flmin;
    pcv[PC_ICEADD * ST_NPT] = seticeq ? dmax(c_dc.EPSMIN, 1.0 - ci) * c_dc.flmin : 0.0;
    pcv[PC_ICEFREE * ST_NPT] = seticeq ? 0.0 : 1.0;
    pcv[PC_SNW * ST_NPT] = snw;
    pcv[PC_CSW * ST_NPT] = csw;
    pcv[PC_BETA * ST_NPT] = (c_dc.licerun && c_dc.lciscal) ? 1.0 - ci : 1.0;
  }
  __syncthreads();
  double sinth[2], costh[2];
  V flm, iaw;    // FLM(k) = FLMC*max(0,cos(TH(k)-WDWAVE))**2 (implsch.F90:236-247); SETICE's noise floor likewise
  {
    const double snw = lds1(sm, L.pc + PC_SNW * 64 + po), csw = lds1(sm, L.pc + PC_CSW * 64 + po);
    const double flmc = lds1(sm, L.pc + PC_FLMC * 64 + po), ia = lds1(sm, L.pc + PC_ICEADD * 64 + po);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      sinth[i] = c_dc.SINTH[k0 + i]; costh[i] = c_dc.COSTH[k0 + i];
      const double cw2 = sq(dmax(0.0, costh[i] * csw + sinth[i] * snw));
      flm.v[i] = flmc * cw2; iaw.v[i] = ia * cw2;
    }
  }
  const int mij = (int)s[S_MIJ * n + pq];
  const bool setice = c_dc.licerun && c_dc.lmaskice;
  const double delt = c_dc.delt, deltm = 1.0 / delt, delt5 = c_dc.ximp * delt;
  const double tmp03 = 1.0 / (c_dc.SDSBR * c_dc.MICHE);
  const int MLSTHG = c_dc.MLSTHG;
  const bool lssource = c_dc.lcflx && c_dc.lwvflx_snl, lsspre = c_dc.lcflx && !c_dc.lwvflx_snl;

  double asl[7][2], afl[7][2];
#pragma unroll
  for (int x = 0; x < 7; ++x)
#pragma unroll
    for (int i = 0; i < 2; ++i) { asl[x][i] = 0.0; afl[x][i] = 0.0; }
#if ST_DP_FIFO
  unsigned fifo_at = L.fifo + (unsigned)t * 16u;   // slot of the oldest quiet row (the rows at positions 1..3 of the slide)
  {
    V z; z.v[0] = 0.0; z.v[1] = 0.0;
#pragma unroll
    for (int x = 0; x < 6; ++x) sts<2>(sm, L.fifo + (unsigned)x * L.SSB + (unsigned)t * 16u, z);
  }
#endif
  V fmij, a_philf, a_ts, a_tu, a_e1, a_e2, a_el;
#pragma unroll
  for (int i = 0; i < 2; ++i) { fmij.v[i] = 0.0; a_philf.v[i] = 0.0; a_ts.v[i] = 0.0; a_tu.v[i] = 0.0; a_e1.v[i] = 0.0; a_e2.v[i] = 0.0; a_el.v[i] = 0.0; }

  auto step = [&](const int st) {
    // ================= phase A =================
    const int rnew = st + 5, rfin = st - 4, rb = st - 2, rtb = st - 1;
    const unsigned par = (unsigned)st & 1u;
    const bool dia = st >= 0 && st < MLSTHG;
    const bool fin = rfin >= 0 && rfin < F;
    const bool sat = ard && rb >= 0 && rb < F;
    const unsigned p3 = (unsigned)(st + 6) % 3u;
    if (rnew >= 0 && rnew < F) {
      const double* g = (rnew < mlo ? src_lo : src_hi) + (size_t)rnew * rstr;
      cp_async<8>(smb + me_s, g); cp_async<8>(smb + me_s + 8, g + P);
    }
    if (fin) {
      const double* g = src_in + (size_t)rfin * rstr;
      cp_async<8>(smb + me_s + L.SSB, g); cp_async<8>(smb + me_s + L.SSB + 8, g + P);
      if (LWFLUX) { const double* gx = src_xl + (size_t)rfin * rstr; cp_async<8>(smb + me_s + 2 * L.SSB, gx); cp_async<8>(smb + me_s + 2 * L.SSB + 8, gx + P); }
    }
    if (rtb >= 0 && rtb < F && t < TQ_N * ST_NPT) {
      const int q = t >> 3, p8 = t & 7;
      cp_async<8>(smb + L.tbs + (unsigned)(((rtb & 7) * TQ_N + q) * 64 + p8 * 8), d.tbg + ((size_t)q * F + rtb) * n + min(pbase + p8, plast));
    }
    mbar_arrive_cp_async(smb + L.mbar);
    if (dia) {
      const int MC0 = st;
      const double* R = c_dc.RNLCOEF[MC0];
      const unsigned bIC = L.ring + (unsigned)c_dc.NLSLOT[MC0][0] * RSB + me_r;
      const unsigned bIP = L.ring + (unsigned)c_dc.NLSLOT[MC0][1] * RSB + me_r;
      const unsigned bIP1 = L.ring + (unsigned)c_dc.NLSLOT[MC0][2] * RSB + me_r;
      const unsigned bIM = L.ring + (unsigned)c_dc.NLSLOT[MC0][3] * RSB + me_r;
      const unsigned bIM1 = L.ring + (unsigned)c_dc.NLSLOT[MC0][4] * RSB + me_r;
      const unsigned cb = L.cur + par * 6u * PSB + me_c;
      const V fc = lds<2>(sm, bIC);
      const double enh = d.enh ? __ldg(d.enh + (size_t)MC0 * n + pq) : lds1(sm, L.pc + PC_ENH * 64 + po);   // ISNONLIN = 1, 2: k_enh's plane
      const double ftemp = c_dc.AF11[MC0] * enh;
      V fij, fcen, fcd1, fcd2;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        fij.v[i] = fc.v[i] * R[0];
        fcen.v[i] = ftemp * fij.v[i];
        fcd1.v[i] = c_dc.DAL1 * fcen.v[i];
        fcd2.v[i] = c_dc.DAL2 * fcen.v[i];
      }
      V csl, cfl;
#pragma unroll
      for (int i = 0; i < 2; ++i) { csl.v[i] = 0.0; cfl.v[i] = 0.0; }
      auto half = [&](auto KH) {
        constexpr int kh = decltype(KH)::value;
        constexpr int s0 = geo_sh(TA, kh, 0), s1 = geo_sh(TA, kh, 1), s2 = geo_sh(TA, kh, 2), s3 = geo_sh(TA, kh, 3);
        constexpr int lp = cmin(s0, s1), lm = cmin(s2, s3);
        double p[3], q[3], m[3], nn[3];
        load3<lp>(sm, bIP, p); load3<lp>(sm, bIP1, q); load3<lm>(sm, bIM, m); load3<lm>(sm, bIM1, nn);
        V vad, vdp, vdm;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const double sap = R[1] * p[s0 - lp + i] + R[2] * p[s1 - lp + i] + R[3] * q[s0 - lp + i] + R[4] * q[s1 - lp + i];
          const double sam = R[13] * m[s2 - lm + i] + R[14] * m[s3 - lm + i] + R[15] * nn[s2 - lm + i] + R[16] * nn[s3 - lm + i];
          double fad1 = fij.v[i] * (sap + sam);
          const double sap2 = 2.0 * sap;
          const double fad2 = fma(-sap2, sam, fad1);
          fad1 = fad1 + fad2;
          vad.v[i] = fad2 * fcen.v[i];
          csl.v[i] += vad.v[i];
          cfl.v[i] = fma(fad1, ftemp, cfl.v[i]);
          vdp.v[i] = fma(-2.0, sam, fij.v[i]) * fcd1.v[i];
          vdm.v[i] = (fij.v[i] - sap2) * fcd2.v[i];
        }
        if (act) {
          sts<2>(sm, cb + (0 + kh) * PSB, vad);
          sts<2>(sm, cb + (2 + kh) * PSB, vdp);
          sts<2>(sm, cb + (4 + kh) * PSB, vdm);
          if (k0 < HC) {
            sts<2>(sm, cb + (0 + kh) * PSB + (unsigned)A * 8u, vad);
            sts<2>(sm, cb + (2 + kh) * PSB + (unsigned)A * 8u, vdp);
            sts<2>(sm, cb + (4 + kh) * PSB + (unsigned)A * 8u, vdm);
          }
          if (k0 >= A - HC) {
            sts<2>(sm, cb + (0 + kh) * PSB - (unsigned)A * 8u, vad);
            sts<2>(sm, cb + (2 + kh) * PSB - (unsigned)A * 8u, vdp);
            sts<2>(sm, cb + (4 + kh) * PSB - (unsigned)A * 8u, vdm);
          }
        }
      };
      half(IC<0>{}); half(IC<1>{});
      const double c2 = c_dc.RNLC2[MC0];
#pragma unroll
      for (int i = 0; i < 2; ++i) { asl[4][i] = fma(-c2, csl.v[i], asl[4][i]); afl[4][i] = fma(-c2, cfl.v[i], afl[4][i]); }
    }
    mbar_arrive(smb + L.mbar);
    // saturation spectrum of row rb (sdissip_ard.F90:142-160): both windows of the pair from HR+1 aligned 16-byte loads
    if (sat) {
      const unsigned wb = L.ring + slot9(rb) * RSB + me_r - (unsigned)HR * 8u;
      V b;
      b.v[0] = 0.0; b.v[1] = 0.0;
#pragma unroll
      for (int jp = 0; jp <= HR; ++jp) {
        const V f = lds<2>(sm, wb + jp * 16);
        // f.v[e] is direction k0 - HR + 2 jp + e; window index of bin i: x = (2 jp + e) - (HR - NSD) - i
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int x = 2 * jp + e - (HR - NSD) - i;
            if (x >= 0 && x < NS) b.v[i] += c_dc.SATW1[x] * f.v[e];
          }
      }
      const double fs = lds1(sm, L.tbs + (unsigned)(((rb & 7) * TQ_N + TQ_FACSAT) * 64) + po);
      V mx;
      mx.v[0] = b.v[0] * fs; mx.v[1] = b.v[1] * fs;
      sts<2>(sm, L.bsl + p3 * L.SSB + (unsigned)t * 16u, mx);
      double m1 = dmax(mx.v[0], mx.v[1]);     // BTH0 = max over direction: the four pairs of this warp, then one partial per warp
      m1 = dmax(m1, __shfl_xor_sync(FULLMASK, m1, 1));
      m1 = dmax(m1, __shfl_xor_sync(FULLMASK, m1, 2));
      if ((lane & 3) == 0) *reinterpret_cast<double*>(sm + L.part + (p3 * (unsigned)nwarp + (unsigned)wq) * 64u + po) = m1;
    }
    mbar_wait(smb + L.mbar, (unsigned)(st + 5) & 1u);
    // ================= phase B =================
    V tot_sl, tot_fl;
    {
      V sl_mm, fl_mm, sl_mm1, fl_mm1, sl_mp, fl_mp, sl_mp1, fl_mp1;
#pragma unroll
      for (int i = 0; i < 2; ++i) { sl_mm.v[i] = fl_mm.v[i] = sl_mm1.v[i] = fl_mm1.v[i] = sl_mp.v[i] = fl_mp.v[i] = sl_mp1.v[i] = fl_mp1.v[i] = 0.0; }
      if (dia) {
        const double* R = c_dc.RNLCOEF[st];
        const unsigned cb = L.cur + par * 6u * PSB + me_c;
        auto half = [&](auto KH) {
          constexpr int kh = decltype(KH)::value;
          constexpr int s0 = -geo_sh(TA, kh, 0), s1 = -geo_sh(TA, kh, 1), s2 = -geo_sh(TA, kh, 2), s3 = -geo_sh(TA, kh, 3);
          constexpr int lp = cmin(s0, s1), lm = cmin(s2, s3);
          const unsigned cA = cb + (0 + kh) * PSB, cP = cb + (2 + kh) * PSB, cM = cb + (4 + kh) * PSB;
          double am[3], mm[3], ap[3], pp[3];
          load3<lm>(sm, cA, am); load3<lm>(sm, cM, mm); load3<lp>(sm, cA, ap); load3<lp>(sm, cP, pp);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const double a2 = am[s2 - lm + i], a21 = am[s3 - lm + i], m2 = mm[s2 - lm + i], m21 = mm[s3 - lm + i];
            const double a1 = ap[s0 - lp + i], a11 = ap[s1 - lp + i], q1 = pp[s0 - lp + i], q11 = pp[s1 - lp + i];
            sl_mm.v[i] = fma(a21, R[19], fma(a2, R[20], sl_mm.v[i]));   fl_mm.v[i] = fma(m21, R[24], fma(m2, R[23], fl_mm.v[i]));
            sl_mm1.v[i] = fma(a21, R[18], fma(a2, R[17], sl_mm1.v[i])); fl_mm1.v[i] = fma(m21, R[22], fma(m2, R[21], fl_mm1.v[i]));
            sl_mp.v[i] = fma(a11, R[7], fma(a1, R[8], sl_mp.v[i]));     fl_mp.v[i] = fma(q11, R[12], fma(q1, R[11], fl_mp.v[i]));
            sl_mp1.v[i] = fma(a11, R[6], fma(a1, R[5], sl_mp1.v[i]));   fl_mp1.v[i] = fma(q11, R[10], fma(q1, R[9], fl_mp1.v[i]));
          }
        };
        half(IC<0>{}); half(IC<1>{});
      }
#if ST_DP_FIFO
      {   // pop the oldest quiet row (position 1), push the row leaving position 4 into the slot it frees
        const V q_sl = lds<2>(sm, fifo_at), q_fl = lds<2>(sm, fifo_at + 3u * L.SSB);
        V p_sl, p_fl;
#pragma unroll
        for (int i = 0; i < 2; ++i) { p_sl.v[i] = asl[4][i]; p_fl.v[i] = afl[4][i]; }
        sts<2>(sm, fifo_at, p_sl); sts<2>(sm, fifo_at + 3u * L.SSB, p_fl);
        fifo_at = (fifo_at == L.fifo + 2u * L.SSB + (unsigned)t * 16u) ? L.fifo + (unsigned)t * 16u : fifo_at + L.SSB;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          tot_sl.v[i] = asl[0][i] + sl_mm.v[i];   tot_fl.v[i] = afl[0][i] + fl_mm.v[i];
          asl[0][i] = q_sl.v[i] + sl_mm1.v[i];    afl[0][i] = q_fl.v[i] + fl_mm1.v[i];
          asl[4][i] = asl[5][i];                  afl[4][i] = afl[5][i];
          asl[5][i] = asl[6][i] + sl_mp.v[i];     afl[5][i] = afl[6][i] + fl_mp.v[i];
          asl[6][i] = sl_mp1.v[i];                afl[6][i] = fl_mp1.v[i];
        }
      }
#else
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        tot_sl.v[i] = asl[0][i] + sl_mm.v[i];   tot_fl.v[i] = afl[0][i] + fl_mm.v[i];
        asl[0][i] = asl[1][i] + sl_mm1.v[i];    afl[0][i] = afl[1][i] + fl_mm1.v[i];
        asl[1][i] = asl[2][i];                  afl[1][i] = afl[2][i];
        asl[2][i] = asl[3][i];                  afl[2][i] = afl[3][i];
        asl[3][i] = asl[4][i];                  afl[3][i] = afl[4][i];
        asl[4][i] = asl[5][i];                  afl[4][i] = afl[5][i];
        asl[5][i] = asl[6][i] + sl_mp.v[i];     afl[5][i] = afl[6][i] + fl_mp.v[i];
        asl[6][i] = sl_mp1.v[i];                afl[6][i] = fl_mp1.v[i];
      }
#endif
    }
    // finish row rfin (implsch.F90:276-395 for these bins)
    if (fin) {
      const int r = rfin;
      const unsigned tq = L.tbs + (unsigned)((r & 7) * TQ_N * 64) + po;
      const V fold = lds<2>(sm, L.ring + slot9(r) * RSB + me_r);
      const V xI = lds<2>(sm, me_s + L.SSB);
      const double usfm = lds1(sm, L.pc + PC_USFMDELT * 64 + po), sdsbk = lds1(sm, L.pc + PC_SDSBK * 64 + po);
      const double tsbo = lds1(sm, tq + TQ_SBO * 64), tcinv = lds1(sm, tq + TQ_CINV * 64), ttail = lds1(sm, tq + TQ_TAIL * 64);
      const double tstf = lds1(sm, tq + TQ_STF * 64), rtail = lds1(sm, L.pc + PC_RTAIL * 64 + po);
      const double beta = c_dc.lciscal ? lds1(sm, L.pc + PC_BETA * 64 + po) : 1.0;
      V dd;
      if (ard) {
        const double b0 = lds1(sm, L.bth0 + (par ^ 1u) * 64u + po);
        const V b_prev = lds<2>(sm, L.bsl + ((unsigned)(r + 8) % 3u) * L.SSB + (unsigned)t * 16u);
        const double ssdsc2_sig = c_dc.SSDSC2 * c_dc.ZPIFR[r];
        const double d0 = ssdsc2_sig * c_dc.SSDSC6 * sq(dmax(0., b0 * tmp03 - c_dc.SSDSC4));
#pragma unroll
        for (int i = 0; i < 2; ++i) dd.v[i] = d0 + ssdsc2_sig * (1. - c_dc.SSDSC6) * sq(dmax(0., b_prev.v[i] * tmp03 - c_dc.SSDSC4));
      } else { const double dj = lds1(sm, tq + TQ_JAN * 64); dd.v[0] = dj; dd.v[1] = dj; }
      const double cofrm4 = c_dc.COFRM4[r], flmax = c_dc.FLMAX[r];
      double rr = (r + 1 > mij) ? 0.0 : c_dc.RHOWG_DFIM[r];      // RHOWGDFTH of frcutindex.F90:99-108
      if (r + 1 == mij && mij != F) rr = 0.5 * rr;
      V xL;
      if (LWFLUX) xL = lds<2>(sm, me_s + 2 * L.SSB);
      V fnv;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double f0 = fold.v[i];
        double fldv = xI.v[i];
        double slv = fldv * f0;
        slv = slv + dd.v[i] * f0; fldv = fldv + dd.v[i];
        slv = slv + tot_sl.v[i]; fldv = fldv + tot_fl.v[i];
        double ssource = 0.0;
        if (lssource) ssource = div_norm(slv, dmax(1.0 - delt5 * fldv, 1.0));
        else if (lsspre) ssource = slv - tot_sl.v[i];    // LWVFLX_SNL = F: SL before SNONLIN, not modulated (implsch.F90:279-288)
        if (r < c_dc.Fr) { slv = slv - sdsbk * f0; fldv = fldv - sdsbk; }            // SDIWBK (0 where it does not apply)
        if (c_dc.lciscal) { slv = slv * beta; fldv = fldv * beta; }                  // LCISCAL (implsch.F90:315-325)
        if (d.ice2) {   // SDICE2 (sdice2.F90:97-113), see k_stencil
          const double e = __ldg(d.ice2 + (size_t)r * n + pq) * sqrt(dmax(c_dc.EPSMIN, f0 * c_dc.DFIM[r]));
          slv = slv - e * f0; fldv = fldv - e;
        }
        slv = slv + tsbo * f0; fldv = fldv + tsbo;                                   // SDICE3 + SBOTTOM (plane is 0 where neither applies)
        const double gtemp1 = dmax(1.0 - delt5 * fldv, 1.0);
        const double gtemp2 = div_norm(delt * slv, gtemp1);
        const double flhab = dmin(fabs(gtemp2), usfm * cofrm4);
        double fn = f0 + copysign(flhab, gtemp2);
        fn = dmax(fn, flm.v[i]);
        ssource = ssource + deltm * dmin(flmax - fn, 0.0);
        fn = dmin(fn, flmax);
        a_philf.v[i] += ssource * rr;
        a_ts.v[i] += ssource * (tcinv * rr);
        if (LWFLUX) {
          const double xf = (xL.v[i] != 0.0) ? fn : 0.0;
          a_e1.v[i] += c_dc.DFIM[r] * xf; a_e2.v[i] += c_dc.DFIMOFR[r] * xf;
          if (r == F - 1) a_el.v[i] += xf;
        }
        if (r == mij - 1) fmij.v[i] = fn;
        if (r > mij - 1) fn = dmax((ttail * rtail) * fmij.v[i], flm.v[i]);
        fnv.v[i] = fn;
      }
      if (setice) {
        const double icefree = lds1(sm, L.pc + PC_ICEFREE * 64 + po);
#pragma unroll
        for (int i = 0; i < 2; ++i) fnv.v[i] = fnv.v[i] * icefree + iaw.v[i];
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        a_tu.v[i] += tstf * fnv.v[i];
        if (r == c_dc.NFRE_ODD - 1) {
          const double cst = 2.0 * c_dc.DELTH * c_dc.ZPI * c_dc.ZPI * c_dc.ZPI / c_dc.G * p4(c_dc.FR[c_dc.NFRE_ODD - 1]);
          a_tu.v[i] += cst * fnv.v[i];
        }
      }
      if (dost) { double* o = dst + (size_t)r * rstr; o[0] = fnv.v[0]; o[P] = fnv.v[1]; }
    }
    if (rnew >= 0 && rnew < F) {
      const V xF = lds<2>(sm, me_s);
      const double fac = lds1(sm, L.pc + PC_FAC * 64 + po);
      V v;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        v.v[i] = dmax(xF.v[i] * fac, c_dc.EPSMIN);
        if (rnew == F - 1) v.v[i] = dmax(v.v[i], flm.v[i]);
      }
      if (act) {
        const unsigned rbse = L.ring + slot9(rnew) * RSB + me_r;
        sts<2>(sm, rbse, v);
        if (k0 < HR) sts<2>(sm, rbse + (unsigned)A * 8u, v);
        if (k0 >= A - HR) sts<2>(sm, rbse - (unsigned)A * 8u, v);
      }
    }
    if (ard && t < ST_NPT && rb - 1 >= 0 && rb - 1 < F) {
      const double* pp = reinterpret_cast<const double*>(sm + L.part + ((unsigned)(st + 5) % 3u) * (unsigned)nwarp * 64u) + t;
      double mx = 0.0;
      for (int w = 0; w < nwarp; ++w) mx = dmax(mx, pp[w * ST_NPT]);
      reinterpret_cast<double*>(sm + L.bth0 + par * 64u)[t] = mx;
    }
  };

#pragma unroll 1
  for (int st = -5; st < MLSTHG; ++st) step(st);
  // ---- per-point sums over direction, then the scalar closures (one thread per point)
  __syncthreads();
  double* red = reinterpret_cast<double*>(sm + L.ring);     // red[q][k][pt]: 8 planes of A*8 doubles in the (now free) ring area
  if (act) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = k0 + i;
      red[(0 * A + k) * ST_NPT + pt] = a_philf.v[i];
      red[(1 * A + k) * ST_NPT + pt] = sinth[i] * a_ts.v[i];
      red[(2 * A + k) * ST_NPT + pt] = costh[i] * a_ts.v[i];
      red[(3 * A + k) * ST_NPT + pt] = a_tu.v[i] * sinth[i];
      red[(4 * A + k) * ST_NPT + pt] = a_tu.v[i] * costh[i];
      if (LWFLUX) { red[(5 * A + k) * ST_NPT + pt] = a_e1.v[i]; red[(6 * A + k) * ST_NPT + pt] = a_e2.v[i]; red[(7 * A + k) * ST_NPT + pt] = a_el.v[i]; }
    }
  }
  __syncthreads();
  double* qs = reinterpret_cast<double*>(sm + L.cur);
  if (t < 8 * ST_NPT) {
    const int x = t >> 3, p8 = t & 7;
    double v = 0.0;
    if (x < 5 || LWFLUX) for (int kk = 0; kk < A; ++kk) v += red[(x * A + kk) * ST_NPT + p8];
    qs[x * ST_NPT + p8] = v;
  }
  __syncthreads();
  if (t < ST_NPT && pbase + t <= plast)
    stencil_closure<LWFLUX>(d, pbase + t, qs + t, reinterpret_cast<const double*>(sm + L.pc) + t);
}

template <int TA, bool LW>
static int launch_stencil_dp(const ImplDev& d, long long p0, long long np, cudaStream_t st) {
  constexpr DpSmem L = dp_smem(TA, LW);
  static_assert(L.total <= 110 * 1024, "k_stencil_dp shared memory");
  static bool attr_done = false;
  if (!attr_done) {
    EW_CUDA_CHECK(cudaFuncSetAttribute(k_stencil_dp<TA, LW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    EW_CUDA_CHECK(cudaFuncSetAttribute(k_stencil_dp<TA, LW>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_done = true;
  }
  k_stencil_dp<TA, LW><<<(unsigned)((np + ST_NPT - 1) / ST_NPT), dp_threads(TA), L.total, st>>>(d, p0, np);
  return 0;
}

template <int TA, int NP, bool LW>
static int launch_stencil(const ImplDev& d, long long p0, long long np, cudaStream_t st) {
  const int nth = stencil_threads(d.A, NP);
  const StencilSmem L = stencil_smem(d.A, TA > 0 ? geo_nsd(TA) : d.halo_r, TA > 0 ? geo_hc(TA) : d.halo_c, nth, NP, LW);
  static bool attr_done = false;
  if (!attr_done) {
    EW_CUDA_CHECK(cudaFuncSetAttribute(k_stencil<TA, NP, LW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    EW_CUDA_CHECK(cudaFuncSetAttribute(k_stencil<TA, NP, LW>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_done = true;
  }
  if (L.total > 110 * 1024) { ew_set_error("k_stencil: shared memory %u B", L.total); return ECWAM_B200_EINVAL; }
  k_stencil<TA, NP, LW><<<(unsigned)((np + ST_NPT - 1) / ST_NPT), nth, L.total, st>>>(d, p0, np, L);
  return 0;
}

// =========================================================================================================
// k_sweep: the default instance of the frequency sweep for the standard grids.  Same algorithm and hazard structure as
// k_stencil_dp (thread = two adjacent directions of one grid point, rows [row][point][direction + cyclic halo], one split
// barrier per step), re-cut for the machine:
//   * CTA = NPT grid points x NANG/2 direction pairs with consecutive threads = consecutive pairs of one point.  For NANG = 36,
//     NPT = 7 gives 126 threads = 4 warps (2 idle lanes instead of 16 of 160) and one warp per SM sub-partition and CTA, so
//     3 CTAs per SM load the four schedulers evenly (5-warp CTAs put 3 + 3 + 2 + 2 warps on them: the slow pair sets the pace).
//     Thread order and point stride are chosen so that every 16-byte shared-memory access of a warp is conflict-free AND the
//     global row loads / stores of a warp are a few 56-byte runs (see sw_step below).
//   * the DIA weights in separable form (DevConst::NLW): the direction interpolation D+(row) = CL11 F(K1W) + ACL1 F(K11W) of a
//     row is formed ONCE, when the row is IP1 of the running centre frequency, and carried in registers to the next step where
//     the same row is IP (likewise D- for IM1 -> IM): half the partner loads of the product phase; the gather phase forms the
//     direction-interpolated sums first and scales them per target row: 24 instead of 32 FMAs per bin.
//   * both KH partners of the P rows are one run of six consecutive directions (three aligned 16-byte loads).
//   * the saturation values of the rows between their window sum and their finish live in registers, not in shared memory.
// Shared memory per CTA (NANG = 36): 9-row ring 30.2 KB + 2 x 6 interaction planes 29.6 KB + scalars = 70.3 KB -> 3 CTAs per SM.
// =========================================================================================================
// Thread order inside the CTA: t = (g * NPT + pt) * 2 + j with direction pair 2 g + j: a quarter-warp is four grid points x two
// adjacent pairs, i.e. four 32-byte pieces of shared memory, and the same pair of seven or eight consecutive grid points is
// 56 / 64 consecutive bytes of global memory (one or two sectors per direction instead of one per lane).  The four pieces use
// distinct bank groups when the point stride is 32 a (mod 128) bytes with a odd, and, where quarter-warps straddle two pair
// groups (2 NPT not a multiple of 8), a = 1 / NPT (mod 4): the pieces then step through the bank groups uniformly.
__host__ __device__ constexpr int sw_step(int npt) { return (2 * npt) % 8 == 0 ? 1 : ((npt % 4 == 1) ? 1 : 3); }
__host__ __device__ constexpr int sw_pad(int need, int npt) {   // >= need, = 4 a (mod 16) doubles (a = 1: also 12 (mod 16) when the groups are aligned)
  int s = need;
  while (!(s % 16 == 4 * sw_step(npt) || ((2 * npt) % 8 == 0 && s % 16 == 12))) ++s;
  return s;
}
__host__ __device__ constexpr int sw_pst(int A, int npt) { return sw_pad(A + 2 * dp_hr(A), npt); }
__host__ __device__ constexpr int sw_pstc(int A, int npt) { return sw_pad(A + 2 * dp_hc(A), npt); }
__host__ __device__ constexpr int sw_threads(int A, int npt) { return (((A / 2) * npt + 31) / 32) * 32; }
__host__ __device__ constexpr int sw_nptp(int npt) { return (npt + 7) / 8 * 8; }
struct SwSmem { unsigned ring, cur, tbs, pc, part, bth0, stage, mbar, total, RSB, PSB, SSB, PRB; };
__host__ __device__ constexpr SwSmem sw_smem(int A, int npt, bool lwflux) {
  SwSmem s{};
  const int nth = sw_threads(A, npt);
  s.RSB = (unsigned)sw_pst(A, npt) * npt * 8;
  s.PSB = (unsigned)sw_pstc(A, npt) * npt * 8;
  s.SSB = (unsigned)nth * 16;
  s.PRB = (unsigned)sw_nptp(npt) * 8;
  unsigned o = 0;
  s.ring = o; o += ST_RING * s.RSB;
  s.cur = o; o += 2 * 6 * s.PSB;
  s.tbs = o; o += 8 * TQ_N * s.PRB;
  s.pc = o; o += PC_N * s.PRB;
  s.part = o; o += 3u * (unsigned)nth * 8;
  s.bth0 = o; o += 2 * s.PRB;
  s.stage = o; o += (lwflux ? 3 : 2) * s.SSB;
  s.mbar = o; o += 16;
  s.total = o;
  return s;
}
__host__ __device__ constexpr int floordiv2(int x) { return x >= 0 ? x / 2 : -((-x + 1) / 2); }
// the directions k0+LO .. k0+HI of a shared-memory row (k0 even): the aligned pairs that cover them, one 16-byte load each
template <int LO, int HI>
struct Run {
  static constexpr int P0 = floordiv2(LO), P1 = floordiv2(HI), N = P1 - P0 + 1;
  double v[2 * N];
  __device__ __forceinline__ void load(const char* sm, unsigned base) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const Vd<2> a = lds<2>(sm, base + (unsigned)((P0 + j) * 16));
      v[2 * j] = a.v[0]; v[2 * j + 1] = a.v[1];
    }
  }
  __device__ __forceinline__ double at(int off) const { return v[off - 2 * P0]; }
};
#ifndef SW_MINB
#define SW_MINB 2   // k_sweep wants ~224 registers: capped at 168 for 3 CTAs per SM it spills and runs 30 % slower
#endif
#ifndef SW_CARRY
#define SW_CARRY 1   // carry the direction interpolation of the IP1 / IM1 rows to the step where they are IP / IM
#endif

template <int TA, int NPT, bool LWFLUX>
__global__ void __launch_bounds__(sw_threads(TA, NPT), SW_MINB) k_sweep(ImplDev d, long long p0, long long np) {
  extern __shared__ __align__(16) char sm[];
  typedef Vd<2> V;
  constexpr int A = TA, NPR = TA / 2, NSD = geo_nsd(TA), NS = 2 * NSD + 1, HR = dp_hr(TA), HC = dp_hc(TA);
  constexpr int NACT = NPT * NPR;
  constexpr unsigned PSTB = sw_pst(TA, NPT) * 8, PSTCB = sw_pstc(TA, NPT) * 8;
  constexpr SwSmem L = sw_smem(TA, NPT, LWFLUX);
  constexpr unsigned RSB = L.RSB, PSB = L.PSB, PRB = L.PRB;
  constexpr int GA = -geo_sh(TA, 0, 0), GB = geo_sh(TA, 0, 2);   // |K1W - K| and |K2W - K|; K11W, K21W are one further out
  const int F = c_dc.F;
  const bool ard = c_dc.iphys == 1;
  const unsigned smb = (unsigned)__cvta_generic_to_shared(sm);
  const int t = threadIdx.x;
  const bool act = t < NACT;                       // lanes beyond NPT*NANG/2 shadow the last thread and never store
  const int ta = act ? t : NACT - 1;
  const int pt = (ta % (2 * NPT)) >> 1;            // grid point of the CTA
  const int k0 = 2 * (2 * (ta / (2 * NPT)) + (ta & 1));   // first direction of the pair
  unsigned po = (unsigned)pt * 8u;
  unsigned me_r = (unsigned)pt * PSTB + (unsigned)(k0 + HR) * 8u;     // own pair inside a ring row
  unsigned me_c = (unsigned)pt * PSTCB + (unsigned)(k0 + HC) * 8u;    // own pair inside an interaction plane
  unsigned me_s = L.stage + (unsigned)t * 8u;                          // own landing slots: [array][element of the pair][thread]
  asm volatile("" : "+r"(po), "+r"(me_r), "+r"(me_c), "+r"(me_s));
  // ---- points
  const long long pbase = p0 + (long long)blockIdx.x * NPT;
  const long long plast = p0 + np - 1;
  long long pq = pbase + pt;
  const bool pvalid = pq <= plast;
  if (!pvalid) pq = plast;
  const bool dost = act && pvalid;
  const long long n = d.npts;
  const double* s = d.scr;
  const size_t P = (size_t)d.P;
  const size_t rstr = P * A;
  const long long pc_ = pq / d.P;
  const int pi_ = (int)(pq - pc_ * d.P);
  const size_t off_hi = (size_t)pi_ + P * A * F * (size_t)pc_ + P * (size_t)k0;
  const double* src_hi = d.f.fl1 + off_hi;
  const double* src_lo = src_hi;
  int mlo = 0;
  if (d.lo_on) {
    src_lo = d.fl_lo + (size_t)pi_ + P * A * d.lo_F * (size_t)pc_ + P * (size_t)k0;
    mlo = d.Fr;
  }
  double* dst = d.f.fl1 + off_hi;
  const double* src_in = d.fldin + off_hi;
  const double* src_xl = d.f.xllws + off_hi;
  // loader of the per-(point, frequency) scalars: thread (q, p8) of the first TQ_N*NPT
  const int tq_q = t / NPT, tq_p = t - tq_q * NPT;
  const bool tq_on = t < TQ_N * NPT;
  const double* tq_g = d.tbg + (size_t)(tq_on ? tq_q : 0) * F * n + min(pbase + tq_p, plast);
  const unsigned tq_s = L.tbs + (unsigned)(tq_on ? tq_q : 0) * PRB + (unsigned)tq_p * 8u;

  // ---- prologue: per-point constants, barrier
  if (t == 0) mbar_init(smb + L.mbar, 2u * blockDim.x);
  if (t < NPT) {
    const long long qp = min(pbase + t, plast);
    double* pcv = reinterpret_cast<double*>(sm + L.pc) + t;
    constexpr int PS = sw_nptp(NPT);
    const double wd = d.f.wdwave[qp], ci = d.f.cicover[qp], dep = d.f.depth[qp];
    double snw, csw;
    sincos(wd, &snw, &csw);
    double enhfr = dmax(0.75 * dep * s[S_AKMEAN * n + qp], 0.5);
    enhfr = 1.0 + (5.5 / enhfr) * (1.0 - .833 * enhfr) * exp(-1.25 * enhfr);
    const int mijq = (int)s[S_MIJ * n + qp];
    const bool seticeq = c_dc.licerun && c_dc.lmaskice && ci > c_dc.cithrsh;
    pcv[PC_FAC * PS] = s[S_FAC * n + qp];
    pcv[PC_ENH * PS] = enhfr;
    pcv[PC_USFMDELT * PS] = s[S_USFM * n + qp] * c_dc.delt;
    pcv[PC_SDSBK * PS] = (c_dc.lbiwbk && dep < 50.0) ? s[S_SDS * n + qp] : 0.0;
    pcv[PC_RTAIL * PS] = 1.0 / d.tbg[((size_t)TQ_TAIL * F + (mijq - 1)) * n + qp];
    pcv[PC_FLMC * PS] = (1. - 0.9 * dmin(ci, 0.99)) * c_dc.flmin;
    pcv[PC_ICEADD * PS] = seticeq ? dmax(c_dc.EPSMIN, 1.0 - ci) * c_dc.flmin : 0.0;
    pcv[PC_ICEFREE * PS] = seticeq ? 0.0 : 1.0;
    pcv[PC_SNW * PS] = snw;
    pcv[PC_CSW * PS] = csw;
    pcv[PC_BETA * PS] = (c_dc.licerun && c_dc.lciscal) ? 1.0 - ci : 1.0;
  }
  __syncthreads();
  V flm, iaw;    // FLM(k) = FLMC*max(0,cos(TH(k)-WDWAVE))**2 (implsch.F90:236-247); SETICE's noise floor likewise
  {
    const double snw = lds1(sm, L.pc + PC_SNW * PRB + po), csw = lds1(sm, L.pc + PC_CSW * PRB + po);
    const double flmc = lds1(sm, L.pc + PC_FLMC * PRB + po), ia = lds1(sm, L.pc + PC_ICEADD * PRB + po);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double cw2 = sq(dmax(0.0, c_dc.COSTH[k0 + i] * csw + c_dc.SINTH[k0 + i] * snw));
      flm.v[i] = flmc * cw2; iaw.v[i] = ia * cw2;
    }
  }
  const int mij = (int)s[S_MIJ * n + pq];
  const bool setice = c_dc.licerun && c_dc.lmaskice;
  const double delt = c_dc.delt, deltm = 1.0 / delt, delt5 = c_dc.ximp * delt;
  const double tmp03 = 1.0 / (c_dc.SDSBR * c_dc.MICHE);
  const int MLSTHG = c_dc.MLSTHG;
  const bool lssource = c_dc.lcflx && c_dc.lwvflx_snl, lsspre = c_dc.lcflx && !c_dc.lwvflx_snl;

  double asl[7][2], afl[7][2];
#pragma unroll
  for (int x = 0; x < 7; ++x)
#pragma unroll
    for (int i = 0; i < 2; ++i) { asl[x][i] = 0.0; afl[x][i] = 0.0; }
  // direction interpolations D+ / D- [kh][bin] of the rows that are IP / IM of the next centre frequency
  double dpc[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, dmc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  V bs1, bs2;    // saturation values of the own bins of rows st-3, st-4
  bs1.v[0] = bs1.v[1] = bs2.v[0] = bs2.v[1] = 0.0;
  V fmij, a_philf, a_ts, a_tu, a_e1, a_e2, a_el;
#pragma unroll
  for (int i = 0; i < 2; ++i) { fmij.v[i] = 0.0; a_philf.v[i] = 0.0; a_ts.v[i] = 0.0; a_tu.v[i] = 0.0; a_e1.v[i] = 0.0; a_e2.v[i] = 0.0; a_el.v[i] = 0.0; }

  auto step = [&](const int st) {
    // ================= phase A =================
    const int rnew = st + 5, rfin = st - 4, rb = st - 2, rtb = st - 1;
    const unsigned par = (unsigned)st & 1u;
    const bool dia = st >= 0 && st < MLSTHG;
    const bool fin = rfin >= 0 && rfin < F;
    const bool sat = ard && rb >= 0 && rb < F;
    const unsigned p3 = (unsigned)(st + 6) % 3u;
    if (rnew >= 0 && rnew < F) {
      const double* g = (rnew < mlo ? src_lo : src_hi) + (size_t)rnew * rstr;
      cp_async<8>(smb + me_s, g); cp_async<8>(smb + me_s + L.SSB / 2, g + P);
    }
    if (fin) {
      const double* g = src_in + (size_t)rfin * rstr;
      cp_async<8>(smb + me_s + L.SSB, g); cp_async<8>(smb + me_s + L.SSB + L.SSB / 2, g + P);
      if (LWFLUX) { const double* gx = src_xl + (size_t)rfin * rstr; cp_async<8>(smb + me_s + 2 * L.SSB, gx); cp_async<8>(smb + me_s + 2 * L.SSB + L.SSB / 2, gx + P); }
    }
    if (rtb >= 0 && rtb < F && tq_on) cp_async<8>(smb + tq_s + (unsigned)((rtb & 7) * TQ_N) * PRB, tq_g + (size_t)rtb * n);
    mbar_arrive_cp_async(smb + L.mbar);
    V bnew;
    bnew.v[0] = 0.0; bnew.v[1] = 0.0;
    if (st >= -1 && st < MLSTHG) {
      // direction interpolation of the rows that join the quadruplets at this centre frequency (IP1, IM1; at st = -1: IP, IM of
      // the first one), both KH families: D+[kh] = CL11 F(k + s0) + ACL1 F(k + s1), D-[kh] = CL21 F(k + s2) + ACL2 F(k + s3)
      double dpn[2][2], dmn[2][2];
      {
        const unsigned bP = L.ring + (unsigned)c_dc.NLS2[st + 1][0] * RSB + me_r;
        const unsigned bM = L.ring + (unsigned)c_dc.NLS2[st + 1][1] * RSB + me_r;
        Run<-GA - 1, GA + 2> rp;
        Run<GB, GB + 2> rm0;
        Run<-GB - 1, -GB + 1> rm1;
        rp.load(sm, bP); rm0.load(sm, bM); rm1.load(sm, bM);
        const double cl11 = c_dc.NLD[0], acl1 = c_dc.NLD[1], cl21 = c_dc.NLD[2], acl2 = c_dc.NLD[3];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          dpn[0][i] = fma(acl1, rp.at(i - GA - 1), cl11 * rp.at(i - GA));
          dpn[1][i] = fma(acl1, rp.at(i + GA + 1), cl11 * rp.at(i + GA));
          dmn[0][i] = fma(acl2, rm0.at(i + GB + 1), cl21 * rm0.at(i + GB));
          dmn[1][i] = fma(acl2, rm1.at(i - GB - 1), cl21 * rm1.at(i - GB));
        }
      }
      if (dia) {
        const int MC0 = st;
        const double* Wc = c_dc.NLW[MC0];
#if !SW_CARRY
        {   // no carry: the IP / IM rows are interpolated again
          const unsigned bP = L.ring + (unsigned)c_dc.NLSLOT[MC0][1] * RSB + me_r;
          const unsigned bM = L.ring + (unsigned)c_dc.NLSLOT[MC0][3] * RSB + me_r;
          Run<-GA - 1, GA + 2> rp;
          Run<GB, GB + 2> rm0;
          Run<-GB - 1, -GB + 1> rm1;
          rp.load(sm, bP); rm0.load(sm, bM); rm1.load(sm, bM);
          const double cl11 = c_dc.NLD[0], acl1 = c_dc.NLD[1], cl21 = c_dc.NLD[2], acl2 = c_dc.NLD[3];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            dpc[0][i] = fma(acl1, rp.at(i - GA - 1), cl11 * rp.at(i - GA));
            dpc[1][i] = fma(acl1, rp.at(i + GA + 1), cl11 * rp.at(i + GA));
            dmc[0][i] = fma(acl2, rm0.at(i + GB + 1), cl21 * rm0.at(i + GB));
            dmc[1][i] = fma(acl2, rm1.at(i - GB - 1), cl21 * rm1.at(i - GB));
          }
        }
#endif
        const unsigned cb = L.cur + par * 6u * PSB + me_c;
        const V fc = lds<2>(sm, L.ring + (unsigned)c_dc.NLSLOT[MC0][0] * RSB + me_r);
        const double enh = d.enh ? __ldg(d.enh + (size_t)MC0 * n + pq) : lds1(sm, L.pc + PC_ENH * PRB + po);   // ISNONLIN = 1, 2: k_enh's plane
        const double ftemp = c_dc.AF11[MC0] * enh;
        const double r0 = c_dc.RNLCOEF[MC0][0];
        V fij, fcen, fcd1, fcd2;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          fij.v[i] = fc.v[i] * r0;
          fcen.v[i] = ftemp * fij.v[i];
          fcd1.v[i] = c_dc.DAL1 * fcen.v[i];
          fcd2.v[i] = c_dc.DAL2 * fcen.v[i];
        }
        V csl, cfl;
        csl.v[0] = csl.v[1] = cfl.v[0] = cfl.v[1] = 0.0;
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {
          V vad, vdp, vdm;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const double sap = fma(Wc[1], dpn[kh][i], Wc[0] * dpc[kh][i]);
            const double sam = fma(Wc[3], dmn[kh][i], Wc[2] * dmc[kh][i]);
            double fad1 = fij.v[i] * (sap + sam);
            const double sap2 = 2.0 * sap;
            const double fad2 = fma(-sap2, sam, fad1);
            fad1 = fad1 + fad2;
            vad.v[i] = fad2 * fcen.v[i];
            csl.v[i] += vad.v[i];
            cfl.v[i] = fma(fad1, ftemp, cfl.v[i]);
            vdp.v[i] = fma(-2.0, sam, fij.v[i]) * fcd1.v[i];
            vdm.v[i] = (fij.v[i] - sap2) * fcd2.v[i];
          }
          if (act) {
            sts<2>(sm, cb + (0 + kh) * PSB, vad);
            sts<2>(sm, cb + (2 + kh) * PSB, vdp);
            sts<2>(sm, cb + (4 + kh) * PSB, vdm);
            if (k0 < HC) {
              sts<2>(sm, cb + (0 + kh) * PSB + (unsigned)A * 8u, vad);
              sts<2>(sm, cb + (2 + kh) * PSB + (unsigned)A * 8u, vdp);
              sts<2>(sm, cb + (4 + kh) * PSB + (unsigned)A * 8u, vdm);
            }
            if (k0 >= A - HC) {
              sts<2>(sm, cb + (0 + kh) * PSB - (unsigned)A * 8u, vad);
              sts<2>(sm, cb + (2 + kh) * PSB - (unsigned)A * 8u, vdp);
              sts<2>(sm, cb + (4 + kh) * PSB - (unsigned)A * 8u, vdm);
            }
          }
        }
        const double c2 = c_dc.RNLC2[MC0];
#pragma unroll
        for (int i = 0; i < 2; ++i) { asl[4][i] = fma(-c2, csl.v[i], asl[4][i]); afl[4][i] = fma(-c2, cfl.v[i], afl[4][i]); }
      }
#if SW_CARRY
#pragma unroll
      for (int kh = 0; kh < 2; ++kh)
#pragma unroll
        for (int i = 0; i < 2; ++i) { dpc[kh][i] = dpn[kh][i]; dmc[kh][i] = dmn[kh][i]; }
#endif
    }
    mbar_arrive(smb + L.mbar);
    // saturation spectrum of row rb (sdissip_ard.F90:142-160): both windows of the pair from HR+1 aligned 16-byte loads
    if (sat) {
      const unsigned wb = L.ring + slot9(rb) * RSB + me_r - (unsigned)HR * 8u;
      V b;
      b.v[0] = 0.0; b.v[1] = 0.0;
#pragma unroll
      for (int jp = 0; jp <= HR; ++jp) {
        const V f = lds<2>(sm, wb + jp * 16);
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int x = 2 * jp + e - (HR - NSD) - i;
            if (x >= 0 && x < NS) b.v[i] += c_dc.SATW1[x] * f.v[e];
          }
      }
      const double fs = lds1(sm, L.tbs + (unsigned)(((rb & 7) * TQ_N + TQ_FACSAT)) * PRB + po);
      bnew.v[0] = b.v[0] * fs; bnew.v[1] = b.v[1] * fs;
      // BTH0 = max over direction: one partial per thread, reduced by four lanes per point in phase B of the next step
      *reinterpret_cast<double*>(sm + L.part + p3 * (unsigned)(blockDim.x * 8u) + (unsigned)t * 8u) = dmax(bnew.v[0], bnew.v[1]);
    }
    mbar_wait(smb + L.mbar, (unsigned)(st + 5) & 1u);
    // ================= phase B =================
    V tot_sl, tot_fl;
    {
      double tap[2] = {0.0, 0.0}, tpp[2] = {0.0, 0.0}, tam[2] = {0.0, 0.0}, tmm[2] = {0.0, 0.0};
      if (dia) {
        // gather of the quadruplet contributions of MC (snonlin.F90:253-308) through the inverse shifts, direction interpolation first
        const unsigned cb = L.cur + par * 6u * PSB + me_c;
        const double cl11 = c_dc.NLD[0], acl1 = c_dc.NLD[1], cl21 = c_dc.NLD[2], acl2 = c_dc.NLD[3];
        const double cl11s = c_dc.NLD[4], acl1s = c_dc.NLD[5], cl21s = c_dc.NLD[6], acl2s = c_dc.NLD[7];
        {   // KH = 1: products at k - s: K1W, K11W families at +GA, +GA+1; K2W, K21W families at -GB, -GB-1
          Run<GA, GA + 2> ap, pp;
          Run<-GB - 1, -GB + 1> am, mm;
          ap.load(sm, cb + 0 * PSB); pp.load(sm, cb + 2 * PSB); am.load(sm, cb + 0 * PSB); mm.load(sm, cb + 4 * PSB);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            tap[i] = fma(acl1, ap.at(i + GA + 1), cl11 * ap.at(i + GA));
            tpp[i] = fma(acl1s, pp.at(i + GA + 1), cl11s * pp.at(i + GA));
            tam[i] = fma(acl2, am.at(i - GB - 1), cl21 * am.at(i - GB));
            tmm[i] = fma(acl2s, mm.at(i - GB - 1), cl21s * mm.at(i - GB));
          }
        }
        {   // KH = 2: mirrored
          Run<-GA - 1, -GA + 1> ap, pp;
          Run<GB, GB + 2> am, mm;
          ap.load(sm, cb + 1 * PSB); pp.load(sm, cb + 3 * PSB); am.load(sm, cb + 1 * PSB); mm.load(sm, cb + 5 * PSB);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            tap[i] = fma(acl1, ap.at(i - GA - 1), fma(cl11, ap.at(i - GA), tap[i]));
            tpp[i] = fma(acl1s, pp.at(i - GA - 1), fma(cl11s, pp.at(i - GA), tpp[i]));
            tam[i] = fma(acl2, am.at(i + GB + 1), fma(cl21, am.at(i + GB), tam[i]));
            tmm[i] = fma(acl2s, mm.at(i + GB + 1), fma(cl21s, mm.at(i + GB), tmm[i]));
          }
        }
      }
      const double* Wc = c_dc.NLW[dia ? st : 0];
      // slide the window of pending rows: row st-3 -> slot 0, ..., row st+3 (first contribution) -> slot 6
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        tot_sl.v[i] = fma(Wc[6], tam[i], asl[0][i]);    tot_fl.v[i] = fma(Wc[10], tmm[i], afl[0][i]);
        asl[0][i] = fma(Wc[7], tam[i], asl[1][i]);      afl[0][i] = fma(Wc[11], tmm[i], afl[1][i]);
        asl[1][i] = asl[2][i];                          afl[1][i] = afl[2][i];
        asl[2][i] = asl[3][i];                          afl[2][i] = afl[3][i];
        asl[3][i] = asl[4][i];                          afl[3][i] = afl[4][i];
        asl[4][i] = asl[5][i];                          afl[4][i] = afl[5][i];
        asl[5][i] = fma(Wc[4], tap[i], asl[6][i]);      afl[5][i] = fma(Wc[8], tpp[i], afl[6][i]);
        asl[6][i] = Wc[5] * tap[i];                     afl[6][i] = Wc[9] * tpp[i];
      }
    }
    // finish row rfin (implsch.F90:276-395 for these bins)
    if (fin) {
      const int r = rfin;
      const unsigned tq = L.tbs + (unsigned)((r & 7) * TQ_N) * PRB + po;
      const V fold = lds<2>(sm, L.ring + slot9(r) * RSB + me_r);
      V xI;
      xI.v[0] = lds1(sm, me_s + L.SSB); xI.v[1] = lds1(sm, me_s + L.SSB + L.SSB / 2);
      const double usfm = lds1(sm, L.pc + PC_USFMDELT * PRB + po), sdsbk = lds1(sm, L.pc + PC_SDSBK * PRB + po);
      const double tsbo = lds1(sm, tq + TQ_SBO * PRB), tcinv = lds1(sm, tq + TQ_CINV * PRB), ttail = lds1(sm, tq + TQ_TAIL * PRB);
      const double tstf = lds1(sm, tq + TQ_STF * PRB), rtail = lds1(sm, L.pc + PC_RTAIL * PRB + po);
      const double beta = c_dc.lciscal ? lds1(sm, L.pc + PC_BETA * PRB + po) : 1.0;
      V dd;
      if (ard) {
        const double b0 = lds1(sm, L.bth0 + (par ^ 1u) * PRB + po);
        const double ssdsc2_sig = c_dc.SSDSC2 * c_dc.ZPIFR[r];
        const double d0 = ssdsc2_sig * c_dc.SSDSC6 * sq(dmax(0., b0 * tmp03 - c_dc.SSDSC4));
#pragma unroll
        for (int i = 0; i < 2; ++i) dd.v[i] = d0 + ssdsc2_sig * (1. - c_dc.SSDSC6) * sq(dmax(0., bs2.v[i] * tmp03 - c_dc.SSDSC4));
      } else { const double dj = lds1(sm, tq + TQ_JAN * PRB); dd.v[0] = dj; dd.v[1] = dj; }
      const double cofrm4 = c_dc.COFRM4[r], flmax = c_dc.FLMAX[r];
      double rr = (r + 1 > mij) ? 0.0 : c_dc.RHOWG_DFIM[r];      // RHOWGDFTH of frcutindex.F90:99-108
      if (r + 1 == mij && mij != F) rr = 0.5 * rr;
      V xL;
      if (LWFLUX) { xL.v[0] = lds1(sm, me_s + 2 * L.SSB); xL.v[1] = lds1(sm, me_s + 2 * L.SSB + L.SSB / 2); }
      V fnv;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double f0 = fold.v[i];
        double fldv = xI.v[i];
        double slv = fldv * f0;
        slv = slv + dd.v[i] * f0; fldv = fldv + dd.v[i];
        slv = slv + tot_sl.v[i]; fldv = fldv + tot_fl.v[i];
        double ssource = 0.0;
        if (lssource) ssource = div_fast(slv, dmax(1.0 - delt5 * fldv, 1.0));
        if (r < c_dc.Fr) { slv = slv - sdsbk * f0; fldv = fldv - sdsbk; }            // SDIWBK (0 where it does not apply)
        if (c_dc.lciscal) { slv = slv * beta; fldv = fldv * beta; }                  // LCISCAL (implsch.F90:315-325)
        slv = slv + tsbo * f0; fldv = fldv + tsbo;                                   // SDICE3 + SBOTTOM (plane is 0 where neither applies)
        const double gtemp1 = dmax(1.0 - delt5 * fldv, 1.0);
        const double gtemp2 = div_fast(delt * slv, gtemp1);
        const double flhab = dmin(fabs(gtemp2), usfm * cofrm4);
        double fn = f0 + copysign(flhab, gtemp2);
        fn = dmax(fn, flm.v[i]);
        ssource = ssource + deltm * dmin(flmax - fn, 0.0);
        fn = dmin(fn, flmax);
        a_philf.v[i] += ssource * rr;
        a_ts.v[i] += ssource * (tcinv * rr);
        if (LWFLUX) {
          const double xf = (xL.v[i] != 0.0) ? fn : 0.0;
          a_e1.v[i] += c_dc.DFIM[r] * xf; a_e2.v[i] += c_dc.DFIMOFR[r] * xf;
          if (r == F - 1) a_el.v[i] += xf;
        }
        if (r == mij - 1) fmij.v[i] = fn;
        if (r > mij - 1) fn = dmax((ttail * rtail) * fmij.v[i], flm.v[i]);
        fnv.v[i] = fn;
      }
      if (setice) {
        const double icefree = lds1(sm, L.pc + PC_ICEFREE * PRB + po);
#pragma unroll
        for (int i = 0; i < 2; ++i) fnv.v[i] = fnv.v[i] * icefree + iaw.v[i];
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        a_tu.v[i] += tstf * fnv.v[i];
        if (r == c_dc.NFRE_ODD - 1) {
          const double cst = 2.0 * c_dc.DELTH * c_dc.ZPI * c_dc.ZPI * c_dc.ZPI / c_dc.G * p4(c_dc.FR[c_dc.NFRE_ODD - 1]);
          a_tu.v[i] += cst * fnv.v[i];
        }
      }
      if (dost) { double* o = dst + (size_t)r * rstr; o[0] = fnv.v[0]; o[P] = fnv.v[1]; }
    }
    bs2 = bs1; bs1 = bnew;
    if (rnew >= 0 && rnew < F) {   // depth-limited (+ floored at NFRE) row st+5 -> ring slot of row st-4 (last read just above)
      V xF;
      xF.v[0] = lds1(sm, me_s); xF.v[1] = lds1(sm, me_s + L.SSB / 2);
      const double fac = lds1(sm, L.pc + PC_FAC * PRB + po);
      V v;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        v.v[i] = dmax(xF.v[i] * fac, c_dc.EPSMIN);
        if (rnew == F - 1) v.v[i] = dmax(v.v[i], flm.v[i]);
      }
      if (act) {
        const unsigned rbse = L.ring + slot9(rnew) * RSB + me_r;
        sts<2>(sm, rbse, v);
        if (k0 < HR) sts<2>(sm, rbse + (unsigned)A * 8u, v);
        if (k0 >= A - HR) sts<2>(sm, rbse - (unsigned)A * 8u, v);
      }
    }
    if (ard && t < 4 * NPT && rb - 1 >= 0 && rb - 1 < F) {   // BTH0 of row st-3 from the partial maxima of the previous step
      const int qp = t >> 2, j = t & 3;
      const double* pp = reinterpret_cast<const double*>(sm + L.part + ((unsigned)(st + 5) % 3u) * (unsigned)(blockDim.x * 8u)) + 2 * qp;
      double mx = 0.0;
#pragma unroll
      for (int x = 0; x < (NPR + 3) / 4; ++x) {   // pair e = j + 4 x of the point sits at thread ((e / 2) * NPT + qp) * 2 + (e & 1)
        const int e = j + 4 * x;
        if (e < NPR) mx = dmax(mx, pp[(e >> 1) * (2 * NPT) + (e & 1)]);
      }
      // only the first 4*NPT lanes are here: the member mask must name exactly them
      constexpr unsigned RM = (4 * NPT >= 32) ? 0xffffffffu : ((1u << ((4 * NPT) & 31)) - 1u);
      mx = dmax(mx, __shfl_xor_sync(RM, mx, 1));
      mx = dmax(mx, __shfl_xor_sync(RM, mx, 2));
      if (j == 0) reinterpret_cast<double*>(sm + L.bth0 + par * PRB)[qp] = mx;
    }
  };

#pragma unroll 1
  for (int st = -5; st < MLSTHG; ++st) step(st);
  // ---- per-point sums over direction, then the scalar closures (one thread per point)
  __syncthreads();
  constexpr int PS = sw_nptp(NPT);
  double* red = reinterpret_cast<double*>(sm + L.ring);     // red[q][k][pt]: 8 planes of A*PS doubles in the (now free) ring area
  static_assert(8u * A * PS * 8u <= ST_RING * L.RSB, "reduction planes must fit the ring");
  if (act) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = k0 + i;
      const double sinth = c_dc.SINTH[k], costh = c_dc.COSTH[k];
      red[(0 * A + k) * PS + pt] = a_philf.v[i];
      red[(1 * A + k) * PS + pt] = sinth * a_ts.v[i];
      red[(2 * A + k) * PS + pt] = costh * a_ts.v[i];
      red[(3 * A + k) * PS + pt] = a_tu.v[i] * sinth;
      red[(4 * A + k) * PS + pt] = a_tu.v[i] * costh;
      if (LWFLUX) { red[(5 * A + k) * PS + pt] = a_e1.v[i]; red[(6 * A + k) * PS + pt] = a_e2.v[i]; red[(7 * A + k) * PS + pt] = a_el.v[i]; }
    }
  }
  __syncthreads();
  double* qs = reinterpret_cast<double*>(sm + L.cur);        // [8][PS] reduced sums
  if (t < 8 * NPT) {
    const int x = t / NPT, p8 = t - x * NPT;
    double v = 0.0;
    if (x < 5 || LWFLUX) for (int kk = 0; kk < A; ++kk) v += red[(x * A + kk) * PS + p8];
    qs[x * PS + p8] = v;
  }
  __syncthreads();
  if (t < NPT && pbase + t <= plast)
    stencil_closure<LWFLUX, sw_nptp(NPT)>(d, pbase + t, qs + t, reinterpret_cast<const double*>(sm + L.pc) + t);
}

// =========================================================================================================
// k_sweep_ws: k_sweep with the step split over two warp roles (producer / consumer warp groups of one CTA).
// What bounds the one-role kernels is neither a pipe nor a memory level but latency at low occupancy: the sweep needs ~220
// registers per thread to be scheduled well (56 of them the pending SNONLIN rows), i.e. 8 warps per SM; capped at 168 registers
// for 12 warps it spills and rematerialises, and either way it runs at the pace of the 10-warp k_stencil_dp (45 ms at O640), with
// 13 % of the warp time in the per-step CTA barrier.  Here
//   * the PRODUCER warp group forms what only reads the spectrum: the direction interpolation of the rows entering the
//     quadruplets, the DIA products of the centre frequency -> interaction planes, the centre bin's own share, and the saturation
//     window of SDISSIP_ARD with its maximum over direction.  Its persistent state is the carried interpolation (16 registers);
//   * the CONSUMER warp group gathers the products into the pending rows (56 registers), finishes one row per step (implicit
//     update, WNFLUXES sums, tail, ice, Stokes), stores it, and feeds the next row of the spectrum into the ring.
//     (Moving the gather to the producers balances the instruction counts but costs a producer barrier in mid-step: 38.1 vs 35.6 ms.)
//   setmaxnreg moves registers from the producers (SW_RP) to the consumers (SW_RC): 2 CTAs = 16 warps per SM, each role with the
//   registers it needs.  The roles run up to two steps apart: the hand-over slots are double-buffered and passed with two mbarriers
//   per buffer (full: producers -> consumers; empty: consumers -> producers — the consumers' "empty" of step s also publishes ring
//   row s+5, the newest row the producers read at step s+2) and so are the interaction planes; the four producer warps meet once
//   per step at a named barrier for the direction maximum.  No CTA-wide barrier inside the sweep; a role that waits for the other
//   one polls with a back-off (a spinning producer group took 16 % of the issue slots).
// =========================================================================================================
#ifndef SW_RP
#define SW_RP 96     // registers per producer thread
#endif
#ifndef SW_RUNPTR
#define SW_RUNPTR 1  // global row addresses of the consumers: 0 rebuilt from the row index (36.1 ms), 1 one running offset (35.6), 2 running pointers (spills: 44.2)
#endif
#ifndef SW_RC
#define SW_RC 160    // registers per consumer thread   (SW_RP + SW_RC <= 256: 2 CTAs of 2 x 128 threads per SM)
#endif
struct WsSmem { unsigned ring, cur, tbs, pc, part, bth0, hand, stage, mbar, total, RSB, PSB, SSB, PRB; };
__host__ __device__ constexpr WsSmem ws_smem(int A, int npt, bool lwflux) {
  WsSmem s{};
  const int nth = sw_threads(A, npt);     // threads per role
  s.RSB = (unsigned)sw_pst(A, npt) * npt * 8;
  s.PSB = (unsigned)sw_pstc(A, npt) * npt * 8;
  s.SSB = (unsigned)nth * 16;
  s.PRB = (unsigned)sw_nptp(npt) * 8;
  unsigned o = 0;
  s.ring = o; o += ST_RING * s.RSB;
  s.cur = o; o += 2 * 6 * s.PSB;
  s.tbs = o; o += 8 * TQ_N * s.PRB;
  s.pc = o; o += PC_N * s.PRB;
  s.part = o; o += 2u * (unsigned)nth * 8;
  s.bth0 = o; o += 2 * s.PRB;
  s.hand = o; o += 2 * 3 * s.SSB;         // per buffer, one pair per thread: saturation values, centre share of SL, centre share of FLD
  s.stage = o; o += (lwflux ? 3 : 2) * s.SSB;
  s.mbar = o; o += 32;                    // full[0], full[1], empty[0], empty[1]
  s.total = o;
  return s;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
template <int ID, int NT> __device__ __forceinline__ void named_bar_sync() { asm volatile("bar.sync %0, %1;\n" ::"n"(ID), "n"(NT) : "memory"); }

template <int TA, int NPT, bool LWFLUX>
__global__ void __launch_bounds__(2 * sw_threads(TA, NPT), 2) k_sweep_ws(ImplDev d, long long p0, long long np) {
  extern __shared__ __align__(16) char sm[];
  typedef Vd<2> V;
  constexpr int A = TA, NPR = TA / 2, NSD = geo_nsd(TA), NS = 2 * NSD + 1, HR = dp_hr(TA), HC = dp_hc(TA);
  constexpr int NACT = NPT * NPR, NTHR = sw_threads(TA, NPT);
  constexpr unsigned PSTB = sw_pst(TA, NPT) * 8, PSTCB = sw_pstc(TA, NPT) * 8;
  constexpr WsSmem L = ws_smem(TA, NPT, LWFLUX);
  constexpr unsigned RSB = L.RSB, PSB = L.PSB, PRB = L.PRB;
  constexpr int GA = -geo_sh(TA, 0, 0), GB = geo_sh(TA, 0, 2);
  static_assert(NTHR == 128, "k_sweep_ws: one warp group (4 warps) per role");
  const int F = c_dc.F;
  const bool ard = c_dc.iphys == 1;
  const unsigned smb = (unsigned)__cvta_generic_to_shared(sm);
  const int t = threadIdx.x;
  const bool consumer = t >= NTHR;
  const int tr = consumer ? t - NTHR : t;          // thread of the role
  const bool act = tr < NACT;                      // lanes beyond NPT*NANG/2 shadow the last thread and never store
  const int ta = act ? tr : NACT - 1;
  const int pt = (ta % (2 * NPT)) >> 1;            // grid point of the CTA
  const int k0 = 2 * (2 * (ta / (2 * NPT)) + (ta & 1));   // first direction of the pair
  unsigned po = (unsigned)pt * 8u;
  unsigned me_r = (unsigned)pt * PSTB + (unsigned)(k0 + HR) * 8u;
  unsigned me_c = (unsigned)pt * PSTCB + (unsigned)(k0 + HC) * 8u;
  unsigned me_h = L.hand + (unsigned)tr * 16u;     // own hand-over slot
  asm volatile("" : "+r"(po), "+r"(me_r), "+r"(me_c), "+r"(me_h));
  const long long pbase = p0 + (long long)blockIdx.x * NPT;
  const long long plast = p0 + np - 1;
  const long long n = d.npts;
  const double* s = d.scr;
  const int MLSTHG = c_dc.MLSTHG;
  const unsigned bar_full = smb + L.mbar, bar_empty = smb + L.mbar + 16;

  // ---- prologue: barriers, per-point constants
  if (t == 0) { mbar_init(bar_full, NTHR); mbar_init(bar_full + 8, NTHR); mbar_init(bar_empty, NTHR); mbar_init(bar_empty + 8, NTHR); }
  if (t < NPT) {
    const long long qp = min(pbase + t, plast);
    double* pcv = reinterpret_cast<double*>(sm + L.pc) + t;
    constexpr int PS = sw_nptp(NPT);
    const double wd = d.f.wdwave[qp], ci = d.f.cicover[qp], dep = d.f.depth[qp];
    double snw, csw;
    sincos(wd, &snw, &csw);
    double enhfr = dmax(0.75 * dep * s[S_AKMEAN * n + qp], 0.5);
    enhfr = 1.0 + (5.5 / enhfr) * (1.0 - .833 * enhfr) * exp(-1.25 * enhfr);
    const int mijq = (int)s[S_MIJ * n + qp];
    const bool seticeq = c_dc.licerun && c_dc.lmaskice && ci > c_dc.cithrsh;
    pcv[PC_FAC * PS] = s[S_FAC * n + qp];
    pcv[PC_ENH * PS] = enhfr;
    pcv[PC_USFMDELT * PS] = s[S_USFM * n + qp] * c_dc.delt;
    pcv[PC_SDSBK * PS] = (c_dc.lbiwbk && dep < 50.0) ? s[S_SDS * n + qp] : 0.0;
    pcv[PC_RTAIL * PS] = 1.0 / d.tbg[((size_t)TQ_TAIL * F + (mijq - 1)) * n + qp];
    pcv[PC_FLMC * PS] = (1. - 0.9 * dmin(ci, 0.99)) * c_dc.flmin;
    pcv[PC_ICEADD * PS] = seticeq ? dmax(c_dc.EPSMIN, 1.0 - ci) * c_dc.flmin : 0.0;
    pcv[PC_ICEFREE * PS] = seticeq ? 0.0 : 1.0;
    pcv[PC_SNW * PS] = snw;
    pcv[PC_CSW * PS] = csw;
    pcv[PC_BETA * PS] = (c_dc.licerun && c_dc.lciscal) ? 1.0 - ci : 1.0;
  }
  __syncthreads();
  V a_philf, a_ts, a_tu, a_e1, a_e2, a_el;     // consumer accumulators (reduced over direction after the sweep)
#pragma unroll
  for (int i = 0; i < 2; ++i) { a_philf.v[i] = 0.0; a_ts.v[i] = 0.0; a_tu.v[i] = 0.0; a_e1.v[i] = 0.0; a_e2.v[i] = 0.0; a_el.v[i] = 0.0; }

  if (!consumer) {
    // ================================================= PRODUCER =================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(SW_RP));
    double dpc[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, dmc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll 1
    for (int st = -5; st < MLSTHG; ++st) {
      const unsigned sidx = (unsigned)(st + 5), b = sidx & 1u, u = sidx >> 1;
      // buffer b is free and ring rows up to st+3 are in place once the consumers have finished step st-2
      if (sidx >= 2u) mbar_wait_backoff(bar_empty + 8u * b, (u - 1u) & 1u);
      const unsigned hb = me_h + b * 3u * L.SSB;
      if (st >= -1) {
        // direction interpolation of the rows that join the quadruplets at this centre frequency (IP1, IM1; at st = -1: IP, IM of
        // the first one), both KH families
        double dpn[2][2], dmn[2][2];
        {
          const unsigned bP = L.ring + (unsigned)c_dc.NLS2[st + 1][0] * RSB + me_r;
          const unsigned bM = L.ring + (unsigned)c_dc.NLS2[st + 1][1] * RSB + me_r;
          Run<-GA - 1, GA + 2> rp;
          Run<GB, GB + 2> rm0;
          Run<-GB - 1, -GB + 1> rm1;
          rp.load(sm, bP); rm0.load(sm, bM); rm1.load(sm, bM);
          const double cl11 = c_dc.NLD[0], acl1 = c_dc.NLD[1], cl21 = c_dc.NLD[2], acl2 = c_dc.NLD[3];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            dpn[0][i] = fma(acl1, rp.at(i - GA - 1), cl11 * rp.at(i - GA));
            dpn[1][i] = fma(acl1, rp.at(i + GA + 1), cl11 * rp.at(i + GA));
            dmn[0][i] = fma(acl2, rm0.at(i + GB + 1), cl21 * rm0.at(i + GB));
            dmn[1][i] = fma(acl2, rm1.at(i - GB - 1), cl21 * rm1.at(i - GB));
          }
        }
        if (st >= 0) {
          const int MC0 = st;
          const double* Wc = c_dc.NLW[MC0];
          const unsigned cb = L.cur + b * 6u * PSB + me_c;
          const V fc = lds<2>(sm, L.ring + (unsigned)c_dc.NLSLOT[MC0][0] * RSB + me_r);
          const double enh = d.enh ? __ldg(d.enh + (size_t)MC0 * n + min(pbase + pt, plast)) : lds1(sm, L.pc + PC_ENH * PRB + po);   // ISNONLIN = 1, 2
          const double ftemp = c_dc.AF11[MC0] * enh;
          const double r0 = c_dc.RNLCOEF[MC0][0];
          V fij, fcen, fcd1, fcd2;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            fij.v[i] = fc.v[i] * r0;
            fcen.v[i] = ftemp * fij.v[i];
            fcd1.v[i] = c_dc.DAL1 * fcen.v[i];
            fcd2.v[i] = c_dc.DAL2 * fcen.v[i];
          }
          V csl, cfl;
          csl.v[0] = csl.v[1] = cfl.v[0] = cfl.v[1] = 0.0;
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {
            V vad, vdp, vdm;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const double sap = fma(Wc[1], dpn[kh][i], Wc[0] * dpc[kh][i]);
              const double sam = fma(Wc[3], dmn[kh][i], Wc[2] * dmc[kh][i]);
              double fad1 = fij.v[i] * (sap + sam);
              const double sap2 = 2.0 * sap;
              const double fad2 = fma(-sap2, sam, fad1);
              fad1 = fad1 + fad2;
              vad.v[i] = fad2 * fcen.v[i];
              csl.v[i] += vad.v[i];
              cfl.v[i] = fma(fad1, ftemp, cfl.v[i]);
              vdp.v[i] = fma(-2.0, sam, fij.v[i]) * fcd1.v[i];
              vdm.v[i] = (fij.v[i] - sap2) * fcd2.v[i];
            }
            if (act) {
              sts<2>(sm, cb + (0 + kh) * PSB, vad);
              sts<2>(sm, cb + (2 + kh) * PSB, vdp);
              sts<2>(sm, cb + (4 + kh) * PSB, vdm);
              if (k0 < HC) {
                sts<2>(sm, cb + (0 + kh) * PSB + (unsigned)A * 8u, vad);
                sts<2>(sm, cb + (2 + kh) * PSB + (unsigned)A * 8u, vdp);
                sts<2>(sm, cb + (4 + kh) * PSB + (unsigned)A * 8u, vdm);
              }
              if (k0 >= A - HC) {
                sts<2>(sm, cb + (0 + kh) * PSB - (unsigned)A * 8u, vad);
                sts<2>(sm, cb + (2 + kh) * PSB - (unsigned)A * 8u, vdp);
                sts<2>(sm, cb + (4 + kh) * PSB - (unsigned)A * 8u, vdm);
              }
            }
          }
          // the centre bin's own share (-2 AD, -2 DELAD; snonlin.F90:253-262) is handed to the consumer thread of the same bins
          const double c2 = c_dc.RNLC2[MC0];
          V ccs, ccf;
#pragma unroll
          for (int i = 0; i < 2; ++i) { ccs.v[i] = -c2 * csl.v[i]; ccf.v[i] = -c2 * cfl.v[i]; }
          sts<2>(sm, hb + 1u * L.SSB, ccs);
          sts<2>(sm, hb + 2u * L.SSB, ccf);
        }
#pragma unroll
        for (int kh = 0; kh < 2; ++kh)
#pragma unroll
          for (int i = 0; i < 2; ++i) { dpc[kh][i] = dpn[kh][i]; dmc[kh][i] = dmn[kh][i]; }
      }
      // saturation spectrum of the row the consumers finish at this step (sdissip_ard.F90:142-160); two partial sums per bin
      // halve the dependent FMA chain
      const int rs = st - 4;
      const bool sat = ard && rs >= 0 && rs < F;
      if (sat) {
        const unsigned wb = L.ring + slot9(rs) * RSB + me_r - (unsigned)HR * 8u;
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int jp = 0; jp <= HR; ++jp) {
          const V f = lds<2>(sm, wb + jp * 16);
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int x = 2 * jp + e - (HR - NSD) - i;
              if (x >= 0 && x < NS) acc[jp & 1][i] = fma(c_dc.SATW1[x], f.v[e], acc[jp & 1][i]);
            }
        }
        V bsat;
        bsat.v[0] = acc[0][0] + acc[1][0]; bsat.v[1] = acc[0][1] + acc[1][1];
        const double fs = lds1(sm, L.tbs + (unsigned)(((rs & 7) * TQ_N + TQ_FACSAT)) * PRB + po);
        bsat.v[0] *= fs; bsat.v[1] *= fs;
        sts<2>(sm, hb, bsat);
        *reinterpret_cast<double*>(sm + L.part + b * (unsigned)(NTHR * 8) + (unsigned)tr * 8u) = dmax(bsat.v[0], bsat.v[1]);
      }
      named_bar_sync<1, NTHR>();          // the partial maxima of all four producer warps are written
      if (sat && tr < 4 * NPT) {          // BTH0 = max over direction: four lanes per grid point
        const int qp = tr >> 2, j = tr & 3;
        const double* pp = reinterpret_cast<const double*>(sm + L.part + b * (unsigned)(NTHR * 8)) + 2 * qp;
        double mx = 0.0;
#pragma unroll
        for (int x = 0; x < (NPR + 3) / 4; ++x) {   // pair e = j + 4 x of the point sits at thread ((e / 2) * NPT + qp) * 2 + (e & 1)
          const int e = j + 4 * x;
          if (e < NPR) mx = dmax(mx, pp[(e >> 1) * (2 * NPT) + (e & 1)]);
        }
        // lanes of this warp that take part: 4 NPT consecutive role threads, possibly over two warps (NPT > 8)
        const int lane0 = tr & ~31, nin = min(32, 4 * NPT - lane0);
        const unsigned RM = nin >= 32 ? 0xffffffffu : ((1u << nin) - 1u);
        mx = dmax(mx, __shfl_xor_sync(RM, mx, 1));
        mx = dmax(mx, __shfl_xor_sync(RM, mx, 2));
        if (j == 0) reinterpret_cast<double*>(sm + L.bth0 + b * PRB)[qp] = mx;
      }
      mbar_arrive(bar_full + 8u * b);
    }
  } else {
    // ================================================= CONSUMER =================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(SW_RC));
    unsigned me_s = L.stage + (unsigned)tr * 8u;       // own landing slots: [array][element of the pair][thread], 8 bytes each, so that
                                                       // a warp's cp.async writes are 256 contiguous bytes (16-byte slots cost 4x the wavefronts)
    asm volatile("" : "+r"(me_s));
    long long pq = pbase + pt;
    const bool pvalid = pq <= plast;
    if (!pvalid) pq = plast;
    const bool dost = act && pvalid;
    const size_t P = (size_t)d.P;
    const size_t rstr = P * A;
    const long long pc_ = pq / d.P;
    const int pi_ = (int)(pq - pc_ * d.P);
    const size_t off_hi = (size_t)pi_ + P * A * F * (size_t)pc_ + P * (size_t)k0;
    const double* src_hi = d.f.fl1 + off_hi;
    const double* src_lo = src_hi;
    int mlo = 0;
    if (d.lo_on) {
      src_lo = d.fl_lo + (size_t)pi_ + P * A * d.lo_F * (size_t)pc_ + P * (size_t)k0;
      mlo = d.Fr;
    }
    double* dst = d.f.fl1 + off_hi;
    const double* src_in = d.fldin + off_hi;
    const double* src_xl = d.f.xllws + off_hi;
    // loader of the per-(point, frequency) scalars: thread (q, p8) of the first TQ_N*NPT consumer threads
    const int tq_q = tr / NPT, tq_p = tr - tq_q * NPT;
    const bool tq_on = tr < TQ_N * NPT;
    const double* tq_g = d.tbg + (size_t)(tq_on ? tq_q : 0) * F * n + min(pbase + tq_p, plast);
    const unsigned tq_s = L.tbs + (unsigned)(tq_on ? tq_q : 0) * PRB + (unsigned)tq_p * 8u;
    V flm, iaw;    // FLM(k) = FLMC*max(0,cos(TH(k)-WDWAVE))**2 (implsch.F90:236-247); SETICE's noise floor likewise
    {
      const double snw = lds1(sm, L.pc + PC_SNW * PRB + po), csw = lds1(sm, L.pc + PC_CSW * PRB + po);
      const double flmc = lds1(sm, L.pc + PC_FLMC * PRB + po), ia = lds1(sm, L.pc + PC_ICEADD * PRB + po);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double cw2 = sq(dmax(0.0, c_dc.COSTH[k0 + i] * csw + c_dc.SINTH[k0 + i] * snw));
        flm.v[i] = flmc * cw2; iaw.v[i] = ia * cw2;
      }
    }
    const int mij = (int)s[S_MIJ * n + pq];
    const bool setice = c_dc.licerun && c_dc.lmaskice;
    const double delt = c_dc.delt, deltm = 1.0 / delt, delt5 = c_dc.ximp * delt;
    const double tmp03 = 1.0 / (c_dc.SDSBR * c_dc.MICHE);
    const bool lssource = c_dc.lcflx && c_dc.lwvflx_snl, lsspre = c_dc.lcflx && !c_dc.lwvflx_snl;
    // pending SNONLIN sums of rows st-4 .. st+2 (row st+3 receives its first contribution, MC -> MC+3, at this step)
    double asl[7][2], afl[7][2];
#pragma unroll
    for (int x = 0; x < 7; ++x)
#pragma unroll
      for (int i = 0; i < 2; ++i) { asl[x][i] = 0.0; afl[x][i] = 0.0; }
    V fmij;
    fmij.v[0] = fmij.v[1] = 0.0;
    // the global rows of a step through ONE running offset (+ one row per step) and a running pointer for the scalar planes:
    // rebuilt from the row index they cost ~100 of the consumers' 596 instructions per step in 64-bit multiplies.
    // roff: element offset of row st-4 (negative until row 0: not dereferenced before); row st+5 = roff + 9 rows.
#if SW_RUNPTR == 1
    ptrdiff_t roff = -9 * (ptrdiff_t)rstr;
    const double* src_lo9 = src_lo + 9 * rstr;
    const double* src_hi9 = src_hi + 9 * rstr;
    const double* g_tq = tq_g - 6 * (ptrdiff_t)n;
#elif SW_RUNPTR == 2
    const double* g_in = src_in - 9 * (ptrdiff_t)rstr;
    const double* g_xl = src_xl - 9 * (ptrdiff_t)rstr;
    double* g_dst = dst - 9 * (ptrdiff_t)rstr;
    const double* g_lo = src_lo;
    const double* g_hi = src_hi;
    const double* g_tq = tq_g - 6 * (ptrdiff_t)n;
#endif

#pragma unroll 1
    for (int st = -5; st < MLSTHG; ++st) {
      const unsigned sidx = (unsigned)(st + 5), b = sidx & 1u, u = sidx >> 1;
      const int rnew = st + 5, rfin = st - 4, rtb = st - 1;
      const bool dia = st >= 0;
      const bool fin = rfin >= 0 && rfin < F;
      // row loads of this step: group 1 = wind-input (and XLLWS) row of the row that is finished, group 2 = FL1 row st+5 for the
      // ring and the per-(point, frequency) scalars of row st-1
#if SW_RUNPTR == 1
      if (fin) {
        const double* g = src_in + roff;
        cp_async<8>(smb + me_s + L.SSB, g); cp_async<8>(smb + me_s + L.SSB + L.SSB / 2, g + P);
        if (LWFLUX) { const double* gx = src_xl + roff; cp_async<8>(smb + me_s + 2 * L.SSB, gx); cp_async<8>(smb + me_s + 2 * L.SSB + L.SSB / 2, gx + P); }
      }
      cp_async_commit();
      if (rnew < F) {
        const double* g = (rnew < mlo ? src_lo9 : src_hi9) + roff;
        cp_async<8>(smb + me_s, g); cp_async<8>(smb + me_s + L.SSB / 2, g + P);
      }
      if (rtb >= 0 && rtb < F && tq_on) cp_async<8>(smb + tq_s + (unsigned)((rtb & 7) * TQ_N) * PRB, g_tq);
      cp_async_commit();
      double* const o_row = dst + roff;
      roff += (ptrdiff_t)rstr; g_tq += n;
#elif SW_RUNPTR == 2
      if (fin) {
        cp_async<8>(smb + me_s + L.SSB, g_in); cp_async<8>(smb + me_s + L.SSB + L.SSB / 2, g_in + P);
        if (LWFLUX) { cp_async<8>(smb + me_s + 2 * L.SSB, g_xl); cp_async<8>(smb + me_s + 2 * L.SSB + L.SSB / 2, g_xl + P); }
      }
      cp_async_commit();
      if (rnew < F) {
        const double* g = rnew < mlo ? g_lo : g_hi;
        cp_async<8>(smb + me_s, g); cp_async<8>(smb + me_s + L.SSB / 2, g + P);
      }
      if (rtb >= 0 && rtb < F && tq_on) cp_async<8>(smb + tq_s + (unsigned)((rtb & 7) * TQ_N) * PRB, g_tq);
      cp_async_commit();
      double* const o_row = g_dst;
      g_in += rstr; g_xl += rstr; g_dst += rstr; g_lo += rstr; g_hi += rstr; g_tq += n;
#else
      if (fin) {
        const double* g = src_in + (size_t)rfin * rstr;
        cp_async<8>(smb + me_s + L.SSB, g); cp_async<8>(smb + me_s + L.SSB + L.SSB / 2, g + P);
        if (LWFLUX) { const double* gx = src_xl + (size_t)rfin * rstr; cp_async<8>(smb + me_s + 2 * L.SSB, gx); cp_async<8>(smb + me_s + 2 * L.SSB + L.SSB / 2, gx + P); }
      }
      cp_async_commit();
      if (rnew < F) {
        const double* g = (rnew < mlo ? src_lo : src_hi) + (size_t)rnew * rstr;
        cp_async<8>(smb + me_s, g); cp_async<8>(smb + me_s + L.SSB / 2, g + P);
      }
      if (rtb >= 0 && rtb < F && tq_on) cp_async<8>(smb + tq_s + (unsigned)((rtb & 7) * TQ_N) * PRB, tq_g + (size_t)rtb * n);
      cp_async_commit();
      double* const o_row = dst + (size_t)rfin * rstr;
#endif
      mbar_wait_backoff(bar_full + 8u * b, u & 1u);
      const unsigned hb = me_h + b * 3u * L.SSB;
      V tot_sl, tot_fl;
      {
        double tap[2] = {0.0, 0.0}, tpp[2] = {0.0, 0.0}, tam[2] = {0.0, 0.0}, tmm[2] = {0.0, 0.0};
        if (dia) {
          // gather of the quadruplet contributions of MC (snonlin.F90:253-308) through the inverse shifts, direction interpolation first
          const unsigned cb = L.cur + b * 6u * PSB + me_c;
          const double cl11 = c_dc.NLD[0], acl1 = c_dc.NLD[1], cl21 = c_dc.NLD[2], acl2 = c_dc.NLD[3];
          const double cl11s = c_dc.NLD[4], acl1s = c_dc.NLD[5], cl21s = c_dc.NLD[6], acl2s = c_dc.NLD[7];
          {   // KH = 1
            Run<GA, GA + 2> ap, pp;
            Run<-GB - 1, -GB + 1> am, mm;
            ap.load(sm, cb + 0 * PSB); pp.load(sm, cb + 2 * PSB); am.load(sm, cb + 0 * PSB); mm.load(sm, cb + 4 * PSB);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              tap[i] = fma(acl1, ap.at(i + GA + 1), cl11 * ap.at(i + GA));
              tpp[i] = fma(acl1s, pp.at(i + GA + 1), cl11s * pp.at(i + GA));
              tam[i] = fma(acl2, am.at(i - GB - 1), cl21 * am.at(i - GB));
              tmm[i] = fma(acl2s, mm.at(i - GB - 1), cl21s * mm.at(i - GB));
            }
          }
          {   // KH = 2: mirrored
            Run<-GA - 1, -GA + 1> ap, pp;
            Run<GB, GB + 2> am, mm;
            ap.load(sm, cb + 1 * PSB); pp.load(sm, cb + 3 * PSB); am.load(sm, cb + 1 * PSB); mm.load(sm, cb + 5 * PSB);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              tap[i] = fma(acl1, ap.at(i - GA - 1), fma(cl11, ap.at(i - GA), tap[i]));
              tpp[i] = fma(acl1s, pp.at(i - GA - 1), fma(cl11s, pp.at(i - GA), tpp[i]));
              tam[i] = fma(acl2, am.at(i + GB + 1), fma(cl21, am.at(i + GB), tam[i]));
              tmm[i] = fma(acl2s, mm.at(i + GB + 1), fma(cl21s, mm.at(i + GB), tmm[i]));
            }
          }
          const V ccs = lds<2>(sm, hb + 1u * L.SSB), ccf = lds<2>(sm, hb + 2u * L.SSB);    // the centre bin's own share
#pragma unroll
          for (int i = 0; i < 2; ++i) { asl[4][i] += ccs.v[i]; afl[4][i] += ccf.v[i]; }
        }
        const double* Wc = c_dc.NLW[dia ? st : 0];
        // slide the window of pending rows: row st-3 -> slot 0, ..., row st+3 (first contribution) -> slot 6
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          tot_sl.v[i] = fma(Wc[6], tam[i], asl[0][i]);    tot_fl.v[i] = fma(Wc[10], tmm[i], afl[0][i]);
          asl[0][i] = fma(Wc[7], tam[i], asl[1][i]);      afl[0][i] = fma(Wc[11], tmm[i], afl[1][i]);
          asl[1][i] = asl[2][i];                          afl[1][i] = afl[2][i];
          asl[2][i] = asl[3][i];                          afl[2][i] = afl[3][i];
          asl[3][i] = asl[4][i];                          afl[3][i] = afl[4][i];
          asl[4][i] = asl[5][i];                          afl[4][i] = afl[5][i];
          asl[5][i] = fma(Wc[4], tap[i], asl[6][i]);      afl[5][i] = fma(Wc[8], tpp[i], afl[6][i]);
          asl[6][i] = Wc[5] * tap[i];                     afl[6][i] = Wc[9] * tpp[i];
        }
      }
      // finish row rfin (implsch.F90:276-395 for these bins)
      if (fin) {
        const int r = rfin;
        const unsigned tq = L.tbs + (unsigned)((r & 7) * TQ_N) * PRB + po;
        const V fold = lds<2>(sm, L.ring + slot9(r) * RSB + me_r);
        cp_async_wait<1>();                              // the wind-input row of this step has landed (own slot)
        V xI;
        xI.v[0] = lds1(sm, me_s + L.SSB); xI.v[1] = lds1(sm, me_s + L.SSB + L.SSB / 2);
        const double usfm = lds1(sm, L.pc + PC_USFMDELT * PRB + po), sdsbk = lds1(sm, L.pc + PC_SDSBK * PRB + po);
        const double tsbo = lds1(sm, tq + TQ_SBO * PRB), tcinv = lds1(sm, tq + TQ_CINV * PRB), ttail = lds1(sm, tq + TQ_TAIL * PRB);
        const double tstf = lds1(sm, tq + TQ_STF * PRB), rtail = lds1(sm, L.pc + PC_RTAIL * PRB + po);
        const double beta = c_dc.lciscal ? lds1(sm, L.pc + PC_BETA * PRB + po) : 1.0;
        V dd;
        if (ard) {
          const double b0 = lds1(sm, L.bth0 + b * PRB + po);
          const V bs = lds<2>(sm, hb);
          const double ssdsc2_sig = c_dc.SSDSC2 * c_dc.ZPIFR[r];
          const double d0 = ssdsc2_sig * c_dc.SSDSC6 * sq(dmax(0., b0 * tmp03 - c_dc.SSDSC4));
#pragma unroll
          for (int i = 0; i < 2; ++i) dd.v[i] = d0 + ssdsc2_sig * (1. - c_dc.SSDSC6) * sq(dmax(0., bs.v[i] * tmp03 - c_dc.SSDSC4));
        } else { const double dj = lds1(sm, tq + TQ_JAN * PRB); dd.v[0] = dj; dd.v[1] = dj; }
        const double cofrm4 = c_dc.COFRM4[r], flmax = c_dc.FLMAX[r];
        double rr = (r + 1 > mij) ? 0.0 : c_dc.RHOWG_DFIM[r];      // RHOWGDFTH of frcutindex.F90:99-108
        if (r + 1 == mij && mij != F) rr = 0.5 * rr;
        V xL;
        if (LWFLUX) { xL.v[0] = lds1(sm, me_s + 2 * L.SSB); xL.v[1] = lds1(sm, me_s + 2 * L.SSB + L.SSB / 2); }
        V fnv;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const double f0 = fold.v[i];
          double fldv = xI.v[i];
          double slv = fldv * f0;
          slv = slv + dd.v[i] * f0; fldv = fldv + dd.v[i];
          slv = slv + tot_sl.v[i]; fldv = fldv + tot_fl.v[i];
          double ssource = 0.0;
          if (lssource) ssource = div_fast(slv, dmax(1.0 - delt5 * fldv, 1.0));
          if (r < c_dc.Fr) { slv = slv - sdsbk * f0; fldv = fldv - sdsbk; }            // SDIWBK (0 where it does not apply)
          if (c_dc.lciscal) { slv = slv * beta; fldv = fldv * beta; }                  // LCISCAL (implsch.F90:315-325)
          slv = slv + tsbo * f0; fldv = fldv + tsbo;                                   // SDICE3 + SBOTTOM (plane is 0 where neither applies)
          const double gtemp1 = dmax(1.0 - delt5 * fldv, 1.0);
          const double gtemp2 = div_fast(delt * slv, gtemp1);
          const double flhab = dmin(fabs(gtemp2), usfm * cofrm4);
          double fn = f0 + copysign(flhab, gtemp2);
          fn = dmax(fn, flm.v[i]);
          ssource = ssource + deltm * dmin(flmax - fn, 0.0);
          fn = dmin(fn, flmax);
          a_philf.v[i] += ssource * rr;
          a_ts.v[i] += ssource * (tcinv * rr);
          if (LWFLUX) {
            const double xf = (xL.v[i] != 0.0) ? fn : 0.0;
            a_e1.v[i] += c_dc.DFIM[r] * xf; a_e2.v[i] += c_dc.DFIMOFR[r] * xf;
            if (r == F - 1) a_el.v[i] += xf;
          }
          if (r == mij - 1) fmij.v[i] = fn;
          if (r > mij - 1) fn = dmax((ttail * rtail) * fmij.v[i], flm.v[i]);
          fnv.v[i] = fn;
        }
        if (setice) {
          const double icefree = lds1(sm, L.pc + PC_ICEFREE * PRB + po);
#pragma unroll
          for (int i = 0; i < 2; ++i) fnv.v[i] = fnv.v[i] * icefree + iaw.v[i];
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          a_tu.v[i] += tstf * fnv.v[i];
          if (r == c_dc.NFRE_ODD - 1) {
            const double cst = 2.0 * c_dc.DELTH * c_dc.ZPI * c_dc.ZPI * c_dc.ZPI / c_dc.G * p4(c_dc.FR[c_dc.NFRE_ODD - 1]);
            a_tu.v[i] += cst * fnv.v[i];
          }
        }
        if (dost) { o_row[0] = fnv.v[0]; o_row[P] = fnv.v[1]; }
      }
      cp_async_wait<0>();          // FL1 row st+5 (own slot) and, for the loader lanes, the scalars of row st-1
      if (rnew < F) {              // depth-limited (+ floored at NFRE) row st+5 -> ring slot of row st-4 (last read just above)
        V xF;
        xF.v[0] = lds1(sm, me_s); xF.v[1] = lds1(sm, me_s + L.SSB / 2);
        const double fac = lds1(sm, L.pc + PC_FAC * PRB + po);
        V v;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          v.v[i] = dmax(xF.v[i] * fac, c_dc.EPSMIN);
          if (rnew == F - 1) v.v[i] = dmax(v.v[i], flm.v[i]);
        }
        if (act) {
          const unsigned rbse = L.ring + slot9(rnew) * RSB + me_r;
          sts<2>(sm, rbse, v);
          if (k0 < HR) sts<2>(sm, rbse + (unsigned)A * 8u, v);
          if (k0 >= A - HR) sts<2>(sm, rbse - (unsigned)A * 8u, v);
        }
      }
      mbar_arrive(bar_empty + 8u * b);
    }
  }
  // ---- per-point sums over direction, then the scalar closures (one thread per point)
  __syncthreads();
  constexpr int PS = sw_nptp(NPT);
  double* red = reinterpret_cast<double*>(sm + L.ring);     // red[q][k][pt]: 8 planes of A*PS doubles in the (now free) ring area
  static_assert(8u * A * PS * 8u <= ST_RING * L.RSB, "reduction planes must fit the ring");
  if (consumer && act) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = k0 + i;
      const double sinth = c_dc.SINTH[k], costh = c_dc.COSTH[k];
      red[(0 * A + k) * PS + pt] = a_philf.v[i];
      red[(1 * A + k) * PS + pt] = sinth * a_ts.v[i];
      red[(2 * A + k) * PS + pt] = costh * a_ts.v[i];
      red[(3 * A + k) * PS + pt] = a_tu.v[i] * sinth;
      red[(4 * A + k) * PS + pt] = a_tu.v[i] * costh;
      if (LWFLUX) { red[(5 * A + k) * PS + pt] = a_e1.v[i]; red[(6 * A + k) * PS + pt] = a_e2.v[i]; red[(7 * A + k) * PS + pt] = a_el.v[i]; }
    }
  }
  __syncthreads();
  double* qs = reinterpret_cast<double*>(sm + L.cur);        // [8][PS] reduced sums
  if (t < 8 * NPT) {
    const int x = t / NPT, p8 = t - x * NPT;
    double v = 0.0;
    if (x < 5 || LWFLUX) for (int kk = 0; kk < A; ++kk) v += red[(x * A + kk) * PS + p8];
    qs[x * PS + p8] = v;
  }
  __syncthreads();
  if (t < NPT && pbase + t <= plast)
    stencil_closure<LWFLUX, sw_nptp(NPT)>(d, pbase + t, qs + t, reinterpret_cast<const double*>(sm + L.pc) + t);
}

template <int TA, int NPT, bool LW>
static int launch_sweep_ws(const ImplDev& d, long long p0, long long np, cudaStream_t st) {
  constexpr WsSmem L = ws_smem(TA, NPT, LW);
  static_assert(L.total <= 113 * 1024, "k_sweep_ws shared memory");
  static_assert(8 * NPT <= 2 * sw_threads(TA, NPT) && TQ_N * NPT <= sw_threads(TA, NPT) && 4 * NPT <= sw_threads(TA, NPT), "k_sweep_ws: too few threads for the per-point loops");
  static bool attr_done = false;
  if (!attr_done) {
    EW_CUDA_CHECK(cudaFuncSetAttribute(k_sweep_ws<TA, NPT, LW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    EW_CUDA_CHECK(cudaFuncSetAttribute(k_sweep_ws<TA, NPT, LW>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_done = true;
  }
  k_sweep_ws<TA, NPT, LW><<<(unsigned)((np + NPT - 1) / NPT), 2 * sw_threads(TA, NPT), L.total, st>>>(d, p0, np);
  return 0;
}


template <int TA, int NPT, bool LW>
static int launch_sweep(const ImplDev& d, long long p0, long long np, cudaStream_t st) {
  constexpr SwSmem L = sw_smem(TA, NPT, LW);
  static_assert(L.total <= 110 * 1024, "k_sweep shared memory");
  static_assert(8 * NPT <= sw_threads(TA, NPT) && TQ_N * NPT <= sw_threads(TA, NPT) && 4 * NPT <= 32, "k_sweep: too few threads for the per-point loops");
  static bool attr_done = false;
  if (!attr_done) {
    EW_CUDA_CHECK(cudaFuncSetAttribute(k_sweep<TA, NPT, LW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    EW_CUDA_CHECK(cudaFuncSetAttribute(k_sweep<TA, NPT, LW>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_done = true;
  }
  k_sweep<TA, NPT, LW><<<(unsigned)((np + NPT - 1) / NPT), sw_threads(TA, NPT), L.total, st>>>(d, p0, np);
  return 0;
}

template <int TA>
static bool geo_matches(const ImplDev& d, int iphys, int nsdsnth) {
  if (d.A != TA) return false;
  for (int kh = 0; kh < 2; ++kh) for (int q = 0; q < 4; ++q) if (d.dsb[kh][q] != geo_sh(TA, kh, q) * ST_NPT * 8) return false;
  return iphys != 1 || nsdsnth == geo_nsd(TA);
}

// =========================================================================================================
// k_enh: ENH(IJ,MC) of ISNONLIN = 1, 2 (snonlin.F90:138-163), one thread per grid point, written to ImplDev::enh [MLSTHG][npts]
// between k_point (SDEPTHLIM's factor) and the frequency sweep, which reads it instead of its per-point constant.
//   ISNONLIN = 1: Janssen & Onorato's shallow-water transfer function of the centre wavenumber (transf.F90:62-110), clipped to
//                 [0.1, 10]; above NFRE the wavenumber continues as deep water on the geometric frequency axis.
//   ISNONLIN = 2: the same with the directional-width correction of the mean-flow term (transf_snl.F90:57-100); the relative
//                 spectral width XNU and the angular width SIG_TH come from PEAK_ANG (peak_ang.F90:72-175) on the depth-limited,
//                 floored spectrum SNONLIN sees (two more reads of FL1 per point: an optional mode).
// =========================================================================================================
__device__ double transf_enh(double xk, double dep) {
  const double EPS = 0.0001, DKMAX = 40.0;
  if (!(dep < c_dc.bathymax && dep > 0.0)) return 1.0;
  const double x = xk * dep;
  if (x > DKMAX) return 1.0;
  const double t0 = tanh(x), om = sqrt(c_dc.G * xk * t0), c0 = om / xk;
  const double vg = x < EPS ? c0 : 0.5 * c0 * (1.0 + 2.0 * x / sinh(2.0 * x));
  const double t2 = t0 * t0;
  const double a = t0 - x * (1.0 - t2);
  const double dvg = a * a + 4.0 * (x * x) * t2 * (1.0 - t2);
  const double xnl1 = (9.0 * (t2 * t2) - 10.0 * t2 + 9.0) / (8.0 * (t2 * t0));
  const double b = 2.0 * vg - 0.5 * c0;
  const double xnl2 = (b * b / (c_dc.G * dep - vg * vg) + 1.0) / x;
  const double xnl = xnl1 - xnl2;
  return xnl * xnl / (dvg * ((t2 * t2) * (t2 * t2)));
}
__device__ double transf_snl_enh(double xk0, double dep, double xnu, double sig_th) {
  const double EPS = 0.0001, DKMAX = 40.0, XKDMIN = 0.75, TMIN = 0.1, TMAX = 10.0;
  if (!(dep < c_dc.bathymax && dep > 0.0)) return 1.0;
  double x = xk0 * dep;
  if (x > DKMAX) return 1.0;
  const double xk = dmax(xk0, XKDMIN / dep);
  x = xk * dep;
  const double t0 = tanh(x), t2 = t0 * t0, om = sqrt(c_dc.G * xk * t0), c0 = om / xk, cs2 = c_dc.G * dep;
  const double vg = x < EPS ? c0 : 0.5 * c0 * (1.0 + 2.0 * x / sinh(2.0 * x));
  const double vg2 = vg * vg;
  const double a = t0 - x * (1. - t2);
  const double dvg = a * a + 4.0 * (x * x) * t2 * (1.0 - t2);
  const double xnl1 = (9.0 * (t2 * t2) - 10.0 * t2 + 9.0) / (8.0 * t2 * t0);
  const double b = 2.0 * vg - 0.5 * c0;
  const double xnl2 = (b * b / (c_dc.G * dep - vg2) + 1.0) / x;
  const double e = 2.0 * c0 + vg * (1.0 - t2);
  const double xnl4 = 1. / (4.0 * t0) * (e * e) / (cs2 - vg2);
  const double alp = (1. - vg2 / cs2) * (c0 * c0) / vg2;
  const double zfac = (sig_th * sig_th) / (sig_th * sig_th + alp * (xnu * xnu));
  const double xnl = xnl1 - xnl2 + zfac * xnl4;
  const double t4 = (t2 * t2) * (t2 * t2);
  return dmax(dmin(TMAX, xnl * xnl / (dvg * t4)), TMIN);
}

__global__ void __launch_bounds__(128) k_enh(ImplDev d, long long p0, long long np) {
  const long long p = p0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= p0 + np) return;
  const int A = c_dc.A, F = c_dc.F;
  const long long n = d.npts;
  const double dep = d.f.depth[p];
  double xnu = 0.0, sig_th = 0.0;
  if (d.isnonlin == 2) {
    const long long c = p / d.P;
    const int i = (int)(p - c * d.P);
    const size_t kstr = (size_t)d.P;
    const double* hi = d.f.fl1 + (size_t)i + (size_t)d.P * A * F * (size_t)c;
    const double* lo = hi;
    int mlo = 0;
    if (d.lo_on) { const int il = (p < d.nloc) ? i : 0; lo = d.fl_lo + (size_t)il + (size_t)d.P * A * d.lo_F * (size_t)c; mlo = d.Fr; }
    const double fac = d.scr[S_FAC * n + p];
    double snw, csw;
    sincos(d.f.wdwave[p], &snw, &csw);
    const double flmc = (1. - 0.9 * dmin(d.f.cicover[p], 0.99)) * c_dc.flmin;
    auto spec = [&](int m, int k) {      // FL1 as SNONLIN sees it: SDEPTHLIM applied, last frequency floored at FLM (sinflx.F90:126-129)
      double f = dmax(__ldg((m < mlo ? lo : hi) + ((size_t)m * A + k) * kstr) * fac, c_dc.EPSMIN);
      if (m == F - 1) f = dmax(f, flmc * sq(dmax(0.0, c_dc.COSTH[k] * csw + c_dc.SINTH[k] * snw)));
      return f;
    };
    const double ZEPS = 10.0 * 2.220446049250313e-16;
    const double frl = c_dc.FR[F - 1];
    const double DELT25 = c_dc.WETAIL * frl * c_dc.DELTH, COEF_FR = c_dc.WP1TAIL * c_dc.DELTH * (frl * frl);
    const double COEF_FR2 = 0.5 * c_dc.DELTH * (frl * frl * frl);     // WP2TAIL = 0.5 (yowfred.F90:54)
    const int NSH = 1 + (int)(log(1.5) / log(c_dc.FRATIO));
    double sum0 = ZEPS, sum1 = 0.0, sum2 = 0.0, temp = 0.0, xmax = 0.0;
    int mmax = 2;
    for (int m = 0; m < F; ++m) {
      temp = spec(m, 0);
      if (m >= 1 && m <= F - 2 && temp > xmax) { mmax = m + 1; xmax = temp; }
      for (int k = 1; k < A; ++k) {
        const double f = spec(m, k);
        temp = temp + f;
        if (m >= 1 && m <= F - 2 && f > xmax) { mmax = m + 1; xmax = f; }
      }
      sum0 = sum0 + temp * c_dc.DFIM[m]; sum1 = sum1 + temp * c_dc.DFIMFR[m]; sum2 = sum2 + temp * (c_dc.DFIM[m] * (c_dc.FR[m] * c_dc.FR[m]));
    }
    sum0 = sum0 + DELT25 * temp; sum1 = sum1 + COEF_FR * temp; sum2 = sum2 + COEF_FR2 * temp;
    xnu = sum0 > ZEPS ? sqrt(dmax(ZEPS, sum2 * sum0 / (sum1 * sum1) - 1.0)) : ZEPS;
    double s1 = ZEPS, s2 = 0.0, sum_s = 0.0, sum_c = ZEPS;
    for (int M = max(1, mmax - NSH); M <= min(F, mmax + NSH); ++M) {
      for (int k = 0; k < A; ++k) { const double f = spec(M - 1, k); sum_s = sum_s + c_dc.SINTH[k] * f; sum_c = sum_c + c_dc.COSTH[k] * f; }
      const double thmean = atan2(sum_s, sum_c);
      for (int k = 0; k < A; ++k) {
        const double f = spec(M - 1, k);
        s1 = s1 + f * c_dc.DFIM[M - 1];
        s2 = s2 + cos(c_dc.TH[k] - thmean) * f * c_dc.DFIM[M - 1];
      }
    }
    sig_th = s1 > ZEPS ? sqrt(2.0 * (1.0 - s2 / s1)) : 0.0;
  }
  for (int mc = 0; mc < c_dc.MLSTHG; ++mc) {
    double xk;
    if (mc < F) xk = d.f.wavnum[idx3(d, p, mc)];
    else {      // XK = GM1*(ZPIFR(NFRE)*FRATIO**(MC-NFRE))**2 (snonlin.F90:145, 157)
      double pw = 1.0;
      for (int j = 0; j < mc + 1 - F; ++j) pw = pw * c_dc.FRATIO;
      const double w = c_dc.ZPIFR[F - 1] * pw;
      xk = c_dc.GM1 * (w * w);
    }
    d.enh[(size_t)mc * n + p] = d.isnonlin == 1 ? dmax(dmin(10.0, transf_enh(xk, dep)), 0.1) : transf_snl_enh(xk, dep, xnu, sig_th);
  }
}

// =========================================================================================================
// k_ice: the per-(point, frequency) parts of SDICE1 and SDICE2 (sdice.F90:99-107), one thread per grid point, between k_point and the
// frequency sweep.
//   LCIWA1 (sdice1.F90:102-187): scattering by ice floes.  ALP = exp(CIDEAC(T, h)) / <D> * ZALPFACB with Kohout & Meylan's table
//     interpolated bilinearly in wave period and ice thickness and the mean floe size <D> of Dumont et al.'s fragmentation cascade
//     (a function of the ice cover).  The term is linear in the spectrum (SL += CICV*FLDICE*F, FLD += CICV*FLDICE) like SBOTTOM and
//     SDICE3, so it is added to their plane tbg[TQ_SBO].
//   LCIWA2 (sdice2.F90:97-113): friction under the ice, ALP = CDICWA*k^2*4*sqrt(max(EPSMIN, F*DFIM))*ZALPFACB depends on the bin:
//     ice2[m][p] = CICV*CDICWA*ZALPFACB*4*k^2*CGROUP, the sqrt is taken in the finish stage of k_stencil / k_stencil_dp.
// ice1 = [NICT*NICH] CIDEAC, then per frequency WT1 (F), IT (F, 0-based, as double), IT1 (F).
// =========================================================================================================
__global__ void __launch_bounds__(128) k_ice(ImplDev d, long long p0, long long np) {
  const long long p = p0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= p0 + np) return;
  const int F = c_dc.F;
  const size_t n = (size_t)d.npts;
  const double ci = d.f.cicover[p], cith = d.f.cithick[p];
  if (d.ice1) {
    const int NT = d.ice_nt, NH = d.ice_nh;
    const double* tabw = d.ice1 + (size_t)NT * NH;
    double dinv = 20.0;       // CIDMIN (sdice1.F90:127: only reached with CITH <= 0, where ALP = 0 anyway)
    int ih = 0, ih1 = 0;
    double wh = 1.0, wh1 = 0.0;
    if (cith > 0.0) {
      const double CIFRGL = 0.955, CIDMIN = 20.0, CIFRGMT = 2.0, A0 = 200.0, C0 = 300.0;
      const int maxicm = (int)(log(A0 / CIDMIN) / log(CIFRGMT));
      const double cidmax = A0 + C0 * ci;
      const int icm = min((int)(log(cidmax / CIDMIN) / log(CIFRGMT)), maxicm);
      double sn = 0.0, sd = 0.0, x = 1.0, fi = 1.0;      // x = (CIFRGMT**2*CIFRGL)**I, fi = CIFRGMT**I
      for (int i = 0; i <= icm; ++i) {
        sn = sn + x * cidmax / fi;
        sd = sd + x;
        x = x * (CIFRGMT * CIFRGMT * CIFRGL); fi = fi * CIFRGMT;
      }
      dinv = 1.0 / (sn / sd);
      ih = (int)floor((cith - d.ice_hmin) / d.ice_dh + 1) ;
      ih = max(1, min(ih, NH));
      ih1 = max(1, min(ih + 1, NH));
      wh1 = dmax(dmin(1.0, (cith - (d.ice_hmin + (ih - 1) * d.ice_dh)) / d.ice_dh), 0.0);
      wh = 1.0 - wh1;
      ih -= 1; ih1 -= 1;
    }
    for (int m = 0; m < F; ++m) {
      double alp = 0.0;
      if (cith > 0.0) {
        const double wt1 = tabw[m], wt = 1.0 - wt1;
        const int it = (int)tabw[F + m], it1 = (int)tabw[2 * F + m];
        const double c = wt * (wh * d.ice1[it + NT * ih] + wh1 * d.ice1[it + NT * ih1]) + wt1 * (wh * d.ice1[it1 + NT * ih] + wh1 * d.ice1[it1 + NT * ih1]);
        alp = exp(c) * dinv * c_dc.zalpfacb;
      }
      d.tbg[((size_t)TQ_SBO * F + m) * n + p] += ci * (-alp * d.f.cgroup[idx3(d, p, m)]);
    }
  }
  if (d.ice2) {
    for (int m = 0; m < F; ++m) {
      const size_t o3 = idx3(d, p, m);
      const double wk = d.f.wavnum[o3];
      d.ice2[(size_t)m * n + p] = ci * ((c_dc.cdicwa * (wk * wk) * 4.0 * c_dc.zalpfacb) * d.f.cgroup[o3]);
    }
  }
  if (!d.nemo) return;
  const NemoDev nd = *d.nemo;
  // LWNEMOCOUIBR (icebreak_modify_attenuation.F90:82-94): where the ice is broken (IBRMEM <= ZIBRW_THRSH) SDICE3's ALPFAC is 1/ZALPFACX
  // instead of ZALPFACX: k_point put the ZALPFACX coefficient on SBOTTOM's plane, the difference is added here.
  double alpfac = c_dc.zalpfacx;
  if (nd.ibr_on && nd.f.ibrmem[p] <= c_dc.zibrw_thrsh) alpfac = 1.0 / c_dc.zalpfacx;
  const double cith125 = (c_dc.lciwa3 && cith > 0.0) ? pow(cith, 1.25) : 0.0;
  if (c_dc.licerun && c_dc.lciwa3 && alpfac != c_dc.zalpfacx)
    for (int m = 0; m < F; ++m)
      d.tbg[((size_t)TQ_SBO * F + m) * n + p] += -ci * ((c_dc.fr45[m] * cith125) * (alpfac - c_dc.zalpfacx)) * d.f.cgroup[idx3(d, p, m)];
  // LWNEMOCOUWRS (wnfluxes.F90:178-196): the wave radiative stress on the sea ice, from SLICE = FL1 * FLDICE / max(1 - DELT5 FLDICE, 1) of
  // the LAST attenuation term that is on (SLICE is INTENT(OUT) of SDICE1, 2, 3 in turn: sdice.F90:99-112), summed over all frequencies.
  // FL1 is the spectrum IMPLSCH works on (depth-limited, floored: sinflx.F90:126-129), still in place before the sweep.
  if (nd.wrs_on) {
    const int A = c_dc.A;
    const long long c = p / d.P;
    const int i = (int)(p - c * d.P);
    const size_t kstr = (size_t)d.P;
    const double* hi = d.f.fl1 + (size_t)i + (size_t)d.P * A * F * (size_t)c;
    const double* lo = hi;
    int mlo = 0;
    if (d.lo_on) { const int il = (p < d.nloc) ? i : 0; lo = d.fl_lo + (size_t)il + (size_t)d.P * A * d.lo_F * (size_t)c; mlo = d.Fr; }
    const double fac = d.scr[S_FAC * n + p];
    double snw, csw;
    sincos(d.f.wdwave[p], &snw, &csw);
    const double flmc = (1. - 0.9 * dmin(ci, 0.99)) * c_dc.flmin;
    const double delt5 = c_dc.ximp * c_dc.delt, eps1000 = c_dc.EPSMIN * 1000.0;
    const bool any = c_dc.licerun && c_dc.lciwa_any;
    // SDICE1's ALP per frequency is needed again when it is the last term: recompute as above
    double xsi = 0.0, ysi = 0.0;
    for (int m = 0; m < F; ++m) {
      const size_t o3 = idx3(d, p, m);
      const double cg = d.f.cgroup[o3], wk = d.f.wavnum[o3];
      double fld_m = 0.0;        // FLDICE of a term that does not depend on the bin
      if (any && c_dc.lciwa3) fld_m = -((c_dc.fr45[m] * cith125) * alpfac) * cg;
      else if (any && !c_dc.lciwa2 && c_dc.lciwa1 && d.ice1 && cith > 0.0) {
        const int NT = d.ice_nt, NH = d.ice_nh;
        const double* tabw = d.ice1 + (size_t)NT * NH;
        int ih = (int)floor((cith - d.ice_hmin) / d.ice_dh + 1);
        ih = max(1, min(ih, NH));
        const int ih1 = max(1, min(ih + 1, NH));
        const double wh1 = dmax(dmin(1.0, (cith - (d.ice_hmin + (ih - 1) * d.ice_dh)) / d.ice_dh), 0.0), wh = 1.0 - wh1;
        const double wt1 = tabw[m], wt = 1.0 - wt1;
        const int it = (int)tabw[F + m], it1 = (int)tabw[2 * F + m];
        const double cc = wt * (wh * d.ice1[it + NT * (ih - 1)] + wh1 * d.ice1[it + NT * (ih1 - 1)]) +
                          wt1 * (wh * d.ice1[it1 + NT * (ih - 1)] + wh1 * d.ice1[it1 + NT * (ih1 - 1)]);
        // mean floe size as in the LCIWA1 block above
        const double CIFRGL = 0.955, CIDMIN = 20.0, CIFRGMT = 2.0, A0 = 200.0, C0 = 300.0;
        const int maxicm = (int)(log(A0 / CIDMIN) / log(CIFRGMT));
        const double cidmax = A0 + C0 * ci;
        const int icm = min((int)(log(cidmax / CIDMIN) / log(CIFRGMT)), maxicm);
        double sn = 0.0, sd = 0.0, x = 1.0, fi = 1.0;
        for (int j = 0; j <= icm; ++j) { sn = sn + x * cidmax / fi; sd = sd + x; x = x * (CIFRGMT * CIFRGMT * CIFRGL); fi = fi * CIFRGMT; }
        fld_m = -(exp(cc) * (1.0 / (sn / sd)) * c_dc.zalpfacb) * cg;
      }
      const bool perbin = any && !c_dc.lciwa3 && c_dc.lciwa2;
      const double c2 = (c_dc.cdicwa * (wk * wk) * 4.0 * c_dc.zalpfacb) * cg;
      double sx = 0.0, sy = 0.0;
      for (int k = 0; k < A; ++k) {
        double f = dmax(__ldg((m < mlo ? lo : hi) + ((size_t)m * A + k) * kstr) * fac, c_dc.EPSMIN);
        if (m == F - 1) f = dmax(f, flmc * sq(dmax(0.0, c_dc.COSTH[k] * csw + c_dc.SINTH[k] * snw)));
        const double fld = perbin ? -c2 * sqrt(dmax(c_dc.EPSMIN, f * c_dc.DFIM[m])) : fld_m;
        const double sl = dmin((f * fld) / dmax(1.0 - delt5 * fld, 1.0), -eps1000);
        if (k == 0) { sx = c_dc.SINTH[0] * sl; sy = c_dc.COSTH[0] * sl; }
        else { sx = sx + c_dc.SINTH[k] * sl; sy = sy + c_dc.COSTH[k] * sl; }
      }
      const double cinv = d.f.cinv[o3];
      xsi = xsi + c_dc.zalpwrs * sx * cinv * c_dc.RHOWG_DFIM[m];
      ysi = ysi + c_dc.zalpwrs * sy * cinv * c_dc.RHOWG_DFIM[m];
    }
    d.scr[S_XSTR * n + p] = xsi; d.scr[S_YSTR * n + p] = ysi;
  }
}

// =========================================================================================================
// k_nemo: what IMPLSCH hands to the ocean model, one thread per grid point after the frequency sweep.
//   LWNEMOCOUSTRN: CIMSSTRN (cimsstrn.F90:83-119), the mean square wave strain in the sea ice from the NEW spectrum, with AKI_ICE's
//     flexural-gravity wavenumber (aki_ice.F90:66-112) -> STRNMS.
//   LWNEMOCOU: the NEMO part of WNFLUXES (wnfluxes.F90:304-330, LNUPD = T): NPHIEPS, NTAUOC, NSWH = 4 sqrt(EM_OC), NMWP = 1 / F1_OC
//     overwritten; NEMOTAUX/Y, NEMOWSWAVE, NEMOPHIF, NEMOTAUICX/Y accumulated; and of STOKESTRN (stokestrn.F90:76-88).  EM_OC, F1_OC
//     (wnfluxes.F90:222-250) are rebuilt from FKMEAN's EMEAN / F1MEAN of the incoming spectrum (scratch of k_point) exactly as the
//     sweep's WNFLUXES closure forms OOVAL and USTAR.
// =========================================================================================================
__device__ double aki_ice(double xk, double depth, double cith) {
  const double YMICE = 5.5e9, RMUICE = 0.3, RHOI = 922.5, EBS = 0.000001, AKI_MAX = 20.0;
  if (cith <= 0.0) return xk;
  const double ficstf = (YMICE * (cith * cith * cith) / (12.0 * (1.0 - RMUICE * RMUICE))) / c_dc.ROWATER;
  const double rdh = (RHOI / c_dc.ROWATER) * cith;
  const double om2 = c_dc.G * xk * tanh(xk * depth);
  double akiold = 0.0;
  double aki = dmin(xk, pow(om2 / dmax(ficstf, 1.0), 0.2));
  while (fabs(aki - akiold) > EBS * akiold && aki < AKI_MAX) {
    akiold = aki;
    const double akid = dmin(depth * aki, 50.0);
    const double a2 = aki * aki, a4 = a2 * a2;
    const double f = ficstf * (a4 * aki) + c_dc.G * aki - om2 * (rdh * aki + 1.0 / tanh(akid));
    const double sh = sinh(akid);
    const double fprime = 5.0 * ficstf * a4 + c_dc.G - om2 * (rdh - depth / (sh * sh));
    aki = aki - f / fprime;
    if (aki <= 0.0) aki = AKI_MAX;
  }
  return aki;
}

__global__ void __launch_bounds__(128) k_nemo(ImplDev d, long long p0, long long np) {
  const long long p = p0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= p0 + np) return;
  const int A = c_dc.A, F = c_dc.F;
  const size_t n = (size_t)d.npts;
  const NemoDev nd = *d.nemo;
  if (nd.strn_on) {
    const long long c = p / d.P;
    const int i = (int)(p - c * d.P);
    const double* fl = d.f.fl1 + (size_t)i + (size_t)d.P * A * F * (size_t)c;
    const double dep = d.f.depth[p], cith = d.f.cithick[p];
    const double f1lim = c_dc.flmin / c_dc.DELTH;
    double strn = 0.0;
    for (int m = 0; m < F; ++m) {
      const double wk = d.f.wavnum[idx3(d, p, m)];
      const double xki = aki_ice(wk, dep, cith);
      const double e = 0.5 * cith * (xki * xki * xki) / wk;
      double sume = 0.0;
      for (int k = 0; k < A; ++k) sume = sume + fl[((size_t)m * A + k) * d.P];
      if (sume > f1lim) strn = strn + e * e * sume * c_dc.DFIM[m];
    }
    d.f.strnms[p] = strn;
  }
  if (nd.wrs_on) { d.f.tauicx[p] = -d.scr[S_XSTR * n + p]; d.f.tauicy[p] = -d.scr[S_YSTR * n + p]; }   // flipped: positive stress on the ice (wnfluxes.F90:266-271)
  if (!nd.nemo_on) return;
  const ecwam_b200_nemo_fields& o = nd.f;
  {   // WNFLUXES
    const double C1 = 1.03e-3, C2 = 0.04e-3, P1 = 1.48, P2 = -0.21, CDMAX_LOC = 0.003, EFD_MIN = 0.0625, EFD_MAX = 6.25;
    const double cicover = d.f.cicover[p], wsw = d.f.wswave[p], ufric = d.f.ufric[p];
    const double em = d.scr[S_EMEAN * n + p], f1 = d.scr[S_F1MEAN * n + p];
    const double cithrsh_inv = c_dc.lciwa_any ? 50.0 : 1.0 / dmax(c_dc.cithrsh, 0.01);
    const double zcithrs = c_dc.lciwa_any ? 0.0 : c_dc.ciblock, zmaxexp = c_dc.lciwa_any ? 20.0 : 10.0;
    double em_oc = em, f1_oc = f1;
    if (c_dc.licerun && c_dc.lwamrsetci && cicover > zcithrs) {
      const double ooval = exp(-dmin(p4(cicover * cithrsh_inv), zmaxexp));
      const double u10p = dmax(wsw, c_dc.EPSU10);
      const double cd_bulk = dmin((C1 + C2 * pow(u10p, P1)) * pow(u10p, P2), CDMAX_LOC);
      const double cd_wave = sq(ufric / u10p);
      const double ustar = dmax(sqrt(ooval * cd_wave + (1.0 - ooval) * cd_bulk) * u10p, c_dc.EPSUS);
      const double efd = dmin((4.0 * c_dc.EGRCRV / (c_dc.G * c_dc.G)) * p4(ustar), EFD_MAX);
      em_oc = dmax(ooval * em + (1.0 - ooval) * efd, EFD_MIN);
      const double ffd = (pow(c_dc.EGRCRV / c_dc.AFCRV, 1.0 / c_dc.BFCRV) * c_dc.G) / ustar;
      f1_oc = dmin(dmax(ooval * f1 + (1.0 - ooval) * ffd, c_dc.FR[1]), c_dc.FR[F - 1]);
    }
    o.nphieps[p] = d.f.phieps[p];
    o.ntauoc[p] = d.f.tauoc[p];
    o.nswh[p] = em_oc != 0.0 ? 4.0 * sqrt(em_oc) : 0.0;
    o.nmwp[p] = f1_oc != 0.0 ? 1.0 / f1_oc : 0.0;
    if (c_dc.lwnemotauoc) { o.nemotaux[p] = o.nemotaux[p] + d.f.tauocxd[p]; o.nemotauy[p] = o.nemotauy[p] + d.f.tauocyd[p]; }
    else { o.nemotaux[p] = o.nemotaux[p] + d.f.tauxd[p]; o.nemotauy[p] = o.nemotauy[p] + d.f.tauyd[p]; }
    o.nemowswave[p] = o.nemowswave[p] + wsw;
    o.nemophif[p] = o.nemophif[p] + d.f.phiocd[p];
    o.nemotauicx[p] = o.nemotauicx[p] + d.f.tauicx[p];
    o.nemotauicy[p] = o.nemotauicy[p] + d.f.tauicy[p];
  }
  if (c_dc.nemo_send) {   // STOKESTRN
    o.nemoustokes[p] = c_dc.lwnemocoustk ? d.f.ustokes[p] : 0.0;
    o.nemovstokes[p] = c_dc.lwnemocoustk ? d.f.vstokes[p] : 0.0;
    if (nd.strn_on) o.nemostrn[p] = d.f.strnms[p];
  }
}

static int launch_sweep_stage(const ImplDev& d, long long p0, long long np, cudaStream_t st);

int launch_implsch_stage(const ImplDev& d, long long p0, long long np, int stage, cudaStream_t st) {
  if (np <= 0) return 0;
  if (stage == 1) {
    const int rc = launch_sweep_stage(d, p0, np, st);
    if (rc == 0 && d.nemo) k_nemo<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(d, p0, np);
    return rc;
  }
  const int A = d.A;
  if (A > 36 || 2 * d.halo_r > A || 2 * d.halo_c > A) { ew_set_error("k_stencil is built for NANG <= 36 and direction halos <= NANG/2"); return ECWAM_B200_EINVAL; }
  if (stage == 0) {
    const size_t smp = (size_t)2 * A * KP_NTH * sizeof(double);
    static bool attr_p = false;
    if (!attr_p) {
      const void* kp[8] = {(const void*)k_point<true, 1, false>, (const void*)k_point<true, 2, false>, (const void*)k_point<false, 1, false>,
                           (const void*)k_point<false, 2, false>, (const void*)k_point<true, 1, true>, (const void*)k_point<true, 2, true>,
                           (const void*)k_point<false, 1, true>, (const void*)k_point<false, 2, true>};
      for (int i = 0; i < 8; ++i) {
        EW_CUDA_CHECK(cudaFuncSetAttribute(kp[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        EW_CUDA_CHECK(cudaFuncSetAttribute(kp[i], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      }
      attr_p = true;
    }
    const unsigned nb = (unsigned)((np + KP_NTH - 1) / KP_NTH);
    const bool cy49 = (d.cy49 & 1) != 0;
    if (d.cy49 & 2) {   // ICODE_WND = 1, 2: the first SINFLX call's AIRSEA is Z0WAVE
      static bool attr_u = false;
      if (!attr_u) {
        const void* ku[4] = {(const void*)k_point<true, 1, false, true>, (const void*)k_point<false, 1, false, true>,
                             (const void*)k_point<true, 1, true, true>, (const void*)k_point<false, 1, true, true>};
        for (int i = 0; i < 4; ++i) {
          EW_CUDA_CHECK(cudaFuncSetAttribute(ku[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
          EW_CUDA_CHECK(cudaFuncSetAttribute(ku[i], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        }
        attr_u = true;
      }
      if (cy49) {
        if (d.iphys == 1) { k_point<true, 1, true, true><<<nb, KP_NTH, smp, st>>>(d, p0, np); k_point<true, 2, true><<<nb, KP_NTH, smp, st>>>(d, p0, np); }
        else { k_point<false, 1, true, true><<<nb, KP_NTH, smp, st>>>(d, p0, np); k_point<false, 2, true><<<nb, KP_NTH, smp, st>>>(d, p0, np); }
      } else if (d.iphys == 1) { k_point<true, 1, false, true><<<nb, KP_NTH, smp, st>>>(d, p0, np); k_point<true, 2, false><<<nb, KP_NTH, smp, st>>>(d, p0, np); }
      else { k_point<false, 1, false, true><<<nb, KP_NTH, smp, st>>>(d, p0, np); k_point<false, 2, false><<<nb, KP_NTH, smp, st>>>(d, p0, np); }
    } else if (cy49) {
      if (d.iphys == 1) { k_point<true, 1, true><<<nb, KP_NTH, smp, st>>>(d, p0, np); k_point<true, 2, true><<<nb, KP_NTH, smp, st>>>(d, p0, np); }
      else { k_point<false, 1, true><<<nb, KP_NTH, smp, st>>>(d, p0, np); k_point<false, 2, true><<<nb, KP_NTH, smp, st>>>(d, p0, np); }
    } else if (d.iphys == 1) { k_point<true, 1, false><<<nb, KP_NTH, smp, st>>>(d, p0, np); k_point<true, 2, false><<<nb, KP_NTH, smp, st>>>(d, p0, np); }
    else { k_point<false, 1, false><<<nb, KP_NTH, smp, st>>>(d, p0, np); k_point<false, 2, false><<<nb, KP_NTH, smp, st>>>(d, p0, np); }
    if (d.isnonlin != 0) k_enh<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(d, p0, np);
    if (d.ice1 || d.ice2 || d.nemo) k_ice<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(d, p0, np);
  } else return ECWAM_B200_EINVAL;
  return 0;
}

static int launch_sweep_stage(const ImplDev& d, long long p0, long long np, cudaStream_t st) {
  {
    // two grid points per thread (16-byte shared / global accesses) need an even NPROMA and an even first point
    const uintptr_t al = (uintptr_t)d.f.fl1 | (uintptr_t)d.f.xllws | (uintptr_t)d.fldin | (uintptr_t)d.fl_lo;
    bool pair = (d.P % 2 == 0) && (p0 % 2 == 0) && (np % 2 == 0) && (al & 15) == 0;
    // test hook: ECWAM_B200_STENCIL=generic (run-time geometry instance) | single (one point per thread, run-time geometry)
    const char* force = getenv("ECWAM_B200_STENCIL");
    const bool generic = force && (!strcmp(force, "generic") || !strcmp(force, "single"));
    if (force && !strcmp(force, "single")) pair = false;
    // default for the standard grids: thread = (two adjacent directions, one point); ECWAM_B200_STENCIL=pp keeps the
    // two-points-per-thread instance (A/B timing, tests)
    // LWVFLX_SNL = F (SSOURCE = SL before SNONLIN, implsch.F90:279-288) is built into k_stencil / k_stencil_dp only
    // LCIWA2 (SDICE2's per-bin attenuation) likewise
    const bool sweep_ok = d.sweep_ok && !d.ssource_pre && !d.ice2;
    // default for NANG = 36: k_sweep_ws (producer / consumer warp groups); ECWAM_B200_STENCIL=sweep: the one-role k_sweep
    if (!force && sweep_ok && geo_matches<36>(d, d.iphys, d.nsdsnth))
      return d.lwflux ? launch_sweep_ws<36, 7, true>(d, p0, np, st) : launch_sweep_ws<36, 7, false>(d, p0, np, st);
    // NANG = 24: k_sweep_ws with 10 points per CTA (120 of 128 lanes per role); ECWAM_B200_STENCIL=sweep keeps the one-role k_sweep
    if ((!force || !strcmp(force, "ws")) && sweep_ok && geo_matches<24>(d, d.iphys, d.nsdsnth))
      return d.lwflux ? launch_sweep_ws<24, 10, true>(d, p0, np, st) : launch_sweep_ws<24, 10, false>(d, p0, np, st);
    // default for the other standard grids: k_sweep (NPT points x NANG/2 direction pairs)
    if ((!force || !strcmp(force, "sweep")) && sweep_ok) {
      if (geo_matches<36>(d, d.iphys, d.nsdsnth)) return d.lwflux ? launch_sweep<36, 7, true>(d, p0, np, st) : launch_sweep<36, 7, false>(d, p0, np, st);
      if (geo_matches<24>(d, d.iphys, d.nsdsnth)) return d.lwflux ? launch_sweep<24, 8, true>(d, p0, np, st) : launch_sweep<24, 8, false>(d, p0, np, st);
      if (geo_matches<12>(d, d.iphys, d.nsdsnth)) return d.lwflux ? launch_sweep<12, 8, true>(d, p0, np, st) : launch_sweep<12, 8, false>(d, p0, np, st);
    }
    // ECWAM_B200_STENCIL=dp: thread = (two adjacent directions, one point), 8 points x 160 threads (the previous default; tests, A/B)
    if (!force || !strcmp(force, "dp")) {
      if (geo_matches<36>(d, d.iphys, d.nsdsnth)) return d.lwflux ? launch_stencil_dp<36, true>(d, p0, np, st) : launch_stencil_dp<36, false>(d, p0, np, st);
      if (geo_matches<24>(d, d.iphys, d.nsdsnth)) return d.lwflux ? launch_stencil_dp<24, true>(d, p0, np, st) : launch_stencil_dp<24, false>(d, p0, np, st);
      if (geo_matches<12>(d, d.iphys, d.nsdsnth)) return d.lwflux ? launch_stencil_dp<12, true>(d, p0, np, st) : launch_stencil_dp<12, false>(d, p0, np, st);
    }
    if (pair) {
      if (!generic) {
        if (geo_matches<36>(d, d.iphys, d.nsdsnth)) return d.lwflux ? launch_stencil<36, 2, true>(d, p0, np, st) : launch_stencil<36, 2, false>(d, p0, np, st);
        if (geo_matches<24>(d, d.iphys, d.nsdsnth)) return d.lwflux ? launch_stencil<24, 2, true>(d, p0, np, st) : launch_stencil<24, 2, false>(d, p0, np, st);
        if (geo_matches<12>(d, d.iphys, d.nsdsnth)) return d.lwflux ? launch_stencil<12, 2, true>(d, p0, np, st) : launch_stencil<12, 2, false>(d, p0, np, st);
      }
      return d.lwflux ? launch_stencil<0, 2, true>(d, p0, np, st) : launch_stencil<0, 2, false>(d, p0, np, st);
    }
    return d.lwflux ? launch_stencil<0, 1, true>(d, p0, np, st) : launch_stencil<0, 1, false>(d, p0, np, st);
  }
}

}  // namespace ew
