// IMPLSCH (src/ecwam/implsch.F90:10-468) for sm_100a: the implicit source-term step of WAMINTGR.
//
// The reference runs the whole call tree per NPROMA chunk with every (IJ,K,M) temporary in memory
// (implsch.F90:152-170).  Here the step is five kernels:
//   k_airsea1   lane = grid point   AIRSEA/TAUT_Z0, first call           (sinflx.F90 ICALL=1; taut_z0.F90:281-341)
//   k_spec<1>   warp = grid point   SDEPTHLIM, FKMEAN, SINPUT (NGST=1), FEMEANWS, FRCUTINDEX, STRESSO sums
//   k_scalar2   lane = grid point   TAU_PHI_HF, TAUW, TAUT_Z0 (2nd call), WSIGSTAR, swell-friction scalars, SDIWBK Q
//   k_spec<2>   warp = grid point   SINPUT (NGST=2, LLSNEG), FEMEANWS, FRCUTINDEX, STRESSO sums, SDISSIP, SNONLIN,
//                                   SDIWBK, SBOTTOM, implicit update, WNFLUXES sums, IMPHFTAIL, SETICE, STOKESDRIFT
//   k_scalar4   lane = grid point   TAU_PHI_HF (stress + PHI), TAUW/TAUWDIR/PHIWA, WNFLUXES closure
// In the warp-per-point kernels the NANG x NFRE spectrum of the point lives in shared memory, lane = direction,
// the per-frequency direction sums are warp-shuffle reductions, and nothing but FL1, XLLWS and the 1-D outputs
// goes back to HBM.  The serial per-point solvers (Newton loops, the 19-point HF integral) run one point per lane.
#include "internal.h"

namespace ew {

__constant__ DevConst c_dc;
int upload_dev_const(const DevConst& h, cudaStream_t st) {
  EW_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_dc, &h, sizeof(DevConst), 0, cudaMemcpyHostToDevice, st));
  return 0;
}

// scalar scratch slots [slot][npts]
enum {
  S_EMEAN = 0, S_FMEAN, S_F1MEAN, S_AKMEAN, S_XKMEAN,   // FKMEAN (first call)
  S_XSTR, S_YSTR, S_F1DCOS3, S_F1DCOS2, S_PHIWA,         // STRESSO partial sums / TAU_PHI_HF inputs
  S_UORBT, S_AORB, S_SIGN, S_TEMP2, S_PTURB, S_PVISC,    // SINPUT_ARD swell-dissipation scalars, WSIGSTAR
  S_SDS,                                                  // SDIWBK
  S_PHILF, S_XSTROC, S_YSTROC,                            // WNFLUXES sums
  S_MIJ, S_USTOLD,
  NSCR
};
size_t implsch_scratch_doubles(long long npts) { return (size_t)NSCR * (size_t)npts; }

#define FULLMASK 0xffffffffu
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(FULLMASK, v, o));
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
  const double tauwact = fmax(tauw * cosdiff, c_dc.EPSMIN);
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
  const double ustold = (1 - iusfg) * utop * sqrt(fmin(c_dc.ACD + c_dc.BCD * utop, c_dc.CDMAX)) + iusfg * ustar;
  double tauold = fmax(sq(ustold), tauweff);
  ustar = sqrt(tauold);
  double ustm1 = 1.0 / fmax(ustar, c_dc.EPSUS);
  double z0ch = 0.0;
  for (int iter = 1; iter <= NITER; ++iter) {
    const double x = fmax(tauwact / tauold, xmin);
    z0ch = alphaog * tauold / sqrt(1.0 - x);
    const double z0vis = c_dc.rnum * ustm1;
    const double z0tot = z0ch + z0vis;
    const double xologz0 = 1.0 / (xlogxl - log(z0tot));
    const double f = ustar - xkutop * xologz0;
    const double zz = ustm1 * (z0ch * (2.0 - TWOXMP1 * x) / (1.0 - x) - z0vis) / z0tot;
    const double delf = 1.0 - xkutop * sq(xologz0) * zz;
    if (delf != 0.0) ustar = ustar - f / delf;
    const double taunew = fmax(sq(ustar), tauweff);
    ustar = sqrt(taunew);
    if (taunew == tauold) break;
    ustm1 = 1.0 / fmax(ustar, c_dc.EPSUS);
    tauold = taunew;
  }
  z0 = z0ch;
  z0b = alphaog * tauold;
  chrnck = fmax(c_dc.G * z0 * sq(ustm1), c_dc.ALPHAMIN);
}

__global__ void __launch_bounds__(128) k_airsea1(ImplDev d, long long p0, long long np) {
  const long long p = p0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= p0 + np) return;
  double ustar = d.f.ufric[p], z0, z0b, ch;
  taut_z0(0, d.f.wswave[p], d.f.wdwave[p], d.f.tauw[p], d.f.tauwdir[p], ustar, z0, z0b, ch);
  d.f.ufric[p] = ustar; d.f.z0m[p] = z0; d.f.z0b[p] = z0b; d.f.chrnck[p] = ch;
}

// ---------------------------------------------------------------------------------------------------------
// tau_phi_hf.F90:111-305 (LLGCBZ0=F, LLNORMAGAM=F: CONST1=CONST2=0 so GAMNORMA=1)
__device__ void tau_phi_hf(int mij, bool shelter, double z0m, double aird, double f1dcos3, double f1dcos2, double& ust,
                           double& tauhf, double& phihf, bool llphihf) {
  const double ZSUPMAX = 0.0;
  const double x0g = c_dc.X0TAUHF * c_dc.G;
  double ustph = ust;
  const double xloggz0 = log(c_dc.G * z0m);
  const double omegacc = fmax(c_dc.ZPIFR[mij - 1], x0g / ust);
  const double sqrtz0og = sqrt(z0m * c_dc.GM1);
  const double sqrtgz0 = 1.0 / sqrtz0og;
  const double yc = omegacc * sqrtz0og;
  const double zinf = log(yc);
  const double consttau = c_dc.ZPI4GM2 * c_dc.FR5[mij - 1];
  double taul = sq(ust);
  double delz = fmax((ZSUPMAX - zinf) / (double)(c_dc.JTOT - 1), 0.0);
  tauhf = 0.0;
  if (shelter) {
    for (int j = 1; j <= c_dc.JTOT; ++j) {
      const double y = exp(zinf + (double)(j - 1) * delz);
      const double omega = y * sqrtgz0;
      const double cm1 = omega * c_dc.GM1;
      const double zx = ust * cm1 + c_dc.ZALP;
      const double zarg = c_dc.XKAPPA / zx;
      double zlog = xloggz0 + 2.0 * log(cm1) + zarg;
      zlog = fmin(zlog, 0.0);
      const double zbeta = p4(zlog) * exp(zlog);
      const double fnc2 = f1dcos3 * consttau * zbeta * taul * c_dc.WTAUHF[j - 1] * delz;
      taul = fmax(taul - c_dc.TAUWSHELTER * fnc2, 0.0);
      ust = sqrt(taul);
      tauhf = tauhf + fnc2;
    }
  } else {
    for (int j = 1; j <= c_dc.JTOT; ++j) {
      const double y = exp(zinf + (double)(j - 1) * delz);
      const double omega = y * sqrtgz0;
      const double cm1 = omega * c_dc.GM1;
      const double zx = ust * cm1 + c_dc.ZALP;
      const double zarg = c_dc.XKAPPA / zx;
      double zlog = xloggz0 + 2.0 * log(cm1) + zarg;
      zlog = fmin(zlog, 0.0);
      const double zbeta = p4(zlog) * exp(zlog);
      tauhf = tauhf + zbeta * c_dc.WTAUHF[j - 1];
    }
    tauhf = f1dcos3 * consttau * taul * tauhf * delz;
  }
  phihf = 0.0;
  if (llphihf) {
    taul = sq(ustph);
    delz = fmax((ZSUPMAX - zinf) / (double)(c_dc.JTOT - 1), 0.0);
    const double constphi = aird * c_dc.ZPI4GM1 * c_dc.FR5[mij - 1];
    if (shelter) {
      for (int j = 1; j <= c_dc.JTOT; ++j) {
        const double y = exp(zinf + (double)(j - 1) * delz);
        const double omega = y * sqrtgz0;
        const double cm1 = omega * c_dc.GM1;
        const double zx = ustph * cm1 + c_dc.ZALP;
        const double zarg = c_dc.XKAPPA / zx;
        double zlog = xloggz0 + 2.0 * log(cm1) + zarg;
        zlog = fmin(zlog, 0.0);
        const double zbeta = p4(zlog) * exp(zlog);
        const double fnc2 = zbeta * taul * c_dc.WTAUHF[j - 1] * delz;
        taul = fmax(taul - c_dc.TAUWSHELTER * f1dcos3 * consttau * fnc2, 0.0);
        ustph = sqrt(taul);
        phihf = phihf + fnc2 / y;
      }
      phihf = f1dcos2 * constphi * sqrtz0og * phihf;
    } else {
      for (int j = 1; j <= c_dc.JTOT; ++j) {
        const double y = exp(zinf + (double)(j - 1) * delz);
        const double omega = y * sqrtgz0;
        const double cm1 = omega * c_dc.GM1;
        const double zx = ustph * cm1 + c_dc.ZALP;
        const double zarg = c_dc.XKAPPA / zx;
        double zlog = xloggz0 + 2.0 * log(cm1) + zarg;
        zlog = fmin(zlog, 0.0);
        const double zbeta = p4(zlog) * exp(zlog);
        phihf = phihf + zbeta * c_dc.WTAUHF[j - 1] / y;
      }
      phihf = f1dcos2 * constphi * sqrtz0og * taul * phihf * delz;
    }
  }
}

// stresso.F90:187-233: closure of the wave stress from the low-frequency sums + HF tail
__device__ void stresso_close(const ImplDev& d, long long p, bool llphiwa, double& tauw, double& tauwdir, double& phiwa) {
  const double* s = d.scr;
  const long long n = d.npts;
  const double aird = d.f.aird[p], ufric = d.f.ufric[p], wdwave = d.f.wdwave[p], z0m = d.f.z0m[p];
  const double am = fmax(aird, 1.0);
  double xstress = s[S_XSTR * n + p] / am, ystress = s[S_YSTR * n + p] / am;
  const int mij = (int)s[S_MIJ * n + p];
  bool shelter;
  double usdirp, ust;
  if (c_dc.iphys == 0 || c_dc.TAUWSHELTER == 0.0) {
    shelter = false; usdirp = wdwave; ust = ufric;
  } else {
    shelter = true;
    const double taux = sq(ufric) * sin(wdwave), tauy = sq(ufric) * cos(wdwave);
    const double taupx = taux - c_dc.TAUWSHELTER * xstress, taupy = tauy - c_dc.TAUWSHELTER * ystress;
    usdirp = atan2(taupx, taupy);
    ust = sqrt(sqrt(taupx * taupx + taupy * taupy));
  }
  double tauhf, phihf;
  tau_phi_hf(mij, shelter, z0m, aird, s[S_F1DCOS3 * n + p], s[S_F1DCOS2 * n + p], ust, tauhf, phihf, llphiwa);
  xstress = xstress + tauhf * sin(usdirp);
  ystress = ystress + tauhf * cos(usdirp);
  tauw = fmax(sqrt(sq(xstress) + sq(ystress)), 0.0);
  tauwdir = atan2(xstress, ystress);
  tauw = fmin(tauw, sq(ufric) * (1.0 / (1.0 + c_dc.EPS1)));   // .NOT. LLGCBZ0 (stresso.F90:218-223)
  phiwa = llphiwa ? s[S_PHIWA * n + p] + phihf : 0.0;
}

// wsigstar.F90:105-129
__device__ double wsigstar(double ufric, double z0m, double wstar) {
  const double ONETHIRD = 1.0 / 3.0, SIG_NMAX = 0.9, C1 = 1.03e-3, C2 = 0.04e-3, P1 = 1.48, P2 = -0.21;
  const double xkappad = 1.0 / c_dc.XKAPPA;
  double u10 = ufric * xkappad * (log(10.0) - log(z0m));
  u10 = fmax(u10, c_dc.wspmin);
  const double u10m1 = 1.0 / u10;
  const double c2u10p1 = C2 * pow(u10, P1);
  const double u10p2 = pow(u10, P2);
  const double c_d = (C1 + c2u10p1) * u10p2;
  const double dc_ddu = (P2 * C1 + (P1 + P2) * c2u10p1) * u10p2 * u10m1;
  const double sig_conv = 1.0 + 0.5 * u10 / c_d * dc_ddu;
  return fmin(SIG_NMAX, sig_conv * u10m1 * pow(0.0 * ufric * ufric * ufric + 0.5 * c_dc.XKAPPA * wstar * wstar * wstar, ONETHIRD));
}

__global__ void __launch_bounds__(128) k_scalar2(ImplDev d, long long p0, long long np) {
  const long long p = p0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= p0 + np) return;
  double* s = d.scr;
  const long long n = d.npts;
  // STRESSO closure of SINFLX call 1 (LLPHIWA = F)
  double tauw, tauwdir, phiwa;
  stresso_close(d, p, false, tauw, tauwdir, phiwa);
  d.f.tauw[p] = tauw; d.f.tauwdir[p] = tauwdir;
  // SINFLX call 2: AIRSEA with IUSFG=1
  double ustar = d.f.ufric[p], z0, z0b, ch;
  taut_z0(1, d.f.wswave[p], d.f.wdwave[p], tauw, tauwdir, ustar, z0, z0b, ch);
  d.f.ufric[p] = ustar; d.f.z0m[p] = z0; d.f.z0b[p] = z0b; d.f.chrnck[p] = ch;
  // WSIGSTAR for NGST=2
  s[S_SIGN * n + p] = wsigstar(ustar, z0, d.f.wstar[p]);
  // swell-dissipation scalars of SINPUT_ARD (sinput_ard.F90:179-271), LLSNEG
  if (c_dc.iphys == 1) {
    const double raorw = fmax(d.f.aird[p], 1.0) * c_dc.ROWATERM1;
    const double uorbt = 2.0 * sqrt(s[S_UORBT * n + p]);
    const double aorb = 2.0 * sqrt(s[S_AORB * n + p]);
    const double re = (4.0 / c_dc.rnu) * uorbt * aorb;
    const double z0vis = c_dc.rnum / fmax(ustar, 0.0001);
    const double z0tub = c_dc.Z0RAT * fmin(c_dc.Z0TUBMAX, z0);
    const double z0noz = fmax(z0vis, z0tub);
    const double zorb = aorb / z0noz;
    const double delabm1 = (double)c_dc.IAB / (c_dc.ABMAX - c_dc.ABMIN);
    const double xi = (log10(fmax(zorb, 3.0)) - c_dc.ABMIN) * delabm1;
    const int ind = min(c_dc.IAB - 1, (int)xi);
    const double deli1 = fmin(1.0, xi - (double)ind);
    const double deli2 = 1.0 - deli1;
    const double fww = d.tab.swellft[ind - 1] * deli2 + d.tab.swellft[ind] * deli1;
    s[S_TEMP2 * n + p] = fww * uorbt;
    double re_c;
    if (c_dc.SWELLF6 == 1.0) re_c = c_dc.SWELLF4;
    else re_c = c_dc.SWELLF4 * pow(2.0 / aorb, 1.0 - c_dc.SWELLF6);
    double pturb, pvisc;
    if (c_dc.SWELLF7 > 0.0) {
      const double smooth = 0.5 * tanh((re - re_c) * c_dc.SWELLF7M1);
      pturb = 0.5 + smooth; pvisc = 0.5 - smooth;
    } else if (re <= re_c) { pturb = 0.0; pvisc = 0.5; }
    else { pturb = 0.5; pvisc = 0.0; }
    s[S_PTURB * n + p] = pturb;
    s[S_PVISC * n + p] = pvisc * raorw;   // AIRD_PVISC
  }
  // SDIWBK (sdiwbk.F90:69-104): Battjes-Janssen fraction of breaking waves
  double sds = 0.0;
  if (c_dc.lbiwbk && d.f.depth[p] < 50.0) {
    const double alph = 2.0 * d.f.emaxdpt[p] / s[S_EMEAN * n + p];
    const double arg = fmin(alph, 50.0);
    double q_old = exp(-arg), q = q_old;
    for (int ic = 1; ic <= 15; ++ic) {
      const double expq = exp(-arg * (1.0 - q_old));
      q = q_old - (expq - q_old) / (arg * expq - 1.0);
      const double rel_err = fabs(q - q_old) / q_old;
      if (rel_err < 0.00001) break;
      q_old = q;
    }
    q = fmin(q, 1.0);
    sds = 2.0 * alph * q * s[S_F1MEAN * n + p];
  }
  s[S_SDS * n + p] = sds;
}

// wnfluxes.F90:222-331 (LWNEMOCOU=F) after the spectral sums
__global__ void __launch_bounds__(128) k_scalar4(ImplDev d, long long p0, long long np) {
  const long long p = p0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= p0 + np) return;
  const double* s = d.scr;
  const long long n = d.npts;
  double tauw, tauwdir, phiwa;
  stresso_close(d, p, true, tauw, tauwdir, phiwa);
  d.f.tauw[p] = tauw; d.f.tauwdir[p] = tauwdir;
  if (!c_dc.lcflx) return;
  const double PHIOC_ICE = -3.75, PHIAW_ICE = 3.75, C1 = 1.03e-3, C2 = 0.04e-3, P1 = 1.48, P2 = -0.21, CDMAX_LOC = 0.003;
  const double epsus3 = c_dc.EPSUS * sqrt(c_dc.EPSUS);
  const double cithrsh_inv = 1.0 / fmax(c_dc.cithrsh, 0.01);
  const double cicover = d.f.cicover[p], ufric = d.f.ufric[p], aird = d.f.aird[p], wdwave = d.f.wdwave[p];
  double ooval = 1.0, ustar = ufric;
  if (c_dc.licerun && c_dc.lwamrsetci && cicover > c_dc.ciblock) {
    ooval = exp(-fmin(p4(cicover * cithrsh_inv), 10.0));
    const double u10p = fmax(d.f.wswave[p], c_dc.EPSU10);
    const double cd_bulk = fmin((C1 + C2 * pow(u10p, P1)) * pow(u10p, P2), CDMAX_LOC);
    const double cd_wave = sq(ufric / u10p);
    const double cd_ice = ooval * cd_wave + (1.0 - ooval) * cd_bulk;
    ustar = fmax(sqrt(cd_ice) * u10p, c_dc.EPSUS);
  }
  const double xstress = s[S_XSTROC * n + p], ystress = s[S_YSTROC * n + p], philf = s[S_PHILF * n + p];
  const double tau = aird * fmax(sq(ustar), c_dc.EPSUS);
  double tauxd = tau * sin(wdwave), tauyd = tau * cos(wdwave);
  double tauocxd = tauxd - ooval * xstress, tauocyd = tauyd - ooval * ystress;
  const double tauo = sqrt(sq(tauocxd) + sq(tauocyd));
  const double tauoc = fmin(fmax(tauo / tau, c_dc.TAUOCMIN), c_dc.TAUOCMAX);
  if (c_dc.lwcouast) {
    const double us = d.f.ustra[p], vs = d.f.vstra[p];
    if (us != 0.0 || vs != 0.0) { tauxd = us; tauocxd = us * tauoc; tauyd = vs; tauocyd = vs * tauoc; }
  }
  d.f.tauxd[p] = tauxd; d.f.tauyd[p] = tauyd; d.f.tauocxd[p] = tauocxd; d.f.tauocyd[p] = tauocyd; d.f.tauoc[p] = tauoc;
  d.f.tauicx[p] = 0.0; d.f.tauicy[p] = 0.0;
  const double xn = aird * fmax(ustar * ustar * ustar, epsus3);
  double phiocd = ooval * (philf - phiwa) + (1.0 - ooval) * PHIOC_ICE * xn;
  double phieps = phiocd / xn;
  phieps = fmin(fmax(phieps, c_dc.PHIEPSMIN), c_dc.PHIEPSMAX);
  phiocd = phieps * xn;
  const double phiaw = ooval * phiwa / xn + (1.0 - ooval) * PHIAW_ICE;
  d.f.phiocd[p] = phiocd; d.f.phieps[p] = phieps; d.f.phiaw[p] = phiaw;
}

// =========================================================================================================
// Warp-per-point spectral kernels
// =========================================================================================================
#define KPL 2   // directions per lane: k = lane, lane + 32  (NANG <= 64)

struct WarpPt {
  double* fl;    // [F][A] spectrum of this point (shared memory)
  double* fld;   // [F][A] (pass 2)
  double* sl;    // [F][A] (pass 2)
  double* tb;    // per-frequency tables [NTB][EW_MAXF]
  unsigned char* xl;   // [F][A] XLLWS flags (pass 2)
};
enum { TB_WAVNUM = 0, TB_CINV, TB_XK2CG, TB_ZCN, TB_A, TB_B, NTB };

// RHOWGDFTH(IJ,M) of frcutindex.F90:99-108 (m 0-based)
__device__ __forceinline__ double rhowgdfth(int m, int mij) {
  if (m + 1 > mij) return 0.0;
  double r = c_dc.RHOWG_DFIM[m];
  if (m + 1 == mij && mij != c_dc.F) r = 0.5 * r;
  return r;
}

template <int PASS>
__global__ void __launch_bounds__(PASS == 1 ? 256 : 192) k_spec(ImplDev d, long long p0, long long np) {
  constexpr int WPB = (PASS == 1) ? 8 : 6;
  extern __shared__ double smem[];
  const int A = c_dc.A, F = c_dc.F, AF = A * F;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long pb = p0 + (long long)blockIdx.x * WPB;
  const long long pend = p0 + np;
  const int nptb = (int)min((long long)WPB, pend - pb);
  // ---- shared memory carve-up
  const int per_pt = (PASS == 1 ? 1 : 3) * AF + NTB * EW_MAXF + (PASS == 2 ? (AF + 7) / 8 : 0);
  WarpPt W;
  {
    double* base = smem + (size_t)w * per_pt;
    W.fl = base;
    W.fld = (PASS == 2) ? base + AF : nullptr;
    W.sl = (PASS == 2) ? base + 2 * AF : nullptr;
    W.tb = base + (PASS == 1 ? 1 : 3) * AF;
    W.xl = (PASS == 2) ? (unsigned char*)(W.tb + NTB * EW_MAXF) : nullptr;
  }
  // ---- cooperative load of the FL1 tile: consecutive threads read consecutive grid points (lane dimension of
  //      the NPROMA-chunked layout) so that every 32-byte sector fetched is fully used
  {
    const int tot = nptb * AF;
    for (int idx = threadIdx.x; idx < tot; idx += blockDim.x) {
      const int i = idx % nptb, bin = idx / nptb;
      const long long p = pb + i;
      const long long c = p / d.P;
      const int ln = (int)(p - c * d.P);
      const int m = bin / A;
      double v;
      if (m < d.Fr && d.lo_F != d.F) {
        // reading the propagation scratch (P,A,Fr,C): padded lanes of the last chunk take lane 0 (propag_wam.F90:388-398)
        v = d.fl_lo[(size_t)ln + (size_t)d.P * ((size_t)bin + (size_t)A * d.lo_F * (size_t)c)];
      } else {
        v = d.f.fl1[(size_t)ln + (size_t)d.P * ((size_t)bin + (size_t)AF * (size_t)c)];
      }
      smem[(size_t)i * per_pt + bin] = v;
    }
  }
  __syncthreads();
  const bool active = w < nptb;
  const long long p = active ? pb + w : pb;   // inactive warps shadow point pb but never store
  const long long n = d.npts;
  double* s = d.scr;

  // ---- per-point scalars
  const double aird = d.f.aird[p], wdwave = d.f.wdwave[p], cicover = d.f.cicover[p];
  const double ufric = d.f.ufric[p], z0m = d.f.z0m[p], depth = d.f.depth[p];
  const double raorw = fmax(aird, 1.0) * c_dc.ROWATERM1;
  double* fl = W.fl;
  // per-frequency tables: lane m loads frequency m
  for (int m = lane; m < F; m += 32) {
    const size_t o = idx3(d, p, m);
    const double wn = d.f.wavnum[o];
    W.tb[TB_WAVNUM * EW_MAXF + m] = wn;
    W.tb[TB_CINV * EW_MAXF + m] = d.f.cinv[o];
    W.tb[TB_XK2CG * EW_MAXF + m] = d.f.xk2cg[o];
    W.tb[TB_ZCN * EW_MAXF + m] = log(wn * z0m);
    const double sqk = sqrt(wn);
    W.tb[TB_A * EW_MAXF + m] = c_dc.DFIM[m] / sqk;   // FKMEAN TEMPA
    W.tb[TB_B * EW_MAXF + m] = sqk * c_dc.DFIM[m];   // FKMEAN TEMPX
  }
  // per-lane direction data
  double coswdif[KPL], sinwdif2[KPL], sinth[KPL], costh[KPL], flm[KPL];
  bool kv[KPL];
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const int k = lane + 32 * j;
    kv[j] = k < A;
    const int kk = kv[j] ? k : 0;
    coswdif[j] = cos(c_dc.TH[kk] - wdwave);
    sinwdif2[j] = sq(sin(c_dc.TH[kk] - wdwave));
    sinth[j] = c_dc.SINTH[kk];
    costh[j] = c_dc.COSTH[kk];
    flm[j] = (1. - 0.9 * fmin(cicover, 0.99)) * c_dc.flmin * sq(fmax(0.0, coswdif[j]));   // implsch.F90:237-242
  }
  __syncwarp();
  const double DELT25 = c_dc.WETAIL * c_dc.FR[F - 1] * c_dc.DELTH;

  // ---- SDEPTHLIM (sdepthlim.F90:50-82 with SEMEAN)
  if (c_dc.lbiwbk) {
    double acc = 0.0, last = 0.0;
    for (int m = 0; m < F; ++m) {
      double t = 0.0;
#pragma unroll
      for (int j = 0; j < KPL; ++j) if (kv[j]) t += fl[m * A + lane + 32 * j];
      acc += c_dc.DFIM[m] * t;
      last = t;
    }
    const double em = c_dc.EPSMIN + wsum(acc) + DELT25 * wsum(last);
    const double fac = fmin(d.f.emaxdpt[p] / em, 1.0);
    for (int m = 0; m < F; ++m)
#pragma unroll
      for (int j = 0; j < KPL; ++j) if (kv[j]) { const int o = m * A + lane + 32 * j; fl[o] = fmax(fl[o] * fac, c_dc.EPSMIN); }
  }
  // ---- FKMEAN (fkmean.F90:60-154), first call only; pass 2 re-reads the scalars
  double emean, fmean, f1mean, akmean, xkmean;
  if (PASS == 1) {
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, last = 0;
    for (int m = 0; m < F; ++m) {
      double t = 0.0;
#pragma unroll
      for (int j = 0; j < KPL; ++j) if (kv[j]) t += fl[m * A + lane + 32 * j];
      a0 += c_dc.DFIM[m] * t; a1 += c_dc.DFIMOFR[m] * t; a2 += c_dc.DFIMFR[m] * t;
      a3 += W.tb[TB_A * EW_MAXF + m] * t; a4 += W.tb[TB_B * EW_MAXF + m] * t;
      last = t;
    }
    a0 = wsum(a0); a1 = wsum(a1); a2 = wsum(a2); a3 = wsum(a3); a4 = wsum(a4); last = wsum(last);
    const double COEFM1 = c_dc.FRTAIL * c_dc.DELTH;
    const double COEF1 = c_dc.WP1TAIL * c_dc.DELTH * sq(c_dc.FR[F - 1]);
    const double COEFA = COEFM1 * sqrt(c_dc.G) / c_dc.ZPI;
    const double COEFX = COEF1 * (c_dc.ZPI / sqrt(c_dc.G));
    emean = c_dc.EPSMIN + a0 + DELT25 * last;
    fmean = emean / (c_dc.EPSMIN + a1 + COEFM1 * last);
    f1mean = (c_dc.EPSMIN + a2 + COEF1 * last) / emean;
    akmean = sq(emean / (c_dc.EPSMIN + a3 + COEFA * last));
    xkmean = sq((c_dc.EPSMIN + a4 + COEFX * last) / emean);
    if (active && lane == 0) {
      s[S_EMEAN * n + p] = emean; s[S_FMEAN * n + p] = fmean; s[S_F1MEAN * n + p] = f1mean;
      s[S_AKMEAN * n + p] = akmean; s[S_XKMEAN * n + p] = xkmean;
    }
  } else {
    emean = s[S_EMEAN * n + p]; fmean = s[S_FMEAN * n + p]; f1mean = s[S_F1MEAN * n + p];
    akmean = s[S_AKMEAN * n + p]; xkmean = s[S_XKMEAN * n + p];
  }
  // ---- SINFLX, first call: FL1(:,:,NFRE) = MAX(FL1, FLM) (sinflx.F90:126-129)
#pragma unroll
  for (int j = 0; j < KPL; ++j) if (kv[j]) { const int o = (F - 1) * A + lane + 32 * j; fl[o] = fmax(fl[o], flm[j]); }
  __syncwarp();

  // ---- SINPUT
  constexpr int NGST = (PASS == 1) ? 1 : 2;
  constexpr bool LLSNEG = (PASS == 2);
  double sig_n = 0.0, temp2_sw = 0.0, pturb = 0.0, aird_pvisc = 0.0;
  if (PASS == 2) {
    sig_n = s[S_SIGN * n + p];
    if (c_dc.iphys == 1) { temp2_sw = s[S_TEMP2 * n + p]; pturb = s[S_PTURB * n + p]; aird_pvisc = s[S_PVISC * n + p]; }
  }
  // lane-distributed per-frequency sums (lane m%32 keeps frequency m): SPOS moments for STRESSO
  double dsumx[2] = {0, 0}, dsumy[2] = {0, 0}, dsumt[2] = {0, 0};
  // lane-accumulated sums
  double ws_em = 0.0, ws_fm = 0.0, ws_last = 0.0;   // FEMEANWS
  double phiwa_acc = 0.0;                            // sum (SL-SPOS)*RHOWG_DFIM
  double uorbt_acc = 0.0, aorb_acc = 0.0;            // pass 1: orbital velocity / amplitude sums for pass 2
  const double CONST1 = c_dc.BETAMAXOXKAPPA2;
  if (c_dc.iphys == 1) {
    // ================= SINPUT_ARD (sinput_ard.F90:149-524) =================
    const double abs_shelter = fabs(c_dc.TAUWSHELTER);
    const bool ltauwshelter = abs_shelter != 0.0;
    double ustp[NGST], xstress[NGST], ystress[NGST], taux[NGST], tauy[NGST];
    if (NGST == 1) ustp[0] = ufric;
    else { ustp[0] = ufric * (1.0 + sig_n); ustp[NGST - 1] = ufric * (1.0 - sig_n); }
    const double snw = sin(wdwave), csw = cos(wdwave);
#pragma unroll
    for (int g = 0; g < NGST; ++g) {
      xstress[g] = 0.0; ystress[g] = 0.0;
      const double usg2 = sq(ustp[g]);
      taux[g] = usg2 * snw; tauy[g] = usg2 * csw;
    }
    const double rogoroair = c_dc.G / raorw;
    const double FU = fabs(c_dc.SWELLF3), FUD = c_dc.SWELLF2;
    for (int m = 0; m < F; ++m) {
      const double sig = c_dc.ZPIFR[m], sig2 = sig * sig;
      const double cinv = W.tb[TB_CINV * EW_MAXF + m], wavnum = W.tb[TB_WAVNUM * EW_MAXF + m];
      const double zcn = W.tb[TB_ZCN * EW_MAXF + m];
      const double cnsn = sig * CONST1 * raorw;
      const double constf = rogoroair * cinv * c_dc.DFIM[m];
      double coef = 0.0, coef5 = 0.0, dstab1 = 0.0, temp1 = 0.0;
      if (LLSNEG) {
        coef = -c_dc.SWELLF * 16. * sig2 / c_dc.G;
        coef5 = -c_dc.SWELLF5 * 2. * sqrt(2. * c_dc.rnu * sig);
        dstab1 = coef5 * aird_pvisc * wavnum;
        temp1 = coef * raorw;
      }
      double cosu[NGST], sinu[NGST], ucn[NGST], ucnzalpd[NGST];
#pragma unroll
      for (int g = 0; g < NGST; ++g) {
        if (ltauwshelter) {
          const double taupx = taux[g] - abs_shelter * xstress[g];
          const double taupy = tauy[g] - abs_shelter * ystress[g];
          // USDIRP = ATAN2(TAUPX,TAUPY); only cos(TH-USDIRP) is needed: use the unit vector instead of the angle
          const double t2 = taupx * taupx + taupy * taupy;
          const double rt = sqrt(t2);
          ustp[g] = sqrt(rt);
          if (rt > 0.0) { const double ri = 1.0 / rt; cosu[g] = taupy * ri; sinu[g] = taupx * ri; }
          else { cosu[g] = 1.0; sinu[g] = 0.0; }
        } else { cosu[g] = csw; sinu[g] = snw; }
        ucn[g] = ustp[g] * cinv;
        ucnzalpd[g] = c_dc.XKAPPA / (ucn[g] + c_dc.ZALP);
      }
      double sx[NGST], sy[NGST], st = 0.0;
#pragma unroll
      for (int g = 0; g < NGST; ++g) { sx[g] = 0.0; sy[g] = 0.0; }
#pragma unroll
      for (int j = 0; j < KPL; ++j) {
        if (!kv[j]) continue;
        const int o = m * A + lane + 32 * j;
        const double f = fl[o];
        double slp_avg = 0.0, flp_avg = 0.0;
        bool xll = false;
#pragma unroll
        for (int g = 0; g < NGST; ++g) {
          const double coslp = ltauwshelter ? (costh[j] * cosu[g] + sinth[j] * sinu[g]) : coswdif[j];
          double gam0 = 0.0;
          if (coslp > 0.01) {
            const double x = coslp * ucn[g];
            const double zlog = zcn + ucnzalpd[g] / coslp;
            if (zlog < 0.0) {
              const double zlog2x = zlog * zlog * x;
              gam0 = exp(zlog) * zlog2x * zlog2x * cnsn;
              xll = true;
            }
          }
          double dstab = 0.0;
          if (LLSNEG) {
            const double dstab2 = temp1 * (temp2_sw + (FU + FUD * coslp) * ustp[g]);
            dstab = dstab1 + pturb * dstab2;
          }
          double slp = gam0;              // GAMNORMA = 1
          const double flp = slp + dstab;
          slp = slp * f;
          sx[g] += slp * sinth[j];
          sy[g] += slp * costh[j];
          slp_avg += slp; flp_avg += flp;
        }
        const double avg = 1.0 / NGST;
        const double spos = avg * slp_avg;
        const double fldv = avg * flp_avg;
        const double slv = fldv * f;
        st += spos;
        if (PASS == 2) { W.fld[o] = fldv; W.sl[o] = slv; W.xl[o] = xll ? 1 : 0; phiwa_acc += (slv - spos) * c_dc.RHOWG_DFIM[m]; }
        const double xf = xll ? f : 0.0;
        ws_em += c_dc.DFIM[m] * xf; ws_fm += c_dc.DFIMOFR[m] * xf;
        if (m == F - 1) ws_last += xf;
        if (PASS == 1) { uorbt_acc += c_dc.DFIM[m] * sig2 * f; aorb_acc += c_dc.DFIM[m] * f; }
      }
      double sxa = 0.0, sya = 0.0;
#pragma unroll
      for (int g = 0; g < NGST; ++g) {
        sx[g] = wsum(sx[g]); sy[g] = wsum(sy[g]);
        xstress[g] += constf * sx[g];      // XSTRESS = XSTRESS + SLP*CONSTF*SINTH(K), summed over K
        ystress[g] += constf * sy[g];
        sxa += sx[g]; sya += sy[g];
      }
      if (PASS == 2) st = wsum(st);
      if (lane == (m & 31)) { dsumx[m >> 5] = sxa * (1.0 / NGST); dsumy[m >> 5] = sya * (1.0 / NGST); dsumt[m >> 5] = st; }
    }
  } else {
    // ================= SINPUT_JAN (sinput_jan.F90:150-400) =================
    const double CONST3 = c_dc.idamping * (2.0 * c_dc.XKAPPA / CONST1);
    const double xkappad = 1.0 / c_dc.XKAPPA;
    double us[NGST], wsin[NGST];
    if (NGST == 1) { us[0] = ufric; wsin[0] = 1.0; }
    else { us[0] = ufric * (1.0 - sig_n); us[NGST - 1] = ufric * (1.0 + sig_n); wsin[0] = 0.5; wsin[NGST - 1] = 0.5; }
    for (int m = 0; m < F; ++m) {
      const double sig = c_dc.ZPIFR[m], sig2 = sig * sig;
      const double cinv = W.tb[TB_CINV * EW_MAXF + m], wavnum = W.tb[TB_WAVNUM * EW_MAXF + m];
      const double ztanhkd = sig2 / (c_dc.G * wavnum);
      const double cnsn = sig * CONST1 * ztanhkd * raorw;
      const double zcn = W.tb[TB_ZCN * EW_MAXF + m];
      double ucn[NGST], const3_ucn2[NGST], ucnd[NGST], xvd[NGST];
#pragma unroll
      for (int g = 0; g < NGST; ++g) {
        ucn[g] = us[g] * cinv + c_dc.ZALP;
        const3_ucn2[g] = CONST3 * sq(ucn[g]);
        ucnd[g] = 1.0 / ucn[g];
        xvd[g] = 1.0 / (-us[g] * xkappad * zcn * cinv);
      }
      double sx = 0.0, sy = 0.0, st = 0.0;
#pragma unroll
      for (int j = 0; j < KPL; ++j) {
        if (!kv[j]) continue;
        const int o = m * A + lane + 32 * j;
        const double f = fl[o];
        double ufac1 = 0.0, ufac2 = 0.0;
        bool xll = false;
#pragma unroll
        for (int g = 0; g < NGST; ++g) {
          double gam0 = 0.0;
          if (coswdif[j] > 0.01) {
            const double zlog = zcn + c_dc.XKAPPA / coswdif[j] * ucnd[g];
            if (zlog < 0.0) {
              const double x = coswdif[j] * ucn[g];
              const double zlog2x = zlog * zlog * x;
              gam0 = zlog2x * zlog2x * exp(zlog) * cnsn;
              xll = true;
            }
          }
          ufac1 += wsin[g] * gam0;
          if (LLSNEG) ufac2 += wsin[g] * (const3_ucn2[g] * (coswdif[j] - xvd[g]));
        }
        const double fldv = ufac1 + ufac2 * cnsn;
        const double spos = ufac1 * f;
        const double slv = fldv * f;
        sx += spos * sinth[j]; sy += spos * costh[j]; st += spos;
        if (PASS == 2) { W.fld[o] = fldv; W.sl[o] = slv; W.xl[o] = xll ? 1 : 0; phiwa_acc += (slv - spos) * c_dc.RHOWG_DFIM[m]; }
        const double xf = xll ? f : 0.0;
        ws_em += c_dc.DFIM[m] * xf; ws_fm += c_dc.DFIMOFR[m] * xf;
        if (m == F - 1) ws_last += xf;
      }
      sx = wsum(sx); sy = wsum(sy);
      if (PASS == 2) st = wsum(st);
      if (lane == (m & 31)) { dsumx[m >> 5] = sx; dsumy[m >> 5] = sy; dsumt[m >> 5] = st; }
    }
  }
  // ---- FEMEANWS (femeanws.F90:50-127)
  ws_em = wsum(ws_em); ws_fm = wsum(ws_fm); ws_last = wsum(ws_last);
  const double emeanws = c_dc.EPSMIN + ws_em + DELT25 * ws_last;
  const double fmeanws = emeanws / (c_dc.EPSMIN + ws_fm + c_dc.FRTAIL * c_dc.DELTH * ws_last);
  // ---- FRCUTINDEX (frcutindex.F90:84-97)
  int mij;
  if (cicover <= c_dc.cithrsh_tail) {
    const double fpmh = c_dc.TAILFACTOR / c_dc.FR[0];
    const double fppm = c_dc.TAILFACTOR_PM * c_dc.G / (28.0 * c_dc.ZPIFR[0]);
    const double fm2 = fmax(fmeanws, fmean) * fpmh;
    const double fpm = fppm / fmax(ufric, c_dc.EPSMIN);
    const double fpm4 = fmax(fm2, fpm);
    const double xr = log10(fpm4) * c_dc.FLOGSPRDM1;
    mij = (int)lround(xr) + 1;     // NINT
    mij = min(max(1, mij), F);
  } else mij = F;
  // ---- STRESSO low-frequency sums (stresso.F90:120-186)
  {
    double xs = 0.0, ys = 0.0, pw = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = lane + 32 * h;
      if (m < F) {
        const double r = rhowgdfth(m, mij);
        const double cm = r * W.tb[TB_CINV * EW_MAXF + m];
        xs += cm * dsumx[h]; ys += cm * dsumy[h]; pw += r * dsumt[h];
      }
    }
    xs = wsum(xs); ys = wsum(ys);
    double f3 = 0.0, f2 = 0.0;
#pragma unroll
    for (int j = 0; j < KPL; ++j) if (kv[j]) {
      const double cw = fmax(coswdif[j], 0.0);
      const double fc2 = fl[(mij - 1) * A + lane + 32 * j] * cw * cw;
      f3 += fc2 * cw; f2 += fc2;
    }
    f3 = c_dc.DELTH * wsum(f3);
    f2 = c_dc.DELTH * wsum(f2);
    if (PASS == 2) { pw = wsum(pw) + wsum(phiwa_acc); }
    if (active && lane == 0) {
      s[S_XSTR * n + p] = xs; s[S_YSTR * n + p] = ys; s[S_F1DCOS3 * n + p] = f3; s[S_F1DCOS2 * n + p] = f2;
      s[S_MIJ * n + p] = (double)mij;
      if (PASS == 2) s[S_PHIWA * n + p] = pw;
    }
  }
  if (PASS == 1) {
    uorbt_acc = wsum(uorbt_acc); aorb_acc = wsum(aorb_acc);
    if (active && lane == 0) { s[S_UORBT * n + p] = c_dc.EPSMIN + uorbt_acc; s[S_AORB * n + p] = c_dc.EPSMIN + aorb_acc; }
    return;
  }

  if (PASS == 2) {
    double* fld = W.fld;
    double* sl = W.sl;
    __syncwarp();
    // ---- SDISSIP
    if (c_dc.iphys == 1) {
      // SDISSIP_ARD (sdissip_ard.F90:131-318), saturation-based part only (SSDSC3 = SSDSC5 = 0)
      const double tpiinv = 1.0 / c_dc.ZPI;
      const double tmp03 = 1.0 / (c_dc.SDSBR * c_dc.MICHE);
      const double ssdsc6m1 = 1. - c_dc.SSDSC6;
      const int ns = 2 * c_dc.NSDSNTH + 1;
      for (int m = 0; m < F; ++m) {
        const double facsat = W.tb[TB_WAVNUM * EW_MAXF + m] * tpiinv * W.tb[TB_XK2CG * EW_MAXF + m];
        double bth[KPL], bmax = 0.0;
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
          bth[j] = 0.0;
          if (!kv[j]) continue;
          const int k = lane + 32 * j;
          double b = 0.0;
          for (int k2 = 0; k2 < ns; ++k2) b += __ldg(d.tab.satweights + k2 * A + k) * fl[m * A + __ldg(d.tab.indicessat + k2 * A + k)];
          bth[j] = b * facsat;
          bmax = fmax(bmax, bth[j]);
        }
        const double bth0 = wmax(bmax);
        const double ssdsc2_sig = c_dc.SSDSC2 * c_dc.ZPIFR[m];
        const double zcoef = ssdsc2_sig * c_dc.SSDSC6, zcoefm1 = ssdsc2_sig * ssdsc6m1;
#pragma unroll
        for (int j = 0; j < KPL; ++j) if (kv[j]) {
          const int o = m * A + lane + 32 * j;
          const double dd = zcoef * sq(fmax(0., bth0 * tmp03 - c_dc.SSDSC4)) + zcoefm1 * sq(fmax(0., bth[j] * tmp03 - c_dc.SSDSC4));
          sl[o] = sl[o] + dd * fl[o];
          fld[o] = fld[o] + dd;
        }
      }
    } else {
      // SDISSIP_JAN (sdissip_jan.F90:96-132)
      const double sds = c_dc.CDIS * c_dc.ZPI * f1mean * sq(emean) * p4(xkmean);
      const double cvis = c_dc.rnu * c_dc.CDISVIS;
      for (int m = 0; m < F; ++m) {
        const double wn = W.tb[TB_WAVNUM * EW_MAXF + m];
        const double x = wn / xkmean;
        const double temp1 = sds * x * ((1.0 - c_dc.DELTA_SDIS) + c_dc.DELTA_SDIS * x) + cvis * sq(wn);
#pragma unroll
        for (int j = 0; j < KPL; ++j) if (kv[j]) {
          const int o = m * A + lane + 32 * j;
          fld[o] = fld[o] + temp1;
          sl[o] = sl[o] + temp1 * fl[o];
        }
      }
    }
    __syncwarp();
    // ---- SNONLIN (snonlin.F90:116-498, ISNONLIN=0): DIA quadruplets, scatter form with warp-synchronous steps
    {
      double enhfr = fmax(0.75 * depth * akmean, 0.5);
      enhfr = 1.0 + (5.5 / enhfr) * (1.0 - .833 * enhfr) * exp(-1.25 * enhfr);
      const int MFR1STFR = -c_dc.MFRSTLW + 1;
      const int MFRLSTFR = F - c_dc.KFRH + MFR1STFR;
      for (int mc = 1; mc <= c_dc.MLSTHG; ++mc) {
        const int MP = c_dc.IKP[mc - 1], MP1 = c_dc.IKP1[mc - 1], MM = c_dc.IKM[mc - 1], MM1 = c_dc.IKM1[mc - 1];
        const int IC = c_dc.INLCOEF[mc - 1][0], IP = c_dc.INLCOEF[mc - 1][1], IP1 = c_dc.INLCOEF[mc - 1][2],
                  IM = c_dc.INLCOEF[mc - 1][3], IM1 = c_dc.INLCOEF[mc - 1][4];
        const double* R = c_dc.RNLCOEF[mc - 1];
        const double FTAIL = R[0], GW1 = R[1], GW2 = R[2], GW3 = R[3], GW4 = R[4];
        const double FKLAMPA = R[5], FKLAMPB = R[6], FKLAMP2 = R[7], FKLAMP1 = R[8];
        const double FKLAPA2 = R[9], FKLAPB2 = R[10], FKLAP12 = R[11], FKLAP22 = R[12];
        const double GW5 = R[13], GW6 = R[14], GW7 = R[15], GW8 = R[16];
        const double FKLAMMA = R[17], FKLAMMB = R[18], FKLAMM2 = R[19], FKLAMM1 = R[20];
        const double FKLAMA2 = R[21], FKLAMB2 = R[22], FKLAM12 = R[23], FKLAM22 = R[24];
        const double ftemp = c_dc.AF11[mc - 1] * enhfr;
        const int branch = (mc > MFR1STFR && mc < MFRLSTFR) ? 0 : (mc >= MFRLSTFR ? 1 : 2);
        bool do_c, do_mm, do_mm1, do_mp, do_mp1;
        if (branch == 0) { do_c = do_mm = do_mm1 = do_mp = do_mp1 = true; }
        else if (branch == 1) {
          do_mm = true; do_mm1 = MM1 <= F; do_c = do_mm1 && mc <= F; do_mp = do_c && MP <= F; do_mp1 = do_mp && MP1 <= F;
        } else { do_mm = false; do_mm1 = MM1 >= 1; do_c = true; do_mp = true; do_mp1 = true; }
        const double* fIP = fl + (IP - 1) * A; const double* fIP1 = fl + (IP1 - 1) * A;
        const double* fIM = fl + (IM - 1) * A; const double* fIM1 = fl + (IM1 - 1) * A;
        const double* fIC = fl + (IC - 1) * A;
        for (int kh = 0; kh < 2; ++kh) {
#pragma unroll
          for (int j = 0; j < KPL; ++j) {
            // all lanes run the step sequence (warp-synchronous); lanes without a direction contribute nothing
            const int k = lane + 32 * j;
            const bool v = k < A;
            if (j > 0 && A <= 32) break;
            const int kk = v ? k : 0;
            const int K1 = __ldg(d.tab.k1w + kh * A + kk), K2 = __ldg(d.tab.k2w + kh * A + kk);
            const int K11 = __ldg(d.tab.k11w + kh * A + kk), K21 = __ldg(d.tab.k21w + kh * A + kk);
            const double sap = GW1 * fIP[K1] + GW2 * fIP[K11] + GW3 * fIP1[K1] + GW4 * fIP1[K11];
            const double sam = GW5 * fIM[K2] + GW6 * fIM[K21] + GW7 * fIM1[K2] + GW8 * fIM1[K21];
            const double fij = (branch == 0) ? fIC[kk] : fIC[kk] * FTAIL;
            double fad1 = fij * (sap + sam);
            const double fad2 = fad1 - 2.0 * sap * sam;
            fad1 = fad1 + fad2;
            const double fcen = ftemp * fij;
            const double ad = fad2 * fcen;
            const double delad = fad1 * ftemp;
            const double delap = (fij - 2.0 * sam) * c_dc.DAL1 * fcen;
            const double delam = (fij - 2.0 * sap) * c_dc.DAL2 * fcen;
            // the nine updates; within one step every lane hits a distinct (direction, frequency) bin
            if (do_c) { if (v) { const int o = (mc - 1) * A + kk; sl[o] -= 2.0 * ad; fld[o] -= 2.0 * delad; } __syncwarp(); }
            if (do_mm) {
              if (v) { const int o = (MM - 1) * A + K2; sl[o] += ad * FKLAMM1; fld[o] += delam * FKLAM12; } __syncwarp();
              if (v) { const int o = (MM - 1) * A + K21; sl[o] += ad * FKLAMM2; fld[o] += delam * FKLAM22; } __syncwarp();
            }
            if (do_mm1) {
              if (v) { const int o = (MM1 - 1) * A + K2; sl[o] += ad * FKLAMMA; fld[o] += delam * FKLAMA2; } __syncwarp();
              if (v) { const int o = (MM1 - 1) * A + K21; sl[o] += ad * FKLAMMB; fld[o] += delam * FKLAMB2; } __syncwarp();
            }
            if (do_mp) {
              if (v) { const int o = (MP - 1) * A + K1; sl[o] += ad * FKLAMP1; fld[o] += delap * FKLAP12; } __syncwarp();
              if (v) { const int o = (MP - 1) * A + K11; sl[o] += ad * FKLAMP2; fld[o] += delap * FKLAP22; } __syncwarp();
            }
            if (do_mp1) {
              if (v) { const int o = (MP1 - 1) * A + K1; sl[o] += ad * FKLAMPA; fld[o] += delap * FKLAPA2; } __syncwarp();
              if (v) { const int o = (MP1 - 1) * A + K11; sl[o] += ad * FKLAMPB; fld[o] += delap * FKLAPB2; } __syncwarp();
            }
          }
        }
      }
    }
    __syncwarp();
    // ---- SSOURCE capture, SDIWBK, SBOTTOM, implicit update (implsch.F90:294-395), WNFLUXES sums
    const double delt = c_dc.delt, deltm = 1.0 / delt, delt5 = c_dc.ximp * delt;
    const double usfm = ufric * fmax(fmeanws, fmean);
    const double sds_bk = s[S_SDS * n + p];
    const bool brk = c_dc.lbiwbk && depth < 50.0;
    double philf = 0.0, xsoc = 0.0, ysoc = 0.0;
    const double sbo_const = -2.0 * 0.038 * c_dc.GM1;
    for (int m = 0; m < F; ++m) {
      const double tempm = usfm * (c_dc.COFRM4[m] * delt);
      double sbo = 0.0;
      if (m < c_dc.Fr && depth < c_dc.bathymax) {
        const double wn = W.tb[TB_WAVNUM * EW_MAXF + m];
        const double arg = fmin(2.0 * depth * wn, 50.0);
        sbo = sbo_const * wn / sinh(arg);
      }
      const double r = rhowgdfth(m, mij);
      const double cmr = W.tb[TB_CINV * EW_MAXF + m] * r;
      double sumt = 0.0, sumx = 0.0, sumy = 0.0;
#pragma unroll
      for (int j = 0; j < KPL; ++j) if (kv[j]) {
        const int o = m * A + lane + 32 * j;
        double slv = sl[o], fldv = fld[o];
        const double f0 = fl[o];
        double ssource = 0.0;
        if (c_dc.lcflx && c_dc.lwvflx_snl) ssource = slv / fmax(1.0 - delt5 * fldv, 1.0);
        if (m < c_dc.Fr) {
          if (brk) { slv = slv - sds_bk * f0; fldv = fldv - sds_bk; }   // SDIWBK
          slv = slv + sbo * f0; fldv = fldv + sbo;                      // SBOTTOM
        }
        const double gtemp1 = fmax(1.0 - delt5 * fldv, 1.0);
        const double gtemp2 = delt * slv / gtemp1;
        const double flhab = fmin(fabs(gtemp2), tempm);
        double fn = f0 + copysign(flhab, gtemp2);
        fn = fmax(fn, flm[j]);
        ssource = ssource + deltm * fmin(c_dc.FLMAX[m] - fn, 0.0);
        fn = fmin(fn, c_dc.FLMAX[m]);
        fld[o] = fn;    // new spectrum parked in FLD (old FL1 no longer needed below, but keep fl intact until all read)
        sumt += ssource; sumx += sinth[j] * ssource; sumy += costh[j] * ssource;
      }
      philf += sumt * r; xsoc += sumx * cmr; ysoc += sumy * cmr;
    }
    __syncwarp();
    if (c_dc.lcflx) {
      philf = wsum(philf); xsoc = wsum(xsoc); ysoc = wsum(ysoc);
      if (active && lane == 0) { s[S_PHILF * n + p] = philf; s[S_XSTROC * n + p] = xsoc; s[S_YSTROC * n + p] = ysoc; }
    }
    // from here FLD holds the new FL1
    double* fn = fld;
    // ---- FEMEANWS on the new spectrum for WSEMEAN/WSFMEAN (implsch.F90:424-446), LWFLUX only
    if (c_dc.lwflux) {
      double e1 = 0.0, e2 = 0.0, el = 0.0;
      for (int m = 0; m < F; ++m)
#pragma unroll
        for (int j = 0; j < KPL; ++j) if (kv[j]) {
          const int o = m * A + lane + 32 * j;
          const double xf = W.xl[o] ? fn[o] : 0.0;
          e1 += c_dc.DFIM[m] * xf; e2 += c_dc.DFIMOFR[m] * xf;
          if (m == F - 1) el += xf;
        }
      e1 = wsum(e1); e2 = wsum(e2); el = wsum(el);
      const double em2 = c_dc.EPSMIN + e1 + DELT25 * el;
      const double fm2 = em2 / (c_dc.EPSMIN + e2 + c_dc.FRTAIL * c_dc.DELTH * el);
      if (active && lane == 0) {
        if (em2 < c_dc.WSEMEAN_MIN) { d.f.wsemean[p] = c_dc.WSEMEAN_MIN; d.f.wsfmean[p] = 2. * c_dc.FR[F - 1]; }
        else { d.f.wsemean[p] = em2; d.f.wsfmean[p] = fm2; }
      }
    }
    // ---- IMPHFTAIL (imphftail.F90:71-87)
    {
      const double temp1 = 1.0 / W.tb[TB_XK2CG * EW_MAXF + mij - 1] / W.tb[TB_WAVNUM * EW_MAXF + mij - 1];
      for (int m = mij; m < F; ++m) {
        double temp2 = 1.0 / W.tb[TB_XK2CG * EW_MAXF + m] / W.tb[TB_WAVNUM * EW_MAXF + m];
        temp2 = temp2 / temp1;
#pragma unroll
        for (int j = 0; j < KPL; ++j) if (kv[j]) {
          const int k = lane + 32 * j;
          fn[m * A + k] = fmax(temp2 * fn[(mij - 1) * A + k], flm[j]);
        }
      }
    }
    // ---- SETICE (setice.F90:64-86)
    if (c_dc.licerun && c_dc.lmaskice) {
      double cireduc, icefree;
      if (cicover > c_dc.cithrsh) { cireduc = fmax(c_dc.EPSMIN, 1.0 - cicover); icefree = 0.0; }
      else { cireduc = 0.0; icefree = 1.0; }
      const double temp = cireduc * c_dc.flmin;
      for (int m = 0; m < F; ++m)
#pragma unroll
        for (int j = 0; j < KPL; ++j) if (kv[j]) {
          const int o = m * A + lane + 32 * j;
          fn[o] = fn[o] * icefree + temp * sq(fmax(0.0, coswdif[j]));
        }
    }
    // ---- STOKESDRIFT (stokesdrift.F90:84-142)
    {
      double us = 0.0, vs = 0.0;
      for (int m = 0; m < c_dc.NFRE_ODD; ++m) {
        const double stfac = d.f.stokfac[idx3(d, p, m)] * c_dc.DFIM_SIM[m];
#pragma unroll
        for (int j = 0; j < KPL; ++j) if (kv[j]) {
          const double fac3 = stfac * fn[m * A + lane + 32 * j];
          us += fac3 * sinth[j]; vs += fac3 * costh[j];
        }
      }
      const double cst = 2.0 * c_dc.DELTH * c_dc.ZPI * c_dc.ZPI * c_dc.ZPI / c_dc.G * p4(c_dc.FR[c_dc.NFRE_ODD - 1]);
#pragma unroll
      for (int j = 0; j < KPL; ++j) if (kv[j]) {
        const double ff = fn[(c_dc.NFRE_ODD - 1) * A + lane + 32 * j];
        us += cst * sinth[j] * ff; vs += cst * costh[j] * ff;
      }
      us = wsum(us); vs = wsum(vs);
      if (c_dc.licerun && c_dc.lwamrsetci && cicover > c_dc.cithrsh) {
        const double wsw = d.f.wswave[p];
        us = 0.016 * wsw * sin(wdwave) * (1.0 - cicover);
        vs = 0.016 * wsw * cos(wdwave) * (1.0 - cicover);
      }
      us = fmin(fmax(us, -1.5), 1.5); vs = fmin(fmax(vs, -1.5), 1.5);
      if (active && lane == 0) { d.f.ustokes[p] = us; d.f.vstokes[p] = vs; d.f.mij[p] = mij; }
    }
    __syncthreads();
    // ---- cooperative store of FL1 and XLLWS (grid-point index fastest)
    {
      const int tot = nptb * AF;
      for (int idx = threadIdx.x; idx < tot; idx += blockDim.x) {
        const int i = idx % nptb, bin = idx / nptb;
        const long long pp = pb + i;
        const long long c = pp / d.P;
        const int ln = (int)(pp - c * d.P);
        const double* base = smem + (size_t)i * per_pt;
        const size_t go = (size_t)ln + (size_t)d.P * ((size_t)bin + (size_t)AF * (size_t)c);
        d.f.fl1[go] = base[AF + bin];
        const unsigned char* xl = (const unsigned char*)(base + 3 * AF + NTB * EW_MAXF);
        d.f.xllws[go] = xl[bin] ? 1.0 : 0.0;
      }
    }
  }
}

int launch_implsch_stage(const ImplDev& d, long long p0, long long np, int stage, cudaStream_t st) {
  if (np <= 0) return 0;
  const int A = d.A, F = d.F, AF = A * F;
  if (A > 32 * KPL) { ew_set_error("NANG too large for the warp mapping"); return ECWAM_B200_EINVAL; }
  const unsigned gs = (unsigned)((np + 127) / 128);
  const size_t sm1 = (size_t)8 * ((size_t)AF + NTB * EW_MAXF) * sizeof(double);
  const size_t sm2 = (size_t)6 * ((size_t)3 * AF + NTB * EW_MAXF + (AF + 7) / 8) * sizeof(double);
  static bool attr_done = false;
  if (!attr_done) {
    EW_CUDA_CHECK(cudaFuncSetAttribute(k_spec<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    EW_CUDA_CHECK(cudaFuncSetAttribute(k_spec<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  if (sm1 > 227 * 1024 || sm2 > 227 * 1024) { ew_set_error("spectrum too large for shared memory"); return ECWAM_B200_EINVAL; }
  switch (stage) {
    case 0: k_airsea1<<<gs, 128, 0, st>>>(d, p0, np); break;
    case 1: k_spec<1><<<(unsigned)((np + 7) / 8), 256, sm1, st>>>(d, p0, np); break;
    case 2: k_scalar2<<<gs, 128, 0, st>>>(d, p0, np); break;
    case 3: k_spec<2><<<(unsigned)((np + 5) / 6), 192, sm2, st>>>(d, p0, np); break;
    case 4: k_scalar4<<<gs, 128, 0, st>>>(d, p0, np); break;
    default: return ECWAM_B200_EINVAL;
  }
  return 0;
}

}  // namespace ew
