// Host-side, one-off construction of the small tables the kernels read: the stand-in for
// ecWAM's INIWCST / SETWAVPHYS / MFREDIR / INITMDL(:437-503) / TABU_SWELLFT / INIT_X0TAUHF / INIT_SDISS_ARDH /
// INISNONLIN(+NLWEIGT, JAFU) when the caller does not bring the module state itself.
// All arrays are 0-based std::vectors; Fortran lower bounds are kept as explicit offsets.
#include "../../include/ecwam_b200.h"
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <vector>

namespace {

struct HostTables {
  ecwam_b200_tables t;
  std::vector<double> fr, dfim, dfimofr, dfimfr, zpifr, fr5, cofrm4, flmax, rhowg_dfim, dfim_sim, th, costh, sinth;
  std::vector<double> satweights, swellft, wtauhf, rnlcoef, af11;
  std::vector<int> indicessat, ikp, ikp1, ikm, ikm1, k1w, k2w, k11w, k21w, inlcoef;
  std::vector<double> xk_gc, omega_gc, cm_gc, c2osqrtvg_gc, xkmsqrtvgoc2_gc, om3gmkm_gc, omxkm3_gc, delkcc_gc_ns, delkcc_omxkm3_gc, delkcc_gc;
  std::vector<double> cideac;
  double delta_theta_rn = 0.75;
};

inline long fnint(double x) { return std::lround(x); }   // Fortran NINT

// ---- src/ecwam/yowpcons.F90:19-79 + iniwcst.F90
void constants(ecwam_b200_tables& t) {
  const double pi = 4.0 * std::atan(1.0);
  t.g = 9.806; t.gm1 = 0.101978381; t.circ = 40007993.95;
  t.zpi = 2.0 * pi;
  t.zpi4gm1 = std::pow(t.zpi, 4) / t.g;
  t.zpi4gm2 = std::pow(t.zpi, 4) / (t.g * t.g);
  t.r_earth = t.circ / t.zpi * 1.0;
  t.rowaterm1 = 1.0 / 1000.0;
  t.epsmin = 0.1e-32; t.epsus = 1.0e-6; t.epsu10 = std::sqrt(1.0e-3);
  t.acd = 8.0e-4; t.bcd = 8.0e-5; t.cdmax = 0.0025;
  t.tauocmin = 0.01; t.tauocmax = 50.0; t.phiepsmin = -3276.80; t.phiepsmax = -0.05; t.wsemean_min = 0.001;
  t.fratio = 1.1; t.wetail = 0.25; t.frtail = 0.2; t.wp1tail = 1.0 / 3.0;
  t.xkappa = 0.40; t.xnlev = 10.0;
  t.swellf = 0.66; t.swellf2 = -0.018; t.swellf3 = 0.022; t.swellf5 = 1.2; t.swellf6 = 1.0; t.abmin = 0.3; t.abmax = 8.0;
  t.sdsbr = 9.0e-4; t.ssdsc2 = -2.2e-5; t.ssdsc3 = 0.0; t.ssdsc4 = 1.0; t.ssdsc6 = 0.3; t.miche = 1.0; t.ssdsc5 = 0.0;
  t.iab = 200; t.eps1 = 0.00001; t.jtot_tauhf = 19;
}

// ---- src/ecwam/setwavphys.F90:46-205
int wave_physics(const ecwam_b200_params& p, ecwam_b200_tables& t, double& alphapmax) {
  const bool gc = p.llgcbz0 != 0, ng = p.llnormagam != 0;
  t.zalp = 0.008; t.tailfactor = 2.5;
  t.alphamax = 0.11; t.acdlin = 0.0008; t.bcdlin = 0.00047;        // yowphys.F90:55, yowpcons.F90:58-59
  t.rn1_rn = 0.25; t.dthrn_a = (p.iphys == 0) ? 0.80 : 0.60; t.dthrn_u = (p.iphys == 0 || gc) ? 33.0 : 200.0;
  if (p.nang <= 24) { t.ang_gc_a = 0.40; t.ang_gc_b = 0.60; } else { t.ang_gc_a = 0.35; t.ang_gc_b = 0.65; }
  t.ang_gc_c = 3.0;
  if (p.iphys == 0) {
    t.alphamin = 0.0001; alphapmax = 0.03; t.tauwshelter = 0.0; t.tailfactor_pm = 0.0;
    if (gc) { t.alpha = 0.0055; t.chnkmin_u = 28.; t.betamaxoxkappa2 = ng ? 1.32 : 1.25; t.cdis = -1.3; t.delta_sdis = 0.6; t.cdisvis = -4.0; }
    else { t.alpha = 0.0065; t.chnkmin_u = 33.; t.betamaxoxkappa2 = 1.20; t.cdis = -1.33; t.delta_sdis = 0.5; t.cdisvis = 0.0; }
    t.egrcrv = 1108.0; t.afcrv = 4.0e-4; t.bfcrv = -3.0;
    t.swellf4 = 1.5e05; t.swellf7 = 3.6e05; t.z0rat = 0.04; t.z0tubmax = 0.0005;
  } else if (p.iphys == 1) {
    t.tailfactor_pm = 3.0;
    if (gc) {
      t.alpha = 0.0055; t.alphamin = 0.0001; t.chnkmin_u = 28.; alphapmax = 0.03; t.z0tubmax = 0.05; t.z0rat = 0.02;
      t.swellf4 = 1.15e05; t.swellf7 = 4.32e05;
      if (ng) { t.betamaxoxkappa2 = 1.39; t.tauwshelter = 0.0; } else { t.betamaxoxkappa2 = 1.44; t.tauwshelter = 0.25; }
    } else {
      t.alpha = 0.0065; alphapmax = 0.031; t.z0tubmax = 0.0005; t.z0rat = 0.04; t.swellf4 = 1.5e05; t.swellf7 = 3.6e05;
      if (ng) { t.betamaxoxkappa2 = 1.39; t.tauwshelter = 0.0; t.alphamin = 0.0005; t.chnkmin_u = 30.; }
      else { t.betamaxoxkappa2 = 1.40; t.tauwshelter = 0.25; t.alphamin = 0.0001; t.chnkmin_u = 33.; }
    }
    t.egrcrv = 1065.0; t.afcrv = 2.453e-4; t.bfcrv = -3.1236;
    t.cdis = -1.33; t.delta_sdis = 0.5; t.cdisvis = 0.0;
  } else return ECWAM_B200_EINVAL;
  t.swellf7m1 = 1.0 / t.swellf7;
  t.alphapmax = alphapmax;
  // betamaxoxkappa2 holds BETAMAX until INIT_X0TAUHF divides it by XKAPPA**2
  return 0;
}

// ---- src/ecwam/initgc.F90:63-110 with gc_dispersion.h: the gravity-capillary wavenumber grid of STRESS_GC / OMEGAGC
double ipow(double x, int m) {   // real ** integer by repeated squaring
  unsigned n = (unsigned)(m < 0 ? -m : m);
  double y = (n & 1) ? x : 1.0;
  while (n >>= 1) { x = x * x; if (n & 1) y = y * x; }
  return m < 0 ? 1.0 / y : y;
}
void gravity_capillary_tables(HostTables& h) {
  ecwam_b200_tables& t = h.t;
  const double KRATIO = 1.2, XKS = 0.006, XKL = 20000.0, SURFT = 0.0717 / 1000.0;   // yowfred.F90:62-65, iniwcst.F90:69
  t.sqrtgosurft = std::sqrt(t.g / SURFT);
  const int N = (int)fnint(std::log(XKL / XKS) / std::log(KRATIO));
  t.nwav_gc = N;
  std::vector<double> xkm(N), vg(N), c(N), delkcc(N);
  for (auto* v : {&h.xk_gc, &h.omega_gc, &h.cm_gc, &h.c2osqrtvg_gc, &h.xkmsqrtvgoc2_gc, &h.om3gmkm_gc, &h.omxkm3_gc, &h.delkcc_gc_ns,
                  &h.delkcc_omxkm3_gc}) v->assign(N, 0.0);
  for (int i = 0; i < N; ++i) {
    const double k = XKS * ipow(KRATIO, i);
    h.xk_gc[i] = k;
    xkm[i] = 1.0 / k;
    const double om = std::sqrt(t.g * k + SURFT * (k * k * k));
    h.omega_gc[i] = om;
    h.omxkm3_gc[i] = om * (xkm[i] * xkm[i] * xkm[i]);
    vg[i] = 0.5 / om * (t.g + 3.0 * SURFT * (k * k));
    c[i] = om / k;
    h.cm_gc[i] = 1.0 / c[i];
    h.c2osqrtvg_gc[i] = (c[i] * c[i]) / std::sqrt(vg[i]);
    h.xkmsqrtvgoc2_gc[i] = xkm[i] / h.c2osqrtvg_gc[i];
    h.om3gmkm_gc[i] = (om * om * om) / (t.g * k);
  }
  delkcc[0] = 0.5 * (h.xk_gc[1] - h.xk_gc[0]) / h.c2osqrtvg_gc[0];
  h.delkcc_gc_ns[0] = delkcc[0];
  for (int i = 1; i < N - 1; ++i) {
    delkcc[i] = 0.5 * (h.xk_gc[i + 1] - h.xk_gc[i - 1]) / h.c2osqrtvg_gc[i];
    h.delkcc_gc_ns[i] = 0.5 * (h.xk_gc[i + 1] - h.xk_gc[i]) / h.c2osqrtvg_gc[i];
  }
  delkcc[N - 1] = 0.5 * (h.xk_gc[N - 1] - h.xk_gc[N - 2]) / h.c2osqrtvg_gc[N - 1];
  h.delkcc_gc_ns[N - 1] = delkcc[N - 1];
  for (int i = 0; i < N; ++i) h.delkcc_omxkm3_gc[i] = delkcc[i] * h.omxkm3_gc[i];
  h.delkcc_gc = delkcc;
}

// ---- Kelvin functions through the modified Bessel functions of complex argument
//      (src/ecwam/kerkei.F90, kzeone.F90 = ACM Algorithm 484, Burrell 1974).  Returns exp(x)*K0(z), exp(x)*K1(z).
void kzeone(double X, double Y, std::complex<double>& K0, std::complex<double>& K1) {
  static const double EXSQ[8] = {0.5641003087264E0, 0.4120286874989E0, 0.1584889157959E0, 0.3078003387255E-1,
                                 0.2778068842913E-2, 0.1000044412325E-3, 0.1059115547711E-5, 0.1522475804254E-8};
  static const double TSQ[8] = {0.0E0, 3.19303633920635E-1, 1.29075862295915E0, 2.95837445869665E0,
                                5.40903159724444E0, 8.80407957805676E0, 1.34685357432515E1, 2.02499163658709E1};
  double re0, im0, re1, im1;
  double r2 = X * X + Y * Y;
  if (r2 >= 1.96e2) {                       // asymptotic expansion
    double rterm = 1.0, iterm = 0.0;
    re0 = 1.0; im0 = 0.0; re1 = 1.0; im1 = 0.0;
    double p1 = 8.0 * r2, p2 = std::sqrt(r2);
    const int nl = (int)fnint(3.91 + 8.12e1 / p2);
    double r1 = 1.0; r2 = 1.0;
    int mm = -8, kk = 3;
    for (int n = 1; n <= nl; ++n) {
      mm += 8; kk -= mm;
      r1 = (double)(kk - 4) * r1; r2 = (double)kk * r2;
      const double t1 = (double)n * p1, t2 = rterm;
      rterm = (t2 * X + iterm * Y) / t1;
      iterm = (-t2 * Y + iterm * X) / t1;
      re0 += r1 * rterm; im0 += r1 * iterm; re1 += r2 * rterm; im1 += r2 * iterm;
    }
    double t1 = std::sqrt(p2 + X), t2 = -Y / t1;
    p1 = 8.86226925452758e-1 / p2;
    rterm = p1 * std::cos(Y); iterm = -p1 * std::sin(Y);
    r1 = re0 * rterm - im0 * iterm; r2 = re0 * iterm + im0 * rterm;
    re0 = t1 * r1 - t2 * r2; im0 = t1 * r2 + t2 * r1;
    r1 = re1 * rterm - im1 * iterm; r2 = re1 * iterm + im1 * rterm;
    re1 = t1 * r1 - t2 * r2; im1 = t1 * r2 + t2 * r1;
  } else if (r2 >= 1.849e1) {               // Gauss-type quadrature
    const double x2 = 2.0 * X, y2 = 2.0 * Y;
    double r1 = y2 * y2;
    double p1 = std::sqrt(x2 * x2 + r1), p2 = std::sqrt(p1 + x2);
    double t1 = EXSQ[0] / (2.0 * p1);
    re0 = t1 * p2; im0 = t1 / p2; re1 = 0.0; im1 = 0.0;
    for (int n = 1; n < 8; ++n) {
      const double t2 = x2 + TSQ[n];
      p1 = std::sqrt(t2 * t2 + r1); p2 = std::sqrt(p1 + t2);
      t1 = EXSQ[n] / p1;
      re0 += t1 * p2; im0 += t1 / p2;
      t1 = EXSQ[n] * TSQ[n];
      re1 += t1 * p2; im1 += t1 / p2;
    }
    double t2 = -y2 * im0;
    re1 = re1 / r2;
    r2 = y2 * im1 / r2;
    const double rterm = 1.41421356237309e0 * std::cos(Y), iterm = -1.41421356237309e0 * std::sin(Y);
    im0 = re0 * iterm + t2 * rterm;
    re0 = re0 * rterm - t2 * iterm;
    t1 = re1 * rterm - r2 * iterm;
    t2 = re1 * iterm + r2 * rterm;
    re1 = t1 * X + t2 * Y;
    im1 = -t1 * Y + t2 * X;
  } else {                                   // power series
    double x2 = X / 2.0, y2 = Y / 2.0;
    double p1 = x2 * x2, p2 = y2 * y2;
    double t1 = -(std::log(p1 + p2) / 2.0 + 0.5772156649015329e0);
    const double t2 = -std::atan2(Y, X);
    x2 = p1 - p2; y2 = X * y2;
    double rterm = 1.0, iterm = 0.0;
    re0 = t1; im0 = t2;
    t1 = t1 + 0.5;
    re1 = t1; im1 = t2;
    p2 = std::sqrt(r2);
    double el = 2.106 * p2 + 4.4;
    if (p2 < 8.0e-1) el = 2.129 * p2 + 4.0;
    const int nl = (int)fnint(el);
    for (int n = 1; n <= nl; ++n) {
      p1 = n; p2 = (double)n * n;
      const double r1 = rterm;
      rterm = (r1 * x2 - iterm * y2) / p2;
      iterm = (r1 * y2 + iterm * x2) / p2;
      t1 = t1 + 0.5 / p1;
      re0 = re0 + t1 * rterm - t2 * iterm;
      im0 = im0 + t1 * iterm + t2 * rterm;
      p1 = p1 + 1.0;
      t1 = t1 + 0.5 / p1;
      re1 = re1 + (t1 * rterm - t2 * iterm) / p1;
      im1 = im1 + (t1 * iterm + t2 * rterm) / p1;
    }
    const double r1 = X / r2 - 0.5 * (X * re1 - Y * im1);
    const double rr2 = -Y / r2 - 0.5 * (X * im1 + Y * re1);
    p1 = std::exp(X);
    re0 = p1 * re0; im0 = p1 * im0; re1 = p1 * r1; im1 = p1 * rr2;
  }
  K0 = {re0, im0};
  K1 = {re1, im1};
}

// ---- src/ecwam/tabu_swellft.F90:64-83: friction factor table for the swell dissipation
void swell_friction_table(HostTables& h) {
  const int IAB = h.t.iab, NITER = 100;
  const double ABMIN = 0.3, ABMAX = 8.0, KAPPA = 0.40;
  h.swellft.assign(IAB, 0.0);
  const double delab = (ABMAX - ABMIN) / (double)IAB, l10 = std::log(10.0);
  double dzeta0 = 0.0;
  for (int i = 1; i <= IAB; ++i) {
    const double abr = std::exp((ABMIN + (double)i * delab) * l10);
    const double fact = 1 / abr / (21.2 * KAPPA);
    double fsubw = 0.05;
    for (int it = 0; it < NITER; ++it) {
      const double fsubw_memo = fsubw, dzeta0_memo = dzeta0;
      dzeta0 = fact * std::pow(fsubw, -0.5);
      const double x = 2.0 * std::sqrt(dzeta0);
      const double zr = x * 0.50 * std::sqrt(2.0);
      std::complex<double> k0, k1;
      kzeone(zr, zr, k0, k1);
      const double ker = k0.real() / std::exp(zr), kei = k0.imag() / std::exp(zr);
      fsubw = 0.08 / (ker * ker + kei * kei);
      fsubw = 0.5 * (fsubw_memo + fsubw);
      dzeta0 = 0.5 * (dzeta0_memo + dzeta0);
    }
    h.swellft[i - 1] = fsubw;
  }
}

// ---- src/ecwam/mfr.F90, mfredir.F90:90-129, initmdl.F90:437-503
void frequency_direction_tables(const ecwam_b200_params& p, int ifre1, double fr1, double alphapmax, HostTables& h) {
  ecwam_b200_tables& t = h.t;
  const int F = p.nfre, A = p.nang;
  const double pi = 4.0 * std::atan(1.0);
  h.fr.assign(F, 0.0);
  h.fr[ifre1 - 1] = fr1;
  for (int m = ifre1 - 2; m >= 0; --m) h.fr[m] = h.fr[m + 1] / t.fratio;
  for (int m = ifre1; m < F; ++m) h.fr[m] = t.fratio * h.fr[m - 1];
  t.delth = t.zpi / (double)A;
  h.th.resize(A); h.costh.resize(A); h.sinth.resize(A);
  for (int k = 0; k < A; ++k) {
    h.th[k] = (double)k * t.delth + 0.5 * t.delth;
    h.costh[k] = std::cos(h.th[k]);
    h.sinth[k] = std::sin(h.th[k]);
  }
  const double co1 = 0.5 * (t.fratio - 1.0) * t.delth;
  h.dfim.resize(F);
  h.dfim[0] = co1 * h.fr[0];
  for (int m = 1; m < F - 1; ++m) h.dfim[m] = co1 * (h.fr[m] + h.fr[m - 1]);
  h.dfim[F - 1] = co1 * h.fr[F - 2];
  h.dfimofr.resize(F); h.dfimfr.resize(F); h.zpifr.resize(F); h.fr5.resize(F); h.cofrm4.resize(F); h.flmax.resize(F);
  h.rhowg_dfim.resize(F); h.dfim_sim.assign(F, 0.0);
  const double COEF4 = 5.0e-07, ROWATER = 1000.0;
  for (int m = 0; m < F; ++m) {
    h.dfimofr[m] = h.dfim[m] / h.fr[m];
    h.dfimfr[m] = h.dfim[m] * h.fr[m];
    h.zpifr[m] = t.zpi * h.fr[m];
    h.fr5[m] = std::pow(h.fr[m], 5);
    h.cofrm4[m] = COEF4 * t.g / std::pow(h.fr[m], 4);
    h.flmax[m] = (alphapmax / pi) / (t.zpi4gm2 * h.fr5[m]);
  }
  t.flogsprdm1 = 1.0 / std::log10(t.fratio);
  const double xlogfratio = std::log(t.fratio);
  for (int m = 0; m < F; ++m) {
    const double w = (m == 0 || m == F - 1) ? 0.5 : 1.0;
    h.rhowg_dfim[m] = (w == 1.0) ? ROWATER * t.g * t.delth * xlogfratio * h.fr[m]
                                 : 0.5 * ROWATER * t.g * t.delth * xlogfratio * h.fr[m];
  }
  t.nfre_odd = F - 1 + (F % 2);
  const int NO = t.nfre_odd;
  h.dfim_sim[0] = t.delth * xlogfratio * h.fr[0] / 3.0;
  for (int M = 2; M <= NO - 1; M += 2) {   // 1-based as in the reference
    h.dfim_sim[M - 1] = 4.0 * t.delth * xlogfratio * h.fr[M - 1] / 3.0;
    h.dfim_sim[M] = 2.0 * t.delth * xlogfratio * h.fr[M] / 3.0;
  }
  h.dfim_sim[NO - 1] = t.delth * xlogfratio * h.fr[NO - 1] / 3.0;
}

// ---- src/ecwam/init_x0tauhf.F90:65-100
void hf_stress_tables(const ecwam_b200_params& p, HostTables& h) {
  ecwam_b200_tables& t = h.t;
  const double betamax = t.betamaxoxkappa2;
  t.betamaxoxkappa2 = betamax / (t.xkappa * t.xkappa);
  t.bmaxokap = h.delta_theta_rn * t.betamaxoxkappa2 / t.xkappa;
  t.gamnconst = t.bmaxokap * 0.5 * std::pow(t.zpi, 4) * std::pow(t.gm1, 3);
  const double alph = (p.llgcbz0 || p.llcapchnk || p.llnormagam) ? t.alphamin : t.alpha;
  double x0 = 0.005;
  for (int j = 0; j < 30; ++j) {
    const double ff = std::exp(t.xkappa / (x0 + t.zalp));
    const double f = alph * x0 * x0 * ff - 1.0;
    if (f == 0.0) break;
    const double q = x0 / (x0 + t.zalp);
    const double df = alph * ff * (2.0 * x0 - t.xkappa * (q * q));
    x0 = x0 - f / df;
  }
  t.x0tauhf = x0;
  const int J = t.jtot_tauhf;
  h.wtauhf.assign(J, 0.0);
  const double c1 = t.betamaxoxkappa2 / 3.0;
  h.wtauhf[0] = c1;
  for (int j = 2; j <= J - 1; j += 2) { h.wtauhf[j - 1] = 4.0 * c1; h.wtauhf[j] = 2.0 * c1; }
  h.wtauhf[J - 1] = c1;
}

// ---- src/ecwam/init_sdiss_ardh.F90:69-96: directional window of the saturation spectrum
void saturation_tables(const ecwam_b200_params& p, HostTables& h) {
  ecwam_b200_tables& t = h.t;
  const int A = p.nang, ISDSDTH = 80;
  const double rad = (4.0 * std::atan(1.0)) / 180.0;
  t.nsdsnth = (int)std::min<long>(fnint(ISDSDTH * rad / t.delth), A / 2 - 1);
  const int N = t.nsdsnth, ns = 2 * N + 1;
  double dtr = (h.th[0] + ISDSDTH * rad) - (h.th[N] - 0.5 * t.delth);
  dtr = std::max(0.0, std::min(dtr, t.delth));
  h.indicessat.assign((size_t)A * ns, 0);
  h.satweights.assign((size_t)A * ns, 0.0);
  for (int k = 1; k <= A; ++k)
    for (int ii = k - N; ii <= k + N; ++ii) {
      int jj = ii;
      if (ii < 1) jj = ii + A;
      if (ii > A) jj = ii - A;
      const int col = ii - (k - N);   // 0-based second index
      h.indicessat[(k - 1) + (size_t)A * col] = jj;
      const double dl = (ii == k - N || ii == k + N) ? dtr : t.delth;
      const double cs = std::cos(h.th[k - 1] - h.th[jj - 1]);
      h.satweights[(k - 1) + (size_t)A * col] = dl * (cs * cs);
    }
}

// ---- src/ecwam/jafu.F90
int jafu(double cl, int j, int ian) {
  const int idph = (int)cl;
  int ja = j + idph;
  if (ja <= 0) ja = ian + ja - 1;
  if (ja >= ian) ja = ja - ian + 1;
  return ja;
}

// ---- src/ecwam/nlweigt.F90:94-262 + inisnonlin.F90:89-270: discrete-interaction (DIA) index and weight tables
void dia_tables(const ecwam_b200_params& p, HostTables& h) {
  ecwam_b200_tables& t = h.t;
  const int A = p.nang, F = p.nfre;
  const double ALAMD = 0.25, CON = 3000.0;
  const double pi = 4.0 * std::atan(1.0), deg = 180. / pi;
  const double f1p1 = std::log10(t.fratio);
  const int isp = (int)(std::log10(1.0 + ALAMD) / f1p1 + .000001);
  const int ism = (int)std::floor(std::log10(1.0 - ALAMD) / f1p1 + .0000001);
  t.mfrstlw = 1 + ism;
  t.mlsthg = F - ism;
  t.kfrh = -ism + isp + 2;
  const int LO = t.mfrstlw, HI = t.mlsthg, NW = HI - LO + 1;
  auto W = [LO](int m) { return m - LO; };   // position of frequency index m in a (MFRSTLW:MLSTHG) array
  // angular offsets
  const double xf = std::pow((1.0 + ALAMD) / (1.0 - ALAMD), 4);
  const double costh3 = (1.0 + 2.0 * ALAMD + 2.0 * ALAMD * ALAMD * ALAMD) / ((1.0 + ALAMD) * (1.0 + ALAMD));
  const double delphi1 = -180.0 / pi * std::acos(costh3);
  const double costh4 = std::sqrt(1.0 - xf + xf * costh3 * costh3);
  const double delphi2 = 180.0 / pi * std::acos(costh4);
  const double deltha = t.delth * deg;
  double cl1 = delphi1 / deltha, cl2 = delphi2 / deltha;
  std::vector<int> ja1((size_t)A * 2, 0), ja2((size_t)A * 2, 0);
  const int klp1 = A + 1;
  for (int kh = 1, ic = 1; kh <= 2; ++kh, ic = -1) {
    const int klh = (kh == 2) ? klp1 : A;
    for (int k = 1; k <= klh; ++k) {
      const int ks = (kh > 1) ? klp1 - k + 1 : k;
      if (ks > A) continue;
      ja1[(ks - 1) + A * (kh - 1)] = jafu(ic * cl1, k, klp1);
      ja2[(ks - 1) + A * (kh - 1)] = jafu(ic * cl2, k, klp1);
    }
  }
  cl1 = cl1 - (int)cl1;
  cl2 = cl2 - (int)cl2;
  const double acl1 = std::fabs(cl1), acl2 = std::fabs(cl2), cl11 = 1.0 - acl1, cl21 = 1.0 - acl2;
  t.dal1 = 1.0 / std::pow(1.0 + ALAMD, 4);
  t.dal2 = 1.0 / std::pow(1.0 - ALAMD, 4);
  h.k1w.assign((size_t)A * 2, 0); h.k2w = h.k1w; h.k11w = h.k1w; h.k21w = h.k1w;
  for (int kh = 1, isg = 1; kh <= 2; ++kh, isg = -1) {
    const double cl1h = isg * cl1, cl2h = isg * cl2;
    for (int k = 1; k <= A; ++k) {
      int ks = (kh == 2) ? A - k + 2 : k;
      if (k == 1) ks = 1;
      const int k1 = ja1[(k - 1) + A * (kh - 1)], k2 = ja2[(k - 1) + A * (kh - 1)];
      int k11, k21;
      if (cl1h < 0.0) { k11 = k1 - 1; if (k11 < 1) k11 = A; } else { k11 = k1 + 1; if (k11 > A) k11 = 1; }
      if (cl2h < 0) { k21 = k2 - 1; if (k21 < 1) k21 = A; } else { k21 = k2 + 1; if (k21 > A) k21 = 1; }
      const size_t o = (ks - 1) + (size_t)A * (kh - 1);
      h.k1w[o] = k1; h.k11w[o] = k11; h.k2w[o] = k2; h.k21w[o] = k21;
    }
  }
  // extended frequency axis MFRSTLW .. NFRE+KFRH
  const int FLO = LO, FHI = F + t.kfrh;
  std::vector<double> frlon(FHI - FLO + 1);
  auto FL = [&](int m) -> double& { return frlon[m - FLO]; };
  for (int m = 1; m <= F; ++m) FL(m) = h.fr[m - 1];
  for (int m = 0; m >= LO; --m) FL(m) = FL(m + 1) / t.fratio;
  for (int m = F + 1; m <= FHI; ++m) FL(m) = t.fratio * FL(m - 1);
  h.ikp.assign(NW, 0); h.ikp1 = h.ikp; h.ikm = h.ikp; h.ikm1 = h.ikp;
  h.af11.assign(NW, 0.0);
  std::vector<double> fklap(NW), fklap1(NW), fklam(NW), fklam1(NW);
  for (int m = LO; m <= HI; ++m) {
    const double frg = FL(m);
    h.af11[W(m)] = CON * std::pow(frg, 11);
    const double flp = frg * (1.0 + ALAMD), flm = frg * (1.0 - ALAMD);
    h.ikp[W(m)] = m + isp;
    const double fkp = FL(h.ikp[W(m)]);
    h.ikp1[W(m)] = h.ikp[W(m)] + 1;
    fklap[W(m)] = (flp - fkp) / (FL(h.ikp1[W(m)]) - fkp);
    fklap1[W(m)] = 1.0 - fklap[W(m)];
    const int ikn = m + ism;
    if (ikn >= LO) {
      h.ikm[W(m)] = ikn;
      const double fkm = FL(ikn);
      h.ikm1[W(m)] = ikn + 1;
      fklam[W(m)] = (flm - fkm) / (FL(ikn + 1) - fkm);
      fklam1[W(m)] = 1.0 - fklam[W(m)];
    } else if (ikn + 1 == LO) {
      h.ikm[W(m)] = 1;
      h.ikm1[W(m)] = LO;
      const double fkm = FL(LO) / t.fratio;
      fklam[W(m)] = (flm - fkm) / (FL(LO) - fkm);
      fklam1[W(m)] = 0.0;
    } else {
      h.ikm[W(m)] = 1; fklam[W(m)] = 0.0; h.ikm1[W(m)] = 1; fklam1[W(m)] = 0.0;
    }
  }
  std::vector<double> frh(t.kfrh);
  for (int i = 1; i <= t.kfrh; ++i) frh[i - 1] = std::pow(FL(F) / FL(F + i - 1), 5);
  // inisnonlin: low-frequency spectral extension factors and the packed coefficients
  auto epmma = [](double x) { return std::exp(-std::min(1.25 * std::pow(x, 4), 50.0)) * std::pow(x, 5); };
  std::vector<double> ftrf(1 - LO + 1);   // (MFRSTLW:1)
  {
    const double alph = 1.0 / epmma(1.0);
    double frr = 1.0;
    for (int mc = 1; mc >= LO; --mc) { ftrf[mc - LO] = alph * epmma(frr); frr = frr * t.fratio; }
  }
  h.inlcoef.assign((size_t)5 * HI, 0);
  h.rnlcoef.assign((size_t)25 * HI, 0.0);
  for (int mc = 1; mc <= HI; ++mc) {
    int ip = h.ikp[W(mc)], ip1 = h.ikp1[W(mc)], im = h.ikm[W(mc)], im1 = h.ikm1[W(mc)], ic = std::max(mc, 1);
    double ffacp = 1.0, ffacp1 = 1.0, ffacm = 1.0, ffacm1 = 1.0, ftail = 1.0;
    if (ip < 1) { ffacp = ftrf[ip - LO]; ip = 1; }
    if (ip1 < 1) { ffacp1 = ftrf[ip1 - LO]; ip1 = 1; }
    if (im < LO) { ffacm = 0.0; im = 1; } else if (im < 1) { ffacm = ftrf[im - LO]; im = 1; }
    if (im1 < LO) { ffacm1 = 0.0; im1 = 1; } else if (im1 < 1) { ffacm1 = ftrf[im1 - LO]; im1 = 1; }
    if (ip1 > F) {
      const int it = std::min(ip1 - F + 1, t.kfrh);
      ffacp1 = frh[it - 1]; ip1 = F;
      if (ip > F) {
        ffacp = frh[ip - F]; ip = F;
        if (ic > F) {
          ftail = frh[ic - F]; ic = F;
          if (im1 > F) { ffacm1 = frh[im1 - F]; im1 = F; }
        }
      }
    }
    int* ii = &h.inlcoef[(size_t)5 * (mc - 1)];
    ii[0] = ic; ii[1] = ip; ii[2] = ip1; ii[3] = im; ii[4] = im1;
    double* r = &h.rnlcoef[(size_t)25 * (mc - 1)];
    const double fklamp = fklap[W(mc)];
    double fklamp1 = fklap1[W(mc)];
    double gw2 = fklamp1 * ffacp * t.dal1;
    const double gw1 = gw2 * cl11;
    gw2 = gw2 * acl1;
    double gw4 = fklamp * ffacp1 * t.dal1;
    const double gw3 = gw4 * cl11;
    gw4 = gw4 * acl1;
    const double fklampa = fklamp * cl11, fklampb = fklamp * acl1, fklamp2 = fklamp1 * acl1;
    fklamp1 = fklamp1 * cl11;
    r[0] = ftail; r[1] = gw1; r[2] = gw2; r[3] = gw3; r[4] = gw4;
    r[5] = fklampa; r[6] = fklampb; r[7] = fklamp2; r[8] = fklamp1;
    r[9] = fklampa * fklampa; r[10] = fklampb * fklampb; r[11] = fklamp1 * fklamp1; r[12] = fklamp2 * fklamp2;
    const double fklamm = fklam[W(mc)];
    double fklamm1 = fklam1[W(mc)];
    double gw6 = fklamm1 * ffacm * t.dal2;
    const double gw5 = gw6 * cl21;
    gw6 = gw6 * acl2;
    double gw8 = fklamm * ffacm1 * t.dal2;
    const double gw7 = gw8 * cl21;
    gw8 = gw8 * acl2;
    const double fklamma = fklamm * cl21, fklammb = fklamm * acl2, fklamm2 = fklamm1 * acl2;
    fklamm1 = fklamm1 * cl21;
    r[13] = gw5; r[14] = gw6; r[15] = gw7; r[16] = gw8;
    r[17] = fklamma; r[18] = fklammb; r[19] = fklamm2; r[20] = fklamm1;
    r[21] = fklamma * fklamma; r[22] = fklammb * fklammb; r[23] = fklamm1 * fklamm1; r[24] = fklamm2 * fklamm2;
  }
}

// ---- src/ecwam/cigetdeac.F90:60-75: ln of the attenuation per floe as a function of wave period (1..16 s) and ice thickness
// (0.2..3.7 m).  Periods 6..16 s are Kohout & Meylan's data; the 1 s column is assumed (-2 at 0.2 m to -1 at 3.7 m, linear in the
// thickness) and the periods in between are a straight line from the 1 s to the 6 s value.
void ice_attenuation_table(HostTables& h) {
  static const double data[36][11] = {
#include "kohout_meylan_fig6.inc"
  };
  ecwam_b200_tables& t = h.t;
  t.nict = 16; t.nich = 36; t.ticmin = 1.0; t.hicmin = 0.2; t.dtic = 1.0; t.dhic = 0.1;
  h.cideac.assign((size_t)t.nict * t.nich, 0.0);
  for (int ih = 0; ih < t.nich; ++ih) {
    double* col = &h.cideac[(size_t)t.nict * ih];      // CIDEAC(:, ih+1)
    col[0] = (ih == 0) ? -2.0 : (ih == t.nich - 1) ? -1.0 : -2.0 + ih * (-1.0 - -2.0) / (t.nich - 1);
    for (int it = 5; it < t.nict; ++it) col[it] = data[ih][it - 5];
    const double dci = col[5] - col[0];
    for (int it = 1; it < 5; ++it) col[it] = col[0] + dci * it * t.dtic / (5 * t.dtic);
  }
}

}  // namespace

struct ecwam_b200_host_tables_s { HostTables h; };

extern "C" {

int ecwam_b200_host_tables_create(const ecwam_b200_params* params, int ifre1, double fr1, ecwam_b200_host_tables_t* out) {
  if (!params || !out || ifre1 < 1 || ifre1 > params->nfre || params->nang < 4 || params->nfre < 6) return ECWAM_B200_EINVAL;
  auto* o = new ecwam_b200_host_tables_s();
  HostTables& h = o->h;
  std::memset(&h.t, 0, sizeof(h.t));
  constants(h.t);
  double alphapmax = 0.0;
  if (wave_physics(*params, h.t, alphapmax)) { delete o; return ECWAM_B200_EINVAL; }
  frequency_direction_tables(*params, ifre1, fr1, alphapmax, h);
  swell_friction_table(h);
  hf_stress_tables(*params, h);
  saturation_tables(*params, h);
  dia_tables(*params, h);
  gravity_capillary_tables(h);
  ice_attenuation_table(h);
  ecwam_b200_tables& t = h.t;
  t.cideac = h.cideac.data();
  t.xk_gc = h.xk_gc.data(); t.omega_gc = h.omega_gc.data(); t.cm_gc = h.cm_gc.data(); t.c2osqrtvg_gc = h.c2osqrtvg_gc.data();
  t.xkmsqrtvgoc2_gc = h.xkmsqrtvgoc2_gc.data(); t.om3gmkm_gc = h.om3gmkm_gc.data(); t.omxkm3_gc = h.omxkm3_gc.data();
  t.delkcc_gc_ns = h.delkcc_gc_ns.data(); t.delkcc_omxkm3_gc = h.delkcc_omxkm3_gc.data(); t.delkcc_gc = h.delkcc_gc.data();
  t.fr = h.fr.data(); t.dfim = h.dfim.data(); t.dfimofr = h.dfimofr.data(); t.dfimfr = h.dfimfr.data();
  t.zpifr = h.zpifr.data(); t.fr5 = h.fr5.data(); t.cofrm4 = h.cofrm4.data(); t.flmax = h.flmax.data();
  t.rhowg_dfim = h.rhowg_dfim.data(); t.dfim_sim = h.dfim_sim.data(); t.th = h.th.data(); t.costh = h.costh.data();
  t.sinth = h.sinth.data(); t.indicessat = h.indicessat.data(); t.satweights = h.satweights.data();
  t.swellft = h.swellft.data(); t.wtauhf = h.wtauhf.data(); t.ikp = h.ikp.data(); t.ikp1 = h.ikp1.data();
  t.ikm = h.ikm.data(); t.ikm1 = h.ikm1.data(); t.k1w = h.k1w.data(); t.k2w = h.k2w.data(); t.k11w = h.k11w.data();
  t.k21w = h.k21w.data(); t.inlcoef = h.inlcoef.data(); t.rnlcoef = h.rnlcoef.data(); t.af11 = h.af11.data();
  *out = o;
  return 0;
}
const ecwam_b200_tables* ecwam_b200_host_tables_get(ecwam_b200_host_tables_t t) { return t ? &t->h.t : nullptr; }
int ecwam_b200_host_tables_free(ecwam_b200_host_tables_t t) { delete t; return 0; }

// ---- src/ecwam/aki.F90:71-91 and depthprpt.F90:60-81
static double aki(const ecwam_b200_tables& t, double om, double beta) {
  const double EBS = 0.0001, DKMAX = 40.0;
  double ao = std::max(om * om / (4.0 * t.g), om / (2.0 * std::sqrt(t.g * beta)));
  for (;;) {
    const double akp = ao, bo = beta * ao;
    if (bo > DKMAX) return om * om / t.g;
    const double th = t.g * ao * std::tanh(bo), sth = std::sqrt(th), ch = std::cosh(bo);
    ao = ao + (om - sth) * sth * 2.0 / (th / ao + t.g * bo / (ch * ch));
    if (!(std::fabs(akp - ao) > EBS * ao)) return ao;
  }
}
int ecwam_b200_host_depthprpt(const ecwam_b200_tables* t, int nfre, long long n, const double* depth, double* wavnum,
                              double* cinv, double* cgroup, double* xk2cg, double* omosnh2kd, double* stokfac) {
  if (!t || !depth || n < 0) return ECWAM_B200_EINVAL;
  const double pi = 4.0 * std::atan(1.0), gh = t->g / (4.0 * pi);
  for (int m = 0; m < nfre; ++m) {
    const double om = t->zpifr[m];
#pragma omp parallel for schedule(static)
    for (long long ij = 0; ij < n; ++ij) {
      const size_t o = (size_t)ij + (size_t)n * m;
      const double ak = aki(*t, om, depth[ij]), akd = ak * depth[ij];
      double cg, osh, stf;
      if (akd <= 10.0) {
        cg = 0.5 * std::sqrt(t->g * std::tanh(akd) / ak) * (1.0 + 2.0 * akd / std::sinh(2.0 * akd));
        osh = om / std::sinh(2.0 * akd);
        stf = 2.0 * t->g * ak * ak / (om * std::tanh(2.0 * akd));
      } else {
        cg = gh / t->fr[m];
        osh = 0.0;
        stf = 2.0 / t->g * om * om * om;
      }
      if (wavnum) wavnum[o] = ak;
      if (cgroup) cgroup[o] = cg;
      if (omosnh2kd) omosnh2kd[o] = osh;
      if (stokfac) stokfac[o] = stf;
      if (cinv) cinv[o] = ak / om;
      if (xk2cg) xk2cg[o] = ak * ak * cg;
    }
  }
  return 0;
}

}  // extern "C"
