// PROPAGS2 (src/ecwam/propags2.F90:99-121) with in-kernel CTU weights, tolerance mode: the default advection kernel.
//
// propag.cu reproduces the reference's stored-weight result bit for bit by recomputing every weight in CTUW's own operation
// order without FMA contraction: ~60 FP64 operations per spectral bin, which puts that kernel between the HBM and the FP64
// floors.  Nothing in the scope asks for bit-identical spectra (only the index tables are bit-exact), so this translation
// unit factors the weights of ctuw.F90:160-275,404-501 into per-(point, frequency, quadrant) terms times per-direction
// constants, and lets the compiler contract:
//   DXUP = |sin th| X1, DXDW = |sin th| X2, DYUP = |cos th| Y1, DYDW = |cos th| Y2   (X, Y >= 0: interface group speed x DELPRO x 360/CIRC)
//   WLONN = (XDELLA - |cos| Y2) |sin| (X1 g)            WLATN = (ZDELLO - |sin| X2) |cos| (Y1 g) {WLAT, 1-WLAT}
//   WCORN = |sin||cos| (X1 Y1 g) {WCOR, 1-WCOR}         SUMWN = |cos| (ZDELLO Y2 g) + |sin| (XDELLA X2 g) - |sin||cos| (X2 Y2 g) + turning
//   turning (great circle [+ depth refraction]): DTHP = T SP(k) [+ OMOSNH2KD DRDP], DTHM likewise, T = tan(phi) CG
// 29 FP64 operations per bin, the two latitude / corner neighbours are interpolated before they are weighted.
// Differences from the exact kernel are rounding only (<= 1e-15 relative per step; tests/test_gpu_parity.py bounds them at 1e-13).
// ECWAM_B200_PROPAG=exact selects the bit-exact kernel (the verifier); IREFRA = 2, 3 always uses the exact current kernel.
#include "internal.h"
#include <cstdlib>
#include <cstring>

namespace ew {

__constant__ PropConst c_pf;
int upload_prop_const_fast(const PropConst& h, cudaStream_t st) {
  EW_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_pf, &h, sizeof(PropConst), 0, cudaMemcpyHostToDevice, st));
  return 0;
}

namespace {
struct Src {
  const double* base;   // spectrum in the NPROMA-chunked layout (P, A, nF, C)
  long long cstride;    // P*A*nF
};
__device__ __forceinline__ void nbr_base(const PropDev& d, const Src& s, int e, int m, const double*& p, int& kstr) {
  const int l = e - d.nbot;
  if ((unsigned)l < (unsigned)d.nloc) {
    const int c = l / d.P;
    const int i = l - c * d.P;
    p = s.base + i + (long long)c * s.cstride + (long long)m * d.P * d.A;
    kstr = d.P;
  } else {
    const int h = (e < d.nbot) ? e : e - d.nloc;
    const int st = __ldg(d.halo_str + h);
    p = d.halo + __ldg(d.halo_off + h) + (long long)m * d.A * st;
    kstr = st;
  }
}
// per-(point, frequency, quadrant) factors of the weights
struct QuadW {
  double a1;    // X1 g          -> WLONN
  double b1;    // Y1 g          -> WLATN
  double c1;    // X1 Y1 g       -> WCORN
  double x2, y2;
  double e1, e2, e3;   // ZDELLO Y2 g, XDELLA X2 g, X2 Y2 g  -> SUMWN
  double wlat, wcor;   // WLAT(JYO), WCOR(KCR) of the quadrant's upwind neighbours
  double t;            // tan(phi) CG
  double omos, ddphi, ddlam_c;   // IREFRA = 1: OMOSNH2KD, DDPHI, DDLAM*COSPHM1
  double zdello;
};
template <bool REFRA>
__device__ __forceinline__ double ctu_bin(const QuadW& q, int k, int idp, double f0, double flon, double flat1, double flat2, double fc1,
                                          double fc2, double fkm, double fkp) {
  const double S = fabs(c_pf.sinth[k]), C = fabs(c_pf.costh[k]), SC = S * C;
  const double wlonn = fma(-C, q.y2, c_pf.xdella) * (S * q.a1);
  const double wl = fma(-S, q.x2, q.zdello) * (C * q.b1);
  const double wc = SC * q.c1;
  const double sxy = fma(-SC, q.e3, fma(S, q.e2, C * q.e1));
  double dthp = q.t * c_pf.sp[idp][k], dthm = q.t * c_pf.sm[idp][k];
  if (REFRA) {   // ctuw.F90:434-439, 487-501 with THDD of propdot.F90:156
    const int kp = c_pf.kpm_p[k], km = c_pf.kpm_m[k];
    const double th = fma(c_pf.sinth[k], q.ddphi, -c_pf.costh[k] * q.ddlam_c);
    const double thp = fma(c_pf.sinth[kp], q.ddphi, -c_pf.costh[kp] * q.ddlam_c);
    const double thm = fma(c_pf.sinth[km], q.ddphi, -c_pf.costh[km] * q.ddlam_c);
    dthp = fma(q.omos, (th + thp) * c_pf.delth0[idp], dthp);
    dthm = fma(q.omos, (th + thm) * c_pf.delth0[idp], dthm);
  }
  const double w0 = (dthp + fabs(dthp)) + (fabs(dthm) - dthm);
  const double wp = fabs(dthp) - dthp;
  const double wm = dthm + fabs(dthm);
  const double c0 = 1.0 - (sxy + w0);
  const double fl = fma(q.wlat, flat1 - flat2, flat2);
  const double fcn = fma(q.wcor, fc1 - fc2, fc2);
  return fma(wp, fkp, fma(wm, fkm, fma(wc, fcn, fma(wl, fl, fma(wlonn, flon, c0 * f0)))));
}

struct PointF {   // k- and quadrant-independent quantities of one (point, frequency)
  double hx[2], hy[2], cxg, cyg, cx, cy, gam1, zdello, t, omos, ddphi, ddlam_c;
};

#ifndef PF_UNR
#define PF_UNR 2   // direction pairs in flight per loop iteration (2 pairs = 24 neighbour gathers + 6 own loads)
#endif
#define PF_STR2(x) #x
#define PF_STR(x) PF_STR2(x)
#define PF_UNROLL _Pragma(PF_STR(unroll PF_UNR))
// One compass quadrant of directions [k0, k1): the upwind selectors JXO / JYO / KCR of ctuwupdt.F90:111-161 are constant inside it.
// jx1, jy1 (1 | 2) and kc (1..4) are run-time values so that the four quadrants share ONE copy of the direction loop (four
// template instances of it, unrolled, were 80 KB of code against a 32 KB instruction cache: the kernel stalled on fetch).
template <bool REFRA>
__device__ __forceinline__ void quadrant(const PropDev& d, const Src& src, const PointF& pf, int l, int m, int idp, int jx1, int jy1, int kc,
                                         int k0, int k1, const double* __restrict__ ps, double* __restrict__ pd) {
  if (k0 >= k1) return;
  const int nl = d.nloc;
  const int e_lon = __ldg(d.nbr + (size_t)(jx1 - 1) * nl + l);
  const int e_la1 = __ldg(d.nbr + (size_t)(2 + (jy1 - 1)) * nl + l);
  const int e_la2 = __ldg(d.nbr + (size_t)(4 + (jy1 - 1)) * nl + l);
  const int e_c1 = __ldg(d.nbr + (size_t)(6 + (kc - 1)) * nl + l);
  const int e_c2 = __ldg(d.nbr + (size_t)(10 + (kc - 1)) * nl + l);
  const double *p_lon, *p_la1, *p_la2, *p_c1, *p_c2;
  int s_lon, s_la1, s_la2, s_c1, s_c2;
  nbr_base(d, src, e_lon, m, p_lon, s_lon);
  nbr_base(d, src, e_la1, m, p_la1, s_la1);
  nbr_base(d, src, e_la2, m, p_la2, s_la2);
  nbr_base(d, src, e_c1, m, p_c1, s_c1);
  nbr_base(d, src, e_c2, m, p_c2, s_c2);
  QuadW q;
  {
    const double hx1 = jx1 == 1 ? pf.hx[0] : pf.hx[1], hx2 = jx1 == 1 ? pf.hx[1] : pf.hx[0];
    const double hy1 = jy1 == 1 ? pf.hy[0] : pf.hy[1], hy2 = jy1 == 1 ? pf.hy[1] : pf.hy[0];
    const double x1 = hx1 * pf.cx, x2 = hx2 * pf.cx;     // DELPRO * interface speed * COSPHM1 * 360/CIRC
    const double y1 = hy1 * pf.cy, y2 = hy2 * pf.cy;
    q.a1 = x1 * pf.gam1; q.b1 = y1 * pf.gam1; q.c1 = x1 * q.b1;
    q.x2 = x2; q.y2 = y2;
    q.e1 = pf.zdello * y2 * pf.gam1; q.e2 = c_pf.xdella * x2 * pf.gam1; q.e3 = x2 * y2 * pf.gam1;
    q.wlat = __ldg(d.wl + (size_t)(jy1 - 1) * nl + l);
    q.wcor = __ldg(d.wl + (size_t)(2 + kc - 1) * nl + l);
    q.t = pf.t; q.omos = pf.omos; q.ddphi = pf.ddphi; q.ddlam_c = pf.ddlam_c; q.zdello = pf.zdello;
  }
  const int P = d.P;
  int k = k0;
  PF_UNROLL
  for (; k + 1 < k1; k += 2) {
    const int ka = k, kb = k + 1;
    const double a0 = ps[(size_t)ka * P], b0 = ps[(size_t)kb * P];
    const double am = ps[(size_t)c_pf.kpm_m[ka] * P], bp = ps[(size_t)c_pf.kpm_p[kb] * P];
    const double a1 = __ldg(p_lon + (size_t)ka * s_lon), b1 = __ldg(p_lon + (size_t)kb * s_lon);
    const double a2 = __ldg(p_la1 + (size_t)ka * s_la1), b2 = __ldg(p_la1 + (size_t)kb * s_la1);
    const double a3 = __ldg(p_la2 + (size_t)ka * s_la2), b3 = __ldg(p_la2 + (size_t)kb * s_la2);
    const double a4 = __ldg(p_c1 + (size_t)ka * s_c1), b4 = __ldg(p_c1 + (size_t)kb * s_c1);
    const double a5 = __ldg(p_c2 + (size_t)ka * s_c2), b5 = __ldg(p_c2 + (size_t)kb * s_c2);
    // KPM(ka,+1) = kb and KPM(kb,-1) = ka inside a quadrant
    const double ra = ctu_bin<REFRA>(q, ka, idp, a0, a1, a2, a3, a4, a5, am, b0);
    const double rb = ctu_bin<REFRA>(q, kb, idp, b0, b1, b2, b3, b4, b5, a0, bp);
    pd[(size_t)ka * P] = ra;
    pd[(size_t)kb * P] = rb;
  }
  for (; k < k1; ++k) {
    const double f0 = ps[(size_t)k * P];
    const double fkm = ps[(size_t)c_pf.kpm_m[k] * P], fkp = ps[(size_t)c_pf.kpm_p[k] * P];
    pd[(size_t)k * P] = ctu_bin<REFRA>(q, k, idp, f0, __ldg(p_lon + (size_t)k * s_lon), __ldg(p_la1 + (size_t)k * s_la1),
                                       __ldg(p_la2 + (size_t)k * s_la2), __ldg(p_c1 + (size_t)k * s_c1), __ldg(p_c2 + (size_t)k * s_c2),
                                       fkm, fkp);
  }
}

#ifndef PF_MINB
#define PF_MINB 4
#endif
// thread = one own grid point x one group of MG frequencies; blockIdx.x (points) fastest so that the rows north and south of
// the running row stay L2-resident for one frequency group at a time (same grid as the exact kernel)
template <bool REFRA>
__global__ void __launch_bounds__(128, PF_MINB) propags2_fast_kernel(PropDev d, Src src, double* __restrict__ dst, long long dcstride, int m0,
                                                                      int m1, int MG, int msplit, int l0, int l1) {
  const int l = l0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= l1) return;
  const int mb = m0 + blockIdx.y * MG;
  const int me = min(mb + MG, m1);
  const int nl = d.nloc;
  int nb[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) nb[j] = __ldg(d.nbr + (size_t)j * nl + l);
  PointF pf;
  const double wlat0 = __ldg(d.wl + l), wlat1 = __ldg(d.wl + nl + l);
  const double cosphm1 = __ldg(d.pt + l);
  const double dp1 = __ldg(d.pt + nl + l), dp2 = __ldg(d.pt + 2 * (size_t)nl + l);
  pf.zdello = __ldg(d.pt + 3 * (size_t)nl + l);
  const double tanph = __ldg(d.pt + 4 * (size_t)nl + l);
  pf.gam1 = 1.0 / (pf.zdello * c_pf.xdella);
  const int e0 = d.nbot + l;
  const int c = l / d.P, i = l - c * d.P;
  const int A = d.A;
  pf.omos = 0.0; pf.ddphi = 0.0; pf.ddlam_c = 0.0;
  if (REFRA) { pf.ddphi = __ldg(d.grad + l); pf.ddlam_c = __ldg(d.grad + nl + l) * cosphm1; }
  for (int m = mb; m < me; ++m) {
    const int idp = (m < msplit) ? 0 : 1;
    const double* cgm = d.cgext + (size_t)m * d.next;
    const double cg = __ldg(cgm + e0);
    if (REFRA) pf.omos = __ldg(d.omos + i + (size_t)d.P * (m + (size_t)d.F * c));
    pf.hx[0] = 0.5 * (cg + __ldg(cgm + nb[0]));
    pf.hx[1] = 0.5 * (cg + __ldg(cgm + nb[1]));
    {
      const double cgyp1 = fma(wlat0, __ldg(cgm + nb[2]) - __ldg(cgm + nb[4]), __ldg(cgm + nb[4]));
      const double cgyp2 = fma(wlat1, __ldg(cgm + nb[3]) - __ldg(cgm + nb[5]), __ldg(cgm + nb[5]));
      pf.hy[0] = 0.5 * fma(dp1, cgyp1, cg);
      pf.hy[1] = 0.5 * fma(dp2, cgyp2, cg);
    }
    pf.cy = c_pf.delpro[idp] * c_pf.cmtodeg;
    pf.cx = pf.cy * cosphm1;
    pf.t = tanph * cg;
    const double* ps = src.base + i + (long long)c * src.cstride + (long long)m * d.P * A;
    double* pd = dst + i + (long long)c * dcstride + (long long)m * d.P * A;
    // quadrants in the order of increasing TH: (west, south, SW) (west, north, NW) (east, north, NE) (east, south, SE)
#pragma unroll 1
    for (int qd = 0; qd < 4; ++qd) {
      const int jx1 = qd < 2 ? 1 : 2, jy1 = (qd == 1 || qd == 2) ? 2 : 1, kc = qd == 0 ? 3 : (qd == 1 ? 4 : (qd == 2 ? 1 : 2));
      quadrant<REFRA>(d, src, pf, l, m, idp, jx1, jy1, kc, c_pf.kq[qd], c_pf.kq[qd + 1], ps, pd);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// propags2_q_kernel: the same arithmetic with the loads decoupled from the arithmetic.
// The kernels above are latency-bound (ncu: 10 long-scoreboard stall cycles per issued instruction at 16 warps per SM, DRAM at
// 17 % of peak): a thread issues the 14 gathers of a direction pair, waits for them, computes, and only then issues the next
// ones.  Here thread = (own grid point, compass quadrant, group of MG frequencies): the neighbour pointers of the quadrant are
// set up once for all its frequencies, the per-thread state is small enough for 4-5 CTAs per SM, and the 14 values of direction
// pair i+1 travel HBM/L2 -> shared memory with cp.async (thread-private slots, [stage][value][thread]) while pair i is computed.
// ---------------------------------------------------------------------------------------------------------------------------
#ifndef PQ_MINB
#define PQ_MINB 4
#endif
#define PQ_NTH 128
#ifndef PQ_ST
#define PQ_ST 3      // stages of the cp.async pipeline: PQ_ST - 1 direction pairs in flight ahead of the one being computed
#endif
#define PQ_NV 14     // a0 b0 am bp | lon a,b | lat1 a,b | lat2 a,b | cor1 a,b | cor2 a,b
__device__ __forceinline__ void nbr_base_m(const PropDev& d, const Src& s, int e, int m, const double*& p, int& kstr, int& mstr) {
  const int l = e - d.nbot;
  if ((unsigned)l < (unsigned)d.nloc) {
    const int c = l / d.P;
    const int i = l - c * d.P;
    p = s.base + i + (long long)c * s.cstride + (long long)m * d.P * d.A;
    kstr = d.P; mstr = d.P * d.A;
  } else {
    const int h = (e < d.nbot) ? e : e - d.nloc;
    const int st = __ldg(d.halo_str + h);
    p = d.halo + __ldg(d.halo_off + h) + (long long)m * d.A * st;
    kstr = st; mstr = d.A * st;
  }
}
__device__ __forceinline__ void cpa8(unsigned sa, const double* g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(g)); }

template <bool REFRA>
__global__ void __launch_bounds__(PQ_NTH, PQ_MINB) propags2_q_kernel(PropDev d, Src src, double* __restrict__ dst, long long dcstride, int m0,
                                                                     int m1, int MG, int msplit, int l0, int l1) {
  extern __shared__ double stage[];     // [PQ_ST][PQ_NV][PQ_NTH]
  const int l = l0 + blockIdx.x * PQ_NTH + threadIdx.x;
  if (l >= l1) return;                  // no CTA-wide synchronisation below: a thread without a point just leaves
  const int qd = blockIdx.y & 3;
  const int mb = m0 + (blockIdx.y >> 2) * MG;
  const int me = min(mb + MG, m1);
  const int k0 = c_pf.kq[qd], k1 = c_pf.kq[qd + 1];
  if (k0 >= k1 || mb >= me) return;
  // quadrants in the order of increasing TH: (west, south, SW) (west, north, NW) (east, north, NE) (east, south, SE)
  const int jx1 = qd < 2 ? 1 : 2, jy1 = (qd == 1 || qd == 2) ? 2 : 1, kc = qd == 0 ? 3 : (qd == 1 ? 4 : (qd == 2 ? 1 : 2));
  const int nl = d.nloc, P = d.P, A = d.A;
  const int npair = (k1 - k0 + 1) >> 1;
  const int c = l / P, i = l - c * P;
  // ---- neighbour pointers of the quadrant at frequency mb
  const double *p_lon, *p_la1, *p_la2, *p_c1, *p_c2;
  int s_lon, s_la1, s_la2, s_c1, s_c2, t_lon, t_la1, t_la2, t_c1, t_c2;
  nbr_base_m(d, src, __ldg(d.nbr + (size_t)(jx1 - 1) * nl + l), mb, p_lon, s_lon, t_lon);
  nbr_base_m(d, src, __ldg(d.nbr + (size_t)(2 + (jy1 - 1)) * nl + l), mb, p_la1, s_la1, t_la1);
  nbr_base_m(d, src, __ldg(d.nbr + (size_t)(4 + (jy1 - 1)) * nl + l), mb, p_la2, s_la2, t_la2);
  nbr_base_m(d, src, __ldg(d.nbr + (size_t)(6 + (kc - 1)) * nl + l), mb, p_c1, s_c1, t_c1);
  nbr_base_m(d, src, __ldg(d.nbr + (size_t)(10 + (kc - 1)) * nl + l), mb, p_c2, s_c2, t_c2);
  const double* ps = src.base + i + (long long)c * src.cstride + (long long)mb * P * A;
  double* pd = dst + i + (long long)c * dcstride + (long long)mb * P * A;
  const int PA = P * A;
  // ---- per-point constants
  int nb[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) nb[j] = __ldg(d.nbr + (size_t)j * nl + l);
  const double wlat0 = __ldg(d.wl + l), wlat1 = __ldg(d.wl + nl + l);
  const double cosphm1 = __ldg(d.pt + l);
  const double dp1 = __ldg(d.pt + nl + l), dp2 = __ldg(d.pt + 2 * (size_t)nl + l);
  const double zdello = __ldg(d.pt + 3 * (size_t)nl + l);
  const double tanph = __ldg(d.pt + 4 * (size_t)nl + l);
  const double gam1 = 1.0 / (zdello * c_pf.xdella);
  const int e0 = d.nbot + l;
  QuadW q;
  q.zdello = zdello;
  q.wlat = __ldg(d.wl + (size_t)(jy1 - 1) * nl + l);
  q.wcor = __ldg(d.wl + (size_t)(2 + kc - 1) * nl + l);
  q.omos = 0.0; q.ddphi = 0.0; q.ddlam_c = 0.0;
  if (REFRA) { q.ddphi = __ldg(d.grad + l); q.ddlam_c = __ldg(d.grad + nl + l) * cosphm1; }
  // group velocity of the 7-point neighbourhood at frequency m (ctuw.F90:160-171, 199-210), loaded one frequency ahead
  double g0, g1, g2, g3, g4, g5, g6;
  auto load_cg = [&](int m) {
    const double* cgm = d.cgext + (size_t)m * d.next;
    g0 = __ldg(cgm + e0); g1 = __ldg(cgm + nb[0]); g2 = __ldg(cgm + nb[1]); g3 = __ldg(cgm + nb[2]); g4 = __ldg(cgm + nb[4]);
    g5 = __ldg(cgm + nb[3]); g6 = __ldg(cgm + nb[5]);
  };
  auto make_quad = [&](int m, int idp) {
    const double hx0 = 0.5 * (g0 + g1), hx1 = 0.5 * (g0 + g2);
    const double hy0 = 0.5 * fma(dp1, fma(wlat0, g3 - g4, g4), g0), hy1 = 0.5 * fma(dp2, fma(wlat1, g5 - g6, g6), g0);
    const double cy = c_pf.delpro[idp] * c_pf.cmtodeg, cx = cy * cosphm1;
    const double x1 = (jx1 == 1 ? hx0 : hx1) * cx, x2 = (jx1 == 1 ? hx1 : hx0) * cx;
    const double y1 = (jy1 == 1 ? hy0 : hy1) * cy, y2 = (jy1 == 1 ? hy1 : hy0) * cy;
    q.a1 = x1 * gam1; q.b1 = y1 * gam1; q.c1 = x1 * q.b1;
    q.x2 = x2; q.y2 = y2;
    q.e1 = zdello * y2 * gam1; q.e2 = c_pf.xdella * x2 * gam1; q.e3 = x2 * y2 * gam1;
    q.t = tanph * g0;
    if (REFRA) q.omos = __ldg(d.omos + i + (size_t)P * (m + (size_t)d.F * c));
  };
  // ---- the pipeline: iteration = (frequency, direction pair), flattened.  The seven source pointers (own + five neighbours) and
  // the destination pointer RUN with the iteration (one IMAD.WIDE each: + two directions, or + the rest of the row at the last
  // pair of a frequency) instead of being rebuilt from (frequency, direction) every time: ~130 of the 351 instructions of an
  // iteration were 64-bit address arithmetic, in a kernel that issues at 59 % with the FP64 pipe at 26 %.
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(stage) + threadIdx.x * 8u;
  const int nm = me - mb;
  const int NI = nm * npair;
  const int last = npair - 1;
  const double* io = ps + (size_t)k0 * P;
  const double* i1 = p_lon + (size_t)k0 * s_lon;
  const double* i2 = p_la1 + (size_t)k0 * s_la1;
  const double* i3 = p_la2 + (size_t)k0 * s_la2;
  const double* i4 = p_c1 + (size_t)k0 * s_c1;
  const double* i5 = p_c2 + (size_t)k0 * s_c2;
  const int eo = PA - last * 2 * P, e1 = t_lon - last * 2 * s_lon, e2 = t_la1 - last * 2 * s_la1, e3 = t_la2 - last * 2 * s_la2,
            e4 = t_c1 - last * 2 * s_c1, e5 = t_c2 - last * 2 * s_c2;
  int ip_i = 0, st_i = 0;            // issue side: pair and stage of the next iteration to issue
  auto issue_next = [&](int itn) {
    if (itn < NI) {
      const int ka = k0 + 2 * ip_i;
      const bool pr = ka + 1 < k1;
      const int kb = pr ? ka + 1 : ka;
      const unsigned sa = sbase + (unsigned)st_i * (PQ_NV * PQ_NTH * 8);
      cpa8(sa + 0 * PQ_NTH * 8, io);
      cpa8(sa + 1 * PQ_NTH * 8, io + (pr ? P : 0));
      cpa8(sa + 2 * PQ_NTH * 8, io + (c_pf.kpm_m[ka] - ka) * P);
      cpa8(sa + 3 * PQ_NTH * 8, io + (c_pf.kpm_p[kb] - ka) * P);
      cpa8(sa + 4 * PQ_NTH * 8, i1); cpa8(sa + 5 * PQ_NTH * 8, i1 + (pr ? s_lon : 0));
      cpa8(sa + 6 * PQ_NTH * 8, i2); cpa8(sa + 7 * PQ_NTH * 8, i2 + (pr ? s_la1 : 0));
      cpa8(sa + 8 * PQ_NTH * 8, i3); cpa8(sa + 9 * PQ_NTH * 8, i3 + (pr ? s_la2 : 0));
      cpa8(sa + 10 * PQ_NTH * 8, i4); cpa8(sa + 11 * PQ_NTH * 8, i4 + (pr ? s_c1 : 0));
      cpa8(sa + 12 * PQ_NTH * 8, i5); cpa8(sa + 13 * PQ_NTH * 8, i5 + (pr ? s_c2 : 0));
      const bool rowend = ip_i == last;
      io += rowend ? eo : 2 * P;
      i1 += rowend ? e1 : 2 * s_lon; i2 += rowend ? e2 : 2 * s_la1; i3 += rowend ? e3 : 2 * s_la2;
      i4 += rowend ? e4 : 2 * s_c1; i5 += rowend ? e5 : 2 * s_c2;
      ip_i = rowend ? 0 : ip_i + 1;
    }
    asm volatile("cp.async.commit_group;\n" ::);     // one group per iteration so that the wait_group counts stay fixed
    if (++st_i == PQ_ST) st_i = 0;
  };
  load_cg(mb);
#pragma unroll
  for (int pre = 0; pre < PQ_ST - 1; ++pre) issue_next(pre);
  int im = 0, ip = 0, st_c = 0;      // compute side
  double* o = pd + (size_t)k0 * P;   // destination of direction ka of the iteration being computed
  for (int it = 0; it < NI; ++it) {
    issue_next(it + PQ_ST - 1);
    asm volatile("cp.async.wait_group %0;\n" ::"n"(PQ_ST - 1) : "memory");
    const int m = mb + im;
    const int idp = (m < msplit) ? 0 : 1;
    if (ip == 0) {
      make_quad(m, idp);
      if (im + 1 < nm) load_cg(m + 1);     // used when the next frequency starts: npair iterations from now
    }
    const int ka = k0 + 2 * ip;
    const bool pair = ka + 1 < k1;
    const int kb = pair ? ka + 1 : ka;
    const double* sv = stage + (size_t)st_c * (PQ_NV * PQ_NTH) + threadIdx.x;
    const double a0 = sv[0 * PQ_NTH], b0 = sv[1 * PQ_NTH], am = sv[2 * PQ_NTH], bp = sv[3 * PQ_NTH];
    const double a1 = sv[4 * PQ_NTH], b1 = sv[5 * PQ_NTH], a2 = sv[6 * PQ_NTH], b2 = sv[7 * PQ_NTH];
    const double a3 = sv[8 * PQ_NTH], b3 = sv[9 * PQ_NTH], a4 = sv[10 * PQ_NTH], b4 = sv[11 * PQ_NTH];
    const double a5 = sv[12 * PQ_NTH], b5 = sv[13 * PQ_NTH];
    // KPM(ka,+1) = kb and KPM(kb,-1) = ka inside a quadrant; a single last direction takes its own KPM(+1) value (slot bp)
    const double ra = ctu_bin<REFRA>(q, ka, idp, a0, a1, a2, a3, a4, a5, am, pair ? b0 : bp);
    o[0] = ra;
    if (pair) {
      const double rb = ctu_bin<REFRA>(q, kb, idp, b0, b1, b2, b3, b4, b5, a0, bp);
      o[P] = rb;
    }
    const bool rowend = ip == last;
    o += rowend ? eo : 2 * P;
    ip = rowend ? 0 : ip + 1;
    im += rowend ? 1 : 0;
    if (++st_c == PQ_ST) st_c = 0;
  }
}
}  // namespace

void launch_propags2_fast(const PropDev& d, const double* src, int srcF, double* dst, int dstF, int m0, int m1, int msplit, cudaStream_t st,
                          int l0, int l1) {
  const int MG = 8;
  Src s{src, (long long)d.P * d.A * srcF};
  static const int mode = []() { const char* e = getenv("ECWAM_B200_PROPAG"); return (e && !strcmp(e, "fast1")) ? 1 : 0; }();
  if (mode == 1) {   // the one-thread-per-(point, frequency group) kernel (A/B)
    dim3 grid((l1 - l0 + 127) / 128, (m1 - m0 + MG - 1) / MG);
    if (d.irefra == 1) propags2_fast_kernel<true><<<grid, 128, 0, st>>>(d, s, dst, (long long)d.P * d.A * dstF, m0, m1, MG, msplit, l0, l1);
    else propags2_fast_kernel<false><<<grid, 128, 0, st>>>(d, s, dst, (long long)d.P * d.A * dstF, m0, m1, MG, msplit, l0, l1);
    return;
  }
  const size_t smem = (size_t)PQ_ST * PQ_NV * PQ_NTH * sizeof(double);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(propags2_q_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(propags2_q_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(propags2_q_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(propags2_q_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr_done = true;
  }
  dim3 grid((l1 - l0 + PQ_NTH - 1) / PQ_NTH, 4 * ((m1 - m0 + MG - 1) / MG));
  if (d.irefra == 1) propags2_q_kernel<true><<<grid, PQ_NTH, smem, st>>>(d, s, dst, (long long)d.P * d.A * dstF, m0, m1, MG, msplit, l0, l1);
  else propags2_q_kernel<false><<<grid, PQ_NTH, smem, st>>>(d, s, dst, (long long)d.P * d.A * dstF, m0, m1, MG, msplit, l0, l1);
}

}  // namespace ew
