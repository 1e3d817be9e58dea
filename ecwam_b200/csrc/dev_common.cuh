// Shared device-side declarations of the B200-native WAMINTGR hot path.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define EW_MAXF 40     // >= NFRE
#define EW_MAXA 48     // >= NANG
#define EW_MAXMC 48    // >= MLSTHG
#define EW_MAXSAT 25   // >= 2*NSDSNTH+1
#define EW_MAXJT 24    // >= JTOT_TAUHF

// Small read-only tables (YOWFRED/YOWPHYS/YOWINDN/... module state) held in __constant__ memory.
// Filled from ecwam_b200_tables + ecwam_b200_params at ecwam_b200_create.
struct DevConst {
  // dimensions / switches (ecwam_b200_params)
  int A, F, Fr, iphys, idamping, llcapchnk, lbiwbk, licerun, lmaskice, lwamrsetci, lwflux, lcflx, lwvflx_snl,
      lwcouast;
  double delt, ximp, rnu, rnum, wspmin, cithrsh, cithrsh_tail, ciblock, flmin, bathymax;
  int lciwa3, lciscal;          // YOWICE LCIWA3 (SDICE3), LCISCAL
  double zalpfacx, fr45[EW_MAXF];   // YOWICE ZALPFACX; 2*CDICE*FR(M)**4.5 of SDICE3 (sdice3.F90:123-129)
  // YOWPCONS
  double G, GM1, ZPI, ZPI4GM1, ZPI4GM2, ROWATERM1, EPSMIN, EPSUS, EPSU10, ACD, BCD, CDMAX, TAUOCMIN, TAUOCMAX,
      PHIEPSMIN, PHIEPSMAX, WSEMEAN_MIN;
  // YOWFRED
  double FRATIO, WETAIL, FRTAIL, WP1TAIL, DELTH, FLOGSPRDM1;
  int NFRE_ODD;
  double FR[EW_MAXF], DFIM[EW_MAXF], DFIMOFR[EW_MAXF], DFIMFR[EW_MAXF], ZPIFR[EW_MAXF], FR5[EW_MAXF],
      COFRM4[EW_MAXF], FLMAX[EW_MAXF], RHOWG_DFIM[EW_MAXF], DFIM_SIM[EW_MAXF];
  double TH[EW_MAXA], COSTH[EW_MAXA], SINTH[EW_MAXA];
  // YOWPHYS
  double XKAPPA, XNLEV, ALPHA, ALPHAMIN, CHNKMIN_U, ZALP, BETAMAXOXKAPPA2, TAUWSHELTER, TAILFACTOR, TAILFACTOR_PM,
      SWELLF, SWELLF2, SWELLF3, SWELLF4, SWELLF5, SWELLF6, SWELLF7, SWELLF7M1, Z0RAT, Z0TUBMAX, ABMIN, ABMAX, CDIS,
      DELTA_SDIS, CDISVIS, SDSBR, SSDSC2, SSDSC3, SSDSC4, SSDSC5, SSDSC6, MICHE, EGRCRV, AFCRV, BFCRV;
  int NSDSNTH;
  // YOWTABL / YOWCOUP
  int IAB, JTOT;
  double EPS1, X0TAUHF, WTAUHF[EW_MAXJT];
  // YOWINDN: per-MC interaction tables (0-based MC index = MC-1), frequency indices are 1-based as in Fortran
  int MLSTHG, MFRSTLW, KFRH;
  double DAL1, DAL2;
  int IKP[EW_MAXMC], IKP1[EW_MAXMC], IKM[EW_MAXMC], IKM1[EW_MAXMC];
  int INLCOEF[EW_MAXMC][5];
  double RNLCOEF[EW_MAXMC][25];
  double AF11[EW_MAXMC];
  // k_stencil: slot (row % 9) of the 9-row shared-memory ring for every frequency row and for INLCOEF(1:5,MC)
  int SLOT9[EW_MAXF + 8];
  int NLSLOT[EW_MAXMC][5];
  double SATW1[EW_MAXSAT];  // SATWEIGHTS(NANG/2, 1:2*NSDSNTH+1): the saturation window weights (they do not depend on the direction)
  double RNLC2[EW_MAXMC];   // 2 (or 0 where the centre row is outside the spectrum): weight of -AD, -DELAD at the centre bin
  // LLGCBZ0 / LLNORMAGAM (gravity-capillary roughness, growth renormalisation; appended so that the other offsets stay put)
  int llgcbz0, llnormagam, NWAV_GC, pad_gc;
  double ALPHAMAX, ALPHAPMAX, ACDLIN, BCDLIN, BMAXOKAP, GAMNCONST, RN1_RN, DTHRN_A, DTHRN_U, ANG_GC_A, ANG_GC_B, ANG_GC_C,
      SQRTGOSURFT, XKM1_GC, XLOGKRATIOM1_GC;
  // k_sweep: the DIA weights in separable form (nlweigt.F90: every RNLCOEF entry is a per-MC frequency weight times one of the
  // two direction-interpolation weights).  NLW[mc] = {WP, WP1, WM, WM1, CMP, CMP1, CMM, CMM1, CMP^2, CMP1^2, CMM^2, CMM1^2}:
  // SAP = WP*D+(IP) + WP1*D+(IP1), SAM = WM*D-(IM) + WM1*D-(IM1) with D+ = CL11*F(K1W) + ACL1*F(K11W), D- = CL21*F(K2W) + ACL2*F(K21W);
  // rows MC+2, MC+3, MC-4, MC-3 receive CMP, CMP1, CMM, CMM1 (squared for FLD) times the direction-interpolated gather.
  // NLD = {CL11, ACL1, CL21, ACL2, CL11^2, ACL1^2, CL21^2, ACL2^2}.  NLS2[st+1] = ring slots of rows IP1(MC0=st), IM1(MC0=st)
  // (entry 0: IP, IM of the first centre frequency).  sweep_ok: the tables have that structure (checked at create).
  double NLW[EW_MAXMC][12];
  double NLD[8];
  int NLS2[EW_MAXMC + 1][2];
  int sweep_ok, pad_sw;
  // SDICE1 / SDICE2 (appended: the offsets above stay put)
  int lciwa1, lciwa2, lciwa_any, pad_ice;   // lciwa_any: LCIWA1 or LCIWA2 or LCIWA3 (WNFLUXES' sea-ice constants, wnfluxes.F90:150-158)
  double zalpfacb, cdicwa;                  // YOWICE ZALPFACB, CDICWA
  int lwnemotauoc, lwnemocoustk, nemo_send, pad_nemo;   // YOWCOUP; nemo_send = (LWNEMOCOUSEND and LWCOU) or not LWCOU (stokestrn.F90:76-78)
  double ROWATER, zalpwrs;                  // 1 / ROWATERM1 (AKI_ICE); YOWICE ZALPWRS
  double zibrw_thrsh;                       // YOWICE ZIBRW_THRSH
  int lwnemocouwrs, lwnemocouibr;           // YOWCOUP.  NOTE: sizeof(DevConst) must stay a multiple of 16 (static_assert in implsch.cu): the
                                            // constants that follow c_dc (c_exp) keep their alignment, and with it k_point's code as measured
};
// rows of the gravity-capillary table ImplDev::gc [GC_NT][NWAV_GC] (YOWFRED *_GC, initgc.F90)
enum { GC_XK = 0, GC_OMEGA, GC_CM, GC_C2OSQRTVG, GC_XKMSQRTVGOC2, GC_OM3GMKM, GC_OMXKM3, GC_DELKCC_NS, GC_DELKCC_OMXKM3, GC_DELKCC,
       GC_LXK /* log(XK_GC), derived at create: STRESS_GC's LOG(XK_GC*Z0) = GC_LXK + log(Z0) */, GC_NT };

// tables too irregular / large for constant memory (per-lane indexed)
struct DevTabPtr {
  const int* k1w;         // [2][A] 0-based direction indices
  const int* k2w;
  const int* k11w;
  const int* k21w;
  const int* ik1w;        // inverse maps: ik1w[kh][k] = K with K1W(K,kh) = k  (gather form of SNONLIN)
  const int* ik2w;
  const int* ik11w;
  const int* ik21w;
  const int* indicessat;  // [2N+1][A] 0-based
  const double* satweights;  // [2N+1][A]
  const double* swellft;  // [IAB]
};

#define EW_CUDA_CHECK(call)                                                              \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      ew_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return ECWAM_B200_ECUDA;                                                           \
    }                                                                                    \
  } while (0)

void ew_set_error(const char* fmt, ...);
