// Internal declarations shared by the translation units of libecwam_b200.so.
#pragma once
#include "../../include/ecwam_b200.h"
#include "dev_common.cuh"

namespace ew {

// Device view of one rank's propagation tables (built once at ecwam_b200_create).
// Extended 0-based index e = IJ - NINF: halo-below [0,nbot), own [nbot,nbot+nloc), halo-above, land = next-1.
struct PropDev {
  int nloc, nbot, ntop, next;   // next = nbot+nloc+ntop+1
  int P, A, F, Fr, nchnk;
  const int* nbr;        // [14][nloc]: KLON(1..2), KLAT(ic,icl) at 2+(ic-1)+2(icl-1), KCOR(icr,icl) at 6+(icr-1)+4(icl-1)
  const double* wl;      // [6][nloc]: WLAT(1..2), WCOR(1..4) after CTUWINI's land edit (ctuwini.F90:61-99)
  const double* pt;      // [5][nloc]: COSPHM1, DP(1), DP(2), ZDELLO(ky), TANPH(ky)
  const double* cgext;   // [nenv][next] group velocity incl. halo and land slot (proenvhalo.F90); row Fr = DEPTH_EXT when IREFRA = 1
  int irefra, nenv;      // YOWSTAT IREFRA (0 | 1 depth refraction); rows of cgext (Fr or Fr+1)
  const double* omos;    // WVPRPT%OMOSNH2KD (P,F,C) (IREFRA = 1)
  const double* grad;    // [7][nloc]: DDPHI, DDLAM (gradi.F90:120-153, IREFRA = 1, 3), DUPHI, DULAM, DVPHI, DVLAM (clamped, :167-229),
                         // OMDD (propdot.F90:134-143)  (IREFRA = 2, 3)
  const double* wavn;    // WVPRPT%WAVNUM (P,F,C) (IREFRA = 2, 3: SDOT)
  const double* curmask; // [nloc] CURMASK of CTUW (ctuw.F90:113-127), IREFRA = 2, 3
  const int* halo_off;   // [nbot+ntop+1] offset of (k=0,m=0) of a halo point in `halo`; land -> a zero element
  const int* halo_str;   // [nbot+ntop+1] direction stride (= points received from that peer); land -> 0
  const double* halo;    // received spectra, per peer block [m][k][ih]
  const double* obs;     // LSUBGRID: [8][Fr][nloc] OBSLON(.,.,1:2), OBSLAT(.,.,1:2), OBSCOR(.,.,1:4) (ctuw.F90:700-733); null = all 1
  const double* pad_obs; // keeps sizeof(PropDev) a multiple of 16 (the kernel arguments after it stay 16-byte aligned)
};

// per-direction tables of the CTU scheme (ctuwupdt.F90:111-161), constant memory
struct PropConst {
  int quad[EW_MAXA];        // 0: cos>=0,sin>=0  1: cos>=0,sin<0  2: cos<0,sin>=0  3: cos<0,sin<0
  int kq[5];                // directions [kq[j],kq[j+1]) form compass quadrant j: quad = 0, 2, 3, 1 (TH increasing)
  int kpm_m[EW_MAXA];       // KPM(K,-1) 0-based
  int kpm_p[EW_MAXA];       // KPM(K,+1)
  double sinth[EW_MAXA], costh[EW_MAXA];
  double sp[2][EW_MAXA];    // DELTH0*(SINTH(K)+SINTH(KP1))/R for the two DELPRO values (ctuw.F90:423-431)
  double sm[2][EW_MAXA];
  double delpro[2];
  double delth0[2];         // 0.25*DELPRO/DELTH (ctuw.F90:407)
  double delfr0[2];         // 0.25*DELPRO/((FRATIO-1)*ZPI) (ctuw.F90:508)
  double fratio;
  double fr[EW_MAXF];       // FR(M)
  double cmtodeg;           // 360/CIRC
  double xdella;
};
int upload_prop_const(const PropConst& h, cudaStream_t st);
int upload_prop_const_fast(const PropConst& h, cudaStream_t st);   // the copy read by propag_fast.cu
// 1: ECWAM_B200_PROPAG=exact, the bit-exact PROPAGS2 kernel of propag.cu is used for IREFRA = 0, 1 as well (verifier)
bool propag_exact_mode();
void launch_propags2_fast(const PropDev& d, const double* src, int srcF, double* dst, int dstF, int m0, int m1, int msplit,
                          cudaStream_t st, int l0, int l1);

// own points [l0, l1) (l1 < 0: all)
// top / topF: IREFRA = 2, 3 in a fast-wave sub-step: the array (layout (P,A,topF,C)) whose row m1 is the frequency neighbour of row m1-1
void launch_propags2(const PropDev& d, const double* src, int srcF, double* dst, int dstF, int m0, int m1, int msplit,
                     cudaStream_t st, int l0 = 0, int l1 = -1, const double* top = nullptr, int topF = 0);
void launch_ctu_check(const PropDev& d, int m0, int m1, int msplit, int* flag, int* count, cudaStream_t st);
void launch_setup_points(const PropDev& d, const double* cosphm1_fld, const double* cosph_m, const double* cosph_p,
                         double* pt, cudaStream_t st);
void launch_fill_cgext(const PropDev& d, const double* cgroup, const double* depth, const double* ucur, const double* vcur, double* cgext,
                       const double* land_cg, cudaStream_t st);
void launch_depth_gradients(const PropDev& d, const double* wlat_raw, const double* dellam, double oneo2delphi, double* grad, cudaStream_t st);
// CTUW's CFL / weight-range scan with currents (ICALL = 1: curmask = 1; ICALL = 2: curmask = 0 where the first scan failed)
void launch_ctu_check_cur(const PropDev& d, int m0, int m1, int msplit, int* flag, int* count, cudaStream_t st);
void launch_curmask(const PropDev& d, const int* flag, double* curmask, int reset, cudaStream_t st);
void launch_pack(const PropDev& d, const double* src, int srcF, const double* cgext, int mode, int nk, int nm, int nfull,
                 const int* send_l, const int* send_pre, const int* send_peer_of, int ntot, double* out, cudaStream_t st);
void launch_unpack_cg(const PropDev& d, const double* in, const int* recv_pre, const int* recv_peer_of, const int* recv_e,
                      int ntot, int nfull, double* cgext, cudaStream_t st);
void launch_copyback(const PropDev& d, const double* fl3, double* fl1, int m0, int m1, cudaStream_t st);
void launch_pad(const PropDev& d, double* fl, int flF, int m0, int m1, cudaStream_t st);

// ---- IMPLSCH -----------------------------------------------------------------------------------------------
struct NemoDev {
  ecwam_b200_nemo_fields f;   // the WAVE2OCEAN fields (valid when nemo_on)
  int nemo_on, strn_on;       // LWNEMOCOU (fields bound); LWNEMOCOUSTRN (CIMSSTRN fills STRNMS)
  int wrs_on, ibr_on;         // LWNEMOCOUWRS (k_ice forms the radiative stress, k_nemo stores it); LWNEMOCOUIBR (k_ice, needs f.ibrmem)
};
struct ImplDev {
  int P, A, F, Fr, nchnk;
  long long npts;          // P*nchnk (padded lanes included, as the reference: KIJL = NPROMA_WAM, wamintgr.F90:120)
  ecwam_b200_fields f;     // device pointers
  const double* fl_lo;     // source of FL1 for m < lo_nf: either f.fl1 (lo_F = F) or the propagation scratch (lo_F = Fr)
  int lo_F;
  int lo_on;               // 1: frequencies < Fr are read from fl_lo (layout (P,A,lo_F,C)), padded lanes from lane 0
  double* scr;             // [NSCR][npts] scalar scratch between the kernels
  double* fldin;           // (P,A,F,C) wind-input linearisation FLD handed from k_point to the stencil kernel
  int lwflux;              // YOWCOUP LWFLUX (selects the k_stencil instance that also forms WSEMEAN/WSFMEAN)
  long long nloc;          // own points (slots >= nloc of the last chunk are padding)
  DevTabPtr tab;
  double* tbg;             // [TQ_N][F][npts] per-(point, frequency) scalars handed from k_point to k_stencil
  int dsb[2][4];           // signed cyclic shifts of K1W, K11W, K2W, K21W (.,KH) (K1W(K,KH) = K + shift mod NANG) in bytes of
                           // k_stencil's shared-memory planes (shift * 8 points * 8 B)
  int iphys, nsdsnth;      // copies of the host-side switches the launcher needs
  int halo_r, halo_c;      // direction halo of the shared-memory spectrum rows / interaction planes of k_stencil
  int cy49;                // launcher flags.  bit 0: LLGCBZ0 or LLNORMAGAM is on (k_point runs its gravity-capillary / renormalised-growth
                           // instance); bit 1: ICODE_WND = 1, 2 (phase 1 runs the Z0WAVE instance).  Not read by the kernels.
  const double* gc;        // [GC_NT][NWAV_GC] gravity-capillary tables (device), read by that instance only
  int sweep_ok;            // the DIA tables have the separable structure k_sweep relies on (DevConst::NLW)
  int ssource_pre;         // LCFLX and not LWVFLX_SNL: WNFLUXES takes SL before SNONLIN (only k_stencil / k_stencil_dp carry that branch)
  int isnonlin;            // YOWSTAT ISNONLIN (0: ENH from the mean wavenumber; 1, 2: per centre frequency, snonlin.F90:138-163)
  double* enh;             // [MLSTHG][npts] ENH(IJ,MC) of ISNONLIN = 1, 2 (k_enh writes it between k_point and the sweep); null otherwise
  const double* ice1;      // LCIWA1: [NICT*NICH] CIDEAC, [F] WT1, [F] IT, [F] IT1 (0-based, as doubles) of SDICE1 (k_ice); null otherwise
  int ice_nt, ice_nh;      // NICT, NICH
  double ice_hmin, ice_dh; // HICMIN, DHIC
  const struct NemoDev* nemo;   // LWNEMOCOU / LWNEMOCOUSTRN: device-resident argument block of k_nemo (after the sweep); null otherwise.
                           // (A pointer, not the 13 fields: sizeof(ImplDev) stays a multiple of 16 -- the kernels' (p0, np) arguments keep
                           // their alignment -- and k_point's register allocation, which moved with the larger struct, stays as measured.)
  double* ice2;            // LCIWA2: [F][npts] per-(point, frequency) factor of SDICE2 (k_ice writes it, k_stencil / k_stencil_dp read it); null otherwise
};
#define EW_TQ_N 6          // number of planes of ImplDev::tbg
int upload_dev_const(const DevConst& h, cudaStream_t st);
// launches stage 0..1 of the IMPLSCH kernel sequence (k_point, k_stencil) for points [p0, p0+np)
#define EW_IMPLSCH_NSTAGE 2
int launch_implsch_stage(const ImplDev& d, long long p0, long long np, int stage, cudaStream_t st);
size_t implsch_scratch_doubles(long long npts);

// ---- NEWWIND / OUTBLOCK core / WAMNORM (outparam.cu) ------------------------------------------------------------
#define EW_OUT_MAXCOL 64
struct OutConst {   // constant memory of outparam.cu
  int A, F, NFRE_ODD, licerun, lmaskice, llsource, ncol;
  int itg[EW_OUT_MAXCOL], icemask[EW_OUT_MAXCOL], seamask[EW_OUT_MAXCOL];
  double EPSMIN, EPSUS, DELTH, WETAIL, FRTAIL, WP1TAIL, WP2TAIL, ZPI, G, GM1, DEG, XKAPPA, XNLEV, ALPHAMIN, ALPHAMAX, ROWATER,
      rnum, flmin, cithrsh, zmiss;
  double FR[EW_MAXF], DFIM[EW_MAXF], DFIMOFR[EW_MAXF], DFIMFR[EW_MAXF], DFIM_SIM[EW_MAXF];
  double TH[EW_MAXA], COSTH[EW_MAXA], SINTH[EW_MAXA];
  // SEBTMEAN (sebtmean.F90) is linear in the 1-D spectrum: EBT = EPSMIN + sum_m SEBT[band][m] * sum_k F(k,m); band 0 = SE10MEAN
  // (T > 10 s), bands 1..6 = the period intervals of mpcrtbl.F90:373-399
  double SEBT[7][EW_MAXF];
  int llgcbz0;             // OUTBETA: no wind-speed cap of the Charnock parameter (outbeta.F90:113-117)
  // MEANSQS (parameter 9): HALPHAP + MEANSQS_GC + MEANSQS_LF (meansqs.F90:80-100)
  int want_mss, NWAV_GC, NE_MSS, NFRE_EFF;
  double ALPHAPMAX, ZPI4GM2_FR5N, SQRTGOSURFT, XKM1_GC, XLOGKRATIOM1_GC, XKMSS, FCUT_MSS;
};
struct OutDev {
  int P, A, F, nchnk;
  long long npts;          // P*nchnk: OUTBLOCK runs over all lanes of every chunk (outbs.F90:100-101)
  ecwam_b200_fields f;
  const int* iodp;         // (P,C) or null (= 1 everywhere)
  double* bout;            // (P, NIPRMOUT, C)
  const double* gc;        // [GC_NT][NWAV_GC] gravity-capillary tables (MEANSQS_GC), null when not supplied
  const double* fl2;       // IREFRA = 2, 3: INTPOL's spectrum on the absolute frequency axis (P,A,F,C), written by k_intpol; null otherwise
};
int upload_out_const(const OutConst& h, cudaStream_t st);
void launch_newwind(long long npts, const ecwam_b200_fields& f, const ecwam_b200_forcing_next& nx, double acd, double bcd, double epsmin,
                    cudaStream_t st);
// ICODE_WND = 1, 2 (newwind.F90:141-150): FF_NEXT%UFRIC instead of FF_NEXT%WSWAVE, first-guess TAUW from the Charnock parameter
void launch_newwind_ustar(long long npts, const ecwam_b200_fields& f, const ecwam_b200_forcing_next& nx, const double* ufric_next, double alpha,
                          cudaStream_t st);
struct GetwndArgs {
  long long npts;
  int nx, lcorrel, licerun, lmaskice;
  double wspmin, zpi;
  const double *ucur, *vcur;
};
void launch_getwnd(const GetwndArgs& a, const ecwam_b200_fieldg& g, const ecwam_b200_getwnd_opts& o, const int* ifromij, const int* jfromij,
                   const ecwam_b200_forcing_next& nx, cudaStream_t st);
void launch_no_source(long long n4, long long n2, double* fl1, double* xllws, int* mij, int nfre, int clip, double epsmin, cudaStream_t st);
int launch_outblock(const OutDev& d, cudaStream_t st);
void launch_intpol(const OutDev& d, double* fla, double fratio, double flogsprdm1, double fr5n, cudaStream_t st);
size_t norm_scratch_doubles(int ncol);
void launch_norm_local(const double* bout, int P, int ncol, long long nloc, double zmiss, double* scratch, double* out4, cudaStream_t st);
void launch_pack_cols(const double* bout, int P, int ncol, long long nloc, double* out, long long ostride, cudaStream_t st);
void launch_norm_seq(const double* zg, const int* ij2new, long long niblo, int ncol, double zmiss, double* out4, cudaStream_t st);

}  // namespace ew
