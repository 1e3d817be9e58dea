// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (double precision, -ffp-contract=off) of ecWAM's WAMINTGR hot path
// (PROPAG_WAM/PROPAGS2/CTUW + IMPLSCH tree) and of the one-off tables that feed it.
// Every routine cites the reference file:line it follows (paths relative to
// /root/reference/src/ecwam unless stated).
//
// PINNING: the reference (Fortran + fiat + field_api + eccodes) cannot be built in this image
// and ships no per-routine golden vectors (only end-to-end swh norms that need GRIB forcing,
// SURVEY.md §8c).  This oracle is pinned against the reference's own source text, executed
// statement by statement through tests/golden/f90run.py (fixtures tests/golden/ref_*.npz):
// the IMPLSCH tree (<= 2.3e-15), the table builders, PROPCONNECT, CTUWUPDT / PROPDOT / GRADI /
// PROPAGS2 for IREFRA 0-3 and LSUBGRID (bit-identical), OUTBLOCK's 51 parameters (<= 1e-12),
// WAMWND + MICEP and NEWWIND (identical), MPDECOMP's sector decomposition and halo lists for
// 1-8 ranks (identical; its two MPL_ALLGATHERVs emulated), MPMINMAXAVG in both flavours for
// 1-5 ranks (identical; MPGATHERSCFLD / MPL_ALLREDUCE emulated, SUM in rank order) and
// TABU_SWELLFT + KERKEI + KZEONE (identical).  Not executed from source: the file formats
// (checked against scipy.io.FortranFile).  WAMODEL's loop + WAMINTGR + NEWWIND's date sequencing
// is executed from source too, against the product's host mirror (not part of this oracle).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.  The product (ecwam_b200/) never does.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <cmath>
#include <algorithm>
#include <stdexcept>
#include <string>
#include <limits>

namespace orc {

// ---------------------------------------------------------------------------
// Fortran-like array with arbitrary lower bounds, column-major.
template <class T>
struct Arr {
  std::vector<T> d;
  long lb[5] = {1, 1, 1, 1, 1};
  long n[5] = {0, 1, 1, 1, 1};
  Arr() {}
  void alloc(long l0, long u0, long l1 = 1, long u1 = 1, long l2 = 1, long u2 = 1, long l3 = 1, long u3 = 1,
             long l4 = 1, long u4 = 1) {
    lb[0] = l0; lb[1] = l1; lb[2] = l2; lb[3] = l3; lb[4] = l4;
    n[0] = std::max(0L, u0 - l0 + 1); n[1] = std::max(0L, u1 - l1 + 1); n[2] = std::max(0L, u2 - l2 + 1);
    n[3] = std::max(0L, u3 - l3 + 1); n[4] = std::max(0L, u4 - l4 + 1);
    d.assign((size_t)n[0] * n[1] * n[2] * n[3] * n[4], T());
  }
  inline T& operator()(long i) { return d[i - lb[0]]; }
  inline T& operator()(long i, long j) { return d[(i - lb[0]) + n[0] * (j - lb[1])]; }
  inline T& operator()(long i, long j, long k) { return d[(i - lb[0]) + n[0] * ((j - lb[1]) + n[1] * (k - lb[2]))]; }
  inline T& operator()(long i, long j, long k, long l) {
    return d[(i - lb[0]) + n[0] * ((j - lb[1]) + n[1] * ((k - lb[2]) + n[2] * (l - lb[3])))];
  }
  inline T& operator()(long i, long j, long k, long l, long m) {
    return d[(i - lb[0]) + n[0] * ((j - lb[1]) + n[1] * ((k - lb[2]) + n[2] * ((l - lb[3]) + n[3] * (m - lb[4]))))];
  }
  inline const T& operator()(long i) const { return d[i - lb[0]]; }
  inline const T& operator()(long i, long j) const { return d[(i - lb[0]) + n[0] * (j - lb[1])]; }
  inline const T& operator()(long i, long j, long k) const {
    return d[(i - lb[0]) + n[0] * ((j - lb[1]) + n[1] * (k - lb[2]))];
  }
  T* data() { return d.data(); }
  size_t size() const { return d.size(); }
};
typedef Arr<double> ArrD;
typedef Arr<int> ArrI;

inline long nint(double x) { return (long)std::lround(x); }  // Fortran NINT: half away from zero
inline double sign(double a, double b) { return std::copysign(std::fabs(a), b); }

// ---------------------------------------------------------------------------
// Run configuration = the namelist values of SURVEY.md Appendix A.
struct Config {
  int nang = 12, nfre = 36, nfre_red = 25;
  int ifre1 = 3;  // 1 when nfre_red == 25 (share/ecwam/scripts/ecwam_configure.sh:52-57)
  double fr1 = 4.177248e-02;
  int iphys = 1;
  int isnonlin = 0, idamping = 1, irefra = 0, icase = 1, ipropags = 2;
  int llgcbz0 = 0, llnormagam = 0, llcapchnk = 1, lbiwbk = 1;
  int licerun = 1, lmaskice = 1, lwamrsetci = 1, lciwa1 = 0, lciwa2 = 0, lciwa3 = 0, lciscal = 0;
  int lwflux = 0, lwfluxout = 1, lwnemocou = 0, lwvflx_snl = 1, lwcouast = 0, lwcou = 0;
  int icode = 3;
  double idelt = 900., idelpro = 900., delpro_lf = 900.;
  int ifrelfmax = 0;
  double ximp = 1.0;
  double rnu = 1.5e-5, rnum = 0.11 * 1.5e-5;  // runwam.F90:232-233
  double wspmin = 1.0;                         // userin.F90:913-918 (LLGCBZ0=F)
  double cithrsh = 0.3, cithrsh_tail = 0.3, ciblock = 0.0;
  double flmin = 1e-5, zalpfacx = 1.0;
  double zalpfacb = 1.0, cdicwa = 0.01;   // YOWICE ZALPFACB (mpuserin.F90:780), CDICWA (userin.F90:962-976: 0.01 with LCIWA2)
  double bathymax = 998.999, deptha = 2.0;
  int nproma = 32;
  int npr = 1;  // emulated MPI ranks (in-process)
  int ll1d = 0;
  int store_all_weights = 0;  // keep the reference's 18 weight arrays (tests only; 150 KB/point at 36x29)
  int nthreads = 1;  // OpenMP threads for the CPU-baseline leg (WAM_NPROMA is NOT applied: LLNO_WAM_NPROMA)
  int llcflcuroff = 1;  // YOWSTAT LLCFLCUROFF (mpuserin.F90:575): retry without current refraction where the CFL check failed
  // YOWCOUP NEMO coupling (with lwnemocou): LWNEMOTAUOC, LWNEMOCOUSTK, LWNEMOCOUSTRN, LWNEMOCOUSEND (yowcoup.F90:24-33)
  int lwnemotauoc = 0, lwnemocoustk = 0, lwnemocoustrn = 0, lwnemocousend = 1;
  int lwnemocouwrs = 0, lwnemocouibr = 0;     // YOWCOUP LWNEMOCOUWRS (radiative stress on the ice), LWNEMOCOUIBR (ice break-up memory)
  double zalpwrs = 1.0, zibrw_thrsh = 0.5;   // YOWICE ZALPWRS, ZIBRW_THRSH (mpuserin.F90:784-786)
};

// ---------------------------------------------------------------------------
// Module state (YOWFRED, YOWPCONS, YOWPHYS, YOWINDN, YOWTABL, YOWCOUP ...)
struct Tables {
  // YOWPCONS (yowpcons.F90:19-79; INIWCST)
  double G = 9.806, GM1 = 0.101978381, CIRC = 40007993.95, ROWATER = 1000.0, ROAIR = 1.225;
  double PI, ZPI, ZPI4GM1, ZPI4GM2, RAD, DEG, R, ROWATERM1;
  double EPSMIN = 0.1e-32, EPSUS = 1.0e-6, EPSU10, ACD = 8.0e-4, BCD = 8.0e-5, ACDLIN = 0.0008, BCDLIN = 0.00047,
         CDMAX = 0.0025, DKMAX = 40.0;
  double TAUOCMIN = 0.01, TAUOCMAX = 50.0, PHIEPSMIN = -3276.80, PHIEPSMAX = -0.05, WSEMEAN_MIN = 0.001;
  // YOWFRED
  double FRATIO = 1.1, WETAIL = 0.25, FRTAIL = 0.2, WP1TAIL = 1.0 / 3.0, WP2TAIL = 0.5, COEF4 = 5.0e-07;
  double DELTH, FLOGSPRDM1, XLOGFRATIO;
  int NFRE_ODD;
  ArrD FR, DFIM, GOM, C, TH, COSTH, SINTH, DFIMOFR, DFIMFR, DFIMFR2, ZPIFR, FR5, FRM5, COFRM4, FLMAX, RHOWG_DFIM,
      DFIM_SIM;
  // YOWPHYS (setwavphys.F90)
  double XKAPPA = 0.40, XNLEV = 10.0, BETAMAX, BETAMAXOXKAPPA2, BMAXOKAP, BMAXOKAPDTH, GAMNCONST, ZALP, ALPHA,
         ALPHAMIN, ALPHAMAX = 0.11, CHNKMIN_U, TAUWSHELTER, ALPHAPMAX, TAILFACTOR, TAILFACTOR_PM, DELTA_THETA_RN,
         DTHRN_A, DTHRN_U, RN1_RN;
  double SWELLF = 0.66, SWELLF2 = -0.018, SWELLF3 = 0.022, SWELLF4, SWELLF5 = 1.2, SWELLF6 = 1.0, SWELLF7, SWELLF7M1,
         Z0RAT, Z0TUBMAX, ABMIN = 0.3, ABMAX = 8.0;
  double CDIS, DELTA_SDIS, CDISVIS;
  double SDSBR = 9.0e-4, SSDSC2 = -2.2e-5, SSDSC4 = 1.0, SSDSC6 = 0.3, MICHE = 1.0, SSDSC3 = 0.0, SSDSBRF1 = 0.5,
         BRKPBCOEF = 28.16, SSDSC5 = 0.0;
  int ISDSDTH = 80, ISB = 2, IPSAT = 2, NSDSNTH;
  ArrI INDICESSAT;
  ArrD SATWEIGHTS;
  double EGRCRV, AFCRV, BFCRV;
  // gravity-capillary model (yowfred.F90:61-65, yowpcons.F90:47-48, initgc.F90, setwavphys.F90: ANG_GC_*)
  double SURFT = 0.0000717, SQRTGOSURFT = 0.0, ANG_GC_A = 0.0, ANG_GC_B = 0.0, ANG_GC_C = 0.0;
  int NWAV_GC = 0;
  ArrD XK_GC, XKM_GC, OMEGA_GC, OMXKM3_GC, VG_GC, C_GC, CM_GC, C2OSQRTVG_GC, XKMSQRTVGOC2_GC, OM3GMKM_GC, DELKCC_GC, DELKCC_GC_NS,
      DELKCC_OMXKM3_GC;
  // YOWTABL
  int IAB = 200;
  double EPS1 = 0.00001;
  ArrD SWELLFT;
  // YOWCOUP (init_x0tauhf.F90)
  int JTOT_TAUHF = 19;
  double X0TAUHF;
  ArrD WTAUHF;
  // YOWINDN (nlweigt.F90 / inisnonlin.F90)
  int MFRSTLW, MLSTHG, KFRH;
  ArrI IKP, IKP1, IKM, IKM1, K1W, K2W, K11W, K21W, INLCOEF;
  ArrD AF11, FKLAP, FKLAP1, FKLAM, FKLAM1, FRH, RNLCOEF, FTRF;
  double ACL1, ACL2, CL11, CL21, DAL1, DAL2;
  // YOWICE (cigetdeac.F90:60-75): Kohout & Meylan's attenuation table of SDICE1
  int NICT = 0, NICH = 0;
  double TICMIN = 1.0, HICMIN = 0.2, DTIC = 1.0, DHIC = 0.1;
  ArrD CIDEAC;   // (NICT,NICH)
};

void init_tables(const Config& c, Tables& t);

// ---------------------------------------------------------------------------
// Grid + decomposition (YOWMAP, YOWGRID, YOWUBUF, YOWSPEC, YOWMPP) for ONE emulated rank.
struct Grid {
  int NGY = 0, NGX = 0, NIBLO = 0, IPER = 1, IRGG = 1;
  double AMOWEP = 0, AMOSOP, AMOEAP, AMONOP, XDELLA, XDELLO;
  ArrI NLONRGG;              // (NGY)
  ArrD ZDELLO, DELLAM, SINPH, COSPH;  // (NGY)
  std::vector<std::vector<unsigned char>> MASK;  // LLOCEANMASK(i,k): MASK[k-1][i-1]
  ArrI IXLG, KXLT;           // BLK2GLO in NEW (relabelled) IJ order, (NIBLO)
  ArrI IXLG0, KXLT0;         // original 1-D order
  ArrI NEWIJ2IJ, IJ2NEWIJ;   // (0:NIBLO)
  std::vector<std::vector<int>> IJMAP;  // original-order ij of (i,k) or 0 (helper for PROPCONNECT's searches)
  int NPR = 1;
  ArrI NSTART, NEND, KLENBOT, KLENTOP;  // (NPR)
};

struct RankDecomp {
  int IRANK = 1, NINF, NSUP, IJS, IJL;
  ArrI KLAT, KLON, KCOR;   // (IJS:IJL,2,2) (IJS:IJL,2) (IJS:IJL,4,2)  -- local halo addressing after MPDECOMP
  ArrD WLAT, WCOR;         // (IJS:IJL,2) (IJS:IJL,4)
  ArrI NFROMPE, NTOPE, NIJSTART, IJTOPE, NTOPELST, NFROMPELST;
  int NTOPEMAX = 0, NFROMPEMAX = 0, NGBTOPE = 0, NGBFROMPE = 0;
  // MCHUNK
  int NPROMA, NCHNK;
  ArrI KIJL4CHNK, IJFROMCHNK;
  // CTUWUPDT index helpers
  ArrI MPM, KPM, JXO, JYO, KCR;
  // CTU weights (stored, as the reference does)
  ArrD SUMWN, WLATN, WLONN, WCORN, WKPMN;  // all 18 arrays: only when cfg.store_all_weights (ctuwupdt.F90:171-178) or IREFRA >= 2
  ArrD WMPMN;                              // frequency-shift weights (IREFRA = 2, 3; ctuw.F90:503-525)
  ArrD W8;  // (IJ,K,M,8): the 8 weights PROPAGS2 reads when IREFRA=0 (propags2.F90:107-116)
  // YOWUBUF OBSLON(IJS:IJL,NFRE_RED,2), OBSLAT(IJS:IJL,NFRE_RED,2), OBSCOR(IJS:IJL,NFRE_RED,4): sub-grid obstruction coefficients
  // (getbobstrct.F90:395-500); empty = LSUBGRID F (all 1)
  ArrD OBSLON, OBSLAT, OBSCOR;
  bool LUPDTWGHT = true;
  int cfl_fail = 0;
};

void build_grid(const Config& c, Grid& g, int ngy, const int* nlonrgg, double amosop, double amonop,
                const unsigned char* maskflat);
void mpdecomp(const Config& c, const Tables& t, Grid& g, std::vector<RankDecomp>& ranks);

// ---------------------------------------------------------------------------
// Per-rank model fields in the NPROMA-chunked layout of yowdrvtype_config.yml.
struct Fields {
  ArrD FL1, XLLWS;                                              // (P,A,F,C)
  ArrD WAVNUM, CINV, CGROUP, XK2CG, OMOSNH2KD, STOKFAC, CIWA;  // (P,F,C)
  ArrD DEPTH, EMAXDPT, DELLAM1, COSPHM1, UCUR, VCUR, IBRMEM;   // (P,C)
  ArrI INDEP, IODP, IOBND;
  ArrD AIRD, WDWAVE, CICOVER, WSWAVE, WSTAR, USTRA, VSTRA, UFRIC, TAUW, TAUWDIR, Z0M, Z0B, CHRNCK, CITHICK;
  ArrD WSEMEAN, WSFMEAN, USTOKES, VSTOKES, STRNMS, TAUXD, TAUYD, TAUOCXD, TAUOCYD, TAUOC, TAUICX, TAUICY, PHIOCD,
      PHIEPS, PHIAW;
  ArrI MIJ;
  // NEMO coupling fields (P,C) (yowdrvtype_config.yml WAVE2OCEAN; JWRO = double here): accumulated / overwritten by WNFLUXES and STOKESTRN
  ArrD NSWH, NMWP, NPHIEPS, NTAUOC, NEMOTAUX, NEMOTAUY, NEMOTAUICX, NEMOTAUICY, NEMOWSWAVE, NEMOPHIF, NEMOUSTOKES, NEMOVSTOKES, NEMOSTRN;
  // test hook (orc_capture): the inputs of WNFLUXES that are internal to IMPLSCH, kept from the last implsch_chunk call
  bool capture = false;
  ArrD DBG_SSOURCE;            // (P,A,F,C)
  ArrD DBG_EM, DBG_F1, DBG_PHIWA;   // (P,C): EMEAN, F1MEAN of the first FKMEAN, PHIWA of the second STRESSO
  // land-point dispersion (WVPRPT_LAND, initdpthflds.F90:80-88)
  ArrD LAND_WAVNUM, LAND_CGROUP, LAND_OMOSNH2KD;
};

// selection of OUTBLOCK output columns (orc_output.cpp)
struct OutSel {
  int n = 0;                            // NIPRMOUT
  std::vector<int> itg, icemask, seamask;  // per BOUT column: reference parameter number, IPRMINFO(:,6), IPRMINFO(:,7)
  double zmiss = -999.0;                // YOWPCONS ZMISS
  int llsource = 1;                     // YOWSTAT LLSOURCE
};

struct Model {
  Config cfg;
  Tables tab;
  Grid grid;
  std::vector<RankDecomp> ranks;
  std::vector<Fields> fld;
  std::vector<double> depth0;  // depth per sea point, original order (1..NIBLO)
  OutSel sel;                              // last orc_outbs selection
  std::vector<std::vector<double>> bout;   // [rank] BOUT (P, NIPRMOUT, C) of the last orc_outbs
};

void alloc_fields(Model& m);
void depthprpt(const Tables& t, const Config& c, long n, const double* depth, double* wavnum, double* cinv,
               double* cgroup, double* xk2cg, double* omosnh2kd, double* stokfac);
void propag_wam(Model& m);  // all ranks, with in-process MPEXCHNG
void implsch_all(Model& m); // all ranks, all chunks (wamintgr.F90:117-146)
void implsch_chunk(const Config& c, const Tables& t, Fields& f, int KIJL, int ichnk);
void femean(const Tables& t, const Config& c, int KIJL, const double* F /*(KIJL,A,F)*/, double* EM, double* FM);
// SNONLIN alone on one chunk (tests: conservation properties of the DIA); SL, FLD (KIJL,A,F) are overwritten
void snonlin_chunk(const Config& c, const Tables& t, Fields& f, int KIJL, int ICHNK, double* SL, double* FLD);
void term_chunk(const Config& c, const Tables& t, Fields& f, int KIJL, int ICHNK, int which, double* SL, double* FLD);
void stresso_chunk(const Config& c, const Tables& t, Fields& f, int KIJL, int ICHNK, double* SL, double* SPOS, double* OUT);

// ---------------------------------------------------------------------------
// The steps either side of the hot path (orc_output.cpp): NEWWIND, OUTBLOCK core parameters, WAMNORM statistics.
bool outparam_supported(int itg);
double aki(const Tables& t, double OM, double BETA);   // aki.F90:71-91
void newwind(Model& m, int ir, const Fields& next);
void outblock(const Config& c, const Tables& t, Fields& f, int KIJL, int ICHNK, const OutSel& sel, double* BOUT /*(KIJL,NIPRMOUT)*/);
void mpminmaxavg(Model& m, const OutSel& sel, const std::vector<std::vector<double>>& bout /*[rank](P,NIPRMOUT,C)*/, bool global,
                 double* WNORM /*(4,NIPRMOUT)*/);

}  // namespace orc
