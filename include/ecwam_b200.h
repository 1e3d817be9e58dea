/* ecwam_b200.h — C ABI of the B200-native WAMINTGR hot path (PROPAG_WAM/PROPAGS2 + IMPLSCH).
 *
 * This is the drop-in boundary: a Fortran ecWAM build binds these entry points with ISO_C_BINDING and keeps
 * the reference signatures of IMPLSCH (src/ecwam/implsch.F90:10-23) and PROPAG_WAM (src/ecwam/propag_wam.F90:10-11)
 * as thin shims (INTEGRATION.md shows the shim).  Plain pointers and sizes only, no torch / C++ types.
 *
 * All paths below are relative to the reference tree (/root/reference).
 *
 * Conventions
 *   - reals are IEEE double (JWRB in a dp build, src/ecwam/parkind_wave.F90:23-35), integers are 32-bit (JWIM);
 *   - arrays keep the reference's column-major NPROMA-blocked layouts (src/ecwam/yowdrvtype_config.yml:12-56):
 *       FL1, XLLWS        (NPROMA, NANG, NFRE, NCHNK)
 *       WAVNUM ... CIWA   (NPROMA, NFRE, NCHNK)
 *       1-D fields        (NPROMA, NCHNK)
 *     the library never re-orders or frees caller memory;
 *   - every function returns 0 on success, a negative ECWAM_B200_E* code on failure
 *     (ecwam_b200_propag returns the number of CFL-violating grid points, >0, where the reference would
 *     call ABORT1 in src/ecwam/ctuwdrv.F90:127-146); ecwam_b200_last_error() gives the message;
 *   - one handle per MPI rank / GPU, driven from one host thread (SURVEY.md 8b "Threading").
 *
 * NOTE for ecwam_b200/lib.py: the struct bodies are parsed to build the ctypes mirrors — keep ONE member
 * per line, of the forms `int x;`, `double x;`, `const double* x;`, `const int* x;`, `double* x;`, `int* x;`.
 */
#ifndef ECWAM_B200_H
#define ECWAM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ECWAM_B200_EINVAL (-1)   /* bad argument / unsupported configuration            */
#define ECWAM_B200_ECUDA (-2)    /* CUDA runtime error                                  */
#define ECWAM_B200_ENCCL (-3)    /* NCCL error                                          */
#define ECWAM_B200_ESTATE (-4)   /* call out of order (e.g. fields not bound)           */
#define ECWAM_B200_EIO (-5)      /* restart / grid-table file: open, size or record error */

/* ---------------------------------------------------------------------------------------------------
 * Run parameters = the NALINE namelist values the hot path reads (src/ecwam/mpuserin.F90:180-260,
 * defaults :540-800; values written by share/ecwam/scripts/ecwam_run_model.sh:211-269).               */
typedef struct ecwam_b200_params {
  int nang;        /* YOWPARAM NANG                                                    */
  int nfre;        /* YOWPARAM NFRE   (physics frequencies, 36)                         */
  int nfre_red;    /* YOWPARAM NFRE_RED (propagated frequencies)                        */
  int iphys;       /* YOWSTAT IPHYS: 0 = Janssen/WAM4, 1 = Ardhuin et al. 2010          */
  int isnonlin;    /* YOWSTAT ISNONLIN 0, 1, 2 (snonlin.F90:127-163; 1, 2: k_enh)          */
  int idamping;    /* YOWSTAT IDAMPING (SINPUT_JAN)                                     */
  int irefra;      /* YOWSTAT IREFRA (0 none, 1 depth, 2 current, 3 depth + current refraction)   */
  int icase;       /* YOWSTAT ICASE (1 = spherical, only)                               */
  int llgcbz0;     /* YOWCOUP LLGCBZ0 (needs the *_gc members of ecwam_b200_tables)      */
  int llnormagam;  /* YOWCOUP LLNORMAGAM (needs the *_gc members of ecwam_b200_tables)   */
  int llcapchnk;   /* YOWCOUP LLCAPCHNK                                                 */
  int lbiwbk;      /* YOWSTAT LBIWBK                                                    */
  int licerun;     /* YOWICE LICERUN                                                    */
  int lmaskice;    /* YOWICE LMASKICE                                                   */
  int lwamrsetci;  /* YOWICE LWAMRSETCI                                                 */
  int lciwa;       /* YOWICE bit mask: 1 LCIWA1 (SDICE1, needs ecwam_b200_tables cideac), 2 LCIWA2 (SDICE2), 4 LCIWA3 (SDICE3), 8 LCISCAL */
  int lwflux;      /* YOWCOUP LWFLUX                                                    */
  int lwfluxout;   /* YOWCOUP LWFLUXOUT (userin.F90:470 sets it .TRUE.)                 */
  int lwnemocou;   /* YOWCOUP LWNEMOCOU (needs ecwam_b200_bind_nemo)                   */
  int lwvflx_snl;  /* YOWCOUP LWVFLX_SNL                                                */
  int lwcouast;    /* YOWCOUP LWCOUAST                                                  */
  int icode_wnd;   /* YOWWNDG ICODE: 3 = 10 m wind; 1, 2 = friction velocity / stress (UFRIC is the forcing) */
  int ifrelfmax;   /* YOWSTAT IFRELFMAX (fast-wave sub-stepping, O1280)                 */
  int nproma;      /* YOWPARAM NPROMA_WAM                                               */
  int nchnk;       /* YOWPARAM NCHNK                                                    */
  double idelt;    /* YOWSTAT IDELT   [s]                                               */
  double idelpro;  /* YOWSTAT IDELPRO [s]                                               */
  double delpro_lf;/* YOWSTAT DELPRO_LF [s]                                             */
  double ximp;     /* YOWSTAT XIMP                                                      */
  double rnu;      /* YOWPHYS RNU                                                       */
  double rnum;     /* YOWPHYS RNUM                                                      */
  double wspmin;   /* YOWWIND WSPMIN                                                    */
  double cithrsh;  /* YOWICE CITHRSH                                                    */
  double cithrsh_tail; /* YOWICE CITHRSH_TAIL                                           */
  double ciblock;  /* YOWICE CIBLOCK                                                    */
  double flmin;    /* YOWICE FLMIN                                                      */
  double bathymax; /* YOWSHAL BATHYMAX                                                  */
  int llcflcuroff; /* YOWSTAT LLCFLCUROFF (IREFRA = 2, 3): retry the CFL check without current refraction */
  double zalpfacx; /* YOWICE ZALPFACX (attenuation factor of SDICE3, 1 = no reduction)    */
  double zalpfacb; /* YOWICE ZALPFACB (scales SDICE1 / SDICE2; mpuserin.F90:780: 1)             */
  double cdicwa;   /* YOWICE CDICWA (ice-water drag coefficient of SDICE2; userin.F90:974: 0.01) */
  int lwnemotauoc;   /* YOWCOUP LWNEMOTAUOC: NEMOTAUX/Y accumulate TAUOCXD/YD instead of TAUXD/YD (wnfluxes.F90:317-323) */
  int lwnemocoustk;  /* YOWCOUP LWNEMOCOUSTK: NEMOUSTOKES/NEMOVSTOKES = USTOKES/VSTOKES, else 0 (stokestrn.F90:79-85)   */
  int lwnemocoustrn; /* YOWCOUP LWNEMOCOUSTRN: CIMSSTRN fills STRNMS (and NEMOSTRN with LWNEMOCOU) (stokestrn.F90:70-74, 87) */
  int lwnemocousend; /* YOWCOUP LWNEMOCOUSEND (read with LWCOU only, stokestrn.F90:76-78)                             */
  int lwcou;         /* YOWCOUP LWCOU (only used in that condition)                                                 */
  int lwnemocouwrs;  /* YOWCOUP LWNEMOCOUWRS: TAUICX/Y = wave radiative stress on the sea ice from SDICE's SLICE (wnfluxes.F90:178-196, 266-271) */
  int lwnemocouibr;  /* YOWCOUP LWNEMOCOUIBR: SDICE3's ALPFAC = 1/ZALPFACX where IBRMEM <= ZIBRW_THRSH (icebreak_modify_attenuation.F90:82-94);
                        needs ecwam_b200_nemo_fields.ibrmem */
  double zalpwrs;    /* YOWICE ZALPWRS (mpuserin.F90:784: 1)                                                        */
  double zibrw_thrsh;/* YOWICE ZIBRW_THRSH (mpuserin.F90:786: 0.5)                                                  */
} ecwam_b200_params;

/* ---------------------------------------------------------------------------------------------------
 * Small read-only tables = module variables of YOWPCONS, YOWFRED, YOWPHYS, YOWTABL, YOWCOUP, YOWINDN
 * (SURVEY.md 8b "Global state consumed").  Host pointers; copied at ecwam_b200_create.
 * ecwam_b200_host_tables() fills one from `params` alone (stand-alone / synthetic runs).              */
typedef struct ecwam_b200_tables {
  /* YOWPCONS (src/ecwam/yowpcons.F90:19-79, iniwcst.F90) */
  double g;
  double gm1;
  double zpi;
  double zpi4gm1;
  double zpi4gm2;
  double circ;
  double r_earth;
  double rowaterm1;
  double epsmin;
  double epsus;
  double epsu10;
  double acd;
  double bcd;
  double cdmax;
  double tauocmin;
  double tauocmax;
  double phiepsmin;
  double phiepsmax;
  double wsemean_min;
  /* YOWFRED (src/ecwam/yowfred.F90:20-120, mfredir.F90, initmdl.F90:437-503) */
  double fratio;
  double wetail;
  double frtail;
  double wp1tail;
  double delth;
  double flogsprdm1;
  int nfre_odd;
  const double* fr;          /* (NFRE) */
  const double* dfim;        /* (NFRE) */
  const double* dfimofr;     /* (NFRE) */
  const double* dfimfr;      /* (NFRE) */
  const double* zpifr;       /* (NFRE) */
  const double* fr5;         /* (NFRE) */
  const double* cofrm4;      /* (NFRE) */
  const double* flmax;       /* (NFRE) */
  const double* rhowg_dfim;  /* (NFRE) */
  const double* dfim_sim;    /* (NFRE) */
  const double* th;          /* (NANG) */
  const double* costh;       /* (NANG) */
  const double* sinth;       /* (NANG) */
  /* YOWPHYS (src/ecwam/yowphys.F90, setwavphys.F90:46-205) */
  double xkappa;
  double xnlev;
  double alpha;
  double alphamin;
  double chnkmin_u;
  double zalp;
  double betamaxoxkappa2;
  double tauwshelter;
  double tailfactor;
  double tailfactor_pm;
  double swellf;
  double swellf2;
  double swellf3;
  double swellf4;
  double swellf5;
  double swellf6;
  double swellf7;
  double swellf7m1;
  double z0rat;
  double z0tubmax;
  double abmin;
  double abmax;
  double cdis;
  double delta_sdis;
  double cdisvis;
  double sdsbr;
  double ssdsc2;
  double ssdsc3;
  double ssdsc4;
  double ssdsc5;
  double ssdsc6;
  double miche;
  double egrcrv;
  double afcrv;
  double bfcrv;
  int nsdsnth;
  const int* indicessat;     /* (NANG, 2*NSDSNTH+1), 1-based direction indices (init_sdiss_ardh.F90:69-96) */
  const double* satweights;  /* (NANG, 2*NSDSNTH+1) */
  /* YOWTABL / YOWCOUP (tabu_swellft.F90:64-83, init_x0tauhf.F90:65-100) */
  int iab;
  double eps1;
  const double* swellft;     /* (IAB) */
  int jtot_tauhf;
  double x0tauhf;
  const double* wtauhf;      /* (JTOT_TAUHF) */
  /* YOWINDN (nlweigt.F90:94-262, inisnonlin.F90:89-270) */
  int mfrstlw;
  int mlsthg;
  int kfrh;
  double dal1;
  double dal2;
  const int* ikp;            /* (MFRSTLW:MLSTHG) */
  const int* ikp1;           /* (MFRSTLW:MLSTHG) */
  const int* ikm;            /* (MFRSTLW:MLSTHG) */
  const int* ikm1;           /* (MFRSTLW:MLSTHG) */
  const int* k1w;            /* (NANG,2) 1-based */
  const int* k2w;            /* (NANG,2) */
  const int* k11w;           /* (NANG,2) */
  const int* k21w;           /* (NANG,2) */
  const int* inlcoef;        /* (5, MLSTHG) */
  const double* rnlcoef;     /* (25, MLSTHG) */
  const double* af11;        /* (MFRSTLW:MLSTHG) */
  /* Gravity-capillary roughness model and growth renormalisation (LLGCBZ0 / LLNORMAGAM, the cy49r1 physics):
   * YOWPHYS (yowphys.F90:45-75, setwavphys.F90:46-205, init_x0tauhf.F90:65-72), YOWPCONS ACDLIN/BCDLIN
   * (yowpcons.F90:58-59), YOWFRED *_GC (yowfred.F90:62-150, initgc.F90:63-110).  Read only when the switches are on. */
  double alphamax;
  double alphapmax;
  double acdlin;
  double bcdlin;
  double bmaxokap;
  double gamnconst;
  double rn1_rn;
  double dthrn_a;
  double dthrn_u;
  double ang_gc_a;
  double ang_gc_b;
  double ang_gc_c;
  double sqrtgosurft;
  int nwav_gc;
  const double* xk_gc;            /* (NWAV_GC) */
  const double* omega_gc;         /* (NWAV_GC) */
  const double* cm_gc;            /* (NWAV_GC) */
  const double* c2osqrtvg_gc;     /* (NWAV_GC) */
  const double* xkmsqrtvgoc2_gc;  /* (NWAV_GC) */
  const double* om3gmkm_gc;       /* (NWAV_GC) */
  const double* omxkm3_gc;        /* (NWAV_GC) */
  const double* delkcc_gc_ns;     /* (NWAV_GC) */
  const double* delkcc_omxkm3_gc; /* (NWAV_GC) */
  const double* delkcc_gc;        /* (NWAV_GC) read by MEANSQS_GC (OUTBLOCK parameter 9) */
  /* YOWICE (yowice.F90:20-31, cigetdeac.F90:60-75): Kohout & Meylan's attenuation table of SDICE1.  Read only with LCIWA1. */
  int nict;                       /* wave-period dimension of CIDEAC (16)      */
  int nich;                       /* ice-thickness dimension of CIDEAC (36)    */
  double ticmin;                  /* first wave period of the table [s] (1)    */
  double hicmin;                  /* first ice thickness of the table [m] (0.2) */
  double dtic;                    /* wave-period increment [s] (1)             */
  double dhic;                    /* ice-thickness increment [m] (0.1)         */
  const double* cideac;           /* (NICT, NICH) column-major                 */
} ecwam_b200_tables;

/* ---------------------------------------------------------------------------------------------------
 * Per-rank grid / decomposition tables = YOWMAP, YOWUBUF, YOWSPEC, YOWMPP after MPDECOMP
 * (src/ecwam/mpdecomp.F90:690-1296, propconnect.F90, yowubuf.F90:59-96, yowspec.F90:17-33).
 * Indices keep the reference's 1-based extended numbering: own points IJS..IJL, halo NINF..IJS-1 and
 * IJL+1..NSUP, land = NSUP+1.  Host pointers; copied at ecwam_b200_create.                            */
typedef struct ecwam_b200_decomp {
  int irank;                 /* 1-based rank (YOWMPP IRANK) */
  int nproc;                 /* YOWMPP NPROC */
  int ijs;                   /* NSTART(IRANK) */
  int ijl;                   /* NEND(IRANK) */
  int ninf;                  /* YOWMPP NINF */
  int nsup;                  /* YOWMPP NSUP */
  int ngy;                   /* YOWPARAM NGY */
  double xdella;             /* YOWMAP XDELLA [deg] */
  const double* zdello;      /* (NGY) YOWMAP ZDELLO [deg] */
  const double* cosph;       /* (NGY) YOWMAP COSPH */
  const double* sinph;       /* (NGY) YOWMAP SINPH */
  const int* kxlt;           /* (IJS:IJL) BLK2GLO%KXLT latitude row of each own point */
  const int* klat;           /* (IJS:IJL,2,2) */
  const int* klon;           /* (IJS:IJL,2)   */
  const int* kcor;           /* (IJS:IJL,4,2) */
  const double* wlat;        /* (IJS:IJL,2)  as left by PROPCONNECT (CTUWINI's edit is applied internally) */
  const double* wcor;        /* (IJS:IJL,4)   */
  const int* nfrompe;        /* (NPROC) */
  const int* ntope;          /* (NPROC) */
  const int* nijstart;       /* (NPROC) */
  int ntopemax;
  const int* ijtope;         /* (NTOPEMAX, NPROC) */
  const double* land_cgroup; /* (NFRE_RED) WVPRPT_LAND%CGROUP (initdpthflds.F90:80-88) */
  /* YOWUBUF sub-grid obstruction coefficients (LSUBGRID = T; getbobstrct.F90:395-500, applied by ctuw.F90:700-733).
   * All three NULL = LSUBGRID F (every coefficient 1, the setting of ecwam_run_model.sh:238).                */
  const double* obslon;      /* (IJS:IJL, NFRE_RED, 2) */
  const double* obslat;      /* (IJS:IJL, NFRE_RED, 2) */
  const double* obscor;      /* (IJS:IJL, NFRE_RED, 4) */
} ecwam_b200_decomp;

/* ---------------------------------------------------------------------------------------------------
 * Model fields (device pointers, or host pointers for ecwam_b200_wamintgr_host) = the arguments of
 * IMPLSCH (implsch.F90:117-143) and PROPAG_WAM (propag_wam.F90:74-77) over ALL chunks.                */
typedef struct ecwam_b200_fields {
  double* fl1;               /* (P,A,F,C) inout */
  double* xllws;             /* (P,A,F,C) out   */
  const double* wavnum;      /* (P,F,C) */
  const double* cinv;
  const double* cgroup;
  const double* xk2cg;
  const double* omosnh2kd;   /* PROPAG_WAM only (read when IREFRA = 1) */
  const double* stokfac;
  const double* ciwa;        /* unused (LCIWA*=F) */
  const double* depth;       /* (P,C) */
  const double* emaxdpt;
  const double* dellam1;
  const double* cosphm1;
  const double* ucur;        /* surface current (read when IREFRA = 2, 3) */
  const double* vcur;
  double* aird;
  double* wdwave;
  double* cicover;
  double* wswave;
  double* wstar;
  double* ustra;
  double* vstra;
  double* ufric;
  double* tauw;
  double* tauwdir;
  double* z0m;
  double* z0b;
  double* chrnck;
  double* cithick;
  double* wsemean;
  double* wsfmean;
  double* ustokes;
  double* vstokes;
  double* strnms;
  double* tauxd;
  double* tauyd;
  double* tauocxd;
  double* tauocyd;
  double* tauoc;
  double* tauicx;
  double* tauicy;
  double* phiocd;
  double* phieps;
  double* phiaw;
  int* mij;                  /* (P,C) out */
} ecwam_b200_fields;

typedef struct ecwam_b200_handle_s* ecwam_b200_handle;

/* Create the per-rank state: uploads tables, builds the device neighbour tables and halo plan.
 * nccl_comm: an ncclComm_t (NULL when nproc==1, or when ecwam_b200_set_exchange supplies the halo exchange);
 * cuda_stream: a cudaStream_t (NULL = default stream).
 * Replaces the one-off set-up the reference does in INITMDL/MPDECOMP/CTUWUPDT index helpers
 * (src/ecwam/ctuwupdt.F90:93-166).                                                                     */
int ecwam_b200_create(const ecwam_b200_params* params, const ecwam_b200_tables* tables,
                      const ecwam_b200_decomp* decomp, void* nccl_comm, void* cuda_stream,
                      ecwam_b200_handle* out);
int ecwam_b200_destroy(ecwam_b200_handle h);

/* NCCL bootstrap helpers so that a host without its own NCCL binding (Fortran, ctypes) can build the
 * communicator: rank 0 calls _nccl_unique_id, broadcasts the 128 bytes by any means (MPL_BROADCAST in
 * ecWAM), every rank calls _nccl_comm_init.                                                            */
int ecwam_b200_nccl_unique_id(char id_out[128]);
int ecwam_b200_nccl_comm_init(const char id[128], int nranks, int rank, void** comm_out);
int ecwam_b200_nccl_comm_destroy(void* comm);

/* MPEXCHNG through the host's own message passing instead of NCCL (src/ecwam/mpexchng.F90:164-206 posts MPL_SEND / MPL_RECV per
 * neighbouring rank; a host that keeps that code passes nccl_comm = NULL to ecwam_b200_create and registers this callback).
 * Whenever the library needs a halo exchange (the spectrum in PROPAG_WAM, the group velocity / depth / currents in PROENVHALO) it
 * packs the send buffer on the device, synchronises its stream and calls fn: for every rank q, send_count[q] doubles starting at
 * sendbuf + send_offset[q] go to q, and recv_count[q] doubles from q land at recvbuf + recv_offset[q] (counts may be 0; the
 * message order between a pair of ranks is the call order).  fn returns 0 when recvbuf is complete.
 * staged = 0: sendbuf / recvbuf are DEVICE pointers (CUDA-aware MPI); staged = 1: the library stages both through pinned HOST
 * buffers and fn sees host pointers (plain MPI, or the gloo-staged two-processes-on-one-GPU test).  fn = NULL restores NCCL.
 * OUTWNORM over several ranks still needs the NCCL communicator.                                                              */
typedef int (*ecwam_b200_exchange_fn)(void* user, int nproc, const double* sendbuf, const long long* send_offset,
                                      const long long* send_count, double* recvbuf, const long long* recv_offset,
                                      const long long* recv_count);
int ecwam_b200_set_exchange(ecwam_b200_handle h, ecwam_b200_exchange_fn fn, void* user, int staged);

/* Bind DEVICE pointers of the model fields (FIELD_API GET_DEVICE_DATA_* + C_LOC on the Fortran side,
 * src/ecwam/wamintgr_loki_gpu.F90:141-157).                                                            */
int ecwam_b200_bind_fields(ecwam_b200_handle h, const ecwam_b200_fields* dev);

/* NEMO coupling fields = the WAVE2OCEAN arguments of IMPLSCH in the reference order (implsch.F90:17-19), DEVICE pointers over all
 * chunks, (P,C) doubles (JWRO = JWRB in the double-precision build).  Required with LWNEMOCOU.  WNFLUXES (wnfluxes.F90:304-330)
 * overwrites NPHIEPS, NTAUOC, NSWH, NMWP and ADDS to NEMOTAUX/Y, NEMOWSWAVE, NEMOPHIF, NEMOTAUICX/Y (the caller zeroes them after a
 * coupling exchange); STOKESTRN (stokestrn.F90:76-88) overwrites NEMOUSTOKES, NEMOVSTOKES and, with LWNEMOCOUSTRN, NEMOSTRN.       */
typedef struct ecwam_b200_nemo_fields {
  double* nemoustokes;
  double* nemovstokes;
  double* nemostrn;
  double* nphieps;
  double* ntauoc;
  double* nswh;
  double* nmwp;
  double* nemotaux;
  double* nemotauy;
  double* nemotauicx;
  double* nemotauicy;
  double* nemowswave;
  double* nemophif;
  const double* ibrmem;   /* OCEAN2WAVE: IBRMEM, the ice break-up memory (read with LWNEMOCOUIBR; may be NULL otherwise) */
} ecwam_b200_nemo_fields;
int ecwam_b200_bind_nemo(ecwam_b200_handle h, const ecwam_b200_nemo_fields* dev);

/* PROPAG_WAM over the whole local block (src/ecwam/propag_wam.F90:10-419, IPROPAGS=2, IREFRA=0..3):
 * halo exchange of FL1 (MPEXCHNG -> NCCL), first-call CTU weight set-up + CFL check (CTUWUPDT),
 * PROPAGS2 (+ fast-wave sub-steps), result back in FL1 with padded lanes refreshed.  IREFRA = 1: depth refraction;
 * IREFRA = 2, 3: advection by / refraction and frequency shift due to the surface current (UCUR, VCUR), with CTUWDRV's
 * LLCFLCUROFF retry.
 * Returns the number of own grid points that violate the CFL / weight-range checks (0 = ok).           */
int ecwam_b200_propag(ecwam_b200_handle h);
/* Force the CTU set-up to be redone at the next ecwam_b200_propag (LUPDTWGHT, getcurr.F90:289).  Also makes the next
 * ecwam_b200_wamintgr_host call upload its static fields again (see there).                             */
int ecwam_b200_invalidate_weights(ecwam_b200_handle h);

/* IMPLSCH for all NCHNK chunks in one go = the chunk loop of WAMINTGR (src/ecwam/wamintgr.F90:117-146). */
int ecwam_b200_implsch_all(ecwam_b200_handle h);
/* IMPLSCH for chunks ichnk0 .. ichnk0+nchnk-1 (1-based): the single-chunk reference call is nchnk=1.   */
int ecwam_b200_implsch(ecwam_b200_handle h, int ichnk0, int nchnk);

/* The reference argument lists, for Fortran bodies that keep the reference's call signatures (fortran/implsch_b200.F90,
 * fortran/propag_wam_b200.F90):
 *   SUBROUTINE IMPLSCH(KIJS, KIJL, FL1, WAVNUM, ..., MIJ, XLLWS)                       src/ecwam/implsch.F90:10-23, 117-143
 *   SUBROUTINE PROPAG_WAM(BLK2GLO, WAVNUM, CGROUP, OMOSNH2KD, FL1, DEPTH, DELLAM1, COSPHM1, UCUR, VCUR)   propag_wam.F90:10-11, 74-77
 * (BLK2GLO is consumed at ecwam_b200_create).  The arguments are the DEVICE addresses of what the reference caller passes —
 * for IMPLSCH the slices of chunk ICHNK, `FL1(:,:,:,ICHNK)` etc. (wamintgr.F90:119-146).  The kernels work on the arrays bound
 * with ecwam_b200_bind_fields: every argument is checked to be chunk ICHNK (derived from FL1's address) of its bound array, so
 * that a FIELD_API re-allocation that was not followed by a re-bind fails with ECWAM_B200_ESTATE instead of using stale
 * memory.  KIJS:KIJL must be 1:NPROMA_WAM as in the reference call.  IOBND, IODP, IBRMEM are accepted and not read; the NEMO
 * accumulators are checked against ecwam_b200_bind_nemo's arrays when LWNEMOCOU is on and ignored otherwise.                                                            */
int ecwam_b200_implsch_f(ecwam_b200_handle h, int kijs, int kijl, double* fl1, const double* wavnum, const double* cgroup,
                         const double* ciwa, const double* cinv, const double* xk2cg, const double* stokfac, const double* emaxdpt,
                         const double* depth, const int* iobnd, const int* iodp, const double* ibrmem, double* aird, double* wdwave,
                         double* cicover, double* wswave, double* wstar, double* ustra, double* vstra, double* ufric, double* tauw,
                         double* tauwdir, double* z0m, double* z0b, double* chrnck, double* cithick, double* nemoustokes,
                         double* nemovstokes, double* nemostrn, double* nphieps, double* ntauoc, double* nswh, double* nmwp,
                         double* nemotaux, double* nemotauy, double* nemotauicx, double* nemotauicy, double* nemowswave,
                         double* nemophif, double* wsemean, double* wsfmean, double* ustokes, double* vstokes, double* strnms,
                         double* tauxd, double* tauyd, double* tauocxd, double* tauocyd, double* tauoc, double* tauicx,
                         double* tauicy, double* phiocd, double* phieps, double* phiaw, int* mij, double* xllws);
int ecwam_b200_propag_wam_f(ecwam_b200_handle h, const double* wavnum, const double* cgroup, const double* omosnh2kd, double* fl1,
                            const double* depth, const double* dellam1, const double* cosphm1, const double* ucur, const double* vcur);

/* One WAMINTGR sub-step with IDELPRO == IDELT: PROPAG_WAM then IMPLSCH (wamintgr.F90:94-146) with the
 * block->chunk copy of PROPAG_WAM fused into IMPLSCH's load.  Same results as _propag + _implsch_all.  */
int ecwam_b200_wamintgr(ecwam_b200_handle h);

/* WAMINTGR's branches without a source-term update: llsource_off != 0 = the LLSOURCE = F branch (wamintgr.F90:163-171:
 * MIJ = NFRE, FL1 = MAX(FL1, EPSMIN), XLLWS = 0); llsource_off == 0 = "not yet time to integrate" (:188-195: MIJ = NFRE,
 * XLLWS = 0).                                                                                           */
int ecwam_b200_no_source(ecwam_b200_handle h, int llsource_off);

/* Same step for callers whose fields live in HOST memory (pinned or pageable): copies the IMPLSCH /
 * PROPAG_WAM inputs host->device, runs the step, copies FL1, XLLWS(optional) and the 1-D outputs back.
 * `host` uses the same struct with host pointers; with_xllws != 0 also returns XLLWS.
 * h2d_bytes / d2h_bytes (optional) receive the bytes moved.
 * Caching: the static fields (WAVNUM, CINV, CGROUP, XK2CG, OMOSNH2KD, STOKFAC, CIWA, DEPTH, EMAXDPT, DELLAM1, COSPHM1,
 * UCUR, VCUR) are uploaded by the first call only; after changing any of them (new currents with IREFRA = 2, 3,
 * getcurr.F90 LUPDTWGHT) call ecwam_b200_invalidate_weights: the next call uploads them again and redoes the CTU set-up.
 * The call works on device mirrors owned by the handle and leaves the binding made with ecwam_b200_bind_fields as it
 * was: the device entry points keep operating on the caller's own tensors.  (The CTU set-up is built from the static
 * fields of one side; a handle that alternates between the two paths rebuilds it at every change of side.)      */
int ecwam_b200_wamintgr_host(ecwam_b200_handle h, const ecwam_b200_fields* host, int with_xllws,
                             long long* h2d_bytes, long long* d2h_bytes);

/* One WAMINTGR sub-step with the model state RESIDENT on the device, as the reference's GPU build keeps it
 * (src/ecwam/wamintgr_loki_gpu.F90:141-201: the spectrum and the fields stay on the device between steps; a step takes the
 * new forcing in and hands the integrated parameters back).  host_next: the eight FF_NEXT fields (struct below) in HOST
 * memory, pinned for full PCIe rate; they are copied in, NEWWIND's field update is applied (ecwam_b200_newwind), then
 * PROPAG_WAM + IMPLSCH run on the tensors bound with ecwam_b200_bind_fields.  host_out (may be NULL): HOST pointers; every
 * non-NULL 1-D output member (UFRIC, TAUW, TAUWDIR, Z0M, Z0B, CHRNCK, WSEMEAN, WSFMEAN, USTOKES, VSTOKES, TAUXD ... PHIAW, MIJ)
 * receives the step's result.  Returns like ecwam_b200_wamintgr; synchronises the handle's stream.           */
struct ecwam_b200_forcing_next;
int ecwam_b200_wamintgr_forced(ecwam_b200_handle h, const struct ecwam_b200_forcing_next* host_next,
                               const ecwam_b200_fields* host_out, long long* h2d_bytes, long long* d2h_bytes);

/* ---------------------------------------------------------------------------------------------------
 * The steps either side of the hot path each time step / output step, kept on the device.   */

/* FF_NEXT members NEWWIND copies (device pointers, (P,C)); src/ecwam/newwind.F90:127-162.                  */
typedef struct ecwam_b200_forcing_next {
  const double* wswave;
  const double* wdwave;
  const double* aird;
  const double* wstar;
  const double* cicover;
  const double* cithick;
  const double* ustra;
  const double* vstra;
} ecwam_b200_forcing_next;
/* NEWWIND's update of FF_NOW from FF_NEXT (src/ecwam/newwind.F90:105-167, ICODE_WND = 3: 10 m wind speed with the
 * low-wind cap on the first-guess wave stress TAUW).  The date bookkeeping (CDATEWH, INCDATE) stays with the caller:
 * call this when NEWWIND's `CDATE >= CDATEWH` test holds.                                                   */
int ecwam_b200_newwind(ecwam_b200_handle h, const ecwam_b200_forcing_next* next);
/* The same with the friction velocity as forcing (ICODE_WND = 1, 2; newwind.F90:141-150): FF_NOW%UFRIC <- ufric_next (DEVICE,
 * (NPROMA, NCHNK)), first-guess TAUW = UFRIC**2 (1 - (ALPHA/CHRNCK)**2), 0 below USTMIN_RESET_TAUW; next->wswave is not read (may be
 * NULL).  IMPLSCH then derives Z0 and U10 with Z0WAVE in the first SINFLX call (airsea.F90:102-120) -- a branch the reference's own
 * GPU build removes (`!$loki remove`); the forcing readers WAMWND / GETWND stay ICODE_WND = 3 only.                                  */
int ecwam_b200_newwind_ustar(ecwam_b200_handle h, const ecwam_b200_forcing_next* next, const double* ufric_next);

/* GETWND's blocking step (src/ecwam/getwnd.F90:196-212): WAMWND (wamwnd.F90:120-300, ICODE_WND = 3) + MICEP
 * (micep.F90:84-240, uncoupled) turn the forcing fields on the forcing grid into the FF_NEXT members that NEWWIND
 * copies.  FIELDG arrays: DEVICE, (NXS:NXE, NYS:NYE) first index fastest; wswave / wdwave are read only with
 * LLWSWAVE / LLWDWAVE.  ifromij / jfromij: DEVICE (NPROMA,NCHNK) = BLK2LOC%IFROMIJ / JFROMIJ.  The relative-wind
 * correction (LRELWIND with IREFRA = 2, 3) reads the bound UCUR / VCUR.  Not built: ICODE_WND = 1, 2; the coupled
 * (LWCOU, NEMO) branches; the swamp domain.                                                                   */
typedef struct ecwam_b200_fieldg {
  const double* uwnd;
  const double* vwnd;
  const double* aird;
  const double* wstar;
  const double* cicover;
  const double* cithick;
  const double* ustra;
  const double* vstra;
  const double* wswave;
  const double* wdwave;
} ecwam_b200_fieldg;
typedef struct ecwam_b200_getwnd_opts {
  int nxs;
  int nxe;
  int nys;
  int nye;
  int llwswave;    /* YOWWIND LLWSWAVE */
  int llwdwave;    /* YOWWIND LLWDWAVE */
  int lrelwind;    /* YOWCURR LRELWIND */
  int iparamci;    /* YOWICE IPARAMCI: 31 sea-ice fraction, 139 sea-surface temperature */
  int liceth;      /* YOWICE LICETH: FIELDG%CITHICK holds a thickness */
  double zmiss;    /* YOWPCONS ZMISS */
} ecwam_b200_getwnd_opts;
/* next: DEVICE (NPROMA,NCHNK) arrays that are WRITTEN (the same struct ecwam_b200_newwind reads afterwards) */
int ecwam_b200_getwnd(ecwam_b200_handle h, const ecwam_b200_fieldg* fieldg, const ecwam_b200_getwnd_opts* opts,
                      const int* ifromij, const int* jfromij, const ecwam_b200_forcing_next* next);

/* Selection of OUTBLOCK's output columns = YOWCOUT after MPCRTBL (src/ecwam/mpcrtbl.F90:89-460).  Host pointers.  */
typedef struct ecwam_b200_outsel {
  int niprmout;              /* YOWCOUT NIPRMOUT: number of BOUT columns                                     */
  const int* itg;            /* (NIPRMOUT) reference parameter number of each column (the inverse of ITOBOUT) */
  const int* icemask;        /* (NIPRMOUT) IPRMINFO(itg,6): sea-ice mask imposed                              */
  const int* seamask;        /* (NIPRMOUT) IPRMINFO(itg,7): too shallow points set to missing                 */
  int llsource;              /* YOWSTAT LLSOURCE                                                              */
  double zmiss;              /* YOWPCONS ZMISS                                                                */
} ecwam_b200_outsel;
/* 1 if OUTBLOCK parameter `itg` is built: 1-16, 20-28, 32, 35-41, 52-56, 62-69, 73-77 (numbering of
 * mpcrtbl.F90 with NTRAIN = 3; 9 = MEANSQS with XKMSS_CUTOFF = XK_GC(NWAV_GC), userin.F90:1213-1215, needs the *_gc
 * tables).  Not built: altimeter (17-19), KURTOSIS family (29-31, 33, 34, 57,
 * 70-72), swell partitions (42-50, LLPARTITION), CIMSSTRN (51), NEMO fields (58-61), W_MAXH (78-81), 82+.     */
int ecwam_b200_outparam_supported(int itg);
/* OUTBS (src/ecwam/outbs.F90:97-122): OUTBLOCK over all chunks (src/ecwam/outblock.F90:150-610, LSECONDORDER = F; with
 * IREFRA = 2, 3 the output spectrum is INTPOL's, intpol.F90:96-271, built in the handle's scratch) with FEMEAN, STHQ, DOMINANT_PERIOD, SEPWISW, MWP1, MWP2, WDIRSPREAD, OUTBETA, WEFLUX and
 * OUTSETWMASK.  bout: DEVICE (NPROMA, NIPRMOUT, NCHNK); iodp: DEVICE (NPROMA, NCHNK) WVENVI%IODP or NULL (= 1).
 * Reads the bound fields (FL1, XLLWS, CINV, CGROUP, forcing, IMPLSCH outputs).                                */
int ecwam_b200_outbs(ecwam_b200_handle h, const ecwam_b200_outsel* sel, const int* iodp, double* bout);
/* OUTWNORM -> MPMINMAXAVG (src/ecwam/mpminmaxavg.F90:68-195) on a BOUT made by ecwam_b200_outbs.
 * wnorm: HOST (4, NIPRMOUT) = average, minimum, maximum, number of non-missing points (every rank gets it).
 * llglobal = 0: per-rank sums combined over ranks (MPL_ALLREDUCE branch, :160-191);
 * llglobal = 1: LLNORMWAMOUT_GLOBAL, one sequential sum over the ORIGINAL global point order (:121-153), bit-
 *   reproducible for any number of ranks; needs niblo and, when nproc > 1, ij2newij (HOST, (0:NIBLO), mpdecomp.F90:667-686)
 *   and nstart/nend (HOST, (NPROC)).                                                                          */
int ecwam_b200_outwnorm(ecwam_b200_handle h, const ecwam_b200_outsel* sel, const double* bout, int llglobal, int niblo,
                        const int* ij2newij, const int* nstart, const int* nend, double* wnorm);

int ecwam_b200_synchronize(ecwam_b200_handle h);
/* Kernel launch counter (all kernels launched through this handle since creation). */
long long ecwam_b200_launch_count(ecwam_b200_handle h);
/* Average device time [ms] of the named kernel class since the last reset ("propags2", "implsch_main", ...),
 * measured with CUDA events on the handle's stream when timing is enabled.                             */
int ecwam_b200_timing_enable(ecwam_b200_handle h, int on);
int ecwam_b200_timing_get(ecwam_b200_handle h, const char* name, double* total_ms, long long* count);
int ecwam_b200_timing_reset(ecwam_b200_handle h);
const char* ecwam_b200_last_error(void);
int ecwam_b200_version(void);
/* Roofline denominators measured on the current device (bench.py): FP64 FMA rate [TFLOP/s] (16 independent DFMA chains
 * per thread, 8 CTAs of 256 threads per SM) — the bound of the IMPLSCH kernels — and a streaming-copy
 * bandwidth [GB/s, read + write].  Either pointer may be NULL.  Not part of the hot path.                */
int ecwam_b200_measure_peaks(double* fp64_tflops, double* copy_gbs);

/* ---------------------------------------------------------------------------------------------------
 * Host-side builders (C++; one-off, init only) for callers that do not bring ecWAM's module state:
 * the equivalents of MFREDIR/INITMDL/SETWAVPHYS/INISNONLIN/... and of PROPCONNECT/MPDECOMP/MCHUNK.
 * The returned objects own their arrays; *_free releases them.                                         */
typedef struct ecwam_b200_host_tables_s* ecwam_b200_host_tables_t;
int ecwam_b200_host_tables_create(const ecwam_b200_params* params, int ifre1, double fr1,
                                  ecwam_b200_host_tables_t* out);
const ecwam_b200_tables* ecwam_b200_host_tables_get(ecwam_b200_host_tables_t t);
int ecwam_b200_host_tables_free(ecwam_b200_host_tables_t t);

typedef struct ecwam_b200_host_grid_s* ecwam_b200_host_grid_t;
/* Reduced grid + sea mask -> sea-point order, MPDECOMP for nproc ranks, PROPCONNECT, halo lists.
 * nlonrgg (NGY) points per row south->north; mask: one byte per grid cell, row-major south->north.      */
int ecwam_b200_host_grid_create(int ngy, const int* nlonrgg, double amosop, double amonop,
                                const unsigned char* mask, int nproc, int ll1d, ecwam_b200_host_grid_t* out);
int ecwam_b200_host_grid_niblo(ecwam_b200_host_grid_t g);
/* decomp of 1-based rank `irank`; land_cgroup must be supplied by the caller before ecwam_b200_create. */
const ecwam_b200_decomp* ecwam_b200_host_grid_decomp(ecwam_b200_host_grid_t g, int irank);
/* maps between the original global sea-point order (1..NIBLO) and the relabelled order (mpdecomp.F90:667-686) */
const int* ecwam_b200_host_grid_ij2newij(ecwam_b200_host_grid_t g);   /* (0:NIBLO) */
const int* ecwam_b200_host_grid_newij2ij(ecwam_b200_host_grid_t g);   /* (0:NIBLO) */
const int* ecwam_b200_host_grid_kxlt(ecwam_b200_host_grid_t g);       /* (NIBLO) relabelled order */
const int* ecwam_b200_host_grid_nstart(ecwam_b200_host_grid_t g);     /* (NPROC) */
const int* ecwam_b200_host_grid_nend(ecwam_b200_host_grid_t g);       /* (NPROC) */
int ecwam_b200_host_grid_free(ecwam_b200_host_grid_t g);
/* DEPTHPRPT/AKI (depthprpt.F90:60-81, aki.F90:71-91): dispersion fields for n points, arrays (n,NFRE). */
int ecwam_b200_host_depthprpt(const ecwam_b200_tables* t, int nfre, long long n, const double* depth, double* wavnum,
                              double* cinv, double* cgroup, double* xk2cg, double* omosnh2kd, double* stokfac);

/* ---------------------------------------------------------------------------------------------------
 * Restart files and the grid-table file in the reference's own on-disk formats (host only, no GPU):
 * Fortran unformatted sequential records, [int32 n][n bytes][int32 n] little endian, gfortran sub-records
 * above 2147483639 bytes; reals are REAL*8 (the double-precision build).
 *   BLS  SAVSPEC/WRITEFL/READFL (savspec.F90:86-166, writefl.F90:86-120, readfl.F90:118-145):
 *        NFRE*NANG records (frequency outside, direction inside; KDEL = MDEL = 1, yowcout.F90:70-71), each
 *        FL(1:NIBLO) of one (K,M) in the ORIGINAL sea-point order (the file position IJ holds the model's
 *        point IJ2NEWIJ(IJ)).  LRSTPARALW: one file per task, FILENAME.%p_%n, one record (IJSG:IJLG,NANG,NFRE).
 *   LAW  SAVSTRESS/WRITESTRESS/READSTRESS (savstress.F90:80-152, writestress.F90:76-109): one record
 *        CDTPRO,CDATEWO,CDAWIFL,CDATEFL (4 x CHARACTER*14), then NREAL records of NIBLO reals in the order
 *        WSWAVE WDWAVE UFRIC TAUW TAUWDIR Z0M Z0B CHRNCK AIRD WSTAR CICOVER CITHICK USTRA VSTRA UCUR VCUR.
 *   wam_grid_tables  OUTCOM/READPRE (outcom.F90:139-144, readpre.F90:262-345): NKIND,IMDLGRBID_G | NGX,NGY |
 *        NLONRGG(NGY) | IPER,IRGG,AMOWEP,AMOSOP,AMOEAP,AMONOP,XDELLA,XDELLO | BATHY(NGX,NGY).
 * Writers are rank-wise: every rank passes its own points (nown of them, ijorig = their 1-based ORIGINAL
 * indices, NULL = 1..NIBLO) and writes them into their places of the shared file; the rank called with
 * create = 1 sizes the file and writes the record markers and must return before the others start.
 * fl is (nown, NANG, NFRE), rfield (nown, NREAL), point index fastest.                                  */
int ecwam_b200_grstname(const char* cdated, const char* cdatef, int ifcst, const char* fileid, const char* cpad,
                        char* filename, int cap);                                 /* grstname.F90:88-142 */
int ecwam_b200_restart_par_name(const char* filename, int irank, int nproc, char* out, int cap);
int ecwam_b200_savspec(const char* filename, long long niblo, int nang, int nfre, long long nown, const int* ijorig,
                       const double* fl, int create);
int ecwam_b200_getspec(const char* filename, long long niblo, int nang, int nfre, long long nown, const int* ijorig,
                       double* fl);
int ecwam_b200_savspec_par(const char* filename, long long nown, int nang, int nfre, const double* fl);
int ecwam_b200_getspec_par(const char* filename, long long nown, int nang, int nfre, double* fl);
int ecwam_b200_savstress(const char* filename, const char* cdtpro, const char* cdatewo, const char* cdawifl,
                         const char* cdatefl, long long niblo, int nreal, long long nown, const int* ijorig,
                         const double* rfield, int create);
/* dates: 4 x 15 bytes, NUL-terminated CDTPRO, CDATEWO, CDAWIFL, CDATEFL (may be NULL) */
int ecwam_b200_getstress(const char* filename, char* dates, long long niblo, int nreal, long long nown,
                         const int* ijorig, double* rfield);
/* amo = AMOWEP, AMOSOP, AMOEAP, AMONOP, XDELLA, XDELLO; bathy (NGX,NGY), land = -999 */
int ecwam_b200_grid_tables_write(const char* filename, int imdlgrbid_g, int ngx, int ngy, const int* nlonrgg, int iper,
                                 int irgg, const double* amo, const double* bathy);
/* nlonrgg = bathy = NULL: dimensions only */
int ecwam_b200_grid_tables_read(const char* filename, int* nkind, int* kmdlgrdid, int* ngx, int* ngy, int* nlonrgg,
                                int nlon_cap, int* iper, int* irgg, double* amo, double* bathy, long long bathy_cap);
/* test hook: the sub-record limit (default and maximum 2147483639) */
int ecwam_b200_io_set_max_subrecord(long long nbytes);

#ifdef __cplusplus
}
#endif
#endif /* ECWAM_B200_H */
