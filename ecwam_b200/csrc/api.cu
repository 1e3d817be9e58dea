// C ABI of libecwam_b200.so (include/ecwam_b200.h): per-rank handle, device tables, NCCL halo plan,
// PROPAG_WAM / IMPLSCH / WAMINTGR entry points.
#include "internal.h"
#include <nccl.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

static thread_local char g_err[1024] = "";
void ew_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
#define EW_FAIL(code, ...) do { ew_set_error(__VA_ARGS__); return (code); } while (0)
#define EW_FAIL_H(h, code, ...) do { ew_set_error(__VA_ARGS__); delete (h); return (code); } while (0)
#define EW_NCCL_CHECK(call)                                                                  \
  do {                                                                                       \
    ncclResult_t r_ = (call);                                                                \
    if (r_ != ncclSuccess) EW_FAIL(ECWAM_B200_ENCCL, "%s:%d: %s: %s", __FILE__, __LINE__, #call, ncclGetErrorString(r_)); \
  } while (0)

using namespace ew;

namespace {
template <class T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  ~DBuf() { free(); }
  int alloc(size_t cnt) {
    free();
    n = cnt;
    if (cnt == 0) return 0;
    cudaError_t e = cudaMalloc((void**)&p, cnt * sizeof(T));
    if (e != cudaSuccess) { ew_set_error("cudaMalloc(%zu bytes): %s", cnt * sizeof(T), cudaGetErrorString(e)); p = nullptr; return ECWAM_B200_ECUDA; }
    return 0;
  }
  int upload(const std::vector<T>& v, cudaStream_t st) {
    int rc = alloc(v.size());
    if (rc) return rc;
    if (v.empty()) return 0;
    EW_CUDA_CHECK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    EW_CUDA_CHECK(cudaStreamSynchronize(st));
    return 0;
  }
  void free() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

struct TimingClass { double total_ms = 0; long long count = 0; std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending; };
}  // namespace

struct ecwam_b200_handle_s {
  ecwam_b200_params par;
  DevConst dc;
  PropConst pc;
  cudaStream_t st = nullptr;
  ncclComm_t comm = nullptr;
  int nproc = 1, irank0 = 0;
  // propagation
  PropDev pd;
  DBuf<int> nbr, halo_off, halo_str, send_l, send_pre, send_peer_of, recv_pre, recv_peer_of, recv_e, flag, count;
  DBuf<double> wl, pt, cgext, halo, sendbuf, fl3, cosph_m, cosph_p, land_cg, cgrecv, obs;
  DBuf<double> wlat_raw, dellam, grad, curmask;   // IREFRA /= 0: WLAT as PROPCONNECT left it, DELLAM/COSPH(KXLT), gradients, CURMASK
  double oneo2delphi = 0.0;
  std::vector<int> h_spre, h_rpre;   // per peer prefix sums (size nproc+1)
  // MPEXCHNG through the host's own message passing instead of NCCL (ecwam_b200_set_exchange)
  ecwam_b200_exchange_fn xchg_fn = nullptr;
  void* xchg_user = nullptr;
  bool xchg_staged = false;
  double *xchg_hs = nullptr, *xchg_hr = nullptr;   // pinned staging of the send / receive buffers (staged mode)
  size_t xchg_ns = 0, xchg_nr = 0;
  // halo / compute overlap of PROPAG_WAM: own points [int_lo, int_hi) have no halo neighbour and are propagated while the
  // exchange is in flight on st_x; the strips below and above follow it
  int int_lo = 0, int_hi = 0;
  bool overlap = false;
  cudaStream_t st_x = nullptr;
  std::vector<cudaStream_t> st_part;   // IMPLSCH in parts on streams of their own (implsch_range)
  std::vector<cudaEvent_t> ev_part;
  cudaEvent_t ev_fork = nullptr;
  cudaEvent_t ev_pack = nullptr, ev_x = nullptr;
  bool send_chunks_sorted = false;
  cudaEvent_t ev_halo = nullptr;
  std::vector<int> send_chunks;      // sorted chunks (0-based) that hold at least one point some peer needs (host-buffer pipeline)
  int nsend = 0, nrecv = 0;
  int msplit = 0;
  bool weights_dirty = true;
  // implsch
  DBuf<double> scr, satw, swellft, fldin, tbg, gctab;
  bool have_gc = false, mss_ok = false;     // gravity-capillary tables supplied; DELKCC_GC too (MEANSQS)
  int gc_n = 0;
  double gc_sqrtgosurft = 0, gc_xk1 = 0, gc_xkn = 0, alphapmax = 0, fratio = 1.1;
  int dsh[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  int halo_r = 0, halo_c = 0, nsdsnth = 0;
  DBuf<int> kw, isat;
  DevTabPtr tab;
  // fields
  ecwam_b200_fields dev;
  bool bound = false;
  // host-call mirrors
  ecwam_b200_fields mir;
  bool mir_alloc = false, mir_static_done = false;
  // h->dev is the caller's device binding except inside ecwam_b200_wamintgr_host, which swaps the mirrors in for the call.
  // The CTU set-up is built from the static fields of one side: cur_side / weights_side = 1 caller's tensors, 2 mirrors.
  int cur_side = 0, weights_side = 0;
  std::vector<void*> mir_bufs;
  // banded copy/compute pipeline of the host-buffer entry point
  cudaStream_t st_up = nullptr, st_dn = nullptr;
  std::vector<cudaEvent_t> ev_up, ev_done;
  cudaEvent_t ev_start = nullptr;
  int nbr_reach = 0;       // max |l' - l| over the own-point neighbours of every own point l
  // resident-state step (ecwam_b200_wamintgr_forced): device staging of the eight FF_NEXT fields
  DBuf<double> frc_next;
  DBuf<double> enhp;       // ENH(IJ,MC) plane of ISNONLIN = 1, 2
  ecwam_b200_nemo_fields nemo{};   // LWNEMOCOU: bound with ecwam_b200_bind_nemo
  bool nemo_bound = false;
  DBuf<NemoDev> nemo_dev;          // k_nemo's argument block (allocated when LWNEMOCOU or LWNEMOCOUSTRN)
  DBuf<double> ice1, ice2; // SDICE1's table (+ per-frequency period interpolation), SDICE2's per-(point, frequency) factor (k_ice)
  int ice_nt = 0, ice_nh = 0;
  double ice_hmin = 0.0, ice_dh = 1.0;
  // NEWWIND / OUTBLOCK / WAMNORM
  DBuf<double> normbuf, zglobal;
  DBuf<int> ij2new_d;
  long long ij2new_n = 0;
  // stats
  long long nlaunch = 0;
  bool timing = false;
  std::map<std::string, TimingClass> tm;
};
typedef ecwam_b200_handle_s H;

static H* g_const_owner = nullptr;
static int ensure_const(H* h) {
  if (g_const_owner == h) return 0;
  int rc = upload_dev_const(h->dc, h->st);
  if (rc) return rc;
  rc = upload_prop_const(h->pc, h->st);
  if (rc) return rc;
  rc = upload_prop_const_fast(h->pc, h->st);
  if (rc) return rc;
  g_const_owner = h;
  return 0;
}

namespace {
struct ScopedTimer {
  H* h; TimingClass* tc = nullptr; cudaEvent_t a = nullptr, b = nullptr; cudaStream_t s = nullptr;
  ScopedTimer(H* h_, const char* name, cudaStream_t on = nullptr) : h(h_) {
    if (!h->timing) return;
    s = on ? on : h->st;
    tc = &h->tm[name];
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, s);
  }
  ~ScopedTimer() {
    if (!tc) return;
    cudaEventRecord(b, s);
    tc->pending.push_back({a, b});
  }
};
void drain_timing(H* h) {
  for (auto& kv : h->tm) {
    for (auto& pr : kv.second.pending) {
      cudaEventSynchronize(pr.second);
      float ms = 0;
      cudaEventElapsedTime(&ms, pr.first, pr.second);
      kv.second.total_ms += ms; kv.second.count++;
      cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
    }
    kv.second.pending.clear();
  }
}
inline long nint_l(double x) { return std::lround(x); }
}  // namespace

extern "C" {

const char* ecwam_b200_last_error(void) { return g_err; }
int ecwam_b200_version(void) { return 100; }

int ecwam_b200_nccl_unique_id(char id_out[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  EW_NCCL_CHECK(ncclGetUniqueId(&id));
  memcpy(id_out, &id, 128);
  return 0;
}
int ecwam_b200_nccl_comm_init(const char id[128], int nranks, int rank, void** comm_out) {
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  ncclComm_t c;
  EW_NCCL_CHECK(ncclCommInitRank(&c, nranks, uid, rank));
  *comm_out = (void*)c;
  return 0;
}
int ecwam_b200_nccl_comm_destroy(void* comm) {
  if (comm) EW_NCCL_CHECK(ncclCommDestroy((ncclComm_t)comm));
  return 0;
}

static int fill_dev_const(const ecwam_b200_params& p, const ecwam_b200_tables& t, DevConst& c) {
  memset(&c, 0, sizeof(c));
  if (p.nang < 4 || p.nang > EW_MAXA || p.nfre < 2 || p.nfre > EW_MAXF || p.nfre_red < 1 || p.nfre_red > p.nfre)
    EW_FAIL(ECWAM_B200_EINVAL, "unsupported spectral dimensions NANG=%d NFRE=%d NFRE_RED=%d", p.nang, p.nfre, p.nfre_red);
  if (p.irefra < 0 || p.irefra > 3 || p.icase != 1) EW_FAIL(ECWAM_B200_EINVAL, "IREFRA must be 0..3 and ICASE = 1 (spherical)");
  if (p.isnonlin < 0 || p.isnonlin > 2) EW_FAIL(ECWAM_B200_EINVAL, "ISNONLIN must be 0, 1 or 2");
  if ((p.llgcbz0 || p.llnormagam) && (t.nwav_gc < 2 || !t.xk_gc || !t.omega_gc || !t.cm_gc || !t.c2osqrtvg_gc || !t.xkmsqrtvgoc2_gc ||
                                      !t.om3gmkm_gc || !t.omxkm3_gc || !t.delkcc_gc_ns || !t.delkcc_omxkm3_gc))
    EW_FAIL(ECWAM_B200_EINVAL, "LLGCBZ0 / LLNORMAGAM need the gravity-capillary tables (ecwam_b200_tables: nwav_gc, *_gc)");
  if ((p.lciwa & 1) && (t.nict < 2 || t.nich < 2 || !t.cideac || !(t.dtic > 0.0) || !(t.dhic > 0.0)))
    EW_FAIL(ECWAM_B200_EINVAL, "LCIWA1 (SDICE1) needs the CIDEAC table (ecwam_b200_tables: nict, nich, ticmin, hicmin, dtic, dhic, cideac)");
  if (p.lciwa & ~15) EW_FAIL(ECWAM_B200_EINVAL, "lciwa: unknown bits");
  if (p.icode_wnd < 1 || p.icode_wnd > 3) EW_FAIL(ECWAM_B200_EINVAL, "ICODE_WND must be 1, 2 (friction velocity / stress forcing) or 3 (10 m wind)");

  if (p.iphys != 0 && p.iphys != 1) EW_FAIL(ECWAM_B200_EINVAL, "IPHYS must be 0 or 1");
  if (t.mlsthg > EW_MAXMC || t.mlsthg < 1 || t.mfrstlw > 1) EW_FAIL(ECWAM_B200_EINVAL, "bad MLSTHG/MFRSTLW");
  if (t.jtot_tauhf > EW_MAXJT) EW_FAIL(ECWAM_B200_EINVAL, "JTOT_TAUHF too large");
  if (p.iphys == 1 && (2 * t.nsdsnth + 1 > EW_MAXSAT)) EW_FAIL(ECWAM_B200_EINVAL, "NSDSNTH too large");
  if (p.iphys == 1 && (t.ssdsc3 != 0.0 || t.ssdsc5 != 0.0))
    EW_FAIL(ECWAM_B200_EINVAL, "SDISSIP_ARD cumulative (SSDSC3) / turbulence (SSDSC5) terms are not implemented");
  c.A = p.nang; c.F = p.nfre; c.Fr = p.nfre_red; c.iphys = p.iphys; c.idamping = p.idamping; c.llcapchnk = p.llcapchnk;
  c.lbiwbk = p.lbiwbk; c.licerun = p.licerun; c.lmaskice = p.lmaskice; c.lwamrsetci = p.lwamrsetci; c.lwflux = p.lwflux;
  c.lcflx = (p.lwflux || p.lwfluxout || p.lwnemocou) ? 1 : 0;   // implsch.F90:187
  c.lciwa3 = (p.lciwa & 4) ? 1 : 0; c.lciscal = (p.lciwa & 8) ? 1 : 0; c.zalpfacx = p.zalpfacx;
  c.lciwa1 = (p.lciwa & 1) ? 1 : 0; c.lciwa2 = (p.lciwa & 2) ? 1 : 0; c.lciwa_any = (p.lciwa & 7) ? 1 : 0;
  c.zalpfacb = p.zalpfacb; c.cdicwa = p.cdicwa;
  c.lwnemotauoc = p.lwnemotauoc ? 1 : 0; c.lwnemocoustk = p.lwnemocoustk ? 1 : 0;
  c.nemo_send = ((p.lwnemocousend && p.lwcou) || !p.lwcou) ? 1 : 0;     // stokestrn.F90:76-78
  c.ROWATER = 1.0 / t.rowaterm1;
  c.lwnemocouwrs = p.lwnemocouwrs ? 1 : 0; c.lwnemocouibr = p.lwnemocouibr ? 1 : 0; c.zalpwrs = p.zalpwrs; c.zibrw_thrsh = p.zibrw_thrsh;
  {   // SDICE3, IMODEL = 2: ALP = (2*CDICE*CITH**1.25*FR(M)**4.5)*ALPFAC with CDICE = 0.1274*(ZPI/SQRT(G))**4.5 (sdice3.F90:123-129)
    const double cdice = 0.1274 * std::pow(t.zpi / std::sqrt(t.g), 4.5);
    for (int m = 0; m < p.nfre && m < EW_MAXF; ++m) c.fr45[m] = 2. * cdice * std::pow(t.fr[m], 4.5);
  }
  c.lwvflx_snl = p.lwvflx_snl; c.lwcouast = p.lwcouast;
  c.delt = p.idelt; c.ximp = p.ximp; c.rnu = p.rnu; c.rnum = p.rnum; c.wspmin = p.wspmin; c.cithrsh = p.cithrsh;
  c.cithrsh_tail = p.cithrsh_tail; c.ciblock = p.ciblock; c.flmin = p.flmin; c.bathymax = p.bathymax;
  c.G = t.g; c.GM1 = t.gm1; c.ZPI = t.zpi; c.ZPI4GM1 = t.zpi4gm1; c.ZPI4GM2 = t.zpi4gm2; c.ROWATERM1 = t.rowaterm1;
  c.EPSMIN = t.epsmin; c.EPSUS = t.epsus; c.EPSU10 = t.epsu10; c.ACD = t.acd; c.BCD = t.bcd; c.CDMAX = t.cdmax;
  c.TAUOCMIN = t.tauocmin; c.TAUOCMAX = t.tauocmax; c.PHIEPSMIN = t.phiepsmin; c.PHIEPSMAX = t.phiepsmax;
  c.WSEMEAN_MIN = t.wsemean_min;
  c.FRATIO = t.fratio; c.WETAIL = t.wetail; c.FRTAIL = t.frtail; c.WP1TAIL = t.wp1tail; c.DELTH = t.delth;
  c.FLOGSPRDM1 = t.flogsprdm1; c.NFRE_ODD = t.nfre_odd;
  for (int m = 0; m < p.nfre; ++m) {
    c.FR[m] = t.fr[m]; c.DFIM[m] = t.dfim[m]; c.DFIMOFR[m] = t.dfimofr[m]; c.DFIMFR[m] = t.dfimfr[m];
    c.ZPIFR[m] = t.zpifr[m]; c.FR5[m] = t.fr5[m]; c.COFRM4[m] = t.cofrm4[m]; c.FLMAX[m] = t.flmax[m];
    c.RHOWG_DFIM[m] = t.rhowg_dfim[m]; c.DFIM_SIM[m] = t.dfim_sim[m];
  }
  for (int k = 0; k < p.nang; ++k) { c.TH[k] = t.th[k]; c.COSTH[k] = t.costh[k]; c.SINTH[k] = t.sinth[k]; }
  c.XKAPPA = t.xkappa; c.XNLEV = t.xnlev; c.ALPHA = t.alpha; c.ALPHAMIN = t.alphamin; c.CHNKMIN_U = t.chnkmin_u;
  c.ZALP = t.zalp; c.BETAMAXOXKAPPA2 = t.betamaxoxkappa2; c.TAUWSHELTER = t.tauwshelter; c.TAILFACTOR = t.tailfactor;
  c.TAILFACTOR_PM = t.tailfactor_pm; c.SWELLF = t.swellf; c.SWELLF2 = t.swellf2; c.SWELLF3 = t.swellf3;
  c.SWELLF4 = t.swellf4; c.SWELLF5 = t.swellf5; c.SWELLF6 = t.swellf6; c.SWELLF7 = t.swellf7; c.SWELLF7M1 = t.swellf7m1;
  c.Z0RAT = t.z0rat; c.Z0TUBMAX = t.z0tubmax; c.ABMIN = t.abmin; c.ABMAX = t.abmax; c.CDIS = t.cdis;
  c.DELTA_SDIS = t.delta_sdis; c.CDISVIS = t.cdisvis; c.SDSBR = t.sdsbr; c.SSDSC2 = t.ssdsc2; c.SSDSC3 = t.ssdsc3;
  c.SSDSC4 = t.ssdsc4; c.SSDSC5 = t.ssdsc5; c.SSDSC6 = t.ssdsc6; c.MICHE = t.miche; c.EGRCRV = t.egrcrv;
  c.AFCRV = t.afcrv; c.BFCRV = t.bfcrv; c.NSDSNTH = t.nsdsnth;
  c.IAB = t.iab; c.JTOT = t.jtot_tauhf; c.EPS1 = t.eps1; c.X0TAUHF = t.x0tauhf;
  c.llgcbz0 = p.llgcbz0 ? 1 : 0; c.llnormagam = p.llnormagam ? 1 : 0;
  if (c.llgcbz0 || c.llnormagam) {
    c.NWAV_GC = t.nwav_gc; c.ALPHAMAX = t.alphamax; c.ALPHAPMAX = t.alphapmax; c.ACDLIN = t.acdlin; c.BCDLIN = t.bcdlin;
    c.BMAXOKAP = t.bmaxokap; c.GAMNCONST = t.gamnconst; c.RN1_RN = t.rn1_rn; c.DTHRN_A = t.dthrn_a; c.DTHRN_U = t.dthrn_u;
    c.ANG_GC_A = t.ang_gc_a; c.ANG_GC_B = t.ang_gc_b; c.ANG_GC_C = t.ang_gc_c; c.SQRTGOSURFT = t.sqrtgosurft;
    c.XKM1_GC = 1.0 / t.xk_gc[0];                 // XKM_GC(1) (initgc.F90:87)
    c.XLOGKRATIOM1_GC = 1.0 / std::log(1.2);      // yowfred.F90:62-63
  }
  for (int j = 0; j < t.jtot_tauhf; ++j) c.WTAUHF[j] = t.wtauhf[j];
  c.MLSTHG = t.mlsthg; c.MFRSTLW = t.mfrstlw; c.KFRH = t.kfrh; c.DAL1 = t.dal1; c.DAL2 = t.dal2;
  const int off = 1 - t.mfrstlw;   // index of MC=1 inside the (MFRSTLW:MLSTHG) arrays
  for (int mc = 0; mc < t.mlsthg; ++mc) {
    c.IKP[mc] = t.ikp[off + mc]; c.IKP1[mc] = t.ikp1[off + mc]; c.IKM[mc] = t.ikm[off + mc]; c.IKM1[mc] = t.ikm1[off + mc];
    c.AF11[mc] = t.af11[off + mc];
    for (int j = 0; j < 5; ++j) { c.INLCOEF[mc][j] = t.inlcoef[j + 5 * mc]; c.NLSLOT[mc][j] = (t.inlcoef[j + 5 * mc] - 1) % 9; }
    for (int j = 0; j < 25; ++j) c.RNLCOEF[mc][j] = t.rnlcoef[j + 25 * mc];
    // The spectrum-edge cases of SNONLIN (snonlin.F90:310-498: rows of the quadruplet outside 1..NFRE receive nothing) are folded
    // into the coefficients: a skipped update has zero weights, so k_stencil's sweep is branch-free.
    {
      const int MC = mc + 1, F = p.nfre;
      const int MFR1STFR = -t.mfrstlw + 1, MFRLSTFR = F - t.kfrh + MFR1STFR;
      const int MP = MC + 2, MP1 = MC + 3, MM1 = MC - 3;
      bool do_c, do_mm, do_mm1, do_mp, do_mp1;
      if (MC > MFR1STFR && MC < MFRLSTFR) { do_c = do_mm = do_mm1 = do_mp = do_mp1 = true; c.RNLCOEF[mc][0] = 1.0; /* FTAIL = 1 there */ }
      else if (MC >= MFRLSTFR) { do_mm = true; do_mm1 = MM1 <= F; do_c = do_mm1 && MC <= F; do_mp = do_c && MP <= F; do_mp1 = do_mp && MP1 <= F; }
      else { do_mm = false; do_mm1 = MM1 >= 1; do_c = true; do_mp = true; do_mp1 = true; }
      double* R = c.RNLCOEF[mc];
      if (!do_mm) R[19] = R[20] = R[23] = R[24] = 0.0;     // FKLAMM2, FKLAMM1, FKLAM12, FKLAM22 -> row MC-4
      if (!do_mm1) R[17] = R[18] = R[21] = R[22] = 0.0;    // FKLAMMA, FKLAMMB, FKLAMA2, FKLAMB2 -> row MC-3
      if (!do_mp) R[7] = R[8] = R[11] = R[12] = 0.0;       // FKLAMP2, FKLAMP1, FKLAP12, FKLAP22 -> row MC+2
      if (!do_mp1) R[5] = R[6] = R[9] = R[10] = 0.0;       // FKLAMPA, FKLAMPB, FKLAPA2, FKLAPB2 -> row MC+3
      c.RNLC2[mc] = do_c ? 2.0 : 0.0;                      // weight of the centre bin (-2 AD, -2 DELAD)
    }
  }
  for (int r = 0; r < EW_MAXF + 8; ++r) c.SLOT9[r] = r % 9;
  // separable form of the DIA weights for k_sweep (see DevConst::NLW); every identity it relies on is checked here
  {
    c.sweep_ok = 0;
    const int NM = t.mlsthg;
    int mi = -1;   // an interior centre frequency (all five rows inside the spectrum)
    for (int mc = 0; mc < NM; ++mc) { const double* R = c.RNLCOEF[mc]; if (R[5] > 0 && R[6] >= 0 && R[17] > 0 && R[18] >= 0 && R[8] > 0 && R[20] > 0) { mi = mc; break; } }
    if (mi >= 0 && NM + 1 <= EW_MAXMC + 1) {
      const double* Ri = c.RNLCOEF[mi];
      const double cl11 = Ri[5] / (Ri[5] + Ri[6]), acl1 = Ri[6] / (Ri[5] + Ri[6]);
      const double cl21 = Ri[17] / (Ri[17] + Ri[18]), acl2 = Ri[18] / (Ri[17] + Ri[18]);
      c.NLD[0] = cl11; c.NLD[1] = acl1; c.NLD[2] = cl21; c.NLD[3] = acl2;
      c.NLD[4] = cl11 * cl11; c.NLD[5] = acl1 * acl1; c.NLD[6] = cl21 * cl21; c.NLD[7] = acl2 * acl2;
      bool ok = true;
      auto same = [&](double a, double b, double scale) { return std::fabs(a - b) <= 1e-13 * std::fabs(scale) + 1e-300; };
      for (int mc = 0; mc < NM; ++mc) {
        const double* R = c.RNLCOEF[mc];
        double* Wt = c.NLW[mc];
        Wt[0] = R[1] + R[2]; Wt[1] = R[3] + R[4]; Wt[2] = R[13] + R[14]; Wt[3] = R[15] + R[16];
        Wt[4] = R[7] + R[8]; Wt[5] = R[5] + R[6]; Wt[6] = R[19] + R[20]; Wt[7] = R[17] + R[18];
        for (int j = 0; j < 4; ++j) Wt[8 + j] = Wt[4 + j] * Wt[4 + j];
        ok = ok && same(Wt[0] * cl11, R[1], Wt[0]) && same(Wt[0] * acl1, R[2], Wt[0]) && same(Wt[1] * cl11, R[3], Wt[1]) && same(Wt[1] * acl1, R[4], Wt[1]);
        ok = ok && same(Wt[2] * cl21, R[13], Wt[2]) && same(Wt[2] * acl2, R[14], Wt[2]) && same(Wt[3] * cl21, R[15], Wt[3]) && same(Wt[3] * acl2, R[16], Wt[3]);
        ok = ok && same(Wt[4] * cl11, R[8], Wt[4]) && same(Wt[4] * acl1, R[7], Wt[4]) && same(Wt[5] * cl11, R[5], Wt[5]) && same(Wt[5] * acl1, R[6], Wt[5]);
        ok = ok && same(Wt[6] * cl21, R[20], Wt[6]) && same(Wt[6] * acl2, R[19], Wt[6]) && same(Wt[7] * cl21, R[17], Wt[7]) && same(Wt[7] * acl2, R[18], Wt[7]);
        ok = ok && same(Wt[8] * c.NLD[4], R[11], Wt[8]) && same(Wt[8] * c.NLD[5], R[12], Wt[8]) && same(Wt[9] * c.NLD[4], R[9], Wt[9]) && same(Wt[9] * c.NLD[5], R[10], Wt[9]);
        ok = ok && same(Wt[10] * c.NLD[6], R[23], Wt[10]) && same(Wt[10] * c.NLD[7], R[24], Wt[10]) && same(Wt[11] * c.NLD[6], R[21], Wt[11]) && same(Wt[11] * c.NLD[7], R[22], Wt[11]);
        // the row that is IP1 (IM1) of one centre frequency is IP (IM) of the next: k_sweep carries its direction interpolation over
        if (mc + 1 < NM) ok = ok && c.INLCOEF[mc + 1][1] == c.INLCOEF[mc][2] && c.INLCOEF[mc + 1][3] == c.INLCOEF[mc][4];
        c.NLS2[mc + 1][0] = c.NLSLOT[mc][2]; c.NLS2[mc + 1][1] = c.NLSLOT[mc][4];
      }
      c.NLS2[0][0] = c.NLSLOT[0][1]; c.NLS2[0][1] = c.NLSLOT[0][3];
      c.sweep_ok = ok ? 1 : 0;
    }
  }
  if (p.iphys == 1) {
    // cos(TH(K)-TH(J))**ISB only depends on K-J (init_sdiss_ardh.F90:69-96): the table rows agree to rounding (checked), so
    // k_stencil takes the weights of the middle direction as warp-uniform constants
    const int ns = 2 * t.nsdsnth + 1, kmid = p.nang / 2;
    for (int x = 0; x < ns; ++x) {
      const double w = t.satweights[kmid + p.nang * x];
      for (int k = 0; k < p.nang; ++k)
        if (std::fabs(t.satweights[k + p.nang * x] - w) > 1e-12 * std::fabs(t.satweights[kmid + p.nang * (ns / 2)]))
          EW_FAIL(ECWAM_B200_EINVAL, "SATWEIGHTS depends on the direction (x=%d, k=%d): unsupported", x, k);
      c.SATW1[x] = w;
    }
  }
  return 0;
}

static int upload_nemo_dev(struct ecwam_b200_handle_s* h);

int ecwam_b200_create(const ecwam_b200_params* params, const ecwam_b200_tables* tables, const ecwam_b200_decomp* dec,
                      void* nccl_comm, void* cuda_stream, ecwam_b200_handle* out) {
  if (!params || !tables || !dec || !out) EW_FAIL(ECWAM_B200_EINVAL, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    EW_FAIL(ECWAM_B200_ECUDA, "no CUDA device: the WAMINTGR hot path has no CPU fallback");
  H* h = new H();
  h->par = *params;
  h->st = (cudaStream_t)cuda_stream;
  h->comm = (ncclComm_t)nccl_comm;
  h->nproc = dec->nproc; h->irank0 = dec->irank - 1;
  // nproc > 1 without a communicator: the halo exchange must be supplied with ecwam_b200_set_exchange before the first step
  int rc = fill_dev_const(*params, *tables, h->dc);
  if (rc) { delete h; return rc; }
  const ecwam_b200_params& p = h->par;
  const int A = p.nang, F = p.nfre, Fr = p.nfre_red, P = p.nproma;
  const int nloc = dec->ijl - dec->ijs + 1;
  const int nbot = dec->ijs - dec->ninf, ntop = dec->nsup - dec->ijl;
  const int next = nbot + nloc + ntop + 1;
  if (nloc <= 0 || P <= 0 || (long long)P * p.nchnk < nloc || (long long)P * (p.nchnk - 1) >= nloc) {
    EW_FAIL_H(h, ECWAM_B200_EINVAL, "NPROMA=%d NCHNK=%d do not cover %d own points (mchunk.F90:64-75)", P, p.nchnk, nloc);
  }
  // ---- direction tables (ctuwupdt.F90:111-161, ctuw.F90:404-431)
  PropConst& pc = h->pc;
  memset(&pc, 0, sizeof(pc));
  for (int k = 0; k < A; ++k) {
    const double cs = tables->costh[k], sn = tables->sinth[k];
    pc.quad[k] = (cs >= 0.0 ? 0 : 2) + (sn >= 0.0 ? 0 : 1);
    pc.kpm_m[k] = (k - 1 < 0) ? A - 1 : k - 1;
    pc.kpm_p[k] = (k + 1 >= A) ? 0 : k + 1;
    pc.sinth[k] = sn; pc.costh[k] = cs;
  }
  {  // the quadrants must be contiguous direction ranges in the order quad = 0, 2, 3, 1 (TH(K) increasing from DELTH/2)
    static const int order[4] = {0, 2, 3, 1};
    int k = 0;
    for (int j = 0; j < 4; ++j) { pc.kq[j] = k; while (k < A && pc.quad[k] == order[j]) ++k; }
    pc.kq[4] = k;
    if (k != A) EW_FAIL_H(h, ECWAM_B200_EINVAL, "directions are not ordered by compass quadrant (TH must increase from DELTH/2)");
  }
  pc.delpro[0] = p.delpro_lf; pc.delpro[1] = p.idelpro;
  for (int v = 0; v < 2; ++v) {
    const double delth0 = 0.25 * pc.delpro[v] / tables->delth;
    pc.delth0[v] = delth0;
    pc.delfr0[v] = 0.25 * pc.delpro[v] / ((tables->fratio - 1) * tables->zpi);
    for (int k = 0; k < A; ++k) {
      pc.sp[v][k] = delth0 * (tables->sinth[k] + tables->sinth[pc.kpm_p[k]]) / tables->r_earth;
      pc.sm[v][k] = delth0 * (tables->sinth[k] + tables->sinth[pc.kpm_m[k]]) / tables->r_earth;
    }
  }
  pc.cmtodeg = 360.0 / tables->circ;
  pc.fratio = tables->fratio;
  for (int m = 0; m < p.nfre && m < EW_MAXF; ++m) pc.fr[m] = tables->fr[m];
  pc.xdella = dec->xdella;
  h->msplit = (p.ifrelfmax > 0) ? std::min(p.ifrelfmax, Fr) : 0;   // ctuwupdt.F90:193-235

  // ---- per-point tables
  const int NLAND = dec->nsup + 1;
  auto ext = [&](int ij) { return ij - dec->ninf; };
  std::vector<int> nbr((size_t)14 * nloc);
  std::vector<double> wl((size_t)6 * nloc), pt((size_t)5 * nloc, 0.0), cpm(nloc), cpp(nloc);
  for (int l = 0; l < nloc; ++l) {
    int klat[2][2], kcor[4][2];
    for (int ic = 0; ic < 2; ++ic) nbr[(size_t)ic * nloc + l] = ext(dec->klon[l + (size_t)nloc * ic]);
    for (int icl = 0; icl < 2; ++icl)
      for (int ic = 0; ic < 2; ++ic) {
        klat[ic][icl] = dec->klat[l + (size_t)nloc * (ic + 2 * icl)];
        nbr[(size_t)(2 + ic + 2 * icl) * nloc + l] = ext(klat[ic][icl]);
      }
    for (int icl = 0; icl < 2; ++icl)
      for (int icr = 0; icr < 4; ++icr) {
        kcor[icr][icl] = dec->kcor[l + (size_t)nloc * (icr + 4 * icl)];
        nbr[(size_t)(6 + icr + 4 * icl) * nloc + l] = ext(kcor[icr][icl]);
      }
    // CTUWINI's edit of WLAT/WCOR next to land (ctuwini.F90:61-99)
    for (int ic = 0; ic < 2; ++ic) {
      double w = dec->wlat[l + (size_t)nloc * ic];
      if (klat[ic][0] < NLAND && klat[ic][1] < NLAND) {}
      else if (klat[ic][0] == NLAND) { if (w <= 0.75) w = 0.0; }
      else { if (w >= 0.5) w = 1.0; }
      wl[(size_t)ic * nloc + l] = w;
    }
    for (int icr = 0; icr < 4; ++icr) {
      double w = dec->wcor[l + (size_t)nloc * icr];
      if (kcor[icr][0] < NLAND && kcor[icr][1] < NLAND) {}
      else if (kcor[icr][0] == NLAND) { if (w <= 0.75) w = 0.0; }
      else { if (w > 0.5) w = 1.0; }
      wl[(size_t)(2 + icr) * nloc + l] = w;
    }
    const int ky = dec->kxlt[l];   // 1-based row
    if (ky < 1 || ky > dec->ngy) { EW_FAIL_H(h, ECWAM_B200_EINVAL, "KXLT out of range"); }
    const int km = std::max(1, std::min(ky - 1, dec->ngy)), kp = std::max(1, std::min(ky + 1, dec->ngy));
    cpm[l] = dec->cosph[km - 1];
    cpp[l] = dec->cosph[kp - 1];
    pt[(size_t)3 * nloc + l] = dec->zdello[ky - 1];
    pt[(size_t)4 * nloc + l] = dec->sinph[ky - 1] / dec->cosph[ky - 1];
  }
  for (size_t i = 0; i < nbr.size(); ++i)
    if (nbr[i] < 0 || nbr[i] >= next) { EW_FAIL_H(h, ECWAM_B200_EINVAL, "neighbour index outside NINF..NSUP+1"); }

  // ---- halo plan (mpdecomp.F90:966-1176, mpexchng.F90:120-249)
  const int np = h->nproc;
  h->h_spre.assign(np + 1, 0); h->h_rpre.assign(np + 1, 0);
  for (int q = 0; q < np; ++q) {
    h->h_spre[q + 1] = h->h_spre[q] + (np > 1 ? dec->ntope[q] : 0);
    h->h_rpre[q + 1] = h->h_rpre[q] + (np > 1 ? dec->nfrompe[q] : 0);
  }
  h->nsend = h->h_spre[np]; h->nrecv = h->h_rpre[np];
  if (h->nrecv != nbot + ntop) { EW_FAIL_H(h, ECWAM_B200_EINVAL, "sum(NFROMPE)=%d != halo size %d", h->nrecv, nbot + ntop); }
  std::vector<int> send_l(h->nsend), send_peer(h->nsend), recv_e(h->nrecv), recv_peer(h->nrecv);
  std::vector<int> halo_off(nbot + ntop + 1, 0), halo_str(nbot + ntop + 1, 0);
  const size_t halo_elems = (size_t)h->nrecv * A * Fr;
  if (halo_elems + 1 > 0x7fffffffull) { EW_FAIL_H(h, ECWAM_B200_EINVAL, "halo too large"); }
  for (int q = 0; q < np && np > 1; ++q) {
    for (int ih = 0; ih < dec->ntope[q]; ++ih) {
      const int ij = dec->ijtope[ih + (size_t)dec->ntopemax * q];
      if (ij < dec->ijs || ij > dec->ijl) { EW_FAIL_H(h, ECWAM_B200_EINVAL, "IJTOPE outside own range"); }
      send_l[h->h_spre[q] + ih] = ij - dec->ijs;
      h->send_chunks.push_back((ij - dec->ijs) / p.nproma);
      send_peer[h->h_spre[q] + ih] = q;
    }
    const int nr = dec->nfrompe[q];
    for (int ih = 0; ih < nr; ++ih) {
      const int e = ext(dec->nijstart[q] + ih);
      if (e < 0 || e >= next - 1 || (e >= nbot && e < nbot + nloc)) { EW_FAIL_H(h, ECWAM_B200_EINVAL, "NIJSTART outside halo"); }
      recv_e[h->h_rpre[q] + ih] = e;
      recv_peer[h->h_rpre[q] + ih] = q;
      const int hh = (e < nbot) ? e : e - nloc;
      halo_off[hh] = (int)((size_t)h->h_rpre[q] * A * Fr + ih);
      halo_str[hh] = nr;
    }
  }
  halo_off[nbot + ntop] = (int)halo_elems;   // land: one zero element, stride 0
  halo_str[nbot + ntop] = 0;

  {
    int reach = 0, lo = 0, hi = nloc;
    for (int j = 0; j < 14; ++j)
      for (int l = 0; l < nloc; ++l) {
        const int e = nbr[(size_t)j * nloc + l];
        if (e >= nbot && e < nbot + nloc) reach = std::max(reach, std::abs(e - nbot - l));
        else if (e < nbot) lo = std::max(lo, l + 1);            // reads the halo below the own block
        else if (e < next - 1) hi = std::min(hi, l);            // ... above it (next-1 is the land slot)
      }
    h->nbr_reach = reach;
    h->int_lo = lo; h->int_hi = std::max(lo, hi);
    const char* ov = getenv("ECWAM_B200_OVERLAP");
    // worth it when most of the block is interior (MPDECOMP's latitude bands: one or two rows at either end are not)
    h->overlap = np > 1 && !(ov && ov[0] == '0') && (long long)(h->int_hi - h->int_lo) * 4 >= (long long)nloc * 3;
    if (h->overlap) {
      EW_CUDA_CHECK(cudaStreamCreateWithFlags(&h->st_x, cudaStreamNonBlocking));
      EW_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_pack, cudaEventDisableTiming));
      EW_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_x, cudaEventDisableTiming));
    }
  }
  cudaStream_t st = h->st;
  bool ok = true;
  ok = ok && !h->nbr.upload(nbr, st) && !h->wl.upload(wl, st) && !h->pt.upload(pt, st) && !h->cosph_m.upload(cpm, st) &&
       !h->cosph_p.upload(cpp, st) && !h->halo_off.upload(halo_off, st) && !h->halo_str.upload(halo_str, st) &&
       !h->send_l.upload(send_l, st) && !h->send_peer_of.upload(send_peer, st) && !h->send_pre.upload(h->h_spre, st) &&
       !h->recv_pre.upload(h->h_rpre, st) && !h->recv_peer_of.upload(recv_peer, st) && !h->recv_e.upload(recv_e, st);
  const int nenv = Fr + (p.irefra == 0 ? 0 : (p.irefra == 1 ? 1 : 3));   // + DEPTH_EXT (+ U_EXT, V_EXT) rows
  std::vector<double> landcg(nenv, 0.0);
  if (dec->land_cgroup) for (int m = 0; m < Fr; ++m) landcg[m] = dec->land_cgroup[m];
  if (p.irefra != 0) {
    landcg[Fr] = p.bathymax;                                   // DEPTH_EXT(NSUP+1); U_EXT, V_EXT = 0 there (proenvhalo.F90:104-106)
    std::vector<double> wr((size_t)2 * nloc), dl((size_t)2 * nloc);
    for (int l = 0; l < nloc; ++l) {
      wr[l] = dec->wlat[l]; wr[(size_t)nloc + l] = dec->wlat[l + (size_t)nloc];
      dl[l] = dec->zdello[dec->kxlt[l] - 1] * tables->circ / 360.0;   // DELLAM(KX) (readmdlconf.F90:153)
      dl[(size_t)nloc + l] = dec->cosph[dec->kxlt[l] - 1];           // COSPH(KX) (gradi.F90:221)
    }
    h->oneo2delphi = 0.5 / (dec->xdella * tables->circ / 360.0);      // gradi.F90:113, readmdlconf.F90:136
    std::vector<double> ones(nloc, 1.0);
    ok = ok && !h->wlat_raw.upload(wr, st) && !h->dellam.upload(dl, st) && !h->grad.alloc((size_t)7 * nloc) && !h->curmask.upload(ones, st);
  }
  ok = ok && !h->land_cg.upload(landcg, st);
  if (dec->obslon || dec->obslat || dec->obscor) {   // LSUBGRID: [8][Fr][nloc] = OBSLON(:,:,1:2), OBSLAT(:,:,1:2), OBSCOR(:,:,1:4)
    if (!dec->obslon || !dec->obslat || !dec->obscor) { EW_FAIL_H(h, ECWAM_B200_EINVAL, "LSUBGRID: OBSLON, OBSLAT and OBSCOR must be given together"); }
    const size_t pl = (size_t)Fr * nloc;
    std::vector<double> ob(8 * pl);
    for (size_t i = 0; i < 2 * pl; ++i) { ob[i] = dec->obslon[i]; ob[2 * pl + i] = dec->obslat[i]; }
    for (size_t i = 0; i < 4 * pl; ++i) ob[4 * pl + i] = dec->obscor[i];
    for (size_t i = 0; i < 8 * pl; ++i)
      if (!(ob[i] >= 0.0 && ob[i] <= 1.0)) { EW_FAIL_H(h, ECWAM_B200_EINVAL, "LSUBGRID: obstruction coefficient outside [0, 1]"); }
    ok = ok && !h->obs.upload(ob, st);
  }
  ok = ok && !h->cgext.alloc((size_t)nenv * next) && !h->halo.alloc(halo_elems + 1) &&
       !h->sendbuf.alloc((size_t)h->nsend * A * Fr) && !h->cgrecv.alloc((size_t)h->nrecv * nenv) &&
       !h->fl3.alloc((size_t)P * A * Fr * p.nchnk) && !h->flag.alloc(nloc) && !h->count.alloc(1);
  // IMPLSCH tables
  std::vector<int> kw((size_t)16 * A, -1);
  for (int kh = 0; kh < 2; ++kh)
    for (int k = 0; k < A; ++k) {
      kw[(size_t)(0 + kh) * A + k] = tables->k1w[k + A * kh] - 1;
      kw[(size_t)(2 + kh) * A + k] = tables->k2w[k + A * kh] - 1;
      kw[(size_t)(4 + kh) * A + k] = tables->k11w[k + A * kh] - 1;
      kw[(size_t)(6 + kh) * A + k] = tables->k21w[k + A * kh] - 1;
    }
  for (size_t i = 0; i < (size_t)8 * A; ++i) if (kw[i] < 0 || kw[i] >= A) { ok = false; ew_set_error("K1W/K2W/K11W/K21W out of range"); }
  // inverse direction maps for the gather form of SNONLIN: each of K1W..K21W(.,KH) is a cyclic shift (jafu.F90)
  for (int tb = 0; tb < 4 && ok; ++tb)
    for (int kh = 0; kh < 2; ++kh) {
      for (int k = 0; k < A; ++k) kw[(size_t)(8 + 2 * tb + kh) * A + kw[(size_t)(2 * tb + kh) * A + k]] = k;
      for (int k = 0; k < A; ++k) if (kw[(size_t)(8 + 2 * tb + kh) * A + k] < 0) { ok = false; ew_set_error("K1W/K2W/K11W/K21W is not a permutation"); }
    }
  // the frequency sweep of k_stencil relies on the DIA offsets of FRATIO=1.1, lambda=0.25 (nlweigt.F90:94-103)
  for (int mc = 1; mc <= tables->mlsthg && ok; ++mc) {
    const int o = mc - tables->mfrstlw;
    if (tables->ikp[o] != mc + 2 || tables->ikp1[o] != mc + 3 || tables->ikm[o] != mc - 4 || tables->ikm1[o] != mc - 3 ||
        tables->mlsthg != F + 4) { ok = false; ew_set_error("unsupported DIA frequency offsets (need IKP=M+2, IKM=M-4, MLSTHG=NFRE+4)"); }
  }
  // K1W, K11W, K2W, K21W(.,KH) as signed cyclic shifts (jafu.F90): k_stencil addresses partners as own bin + shift
  {
    static const int order[4] = {0, 2, 1, 3};   // kw blocks are K1W, K2W, K11W, K21W
    int maxsh = 0;
    for (int kh = 0; kh < 2 && ok; ++kh)
      for (int q = 0; q < 4 && ok; ++q) {
        const int* tbl = &kw[(size_t)(2 * order[q] + kh) * A];
        int sh = tbl[0];
        if (sh > A / 2) sh -= A;
        for (int k = 0; k < A; ++k) if (tbl[k] != ((k + sh) % A + A) % A) { ok = false; ew_set_error("K1W/K2W/K11W/K21W is not a cyclic shift"); }
        h->dsh[kh][q] = sh;
        maxsh = std::max(maxsh, std::abs(sh));
      }
    h->halo_c = maxsh;
    h->nsdsnth = tables->nsdsnth;
    h->halo_r = std::max(maxsh, p.iphys == 1 ? tables->nsdsnth : 0);
    if (ok && (2 * h->halo_r > A || (p.iphys == 1 && 2 * tables->nsdsnth + 1 > 17))) { ok = false; ew_set_error("direction halo %d too wide for NANG=%d", h->halo_r, A); }
  }
  ok = ok && !h->kw.upload(kw, st);
  if (p.iphys == 1) {
    const int ns = 2 * tables->nsdsnth + 1;
    std::vector<int> isat((size_t)ns * A);
    std::vector<double> satw((size_t)ns * A);
    for (int j = 0; j < ns; ++j)
      for (int k = 0; k < A; ++k) { isat[(size_t)j * A + k] = tables->indicessat[k + A * j] - 1; satw[(size_t)j * A + k] = tables->satweights[k + A * j]; }
    ok = ok && !h->isat.upload(isat, st) && !h->satw.upload(satw, st);
  }
  std::vector<double> sw(tables->swellft, tables->swellft + tables->iab);
  ok = ok && !h->swellft.upload(sw, st);
  h->have_gc = tables->nwav_gc >= 2 && tables->xk_gc && tables->omega_gc && tables->cm_gc && tables->c2osqrtvg_gc &&
               tables->xkmsqrtvgoc2_gc && tables->om3gmkm_gc && tables->omxkm3_gc && tables->delkcc_gc_ns && tables->delkcc_omxkm3_gc;
  if (h->have_gc) {   // [GC_NT][NWAV_GC]; the last row (DELKCC_GC) is only read by MEANSQS_GC
    const int ng = tables->nwav_gc;
    const double* src[GC_NT] = {tables->xk_gc, tables->omega_gc, tables->cm_gc, tables->c2osqrtvg_gc, tables->xkmsqrtvgoc2_gc,
                                tables->om3gmkm_gc, tables->omxkm3_gc, tables->delkcc_gc_ns, tables->delkcc_omxkm3_gc, tables->delkcc_gc,
                                nullptr};
    std::vector<double> gc((size_t)GC_NT * ng, 0.0);
    for (int r = 0; r < GC_NT; ++r) if (src[r]) for (int i = 0; i < ng; ++i) gc[(size_t)r * ng + i] = src[r][i];
    for (int i = 0; i < ng; ++i) gc[(size_t)GC_LXK * ng + i] = std::log(tables->xk_gc[i]);
    ok = ok && !h->gctab.upload(gc, st);
    h->mss_ok = tables->delkcc_gc != nullptr;
    h->gc_n = ng; h->gc_sqrtgosurft = tables->sqrtgosurft; h->gc_xk1 = tables->xk_gc[0]; h->gc_xkn = tables->xk_gc[ng - 1];
    h->alphapmax = tables->alphapmax; h->fratio = tables->fratio;
  }
  const long long npts = (long long)P * p.nchnk;
  ok = ok && !h->scr.alloc(implsch_scratch_doubles(npts)) && !h->fldin.alloc((size_t)npts * A * F) &&
       !h->tbg.alloc((size_t)EW_TQ_N * F * npts);
  if (p.isnonlin != 0) ok = ok && !h->enhp.alloc((size_t)tables->mlsthg * npts);
  if (p.licerun && (p.lciwa & 1)) {   // SDICE1: the table and, per frequency, the wave-period interpolation (sdice1.F90:144-151)
    const int NT = tables->nict, NH = tables->nich;
    std::vector<double> ice((size_t)NT * NH + 3 * (size_t)F);
    for (size_t i = 0; i < (size_t)NT * NH; ++i) ice[i] = tables->cideac[i];
    for (int m = 0; m < F; ++m) {
      const double tw = 1.0 / tables->fr[m];
      int it = (int)std::floor((tw - tables->ticmin) / tables->dtic + 1);
      it = std::max(1, std::min(it, NT));
      const int it1 = std::max(1, std::min(it + 1, NT));
      const double wt1 = std::max(std::min(1.0, (tw - (tables->ticmin + (it - 1) * tables->dtic)) / tables->dtic), 0.0);
      ice[(size_t)NT * NH + m] = wt1; ice[(size_t)NT * NH + F + m] = it - 1; ice[(size_t)NT * NH + 2 * F + m] = it1 - 1;
    }
    ok = ok && !h->ice1.upload(ice, st);
    h->ice_nt = NT; h->ice_nh = NH; h->ice_hmin = tables->hicmin; h->ice_dh = tables->dhic;
  }
  if (p.licerun && (p.lciwa & 2)) ok = ok && !h->ice2.alloc((size_t)F * npts);
  ok = ok && !upload_nemo_dev(h);   // LWNEMOCOUSTRN without LWNEMOCOU: CIMSSTRN alone
  if (!ok) { ecwam_b200_destroy(h); return ECWAM_B200_ECUDA; }
  cudaMemsetAsync(h->halo.p, 0, (halo_elems + 1) * sizeof(double), st);
  cudaMemsetAsync(h->fl3.p, 0, h->fl3.n * sizeof(double), st);
  cudaMemsetAsync(h->cgext.p, 0, h->cgext.n * sizeof(double), st);
  cudaMemsetAsync(h->scr.p, 0, h->scr.n * sizeof(double), st);

  PropDev& d = h->pd;
  d.nloc = nloc; d.nbot = nbot; d.ntop = ntop; d.next = next; d.P = P; d.A = A; d.F = F; d.Fr = Fr; d.nchnk = p.nchnk;
  d.nbr = h->nbr.p; d.wl = h->wl.p; d.pt = h->pt.p; d.cgext = h->cgext.p; d.halo_off = h->halo_off.p;
  d.irefra = p.irefra; d.nenv = nenv; d.omos = nullptr; d.grad = h->grad.p; d.wavn = nullptr; d.curmask = h->curmask.p;
  d.halo_str = h->halo_str.p; d.halo = h->halo.p; d.obs = h->obs.p; d.pad_obs = nullptr;
  h->tab.k1w = h->kw.p; h->tab.k2w = h->kw.p + 2 * A; h->tab.k11w = h->kw.p + 4 * A; h->tab.k21w = h->kw.p + 6 * A;
  h->tab.ik1w = h->kw.p + 8 * A; h->tab.ik2w = h->kw.p + 10 * A; h->tab.ik11w = h->kw.p + 12 * A; h->tab.ik21w = h->kw.p + 14 * A;
  h->tab.indicessat = h->isat.p; h->tab.satweights = h->satw.p; h->tab.swellft = h->swellft.p;
  memset(&h->dev, 0, sizeof(h->dev));
  memset(&h->mir, 0, sizeof(h->mir));
  if (cudaStreamSynchronize(st) != cudaSuccess) { ecwam_b200_destroy(h); EW_FAIL(ECWAM_B200_ECUDA, "create: stream sync failed"); }
  *out = h;
  return 0;
}

int ecwam_b200_destroy(ecwam_b200_handle h) {
  if (!h) return 0;
  cudaStreamSynchronize(h->st);
  drain_timing(h);
  if (g_const_owner == h) g_const_owner = nullptr;
  h->nbr.free(); h->halo_off.free(); h->halo_str.free(); h->send_l.free(); h->send_pre.free(); h->send_peer_of.free();
  h->recv_pre.free(); h->recv_peer_of.free(); h->recv_e.free(); h->flag.free(); h->count.free(); h->wl.free(); h->pt.free();
  h->cgext.free(); h->halo.free(); h->sendbuf.free(); h->fl3.free(); h->cosph_m.free(); h->cosph_p.free();
  h->land_cg.free(); h->cgrecv.free(); h->scr.free(); h->fldin.free(); h->tbg.free(); h->satw.free(); h->swellft.free(); h->gctab.free(); h->kw.free(); h->isat.free();
  for (void* b : h->mir_bufs) cudaFree(b);
  for (cudaEvent_t e : h->ev_up) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_done) cudaEventDestroy(e);
  if (h->ev_start) cudaEventDestroy(h->ev_start);
  if (h->ev_halo) cudaEventDestroy(h->ev_halo);
  for (cudaStream_t x : h->st_part) cudaStreamDestroy(x);
  for (cudaEvent_t e : h->ev_part) cudaEventDestroy(e);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_pack) cudaEventDestroy(h->ev_pack);
  if (h->ev_x) cudaEventDestroy(h->ev_x);
  if (h->st_x) cudaStreamDestroy(h->st_x);
  if (h->xchg_hs) cudaFreeHost(h->xchg_hs);
  if (h->xchg_hr) cudaFreeHost(h->xchg_hr);
  if (h->st_up) cudaStreamDestroy(h->st_up);
  if (h->st_dn) cudaStreamDestroy(h->st_dn);
  delete h;
  return 0;
}

int ecwam_b200_bind_fields(ecwam_b200_handle h, const ecwam_b200_fields* dev) {
  if (!h || !dev) EW_FAIL(ECWAM_B200_EINVAL, "null argument");
  if (!dev->fl1 || !dev->xllws || !dev->wavnum || !dev->cinv || !dev->cgroup || !dev->xk2cg || !dev->stokfac || !dev->depth ||
      !dev->emaxdpt || !dev->cosphm1 || !dev->aird || !dev->wdwave || !dev->cicover || !dev->wswave || !dev->wstar ||
      !dev->ufric || !dev->tauw || !dev->tauwdir || !dev->z0m || !dev->z0b || !dev->chrnck || !dev->ustokes || !dev->vstokes ||
      !dev->tauxd || !dev->tauyd || !dev->tauocxd || !dev->tauocyd || !dev->tauoc || !dev->tauicx || !dev->tauicy ||
      !dev->phiocd || !dev->phieps || !dev->phiaw || !dev->mij || !dev->wsemean || !dev->wsfmean || !dev->ustra || !dev->vstra)
    EW_FAIL(ECWAM_B200_EINVAL, "bind_fields: a required field pointer is NULL");
  if (h->par.irefra != 0 && !dev->omosnh2kd) EW_FAIL(ECWAM_B200_EINVAL, "bind_fields: IREFRA /= 0 needs OMOSNH2KD");
  if (h->par.irefra >= 2 && (!dev->ucur || !dev->vcur)) EW_FAIL(ECWAM_B200_EINVAL, "bind_fields: IREFRA = 2, 3 needs UCUR and VCUR");
  h->dev = *dev;
  h->pd.omos = dev->omosnh2kd;
  h->pd.wavn = dev->wavnum;
  h->bound = true;
  h->weights_dirty = true;
  h->cur_side = 1;
  return 0;
}

static int upload_nemo_dev(H* h) {
  if (!h->par.lwnemocou && !h->par.lwnemocoustrn && !h->par.lwnemocouwrs && !h->par.lwnemocouibr) return 0;
  NemoDev nd;
  nd.f = h->nemo; nd.nemo_on = (h->par.lwnemocou && h->nemo_bound) ? 1 : 0; nd.strn_on = h->par.lwnemocoustrn ? 1 : 0;
  nd.wrs_on = h->par.lwnemocouwrs ? 1 : 0; nd.ibr_on = (h->par.lwnemocouibr && h->nemo_bound && h->nemo.ibrmem) ? 1 : 0;
  return h->nemo_dev.upload(std::vector<NemoDev>(1, nd), h->st);
}

int ecwam_b200_bind_nemo(ecwam_b200_handle h, const ecwam_b200_nemo_fields* dev) {
  if (!h || !dev) EW_FAIL(ECWAM_B200_EINVAL, "null argument");
  const void* const* pp = (const void* const*)dev;
  for (size_t i = 0; i + 1 < sizeof(*dev) / sizeof(void*); ++i) if (!pp[i]) EW_FAIL(ECWAM_B200_EINVAL, "bind_nemo: member %zu is null", i);
  if (h->par.lwnemocouibr && !dev->ibrmem) EW_FAIL(ECWAM_B200_EINVAL, "bind_nemo: LWNEMOCOUIBR needs IBRMEM");
  h->nemo = *dev;
  h->nemo_bound = true;
  return upload_nemo_dev(h);
}

int ecwam_b200_set_exchange(ecwam_b200_handle h, ecwam_b200_exchange_fn fn, void* user, int staged) {
  if (!h) EW_FAIL(ECWAM_B200_EINVAL, "null handle");
  h->xchg_fn = fn; h->xchg_user = user; h->xchg_staged = staged != 0;
  return 0;
}

int ecwam_b200_invalidate_weights(ecwam_b200_handle h) {
  if (!h) EW_FAIL(ECWAM_B200_EINVAL, "null handle");
  h->weights_dirty = true;
  h->mir_static_done = false;   // the host-buffer path uploads its static fields (currents, depth, dispersion) again
  return 0;
}

// MPEXCHNG (mpexchng.F90:164-206) over NCCL: one grouped send/recv per neighbouring rank.
static int exchange(H* h, const double* sendbuf, double* recvbuf, size_t per_point_full, size_t per_point_now, cudaStream_t xs = nullptr) {
  if (h->nproc <= 1) return 0;
  if (!xs) xs = h->st;
  if (h->xchg_fn) {
    const int np = h->nproc;
    std::vector<long long> so(np), sc(np), ro(np), rcn(np);
    for (int q = 0; q < np; ++q) {
      so[q] = (long long)h->h_spre[q] * (long long)per_point_full; sc[q] = (long long)(h->h_spre[q + 1] - h->h_spre[q]) * (long long)per_point_now;
      ro[q] = (long long)h->h_rpre[q] * (long long)per_point_full; rcn[q] = (long long)(h->h_rpre[q + 1] - h->h_rpre[q]) * (long long)per_point_now;
    }
    if (!h->xchg_staged) {   // device pointers: a CUDA-aware message-passing library moves them
      EW_CUDA_CHECK(cudaStreamSynchronize(xs));
      const int rc = h->xchg_fn(h->xchg_user, np, sendbuf, so.data(), sc.data(), recvbuf, ro.data(), rcn.data());
      if (rc) EW_FAIL(ECWAM_B200_ESTATE, "the exchange callback returned %d", rc);
      return 0;
    }
    const size_t ns = (size_t)h->h_spre[np] * per_point_full, nr = (size_t)h->h_rpre[np] * per_point_full;
    if (ns > h->xchg_ns) { if (h->xchg_hs) cudaFreeHost(h->xchg_hs); h->xchg_hs = nullptr; EW_CUDA_CHECK(cudaMallocHost((void**)&h->xchg_hs, ns * 8)); h->xchg_ns = ns; }
    if (nr > h->xchg_nr) { if (h->xchg_hr) cudaFreeHost(h->xchg_hr); h->xchg_hr = nullptr; EW_CUDA_CHECK(cudaMallocHost((void**)&h->xchg_hr, nr * 8)); h->xchg_nr = nr; }
    for (int q = 0; q < np; ++q)
      if (sc[q] > 0) EW_CUDA_CHECK(cudaMemcpyAsync(h->xchg_hs + so[q], sendbuf + so[q], (size_t)sc[q] * 8, cudaMemcpyDeviceToHost, xs));
    EW_CUDA_CHECK(cudaStreamSynchronize(xs));
    const int rc = h->xchg_fn(h->xchg_user, np, h->xchg_hs, so.data(), sc.data(), h->xchg_hr, ro.data(), rcn.data());
    if (rc) EW_FAIL(ECWAM_B200_ESTATE, "the exchange callback returned %d", rc);
    for (int q = 0; q < np; ++q)
      if (rcn[q] > 0) EW_CUDA_CHECK(cudaMemcpyAsync(recvbuf + ro[q], h->xchg_hr + ro[q], (size_t)rcn[q] * 8, cudaMemcpyHostToDevice, xs));
    EW_CUDA_CHECK(cudaStreamSynchronize(xs));   // the staging buffer is reused by the next exchange
    return 0;
  }
  if (!h->comm) EW_FAIL(ECWAM_B200_ESTATE, "nproc > 1: neither an NCCL communicator (ecwam_b200_create) nor an exchange callback (ecwam_b200_set_exchange)");
  EW_NCCL_CHECK(ncclGroupStart());
  for (int q = 0; q < h->nproc; ++q) {
    const int ns = h->h_spre[q + 1] - h->h_spre[q], nr = h->h_rpre[q + 1] - h->h_rpre[q];
    if (ns > 0) EW_NCCL_CHECK(ncclSend(sendbuf + (size_t)h->h_spre[q] * per_point_full, (size_t)ns * per_point_now, ncclDouble, q, h->comm, xs));
    if (nr > 0) EW_NCCL_CHECK(ncclRecv(recvbuf + (size_t)h->h_rpre[q] * per_point_full, (size_t)nr * per_point_now, ncclDouble, q, h->comm, xs));
  }
  EW_NCCL_CHECK(ncclGroupEnd());
  return 0;
}

// halo exchange of the spectrum held in `src` (chunked layout with srcF frequencies), frequencies [0, nm)
static int halo_spectrum(H* h, const double* src, int srcF, int nm) {
  if (h->nproc <= 1) return 0;
  ScopedTimer t(h, "halo");
  const PropDev& d = h->pd;
  launch_pack(d, src, srcF, nullptr, 0, d.A, nm, d.Fr, h->send_l.p, h->send_pre.p, h->send_peer_of.p, h->nsend, h->sendbuf.p, h->st);
  h->nlaunch += (h->nsend > 0);
  return exchange(h, h->sendbuf.p, h->halo.p, (size_t)d.A * d.Fr, (size_t)d.A * nm);
}

// MPEXCHNG + PROPAGS2 of frequencies [0, m1) (propag_wam.F90:166, 245-251).  With several ranks the exchange runs on its own
// stream while the own points without a halo neighbour are propagated; the two strips that read the halo follow it
// (the reference posts non-blocking receives the same way, mpexchng.F90:164-206).
static int halo_propags2(H* h, const double* src, int srcF, double* dst, int dstF, int m1, const double* top = nullptr, int topF = 0) {
  const PropDev& d = h->pd;
  if (!h->overlap) {
    int rc = halo_spectrum(h, src, srcF, m1);
    if (rc) return rc;
    ScopedTimer t(h, "propags2");
    launch_propags2(d, src, srcF, dst, dstF, 0, m1, h->msplit, h->st, 0, -1, top, topF);
    h->nlaunch++;
    return 0;
  }
  ScopedTimer t(h, "propags2");      // the whole overlapped region on the main stream: pack, interior, wait, strips
  launch_pack(d, src, srcF, nullptr, 0, d.A, m1, d.Fr, h->send_l.p, h->send_pre.p, h->send_peer_of.p, h->nsend, h->sendbuf.p, h->st);
  h->nlaunch += (h->nsend > 0);
  EW_CUDA_CHECK(cudaEventRecord(h->ev_pack, h->st));
  EW_CUDA_CHECK(cudaStreamWaitEvent(h->st_x, h->ev_pack, 0));
  launch_propags2(d, src, srcF, dst, dstF, 0, m1, h->msplit, h->st, h->int_lo, h->int_hi, top, topF);
  h->nlaunch++;
  {
    ScopedTimer tx(h, "halo", h->st_x);
    int rc = exchange(h, h->sendbuf.p, h->halo.p, (size_t)d.A * d.Fr, (size_t)d.A * m1, h->st_x);
    if (rc) return rc;
  }
  EW_CUDA_CHECK(cudaEventRecord(h->ev_x, h->st_x));
  EW_CUDA_CHECK(cudaStreamWaitEvent(h->st, h->ev_x, 0));
  if (h->int_lo > 0) { launch_propags2(d, src, srcF, dst, dstF, 0, m1, h->msplit, h->st, 0, h->int_lo, top, topF); h->nlaunch++; }
  if (h->int_hi < d.nloc) { launch_propags2(d, src, srcF, dst, dstF, 0, m1, h->msplit, h->st, h->int_hi, d.nloc, top, topF); h->nlaunch++; }
  return 0;
}

// CTUWUPDT equivalent: PROENVHALO of the group velocity + per-point set-up + CFL scan (propag_wam.F90:221-236)
static int update_weights(H* h, int* cfl) {
  const PropDev& d = h->pd;
  launch_setup_points(d, h->dev.cosphm1, h->cosph_m.p, h->cosph_p.p, h->pt.p, h->st);
  launch_fill_cgext(d, h->dev.cgroup, h->dev.depth, h->dev.ucur, h->dev.vcur, h->cgext.p, h->land_cg.p, h->st);
  h->nlaunch += 3;
  if (h->nproc > 1) {   // PROENVHALO: group velocity (+ depth when IREFRA = 1) of the halo points
    launch_pack(d, nullptr, 0, h->cgext.p, 1, 1, d.nenv, d.nenv, h->send_l.p, h->send_pre.p, h->send_peer_of.p, h->nsend,
                h->sendbuf.p, h->st);
    int rc = exchange(h, h->sendbuf.p, h->cgrecv.p, (size_t)d.nenv, (size_t)d.nenv);
    if (rc) return rc;
    launch_unpack_cg(d, h->cgrecv.p, h->recv_pre.p, h->recv_peer_of.p, h->recv_e.p, h->nrecv, d.nenv, h->cgext.p, h->st);
    h->nlaunch += 2;
  }
  if (d.irefra != 0) {   // PROPDOT/GRADI (propag_wam.F90:171-216)
    launch_depth_gradients(d, h->wlat_raw.p, h->dellam.p, h->oneo2delphi, h->grad.p, h->st);
    h->nlaunch++;
  }
  int cnt = 0;
  if (d.irefra >= 2) {   // CTUWDRV with currents (ctuwdrv.F90:83-123): ICALL = 1, then ICALL = 2 with CURMASK where it failed
    launch_curmask(d, h->flag.p, h->curmask.p, 1, h->st);
    launch_ctu_check_cur(d, 0, d.Fr, h->msplit, h->flag.p, h->count.p, h->st);
    h->nlaunch += 3;
    EW_CUDA_CHECK(cudaMemcpyAsync(&cnt, h->count.p, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    EW_CUDA_CHECK(cudaStreamSynchronize(h->st));
    if (cnt > 0 && h->par.llcflcuroff) {
      launch_curmask(d, h->flag.p, h->curmask.p, 0, h->st);
      launch_ctu_check_cur(d, 0, d.Fr, h->msplit, h->flag.p, h->count.p, h->st);
      h->nlaunch += 3;
    }
  } else {
    launch_ctu_check(d, 0, d.Fr, h->msplit, h->flag.p, h->count.p, h->st);
    h->nlaunch += 2;
  }
  EW_CUDA_CHECK(cudaMemcpyAsync(&cnt, h->count.p, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  EW_CUDA_CHECK(cudaStreamSynchronize(h->st));
  *cfl = cnt;
  h->weights_dirty = false;
  h->weights_side = h->cur_side;
  return 0;
}

// returns CFL count (>=0) or error (<0); leaves m>=msplit results in FL3 and tells where the fast-wave
// frequencies ended up (in_fl3)
static int propag_core(H* h, bool* lf_in_fl3) {
  if (!h->bound) EW_FAIL(ECWAM_B200_ESTATE, "fields not bound");
  int rc = ensure_const(h);
  if (rc) return rc;
  const PropDev& d = h->pd;
  const ecwam_b200_params& p = h->par;
  int cfl = 0;
  if (h->weights_dirty || h->weights_side != h->cur_side) {
    rc = update_weights(h, &cfl);
    if (rc) return rc;
    if (cfl > 0) { ew_set_error("CTUW: CFL / weight-range check failed at %d grid points (ctuwdrv.F90:127-146)", cfl); h->weights_dirty = true; return cfl; }
  }
  rc = halo_propags2(h, h->dev.fl1, d.F, h->fl3.p, d.Fr, d.Fr);   // propag_wam.F90:166, 245-251
  if (rc) return rc;
  *lf_in_fl3 = true;
  if (p.ifrelfmax > 0 && p.ifrelfmax < d.Fr) {   // propag_wam.F90:257-313
    const int nstep = (int)nint_l(p.idelpro / p.delpro_lf);
    for (int isub = 2; isub <= nstep; ++isub) {
      const double* src = *lf_in_fl3 ? h->fl3.p : h->dev.fl1;
      double* dst = *lf_in_fl3 ? h->dev.fl1 : h->fl3.p;
      const int sF = *lf_in_fl3 ? d.Fr : d.F, dF = *lf_in_fl3 ? d.F : d.Fr;
      // with currents the frequency shift of row IFRELFMAX reads row IFRELFMAX + 1 of the spectrum of the start of the step: the bound
      // FL1, whose rows >= IFRELFMAX no sub-step writes
      rc = halo_propags2(h, src, sF, dst, dF, p.ifrelfmax, p.irefra >= 2 ? h->dev.fl1 : nullptr, d.F);
      if (rc) return rc;
      *lf_in_fl3 = !*lf_in_fl3;
    }
  }
  return 0;
}

int ecwam_b200_propag(ecwam_b200_handle h) {
  if (!h) EW_FAIL(ECWAM_B200_EINVAL, "null handle");
  bool lf_in_fl3 = true;
  int rc = propag_core(h, &lf_in_fl3);
  if (rc) return rc;
  const PropDev& d = h->pd;
  const int ms = (h->par.ifrelfmax > 0 && h->par.ifrelfmax < d.Fr) ? h->par.ifrelfmax : 0;
  ScopedTimer t(h, "copyback");
  if (lf_in_fl3) {
    launch_copyback(d, h->fl3.p, h->dev.fl1, 0, d.Fr, h->st);
    h->nlaunch++;
  } else {
    launch_copyback(d, h->fl3.p, h->dev.fl1, ms, d.Fr, h->st);
    launch_pad(d, h->dev.fl1, d.F, 0, ms, h->st);
    h->nlaunch += 2;
  }
  EW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

static ImplDev make_impl(H* h, bool from_fl3) {
  ImplDev d;
  d.P = h->par.nproma; d.A = h->par.nang; d.F = h->par.nfre; d.Fr = h->par.nfre_red; d.nchnk = h->par.nchnk;
  d.npts = (long long)d.P * d.nchnk;
  d.f = h->dev;
  d.fl_lo = from_fl3 ? h->fl3.p : h->dev.fl1;
  d.lo_F = from_fl3 ? d.Fr : d.F;
  d.lo_on = from_fl3 ? 1 : 0;
  d.scr = h->scr.p;
  d.fldin = h->fldin.p;
  d.nloc = h->pd.nloc;
  d.lwflux = h->par.lwflux;
  d.tab = h->tab;
  d.tbg = h->tbg.p;
  for (int kh = 0; kh < 2; ++kh) for (int q = 0; q < 4; ++q) d.dsb[kh][q] = h->dsh[kh][q] * 64;
  d.halo_r = h->halo_r; d.halo_c = h->halo_c;
  d.iphys = h->par.iphys; d.nsdsnth = h->nsdsnth;
  d.cy49 = ((h->par.llgcbz0 || h->par.llnormagam) ? 1 : 0) | (h->par.icode_wnd != 3 ? 2 : 0);
  d.gc = h->gctab.p;
  d.sweep_ok = h->dc.sweep_ok;
  d.ssource_pre = (h->dc.lcflx && !h->par.lwvflx_snl) ? 1 : 0;
  d.isnonlin = h->par.isnonlin;
  d.enh = h->enhp.p;
  d.ice1 = h->ice1.p; d.ice_nt = h->ice_nt; d.ice_nh = h->ice_nh; d.ice_hmin = h->ice_hmin; d.ice_dh = h->ice_dh;
  d.ice2 = h->ice2.p;
  d.nemo = h->nemo_dev.p;
  return d;
}

static int implsch_range(H* h, int ichnk0, int nchnk, bool from_fl3) {
  if (!h->bound) EW_FAIL(ECWAM_B200_ESTATE, "fields not bound");
  if (ichnk0 < 1 || nchnk < 1 || ichnk0 + nchnk - 1 > h->par.nchnk) EW_FAIL(ECWAM_B200_EINVAL, "chunk range out of bounds");
  int rc = ensure_const(h);
  if (rc) return rc;
  if ((h->par.lwnemocou || h->par.lwnemocouibr) && !h->nemo_bound)
    EW_FAIL(ECWAM_B200_ESTATE, "IMPLSCH: LWNEMOCOU / LWNEMOCOUIBR need the NEMO coupling fields (ecwam_b200_bind_nemo)");
  if (h->par.lwnemocoustrn && (!h->dev.strnms || !h->dev.cithick)) EW_FAIL(ECWAM_B200_ESTATE, "IMPLSCH: LWNEMOCOUSTRN needs STRNMS and CITHICK bound");
  if (h->par.licerun && (h->par.lciwa & 5) && !h->dev.cithick) EW_FAIL(ECWAM_B200_ESTATE, "IMPLSCH: LCIWA1 / LCIWA3 need CITHICK bound");
  ImplDev d = make_impl(h, from_fl3);
  static const char* kStage[EW_IMPLSCH_NSTAGE] = {"implsch_point", "implsch_stencil"};
  // Small blocks (several ranks, or O320 and below): the chunk range is cut into parts that run the kernel sequence on streams of
  // their own.  IMPLSCH is point-wise, so the parts are independent, and a kernel boundary stops being a device-wide barrier: the
  // partial last wave of a k_point phase (2.41 waves at 8 GPUs: the third one is 41 % full but takes as long as a full one) shares
  // the SMs with the next kernel of the other part.  ECWAM_B200_IMPLSCH_SPLIT = 0 | n overrides the automatic choice.
  static const int split_env = []() { const char* e = getenv("ECWAM_B200_IMPLSCH_SPLIT"); return e ? atoi(e) : -1; }();
  const long long np_all = (long long)nchnk * d.P;
  int ns = split_env >= 0 ? split_env : ((np_all >= 60000 && np_all <= 700000) ? 2 : 1);
  ns = std::max(1, std::min(std::min(ns, 4), nchnk));
  if (ns > 1) {
    while ((int)h->st_part.size() < ns - 1) {
      cudaStream_t x; cudaEvent_t e;
      EW_CUDA_CHECK(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
      EW_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->st_part.push_back(x); h->ev_part.push_back(e);
    }
    if (!h->ev_fork) EW_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    ScopedTimer t(h, "implsch");
    EW_CUDA_CHECK(cudaEventRecord(h->ev_fork, h->st));
    for (int k = 1; k < ns; ++k) EW_CUDA_CHECK(cudaStreamWaitEvent(h->st_part[k - 1], h->ev_fork, 0));
    for (int s = 0; s < EW_IMPLSCH_NSTAGE; ++s)
      for (int k = 0; k < ns; ++k) {
        const long long c0 = (long long)nchnk * k / ns, c1 = (long long)nchnk * (k + 1) / ns;
        rc = launch_implsch_stage(d, ((long long)(ichnk0 - 1) + c0) * d.P, (c1 - c0) * d.P, s, k == 0 ? h->st : h->st_part[k - 1]);
        if (rc) return rc;
        h->nlaunch += (s == 0) ? 2 : 1;
      }
    for (int k = 1; k < ns; ++k) {
      EW_CUDA_CHECK(cudaEventRecord(h->ev_part[k - 1], h->st_part[k - 1]));
      EW_CUDA_CHECK(cudaStreamWaitEvent(h->st, h->ev_part[k - 1], 0));
    }
    EW_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  for (int s = 0; s < EW_IMPLSCH_NSTAGE; ++s) {
    ScopedTimer t(h, kStage[s]);
    rc = launch_implsch_stage(d, (long long)(ichnk0 - 1) * d.P, (long long)nchnk * d.P, s, h->st);
    if (rc) return rc;
    h->nlaunch += (s == 0) ? 2 : 1;   // stage 0 = the two k_point kernels
  }
  EW_CUDA_CHECK(cudaGetLastError());
  return 0;
}
int ecwam_b200_implsch(ecwam_b200_handle h, int ichnk0, int nchnk) {
  if (!h) EW_FAIL(ECWAM_B200_EINVAL, "null handle");
  return implsch_range(h, ichnk0, nchnk, false);
}
int ecwam_b200_implsch_all(ecwam_b200_handle h) {
  if (!h) EW_FAIL(ECWAM_B200_EINVAL, "null handle");
  return implsch_range(h, 1, h->par.nchnk, false);
}

// ---- the reference argument lists (fortran/implsch_b200.F90, fortran/propag_wam_b200.F90) --------------------------------
// The Fortran bodies pass what the reference callers pass: the chunk's slices of the FIELD_API arrays.  The library works on
// the arrays bound with ecwam_b200_bind_fields, so every argument is checked against the bound array (a re-allocated FIELD_API
// buffer that was not re-bound is an error, not a silent use of stale memory) and the chunk is derived from FL1's address.
namespace {
struct ArgCheck {
  H* h; long long ichnk0; const char* who; int bad = 0; char first[48] = "";
  void chk(const void* passed, const void* bound, size_t chunk_bytes, const char* name) {
    if (!passed || !bound) return;   // optional argument / field that is not bound
    if ((const char*)passed != (const char*)bound + chunk_bytes * (size_t)ichnk0) { if (!bad) snprintf(first, sizeof(first), "%s", name); ++bad; }
  }
};
}  // namespace

int ecwam_b200_implsch_f(ecwam_b200_handle h, int kijs, int kijl, double* fl1, const double* wavnum, const double* cgroup,
                         const double* ciwa, const double* cinv, const double* xk2cg, const double* stokfac, const double* emaxdpt,
                         const double* depth, const int* iobnd, const int* iodp, const double* ibrmem, double* aird, double* wdwave,
                         double* cicover, double* wswave, double* wstar, double* ustra, double* vstra, double* ufric, double* tauw,
                         double* tauwdir, double* z0m, double* z0b, double* chrnck, double* cithick, double* nemoustokes,
                         double* nemovstokes, double* nemostrn, double* nphieps, double* ntauoc, double* nswh, double* nmwp,
                         double* nemotaux, double* nemotauy, double* nemotauicx, double* nemotauicy, double* nemowswave,
                         double* nemophif, double* wsemean, double* wsfmean, double* ustokes, double* vstokes, double* strnms,
                         double* tauxd, double* tauyd, double* tauocxd, double* tauocyd, double* tauoc, double* tauicx,
                         double* tauicy, double* phiocd, double* phieps, double* phiaw, int* mij, double* xllws) {
  if (!h || !fl1) EW_FAIL(ECWAM_B200_EINVAL, "IMPLSCH: null handle or FL1");
  if (!h->bound) EW_FAIL(ECWAM_B200_ESTATE, "IMPLSCH: fields not bound");
  const ecwam_b200_params& p = h->par;
  if (kijs != 1 || kijl != p.nproma) EW_FAIL(ECWAM_B200_EINVAL, "IMPLSCH: KIJS:KIJL must be 1:NPROMA_WAM = 1:%d (got %d:%d)", p.nproma, kijs, kijl);
  const size_t b4 = (size_t)p.nproma * p.nang * p.nfre * 8, b3 = (size_t)p.nproma * p.nfre * 8, b2 = (size_t)p.nproma * 8, bi = (size_t)p.nproma * 4;
  const ptrdiff_t off = (const char*)fl1 - (const char*)h->dev.fl1;
  if (off < 0 || off % (ptrdiff_t)b4 != 0 || off / (ptrdiff_t)b4 >= p.nchnk)
    EW_FAIL(ECWAM_B200_ESTATE, "IMPLSCH: FL1 is not a chunk of the bound FL1 array (re-bind the fields after FIELD_API re-allocates)");
  ArgCheck a{h, (long long)(off / (ptrdiff_t)b4), "IMPLSCH"};
  const ecwam_b200_fields& f = h->dev;
  a.chk(xllws, f.xllws, b4, "XLLWS");
  a.chk(wavnum, f.wavnum, b3, "WAVNUM"); a.chk(cgroup, f.cgroup, b3, "CGROUP"); a.chk(ciwa, f.ciwa, b3, "CIWA"); a.chk(cinv, f.cinv, b3, "CINV");
  a.chk(xk2cg, f.xk2cg, b3, "XK2CG"); a.chk(stokfac, f.stokfac, b3, "STOKFAC");
  a.chk(emaxdpt, f.emaxdpt, b2, "EMAXDPT"); a.chk(depth, f.depth, b2, "DEPTH");
  a.chk(aird, f.aird, b2, "AIRD"); a.chk(wdwave, f.wdwave, b2, "WDWAVE"); a.chk(cicover, f.cicover, b2, "CICOVER"); a.chk(wswave, f.wswave, b2, "WSWAVE");
  a.chk(wstar, f.wstar, b2, "WSTAR"); a.chk(ustra, f.ustra, b2, "USTRA"); a.chk(vstra, f.vstra, b2, "VSTRA"); a.chk(ufric, f.ufric, b2, "UFRIC");
  a.chk(tauw, f.tauw, b2, "TAUW"); a.chk(tauwdir, f.tauwdir, b2, "TAUWDIR"); a.chk(z0m, f.z0m, b2, "Z0M"); a.chk(z0b, f.z0b, b2, "Z0B");
  a.chk(chrnck, f.chrnck, b2, "CHRNCK"); a.chk(cithick, f.cithick, b2, "CITHICK");
  a.chk(wsemean, f.wsemean, b2, "WSEMEAN"); a.chk(wsfmean, f.wsfmean, b2, "WSFMEAN"); a.chk(ustokes, f.ustokes, b2, "USTOKES"); a.chk(vstokes, f.vstokes, b2, "VSTOKES");
  a.chk(strnms, f.strnms, b2, "STRNMS"); a.chk(tauxd, f.tauxd, b2, "TAUXD"); a.chk(tauyd, f.tauyd, b2, "TAUYD"); a.chk(tauocxd, f.tauocxd, b2, "TAUOCXD");
  a.chk(tauocyd, f.tauocyd, b2, "TAUOCYD"); a.chk(tauoc, f.tauoc, b2, "TAUOC"); a.chk(tauicx, f.tauicx, b2, "TAUICX"); a.chk(tauicy, f.tauicy, b2, "TAUICY");
  a.chk(phiocd, f.phiocd, b2, "PHIOCD"); a.chk(phieps, f.phieps, b2, "PHIEPS"); a.chk(phiaw, f.phiaw, b2, "PHIAW");
  a.chk(mij, f.mij, bi, "MIJ");
  if (a.bad) EW_FAIL(ECWAM_B200_ESTATE, "IMPLSCH: %d argument(s) are not chunk %lld of the bound arrays (first: %s): re-bind the fields", a.bad, a.ichnk0 + 1, a.first);
  // IOBND, IODP, IBRMEM are not read on this path; the NEMO accumulators must be chunk ICHNK of the arrays bound with bind_nemo
  (void)iobnd; (void)iodp; (void)ibrmem;
  if (h->par.lwnemocou) {
    if (!h->nemo_bound) EW_FAIL(ECWAM_B200_ESTATE, "IMPLSCH: LWNEMOCOU needs ecwam_b200_bind_nemo");
    const ecwam_b200_nemo_fields& o = h->nemo;
    a.chk(nemoustokes, o.nemoustokes, b2, "NEMOUSTOKES"); a.chk(nemovstokes, o.nemovstokes, b2, "NEMOVSTOKES"); a.chk(nemostrn, o.nemostrn, b2, "NEMOSTRN");
    a.chk(nphieps, o.nphieps, b2, "NPHIEPS"); a.chk(ntauoc, o.ntauoc, b2, "NTAUOC"); a.chk(nswh, o.nswh, b2, "NSWH"); a.chk(nmwp, o.nmwp, b2, "NMWP");
    a.chk(nemotaux, o.nemotaux, b2, "NEMOTAUX"); a.chk(nemotauy, o.nemotauy, b2, "NEMOTAUY"); a.chk(nemotauicx, o.nemotauicx, b2, "NEMOTAUICX");
    a.chk(nemotauicy, o.nemotauicy, b2, "NEMOTAUICY"); a.chk(nemowswave, o.nemowswave, b2, "NEMOWSWAVE"); a.chk(nemophif, o.nemophif, b2, "NEMOPHIF");
    if (a.bad) EW_FAIL(ECWAM_B200_ESTATE, "IMPLSCH: %d NEMO argument(s) are not chunk %lld of the arrays bound with bind_nemo (first: %s)", a.bad, a.ichnk0 + 1, a.first);
  } else {
    (void)nemoustokes; (void)nemovstokes; (void)nemostrn; (void)nphieps; (void)ntauoc; (void)nswh; (void)nmwp;
    (void)nemotaux; (void)nemotauy; (void)nemotauicx; (void)nemotauicy; (void)nemowswave; (void)nemophif;
  }
  return implsch_range(h, (int)a.ichnk0 + 1, 1, false);
}

int ecwam_b200_propag_wam_f(ecwam_b200_handle h, const double* wavnum, const double* cgroup, const double* omosnh2kd, double* fl1,
                            const double* depth, const double* dellam1, const double* cosphm1, const double* ucur, const double* vcur) {
  if (!h || !fl1) EW_FAIL(ECWAM_B200_EINVAL, "PROPAG_WAM: null handle or FL1");
  if (!h->bound) EW_FAIL(ECWAM_B200_ESTATE, "PROPAG_WAM: fields not bound");
  ArgCheck a{h, 0, "PROPAG_WAM"};
  const ecwam_b200_fields& f = h->dev;
  a.chk(fl1, f.fl1, 0, "FL1"); a.chk(wavnum, f.wavnum, 0, "WAVNUM"); a.chk(cgroup, f.cgroup, 0, "CGROUP"); a.chk(omosnh2kd, f.omosnh2kd, 0, "OMOSNH2KD");
  a.chk(depth, f.depth, 0, "DEPTH"); a.chk(dellam1, f.dellam1, 0, "DELLAM1"); a.chk(cosphm1, f.cosphm1, 0, "COSPHM1");
  a.chk(ucur, f.ucur, 0, "UCUR"); a.chk(vcur, f.vcur, 0, "VCUR");
  if (a.bad) EW_FAIL(ECWAM_B200_ESTATE, "PROPAG_WAM: %d argument(s) are not the bound arrays (first: %s): re-bind the fields", a.bad, a.first);
  return ecwam_b200_propag(h);
}

// PROPAG_WAM + IMPLSCH with the block->chunk copy of PROPAG_WAM (propag_wam.F90:368-405) folded into IMPLSCH's loads:
// the IMPLSCH kernels read the propagated frequencies straight from the propagation scratch and write FL1.
int ecwam_b200_wamintgr(ecwam_b200_handle h) {
  if (!h) EW_FAIL(ECWAM_B200_EINVAL, "null handle");
  bool lf_in_fl3 = true;
  int rc = propag_core(h, &lf_in_fl3);
  if (rc) return rc;
  if (!lf_in_fl3) {   // odd number of fast-wave sub-steps left the low frequencies in FL1: finish PROPAG_WAM the plain way
    const PropDev& d = h->pd;
    launch_copyback(d, h->fl3.p, h->dev.fl1, h->par.ifrelfmax, d.Fr, h->st);
    launch_pad(d, h->dev.fl1, d.F, 0, h->par.ifrelfmax, h->st);
    h->nlaunch += 2;
    return implsch_range(h, 1, h->par.nchnk, false);
  }
  launch_pad(h->pd, h->fl3.p, h->pd.Fr, 0, h->pd.Fr, h->st);   // padded lanes of the last chunk (propag_wam.F90:388-398)
  return implsch_range(h, 1, h->par.nchnk, true);
}

// ---- the steps either side of the hot path: NEWWIND, OUTBS/OUTBLOCK core, OUTWNORM -------------------------------
int ecwam_b200_newwind(ecwam_b200_handle h, const ecwam_b200_forcing_next* next) {
  if (!h || !next) EW_FAIL(ECWAM_B200_EINVAL, "null argument");
  if (!h->bound) EW_FAIL(ECWAM_B200_ESTATE, "fields not bound");
  if (h->par.icode_wnd != 3) EW_FAIL(ECWAM_B200_EINVAL, "NEWWIND: with ICODE_WND = 1, 2 the forcing is the friction velocity: use ecwam_b200_newwind_ustar");
  const void* const* pp = (const void* const*)next;
  for (size_t i = 0; i < sizeof(*next) / sizeof(void*); ++i) if (!pp[i]) EW_FAIL(ECWAM_B200_EINVAL, "NEWWIND: FF_NEXT member %zu is null", i);
  ScopedTimer t(h, "newwind");
  launch_newwind((long long)h->par.nproma * h->par.nchnk, h->dev, *next, h->dc.ACD, h->dc.BCD, h->dc.EPSMIN, h->st);
  h->nlaunch++;
  EW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ecwam_b200_newwind_ustar(ecwam_b200_handle h, const ecwam_b200_forcing_next* next, const double* ufric_next) {
  if (!h || !next || !ufric_next) EW_FAIL(ECWAM_B200_EINVAL, "null argument");
  if (!h->bound) EW_FAIL(ECWAM_B200_ESTATE, "fields not bound");
  if (h->par.icode_wnd == 3) EW_FAIL(ECWAM_B200_EINVAL, "NEWWIND: ICODE_WND = 3 takes the 10 m wind: use ecwam_b200_newwind");
  if (!next->wdwave || !next->aird || !next->wstar || !next->cicover || !next->cithick || !next->ustra || !next->vstra)
    EW_FAIL(ECWAM_B200_EINVAL, "NEWWIND: a FF_NEXT member is null");
  ScopedTimer t(h, "newwind");
  launch_newwind_ustar((long long)h->par.nproma * h->par.nchnk, h->dev, *next, ufric_next, h->dc.ALPHA, h->st);
  h->nlaunch++;
  EW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ecwam_b200_getwnd(ecwam_b200_handle h, const ecwam_b200_fieldg* g, const ecwam_b200_getwnd_opts* o, const int* ifromij,
                      const int* jfromij, const ecwam_b200_forcing_next* next) {
  if (!h || !g || !o || !ifromij || !jfromij || !next) EW_FAIL(ECWAM_B200_EINVAL, "null argument");
  if (!h->bound) EW_FAIL(ECWAM_B200_ESTATE, "fields not bound");
  if (h->par.icode_wnd != 3) EW_FAIL(ECWAM_B200_EINVAL, "GETWND: only ICODE_WND = 3 (10 m wind components) is built");
  if (o->nxe < o->nxs || o->nye < o->nys) EW_FAIL(ECWAM_B200_EINVAL, "GETWND: empty forcing grid");
  if (o->iparamci != 31 && o->iparamci != 139) EW_FAIL(ECWAM_B200_EINVAL, "GETWND: IPARAMCI must be 31 (sea-ice fraction) or 139 (SST)");
  if (!g->uwnd || !g->vwnd || !g->aird || !g->wstar || !g->cicover || !g->cithick || !g->ustra || !g->vstra)
    EW_FAIL(ECWAM_B200_EINVAL, "GETWND: a FIELDG member is null");
  if ((o->llwswave && !g->wswave) || (o->llwdwave && !g->wdwave)) EW_FAIL(ECWAM_B200_EINVAL, "GETWND: LLWSWAVE / LLWDWAVE need FIELDG%%WSWAVE / WDWAVE");
  const void* const* pp = (const void* const*)next;
  for (size_t i = 0; i < sizeof(*next) / sizeof(void*); ++i) if (!pp[i]) EW_FAIL(ECWAM_B200_EINVAL, "GETWND: FF_NEXT member %zu is null", i);
  GetwndArgs a;
  a.npts = (long long)h->par.nproma * h->par.nchnk;
  a.nx = o->nxe - o->nxs + 1;
  a.lcorrel = (o->lrelwind && (h->par.irefra == 2 || h->par.irefra == 3)) ? 1 : 0;      // wamwnd.F90:122-127 (LWCOU = F)
  a.licerun = h->par.licerun; a.lmaskice = h->par.lmaskice;
  a.wspmin = h->par.wspmin; a.zpi = h->dc.ZPI;
  a.ucur = h->dev.ucur; a.vcur = h->dev.vcur;
  if (a.lcorrel && (!a.ucur || !a.vcur)) EW_FAIL(ECWAM_B200_ESTATE, "GETWND: the relative-wind correction needs UCUR / VCUR bound");
  ScopedTimer t(h, "getwnd");
  launch_getwnd(a, *g, *o, ifromij, jfromij, *next, h->st);
  h->nlaunch++;
  EW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ecwam_b200_no_source(ecwam_b200_handle h, int llsource_off) {
  if (!h) EW_FAIL(ECWAM_B200_EINVAL, "null handle");
  if (!h->bound) EW_FAIL(ECWAM_B200_ESTATE, "fields not bound");
  const long long n2 = (long long)h->par.nproma * h->par.nchnk;
  launch_no_source(n2 * h->par.nang * h->par.nfre, n2, h->dev.fl1, h->dev.xllws, h->dev.mij, h->par.nfre, llsource_off != 0, h->dc.EPSMIN, h->st);
  h->nlaunch++;
  EW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ecwam_b200_outparam_supported(int itg) {
  static const int ok[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 20, 21, 22, 23, 24, 25, 26, 27, 28, 32, 35, 36, 37, 38,
                           39, 40, 41, 52, 53, 54, 55, 56, 62, 63, 64, 65, 66, 67, 68, 69, 73, 74, 75, 76, 77};
  for (int v : ok) if (v == itg) return 1;
  return 0;
}

// SEBTMEAN (sebtmean.F90:66-198) as weights on the 1-D spectrum rows: EBT = EPSMIN + sum_m w[m] * sum_k F(k,m).  The routine is
// linear in F (trapezoid between the cut-offs, linear interpolation at the two ends, f**-5 tail), so the weights depend on
// (TB, TT) and the frequency grid only.
static void sebtmean_weights(const DevConst& c, double TB, double TT, double* w) {
  const int NFRE = c.F;
  auto FR = [&](int m) { return c.FR[m - 1]; };
  for (int m = 0; m < NFRE; ++m) w[m] = 0.0;
  std::vector<double> FRLOC(NFRE + 2, 0.0);
  std::vector<std::vector<double>> node(NFRE + 2, std::vector<double>(NFRE + 1, 0.0));   // node[M][row]: F1D(:,M) as a combination of rows
  double FBOT = 1.0 / std::max(TT, c.EPSMIN);
  const double FCUTB_FT = std::min(FBOT, FR(NFRE));
  const double FCUTB = std::max(FR(1), FCUTB_FT);
  FBOT = std::max(FBOT, FR(NFRE));
  int MCUTB = 1;
  while (FR(MCUTB) < FCUTB && MCUTB < NFRE) ++MCUTB;
  double FTOP = 1.0 / std::max(TB, c.EPSMIN);
  const double FCUTT = std::max(FR(1), std::min(FTOP, FR(NFRE)));
  FTOP = std::max(FTOP, FR(NFRE));
  int MCUTT = NFRE;
  while (FR(MCUTT) > FCUTT && MCUTT > 1) --MCUTT;
  if (FCUTB == FCUTT) MCUTT = MCUTB - 1;
  if (MCUTB > 1) {
    FRLOC[MCUTB - 1] = FCUTB;
    const double WL = (FR(MCUTB) - FCUTB) / (FR(MCUTB) - FR(MCUTB - 1));
    node[MCUTB - 1][MCUTB - 1] = WL; node[MCUTB - 1][MCUTB] = 1.0 - WL;
  }
  for (int M = MCUTB; M <= MCUTT; ++M) { FRLOC[M] = FR(M); std::fill(node[M].begin(), node[M].end(), 0.0); node[M][M] = 1.0; }
  if (MCUTT < NFRE) {
    FRLOC[MCUTT + 1] = FCUTT;
    // MCUTT = 0 (band entirely below FR(1)): the reference's WL = 0/(FR(1)-FR(0)) reads FR(0) out of bounds; take WL = 0
    const double WL = MCUTT >= 1 ? (FR(MCUTT + 1) - FCUTT) / (FR(MCUTT + 1) - FR(MCUTT)) : 0.0;
    std::fill(node[MCUTT + 1].begin(), node[MCUTT + 1].end(), 0.0);
    if (MCUTT >= 1) node[MCUTT + 1][MCUTT] = WL;
    node[MCUTT + 1][MCUTT + 1] = 1.0 - WL;
  }
  for (int M = std::max(MCUTB - 1, 1); M <= std::min(MCUTT, NFRE - 1); ++M) {
    const double DF = 0.5 * (FRLOC[M + 1] - FRLOC[M]);
    for (int r = 1; r <= NFRE; ++r) w[r - 1] += DF * (node[M + 1][r] + node[M][r]);
  }
  if (FCUTB_FT < FCUTB && FCUTB == FR(1)) {
    const double WL = (FR(1) - FCUTB_FT) / FR(1), WR = 1.0 - WL;
    const double DF = 0.5 * (FR(1) - FCUTB_FT) * (1.0 + WR);
    for (int r = 1; r <= NFRE; ++r) w[r - 1] += DF * node[1][r];
  }
  if (FBOT < FTOP) w[NFRE - 1] += 0.25 * c.FR5[NFRE - 1] * (1.0 / (FBOT * FBOT * FBOT * FBOT) - 1.0 / (FTOP * FTOP * FTOP * FTOP));
  for (int m = 0; m < NFRE; ++m) w[m] *= c.DELTH;
}

static int check_outsel(const ecwam_b200_outsel* sel) {
  if (!sel || !sel->itg || !sel->icemask || !sel->seamask) EW_FAIL(ECWAM_B200_EINVAL, "null output selection");
  if (sel->niprmout < 1 || sel->niprmout > EW_OUT_MAXCOL) EW_FAIL(ECWAM_B200_EINVAL, "NIPRMOUT must be 1..%d", EW_OUT_MAXCOL);
  for (int i = 0; i < sel->niprmout; ++i)
    if (!ecwam_b200_outparam_supported(sel->itg[i])) EW_FAIL(ECWAM_B200_EINVAL, "OUTBLOCK parameter %d is not built (see ecwam_b200.h)", sel->itg[i]);
  return 0;
}
static bool wants(const ecwam_b200_outsel* sel, int itg) {
  for (int i = 0; i < sel->niprmout; ++i) if (sel->itg[i] == itg) return true;
  return false;
}

int ecwam_b200_outbs(ecwam_b200_handle h, const ecwam_b200_outsel* sel, const int* iodp, double* bout) {
  if (!h || !bout) EW_FAIL(ECWAM_B200_EINVAL, "null argument");
  if (!h->bound) EW_FAIL(ECWAM_B200_ESTATE, "fields not bound");
  int rc = check_outsel(sel);
  if (rc) return rc;
  const DevConst& c = h->dc;
  OutConst oc;
  memset(&oc, 0, sizeof(oc));
  oc.A = c.A; oc.F = c.F; oc.NFRE_ODD = c.NFRE_ODD; oc.licerun = c.licerun; oc.lmaskice = c.lmaskice; oc.llsource = sel->llsource;
  oc.ncol = sel->niprmout;
  for (int i = 0; i < oc.ncol; ++i) { oc.itg[i] = sel->itg[i]; oc.icemask[i] = sel->icemask[i]; oc.seamask[i] = sel->seamask[i]; }
  oc.EPSMIN = c.EPSMIN; oc.EPSUS = c.EPSUS; oc.DELTH = c.DELTH; oc.WETAIL = c.WETAIL; oc.FRTAIL = c.FRTAIL; oc.WP1TAIL = c.WP1TAIL;
  oc.WP2TAIL = 0.5;          // yowfred.F90:54
  oc.ZPI = c.ZPI; oc.G = c.G; oc.GM1 = c.GM1;
  oc.DEG = 360.0 / c.ZPI;    // = 180/PI bit for bit (iniwcst.F90:63)
  oc.XKAPPA = c.XKAPPA; oc.XNLEV = c.XNLEV; oc.ALPHAMIN = c.ALPHAMIN;
  oc.ALPHAMAX = 0.11;        // yowphys.F90:55
  oc.llgcbz0 = c.llgcbz0;
  if (wants(sel, 9)) {   // MEANSQS (meansqs.F90:80-100) with XKMSS_CUTOFF = XK_GC(NWAV_GC) (userin.F90:1213-1215)
    if (!h->mss_ok) EW_FAIL(ECWAM_B200_EINVAL, "OUTBLOCK parameter 9 (MEANSQS) needs the gravity-capillary tables incl. delkcc_gc");
    oc.want_mss = 1; oc.NWAV_GC = h->gc_n; oc.SQRTGOSURFT = h->gc_sqrtgosurft; oc.XKM1_GC = 1.0 / h->gc_xk1;
    oc.XLOGKRATIOM1_GC = 1.0 / std::log(1.2);
    oc.XKMSS = h->gc_xkn;
    oc.NE_MSS = std::min(std::max((int)std::lround(std::log(oc.XKMSS * oc.XKM1_GC) * oc.XLOGKRATIOM1_GC), 1), oc.NWAV_GC);   // meansqs_gc.F90:60
    oc.FCUT_MSS = std::sqrt(c.G * oc.XKMSS) / c.ZPI;
    const int nfre_mss = (int)(std::log(oc.FCUT_MSS / c.FR[0]) / std::log(h->fratio)) + 1;
    oc.NFRE_EFF = std::min(c.F, nfre_mss);
    oc.ALPHAPMAX = h->alphapmax; oc.ZPI4GM2_FR5N = c.ZPI4GM2 * c.FR5[c.F - 1];
  }
  oc.ROWATER = 1000.0;       // yowpcons.F90
  oc.rnum = c.rnum; oc.flmin = c.flmin; oc.cithrsh = c.cithrsh; oc.zmiss = sel->zmiss;
  for (int m = 0; m < c.F; ++m) { oc.FR[m] = c.FR[m]; oc.DFIM[m] = c.DFIM[m]; oc.DFIMOFR[m] = c.DFIMOFR[m]; oc.DFIMFR[m] = c.DFIMFR[m]; oc.DFIM_SIM[m] = c.DFIM_SIM[m]; }
  for (int k = 0; k < c.A; ++k) { oc.TH[k] = c.TH[k]; oc.COSTH[k] = c.COSTH[k]; oc.SINTH[k] = c.SINTH[k]; }
  {   // SE10MEAN (se10mean.F90:60-66: 10 s .. 1/FR(1)) and the six bands of mpcrtbl.F90:373-399
    static const double BANDS[7][2] = {{10, 0}, {10, 12}, {12, 14}, {14, 17}, {17, 21}, {21, 25}, {25, 30}};
    for (int b = 0; b < 7; ++b) sebtmean_weights(c, BANDS[b][0], b == 0 ? 1.0 / c.FR[0] : BANDS[b][1], oc.SEBT[b]);
  }
  rc = upload_out_const(oc, h->st);
  if (rc) return rc;
  OutDev d;
  d.P = h->par.nproma; d.A = c.A; d.F = c.F; d.nchnk = h->par.nchnk;
  d.npts = (long long)d.P * d.nchnk;
  d.f = h->dev; d.iodp = iodp; d.bout = bout; d.gc = h->gctab.p; d.fl2 = nullptr;
  ScopedTimer t(h, "outblock");
  if (h->par.irefra >= 2) {   // INTPOL into the wind-input scratch of IMPLSCH (free between steps), outblock.F90:168-169
    launch_intpol(d, h->fldin.p, h->dc.FRATIO, h->dc.FLOGSPRDM1, h->dc.FR5[c.F - 1], h->st);
    d.fl2 = h->fldin.p;
    h->nlaunch++;
  }
  rc = launch_outblock(d, h->st);
  if (rc) return rc;
  h->nlaunch++;
  EW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ecwam_b200_outwnorm(ecwam_b200_handle h, const ecwam_b200_outsel* sel, const double* bout, int llglobal, int niblo,
                        const int* ij2newij, const int* nstart, const int* nend, double* wnorm) {
  if (!h || !bout || !wnorm) EW_FAIL(ECWAM_B200_EINVAL, "null argument");
  int rc = check_outsel(sel);
  if (rc) return rc;
  const int ncol = sel->niprmout, P = h->par.nproma, np = h->nproc;
  const long long nloc = h->pd.nloc;
  const double HUGE_ = 1.7976931348623157e308;
  if (np > 1 && !h->comm) EW_FAIL(ECWAM_B200_ESTATE, "outwnorm over %d ranks needs the NCCL communicator (its gather is not part of the exchange callback)", np);
  ScopedTimer t(h, "outwnorm");
  if (!llglobal) {   // mpminmaxavg.F90:160-191
    const size_t ns = norm_scratch_doubles(ncol);
    if (h->normbuf.n < ns + (size_t)4 * ncol * np) { rc = h->normbuf.alloc(ns + (size_t)4 * ncol * np); if (rc) return rc; }
    double* out4 = h->normbuf.p + ns - (size_t)4 * ncol;
    double* gath = h->normbuf.p + ns;
    launch_norm_local(bout, P, ncol, nloc, sel->zmiss, h->normbuf.p, out4, h->st);
    h->nlaunch += 2;
    if (np > 1) EW_NCCL_CHECK(ncclAllGather(out4, gath, (size_t)4 * ncol, ncclDouble, h->comm, h->st));
    else EW_CUDA_CHECK(cudaMemcpyAsync(gath, out4, sizeof(double) * 4 * ncol, cudaMemcpyDeviceToDevice, h->st));
    std::vector<double> hg((size_t)4 * ncol * np);
    EW_CUDA_CHECK(cudaMemcpyAsync(hg.data(), gath, hg.size() * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    EW_CUDA_CHECK(cudaStreamSynchronize(h->st));
    for (int i = 0; i < ncol; ++i) {
      double s = 0.0, c = 0.0, mn = HUGE_, mx = -HUGE_;
      for (int q = 0; q < np; ++q) {   // fixed rank order (the reference's MPL_ALLREDUCE with LDREPROD)
        const double* v = &hg[((size_t)q * ncol + i) * 4];
        s += v[0]; c += v[1]; mn = std::min(mn, v[2]); mx = std::max(mx, v[3]);
      }
      wnorm[4 * i + 1] = mn; wnorm[4 * i + 2] = mx; wnorm[4 * i + 3] = c;
      wnorm[4 * i + 0] = c < 1.0 ? -HUGE_ : s / c;
    }
    return 0;
  }
  // LLNORMWAMOUT_GLOBAL (mpminmaxavg.F90:121-153)
  if (niblo < nloc) EW_FAIL(ECWAM_B200_EINVAL, "outwnorm: niblo < own points");
  if (np > 1 && (!nstart || !nend)) EW_FAIL(ECWAM_B200_EINVAL, "outwnorm: the global norm over %d ranks needs nstart/nend", np);
  if (np == 1 && niblo != nloc) EW_FAIL(ECWAM_B200_EINVAL, "outwnorm: niblo must equal the own points on one rank");
  const size_t nz = (size_t)ncol * niblo + (size_t)4 * ncol;
  if (h->zglobal.n < nz) { rc = h->zglobal.alloc(nz); if (rc) return rc; }
  double* zg = h->zglobal.p;
  double* out4 = zg + (size_t)ncol * niblo;
  const long long off = np > 1 ? nstart[h->irank0] - 1 : 0;
  launch_pack_cols(bout, P, ncol, nloc, zg + off, niblo, h->st);
  h->nlaunch++;
  const int* ijd = nullptr;
  if (np > 1) {
    EW_NCCL_CHECK(ncclGroupStart());
    for (int q = 0; q < np; ++q) {
      const size_t cnt = (size_t)(nend[q] - nstart[q] + 1);
      for (int i = 0; i < ncol; ++i) {
        double* seg = zg + (size_t)i * niblo + (nstart[q] - 1);
        EW_NCCL_CHECK(ncclBroadcast(seg, seg, cnt, ncclDouble, q, h->comm, h->st));
      }
    }
    EW_NCCL_CHECK(ncclGroupEnd());
    if (ij2newij) {   // IJ = IJ2NEWIJ(IJOLD) unless LL1D (mpminmaxavg.F90:134-138)
      if (h->ij2new_n != niblo + 1) {
        std::vector<int> v(ij2newij, ij2newij + niblo + 1);
        rc = h->ij2new_d.upload(v, h->st);
        if (rc) return rc;
        h->ij2new_n = niblo + 1;
      }
      ijd = h->ij2new_d.p;
    }
  }
  launch_norm_seq(zg, ijd, niblo, ncol, sel->zmiss, out4, h->st);
  h->nlaunch++;
  EW_CUDA_CHECK(cudaMemcpyAsync(wnorm, out4, sizeof(double) * 4 * ncol, cudaMemcpyDeviceToHost, h->st));
  EW_CUDA_CHECK(cudaStreamSynchronize(h->st));
  EW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ecwam_b200_synchronize(ecwam_b200_handle h) {
  if (!h) EW_FAIL(ECWAM_B200_EINVAL, "null handle");
  EW_CUDA_CHECK(cudaStreamSynchronize(h->st));
  EW_CUDA_CHECK(cudaGetLastError());
  return 0;
}
long long ecwam_b200_launch_count(ecwam_b200_handle h) { return h ? h->nlaunch : 0; }
int ecwam_b200_timing_enable(ecwam_b200_handle h, int on) { if (!h) return ECWAM_B200_EINVAL; h->timing = on != 0; return 0; }
int ecwam_b200_timing_get(ecwam_b200_handle h, const char* name, double* total_ms, long long* count) {
  if (!h || !name) EW_FAIL(ECWAM_B200_EINVAL, "null argument");
  drain_timing(h);
  auto it = h->tm.find(name);
  if (total_ms) *total_ms = it == h->tm.end() ? 0.0 : it->second.total_ms;
  if (count) *count = it == h->tm.end() ? 0 : it->second.count;
  return 0;
}
int ecwam_b200_timing_reset(ecwam_b200_handle h) {
  if (!h) return ECWAM_B200_EINVAL;
  drain_timing(h);
  h->tm.clear();
  return 0;
}

// ---- host-buffer entry point ----------------------------------------------------------------------------
namespace {
struct FieldDesc { size_t off; int kind; /* 0: (P,A,F,C) 1: (P,F,C) 2: (P,C) double 3: (P,C) int */ int dir; /* 1 in(static) 2 in(per step) 4 out */ };
#define FD(member, kind, dir) {offsetof(ecwam_b200_fields, member), kind, dir}
const FieldDesc kFields[] = {
    FD(fl1, 0, 2 | 4), FD(xllws, 0, 8),
    FD(wavnum, 1, 1), FD(cinv, 1, 1), FD(cgroup, 1, 1), FD(xk2cg, 1, 1), FD(omosnh2kd, 1, 1), FD(stokfac, 1, 1), FD(ciwa, 1, 1),
    FD(depth, 2, 1), FD(emaxdpt, 2, 1), FD(dellam1, 2, 1), FD(cosphm1, 2, 1), FD(ucur, 2, 1), FD(vcur, 2, 1),
    FD(aird, 2, 2), FD(wdwave, 2, 2), FD(cicover, 2, 2), FD(wswave, 2, 2), FD(wstar, 2, 2), FD(ustra, 2, 2), FD(vstra, 2, 2),
    FD(ufric, 2, 2 | 4), FD(tauw, 2, 2 | 4), FD(tauwdir, 2, 2 | 4), FD(z0m, 2, 2 | 4), FD(z0b, 2, 2 | 4), FD(chrnck, 2, 2 | 4),
    FD(cithick, 2, 2),
    FD(wsemean, 2, 4), FD(wsfmean, 2, 4), FD(ustokes, 2, 4), FD(vstokes, 2, 4), FD(strnms, 2, 0), FD(tauxd, 2, 4), FD(tauyd, 2, 4),
    FD(tauocxd, 2, 4), FD(tauocyd, 2, 4), FD(tauoc, 2, 4), FD(tauicx, 2, 4), FD(tauicy, 2, 4), FD(phiocd, 2, 4), FD(phieps, 2, 4),
    FD(phiaw, 2, 4), FD(mij, 3, 4)};
#undef FD
inline void*& fptr(ecwam_b200_fields& f, size_t off) { return *(void**)((char*)&f + off); }
inline void* fptr_c(const ecwam_b200_fields& f, size_t off) { return *(void* const*)((const char*)&f + off); }
}  // namespace

// ---- resident-state step: only the forcing goes in and the 1-D results come out ------------------------------------------
// How the reference's GPU build moves data (wamintgr_loki_gpu.F90:141-201): the spectrum and the model fields live on the device,
// a step receives the new forcing and hands back the integrated parameters.  host_next: the eight FF_NEXT fields in (pinned)
// host memory; host_out: a fields struct whose non-NULL 1-D members (UFRIC ... PHIAW, MIJ) receive the step's results.
int ecwam_b200_wamintgr_forced(ecwam_b200_handle h, const ecwam_b200_forcing_next* host_next, const ecwam_b200_fields* host_out,
                               long long* h2d_bytes, long long* d2h_bytes) {
  if (!h || !host_next) EW_FAIL(ECWAM_B200_EINVAL, "null argument");
  if (!h->bound) EW_FAIL(ECWAM_B200_ESTATE, "fields not bound");
  const size_t npts = (size_t)h->par.nproma * h->par.nchnk;
  const int NF = (int)(sizeof(*host_next) / sizeof(void*));
  if (!h->frc_next.p) { int rc = h->frc_next.alloc(npts * NF); if (rc) return rc; }
  const void* const* src = (const void* const*)host_next;
  ecwam_b200_forcing_next dn;
  const void** dpp = (const void**)&dn;
  long long nin = 0, nout = 0;
  for (int i = 0; i < NF; ++i) {
    if (!src[i]) EW_FAIL(ECWAM_B200_EINVAL, "wamintgr_forced: FF_NEXT member %d is null", i);
    EW_CUDA_CHECK(cudaMemcpyAsync(h->frc_next.p + (size_t)i * npts, src[i], npts * 8, cudaMemcpyHostToDevice, h->st));
    dpp[i] = h->frc_next.p + (size_t)i * npts;
    nin += (long long)(npts * 8);
  }
  int rc = ecwam_b200_newwind(h, &dn);
  if (rc) return rc;
  rc = ecwam_b200_wamintgr(h);
  if (rc) return rc;
  if (host_out) {
    for (const FieldDesc& fd : kFields) {
      if (!(fd.dir & 4) || fd.kind < 2) continue;                 // the 1-D outputs of IMPLSCH (and MIJ)
      char* dst = (char*)fptr_c(*host_out, fd.off);
      const char* dsrc = (const char*)fptr_c(h->dev, fd.off);
      if (!dst || !dsrc) continue;
      const size_t nb = npts * (fd.kind == 3 ? 4 : 8);
      EW_CUDA_CHECK(cudaMemcpyAsync(dst, dsrc, nb, cudaMemcpyDeviceToHost, h->st));
      nout += (long long)nb;
    }
  }
  EW_CUDA_CHECK(cudaStreamSynchronize(h->st));
  if (h2d_bytes) *h2d_bytes = nin;
  if (d2h_bytes) *d2h_bytes = nout;
  return 0;
}

// Host-buffer entry: the caller's arrays live in (pinned) host memory.  Copies and kernels are pipelined over bands of
// NPROMA chunks on three streams: upload of band b+2 | PROPAGS2 of band b+1 | IMPLSCH of band b | download of band b-1.
//   * full pipeline (one rank, no fast-wave sub-steps, CTU set-up done): PROPAGS2 of a band only needs the bands next to it
//     (a band is at least as long as the largest neighbour distance of the grid), and IMPLSCH of band b is issued after
//     PROPAGS2 of band b+1, which is the last reader of the old spectrum of band b;
//   * otherwise: upload everything, PROPAG_WAM as a whole (it needs the halo), then IMPLSCH / download band by band.
int ecwam_b200_wamintgr_host(ecwam_b200_handle h, const ecwam_b200_fields* host, int with_xllws, long long* h2d_bytes,
                             long long* d2h_bytes) {
  if (!h || !host) EW_FAIL(ECWAM_B200_EINVAL, "null argument");
  const ecwam_b200_params& p = h->par;
  const size_t npts = (size_t)p.nproma * p.nchnk;
  auto chunk_bytes = [&](int kind) -> size_t {
    switch (kind) { case 0: return (size_t)p.nproma * p.nang * p.nfre * 8; case 1: return (size_t)p.nproma * p.nfre * 8; case 2: return (size_t)p.nproma * 8; default: return (size_t)p.nproma * 4; }
  };
  if (!h->mir_alloc) {
    for (const FieldDesc& fd : kFields) {
      void* b = nullptr;
      EW_CUDA_CHECK(cudaMalloc(&b, chunk_bytes(fd.kind) * p.nchnk));
      EW_CUDA_CHECK(cudaMemsetAsync(b, 0, chunk_bytes(fd.kind) * p.nchnk, h->st));
      h->mir_bufs.push_back(b);
      fptr(h->mir, fd.off) = b;
    }
    EW_CUDA_CHECK(cudaStreamCreateWithFlags(&h->st_up, cudaStreamNonBlocking));
    EW_CUDA_CHECK(cudaStreamCreateWithFlags(&h->st_dn, cudaStreamNonBlocking));
    EW_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
    EW_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
    h->mir_alloc = true;
  }
  (void)npts;
  // The host path works on the library's mirrors, the device entry points on the caller's bound tensors: the binding is swapped
  // for the duration of this call only.  The CTU set-up depends on the bound static fields, so a handle that uses both paths
  // rebuilds it whenever it changes sides.
  struct BindGuard {
    H* h; ecwam_b200_fields saved; bool saved_bound;
    ~BindGuard() { h->dev = saved; h->bound = saved_bound; h->pd.omos = saved.omosnh2kd; h->pd.wavn = saved.wavnum; h->cur_side = saved_bound ? 1 : 0; }
  } guard{h, h->dev, h->bound};
  h->dev = h->mir; h->pd.omos = h->mir.omosnh2kd; h->pd.wavn = h->mir.wavnum; h->bound = true; h->cur_side = 2;
  // ---- bands of chunks
  const int nchnk = p.nchnk, P = p.nproma;
  const int reach_chunks = (h->nbr_reach + P - 1) / P + 1;
  static const int kBands = []() { const char* e = getenv("ECWAM_B200_HOST_BANDS"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 24; }();
  int band = std::max((nchnk + kBands - 1) / kBands, reach_chunks);
  const int nband = (nchnk + band - 1) / band;
  const bool substeps = p.ifrelfmax > 0 && p.ifrelfmax < p.nfre_red;
  const bool full = !substeps && !h->weights_dirty && h->weights_side == h->cur_side && h->mir_static_done && nband >= 3 && h->par.irefra < 2;
  while ((int)h->ev_up.size() < nband) { cudaEvent_t e; EW_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->ev_up.push_back(e); }
  while ((int)h->ev_done.size() < nband) { cudaEvent_t e; EW_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->ev_done.push_back(e); }
  long long nin = 0, nout = 0;
  const bool first = !h->mir_static_done;
  auto upload = [&](int c0, int c1, cudaStream_t s) -> int {
    for (const FieldDesc& fd : kFields) {
      const char* src = (const char*)fptr_c(*host, fd.off);
      const bool is_static = (fd.dir & 1) != 0, per_step = (fd.dir & 2) != 0;
      if (!src || !((is_static && first) || per_step)) continue;
      const size_t cb = chunk_bytes(fd.kind);
      EW_CUDA_CHECK(cudaMemcpyAsync((char*)fptr(h->mir, fd.off) + cb * c0, src + cb * c0, cb * (c1 - c0), cudaMemcpyHostToDevice, s));
      nin += (long long)(cb * (c1 - c0));
    }
    return 0;
  };
  auto download = [&](int c0, int c1, cudaStream_t s) -> int {
    for (const FieldDesc& fd : kFields) {
      char* dst = (char*)fptr_c(*host, fd.off);
      if (!dst || !((fd.dir & 4) || ((fd.dir & 8) && with_xllws))) continue;
      const size_t cb = chunk_bytes(fd.kind);
      EW_CUDA_CHECK(cudaMemcpyAsync(dst + cb * c0, (char*)fptr(h->mir, fd.off) + cb * c0, cb * (c1 - c0), cudaMemcpyDeviceToHost, s));
      nout += (long long)(cb * (c1 - c0));
    }
    return 0;
  };
  auto c0_of = [&](int b) { return b * band; };
  auto c1_of = [&](int b) { return std::min(nchnk, (b + 1) * band); };
  // the side streams start after everything already queued on the compute stream
  EW_CUDA_CHECK(cudaEventRecord(h->ev_start, h->st));
  EW_CUDA_CHECK(cudaStreamWaitEvent(h->st_up, h->ev_start, 0));
  EW_CUDA_CHECK(cudaStreamWaitEvent(h->st_dn, h->ev_start, 0));
  int rc = 0;
  if (full) {
    rc = ensure_const(h);
    if (rc) return rc;
    const PropDev& d = h->pd;
    if (h->nproc > 1) {
      // MPEXCHNG needs the spectra of the points the neighbouring ranks read: the chunks that hold them go up first, the halo
      // exchange (pack + NCCL) follows, and the band pipeline below runs as on one rank — PROPAGS2 finds the halo in place.
      if (!h->send_chunks_sorted) {
        std::sort(h->send_chunks.begin(), h->send_chunks.end());
        h->send_chunks.erase(std::unique(h->send_chunks.begin(), h->send_chunks.end()), h->send_chunks.end());
        h->send_chunks_sorted = true;
      }
      const size_t cb4 = chunk_bytes(0);
      const char* src = (const char*)host->fl1;
      for (size_t i = 0; i < h->send_chunks.size();) {      // runs of consecutive chunks: one copy each
        size_t j = i + 1;
        while (j < h->send_chunks.size() && h->send_chunks[j] == h->send_chunks[j - 1] + 1) ++j;
        const int c0 = h->send_chunks[i], n = (int)(j - i);
        EW_CUDA_CHECK(cudaMemcpyAsync((char*)h->mir.fl1 + cb4 * c0, src + cb4 * c0, cb4 * n, cudaMemcpyHostToDevice, h->st_up));
        nin += (long long)(cb4 * n);
        i = j;
      }
      EW_CUDA_CHECK(cudaEventRecord(h->ev_halo, h->st_up));
      EW_CUDA_CHECK(cudaStreamWaitEvent(h->st, h->ev_halo, 0));
      rc = halo_spectrum(h, h->dev.fl1, d.F, d.Fr);
      if (rc) return rc;
    }
    for (int b = 0; b < nband; ++b) {
      if ((rc = upload(c0_of(b), c1_of(b), h->st_up))) return rc;
      EW_CUDA_CHECK(cudaEventRecord(h->ev_up[b], h->st_up));
    }
    for (int b = 0; b <= nband; ++b) {
      if (b < nband) {   // PROPAGS2 of band b reads bands b-1 .. b+1
        EW_CUDA_CHECK(cudaStreamWaitEvent(h->st, h->ev_up[std::min(b + 1, nband - 1)], 0));
        ScopedTimer t(h, "propags2");
        launch_propags2(d, h->dev.fl1, d.F, h->fl3.p, d.Fr, 0, d.Fr, h->msplit, h->st, c0_of(b) * P, c1_of(b) * P);
        h->nlaunch++;
        if (b == nband - 1) launch_pad(d, h->fl3.p, d.Fr, 0, d.Fr, h->st);   // padded lanes of the last chunk (propag_wam.F90:388-398)
      }
      if (b >= 1) {      // IMPLSCH of band b-1, then its way back to the host
        const int bb = b - 1;
        if ((rc = implsch_range(h, c0_of(bb) + 1, c1_of(bb) - c0_of(bb), true))) return rc;
        EW_CUDA_CHECK(cudaEventRecord(h->ev_done[bb], h->st));
        EW_CUDA_CHECK(cudaStreamWaitEvent(h->st_dn, h->ev_done[bb], 0));
        if ((rc = download(c0_of(bb), c1_of(bb), h->st_dn))) return rc;
      }
    }
  } else {
    if ((rc = upload(0, nchnk, h->st))) return rc;
    h->mir_static_done = true;
    bool lf_in_fl3 = true;
    rc = propag_core(h, &lf_in_fl3);
    if (rc) return rc;
    const PropDev& d = h->pd;
    if (!lf_in_fl3) {   // odd number of fast-wave sub-steps left the low frequencies in FL1: finish PROPAG_WAM the plain way
      launch_copyback(d, h->fl3.p, h->dev.fl1, p.ifrelfmax, d.Fr, h->st);
      launch_pad(d, h->dev.fl1, d.F, 0, p.ifrelfmax, h->st);
      h->nlaunch += 2;
    } else launch_pad(d, h->fl3.p, d.Fr, 0, d.Fr, h->st);
    for (int b = 0; b < nband; ++b) {
      if ((rc = implsch_range(h, c0_of(b) + 1, c1_of(b) - c0_of(b), lf_in_fl3))) return rc;
      EW_CUDA_CHECK(cudaEventRecord(h->ev_done[b], h->st));
      EW_CUDA_CHECK(cudaStreamWaitEvent(h->st_dn, h->ev_done[b], 0));
      if ((rc = download(c0_of(b), c1_of(b), h->st_dn))) return rc;
    }
  }
  EW_CUDA_CHECK(cudaStreamSynchronize(h->st_dn));
  EW_CUDA_CHECK(cudaStreamSynchronize(h->st));
  EW_CUDA_CHECK(cudaStreamSynchronize(h->st_up));
  if (h2d_bytes) *h2d_bytes = nin;
  if (d2h_bytes) *d2h_bytes = nout;
  return 0;
}

}  // extern "C"
