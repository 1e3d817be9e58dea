// ORACLE (test infrastructure) — C API used by tests/, smoke() and bench.py's CPU-baseline legs through ctypes.
// All per-point arrays cross this API in the ORIGINAL global sea-point order (south->north, west->east,
// mblock.F90:126-135), independent of the emulated decomposition; spectra as [m][k][ij] (ij fastest).
#include "oracle.h"
#include <cstring>
#include <map>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

static ArrD* field2d(Fields& f, const std::string& n) {
  static const std::map<std::string, ArrD Fields::*> mp = {
      {"DEPTH", &Fields::DEPTH}, {"EMAXDPT", &Fields::EMAXDPT}, {"DELLAM1", &Fields::DELLAM1},
      {"COSPHM1", &Fields::COSPHM1}, {"UCUR", &Fields::UCUR}, {"VCUR", &Fields::VCUR}, {"AIRD", &Fields::AIRD},
      {"WDWAVE", &Fields::WDWAVE}, {"CICOVER", &Fields::CICOVER}, {"WSWAVE", &Fields::WSWAVE},
      {"WSTAR", &Fields::WSTAR}, {"USTRA", &Fields::USTRA}, {"VSTRA", &Fields::VSTRA}, {"UFRIC", &Fields::UFRIC},
      {"TAUW", &Fields::TAUW}, {"TAUWDIR", &Fields::TAUWDIR}, {"Z0M", &Fields::Z0M}, {"Z0B", &Fields::Z0B},
      {"CHRNCK", &Fields::CHRNCK}, {"CITHICK", &Fields::CITHICK}, {"WSEMEAN", &Fields::WSEMEAN},
      {"WSFMEAN", &Fields::WSFMEAN}, {"USTOKES", &Fields::USTOKES}, {"VSTOKES", &Fields::VSTOKES},
      {"STRNMS", &Fields::STRNMS}, {"TAUXD", &Fields::TAUXD}, {"TAUYD", &Fields::TAUYD},
      {"TAUOCXD", &Fields::TAUOCXD}, {"TAUOCYD", &Fields::TAUOCYD}, {"TAUOC", &Fields::TAUOC},
      {"TAUICX", &Fields::TAUICX}, {"TAUICY", &Fields::TAUICY}, {"PHIOCD", &Fields::PHIOCD},
      {"PHIEPS", &Fields::PHIEPS}, {"PHIAW", &Fields::PHIAW},
      {"IBRMEM", &Fields::IBRMEM}, {"NSWH", &Fields::NSWH}, {"NMWP", &Fields::NMWP}, {"NPHIEPS", &Fields::NPHIEPS}, {"NTAUOC", &Fields::NTAUOC},
      {"NEMOTAUX", &Fields::NEMOTAUX}, {"NEMOTAUY", &Fields::NEMOTAUY}, {"NEMOTAUICX", &Fields::NEMOTAUICX},
      {"NEMOTAUICY", &Fields::NEMOTAUICY}, {"NEMOWSWAVE", &Fields::NEMOWSWAVE}, {"NEMOPHIF", &Fields::NEMOPHIF},
      {"NEMOUSTOKES", &Fields::NEMOUSTOKES}, {"NEMOVSTOKES", &Fields::NEMOVSTOKES}, {"NEMOSTRN", &Fields::NEMOSTRN}};
  auto it = mp.find(n);
  return it == mp.end() ? nullptr : &(f.*(it->second));
}
static ArrD* field3d(Fields& f, const std::string& n) {
  static const std::map<std::string, ArrD Fields::*> mp = {
      {"WAVNUM", &Fields::WAVNUM}, {"CINV", &Fields::CINV}, {"CGROUP", &Fields::CGROUP}, {"XK2CG", &Fields::XK2CG},
      {"OMOSNH2KD", &Fields::OMOSNH2KD}, {"STOKFAC", &Fields::STOKFAC}, {"CIWA", &Fields::CIWA}};
  auto it = mp.find(n);
  return it == mp.end() ? nullptr : &(f.*(it->second));
}

template <class T>
static long copy_out(const std::vector<T>& v, T* out, long cap) {
  if ((long)v.size() > cap) return -(long)v.size();
  std::memcpy(out, v.data(), v.size() * sizeof(T));
  return (long)v.size();
}

extern "C" {

void* orc_create(const Config* cfg, int ngy, const int* nlonrgg, double amosop, double amonop,
                 const unsigned char* maskflat, const double* depth_sea) {
  try {
    Model* m = new Model();
    m->cfg = *cfg;
#ifdef _OPENMP
    omp_set_num_threads(std::max(1, cfg->nthreads));
#endif
    init_tables(m->cfg, m->tab);
    build_grid(m->cfg, m->grid, ngy, nlonrgg, amosop, amonop, maskflat);
    m->depth0.assign(depth_sea, depth_sea + m->grid.NIBLO);
    mpdecomp(m->cfg, m->tab, m->grid, m->ranks);
    alloc_fields(*m);
    return m;
  } catch (std::exception& e) {
    fprintf(stderr, "orc_create: %s\n", e.what());
    return nullptr;
  }
}
void orc_destroy(void* h) { delete (Model*)h; }
int orc_niblo(void* h) { return ((Model*)h)->grid.NIBLO; }
int orc_config_size() { return (int)sizeof(Config); }

// named scalar / small-table getters ------------------------------------------------------------------
long orc_get_table(void* h, const char* name, double* out, long cap) {
  Model* m = (Model*)h;
  Tables& t = m->tab;
  std::string n(name);
  std::map<std::string, ArrD*> mp = {
      {"FR", &t.FR}, {"DFIM", &t.DFIM}, {"TH", &t.TH}, {"COSTH", &t.COSTH}, {"SINTH", &t.SINTH},
      {"DFIMOFR", &t.DFIMOFR}, {"DFIMFR", &t.DFIMFR}, {"ZPIFR", &t.ZPIFR}, {"FR5", &t.FR5}, {"COFRM4", &t.COFRM4},
      {"FLMAX", &t.FLMAX}, {"RHOWG_DFIM", &t.RHOWG_DFIM}, {"DFIM_SIM", &t.DFIM_SIM}, {"SATWEIGHTS", &t.SATWEIGHTS},
      {"SWELLFT", &t.SWELLFT}, {"WTAUHF", &t.WTAUHF}, {"AF11", &t.AF11}, {"FKLAP", &t.FKLAP}, {"FKLAP1", &t.FKLAP1},
      {"FKLAM", &t.FKLAM}, {"FKLAM1", &t.FKLAM1}, {"FRH", &t.FRH}, {"RNLCOEF", &t.RNLCOEF}, {"FTRF", &t.FTRF},
      {"XK_GC", &t.XK_GC}, {"OMEGA_GC", &t.OMEGA_GC}, {"CM_GC", &t.CM_GC}, {"C2OSQRTVG_GC", &t.C2OSQRTVG_GC},
      {"XKMSQRTVGOC2_GC", &t.XKMSQRTVGOC2_GC}, {"OM3GMKM_GC", &t.OM3GMKM_GC}, {"OMXKM3_GC", &t.OMXKM3_GC},
      {"DELKCC_GC_NS", &t.DELKCC_GC_NS}, {"DELKCC_OMXKM3_GC", &t.DELKCC_OMXKM3_GC}, {"CIDEAC", &t.CIDEAC},
      {"DFIMFR2", &t.DFIMFR2}, {"XKM_GC", &t.XKM_GC}, {"VG_GC", &t.VG_GC}, {"C_GC", &t.C_GC}, {"DELKCC_GC", &t.DELKCC_GC}, {"GOM", &t.GOM}, {"FRM5", &t.FRM5},
      {"ZDELLO", &m->grid.ZDELLO}, {"COSPH", &m->grid.COSPH}, {"SINPH", &m->grid.SINPH}, {"DELLAM", &m->grid.DELLAM}};
  auto it = mp.find(n);
  if (it != mp.end()) return copy_out(it->second->d, out, cap);
  std::map<std::string, double> sc = {
      {"X0TAUHF", t.X0TAUHF}, {"DELTH", t.DELTH}, {"FLOGSPRDM1", t.FLOGSPRDM1}, {"BETAMAXOXKAPPA2", t.BETAMAXOXKAPPA2},
      {"DAL1", t.DAL1}, {"DAL2", t.DAL2}, {"ACL1", t.ACL1}, {"ACL2", t.ACL2}, {"CL11", t.CL11}, {"CL21", t.CL21},
      {"XDELLA", m->grid.XDELLA}, {"R", t.R}, {"ZPI", t.ZPI}, {"TAUWSHELTER", t.TAUWSHELTER}, {"BETAMAX", t.BETAMAX},
      {"ALPHA", t.ALPHA}, {"ALPHAMIN", t.ALPHAMIN}, {"ALPHAMAX", t.ALPHAMAX}, {"ALPHAPMAX", t.ALPHAPMAX}, {"CHNKMIN_U", t.CHNKMIN_U},
      {"ACDLIN", t.ACDLIN}, {"BCDLIN", t.BCDLIN}, {"BMAXOKAP", t.BMAXOKAP}, {"GAMNCONST", t.GAMNCONST}, {"RN1_RN", t.RN1_RN},
      {"DTHRN_A", t.DTHRN_A}, {"DTHRN_U", t.DTHRN_U}, {"ANG_GC_A", t.ANG_GC_A}, {"ANG_GC_B", t.ANG_GC_B}, {"ANG_GC_C", t.ANG_GC_C},
      {"SQRTGOSURFT", t.SQRTGOSURFT}, {"NWAV_GC", (double)t.NWAV_GC}, {"Z0RAT", t.Z0RAT}, {"Z0TUBMAX", t.Z0TUBMAX},
      {"SWELLF4", t.SWELLF4}, {"SWELLF7", t.SWELLF7}, {"CDIS", t.CDIS}, {"DELTA_SDIS", t.DELTA_SDIS}, {"CDISVIS", t.CDISVIS},
      // every other scalar of the module state (tests/golden/make_ref_golden.py hands them to the translated reference source)
      {"G", t.G}, {"GM1", t.GM1}, {"ROWATER", t.ROWATER}, {"ROWATERM1", t.ROWATERM1}, {"ZPI4GM1", t.ZPI4GM1}, {"ZPI4GM2", t.ZPI4GM2},
      {"EPSMIN", t.EPSMIN}, {"EPSUS", t.EPSUS}, {"EPSU10", t.EPSU10}, {"ACD", t.ACD}, {"BCD", t.BCD}, {"CDMAX", t.CDMAX}, {"DKMAX", t.DKMAX},
      {"TAUOCMIN", t.TAUOCMIN}, {"TAUOCMAX", t.TAUOCMAX}, {"PHIEPSMIN", t.PHIEPSMIN}, {"PHIEPSMAX", t.PHIEPSMAX}, {"WSEMEAN_MIN", t.WSEMEAN_MIN},
      {"FRATIO", t.FRATIO}, {"WETAIL", t.WETAIL}, {"FRTAIL", t.FRTAIL}, {"WP1TAIL", t.WP1TAIL}, {"WP2TAIL", t.WP2TAIL},
      {"XKAPPA", t.XKAPPA}, {"XNLEV", t.XNLEV}, {"ZALP", t.ZALP}, {"TAILFACTOR", t.TAILFACTOR}, {"TAILFACTOR_PM", t.TAILFACTOR_PM},
      {"SWELLF", t.SWELLF}, {"SWELLF2", t.SWELLF2}, {"SWELLF3", t.SWELLF3}, {"SWELLF5", t.SWELLF5}, {"SWELLF6", t.SWELLF6},
      {"SWELLF7M1", t.SWELLF7M1}, {"ABMIN", t.ABMIN}, {"ABMAX", t.ABMAX}, {"SDSBR", t.SDSBR}, {"SSDSC2", t.SSDSC2}, {"SSDSC3", t.SSDSC3},
      {"SSDSC4", t.SSDSC4}, {"SSDSC5", t.SSDSC5}, {"SSDSC6", t.SSDSC6}, {"MICHE", t.MICHE}, {"SSDSBRF1", t.SSDSBRF1}, {"BRKPBCOEF", t.BRKPBCOEF},
      {"ISDSDTH", (double)t.ISDSDTH}, {"ISB", (double)t.ISB}, {"IPSAT", (double)t.IPSAT}, {"EGRCRV", t.EGRCRV}, {"AFCRV", t.AFCRV}, {"BFCRV", t.BFCRV},
      {"SURFT", t.SURFT}, {"IAB", (double)t.IAB}, {"EPS1", t.EPS1}, {"JTOT_TAUHF", (double)t.JTOT_TAUHF},
      {"TICMIN", t.TICMIN}, {"HICMIN", t.HICMIN}, {"DTIC", t.DTIC}, {"DHIC", t.DHIC}, {"NICT", (double)t.NICT}, {"NICH", (double)t.NICH}};
  auto is = sc.find(n);
  if (is != sc.end()) { if (cap < 1) return -1; out[0] = is->second; return 1; }
  return 0;
}
long orc_get_itable(void* h, const char* name, int rank, int* out, long cap) {
  Model* m = (Model*)h;
  Tables& t = m->tab;
  RankDecomp& r = m->ranks[rank];
  std::string n(name);
  std::map<std::string, ArrI*> mp = {
      {"INDICESSAT", &t.INDICESSAT}, {"IKP", &t.IKP}, {"IKP1", &t.IKP1}, {"IKM", &t.IKM}, {"IKM1", &t.IKM1},
      {"K1W", &t.K1W}, {"K2W", &t.K2W}, {"K11W", &t.K11W}, {"K21W", &t.K21W}, {"INLCOEF", &t.INLCOEF},
      {"KLAT", &r.KLAT}, {"KLON", &r.KLON}, {"KCOR", &r.KCOR}, {"NFROMPE", &r.NFROMPE}, {"NTOPE", &r.NTOPE},
      {"NIJSTART", &r.NIJSTART}, {"IJTOPE", &r.IJTOPE}, {"NTOPELST", &r.NTOPELST}, {"NFROMPELST", &r.NFROMPELST},
      {"KIJL4CHNK", &r.KIJL4CHNK}, {"IJFROMCHNK", &r.IJFROMCHNK}, {"MPM", &r.MPM}, {"KPM", &r.KPM}, {"JXO", &r.JXO},
      {"JYO", &r.JYO}, {"KCR", &r.KCR}, {"NSTART", &m->grid.NSTART}, {"NEND", &m->grid.NEND},
      {"NEWIJ2IJ", &m->grid.NEWIJ2IJ}, {"IJ2NEWIJ", &m->grid.IJ2NEWIJ}, {"IXLG", &m->grid.IXLG}, {"KXLT", &m->grid.KXLT},
      {"KLENBOT", &m->grid.KLENBOT}, {"KLENTOP", &m->grid.KLENTOP}};
  auto it = mp.find(n);
  if (it != mp.end()) return copy_out(it->second->d, out, cap);
  std::map<std::string, int> sc = {
      {"NINF", r.NINF}, {"NSUP", r.NSUP}, {"IJS", r.IJS}, {"IJL", r.IJL}, {"NPROMA", r.NPROMA}, {"NCHNK", r.NCHNK},
      {"NTOPEMAX", r.NTOPEMAX}, {"NFROMPEMAX", r.NFROMPEMAX}, {"NGBTOPE", r.NGBTOPE}, {"NGBFROMPE", r.NGBFROMPE},
      {"NSDSNTH", t.NSDSNTH}, {"MFRSTLW", t.MFRSTLW}, {"MLSTHG", t.MLSTHG}, {"KFRH", t.KFRH}, {"NFRE_ODD", t.NFRE_ODD},
      {"CFL_FAIL", r.cfl_fail}, {"NIBLO", m->grid.NIBLO}};
  auto is = sc.find(n);
  if (is != sc.end()) { if (cap < 1) return -1; out[0] = is->second; return 1; }
  return 0;
}
long orc_get_rank_double(void* h, const char* name, int rank, double* out, long cap) {
  Model* m = (Model*)h;
  RankDecomp& r = m->ranks[rank];
  std::string n(name);
  std::map<std::string, ArrD*> mp = {{"WLAT", &r.WLAT}, {"WCOR", &r.WCOR}, {"W8", &r.W8}, {"SUMWN", &r.SUMWN},
                                     {"WLATN", &r.WLATN}, {"WLONN", &r.WLONN}, {"WCORN", &r.WCORN}, {"WKPMN", &r.WKPMN}, {"WMPMN", &r.WMPMN}};
  auto it = mp.find(n);
  if (it != mp.end()) return copy_out(it->second->d, out, cap);
  return 0;
}

// per-point fields in original global order ------------------------------------------------------------
static inline void locate(Model* m, int ij0 /*1-based original*/, int& ir, int& ip, int& ic) {
  int nij = m->grid.IJ2NEWIJ(ij0);
  ir = 0;
  while (nij > m->grid.NEND(ir + 1)) ++ir;
  RankDecomp& r = m->ranks[ir];
  int l = nij - r.IJS;
  ic = l / r.NPROMA + 1;
  ip = l % r.NPROMA + 1;
}
// after setting real points, replicate point 1 of the last chunk into its padding lanes (mpdecomp.F90:1443-1456)
static void pad2d(Model* m, ArrD Fields::*mem) {
  for (int ir = 0; ir < m->cfg.npr; ++ir) {
    RankDecomp& r = m->ranks[ir];
    ArrD& a = m->fld[ir].*mem;
    int C = r.NCHNK;
    for (int j = r.KIJL4CHNK(C) + 1; j <= r.NPROMA; ++j) a(j, C) = a(1, C);
  }
}
int orc_set_field(void* h, const char* name, const double* v) {
  Model* m = (Model*)h;
  std::string n(name);
  if (!field2d(m->fld[0], n)) return -1;
  for (int ij = 1; ij <= m->grid.NIBLO; ++ij) {
    int ir, ip, ic; locate(m, ij, ir, ip, ic);
    (*field2d(m->fld[ir], n))(ip, ic) = v[ij - 1];
  }
  for (int ir = 0; ir < m->cfg.npr; ++ir) {
    RankDecomp& r = m->ranks[ir];
    ArrD& a = *field2d(m->fld[ir], n);
    int C = r.NCHNK;
    for (int j = r.KIJL4CHNK(C) + 1; j <= r.NPROMA; ++j) a(j, C) = a(1, C);
  }
  (void)pad2d;
  return 0;
}
// sub-grid obstruction coefficients in the ORIGINAL point order: lon[ic][m][ij] (2, NFRE_RED, NIBLO), lat likewise, cor (4, NFRE_RED, NIBLO)
int orc_set_obstructions(void* h, const double* lon, const double* lat, const double* cor) {
  Model* m = (Model*)h;
  const int FR = m->cfg.nfre_red;
  const long N = m->grid.NIBLO;
  for (int ir = 0; ir < m->cfg.npr; ++ir) {
    RankDecomp& r = m->ranks[ir];
    r.OBSLON.alloc(r.IJS, r.IJL, 1, FR, 1, 2); r.OBSLAT.alloc(r.IJS, r.IJL, 1, FR, 1, 2); r.OBSCOR.alloc(r.IJS, r.IJL, 1, FR, 1, 4);
    r.LUPDTWGHT = true;
  }
  for (int ij = 1; ij <= N; ++ij) {
    int ir, ip, ic; locate(m, ij, ir, ip, ic);
    RankDecomp& r = m->ranks[ir];
    const int IJ = r.IJFROMCHNK(1, ic) + ip - 1;
    for (int M = 1; M <= FR; ++M) {
      for (int a = 1; a <= 2; ++a) {
        r.OBSLON(IJ, M, a) = lon[((a - 1) * (long)FR + (M - 1)) * N + ij - 1];
        r.OBSLAT(IJ, M, a) = lat[((a - 1) * (long)FR + (M - 1)) * N + ij - 1];
      }
      for (int a = 1; a <= 4; ++a) r.OBSCOR(IJ, M, a) = cor[((a - 1) * (long)FR + (M - 1)) * N + ij - 1];
    }
  }
  return 0;
}
int orc_get_field(void* h, const char* name, double* v) {
  Model* m = (Model*)h;
  std::string n(name);
  if (n == "MIJ") {
    for (int ij = 1; ij <= m->grid.NIBLO; ++ij) { int ir, ip, ic; locate(m, ij, ir, ip, ic); v[ij - 1] = m->fld[ir].MIJ(ip, ic); }
    return 0;
  }
  if (!field2d(m->fld[0], n)) return -1;
  for (int ij = 1; ij <= m->grid.NIBLO; ++ij) {
    int ir, ip, ic; locate(m, ij, ir, ip, ic);
    v[ij - 1] = (*field2d(m->fld[ir], n))(ip, ic);
  }
  return 0;
}
// (NFRE, NIBLO) arrays as [m][ij]
int orc_get_field3(void* h, const char* name, double* v) {
  Model* m = (Model*)h;
  std::string n(name);
  if (!field3d(m->fld[0], n)) return -1;
  const long N = m->grid.NIBLO;
  for (int ij = 1; ij <= N; ++ij) {
    int ir, ip, ic; locate(m, ij, ir, ip, ic);
    ArrD& a = *field3d(m->fld[ir], n);
    for (int M = 1; M <= m->cfg.nfre; ++M) v[(M - 1) * N + ij - 1] = a(ip, M, ic);
  }
  return 0;
}
static int spec_io(Model* m, bool xllws, double* v, bool set) {
  const long N = m->grid.NIBLO;
  const int A = m->cfg.nang, F = m->cfg.nfre;
  for (int ij = 1; ij <= N; ++ij) {
    int ir, ip, ic; locate(m, ij, ir, ip, ic);
    ArrD& a = xllws ? m->fld[ir].XLLWS : m->fld[ir].FL1;
    for (int M = 1; M <= F; ++M)
      for (int K = 1; K <= A; ++K) {
        size_t o = ((size_t)(M - 1) * A + (K - 1)) * N + ij - 1;
        if (set) a(ip, K, M, ic) = v[o]; else v[o] = a(ip, K, M, ic);
      }
  }
  if (set)
    for (int ir = 0; ir < m->cfg.npr; ++ir) {
      RankDecomp& r = m->ranks[ir];
      ArrD& a = m->fld[ir].FL1;
      int C = r.NCHNK;
      for (int M = 1; M <= F; ++M)
        for (int K = 1; K <= A; ++K)
          for (int j = r.KIJL4CHNK(C) + 1; j <= r.NPROMA; ++j) a(j, K, M, C) = a(1, K, M, C);
    }
  return 0;
}
int orc_set_fl1(void* h, const double* v) { return spec_io((Model*)h, false, const_cast<double*>(v), true); }
int orc_get_fl1(void* h, double* v) { return spec_io((Model*)h, false, v, false); }
int orc_get_xllws(void* h, double* v) { return spec_io((Model*)h, true, v, false); }

// test hook: keep WNFLUXES' IMPLSCH-internal inputs of the next orc_implsch; read them back in original point order
int orc_capture(void* h, int on) {
  Model* m = (Model*)h;
  for (int ir = 0; ir < m->cfg.npr; ++ir) {
    Fields& f = m->fld[ir];
    RankDecomp& r = m->ranks[ir];
    f.capture = on != 0;
    if (on) {
      f.DBG_SSOURCE.alloc(1, r.NPROMA, 1, m->cfg.nang, 1, m->cfg.nfre, 1, r.NCHNK);
      for (ArrD* a : {&f.DBG_EM, &f.DBG_F1, &f.DBG_PHIWA}) a->alloc(1, r.NPROMA, 1, r.NCHNK);
    }
  }
  return 0;
}
int orc_get_capture(void* h, double* ssource, double* em, double* f1, double* phiwa) {
  Model* m = (Model*)h;
  const long N = m->grid.NIBLO;
  const int A = m->cfg.nang, F = m->cfg.nfre;
  if (!m->fld[0].capture) return -1;
  for (int ij = 1; ij <= N; ++ij) {
    int ir, ip, ic; locate(m, ij, ir, ip, ic);
    Fields& f = m->fld[ir];
    em[ij - 1] = f.DBG_EM(ip, ic); f1[ij - 1] = f.DBG_F1(ip, ic); phiwa[ij - 1] = f.DBG_PHIWA(ip, ic);
    for (int M = 1; M <= F; ++M)
      for (int K = 1; K <= A; ++K) ssource[((size_t)(M - 1) * A + (K - 1)) * N + ij - 1] = f.DBG_SSOURCE(ip, K, M, ic);
  }
  return 0;
}

// the hot path ---------------------------------------------------------------------------------------
int orc_propag(void* h) {
  try { propag_wam(*(Model*)h); } catch (std::exception& e) { fprintf(stderr, "orc_propag: %s\n", e.what()); return -1; }
  int cfl = 0;
  for (auto& r : ((Model*)h)->ranks) cfl += r.cfl_fail;
  return cfl;
}
int orc_implsch(void* h) {
  try { implsch_all(*(Model*)h); } catch (std::exception& e) { fprintf(stderr, "orc_implsch: %s\n", e.what()); return -1; }
  return 0;
}
// Hs = 4 sqrt(EM), mean frequency FM (femean.F90, outblock.F90:223-244) in original order
int orc_get_hs_fm(void* h, double* hs, double* fm) {
  Model* m = (Model*)h;
  const int A = m->cfg.nang, F = m->cfg.nfre;
  std::vector<double> spec((size_t)A * F);
  for (int ij = 1; ij <= m->grid.NIBLO; ++ij) {
    int ir, ip, ic; locate(m, ij, ir, ip, ic);
    ArrD& a = m->fld[ir].FL1;
    for (int M = 1; M <= F; ++M) for (int K = 1; K <= A; ++K) spec[(M - 1) * A + K - 1] = a(ip, K, M, ic);
    double EM, FM;
    femean(m->tab, m->cfg, 1, spec.data(), &EM, &FM);
    hs[ij - 1] = 4.0 * std::sqrt(EM);
    fm[ij - 1] = FM;
  }
  return 0;
}

// SNONLIN source term SL and its functional derivative FLD of the current spectra, [m][k][ij] in original global order
int orc_snonlin(void* h, double* sl, double* fld) {
  Model* m = (Model*)h;
  const long N = m->grid.NIBLO;
  const int A = m->cfg.nang, F = m->cfg.nfre;
  for (int ir = 0; ir < m->cfg.npr; ++ir) {
    RankDecomp& r = m->ranks[ir];
    const int P = r.NPROMA;
    std::vector<double> S((size_t)P * A * F), D((size_t)P * A * F);
    for (int ic = 1; ic <= r.NCHNK; ++ic) {
      snonlin_chunk(m->cfg, m->tab, m->fld[ir], P, ic, S.data(), D.data());
      for (int ip = 1; ip <= r.KIJL4CHNK(ic); ++ip) {
        const int ij0 = m->grid.NEWIJ2IJ(r.IJFROMCHNK(ip, ic));
        for (int M = 1; M <= F; ++M)
          for (int K = 1; K <= A; ++K) {
            const size_t o = ((size_t)(M - 1) * A + (K - 1)) * N + ij0 - 1, q = (ip - 1) + (size_t)P * ((K - 1) + (size_t)A * (M - 1));
            sl[o] = S[q]; fld[o] = D[q];
          }
      }
    }
  }
  return 0;
}

// second-call wind input + STRESSO alone: sl, spos [m][k][ij]; out [3][ij] = TAUW, TAUWDIR, PHIWA
int orc_stresso(void* h, double* sl, double* spos, double* out) {
  Model* m = (Model*)h;
  const long N = m->grid.NIBLO;
  const int A = m->cfg.nang, F = m->cfg.nfre;
  try {
    for (int ir = 0; ir < m->cfg.npr; ++ir) {
      RankDecomp& r = m->ranks[ir];
      const int P = r.NPROMA;
      std::vector<double> S((size_t)P * A * F), D((size_t)P * A * F), O((size_t)3 * P);
      for (int ic = 1; ic <= r.NCHNK; ++ic) {
        stresso_chunk(m->cfg, m->tab, m->fld[ir], P, ic, S.data(), D.data(), O.data());
        for (int ip = 1; ip <= r.KIJL4CHNK(ic); ++ip) {
          const int ij0 = m->grid.NEWIJ2IJ(r.IJFROMCHNK(ip, ic));
          for (int q = 0; q < 3; ++q) out[(size_t)q * N + ij0 - 1] = O[(size_t)q * P + ip - 1];
          for (int M = 1; M <= F; ++M)
            for (int K = 1; K <= A; ++K) {
              const size_t o = ((size_t)(M - 1) * A + (K - 1)) * N + ij0 - 1, q = (ip - 1) + (size_t)P * ((K - 1) + (size_t)A * (M - 1));
              sl[o] = S[q]; spos[o] = D[q];
            }
        }
      }
    }
  } catch (const std::exception& e) { fprintf(stderr, "orc_stresso: %s\n", e.what()); return 1; }
  return 0;
}

// SINPUT (which = 1) or SDISSIP (which = 2) alone, same layout as orc_snonlin
int orc_term(void* h, int which, double* sl, double* fld) {
  Model* m = (Model*)h;
  const long N = m->grid.NIBLO;
  const int A = m->cfg.nang, F = m->cfg.nfre;
  try {
    for (int ir = 0; ir < m->cfg.npr; ++ir) {
      RankDecomp& r = m->ranks[ir];
      const int P = r.NPROMA;
      std::vector<double> S((size_t)P * A * F), D((size_t)P * A * F);
      for (int ic = 1; ic <= r.NCHNK; ++ic) {
        term_chunk(m->cfg, m->tab, m->fld[ir], P, ic, which, S.data(), D.data());
        for (int ip = 1; ip <= r.KIJL4CHNK(ic); ++ip) {
          const int ij0 = m->grid.NEWIJ2IJ(r.IJFROMCHNK(ip, ic));
          for (int M = 1; M <= F; ++M)
            for (int K = 1; K <= A; ++K) {
              const size_t o = ((size_t)(M - 1) * A + (K - 1)) * N + ij0 - 1, q = (ip - 1) + (size_t)P * ((K - 1) + (size_t)A * (M - 1));
              sl[o] = S[q]; fld[o] = D[q];
            }
        }
      }
    }
  } catch (const std::exception& e) { fprintf(stderr, "orc_term: %s\n", e.what()); return 1; }
  return 0;
}

// NEWWIND (newwind.F90) with FF_NEXT given in original global order: wswave, wdwave, aird, wstar, cicover, cithick, ustra, vstra
int orc_newwind(void* h, const double* const* next8) {
  Model* m = (Model*)h;
  std::vector<Fields> nx(m->cfg.npr);
  ArrD Fields::*mem[8] = {&Fields::WSWAVE, &Fields::WDWAVE, &Fields::AIRD, &Fields::WSTAR, &Fields::CICOVER, &Fields::CITHICK,
                          &Fields::USTRA, &Fields::VSTRA};
  for (int ir = 0; ir < m->cfg.npr; ++ir) for (auto mp : mem) (nx[ir].*mp).alloc(1, m->ranks[ir].NPROMA, 1, m->ranks[ir].NCHNK);
  for (int q = 0; q < 8; ++q) {
    for (int ij = 1; ij <= m->grid.NIBLO; ++ij) { int ir, ip, ic; locate(m, ij, ir, ip, ic); (nx[ir].*mem[q])(ip, ic) = next8[q][ij - 1]; }
    for (int ir = 0; ir < m->cfg.npr; ++ir) {
      RankDecomp& r = m->ranks[ir];
      ArrD& a = nx[ir].*mem[q];
      for (int j = r.KIJL4CHNK(r.NCHNK) + 1; j <= r.NPROMA; ++j) a(j, r.NCHNK) = a(1, r.NCHNK);
    }
  }
  for (int ir = 0; ir < m->cfg.npr; ++ir) newwind(*m, ir, nx[ir]);
  return 0;
}
// OUTBS: OUTBLOCK for every chunk of every rank (outbs.F90:97-122); out[(column) * NIBLO + ij] in original global order
int orc_outbs(void* h, int n, const int* itg, const int* icemask, const int* seamask, double zmiss, int llsource, double* out) {
  Model* m = (Model*)h;
  for (int i = 0; i < n; ++i) if (!outparam_supported(itg[i])) return -1;
  OutSel& s = m->sel;
  s.n = n; s.itg.assign(itg, itg + n); s.icemask.assign(icemask, icemask + n); s.seamask.assign(seamask, seamask + n);
  s.zmiss = zmiss; s.llsource = llsource;
  m->bout.assign(m->cfg.npr, {});
  for (int ir = 0; ir < m->cfg.npr; ++ir) {
    RankDecomp& r = m->ranks[ir];
    m->bout[ir].assign((size_t)r.NPROMA * n * r.NCHNK, 0.0);
#pragma omp parallel for schedule(dynamic, 1)
    for (int ic = 1; ic <= r.NCHNK; ++ic)
      outblock(m->cfg, m->tab, m->fld[ir], r.NPROMA, ic, s, m->bout[ir].data() + (size_t)r.NPROMA * n * (ic - 1));
  }
  const long N = m->grid.NIBLO;
  for (int ij = 1; ij <= N; ++ij) {
    int ir, ip, ic; locate(m, ij, ir, ip, ic);
    const RankDecomp& r = m->ranks[ir];
    for (int i = 0; i < n; ++i) out[(size_t)i * N + ij - 1] = m->bout[ir][(ip - 1) + (size_t)r.NPROMA * (i + (size_t)n * (ic - 1))];
  }
  return 0;
}
// OUTWNORM -> MPMINMAXAVG on the BOUT of the last orc_outbs; wnorm (4, NIPRMOUT): average, minimum, maximum, count
int orc_outwnorm(void* h, int global, double* wnorm) {
  Model* m = (Model*)h;
  if (m->bout.empty()) return -1;
  mpminmaxavg(*m, m->sel, m->bout, global != 0, wnorm);
  return 0;
}

}  // extern "C"
