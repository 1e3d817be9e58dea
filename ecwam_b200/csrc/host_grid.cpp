// Host-side, one-off grid set-up for callers without ecWAM's module state:
//   sea-point ordering               src/ecwam/mblock.F90:126-135, readmdlconf.F90:107-164
//   MPDECOMP 1-D / 2-D decomposition src/ecwam/mpdecomp.F90:341-686
//   PROPCONNECT neighbour tables     src/ecwam/propconnect.F90:69-430 (indices), :653-971 (WLAT/WCOR)
//   halo send/receive lists          src/ecwam/mpdecomp.F90:731-1176, land slot :1264-1296
// Integer outputs have to be bit-identical to the reference's; the searches the reference does linearly are
// replaced by a (row, column) -> sea-point lookup table and sorts, which give the same unique answers.
#include "../../include/ecwam_b200.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace {

inline long fnint(double x) { return std::lround(x); }

struct RankTables {
  ecwam_b200_decomp d;
  std::vector<int> kxlt, klat, klon, kcor, nfrompe, ntope, nijstart, ijtope;
  std::vector<double> wlat, wcor;
};

struct HostGrid {
  int ngy = 0, ngx = 0, niblo = 0, nproc = 1;
  double amosop = 0, amonop = 0, xdella = 0, xdello = 0;
  std::vector<int> nlon;                       // points per row
  std::vector<double> zdello, cosph, sinph;    // per row
  std::vector<int64_t> rowoff;                 // first cell of each row in the flat mask
  std::vector<int> cell2ij;                    // flat cell -> ORIGINAL sea-point number (1-based) or 0
  std::vector<int> ix0, ky0;                   // (1..niblo) original order: column, row (1-based)
  std::vector<int> ij2new, new2ij;             // (0:niblo)
  std::vector<int> ixn, kyn;                   // relabelled order (index 1..niblo stored at [ij-1])
  std::vector<int> nstart, nend;               // (nproc) 1-based
  std::vector<RankTables> ranks;

  bool sea(int i, int k) const { return cell2ij[rowoff[k - 1] + i - 1] != 0; }
  int newij(int i, int k) const { return ij2new[cell2ij[rowoff[k - 1] + i - 1]]; }
};

// ---- grid geometry (readmdlconf.F90:136-164): the polar rows are clamped at 87.5 degrees
void geometry(HostGrid& g) {
  const double pi = 4.0 * std::atan(1.0), rad = pi / 180.0, XLATMAX = 87.5;
  g.xdella = (g.amonop - g.amosop) / (double)(g.ngy - 1);
  g.ngx = *std::max_element(g.nlon.begin(), g.nlon.end());
  g.xdello = 360.0 / (double)g.ngx;
  g.zdello.resize(g.ngy); g.cosph.resize(g.ngy); g.sinph.resize(g.ngy);
  const double cmin = std::cos(XLATMAX * rad);
  for (int k = 0; k < g.ngy; ++k) {
    const double xlat = (g.amosop + (double)k * g.xdella) * rad;
    g.sinph[k] = std::sin(xlat);
    g.cosph[k] = std::cos(xlat);
    g.zdello[k] = 360.0 / (double)g.nlon[k];
    if (g.cosph[k] <= cmin) { g.cosph[k] = std::cos(XLATMAX * rad); g.sinph[k] = std::sin(XLATMAX * rad); }
  }
}

// ---- decomposition shape (mpdecomp.F90:341-395)
bool decomposition_shape(int npr, bool ll1d, int& nx, int& ny, int& nycut) {
  if (ll1d) { nx = 1; ny = npr; nycut = ny; return true; }
  if (npr == 1) { nx = ny = nycut = 1; return true; }
  if (npr == 2) { nx = 2; ny = 1; nycut = 1; return true; }
  int ic = 0, ip = 0;
  while (ip < npr) { ++ic; ip = 2 * ic * ic; }
  if (ip == npr) { ny = (int)std::sqrt((double)npr / 2.0); nx = 2 * ny; nycut = ny; return true; }
  ny = (int)std::sqrt((double)npr / 2.0) + 1;
  for (nx = 2 * ny; nx >= ny; --nx)
    for (nycut = ny; nycut >= 1; --nycut)
      if (ny * (nx - 1) + nycut == npr) return true;
  return false;
}

// ---- latitude bands of (nearly) equal sea-point count (mpdecomp.F90:397-463)
void latitude_bands(int ijl, int nx, int ny, int nycut, std::vector<int>& s1, std::vector<int>& e1) {
  s1.assign(ny, 0); e1.assign(ny, 0);
  if (nycut == ny) {
    const int nmean = ijl / ny;
    int nrest = ijl - nmean * ny, next = 1;
    for (int b = 0; b < ny; ++b) {
      const int npts = nmean + (nrest > 0 ? 1 : 0);
      if (nrest > 0) --nrest;
      s1[b] = next; e1[b] = next + npts - 1; next += npts;
    }
  } else {
    int nmean = (int)((double)ijl * ((double)nx / (double)((nx - 1) * ny + nycut)));
    int next = 1;
    for (int b = 0; b < nycut; ++b) { s1[b] = next; e1[b] = next + nmean - 1; next += nmean; }
    const int left = ijl - e1[nycut - 1];
    nmean = left / (ny - nycut);
    int nrest = left - nmean * (ny - nycut);
    for (int b = nycut; b < ny; ++b) {
      const int npts = nmean + (nrest > 0 ? 1 : 0);
      if (nrest > 0) --nrest;
      s1[b] = next; e1[b] = next + npts - 1; next += npts;
    }
  }
}

// ---- 2-D split of every band into longitude sectors + relabelling (mpdecomp.F90:484-686)
void split_bands(HostGrid& g, int nx, int ny, int nycut, const std::vector<int>& s1, const std::vector<int>& e1) {
  const int N = g.niblo;
  const double AMOWEP = 0.0, AMOEAP = 360.0 - g.xdello;
  const double xdelloinv = 1.0 / g.xdello;
  double stagger = 0.5 * (AMOEAP - AMOWEP + 1 * g.xdello) / nx;
  stagger = (double)fnint(100 * stagger) / 100.0;
  const int istagger = (int)fnint(stagger * xdelloinv);
  int iproc = 0, nij = 0;   // iproc: 1-based rank being filled
  std::vector<int> order, key;
  for (int b = 1; b <= ny; ++b) {
    ++iproc;
    g.nstart[iproc - 1] = nij + 1;
    const int lo = s1[b - 1], hi = e1[b - 1], ntot = hi - lo + 1;
    const int narea = (b <= nycut) ? nx : nx - 1;
    std::vector<int> quota(narea);
    {
      const int nmean = ntot / narea;
      int nrest = ntot - nmean * narea;
      for (int a = 0; a < narea; ++a) { quota[a] = nmean + (nrest > 0 ? 1 : 0); if (nrest > 0) --nrest; }
    }
    // west->east merge of the band's rows: the reference repeatedly takes, over all rows, the next point with the
    // smallest longitude index (lowest row on ties) == a stable sort on (longitude index, row)
    order.resize(ntot); key.resize(ntot);
    for (int t = 0; t < ntot; ++t) {
      const int ij = lo + t;
      double xlon = AMOWEP + (g.ix0[ij - 1] - 1) * g.zdello[g.ky0[ij - 1] - 1];
      xlon = (double)fnint(100 * xlon) / 100.0;
      key[t] = (int)fnint(xlon * xdelloinv);
      order[t] = t;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int c) {
      if (key[a] != key[c]) return key[a] < key[c];
      return g.ky0[lo + a - 1] < g.ky0[lo + c - 1];
    });
    int jcm = 1;   // even bands start half a sector further east
    if (b % 2 == 0) for (int t = 0; t < ntot; ++t) if (key[t] < istagger) ++jcm;
    int iar = 0, ic = 0;
    auto place = [&](int jc /*1-based*/) {
      ++nij; ++ic;
      if (ic == quota[iar]) g.nend[iproc - 1] = nij;
      else if (ic > quota[iar]) { ic = 1; ++iar; ++iproc; g.nstart[iproc - 1] = nij; }
      const int ij = lo + order[jc - 1];
      g.new2ij[nij] = ij; g.ij2new[ij] = nij;
    };
    for (int jc = jcm; jc <= ntot; ++jc) place(jc);
    for (int jc = 1; jc <= jcm - 1; ++jc) place(jc);
  }
  (void)N;
}

// ---- PROPCONNECT for the own points of one rank, global (relabelled) numbering, 0 = land/outside
void connect(const HostGrid& g, int ijs, int ijl, RankTables& r) {
  const int n = ijl - ijs + 1, NGY = g.ngy;
  r.klat.assign((size_t)n * 4, 0); r.klon.assign((size_t)n * 2, 0); r.kcor.assign((size_t)n * 8, 0);
  r.wlat.assign((size_t)n * 2, 1.0); r.wcor.assign((size_t)n * 4, 1.0);
  auto KLAT = [&](int l, int ic, int icl) -> int& { return r.klat[l + (size_t)n * ((ic - 1) + 2 * (icl - 1))]; };
  auto KLON = [&](int l, int ic) -> int& { return r.klon[l + (size_t)n * (ic - 1)]; };
  auto KCOR = [&](int l, int icr, int icl) -> int& { return r.kcor[l + (size_t)n * ((icr - 1) + 4 * (icl - 1))]; };
  for (int l = 0; l < n; ++l) {
    const int ip = ijs + l, I = g.ixn[ip - 1], K = g.kyn[ip - 1];
    const double zk = g.zdello[K - 1];
    // closest and second closest point on the neighbouring rows (propconnect.F90:69-165)
    for (int side = 1; side <= 2; ++side) {
      const int kn = (side == 1) ? K - 1 : K + 1;
      if (kn < 1 || kn > NGY) continue;
      const double zn = g.zdello[kn - 1];
      const int nl = g.nlon[kn - 1];
      const double xm = (double)(I - 1) * zk / zn;
      const int im = (int)fnint(xm) + 1;
      if (g.sea(im, kn)) KLAT(l, side, 1) = g.newij(im, kn);
      int im2;
      if (xm <= (double)(im - 1)) im2 = (im <= 1) ? 1 : im - 1;
      else im2 = (im >= nl) ? nl : im + 1;
      if (g.sea(im2, kn)) KLAT(l, side, 2) = g.newij(im2, kn);
    }
    // west / east with periodic wrap (propconnect.F90:167-202)
    {
      const int nl = g.nlon[K - 1];
      const int iw = (I > 1) ? I - 1 : nl, ie = (I < nl) ? I + 1 : 1;
      if (g.sea(iw, K)) KLON(l, 1) = g.newij(iw, K);
      if (g.sea(ie, K)) KLON(l, 2) = g.newij(ie, K);
    }
    // corners: 1=NE 2=SE 3=SW 4=NW (propconnect.F90:205-430)
    const double xlon = (double)(I - 1) * zk;
    for (int icr = 1; icr <= 4; ++icr) {
      const bool north = (icr == 1 || icr == 4), east = (icr == 1 || icr == 2);
      const int kn = north ? K + 1 : K - 1;
      if (kn < 1 || kn > NGY) continue;
      const int nl = g.nlon[kn - 1];
      const double xl = east ? xlon + zk : xlon - zk;
      double xm = xl / g.zdello[kn - 1];
      int im = (int)fnint(xm) + 1;
      if (!east && im < 1) { im += nl; xm += (double)nl; }
      if (east && im > nl) { im -= nl; xm -= (double)nl; }
      if (im < 1 || im > nl) continue;
      if (g.sea(im, kn)) KCOR(l, icr, 1) = g.newij(im, kn);
      int im2;
      if (xm <= (double)(im - 1)) im2 = (im <= 1) ? nl : im - 1;
      else im2 = (im >= nl) ? 1 : im + 1;
      if (g.sea(im2, kn)) KCOR(l, icr, 2) = g.newij(im2, kn);
    }
    // interpolation weights of the closest point (propconnect.F90:653-971): overlap of the grid boxes
    const double d0 = (double)(I - 1) * zk, d3 = d0 - 0.5 * zk, d5 = d0 + 0.5 * zk;
    for (int side = 1; side <= 2; ++side) {
      const int kn = (side == 1) ? K - 1 : K + 1;
      if (kn < 1 || kn > NGY) continue;
      const double zn = g.zdello[kn - 1];
      const int im = (int)fnint(d0 / zn) + 1;
      const double xp = (double)(im - 1) * zn, d4 = xp - 0.5 * zn, d6 = xp + 0.5 * zn;
      double w = 1.0;
      if (d0 <= xp) { if (!(d4 <= d3 || d6 <= d5)) { const double d1 = zk - (d4 - d3); w = std::min(1.0, d1 / zk); } }
      else { if (!(d4 >= d3 || d6 >= d5)) { const double d1 = zk - (d5 - d6); w = std::min(1.0, d1 / zk); } }
      r.wlat[l + (size_t)n * (side - 1)] = w;
      for (int ew = 0; ew < 2; ++ew) {
        const int icr = (side == 1) ? (ew == 0 ? 3 : 2) : (ew == 0 ? 4 : 1);
        const double xl = (ew == 0) ? d0 - zk : d0 + zk;
        const double xll = xl - 0.5 * zk, xlr = xl + 0.5 * zk;
        const int im2 = (int)fnint(xl / zn) + 1;
        const double xp2 = (double)(im2 - 1) * zn, xpl = xp2 - 0.5 * zn, xpr = xp2 + 0.5 * zn;
        const double d1 = (xpl > xll && xpr < xlr) ? zk : std::min(xlr, xpr) - std::max(xll, xpl);
        r.wcor[l + (size_t)n * (icr - 1)] = std::min(1.0, d1 / zk);
      }
    }
  }
}

}  // namespace

struct ecwam_b200_host_grid_s { HostGrid g; };

extern "C" {

int ecwam_b200_host_grid_create(int ngy, const int* nlonrgg, double amosop, double amonop, const unsigned char* mask,
                                int nproc, int ll1d, ecwam_b200_host_grid_t* out) {
  if (!nlonrgg || !mask || !out || ngy < 3 || nproc < 1) return ECWAM_B200_EINVAL;
  auto* o = new ecwam_b200_host_grid_s();
  HostGrid& g = o->g;
  g.ngy = ngy; g.amosop = amosop; g.amonop = amonop; g.nproc = nproc;
  g.nlon.assign(nlonrgg, nlonrgg + ngy);
  geometry(g);
  // sea points: rows south -> north, west -> east inside a row
  g.rowoff.resize(ngy + 1);
  g.rowoff[0] = 0;
  for (int k = 0; k < ngy; ++k) g.rowoff[k + 1] = g.rowoff[k] + g.nlon[k];
  g.cell2ij.assign((size_t)g.rowoff[ngy], 0);
  for (int k = 1; k <= ngy; ++k)
    for (int i = 1; i <= g.nlon[k - 1]; ++i)
      if (mask[g.rowoff[k - 1] + i - 1]) {
        g.cell2ij[g.rowoff[k - 1] + i - 1] = ++g.niblo;
        g.ix0.push_back(i); g.ky0.push_back(k);
      }
  const int N = g.niblo;
  if (N < nproc) { delete o; return ECWAM_B200_EINVAL; }
  int nx, ny, nycut;
  if (!decomposition_shape(nproc, ll1d != 0, nx, ny, nycut)) { delete o; return ECWAM_B200_EINVAL; }
  std::vector<int> s1, e1;
  latitude_bands(N, nx, ny, nycut, s1, e1);
  g.nstart.assign(nproc, 0); g.nend.assign(nproc, 0);
  g.ij2new.assign(N + 1, 0); g.new2ij.assign(N + 1, 0);
  if (ll1d || nproc == 1) {
    for (int b = 0; b < ny; ++b) { g.nstart[b] = s1[b]; g.nend[b] = e1[b]; }
    std::iota(g.ij2new.begin(), g.ij2new.end(), 0);
    std::iota(g.new2ij.begin(), g.new2ij.end(), 0);
  } else {
    split_bands(g, nx, ny, nycut, s1, e1);
  }
  g.ixn.resize(N); g.kyn.resize(N);
  for (int nij = 1; nij <= N; ++nij) { g.ixn[nij - 1] = g.ix0[g.new2ij[nij] - 1]; g.kyn[nij - 1] = g.ky0[g.new2ij[nij] - 1]; }

  // ---- per rank: neighbours, then the halo = sorted set of referenced points owned by someone else
  g.ranks.resize(nproc);
  std::vector<std::vector<int>> want(nproc);   // global indices each rank needs, sorted + unique (mpdecomp.F90:764-910)
  for (int ir = 0; ir < nproc; ++ir) {
    RankTables& r = g.ranks[ir];
    const int ijs = g.nstart[ir], ijl = g.nend[ir];
    connect(g, ijs, ijl, r);
    std::vector<int>& w = want[ir];
    auto scan = [&](const std::vector<int>& a) { for (int v : a) if (v > 0 && (v < ijs || v > ijl)) w.push_back(v); };
    scan(r.klon); scan(r.klat); scan(r.kcor);
    if (w.size() > 1) { std::sort(w.begin(), w.end()); w.erase(std::unique(w.begin(), w.end()), w.end()); }
    else w.clear();   // the reference's NH > 1 tests drop a lone halo point (mpdecomp.F90:880-897)
  }
  auto owner = [&](int ij) { return (int)(std::upper_bound(g.nstart.begin(), g.nstart.end(), ij) - g.nstart.begin()) - 1; };
  std::vector<int> klenbot(nproc, 0), klentop(nproc, 0);
  std::vector<std::vector<int>> own(nproc);
  for (int ir = 0; ir < nproc; ++ir) {
    own[ir].resize(want[ir].size());
    for (size_t h = 0; h < want[ir].size(); ++h) {
      const int q = owner(want[ir][h]);
      own[ir][h] = q;
      if (q < ir) ++klenbot[ir]; else if (q > ir) ++klentop[ir];
    }
  }
  for (int ir = 0; ir < nproc; ++ir) {
    RankTables& r = g.ranks[ir];
    ecwam_b200_decomp& d = r.d;
    const int ijs = g.nstart[ir], ijl = g.nend[ir], n = ijl - ijs + 1;
    d.irank = ir + 1; d.nproc = nproc; d.ijs = ijs; d.ijl = ijl;
    d.ninf = ijs - klenbot[ir]; d.nsup = ijl + klentop[ir];
    d.ngy = ngy; d.xdella = g.xdella;
    const int nland = d.nsup + 1;
    // counts and lists (mpdecomp.F90:990-1066): what I receive from q / what q needs from me, in q's sorted order
    r.nfrompe.assign(nproc, 0); r.ntope.assign(nproc, 0); r.nijstart.assign(nproc, nland);
    for (int q : own[ir]) ++r.nfrompe[q];
    for (int q = 0; q < nproc; ++q) for (int o2 : own[q]) if (o2 == ir) ++r.ntope[q];
    d.ntopemax = *std::max_element(r.ntope.begin(), r.ntope.end());
    const int ld = std::max(1, d.ntopemax);
    r.ijtope.assign((size_t)ld * nproc, nland);
    for (int q = 0; q < nproc; ++q) {
      int jh = 0;
      for (size_t h = 0; h < want[q].size(); ++h) if (own[q][h] == ir) r.ijtope[(size_t)ld * q + jh++] = want[q][h];
    }
    d.ntopemax = ld;
    // local slot of every halo point: lower-rank owners below NSTART, higher-rank owners above NEND (:1068-1087)
    std::vector<int> slot(want[ir].size());
    for (size_t h = 0; h < want[ir].size(); ++h) {
      slot[h] = (own[ir][h] < ir) ? d.ninf + (int)h : ijl + (int)h + 1 - klenbot[ir];
      if (h == 0 || own[ir][h] != own[ir][h - 1]) r.nijstart[own[ir][h]] = slot[h];
    }
    // re-address the neighbour tables: halo -> local slot, land/outside -> NLAND (:1089-1156, :1264-1296)
    auto relocate = [&](std::vector<int>& a) {
      for (int& v : a) {
        if (v == 0) { v = nland; continue; }
        if (v >= ijs && v <= ijl) continue;
        auto it = std::lower_bound(want[ir].begin(), want[ir].end(), v);
        if (it != want[ir].end() && *it == v) v = slot[it - want[ir].begin()];
      }
    };
    relocate(r.klon); relocate(r.klat); relocate(r.kcor);
    r.kxlt.assign(g.kyn.begin() + (ijs - 1), g.kyn.begin() + ijl);
    d.zdello = g.zdello.data(); d.cosph = g.cosph.data(); d.sinph = g.sinph.data();
    d.kxlt = r.kxlt.data(); d.klat = r.klat.data(); d.klon = r.klon.data(); d.kcor = r.kcor.data();
    d.wlat = r.wlat.data(); d.wcor = r.wcor.data(); d.nfrompe = r.nfrompe.data(); d.ntope = r.ntope.data();
    d.nijstart = r.nijstart.data(); d.ijtope = r.ijtope.data(); d.land_cgroup = nullptr;
    (void)n;
  }
  *out = o;
  return 0;
}

int ecwam_b200_host_grid_niblo(ecwam_b200_host_grid_t g) { return g ? g->g.niblo : 0; }
const ecwam_b200_decomp* ecwam_b200_host_grid_decomp(ecwam_b200_host_grid_t g, int irank) {
  if (!g || irank < 1 || irank > g->g.nproc) return nullptr;
  return &g->g.ranks[irank - 1].d;
}
const int* ecwam_b200_host_grid_ij2newij(ecwam_b200_host_grid_t g) { return g ? g->g.ij2new.data() : nullptr; }
const int* ecwam_b200_host_grid_newij2ij(ecwam_b200_host_grid_t g) { return g ? g->g.new2ij.data() : nullptr; }
const int* ecwam_b200_host_grid_kxlt(ecwam_b200_host_grid_t g) { return g ? g->g.kyn.data() : nullptr; }
const int* ecwam_b200_host_grid_nstart(ecwam_b200_host_grid_t g) { return g ? g->g.nstart.data() : nullptr; }
const int* ecwam_b200_host_grid_nend(ecwam_b200_host_grid_t g) { return g ? g->g.nend.data() : nullptr; }
int ecwam_b200_host_grid_free(ecwam_b200_host_grid_t g) { delete g; return 0; }

}  // extern "C"
