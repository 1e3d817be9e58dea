// ORACLE (test infrastructure) — grid description, PROPCONNECT, MPDECOMP (N emulated ranks), MCHUNK.
#include "oracle.h"

namespace orc {

// readmdlconf.F90:107-164 (sea-point order south->north / west->east; ZDELLO, DELLAM, COSPH clamp at 87.5 deg)
void build_grid(const Config& c, Grid& g, int ngy, const int* nlonrgg, double amosop, double amonop,
                const unsigned char* maskflat) {
  Tables t;  // only for constants
  double PI = 4.0 * std::atan(1.0), RAD = PI / 180.0, CIRC = 40007993.95;
  g.NGY = ngy;
  g.NLONRGG.alloc(1, ngy);
  g.NGX = 0;
  for (int k = 1; k <= ngy; ++k) { g.NLONRGG(k) = nlonrgg[k - 1]; g.NGX = std::max(g.NGX, nlonrgg[k - 1]); }
  g.AMOSOP = amosop; g.AMONOP = amonop; g.AMOWEP = 0.0;
  g.XDELLA = (amonop - amosop) / (double)(ngy - 1);
  g.XDELLO = 360.0 / (double)g.NGX;
  g.AMOEAP = 360.0 - g.XDELLO;
  g.IPER = 1; g.IRGG = 1;
  g.MASK.resize(ngy);
  g.IJMAP.resize(ngy);
  size_t off = 0;
  int ip = 0;
  for (int k = 1; k <= ngy; ++k) {
    int n = nlonrgg[k - 1];
    g.MASK[k - 1].assign(maskflat + off, maskflat + off + n);
    off += n;
    for (int i = 1; i <= n; ++i) if (g.MASK[k - 1][i - 1]) ++ip;
  }
  g.NIBLO = ip;
  g.IXLG0.alloc(1, ip); g.KXLT0.alloc(1, ip);
  ip = 0;
  for (int k = 1; k <= ngy; ++k) {
    int n = nlonrgg[k - 1];
    g.IJMAP[k - 1].assign(n, 0);
    for (int i = 1; i <= n; ++i)
      if (g.MASK[k - 1][i - 1]) { ++ip; g.IXLG0(ip) = i; g.KXLT0(ip) = k; g.IJMAP[k - 1][i - 1] = ip; }
  }
  g.IXLG = g.IXLG0; g.KXLT = g.KXLT0;
  g.ZDELLO.alloc(1, ngy); g.DELLAM.alloc(1, ngy); g.SINPH.alloc(1, ngy); g.COSPH.alloc(1, ngy);
  const double XLATMAX = 87.5;
  for (int K = 1; K <= ngy; ++K) {
    double XLAT = (g.AMOSOP + (double)(K - 1) * g.XDELLA) * RAD;
    g.SINPH(K) = std::sin(XLAT);
    g.COSPH(K) = std::cos(XLAT);
    g.ZDELLO(K) = 360.0 / (double)g.NLONRGG(K);
    g.DELLAM(K) = g.ZDELLO(K) * CIRC / 360.0;
  }
  double COSPHMIN = std::cos(XLATMAX * RAD);
  for (int K = 1; K <= ngy; ++K)
    if (g.COSPH(K) <= COSPHMIN) { g.COSPH(K) = std::cos(XLATMAX * RAD); g.SINPH(K) = std::sin(XLATMAX * RAD); }
  (void)c; (void)t;
}

// The reference finds the 1-D index of grid cell (I,K) by a linear search through BLK2GLO starting at the
// current point (propconnect.F90:84-87 and every similar loop).  The search result is, by construction, the
// unique sea point with those coordinates; the oracle looks it up in IJMAP instead (identical result, O(1)).
static inline int find_ij(const Grid& g, int I, int K) { return g.IJ2NEWIJ.d[g.IJMAP[K - 1][I - 1]]; }
static inline bool sea(const Grid& g, int I, int K) { return g.MASK[K - 1][I - 1] != 0; }

// propconnect.F90:69-430 (KLAT/KLON/KCOR) and :653-971 (WLAT/WCOR), IPROPAGS=2, global (new-IJ) indices.
static void propconnect(const Grid& g, int IJS, int IJL, RankDecomp& r) {
  const ArrI& NLONRGG = g.NLONRGG;
  const ArrD& ZDELLO = g.ZDELLO;
  const int NGY = g.NGY, IPER = g.IPER, IRGG = g.IRGG;
  for (int IP = IJS; IP <= IJL; ++IP) {
    int I = g.IXLG(IP), K = g.KXLT(IP);
    // ---- KLAT (:69-165)
    if (K > 1) {
      double XMIN = (double)(I - 1) * ZDELLO(K) / ZDELLO(K - 1);
      int IMIN = (int)nint(XMIN) + 1;
      if (sea(g, IMIN, K - 1)) r.KLAT(IP, 1, 1) = find_ij(g, IMIN, K - 1);
      if (IRGG == 1) {
        int IMIN2;
        if (XMIN <= (double)(IMIN - 1)) { IMIN2 = (IMIN <= 1) ? 1 : IMIN - 1; }
        else { IMIN2 = (IMIN >= NLONRGG(K - 1)) ? NLONRGG(K - 1) : IMIN + 1; }
        if (sea(g, IMIN2, K - 1)) r.KLAT(IP, 1, 2) = find_ij(g, IMIN2, K - 1);
      } else {
        r.KLAT(IP, 1, 2) = r.KLAT(IP, 1, 1);
      }
    }
    if (K < NGY) {
      double XPLUS = (double)(I - 1) * ZDELLO(K) / ZDELLO(K + 1);
      int IPLUS = (int)nint(XPLUS) + 1;
      if (sea(g, IPLUS, K + 1)) r.KLAT(IP, 2, 1) = find_ij(g, IPLUS, K + 1);
      if (IRGG == 1) {
        int IPLUS2;
        if (XPLUS <= (double)(IPLUS - 1)) { IPLUS2 = (IPLUS <= 1) ? 1 : IPLUS - 1; }
        else { IPLUS2 = (IPLUS >= NLONRGG(K + 1)) ? NLONRGG(K + 1) : IPLUS + 1; }
        if (sea(g, IPLUS2, K + 1)) r.KLAT(IP, 2, 2) = find_ij(g, IPLUS2, K + 1);
      } else {
        r.KLAT(IP, 2, 2) = r.KLAT(IP, 2, 1);
      }
    }
    // ---- KLON (:167-202)
    int IP1D = g.NEWIJ2IJ(IP);
    if (I > 1) {
      if (sea(g, I - 1, K)) r.KLON(IP, 1) = g.IJ2NEWIJ(IP1D - 1);
    } else if (IPER == 1 && sea(g, NLONRGG(K), K)) {
      int kl = IP1D;
      for (int IH = 2; IH <= NLONRGG(K); ++IH) if (sea(g, IH, K)) kl = kl + 1;
      r.KLON(IP, 1) = g.IJ2NEWIJ(kl);
    }
    if (I < NLONRGG(K)) {
      if (sea(g, I + 1, K)) r.KLON(IP, 2) = g.IJ2NEWIJ(IP1D + 1);
    } else if (IPER == 1 && sea(g, 1, K)) {
      int kl = IP1D;
      for (int IH = NLONRGG(K) - 1; IH >= 1; --IH) if (sea(g, IH, K)) kl = kl - 1;
      r.KLON(IP, 2) = g.IJ2NEWIJ(kl);
    }
    // ---- KCOR (:205-430)
    double XLON = (double)(I - 1) * ZDELLO(K);
    for (int side = 0; side < 2; ++side) {  // side 0: K-1 (corners 3=SW, 2=SE); side 1: K+1 (corners 4=NW, 1=NE)
      int KN = side == 0 ? K - 1 : K + 1;
      if (KN < 1 || KN > NGY) continue;
      for (int ew = 0; ew < 2; ++ew) {  // ew 0: west (XLON-ZDELLO), ew 1: east
        int ICR = side == 0 ? (ew == 0 ? 3 : 2) : (ew == 0 ? 4 : 1);
        double XL = ew == 0 ? XLON - ZDELLO(K) : XLON + ZDELLO(K);
        double XM = XL / ZDELLO(KN);
        int IM = (int)nint(XM) + 1;
        bool ok;
        if (ew == 0) {
          if (IPER == 1 && IM < 1) { IM = IM + NLONRGG(KN); XM = XM + (double)NLONRGG(KN); }
          ok = IM >= 1;
        } else {
          if (IPER == 1 && IM > NLONRGG(KN)) { IM = IM - NLONRGG(KN); XM = XM - (double)NLONRGG(KN); }
          ok = IM <= NLONRGG(KN);
        }
        if (!ok) continue;
        if (sea(g, IM, KN)) r.KCOR(IP, ICR, 1) = find_ij(g, IM, KN);
        int IM2;
        if (XM <= (double)(IM - 1)) { IM2 = (IM <= 1) ? NLONRGG(KN) : IM - 1; }
        else { IM2 = (IM >= NLONRGG(KN)) ? 1 : IM + 1; }
        if (sea(g, IM2, KN)) r.KCOR(IP, ICR, 2) = find_ij(g, IM2, KN);
      }
    }
  }
  // ---- weights (:653-971, IRGG=1, IPROPAGS=2)
  for (int IP = IJS; IP <= IJL; ++IP) {
    for (int j = 1; j <= 2; ++j) r.WLAT(IP, j) = 1.0;
    for (int j = 1; j <= 4; ++j) r.WCOR(IP, j) = 1.0;
  }
  if (IRGG == 1) {
    for (int IP = IJS; IP <= IJL; ++IP) {
      int I = g.IXLG(IP), K = g.KXLT(IP);
      double D0 = (double)(I - 1) * ZDELLO(K);
      double D3 = D0 - 0.5 * ZDELLO(K), D5 = D0 + 0.5 * ZDELLO(K);
      for (int side = 0; side < 2; ++side) {
        int KN = side == 0 ? K - 1 : K + 1;
        if (KN < 1 || KN > NGY) continue;
        double XM = D0 / ZDELLO(KN);
        int IM = (int)nint(XM) + 1;
        double XP = (double)(IM - 1) * ZDELLO(KN);
        double D4 = XP - 0.5 * ZDELLO(KN), D6 = XP + 0.5 * ZDELLO(KN);
        double w;
        if (D0 <= XP) {
          if (D4 <= D3 || D6 <= D5) w = 1.0;
          else { double D2 = D4 - D3; double D1 = ZDELLO(K) - D2; w = std::min(1.0, D1 / ZDELLO(K)); }
        } else {
          if (D4 >= D3 || D6 >= D5) w = 1.0;
          else { double D2 = D5 - D6; double D1 = ZDELLO(K) - D2; w = std::min(1.0, D1 / ZDELLO(K)); }
        }
        r.WLAT(IP, side + 1) = w;
        for (int ew = 0; ew < 2; ++ew) {
          int ICR = side == 0 ? (ew == 0 ? 3 : 2) : (ew == 0 ? 4 : 1);
          double XL = ew == 0 ? D0 - ZDELLO(K) : D0 + ZDELLO(K);
          double XLL = XL - 0.5 * ZDELLO(K), XLR = XL + 0.5 * ZDELLO(K);
          double XM2 = XL / ZDELLO(KN);
          int IM2 = (int)nint(XM2) + 1;
          double XP2 = (double)(IM2 - 1) * ZDELLO(KN);
          double XPL = XP2 - 0.5 * ZDELLO(KN), XPR = XP2 + 0.5 * ZDELLO(KN);
          double D1;
          if (XPL > XLL && XPR < XLR) D1 = ZDELLO(K);
          else D1 = std::min(XLR, XPR) - std::max(XLL, XPL);
          r.WCOR(IP, ICR) = std::min(1.0, D1 / ZDELLO(K));
        }
      }
    }
  }
}

// mchunk.F90:33-75
static void mchunk(RankDecomp& r) {
  int IJS = r.IJS, IJL = r.IJL, P = r.NPROMA;
  r.NCHNK = (IJL - IJS + 1) / P;
  if (r.NCHNK * P <= (IJL - IJS)) r.NCHNK += 1;
  r.KIJL4CHNK.alloc(1, r.NCHNK);
  for (int ICHNK = 1; ICHNK <= r.NCHNK; ++ICHNK)
    r.KIJL4CHNK(ICHNK) = std::min(P, IJL - IJS + 1 - (ICHNK - 1) * P);
  r.IJFROMCHNK.alloc(1, P, 1, r.NCHNK);
  for (int ICHNK = 1; ICHNK <= r.NCHNK; ++ICHNK) {
    for (int IPRM = 1; IPRM <= r.KIJL4CHNK(ICHNK); ++IPRM) r.IJFROMCHNK(IPRM, ICHNK) = IJS + IPRM - 1 + (ICHNK - 1) * P;
    for (int IPRM = r.KIJL4CHNK(ICHNK) + 1; IPRM <= P; ++IPRM) r.IJFROMCHNK(IPRM, ICHNK) = 0;
  }
}

// mpdecomp.F90:341-1296 (+ :1343-1356 chunking).  All NPR ranks are built in-process; the two
// MPL_ALLGATHERV calls (:914,:924) become plain loops over ranks.
void mpdecomp(const Config& c, const Tables& t, Grid& g, std::vector<RankDecomp>& ranks) {
  (void)t;
  const int NPR = c.npr, NIBLO = g.NIBLO, IJL = NIBLO;
  g.NPR = NPR;
  g.NSTART.alloc(1, NPR); g.NEND.alloc(1, NPR); g.KLENBOT.alloc(1, NPR); g.KLENTOP.alloc(1, NPR);
  int NXDECOMP, NYDECOMP, NYCUT;
  // :341-395
  if (c.ll1d) { NXDECOMP = 1; NYDECOMP = NPR; NYCUT = NYDECOMP; }
  else if (NPR == 1) { NXDECOMP = 1; NYDECOMP = 1; NYCUT = 1; }
  else if (NPR == 2) { NXDECOMP = 2; NYDECOMP = 1; NYCUT = 1; }
  else {
    int IPROC = 0, ICOUNT = 0;
    while (IPROC < NPR) { ICOUNT++; IPROC = 2 * ICOUNT * ICOUNT; }
    if (IPROC == NPR) {
      NYDECOMP = (int)std::sqrt((double)NPR / 2.0);
      NXDECOMP = 2 * NYDECOMP;
      NYCUT = NYDECOMP;
    } else {
      IPROC = 0;
      NYDECOMP = (int)std::sqrt((double)NPR / 2.0) + 1;
      NXDECOMP = 0; NYCUT = 0;
      for (NXDECOMP = 2 * NYDECOMP; NXDECOMP >= NYDECOMP; --NXDECOMP) {
        for (NYCUT = NYDECOMP; NYCUT >= 1; --NYCUT) {
          IPROC = NYDECOMP * (NXDECOMP - 1) + NYCUT;
          if (IPROC == NPR) break;
        }
        if (IPROC == NPR) break;
      }
      if (IPROC != NPR) throw std::runtime_error("MPDECOMP: decomposition failed");
    }
  }
  // :397-463 latitude bands
  ArrI NSTART1D, NEND1D;
  NSTART1D.alloc(1, NYDECOMP); NEND1D.alloc(1, NYDECOMP);
  if (NYCUT == NYDECOMP) {
    int NMEAN = IJL / NYDECOMP, NREST = IJL - NMEAN * NYDECOMP, NPTS;
    NSTART1D(1) = 1;
    if (NREST > 0) { NPTS = NMEAN + 1; NREST--; } else NPTS = NMEAN;
    NEND1D(1) = NSTART1D(1) + NPTS - 1;
    for (int IP = 2; IP <= NYDECOMP; ++IP) {
      NSTART1D(IP) = NSTART1D(IP - 1) + NPTS;
      if (NREST > 0) { NPTS = NMEAN + 1; NREST--; } else NPTS = NMEAN;
      NEND1D(IP) = NSTART1D(IP) + NPTS - 1;
    }
  } else {
    int NMEAN = (int)((double)IJL * ((double)NXDECOMP / (double)((NXDECOMP - 1) * NYDECOMP + NYCUT)));
    NSTART1D(1) = 1;
    int NPTS = NMEAN;
    NEND1D(1) = NSTART1D(1) + NPTS - 1;
    for (int IP = 2; IP <= NYCUT; ++IP) { NSTART1D(IP) = NSTART1D(IP - 1) + NPTS; NPTS = NMEAN; NEND1D(IP) = NSTART1D(IP) + NPTS - 1; }
    NMEAN = (IJL - NEND1D(NYCUT)) / (NYDECOMP - NYCUT);
    int NREST = (IJL - NEND1D(NYCUT)) - NMEAN * (NYDECOMP - NYCUT);
    for (int IP = NYCUT + 1; IP <= NYDECOMP; ++IP) {
      NSTART1D(IP) = NSTART1D(IP - 1) + NPTS;
      if (NREST > 0) { NPTS = NMEAN + 1; NREST--; } else NPTS = NMEAN;
      NEND1D(IP) = NSTART1D(IP) + NPTS - 1;
    }
  }
  g.NEWIJ2IJ.alloc(0, NIBLO); g.IJ2NEWIJ.alloc(0, NIBLO);
  if (c.ll1d || NPR == 1) {
    for (int IP = 1; IP <= NYDECOMP; ++IP) { g.NSTART(IP) = NSTART1D(IP); g.NEND(IP) = NEND1D(IP); }
    for (int IJ = 0; IJ <= NIBLO; ++IJ) { g.NEWIJ2IJ(IJ) = IJ; g.IJ2NEWIJ(IJ) = IJ; }
  } else {
    // :465-686 2-D split + relabel
    g.NEWIJ2IJ(0) = 0; g.IJ2NEWIJ(0) = 0;
    double XDELLOINV = 1.0 / g.XDELLO;
    double STAGGER = 0.5 * (g.AMOEAP - g.AMOWEP + g.IPER * g.XDELLO) / NXDECOMP;
    STAGGER = (double)nint(100 * STAGGER) / 100.0;
    int ISTAGGER = (int)nint(STAGGER * XDELLOINV);
    int IPROC = 0, NIJ = 0;
    const ArrI& KXLT = g.KXLT0; const ArrI& IXLG = g.IXLG0;
    for (int IPR = 1; IPR <= NYDECOMP; ++IPR) {
      IPROC++;
      g.NSTART(IPROC) = NIJ + 1;
      int NTOT = NEND1D(IPR) - NSTART1D(IPR) + 1;
      int NAREA = (IPR <= NYCUT) ? NXDECOMP : NXDECOMP - 1;
      ArrI NTOTSUB; NTOTSUB.alloc(1, NAREA);
      int NMEAN = NTOT / NAREA, NREST = NTOT - NMEAN * NAREA;
      for (int IAR = 1; IAR <= NAREA; ++IAR) { if (NREST > 0) { NTOTSUB(IAR) = NMEAN + 1; NREST--; } else NTOTSUB(IAR) = NMEAN; }
      int KLATBOT = KXLT(NSTART1D(IPR)), KLATTOP = KXLT(NEND1D(IPR));
      ArrI KSTART1, KEND1, NLON, ILON;
      KSTART1.alloc(KLATBOT, KLATTOP); KEND1.alloc(KLATBOT, KLATTOP); NLON.alloc(KLATBOT, KLATTOP); ILON.alloc(KLATBOT, KLATTOP);
      int KXLAT = KLATBOT;
      KSTART1(KXLAT) = NSTART1D(IPR);
      for (int IJ = NSTART1D(IPR) + 1; IJ <= NEND1D(IPR); ++IJ) {
        if (KXLAT < KXLT(IJ)) { KXLAT = KXLT(IJ); KSTART1(KXLAT) = IJ; KEND1(KXLAT - 1) = IJ - 1; }
      }
      KEND1(KLATTOP) = NEND1D(IPR);
      int NLONGMAX = 0;
      for (KXLAT = KLATBOT; KXLAT <= KLATTOP; ++KXLAT) NLONGMAX = std::max(KEND1(KXLAT) - KSTART1(KXLAT) + 1, NLONGMAX);
      ArrI IXLON; IXLON.alloc(1, std::max(1, NLONGMAX), KLATBOT, KLATTOP);
      int IXLONMAX = (int)(g.AMOWEP * XDELLOINV) - 1;
      for (KXLAT = KLATBOT; KXLAT <= KLATTOP; ++KXLAT) NLON(KXLAT) = 0;
      KXLAT = KLATBOT;
      for (int IJ = NSTART1D(IPR); IJ <= NEND1D(IPR); ++IJ) {
        if (KXLAT < KXLT(IJ)) KXLAT = KXLT(IJ);
        NLON(KXLAT) = NLON(KXLAT) + 1;
        int IX = IXLG(IJ), JSN = KXLT(IJ);
        double XLON = g.AMOWEP + (IX - 1) * g.ZDELLO(JSN);
        XLON = (double)nint(100 * XLON) / 100.0;
        IXLON(NLON(KXLAT), KXLAT) = (int)nint(XLON * XDELLOINV);
        IXLONMAX = std::max(IXLONMAX, IXLON(NLON(KXLAT), KXLAT));
      }
      ArrI IJNDEX; IJNDEX.alloc(1, NTOT);
      for (KXLAT = KLATBOT; KXLAT <= KLATTOP; ++KXLAT) ILON(KXLAT) = 1;
      int JC = 0, KMIN = KLATBOT;
      while (KMIN > 0) {
        int IXLONMIN = IXLONMAX + 1;
        KMIN = 0;
        for (KXLAT = KLATBOT; KXLAT <= KLATTOP; ++KXLAT) {
          if (ILON(KXLAT) <= NLON(KXLAT)) {
            if (IXLON(ILON(KXLAT), KXLAT) < IXLONMIN) { KMIN = KXLAT; IXLONMIN = IXLON(ILON(KXLAT), KXLAT); }
          }
        }
        if (KMIN > 0) { int IJ = KSTART1(KMIN) + ILON(KMIN) - 1; JC++; IJNDEX(JC) = IJ; ILON(KMIN)++; }
      }
      int JCS = 1, JCM;
      if (IPR % 2 == 0) {
        JCM = 1;
        for (KXLAT = KLATBOT; KXLAT <= KLATTOP; ++KXLAT) {
          int IIL = 1;
          while (NLON(KXLAT) > 0 && IIL <= NLON(KXLAT) && IXLON(std::min(IIL, NLON(KXLAT)), KXLAT) < ISTAGGER) { IIL++; JCM++; }
        }
      } else JCM = 1;
      int IAR = 1, IC = 0;
      auto place = [&](int jc) {
        NIJ++; IC++;
        if (IC == NTOTSUB(IAR)) g.NEND(IPROC) = NIJ;
        else if (IC > NTOTSUB(IAR)) { IC = 1; IAR++; IPROC++; g.NSTART(IPROC) = NIJ; }
        int IJ = IJNDEX(jc);
        g.NEWIJ2IJ(NIJ) = IJ; g.IJ2NEWIJ(IJ) = NIJ;
      };
      for (int jc = JCM; jc <= NTOT; ++jc) place(jc);
      for (int jc = JCS; jc <= JCM - 1; ++jc) place(jc);
      // NB (:646-650): when NTOTSUB(IAR)==1 the reference's IF/ELSEIF sets NEND only on IC==NTOTSUB; keep as is.
    }
    for (int N = 1; N <= NIBLO; ++N) { g.IXLG(N) = g.IXLG0(g.NEWIJ2IJ(N)); g.KXLT(N) = g.KXLT0(g.NEWIJ2IJ(N)); }
  }

  // ---- per rank: PROPCONNECT + halo lists (:690-910)
  ranks.assign(NPR, RankDecomp());
  int MAXLEN = 0;
  for (int IP = 1; IP <= NPR; ++IP) MAXLEN = std::max(MAXLEN, g.NEND(IP) - g.NSTART(IP) + 1);
  std::vector<std::vector<int>> IJFROMPE(NPR + 1);
  ArrI NLENHALO; NLENHALO.alloc(1, NPR);
  for (int IR = 1; IR <= NPR; ++IR) {
    RankDecomp& r = ranks[IR - 1];
    r.IRANK = IR; r.IJS = g.NSTART(IR); r.IJL = g.NEND(IR);
    r.KLAT.alloc(r.IJS, r.IJL, 1, 2, 1, 2); r.KLON.alloc(r.IJS, r.IJL, 1, 2); r.KCOR.alloc(r.IJS, r.IJL, 1, 4, 1, 2);
    r.WLAT.alloc(r.IJS, r.IJL, 1, 2); r.WCOR.alloc(r.IJS, r.IJL, 1, 4);
    propconnect(g, r.IJS, r.IJL, r);
    std::vector<int> ITEMP;
    auto consider = [&](int v) { if (v > 0 && v <= NIBLO && (v < r.IJS || v > r.IJL)) ITEMP.push_back(v); };
    for (int IC = 1; IC <= 2; ++IC) for (int IJ = r.IJS; IJ <= r.IJL; ++IJ) consider(r.KLON(IJ, IC));
    for (int ICL = 1; ICL <= 2; ++ICL) for (int IC = 1; IC <= 2; ++IC) for (int IJ = r.IJS; IJ <= r.IJL; ++IJ) consider(r.KLAT(IJ, IC, ICL));
    for (int ICL = 1; ICL <= 2; ++ICL) for (int ICR = 1; ICR <= 4; ++ICR) for (int IJ = r.IJS; IJ <= r.IJL; ++IJ) consider(r.KCOR(IJ, ICR, ICL));
    int NH = (int)ITEMP.size();
    std::vector<int>& out = IJFROMPE[IR];
    if (NH > 1) {  // (:880-897: a single halo point is dropped by the reference's NH>1 tests)
      std::sort(ITEMP.begin(), ITEMP.end());
      out.push_back(ITEMP[0]);
      for (int IH = 1; IH < NH; ++IH) if (ITEMP[IH] > ITEMP[IH - 1]) out.push_back(ITEMP[IH]);
    }
    NLENHALO(IR) = (int)out.size();
  }
  // :966-988 IPROCFROM, KLENBOT/KLENTOP
  std::vector<std::vector<int>> IPROCFROM(NPR + 1);
  for (int IP = 1; IP <= NPR; ++IP) {
    g.KLENBOT(IP) = 0; g.KLENTOP(IP) = 0;
    IPROCFROM[IP].assign(NLENHALO(IP), NPR + 1);
    for (int IH = 1; IH <= NLENHALO(IP); ++IH) {
      for (int IPROC = 1; IPROC <= NPR; ++IPROC) {
        if (IJFROMPE[IP][IH - 1] >= g.NSTART(IPROC) && IJFROMPE[IP][IH - 1] <= g.NEND(IPROC)) {
          IPROCFROM[IP][IH - 1] = IPROC;
          if (IPROC < IP) g.KLENBOT(IP)++; else if (IPROC > IP) g.KLENTOP(IP)++;
          break;
        }
      }
    }
  }
  for (int IR = 1; IR <= NPR; ++IR) {
    RankDecomp& r = ranks[IR - 1];
    const int IRANK = IR;
    r.NINF = g.NSTART(IRANK) - g.KLENBOT(IRANK);
    r.NSUP = g.NEND(IRANK) + g.KLENTOP(IRANK);
    const int NLAND = r.NSUP + 1;
    r.NTOPE.alloc(1, NPR); r.NFROMPE.alloc(1, NPR); r.NIJSTART.alloc(1, NPR);
    // :990-1013
    for (int IP = 1; IP <= NPR; ++IP) {
      r.NTOPE(IP) = 0;
      for (int IH = 1; IH <= NLENHALO(IP); ++IH) if (IPROCFROM[IP][IH - 1] == IRANK) r.NTOPE(IP)++;
    }
    r.NTOPEMAX = 0;
    for (int IP = 1; IP <= NPR; ++IP) r.NTOPEMAX = std::max(r.NTOPEMAX, r.NTOPE(IP));
    for (int IP = 1; IP <= NPR; ++IP) r.NFROMPE(IP) = 0;
    for (int IH = 1; IH <= NLENHALO(IRANK); ++IH) r.NFROMPE(IPROCFROM[IRANK][IH - 1])++;
    r.NFROMPEMAX = 0;
    for (int IP = 1; IP <= NPR; ++IP) r.NFROMPEMAX = std::max(r.NFROMPEMAX, r.NFROMPE(IP));
    // :1015-1047 neighbour lists
    r.NGBTOPE = 0;
    for (int IP = 1; IP <= NPR; ++IP) if (r.NTOPE(IP) > 0) r.NGBTOPE++;
    r.NTOPELST.alloc(1, std::max(1, r.NGBTOPE));
    int INBNGH = 0;
    for (int IP = 1; IP <= NPR; ++IP) if (r.NTOPE(IP) > 0) r.NTOPELST(++INBNGH) = IP;
    r.NGBFROMPE = 0;
    for (int IP = 1; IP <= NPR; ++IP) if (r.NFROMPE(IP) > 0) r.NGBFROMPE++;
    r.NFROMPELST.alloc(1, std::max(1, r.NGBFROMPE));
    INBNGH = 0;
    for (int IP = 1; IP <= NPR; ++IP) if (r.NFROMPE(IP) > 0) r.NFROMPELST(++INBNGH) = IP;
    // :1049-1066 IJTOPE
    r.IJTOPE.alloc(1, std::max(1, r.NTOPEMAX), 1, NPR);
    for (int IP = 1; IP <= NPR; ++IP) for (int JH = 1; JH <= r.NTOPEMAX; ++JH) r.IJTOPE(JH, IP) = NLAND;
    for (int IP = 1; IP <= NPR; ++IP) {
      int JH = 0;
      for (int IH = 1; IH <= NLENHALO(IP); ++IH) if (IPROCFROM[IP][IH - 1] == IRANK) r.IJTOPE(++JH, IP) = IJFROMPE[IP][IH - 1];
    }
    // :1068-1087 IJHALO
    std::vector<int> IJHALO(std::max(1, NLENHALO(IRANK)));
    for (int IH = 1; IH <= NLENHALO(IRANK); ++IH) {
      if (IPROCFROM[IRANK][IH - 1] < IRANK) IJHALO[IH - 1] = r.NINF + IH - 1;
      else if (IPROCFROM[IRANK][IH - 1] > IRANK) IJHALO[IH - 1] = g.NEND(IRANK) + IH - g.KLENBOT(IRANK);
    }
    // :1089-1156 local re-addressing (binary search == the reference's first-match linear search on a sorted,
    // de-duplicated list)
    if (NPR > 1) {
      const std::vector<int>& lst = IJFROMPE[IRANK];
      auto remap = [&](int& v) {
        auto it = std::lower_bound(lst.begin(), lst.end(), v);
        if (it != lst.end() && *it == v) v = IJHALO[it - lst.begin()];
      };
      for (int IC = 1; IC <= 2; ++IC) for (int IJ = r.IJS; IJ <= r.IJL; ++IJ) remap(r.KLON(IJ, IC));
      for (int ICL = 1; ICL <= 2; ++ICL) for (int IC = 1; IC <= 2; ++IC) for (int IJ = r.IJS; IJ <= r.IJL; ++IJ) remap(r.KLAT(IJ, IC, ICL));
      for (int ICL = 1; ICL <= 2; ++ICL) for (int ICR = 1; ICR <= 4; ++ICR) for (int IJ = r.IJS; IJ <= r.IJL; ++IJ) remap(r.KCOR(IJ, ICR, ICL));
    }
    // :1158-1176 NIJSTART
    for (int IP = 1; IP <= NPR; ++IP) r.NIJSTART(IP) = NLAND;
    if (NPR > 1 && NLENHALO(IRANK) > 0) {
      const std::vector<int>& pf = IPROCFROM[IRANK];
      if (pf[0] < IRANK) r.NIJSTART(pf[0]) = r.NINF; else if (pf[0] > IRANK) r.NIJSTART(pf[0]) = g.NEND(IRANK) + 1;
      for (int IH = 2; IH <= NLENHALO(IRANK); ++IH) {
        if (pf[IH - 1] != pf[IH - 2]) {
          if (pf[IH - 1] < IRANK) r.NIJSTART(pf[IH - 1]) = r.NINF + IH - 1;
          else if (pf[IH - 1] > IRANK) r.NIJSTART(pf[IH - 1]) = g.NEND(IRANK) + IH - g.KLENBOT(IRANK);
        }
      }
    }
    // :1264-1296 land -> NLAND
    for (size_t i = 0; i < r.KLAT.d.size(); ++i) if (r.KLAT.d[i] == 0) r.KLAT.d[i] = NLAND;
    for (size_t i = 0; i < r.KLON.d.size(); ++i) if (r.KLON.d[i] == 0) r.KLON.d[i] = NLAND;
    for (size_t i = 0; i < r.KCOR.d.size(); ++i) if (r.KCOR.d[i] == 0) r.KCOR.d[i] = NLAND;
    // :1343-1356 chunking (WAM_NPROMA's thread re-balancing is a CPU-threading detail and is not applied:
    // LLNO_WAM_NPROMA=.TRUE. path)
    r.NPROMA = std::min(c.nproma, r.IJL - r.IJS + 1);
    mchunk(r);
  }
}

}  // namespace orc
