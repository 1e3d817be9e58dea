// Restart files and the grid-table file of the reference, host side:
//   BLS  spectrum restart   SAVSPEC / WRITEFL / READFL          (savspec.F90:86-166, writefl.F90:86-120, readfl.F90:118-145)
//   LAW  stress restart     SAVSTRESS / WRITESTRESS / READSTRESS (savstress.F90:80-152, writestress.F90:76-109, readstress.F90:97-124)
//   wam_grid_tables         OUTCOM / READPRE                     (outcom.F90:139-144, readpre.F90:186-215, 262-345)
//   file names              GRSTNAME, EXPAND_STRING              (grstname.F90:88-142, expand_string.F90:108-180)
// All of them are Fortran unformatted sequential files: every record is  [int32 n][n bytes][int32 n]  (little endian);
// a record longer than 2147483639 bytes is a chain of sub-records whose leading marker is negative when another
// sub-record follows and whose trailing marker is negative when one precedes (the gfortran convention).
//
// The global BLS / LAW files hold fixed-size records in the ORIGINAL sea-point order (the reference gathers the
// spectrum on one task, MPGATHERFL, and writes FL(IJ2NEWIJ(IJ),K,M)).  Here every rank writes its own points straight
// into their places of the shared file with pwrite (no gather, any number of ranks at once); one rank sizes the file
// and writes the record markers first.  The result is byte-identical to the single-task file.
#include "../../include/ecwam_b200.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

void ew_set_error(const char* fmt, ...);
#define IO_FAIL(...) do { ew_set_error(__VA_ARGS__); return ECWAM_B200_EIO; } while (0)

namespace {

long long g_maxsub = 2147483639LL;   // gfortran's maximum sub-record length (tests lower it)

struct Fd {
  int fd = -1;
  ~Fd() { if (fd >= 0) ::close(fd); }
};
int pwrite_all(int fd, const void* buf, size_t n, long long off) {
  const char* p = (const char*)buf;
  while (n) {
    const ssize_t w = ::pwrite(fd, p, n, (off_t)off);
    if (w < 0) { if (errno == EINTR) continue; return -1; }
    p += w; n -= (size_t)w; off += w;
  }
  return 0;
}
int pread_all(int fd, void* buf, size_t n, long long off) {
  char* p = (char*)buf;
  while (n) {
    const ssize_t r = ::pread(fd, p, n, (off_t)off);
    if (r < 0) { if (errno == EINTR) continue; return -1; }
    if (r == 0) return -2;   // short file
    p += r; n -= (size_t)r; off += r;
  }
  return 0;
}
// bytes a record of n data bytes occupies on disk, markers included
long long record_span(long long n) {
  if (n <= g_maxsub) return n + 8;
  const long long nsub = (n + g_maxsub - 1) / g_maxsub;
  return n + 8 * nsub;
}
// write one record at `off`; returns the offset behind it (or -1)
long long write_record(int fd, long long off, const void* data, long long n) {
  const char* p = (const char*)data;
  long long left = n;
  bool first = true;
  do {
    const long long len = left > g_maxsub ? g_maxsub : left;
    const bool more = left > len;
    const int32_t head = (int32_t)(more ? -len : len), tail = (int32_t)(first ? len : -len);
    if (pwrite_all(fd, &head, 4, off) || pwrite_all(fd, p, (size_t)len, off + 4) || pwrite_all(fd, &tail, 4, off + 4 + len)) return -1;
    off += len + 8; p += len; left -= len; first = false;
  } while (left > 0);
  return off;
}
// read one record of exactly n data bytes at `off`; returns the offset behind it, -1 I/O error, -2 wrong length / short file
long long read_record(int fd, long long off, void* data, long long n) {
  char* p = (char*)data;
  long long got = 0;
  for (;;) {
    int32_t head, tail;
    int rc = pread_all(fd, &head, 4, off);
    if (rc) return rc;
    const long long len = head < 0 ? -(long long)head : head;
    if (got + len > n) return -2;
    if ((rc = pread_all(fd, p, (size_t)len, off + 4))) return rc;
    if ((rc = pread_all(fd, &tail, 4, off + 4 + len))) return rc;
    if ((tail < 0 ? -(long long)tail : tail) != len) return -2;
    off += len + 8; p += len; got += len;
    if (head >= 0) break;   // no further sub-record
  }
  return got == n ? off : -2;
}

// The points of a rank sorted by original index fall into a few runs of consecutive file positions (an MPDECOMP rank is one
// run per latitude row), although the rank's own order walks through them differently: `perm` lists the local points in
// file order, a run is a stretch of it that is contiguous in the file.
struct Run { long long src, dst, len; };   // src: first position in perm, dst: first original index (0-based)
struct Plan {
  std::vector<long long> perm;             // empty = identity
  std::vector<Run> runs;
};
Plan plan_of(long long nown, const int* ijorig) {
  Plan pl;
  if (!ijorig) { pl.runs.push_back({0, 0, nown}); return pl; }
  bool sorted = true;
  for (long long i = 1; i < nown && sorted; ++i) sorted = ijorig[i] > ijorig[i - 1];
  if (!sorted) {
    pl.perm.resize((size_t)nown);
    for (long long i = 0; i < nown; ++i) pl.perm[(size_t)i] = i;
    std::sort(pl.perm.begin(), pl.perm.end(), [&](long long a, long long b) { return ijorig[a] < ijorig[b]; });
  }
  auto at = [&](long long j) { return (long long)ijorig[pl.perm.empty() ? j : pl.perm[(size_t)j]]; };
  long long i = 0;
  while (i < nown) {
    long long j = i + 1;
    while (j < nown && at(j) == at(j - 1) + 1) ++j;
    pl.runs.push_back({i, at(i) - 1, j - i});
    i = j;
  }
  return pl;
}
int check_points(long long niblo, long long nown, const int* ijorig) {
  if (niblo < 1 || nown < 0 || nown > niblo) return 1;
  if (!ijorig) return nown == niblo ? 0 : 1;
  for (long long i = 0; i < nown; ++i) if (ijorig[i] < 1 || ijorig[i] > niblo) return 1;
  return 0;   // (a repeated index would only make the last writer win; MPDECOMP's maps are permutations)
}
// `nrec` records of niblo doubles starting at byte `base`: size the file and write every marker
int lay_out_records(int fd, long long base, long long nrec, long long niblo) {
  const long long nb = niblo * 8, span = record_span(nb);
  if (nb > g_maxsub) return -3;
  if (::ftruncate(fd, (off_t)(base + nrec * span))) return -1;
  const int32_t m = (int32_t)nb;
  for (long long r = 0; r < nrec; ++r)
    if (pwrite_all(fd, &m, 4, base + r * span) || pwrite_all(fd, &m, 4, base + r * span + 4 + nb)) return -1;
  return 0;
}
// scatter / gather the points of this rank into / out of record r (data of point i at data[i + nown*r])
int rw_points(int fd, bool wr, long long base, long long nrec, long long niblo, long long nown, const Plan& pl, double* data) {
  const long long span = record_span(niblo * 8);
  std::vector<double> tmp(pl.perm.empty() ? 0 : (size_t)nown);
  for (long long r = 0; r < nrec; ++r) {
    const long long rb = base + r * span + 4;
    double* row = data + nown * r;
    if (!wr) {   // the record markers must be what a record of NIBLO reals carries
      int32_t m[1];
      if (pread_all(fd, m, 4, rb - 4) || m[0] != (int32_t)(niblo * 8)) return -2;
    } else if (!pl.perm.empty()) {
      for (long long j = 0; j < nown; ++j) tmp[(size_t)j] = row[pl.perm[(size_t)j]];
    }
    double* buf = pl.perm.empty() ? row : tmp.data();
    for (const Run& q : pl.runs) {
      const int rc = wr ? pwrite_all(fd, buf + q.src, (size_t)q.len * 8, rb + q.dst * 8) : pread_all(fd, buf + q.src, (size_t)q.len * 8, rb + q.dst * 8);
      if (rc) return rc;
    }
    if (!wr && !pl.perm.empty()) for (long long j = 0; j < nown; ++j) row[pl.perm[(size_t)j]] = tmp[(size_t)j];
  }
  return 0;
}

// days from 1970-01-01 of a proleptic Gregorian date
long long days_from_civil(long long y, int m, int d) {
  y -= m <= 2;
  const long long era = (y >= 0 ? y : y - 399) / 400;
  const long long yoe = y - era * 400;
  const long long doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  const long long doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + doe - 719468;
}
int parse_cdate(const char* c, long long& sec) {   // YYYYMMDDHHmmss
  if (!c || strlen(c) < 14) return 1;
  int v[6];
  const int w[6] = {4, 2, 2, 2, 2, 2};
  int o = 0;
  for (int i = 0; i < 6; ++i) {
    v[i] = 0;
    for (int j = 0; j < w[i]; ++j) { const char ch = c[o++]; if (ch < '0' || ch > '9') return 1; v[i] = v[i] * 10 + (ch - '0'); }
  }
  sec = days_from_civil(v[0], v[1], v[2]) * 86400LL + v[3] * 3600LL + v[4] * 60LL + v[5];
  return 0;
}

}  // namespace

extern "C" {

int ecwam_b200_io_set_max_subrecord(long long nbytes) {
  if (nbytes < 8 || nbytes > 2147483639LL) return ECWAM_B200_EINVAL;
  g_maxsub = nbytes;
  return 0;
}

int ecwam_b200_grstname(const char* cdated, const char* cdatef, int ifcst, const char* fileid, const char* cpad, char* filename, int cap) {
  long long sd, sf;
  if (!fileid || strlen(fileid) != 3 || !filename || parse_cdate(cdated, sd) || parse_cdate(cdatef, sf)) {
    ew_set_error("grstname: dates are YYYYMMDDHHmmss, the file id has 3 characters");
    return ECWAM_B200_EINVAL;
  }
  const bool before = strncmp(cdated, cdatef, 14) < 0;          // grstname.F90:99-122
  const char* cdateh = before ? cdated : cdatef;
  // 64-bit seconds: the reference's ISHIFTDAY / IMAXYEAR chunking (grstname.F90:108-120) only keeps its default-kind DIFDATE from
  // overflowing beyond ~67 years; day count + remainder are the same numbers
  const long long ishift = before ? ifcst : sd - sf;
  const long long dd = ishift / 86400, hh = (ishift - dd * 86400) / 3600, mi = (ishift - dd * 86400 - hh * 3600) / 60;
  const long long ss = ishift - dd * 86400 - hh * 3600 - mi * 60;
  char name[128];
  snprintf(name, sizeof(name), "%.3s%.14s_%06lld%02lld%02lld%02lld", fileid, cdateh, dd, hh, mi, ss);
  std::string out = (cpad && *cpad) ? std::string(cpad) + "/" + name : std::string(name);
  if ((int)out.size() + 1 > cap) { ew_set_error("grstname: file name buffer too small"); return ECWAM_B200_EINVAL; }
  memcpy(filename, out.c_str(), out.size() + 1);
  return 0;
}

int ecwam_b200_restart_par_name(const char* filename, int irank, int nproc, char* out, int cap) {
  if (!filename || !out || irank < 1 || nproc < irank) return ECWAM_B200_EINVAL;
  const int n = snprintf(out, (size_t)cap, "%s.%d_%d", filename, irank, nproc);   // FILENAME//'.%p_%n' (savspec.F90:97-98)
  return (n < 0 || n >= cap) ? ECWAM_B200_EINVAL : 0;
}

int ecwam_b200_savspec(const char* filename, long long niblo, int nang, int nfre, long long nown, const int* ijorig,
                       const double* fl, int create) {
  if (!filename || nang < 1 || nfre < 1 || (nown > 0 && !fl) || check_points(niblo, nown, ijorig)) {
    ew_set_error("savspec: bad arguments");
    return ECWAM_B200_EINVAL;
  }
  Fd f;
  f.fd = ::open(filename, create ? (O_RDWR | O_CREAT | O_TRUNC) : O_RDWR, 0644);
  if (f.fd < 0) IO_FAIL("savspec: cannot open %s: %s", filename, strerror(errno));
  const long long nrec = (long long)nang * nfre;   // MLOOP outside, KLOOP inside, KDEL = MDEL = 1 (savspec.F90:120-160, yowcout.F90:70-71)
  if (create) {
    const int rc = lay_out_records(f.fd, 0, nrec, niblo);
    if (rc) IO_FAIL("savspec: cannot lay out %s (%s)", filename, rc == -3 ? "record above the sub-record limit" : strerror(errno));
  } else {
    struct stat st;
    if (fstat(f.fd, &st) || st.st_size != nrec * record_span(niblo * 8)) IO_FAIL("savspec: %s was not laid out for NIBLO=%lld, %dx%d", filename, niblo, nang, nfre);
  }
  if (rw_points(f.fd, true, 0, nrec, niblo, nown, plan_of(nown, ijorig), const_cast<double*>(fl))) IO_FAIL("savspec: write to %s failed: %s", filename, strerror(errno));
  return 0;
}

int ecwam_b200_getspec(const char* filename, long long niblo, int nang, int nfre, long long nown, const int* ijorig, double* fl) {
  if (!filename || nang < 1 || nfre < 1 || (nown > 0 && !fl) || check_points(niblo, nown, ijorig)) {
    ew_set_error("getspec: bad arguments");
    return ECWAM_B200_EINVAL;
  }
  Fd f;
  f.fd = ::open(filename, O_RDONLY);
  if (f.fd < 0) IO_FAIL("getspec: could not find file %s", filename);          // readfl.F90:92-110
  const long long nrec = (long long)nang * nfre;
  struct stat st;
  if (fstat(f.fd, &st) || st.st_size != nrec * record_span(niblo * 8))
    IO_FAIL("getspec: %s does not hold %lld records of %lld reals (size %lld)", filename, nrec, niblo, (long long)st.st_size);
  if (rw_points(f.fd, false, 0, nrec, niblo, nown, plan_of(nown, ijorig), fl)) IO_FAIL("getspec: %s: record markers do not match NIBLO=%lld", filename, niblo);
  return 0;
}

int ecwam_b200_savspec_par(const char* filename, long long nown, int nang, int nfre, const double* fl) {
  if (!filename || nown < 1 || nang < 1 || nfre < 1 || !fl) { ew_set_error("savspec_par: bad arguments"); return ECWAM_B200_EINVAL; }
  Fd f;
  f.fd = ::open(filename, O_RDWR | O_CREAT | O_TRUNC, 0644);
  if (f.fd < 0) IO_FAIL("savspec_par: cannot open %s: %s", filename, strerror(errno));
  if (write_record(f.fd, 0, fl, nown * nang * nfre * 8) < 0) IO_FAIL("savspec_par: write to %s failed: %s", filename, strerror(errno));
  return 0;
}

int ecwam_b200_getspec_par(const char* filename, long long nown, int nang, int nfre, double* fl) {
  if (!filename || nown < 1 || nang < 1 || nfre < 1 || !fl) { ew_set_error("getspec_par: bad arguments"); return ECWAM_B200_EINVAL; }
  Fd f;
  f.fd = ::open(filename, O_RDONLY);
  if (f.fd < 0) IO_FAIL("getspec_par: could not find file %s", filename);
  if (read_record(f.fd, 0, fl, nown * nang * nfre * 8) < 0) IO_FAIL("getspec_par: %s does not hold one record of %lldx%dx%d reals", filename, nown, nang, nfre);
  return 0;
}

int ecwam_b200_savstress(const char* filename, const char* cdtpro, const char* cdatewo, const char* cdawifl, const char* cdatefl,
                         long long niblo, int nreal, long long nown, const int* ijorig, const double* rfield, int create) {
  if (!filename || nreal < 1 || (nown > 0 && !rfield) || check_points(niblo, nown, ijorig)) { ew_set_error("savstress: bad arguments"); return ECWAM_B200_EINVAL; }
  Fd f;
  f.fd = ::open(filename, create ? (O_RDWR | O_CREAT | O_TRUNC) : O_RDWR, 0644);
  if (f.fd < 0) IO_FAIL("savstress: cannot open %s: %s", filename, strerror(errno));
  const long long base = record_span(56);           // WRITE(IUNIT) CDTPRO, CDATEWO, CDAWIFL, CDATEFL: 4 x CHARACTER*14
  if (create) {
    const char* d[4] = {cdtpro, cdatewo, cdawifl, cdatefl};
    char hdr[56];
    memset(hdr, ' ', sizeof(hdr));
    for (int i = 0; i < 4; ++i) { if (!d[i]) { ew_set_error("savstress: dates missing"); return ECWAM_B200_EINVAL; } memcpy(hdr + 14 * i, d[i], strnlen(d[i], 14)); }
    if (write_record(f.fd, 0, hdr, 56) < 0) IO_FAIL("savstress: write to %s failed: %s", filename, strerror(errno));
    const int rc = lay_out_records(f.fd, base, nreal, niblo);
    if (rc) IO_FAIL("savstress: cannot lay out %s", filename);
  } else {   // the laying-out rank must have finished (the caller's barrier): header record and size as laid out for this NIBLO/NREAL
    struct stat st;
    int32_t mk[2] = {0, 0};
    if (fstat(f.fd, &st) || st.st_size != base + nreal * record_span(niblo * 8) || pread(f.fd, &mk[0], 4, 0) != 4 ||
        pread(f.fd, &mk[1], 4, 4 + 56) != 4 || mk[0] != 56 || mk[1] != 56)
      IO_FAIL("savstress: %s was not laid out for NIBLO=%lld, NREAL=%d", filename, niblo, nreal);
  }
  if (rw_points(f.fd, true, base, nreal, niblo, nown, plan_of(nown, ijorig), const_cast<double*>(rfield))) IO_FAIL("savstress: write to %s failed: %s", filename, strerror(errno));
  return 0;
}

int ecwam_b200_getstress(const char* filename, char* dates, long long niblo, int nreal, long long nown, const int* ijorig, double* rfield) {
  if (!filename || nreal < 1 || (nown > 0 && !rfield) || check_points(niblo, nown, ijorig)) { ew_set_error("getstress: bad arguments"); return ECWAM_B200_EINVAL; }
  Fd f;
  f.fd = ::open(filename, O_RDONLY);
  if (f.fd < 0) IO_FAIL("getstress: could not find file %s", filename);         // readstress.F90:78-96
  char hdr[56];
  const long long base = read_record(f.fd, 0, hdr, 56);
  if (base < 0) IO_FAIL("getstress: %s does not start with the 4 x CHARACTER*14 date record", filename);
  if (dates) for (int i = 0; i < 4; ++i) { memcpy(dates + 15 * i, hdr + 14 * i, 14); dates[15 * i + 14] = 0; }
  struct stat st;
  if (fstat(f.fd, &st) || st.st_size != base + nreal * record_span(niblo * 8)) IO_FAIL("getstress: %s does not hold %d records of %lld reals", filename, nreal, niblo);
  if (rw_points(f.fd, false, base, nreal, niblo, nown, plan_of(nown, ijorig), rfield)) IO_FAIL("getstress: %s: record markers do not match NIBLO=%lld", filename, niblo);
  return 0;
}

int ecwam_b200_grid_tables_write(const char* filename, int imdlgrbid_g, int ngx, int ngy, const int* nlonrgg, int iper, int irgg,
                                 const double* amo, const double* bathy) {
  if (!filename || ngx < 1 || ngy < 1 || !nlonrgg || !amo || !bathy) { ew_set_error("grid_tables_write: bad arguments"); return ECWAM_B200_EINVAL; }
  Fd f;
  f.fd = ::open(filename, O_RDWR | O_CREAT | O_TRUNC, 0644);
  if (f.fd < 0) IO_FAIL("grid_tables_write: cannot open %s: %s", filename, strerror(errno));
  long long off = 0;
  const int32_t r1[2] = {8, imdlgrbid_g}, r2[2] = {ngx, ngy};      // NKIND = KIND(AMOSOP) = 8 in the double-precision build
  char r4[8 + 48];
  const int32_t ii[2] = {iper, irgg};
  memcpy(r4, ii, 8); memcpy(r4 + 8, amo, 48);                      // IPER, IRGG, AMOWEP, AMOSOP, AMOEAP, AMONOP, XDELLA, XDELLO
  if ((off = write_record(f.fd, off, r1, 8)) < 0 || (off = write_record(f.fd, off, r2, 8)) < 0 ||
      (off = write_record(f.fd, off, nlonrgg, (long long)ngy * 4)) < 0 || (off = write_record(f.fd, off, r4, 56)) < 0 ||
      (off = write_record(f.fd, off, bathy, (long long)ngx * ngy * 8)) < 0)
    IO_FAIL("grid_tables_write: write to %s failed: %s", filename, strerror(errno));
  return 0;
}

int ecwam_b200_grid_tables_read(const char* filename, int* nkind, int* kmdlgrdid, int* ngx, int* ngy, int* nlonrgg, int nlon_cap,
                                int* iper, int* irgg, double* amo, double* bathy, long long bathy_cap) {
  if (!filename || !ngx || !ngy) { ew_set_error("grid_tables_read: bad arguments"); return ECWAM_B200_EINVAL; }
  Fd f;
  f.fd = ::open(filename, O_RDONLY);
  if (f.fd < 0) IO_FAIL("grid_tables_read: could not find file %s", filename);
  int32_t r1[2], r2[2];
  long long off = 0;
  if ((off = read_record(f.fd, off, r1, 8)) < 0 || (off = read_record(f.fd, off, r2, 8)) < 0) IO_FAIL("grid_tables_read: %s: bad header records", filename);
  if (nkind) *nkind = r1[0];
  if (kmdlgrdid) *kmdlgrdid = r1[1];
  *ngx = r2[0]; *ngy = r2[1];
  if (r1[0] != 8) IO_FAIL("grid_tables_read: %s was written with REAL*%d, the model runs in REAL*8 (readpre.F90:201-212)", filename, r1[0]);
  if (!nlonrgg && !bathy) return 0;                                 // dimensions only
  if (r2[0] < 1 || r2[1] < 1 || nlon_cap < r2[1] || !nlonrgg) { ew_set_error("grid_tables_read: NLONRGG buffer too small (NGY=%d)", r2[1]); return ECWAM_B200_EINVAL; }
  char r4[56];
  if ((off = read_record(f.fd, off, nlonrgg, (long long)r2[1] * 4)) < 0 || (off = read_record(f.fd, off, r4, 56)) < 0) IO_FAIL("grid_tables_read: %s: bad NLONRGG / grid records", filename);
  int32_t ii[2];
  memcpy(ii, r4, 8);
  if (iper) *iper = ii[0];
  if (irgg) *irgg = ii[1];
  if (amo) memcpy(amo, r4 + 8, 48);
  if (bathy) {
    if (bathy_cap < (long long)r2[0] * r2[1]) { ew_set_error("grid_tables_read: BATHY buffer too small"); return ECWAM_B200_EINVAL; }
    if (read_record(f.fd, off, bathy, (long long)r2[0] * r2[1] * 8) < 0) IO_FAIL("grid_tables_read: %s: bad BATHY record", filename);
  }
  return 0;
}

}  // extern "C"
