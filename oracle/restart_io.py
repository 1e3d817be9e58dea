"""TEST INFRASTRUCTURE (oracle): numpy / scipy restatement of the reference's restart and grid-table file formats.

Independent of ecwam_b200/csrc/host_io.cpp: the records are produced and parsed by scipy.io.FortranFile (a third-party
implementation of the Fortran unformatted sequential layout, 4-byte record markers), statement by statement as the
reference's own WRITE / READ lists.  Only tests/ may import this module.

  writefl / readfl         writefl.F90:86-120, readfl.F90:118-145 called as savspec.F90:120-160 (KDEL = MDEL = 1)
  writestress / readstress writestress.F90:76-109, readstress.F90:97-124 (NREAL = 16, savstress.F90:110-125)
  outcom / readpre         outcom.F90:139-144, readpre.F90:262-345 (binary branch)
  grstname                 grstname.F90:88-142

Parity: unpinned against the Fortran runtime itself (no Fortran compiler in this image); the layout is the documented
gfortran one, which scipy.io.FortranFile reads and writes.
"""
import datetime

import numpy as np
from scipy.io import FortranFile


def writefl(path, fl_new, ij2newij=None):
    """fl_new[m, k, ij_new] (the model's relabelled order) -> BLS file.  ij2newij: 1-based map original -> new (NPROC > 1)."""
    F, A, n = fl_new.shape
    with FortranFile(path, "w") as f:
        for m in range(F):            # DO MLOOP (savspec.F90:121)
            for k in range(A):        # DO KLOOP (:124); LOUNIT only for the first block, later ones append (:127-128)
                row = fl_new[m, k]
                if ij2newij is not None:
                    row = row[np.asarray(ij2newij) - 1]     # WRITE(IUNIT) (((FL(IJ2NEWIJ(IJ),J2,J3),IJ=..),..),..) (writefl.F90:114)
                f.write_record(np.ascontiguousarray(row, dtype="<f8"))


def readfl(path, n, A, F, ij2newij=None):
    out = np.empty((F, A, n))
    with FortranFile(path, "r") as f:
        for m in range(F):
            for k in range(A):
                row = f.read_record("<f8")
                assert row.size == n
                if ij2newij is not None:
                    out[m, k, np.asarray(ij2newij) - 1] = row
                else:
                    out[m, k] = row
    return out


def writestress(path, dates, rfield_new, ij2newij=None):
    """dates = (CDTPRO, CDATEWO, CDAWIFL, CDATEFL); rfield_new[field, ij_new]."""
    with FortranFile(path, "w") as f:
        f.write_record(np.frombuffer("".join("%-14.14s" % d for d in dates).encode(), dtype=np.uint8))
        for r in rfield_new:
            if ij2newij is not None:
                r = r[np.asarray(ij2newij) - 1]
            f.write_record(np.ascontiguousarray(r, dtype="<f8"))


def readstress(path, n, nreal, ij2newij=None):
    out = np.empty((nreal, n))
    with FortranFile(path, "r") as f:
        h = f.read_record(np.uint8).tobytes().decode()
        dates = tuple(h[14 * i: 14 * i + 14] for i in range(4))
        for i in range(nreal):
            r = f.read_record("<f8")
            if ij2newij is not None:
                out[i, np.asarray(ij2newij) - 1] = r
            else:
                out[i] = r
    return dates, out


def outcom(path, imdlgrbid_g, nlonrgg, iper, irgg, amo, bathy):
    """bathy[ngy, ngx] (C order = Fortran BATHY(NGX,NGY)); amo = AMOWEP, AMOSOP, AMOEAP, AMONOP, XDELLA, XDELLO."""
    ngy, ngx = bathy.shape
    with FortranFile(path, "w") as f:
        f.write_record(np.array([8, imdlgrbid_g], dtype="<i4"))                       # NKIND, IMDLGRBID_G
        f.write_record(np.array([ngx, ngy], dtype="<i4"))
        f.write_record(np.asarray(nlonrgg, dtype="<i4"))
        f.write_record(np.array([iper, irgg], dtype="<i4"), np.asarray(amo, dtype="<f8"))
        f.write_record(np.ascontiguousarray(bathy, dtype="<f8"))


def readpre(path):
    with FortranFile(path, "r") as f:
        nkind, kmdl = f.read_record("<i4")
        ngx, ngy = f.read_record("<i4")
        nlonrgg = f.read_record("<i4")
        ii, amo = f.read_record("(2,)<i4", "(6,)<f8")
        bathy = f.read_record("<f8").reshape(ngy, ngx)
    return dict(nkind=int(nkind), kmdlgrdid=int(kmdl), ngx=int(ngx), ngy=int(ngy), nlonrgg=nlonrgg, iper=int(ii[0]), irgg=int(ii[1]),
                amo=amo, bathy=bathy)


def _t(c):
    return datetime.datetime.strptime(c, "%Y%m%d%H%M%S")


def grstname(cdated, cdatef, ifcst, fileid, cpad=""):
    if cdated < cdatef:
        cdateh, shift = cdated, ifcst
    else:
        cdateh, shift = cdatef, int((_t(cdated) - _t(cdatef)).total_seconds())
    dd, rem = divmod(shift, 86400)
    hh, rem = divmod(rem, 3600)
    mi, ss = divmod(rem, 60)
    name = "%s%s_%06d%02d%02d%02d" % (fileid, cdateh, dd, hh, mi, ss)
    return (cpad + "/" + name) if cpad else name
