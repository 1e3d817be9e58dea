"""Restart files (BLS / LAW) and the grid-table file in the reference's on-disk formats: the C-ABI
host functions of ecwam_b200/csrc/host_io.cpp against oracle/restart_io.py (scipy.io.FortranFile, statement by statement
as WRITEFL / READFL / WRITESTRESS / READSTRESS / OUTCOM / READPRE).  Files must be byte-identical, whichever side wrote them
and however many ranks wrote them.  No GPU needed (the GPU round trip through a restart is in test_gpu_output.py)."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from ecwam_b200 import lib as L, model as M, synth
from oracle import restart_io as R

dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


def _ptr(a, t=dp):
    return a.ctypes.data_as(t)


def _setup(nproc):
    g = synth.make_grid(12, "continents")
    return g, M.WamSetup(g, nproc=nproc, nang=12, nfre_red=25)


def _own(s, r):
    """original 1-based indices of rank r's points, in the rank's own (relabelled) order"""
    return np.ascontiguousarray(s.new2ij[s.nstart[r]: s.nend[r] + 1], dtype=np.int32)


@pytest.mark.parametrize("nproc", [1, 3, 4])
def test_bls_file_is_byte_identical_to_the_reference_layout(built, tmp_path, nproc):
    lib = L.load()
    g, s = _setup(nproc)
    n, A, F = s.niblo, 12, 36
    rng = np.random.default_rng(3)
    fl_new = rng.random((F, A, n))                          # the model's (relabelled) point order
    ij2new = np.asarray(s.ij2new[1:], dtype=np.int64)       # original -> new, 1-based
    ref = str(tmp_path / "ref.bls")
    R.writefl(ref, fl_new, ij2new if nproc > 1 else None)
    out = str(tmp_path / "out.bls").encode()
    for r in range(nproc):                                  # every rank writes its own points in place
        ij = _own(s, r)
        mine = np.ascontiguousarray(fl_new[:, :, s.nstart[r] - 1: s.nend[r]])
        L.check(lib.ecwam_b200_savspec(out, n, A, F, ij.size, _ptr(ij, ip), _ptr(mine), int(r == 0)), "savspec")
    assert open(out, "rb").read() == open(ref, "rb").read()
    assert os.path.getsize(out) == A * F * (8 * n + 8)
    # ... and each rank reads its own points back out of the reference-written file
    for r in range(nproc):
        ij = _own(s, r)
        got = np.empty((F, A, ij.size))
        L.check(lib.ecwam_b200_getspec(ref.encode(), n, A, F, ij.size, _ptr(ij, ip), _ptr(got)), "getspec")
        np.testing.assert_array_equal(got, fl_new[:, :, s.nstart[r] - 1: s.nend[r]])
    np.testing.assert_array_equal(R.readfl(out.decode(), n, A, F, ij2new if nproc > 1 else None), fl_new)


@pytest.mark.parametrize("nproc", [1, 4])
def test_law_file_is_byte_identical_to_the_reference_layout(built, tmp_path, nproc):
    lib = L.load()
    g, s = _setup(nproc)
    n, NREAL = s.niblo, 16
    r_new = np.random.default_rng(5).normal(size=(NREAL, n))
    ij2new = np.asarray(s.ij2new[1:], dtype=np.int64)
    dates = ("20220101060000", "20220101070000", "20220101000000", "20220102000000")
    ref = str(tmp_path / "ref.law")
    R.writestress(ref, dates, r_new, ij2new if nproc > 1 else None)
    out = str(tmp_path / "out.law").encode()
    for r in range(nproc):
        ij = _own(s, r)
        mine = np.ascontiguousarray(r_new[:, s.nstart[r] - 1: s.nend[r]])
        L.check(lib.ecwam_b200_savstress(out, *[d.encode() for d in dates], n, NREAL, ij.size, _ptr(ij, ip), _ptr(mine), int(r == 0)),
                "savstress")
    assert open(out, "rb").read() == open(ref, "rb").read()
    for r in range(nproc):
        ij = _own(s, r)
        got = np.empty((NREAL, ij.size))
        d = C.create_string_buffer(60)
        L.check(lib.ecwam_b200_getstress(ref.encode(), d, n, NREAL, ij.size, _ptr(ij, ip), _ptr(got)), "getstress")
        np.testing.assert_array_equal(got, r_new[:, s.nstart[r] - 1: s.nend[r]])
        assert tuple(d.raw[15 * i: 15 * i + 14].decode() for i in range(4)) == dates
    d2, back = R.readstress(out.decode(), n, NREAL, ij2new if nproc > 1 else None)
    assert d2 == dates
    np.testing.assert_array_equal(back, r_new)


def test_per_rank_restart_files_and_subrecords(built, tmp_path):
    """LRSTPARALW (savspec.F90:96-116): FILENAME.<irank>_<nproc>, one record per file.  A record above the sub-record limit
    is a chain of sub-records: leading marker negative while another follows, trailing marker negative when one precedes."""
    lib = L.load()
    name = C.create_string_buffer(256)
    L.check(lib.ecwam_b200_restart_par_name(b"/x/BLS20220101000000_000000060000", 3, 16, name, 256), "par_name")
    assert name.value == b"/x/BLS20220101000000_000000060000.3_16"            # expand_string.F90: plain integers
    nown, A, F = 37, 12, 36
    fl = np.random.default_rng(1).random((F, A, nown))
    p = str(tmp_path / "bls.1_2")
    L.check(lib.ecwam_b200_savspec_par(p.encode(), nown, A, F, _ptr(fl)), "savspec_par")
    from scipy.io import FortranFile
    with FortranFile(p, "r") as f:
        np.testing.assert_array_equal(f.read_record("<f8").reshape(F, A, nown), fl)
    # the same record split into sub-records of at most 1000 bytes
    try:
        L.check(lib.ecwam_b200_io_set_max_subrecord(1000), "set_max_subrecord")
        q = str(tmp_path / "bls_sub.1_2")
        L.check(lib.ecwam_b200_savspec_par(q.encode(), nown, A, F, _ptr(fl)), "savspec_par")
        raw = open(q, "rb").read()
        nbytes, off, parts, i = fl.nbytes, 0, [], 0
        nsub = -(-nbytes // 1000)
        while off < len(raw):
            head, = struct.unpack_from("<i", raw, off)
            ln = abs(head)
            tail, = struct.unpack_from("<i", raw, off + 4 + ln)
            assert (head < 0) == (i < nsub - 1) and (tail < 0) == (i > 0) and abs(tail) == ln and ln <= 1000
            parts.append(raw[off + 4: off + 4 + ln])
            off += ln + 8
            i += 1
        assert i == nsub and b"".join(parts) == fl.tobytes()
        back = np.empty_like(fl)
        L.check(lib.ecwam_b200_getspec_par(q.encode(), nown, A, F, _ptr(back)), "getspec_par")
        np.testing.assert_array_equal(back, fl)
    finally:
        L.check(lib.ecwam_b200_io_set_max_subrecord(2147483639), "set_max_subrecord")


def test_grid_tables_file(built, tmp_path):
    """wam_grid_tables (outcom.F90:139-144 / readpre.F90:262-345): written here == written record by record by the oracle;
    the synthetic octahedral grid survives the round trip."""
    lib = L.load()
    g = synth.make_grid(12, "continents")
    ngy, ngx = int(g.ngy), int(max(g.nlonrgg))
    bathy = np.full((ngy, ngx), -999.0)
    k = 0
    mask = np.asarray(g.mask).reshape(-1) if not isinstance(g.mask, (bytes, bytearray)) else np.frombuffer(g.mask, np.uint8)
    cell = 0
    for j in range(ngy):
        for i in range(int(g.nlonrgg[j])):
            if mask[cell]:
                bathy[j, i] = g.depth[k]
                k += 1
            cell += 1
    assert k == g.niblo
    nl = np.ascontiguousarray(g.nlonrgg, dtype=np.int32)
    amo = np.array([0.0, g.amosop, 360.0 - 360.0 / ngx, g.amonop, (g.amonop - g.amosop) / (ngy - 1), 360.0 / ngx])
    ref, out = str(tmp_path / "ref_grid"), str(tmp_path / "out_grid")
    R.outcom(ref, 108, nl, 1, 1, amo, bathy)
    L.check(lib.ecwam_b200_grid_tables_write(out.encode(), 108, ngx, ngy, _ptr(nl, ip), 1, 1, _ptr(amo), _ptr(bathy)), "grid_tables_write")
    assert open(out, "rb").read() == open(ref, "rb").read()
    v = [C.c_int() for _ in range(6)]
    L.check(lib.ecwam_b200_grid_tables_read(ref.encode(), C.byref(v[0]), C.byref(v[1]), C.byref(v[2]), C.byref(v[3]), None, 0, None, None,
                                            None, None, 0), "grid_tables_read (dimensions)")
    assert (v[0].value, v[1].value, v[2].value, v[3].value) == (8, 108, ngx, ngy)
    nl2, amo2, b2 = np.empty(ngy, np.int32), np.empty(6), np.empty((ngy, ngx))
    L.check(lib.ecwam_b200_grid_tables_read(ref.encode(), C.byref(v[0]), C.byref(v[1]), C.byref(v[2]), C.byref(v[3]), _ptr(nl2, ip), ngy,
                                            C.byref(v[4]), C.byref(v[5]), _ptr(amo2), _ptr(b2), b2.size), "grid_tables_read")
    np.testing.assert_array_equal(nl2, nl); np.testing.assert_array_equal(amo2, amo); np.testing.assert_array_equal(b2, bathy)
    assert (v[4].value, v[5].value) == (1, 1)
    d = R.readpre(out)
    np.testing.assert_array_equal(d["bathy"], bathy)
    # the sea mask the model derives from it (mgrid.F90:72: BATHY > ZMISS) rebuilds the same decomposition
    g2 = synth.make_grid(12, "continents")
    assert ((b2 > -990.0).sum(), int(g2.niblo)) == (g.niblo, g.niblo)
    # a REAL*4 file is refused like READPRE does (readpre.F90:201-212)
    raw = bytearray(open(ref, "rb").read())
    struct.pack_into("<i", raw, 4, 4)
    bad = str(tmp_path / "r4_grid")
    open(bad, "wb").write(raw)
    rc = lib.ecwam_b200_grid_tables_read(bad.encode(), C.byref(v[0]), C.byref(v[1]), C.byref(v[2]), C.byref(v[3]), None, 0, None, None, None, None, 0)
    assert rc == -5 and b"REAL*4" in lib.ecwam_b200_last_error()


@pytest.mark.parametrize("cdated,cdatef,ifcst", [("20220101060000", "20220101000000", 0), ("20220103120530", "20220101000000", 0),
                                                 ("20211231180000", "20220101000000", 0), ("20240301000000", "20240228120000", 0),
                                                 ("20220101000000", "20220101000000", 0), ("20991231235959", "19700101000000", 0)])
def test_restart_file_names(built, cdated, cdatef, ifcst):
    """GRSTNAME (grstname.F90:88-142): <ID><analysis date>_<dddddd hh mm ss forecast range>."""
    lib = L.load()
    for cpad in ("", "/scratch/run"):
        buf = C.create_string_buffer(300)
        L.check(lib.ecwam_b200_grstname(cdated.encode(), cdatef.encode(), ifcst, b"BLS", cpad.encode(), buf, 300), "grstname")
        assert buf.value.decode() == R.grstname(cdated, cdatef, ifcst, "BLS", cpad)
    assert R.grstname("20220101060000", "20220101000000", 0, "LAW") == "LAW20220101000000_000000060000"


def test_restart_errors_are_loud(built, tmp_path):
    lib = L.load()
    x = np.zeros((36, 12, 10))
    rc = lib.ecwam_b200_getspec(str(tmp_path / "missing").encode(), 10, 12, 36, 10, None, _ptr(x))
    assert rc == -5 and b"could not find file" in lib.ecwam_b200_last_error()          # readfl.F90:92-110
    p = str(tmp_path / "short.bls")
    R.writefl(p, np.zeros((36, 12, 9)))
    rc = lib.ecwam_b200_getspec(p.encode(), 10, 12, 36, 10, None, _ptr(x))
    assert rc == -5 and b"does not hold" in lib.ecwam_b200_last_error()
    ij = np.array([0, 1, 2], dtype=np.int32)                                            # index outside 1..NIBLO
    assert lib.ecwam_b200_savspec(p.encode(), 10, 12, 36, 3, _ptr(ij, ip), _ptr(x), 1) == -1
    # a rank that is not the creating one must find the file laid out for the same dimensions
    assert lib.ecwam_b200_savspec(p.encode(), 10, 12, 36, 10, None, _ptr(x), 0) == -5


def test_grid_from_grid_tables_file_rebuilds_the_decomposition(built, tmp_path):
    """synth.write_grid_tables -> synth.grid_from_tables: a grid that only exists as a `wam_grid_tables` file gives the same sea
    points, depths and MPDECOMP tables as the grid it was written from."""
    g = synth.make_grid(16, "continents")
    p = str(tmp_path / "wam_grid_tables")
    synth.write_grid_tables(g, p)
    h = synth.grid_from_tables(p)
    assert (h.ngy, h.niblo, h.amosop, h.amonop) == (g.ngy, g.niblo, g.amosop, g.amonop)
    for k in ("nlonrgg", "mask", "row_of", "lon", "lat", "depth"):
        np.testing.assert_array_equal(np.asarray(getattr(h, k)), np.asarray(getattr(g, k)), err_msg=k)
    a, b = M.WamSetup(g, nproc=3, nang=12, nfre_red=25), M.WamSetup(h, nproc=3, nang=12, nfre_red=25)
    np.testing.assert_array_equal(a.ij2new, b.ij2new)
    for r in range(3):
        da, db = a.decomp_arrays(r), b.decomp_arrays(r)
        for nm in ("klat", "klon", "kcor", "wlat", "wcor", "ntope", "nfrompe"):
            np.testing.assert_array_equal(da[nm], db[nm], err_msg=nm)
