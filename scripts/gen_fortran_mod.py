#!/usr/bin/env python
"""Generates fortran/ecwam_b200_mod.F90 — the ISO_C_BINDING mirror of include/ecwam_b200.h (struct members in header order, one
INTERFACE per exported function) — so that the Fortran side cannot drift from the C ABI.  tests/test_abi_and_host.py checks that
the committed file is what this script produces.    usage: python scripts/gen_fortran_mod.py [--check]"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ecwam_b200.h")
OUT = os.path.join(ROOT, "fortran", "ecwam_b200_mod.F90")

STRUCTS = ["ecwam_b200_params", "ecwam_b200_tables", "ecwam_b200_decomp", "ecwam_b200_fields", "ecwam_b200_nemo_fields", "ecwam_b200_forcing_next",
           "ecwam_b200_outsel", "ecwam_b200_fieldg", "ecwam_b200_getwnd_opts"]
# written by hand below (assumed-size array dummies so that the reference's actual arguments can be passed as they are)
HAND = {"ecwam_b200_implsch_f", "ecwam_b200_propag_wam_f"}


def strip_comments(s):
    return re.sub(r"/\*.*?\*/", "", s, flags=re.S)


def wrap(line, indent="      "):
    """Fortran free-form continuation at <= 120 columns."""
    out = []
    while len(line) > 118:
        cut = line.rfind(",", 0, 116)
        out.append(line[:cut + 1] + " &")
        line = indent + " &  " + line[cut + 1:].lstrip()
    out.append(line)
    return "\n".join(out)


def struct_members(src, name):
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, re.S)
    body = strip_comments(m.group(1))
    mem = []
    for line in body.split(";"):
        line = " ".join(line.split())
        if not line:
            continue
        mm = re.match(r"^(const double\*|const int\*|double\*|int\*|double|int) (\w+)$", line)
        mem.append((mm.group(1), mm.group(2)))
    return mem


def f_struct(name, mem):
    lines = ["  TYPE, BIND(C) :: %s" % name.upper()]
    for t, n in mem:
        ft = {"int": "INTEGER(C_INT)", "double": "REAL(C_DOUBLE)"}.get(t, "TYPE(C_PTR)")
        init = " = C_NULL_PTR" if ft == "TYPE(C_PTR)" else ""
        lines.append("    %s :: %s%s" % (ft, n.upper(), init))
    lines.append("  END TYPE %s" % name.upper())
    return "\n".join(lines)


def prototypes(src):
    s = strip_comments(src)
    s = re.sub(r"typedef struct \w+ \{.*?\} \w+;", "", s, flags=re.S)
    protos = []
    for m in re.finditer(r"\n\s*(const char\*|const int\*|const ecwam_b200_\w+\*|int|long long)\s+(ecwam_b200_\w+)\s*\((.*?)\)\s*;", s, re.S):
        protos.append((m.group(1), m.group(2), " ".join(m.group(3).split())))
    return protos


def f_arg(decl):
    """C parameter -> (Fortran declaration, name)."""
    decl = decl.strip()
    if decl == "void":
        return None
    m = re.match(r"^(.*?)(\w+)(\[\d+\])?$", decl)
    ctype, name, arr = m.group(1).strip().replace("struct ", ""), m.group(2).upper(), m.group(3)
    if name in ("OUT", "IN"):       # not reserved words, but keep the generated code easy on the eye
        name = "P" + name
    if arr and "char" in ctype:
        return "CHARACTER(KIND=C_CHAR) :: %s(%s)" % (name, arr[1:-1]), name
    if ctype in ("const char*", "char*"):
        return "CHARACTER(KIND=C_CHAR) :: %s(*)" % name, name
    if ctype == "int":
        return "INTEGER(C_INT), VALUE :: %s" % name, name
    if ctype == "long long":
        return "INTEGER(C_LONG_LONG), VALUE :: %s" % name, name
    if ctype == "double":
        return "REAL(C_DOUBLE), VALUE :: %s" % name, name
    if ctype in ("int*",):
        return "INTEGER(C_INT) :: %s(*)" % name, name
    if ctype == "long long*":
        return "INTEGER(C_LONG_LONG) :: %s" % name, name
    if ctype == "ecwam_b200_exchange_fn":     # a BIND(C) procedure of the host: pass C_FUNLOC(...)
        return "TYPE(C_FUNPTR), VALUE :: %s" % name, name
    if ctype == "ecwam_b200_handle" or ctype in ("void*", "const void*", "ecwam_b200_host_tables_t", "ecwam_b200_host_grid_t"):
        return "TYPE(C_PTR), VALUE :: %s" % name, name
    if ctype in ("ecwam_b200_handle*", "void**", "ecwam_b200_host_tables_t*", "ecwam_b200_host_grid_t*"):
        return "TYPE(C_PTR) :: %s" % name, name
    ms = re.match(r"^(const )?(ecwam_b200_\w+)\*$", ctype)
    if ms and ms.group(2) in STRUCTS:
        return "TYPE(%s) :: %s" % (ms.group(2).upper(), name), name
    if ctype.endswith("*"):          # device or host data pointers: pass C_LOC(...) / the address obtained under host_data use_device
        return "TYPE(C_PTR), VALUE :: %s" % name, name
    raise ValueError("cannot map C parameter %r" % decl)


def f_interface(ret, name, args):
    decls = [f_arg(a) for a in args.split(",")] if args.strip() else []
    decls = [d for d in decls if d]
    names = ", ".join(n for _, n in decls)
    fret = {"int": "INTEGER(C_INT)", "long long": "INTEGER(C_LONG_LONG)"}.get(ret, "TYPE(C_PTR)")   # pointers: C_F_POINTER on the caller's side
    lines = [wrap("    %s FUNCTION %s(%s) BIND(C, NAME='%s')" % (fret, name.upper(), names, name))]
    lines.append("      IMPORT")
    for d, _ in decls:
        lines.append("      " + d)
    lines.append("    END FUNCTION %s" % name.upper())
    return "\n".join(lines)


IMPLSCH_ARGS = ("FL1, WAVNUM, CGROUP, CIWA, CINV, XK2CG, STOKFAC, EMAXDPT, DEPTH, IOBND, IODP, IBRMEM, AIRD, WDWAVE, CICOVER, WSWAVE, "
                "WSTAR, USTRA, VSTRA, UFRIC, TAUW, TAUWDIR, Z0M, Z0B, CHRNCK, CITHICK, NEMOUSTOKES, NEMOVSTOKES, NEMOSTRN, NPHIEPS, NTAUOC, "
                "NSWH, NMWP, NEMOTAUX, NEMOTAUY, NEMOTAUICX, NEMOTAUICY, NEMOWSWAVE, NEMOPHIF, WSEMEAN, WSFMEAN, USTOKES, VSTOKES, STRNMS, "
                "TAUXD, TAUYD, TAUOCXD, TAUOCYD, TAUOC, TAUICX, TAUICY, PHIOCD, PHIEPS, PHIAW, MIJ, XLLWS")


def hand_written():
    ints = {"IOBND", "IODP", "MIJ"}
    names = [a.strip() for a in IMPLSCH_ARGS.split(",")]
    lines = ["    ! the reference argument lists (implsch.F90:10-23, propag_wam.F90:10-11): assumed-size dummies, so the actual arguments of",
             "    ! the reference call sites are passed as they are (their device addresses inside !$acc host_data use_device)",
             wrap("    INTEGER(C_INT) FUNCTION ECWAM_B200_IMPLSCH_F(HANDLE, KIJS, KIJL, %s) BIND(C, NAME='ecwam_b200_implsch_f')" % IMPLSCH_ARGS),
             "      IMPORT", "      TYPE(C_PTR), VALUE :: HANDLE", "      INTEGER(C_INT), VALUE :: KIJS, KIJL"]
    reals = [n for n in names if n not in ints]
    lines.append(wrap("      REAL(C_DOUBLE) :: " + ", ".join(n + "(*)" for n in reals)))
    lines.append("      INTEGER(C_INT) :: " + ", ".join(n + "(*)" for n in names if n in ints))
    lines.append("    END FUNCTION ECWAM_B200_IMPLSCH_F")
    lines.append(wrap("    INTEGER(C_INT) FUNCTION ECWAM_B200_PROPAG_WAM_F(HANDLE, WAVNUM, CGROUP, OMOSNH2KD, FL1, DEPTH, DELLAM1, COSPHM1, UCUR, VCUR) "
                      "BIND(C, NAME='ecwam_b200_propag_wam_f')"))
    lines += ["      IMPORT", "      TYPE(C_PTR), VALUE :: HANDLE",
              "      REAL(C_DOUBLE) :: WAVNUM(*), CGROUP(*), OMOSNH2KD(*), FL1(*), DEPTH(*), DELLAM1(*), COSPHM1(*), UCUR(*), VCUR(*)",
              "    END FUNCTION ECWAM_B200_PROPAG_WAM_F"]
    return "\n".join(lines)


def generate():
    with open(HEADER) as f:
        src = f.read()
    parts = ["""! ecwam_b200_mod.F90 -- ISO_C_BINDING mirror of include/ecwam_b200.h (GENERATED by scripts/gen_fortran_mod.py: do not edit).
!
! The thin layer between ecWAM's Fortran host code and libecwam_b200.so: BIND(C) derived types with the members of the C
! structs in header order, one interface per exported function, and the handle of this MPI task.  The bodies that keep the
! reference's call signatures are fortran/implsch_b200.F90 (IMPLSCH, src/ecwam/implsch.F90:10-23) and
! fortran/propag_wam_b200.F90 (PROPAG_WAM, src/ecwam/propag_wam.F90:10-11); the one-off set-up is fortran/ecwam_b200_setup.F90.
! Reals are C_DOUBLE = JWRB of the double-precision build, integers C_INT = JWIM (parkind_wave.F90:23-35).
MODULE ECWAM_B200_MOD
  USE, INTRINSIC :: ISO_C_BINDING
  IMPLICIT NONE
  PUBLIC

  INTEGER(C_INT), PARAMETER :: ECWAM_B200_OK = 0, ECWAM_B200_EINVAL = -1, ECWAM_B200_ECUDA = -2, ECWAM_B200_ENCCL = -3, &
 &                             ECWAM_B200_ESTATE = -4, ECWAM_B200_EIO = -5
"""]
    for st in STRUCTS:
        parts.append(f_struct(st, struct_members(src, st)) + "\n")
    parts.append("  INTERFACE")
    for ret, name, args in prototypes(src):
        if name in HAND:
            continue
        parts.append(f_interface(ret, name, args))
    parts.append(hand_written())
    parts.append("  END INTERFACE\n")
    parts.append("""  ! one handle per MPI task / GPU (the library entry points are not re-entrant per handle: call them from the master thread)
  TYPE(C_PTR), SAVE :: B200_HANDLE = C_NULL_PTR
  TYPE(ECWAM_B200_FIELDS), SAVE :: B200_FIELDS        ! the device addresses bound with ECWAM_B200_BIND_FIELDS

CONTAINS

  ! last error message of the library as a Fortran string (for WAM_ABORT)
  FUNCTION ECWAM_B200_ERRMSG() RESULT(MSG)
    CHARACTER(LEN=512) :: MSG
    CHARACTER(KIND=C_CHAR), POINTER :: P(:)
    TYPE(C_PTR) :: CP
    INTEGER :: I
    MSG = ' '
    CP = ECWAM_B200_LAST_ERROR()
    IF (.NOT. C_ASSOCIATED(CP)) RETURN
    CALL C_F_POINTER(CP, P, [512])
    DO I = 1, 512
      IF (P(I) == C_NULL_CHAR) EXIT
      MSG(I:I) = P(I)
    ENDDO
  END FUNCTION ECWAM_B200_ERRMSG

END MODULE ECWAM_B200_MOD
""")
    return "\n".join(parts)


if __name__ == "__main__":
    text = generate()
    if "--check" in sys.argv:
        with open(OUT) as f:
            sys.exit(0 if f.read() == text else 1)
    with open(OUT, "w") as f:
        f.write(text)
    print("wrote", OUT, len(text.splitlines()), "lines")
