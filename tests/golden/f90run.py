"""TEST INFRASTRUCTURE: a translator from the Fortran 90 subset the ecWAM physics routines are written in to Python, so that the
REFERENCE'S OWN SOURCE TEXT can be executed in the build container (no Fortran compiler exists in the image) and its outputs
committed as golden vectors (tests/golden/ref_*.npz, made by tests/golden/make_ref_golden.py).

It executes the statements of a routine in source order with IEEE double arithmetic in Python's (= Fortran's) operator
precedence and left-to-right order; integers keep Fortran semantics (truncating division) through `FInt`; arrays are `FArr`
(arbitrary lower bounds, column-major meaning irrelevant here: every access is by subscript).  Nothing is restated by hand: the
only inputs besides the reference file are the values of the module variables the routine USEs.

Supported: SUBROUTINE / FUNCTION, declarations with DIMENSION / inline dimensions / PARAMETER, DO / DO WHILE / CYCLE / EXIT,
block and one-line IF, SELECT CASE, CALL (scalar INTENT(OUT/INOUT) dummies are returned and re-assigned), array sections,
whole-array assignment, statement functions, `#include` of the two function headers, cpp `#ifdef` blocks (no macro defined).
Not supported (raises if reached): WHERE, derived types, I/O.
"""
from __future__ import annotations

import math
import os
import re

import numpy as np

REF = "/root/reference/src/ecwam"


# ---------------------------------------------------------------------------------------------------------------- runtime
class FInt(int):
    """Fortran INTEGER: arithmetic between integers stays integer, `/` truncates toward zero."""
    __slots__ = ()

    def _w(self, v):
        return FInt(v) if isinstance(v, int) and not isinstance(v, bool) else v

    def __add__(self, o): return self._w(int.__add__(self, o)) if isinstance(o, int) else o.__radd__(int(self)) if not isinstance(o, float) else float(self) + o
    def __radd__(self, o): return self._w(int.__radd__(self, o)) if isinstance(o, int) else o + float(self)
    def __sub__(self, o): return self._w(int.__sub__(self, o)) if isinstance(o, int) else float(self) - o
    def __rsub__(self, o): return self._w(int.__rsub__(self, o)) if isinstance(o, int) else o - float(self)
    def __mul__(self, o): return self._w(int.__mul__(self, o)) if isinstance(o, int) else float(self) * o
    def __rmul__(self, o): return self._w(int.__rmul__(self, o)) if isinstance(o, int) else o * float(self)
    def __neg__(self): return FInt(int.__neg__(self))
    def __pos__(self): return self
    def __abs__(self): return FInt(int.__abs__(self))

    def __truediv__(self, o):
        if isinstance(o, int):
            q = abs(int(self)) // abs(int(o))
            return FInt(q if (int(self) >= 0) == (int(o) >= 0) else -q)
        return float(self) / o

    def __rtruediv__(self, o):
        if isinstance(o, int):
            q = abs(int(o)) // abs(int(self))
            return FInt(q if (int(self) >= 0) == (int(o) >= 0) else -q)
        return o / float(self)

    def __pow__(self, o, mod=None):
        if isinstance(o, int):
            return FInt(int.__pow__(self, o)) if o >= 0 else FInt(0 if abs(int(self)) > 1 else int(self) ** o)
        return float(self) ** o

    def __rpow__(self, o, mod=None):
        if isinstance(o, int):
            return FInt(int(o) ** int(self)) if self >= 0 else 1.0 / (int(o) ** -int(self))
        return fpow(o, self)


def fpow(x, n):
    """X**N: integer exponents by repeated multiplication (what compilers emit for small constant powers), else pow()."""
    if isinstance(n, int) and not isinstance(x, np.ndarray):
        if isinstance(x, int):
            return FInt(int(x) ** int(n)) if n >= 0 else FInt(0)
        k = abs(int(n))
        if k <= 8:
            r = 1.0
            for _ in range(k):
                r = r * x
            return r if n >= 0 else 1.0 / r
        return math.pow(x, int(n))
    if isinstance(x, np.ndarray) or isinstance(n, np.ndarray):
        return np.power(x, n)
    return math.pow(float(x), float(n))


def frange(a, b, c=1):
    a, b, c = int(a), int(b), int(c)
    v = a
    if c > 0:
        while v <= b:
            yield FInt(v)
            v += c
    else:
        while v >= b:
            yield FInt(v)
            v += c


class FArr:
    """Fortran array with explicit bounds; subscripts are Fortran subscripts, sections are inclusive."""

    def __init__(self, bounds, dtype=float, data=None):
        self.lb = [int(lo) for lo, _ in bounds]
        shape = [max(0, int(hi) - int(lo) + 1) for lo, hi in bounds]
        self.a = np.zeros(shape, dtype=dtype) if data is None else data
        if dtype is float and data is None:
            self.a.fill(np.nan)          # Fortran leaves locals undefined: reading one before it is set shows up as NaN
        self.kind = dtype

    @classmethod
    def of(cls, arr, lb=None):
        arr = np.array(arr, copy=True)
        o = cls.__new__(cls)
        o.a = arr
        o.lb = list(lb) if lb is not None else [1] * arr.ndim
        o.kind = float if arr.dtype.kind == "f" else (bool if arr.dtype.kind == "b" else int)
        return o

    def _ix(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        out, scalar = [], True
        for d, i in enumerate(idx):
            if isinstance(i, slice):
                scalar = False
                lo = None if i.start is None else int(i.start) - self.lb[d]
                hi = None if i.stop is None else int(i.stop) - self.lb[d] + 1
                if lo is not None and lo < 0:
                    raise IndexError("section below the lower bound")
                out.append(slice(lo, hi, i.step))
            else:
                j = int(i) - self.lb[d]
                if j < 0 or j >= self.a.shape[d]:
                    raise IndexError("subscript %d of dimension %d out of bounds %d:%d" % (int(i), d + 1, self.lb[d], self.lb[d] + self.a.shape[d] - 1))
                out.append(j)
        return tuple(out), scalar

    oob_read_zero = False      # set by a driver for a routine known to read one element outside an array with weight 0 (SEBTMEAN's FR(0))

    def __getitem__(self, idx):
        try:
            ix, scalar = self._ix(idx)
        except IndexError:
            if FArr.oob_read_zero:
                return 0.0 if self.kind is float else FInt(0)
            raise
        v = self.a[ix]
        if scalar:
            if self.kind is float:
                return float(v)
            if self.kind is bool:
                return bool(v)
            return FInt(int(v))
        return v

    def __setitem__(self, idx, val):
        ix, _ = self._ix(idx)
        self.a[ix] = val

    def assign(self, val):
        self.a[...] = val.a if isinstance(val, FArr) else val

    @classmethod
    def view(cls, arr):
        """an array section passed as an actual argument: shares the storage (numpy basic slicing), lower bounds 1"""
        o = cls.__new__(cls)
        o.a, o.lb = arr, [1] * arr.ndim
        o.kind = float if arr.dtype.kind == "f" else (bool if arr.dtype.kind == "b" else int)
        return o

    def rebase(self, lbs):
        """the same storage seen with the callee's declared lower bounds (dummy-array bounds are local to the routine)"""
        if list(lbs) == self.lb or len(lbs) != self.a.ndim:
            return self
        o = FArr.__new__(FArr)
        o.a, o.lb, o.kind = self.a, [int(x) for x in lbs], self.kind
        return o

    def elemview(self, i):
        """actual argument A(i) for an array dummy (sequence association, rank 1): the storage from element i on"""
        assert self.a.ndim == 1
        o = FArr.__new__(FArr)
        o.a, o.lb, o.kind = self.a[int(i) - self.lb[0]:], [1], self.kind
        return o


def _isarr(x):
    return isinstance(x, np.ndarray)


def _fun1(m, n):
    def f(x):
        return n(x) if _isarr(x) else m(float(x))
    return f


def F_MAX(*a):
    if all(isinstance(v, int) for v in a):
        return FInt(max(int(v) for v in a))
    r = a[0]
    for v in a[1:]:
        r = np.maximum(r, v) if (_isarr(r) or _isarr(v)) else (float(v) if float(v) > float(r) else float(r))
    return r


def F_MIN(*a):
    if all(isinstance(v, int) for v in a):
        return FInt(min(int(v) for v in a))
    r = a[0]
    for v in a[1:]:
        r = np.minimum(r, v) if (_isarr(r) or _isarr(v)) else (float(v) if float(v) < float(r) else float(r))
    return r


def F_SIGN(a, b):
    if isinstance(a, int) and isinstance(b, int):
        return FInt(abs(int(a)) if b >= 0 else -abs(int(a)))
    if _isarr(a) or _isarr(b):
        return np.copysign(np.abs(a), b)
    return math.copysign(abs(float(a)), float(b))


def F_MOD(a, b):
    if isinstance(a, int) and isinstance(b, int):
        return FInt(int(math.fmod(int(a), int(b))))
    if _isarr(a) or _isarr(b):
        return np.fmod(a, b)
    return math.fmod(float(a), float(b))


def F_ABS(x):
    if isinstance(x, int):
        return FInt(abs(int(x)))
    return np.abs(x) if _isarr(x) else abs(float(x))


def F_INT(x, *k): return FInt(int(x))
def F_NINT(x, *k): return FInt(int(math.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1))
def F_FLOOR(x, *k): return FInt(math.floor(x))
def F_CEILING(x, *k): return FInt(math.ceil(x))
def F_REAL(x, *k): return x.astype(float) if _isarr(x) else float(x)
def F_MERGE(a, b, c): return np.where(c, a, b) if _isarr(c) else (a if c else b)
def F_SUM(x): return float(np.sum(x.a if isinstance(x, FArr) else x))
def F_MAXVAL(x): return float(np.max(x.a if isinstance(x, FArr) else x))
def F_MINVAL(x): return float(np.min(x.a if isinstance(x, FArr) else x))
def F_EPSILON(x): return 2.220446049250313e-16
def F_TINY(x): return 2.2250738585072014e-308
def F_HUGE(x): return 1.7976931348623157e308 if isinstance(x, float) else FInt(2147483647)
def F_ATAN2(a, b): return np.arctan2(a, b) if (_isarr(a) or _isarr(b)) else math.atan2(float(a), float(b))
def F_ALLOCATED(x): return x is not None
def F_ICONV(v): return v if isinstance(v, int) else FInt(int(v))          # REAL -> INTEGER on assignment: truncation
def F_RCONV(v): return float(v) if isinstance(v, int) and not isinstance(v, bool) else v     # INTEGER -> REAL on assignment
def F_PRESENT(x): return x is not None
def F_SIZE(x, d=None): return FInt(x.a.size if d is None else x.a.shape[int(d) - 1])
def F_KIND(x): return FInt(8 if (isinstance(x, FArr) and x.kind is float) or isinstance(x, float) else 4)


RUNTIME = dict(JWRB=8, JWRU=8, JWRO=8, JWIM=4, JPHOOK=8, FInt=FInt, FArr=FArr, frange=frange, fpow=fpow, np=np, math=math,
               F_MAX=F_MAX, F_MIN=F_MIN, F_SIGN=F_SIGN, F_MOD=F_MOD, F_ABS=F_ABS, F_INT=F_INT, F_NINT=F_NINT, F_FLOOR=F_FLOOR,
               F_CEILING=F_CEILING, F_REAL=F_REAL, F_MERGE=F_MERGE, F_SUM=F_SUM, F_MAXVAL=F_MAXVAL, F_MINVAL=F_MINVAL, F_EPSILON=F_EPSILON,
               F_TINY=F_TINY, F_HUGE=F_HUGE, F_ATAN2=F_ATAN2, F_ALLOCATED=F_ALLOCATED, F_ICONV=F_ICONV, F_RCONV=F_RCONV, F_PRESENT=F_PRESENT, F_KIND=F_KIND, F_SIZE=F_SIZE,
               F_SQRT=_fun1(math.sqrt, np.sqrt), F_EXP=_fun1(math.exp, np.exp), F_LOG=_fun1(math.log, np.log),
               F_LOG10=_fun1(math.log10, np.log10), F_TANH=_fun1(math.tanh, np.tanh), F_SINH=_fun1(math.sinh, np.sinh),
               F_COSH=_fun1(math.cosh, np.cosh), F_COS=_fun1(math.cos, np.cos), F_SIN=_fun1(math.sin, np.sin),
               F_TAN=_fun1(math.tan, np.tan), F_ATAN=_fun1(math.atan, np.arctan), F_ACOS=_fun1(math.acos, np.arccos),
               F_ASIN=_fun1(math.asin, np.arcsin))
INTRINSICS = {k[2:] for k in RUNTIME if k.startswith("F_")}


# ------------------------------------------------------------------------------------------------------------- translator
class Routine:
    def __init__(self, name, kind, args):
        self.name, self.kind, self.args = name, kind, args
        self.arrays = set()          # names subscripted as arrays inside this routine (dummy, local)
        self.decl = {}               # name -> dict(type, dims (list of (lo, hi) source strings) or None, intent, param)
        self.uses = {}               # module -> [names]
        self.body = []               # logical source lines
        self.stmtfun = set()
        self.py = None


def _logical_lines(path, include_dirs):
    """cpp (#ifdef with nothing defined, #include of non-interface headers), comments, continuation lines -> upper-case statements."""
    raw = []

    def read(p):
        with open(p, errors="replace") as fh:
            for ln in fh:
                m = re.match(r'\s*#include\s+"([^"]+)"', ln)
                if m:
                    inc = m.group(1)
                    if inc.endswith(".intfb.h"):
                        continue
                    for d in include_dirs:
                        q = os.path.join(d, inc)
                        if os.path.exists(q):
                            read(q)
                            break
                    continue
                raw.append(ln.rstrip("\n"))
    read(path)
    out, stack, cur = [], [], ""
    for ln in raw:
        s = ln.strip()
        if s.startswith("#"):
            d = s[1:].strip()
            if d.startswith("ifdef") or (d.startswith("if ") and "defined" in d):
                stack.append(False)
            elif d.startswith("ifndef"):
                stack.append(True)
            elif d.startswith("else"):
                stack[-1] = not stack[-1]
            elif d.startswith("endif"):
                stack.pop()
            continue
        if stack and not all(stack):
            continue
        # strip comments (no '!' inside the strings of the statements we keep)
        if "!" in ln:
            q, o = False, []
            for ch in ln:
                if ch in "'\"":
                    q = not q
                if ch == "!" and not q:
                    break
                o.append(ch)
            ln = "".join(o)
        s = ln.strip()
        if not s:
            continue
        if s.startswith("&"):
            s = s[1:].strip()
        cont = s.endswith("&")
        if cont:
            s = s[:-1].rstrip()
        cur = (cur + " " + s) if cur else s
        if not cont:
            for part in _split_semicolon(cur):
                part = part.upper()
                if "'" not in part and '"' not in part:
                    part = re.sub(r"\s*%\s*", "_", part)          # derived-type components become plain names (BLK2GLO%KXLT -> BLK2GLO_KXLT)
                out.append(part)
            cur = ""
    # `IF (c) GO TO n` ... `n CONTINUE` directly in front of END DO is a CYCLE; other GO TOs are not supported
    labels = {}
    for i, ln in enumerate(out):
        m = re.match(r"^(\d+)\s+CONTINUE$", ln)
        if m:
            labels[m.group(1)] = i
    # a backward `GO TO n` to a label that opens the rest of the routine (AKI's iteration) is a loop: `n CONTINUE` -> LABEL_LOOP
    back = set()
    for i, ln in enumerate(out):
        m = re.search(r"\bGO\s*TO\s+(\d+)$", ln)
        if m and labels.get(m.group(1), 1 << 30) < i:
            back.add(m.group(1))
    # a forward `GO TO n` to a labelled executable statement (KZEONE's three branches): the routine is cut into segments at those
    # statements; `GO TO n` outside any DO loop -> GOTOSEG n (skip to the segment), falling off a segment's end enters the next one
    seg = {}
    for i, ln in enumerate(out):
        m = re.match(r"^(\d+)\s+(?!CONTINUE$|FORMAT\b)(\S.*)$", ln)
        if m:
            seg[m.group(1)] = i
    res = []
    for i, ln in enumerate(out):
        m = re.match(r"^(\d+)\s+CONTINUE$", ln)
        if m:
            if m.group(1) in back:
                res.append("LABEL_LOOP")
            continue
        m = re.match(r"^(\d+)\s+(?!FORMAT\b)(\S.*)$", ln)
        if m and m.group(1) in seg:
            res.append("LABEL_SEG " + m.group(1))
            ln = m.group(2)
        m = re.search(r"\bGO\s*TO\s+(\d+)$", ln)
        if m and seg.get(m.group(1), -1) > i:
            res.append(ln[:m.start()] + "GOTOSEG " + m.group(1))
            continue
        if m:
            j = labels.get(m.group(1))
            if m.group(1) in back:
                ln = ln[:m.start()] + "CYCLE"
            elif j is not None and j + 1 < len(out) and re.match(r"^END\s*DO$", out[j + 1]):
                ln = ln[:m.start()] + "CYCLE"
            else:
                ln = ln[:m.start()] + "CALL ABORT1"      # unsupported GO TO: fails if reached
        res.append(ln)
    return res


def _split_semicolon(s):
    if ";" not in s:
        return [s]
    parts, depth, q, cur = [], 0, False, []
    for ch in s:
        if ch in "'\"":
            q = not q
        if ch == ";" and not q:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur).strip())
    return [p for p in parts if p]


def _split_top(s, sep=","):
    parts, depth, cur = [], 0, []
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur).strip())
    return parts


_TOK = re.compile(r"\s*(\d+\.\d*(?:[ED][+-]?\d+)?(?:_\w+)?|\.\d+(?:[ED][+-]?\d+)?(?:_\w+)?|\d+(?:[ED][+-]?\d+)(?:_\w+)?|\d+(?:_\w+)?|"
                  r"\.(?:AND|OR|NOT|TRUE|FALSE|EQ|NE|LT|LE|GT|GE|EQV|NEQV)\.|[A-Z_][A-Z0-9_]*|\*\*|==|/=|<=|>=|//|'[^']*'|\"[^\"]*\"|.)")
_DOT = {".AND.": " and ", ".OR.": " or ", ".NOT.": " not ", ".TRUE.": " True ", ".FALSE.": " False ", ".EQ.": "==", ".NE.": "!=",
        ".LT.": "<", ".LE.": "<=", ".GT.": ">", ".GE.": ">=", ".EQV.": "==", ".NEQV.": "!="}


def module_registry(modules, include_dirs=(REF,)):
    """name -> (dtype, is_array) of the variables declared in the given YOW* module files"""
    reg = {}
    for m in modules:
        p = m if os.path.isabs(m) else os.path.join(REF, m + ".F90")
        if not os.path.exists(p):
            continue
        for ln in _logical_lines(p, list(include_dirs)):
            mm = re.match(r"^(REAL|INTEGER|LOGICAL)\s*(\((?:[^()]|\([^()]*\))*\))?\s*(.*?)::\s*(.*)$", ln)
            if not mm:
                continue
            typ = {"REAL": float, "INTEGER": int, "LOGICAL": bool}[mm.group(1)]
            arr_attr = "DIMENSION" in mm.group(3)
            for ent in _split_top(mm.group(4)):
                em = re.match(r"^(\w+)\s*(\(.*\))?", ent)
                if em:
                    reg[em.group(1)] = (typ, arr_attr or em.group(2) is not None)
                    if em.group(2) is not None and ":" not in em.group(2):
                        STATIC_DIMS[em.group(1)] = (typ, _split_top(em.group(2)[1:-1]))
    return reg


STATIC_DIMS = {}     # module arrays with explicit shape (WTAUHF(JTOT_TAUHF), SWELLFT(IAB)): name -> (dtype, [dim expressions])


class Translator:
    def __init__(self, files, include_dirs=(REF,), registry=None, stubs=(), externals=()):
        self.routines = {}
        self.registry = registry or {}
        self.stubs = set(stubs)       # routines whose CALL is dropped (their results are not used by the caller's selected outputs)
        self.external_outs = {}           # external routine -> positions of the scalar arguments it returns (INCDATE: (0,))
        self.no_intent_outs = {}          # routine -> scalar dummies without an INTENT that it assigns (F77-style: KZEONE)
        self.externals = set(externals)   # CALLs handed to a Python callable of that name in the namespace (MPL_ALLGATHERV emulation)
        for f in files:
            self._parse_file(f if os.path.isabs(f) else os.path.join(REF, f), list(include_dirs))
        self.global_arrays = set()        # module arrays (names bound to FArr in the namespace)

    # ---- pass 1: split into routines, read declarations
    def _parse_file(self, path, incs):
        lines = _logical_lines(path, incs)
        cur, outer = None, None      # outer: the host routine while its CONTAINed procedures are read
        for ln in lines:
            if cur is not None and ln.strip() == "CONTAINS":
                outer, cur = cur, None
                continue
            if cur is None and outer is not None and re.match(r"^END\s*(SUBROUTINE|FUNCTION)\s+%s$" % outer.name, ln):
                self.routines[outer.name] = outer
                outer = None
                continue
            m = re.match(r"^(?:(?:REAL|INTEGER|LOGICAL)\s*(?:\([^)]*\))?\s+)?(SUBROUTINE|FUNCTION)\s+(\w+)\s*(?:\((.*?)\))?\s*(?:RESULT\s*\((\w+)\))?$", ln)
            if m and cur is None:
                args = [a.strip() for a in (m.group(3) or "").split(",") if a.strip()]
                cur = Routine(m.group(2), m.group(1), args)
                cur.result = m.group(4) or m.group(2)
                continue
            if cur is None:
                continue
            if re.match(r"^END\s*(SUBROUTINE|FUNCTION)?(\s+\w+)?$", ln) and not re.match(r"^END\s*(DO|IF|SELECT|WHERE|INTERFACE)", ln):
                if ln.strip() == "END" or re.match(r"^END\s*(SUBROUTINE|FUNCTION)", ln):
                    self.routines[cur.name] = cur
                    cur = None
                    continue
            cur.body.append(ln)
        for r in self.routines.values():
            self._read_decls(r)

    def _read_decls(self, r):
        body, in_iface = [], False
        for ln in r.body:
            if re.match(r"^INTERFACE\b", ln):
                in_iface = True
                continue
            if in_iface:
                if re.match(r"^END\s*INTERFACE", ln):
                    in_iface = False
                continue
            if re.match(r"^USE\s*,\s*INTRINSIC", ln):
                continue
            m = re.match(r"^USE\s+(\w+)\s*(?:,\s*ONLY\s*:\s*(.*))?$", ln)
            if m:
                r.uses[m.group(1)] = [x.strip() for x in (m.group(2) or "").split(",") if x.strip()]
                continue
            if re.match(r"^IMPLICIT\b", ln) or re.match(r"^(EXTERNAL|SAVE|INTRINSIC)\b", ln) or re.match(r"^TYPE\s*\(", ln):
                continue
            m = re.match(r"^(REAL|INTEGER|LOGICAL|CHARACTER|DOUBLE\s+PRECISION)\s*(\((?:[^()]|\([^()]*\))*\))?\s*(.*?)::\s*(.*)$", ln)
            if m:
                typ = {"REAL": float, "DOUBLE PRECISION": float, "INTEGER": int, "LOGICAL": bool, "CHARACTER": str}[re.sub(r"\s+", " ", m.group(1))]
                attrs = m.group(3)
                dims = None
                dm = re.search(r"DIMENSION\s*\(", attrs)
                if dm:
                    depth, j = 1, dm.end()
                    while depth:
                        depth += {"(": 1, ")": -1}.get(attrs[j], 0)
                        j += 1
                    dims = _split_top(attrs[dm.end(): j - 1])
                param = "PARAMETER" in attrs
                im = re.search(r"INTENT\s*\(\s*(\w+)\s*\)", attrs)
                for ent in _split_top(m.group(4)):
                    em = re.match(r"^(\w+)\s*(?:\((.*)\))?\s*(?:=\s*(.*))?$", ent)
                    nm = em.group(1)
                    d = _split_top(em.group(2)) if em.group(2) else dims
                    r.decl[nm] = dict(type=typ, dims=d, intent=im.group(1) if im else None, param=param, init=em.group(3), optional="OPTIONAL" in attrs)
                    if d:
                        r.arrays.add(nm)
                continue
            body.append(ln)
        r.body = body

    # ---- expressions
    def expr(self, r, s):
        toks = _TOK.findall(s)
        out, closers = [], []
        i = 0
        while i < len(toks):
            t = toks[i]
            nxt = toks[i + 1] if i + 1 < len(toks) else ""
            if t in _DOT:
                out.append(_DOT[t])
            elif re.match(r"^[\d.]", t) and t != ".":
                m = re.match(r"^([\d.]+(?:[ED][+-]?\d+)?)(?:_\w+)?$", t)
                if m is None:
                    raise SyntaxError("%s: bad number token %r in: %s" % (r.name, t, s))
                num = m.group(1).replace("D", "E")
                if re.match(r"^\d+$", num):
                    out.append("FInt(%s)" % num)
                else:
                    out.append("(%r)" % float(num))
            elif t in _DOT:
                out.append(_DOT[t])
            elif t == "/=":
                out.append("!=")
            elif t == "**":
                # X**N -> fpow(X, N): rewrite afterwards (needs operand boundaries); mark
                out.append("**")
            elif re.match(r"^[A-Z_]", t):
                if nxt == "(":
                    if t in r.arrays or (t in self.global_arrays and t not in getattr(r, 'decl', {})):
                        out.append(t + "[")
                        closers.append("]")
                    elif t in r.stmtfun or t in self.routines:
                        if t in self.routines and t not in r.stmtfun and t != r.name:
                            self.translate(t)
                        out.append(("SF_" if t in r.stmtfun else "") + t + "(")
                        closers.append(")")
                    elif t in INTRINSICS:
                        out.append("F_" + t + "(")
                        closers.append(")")
                    elif t == "OML_GET_MAX_THREADS":      # one thread
                        out.append("FInt(1 + 0*len(")
                        closers.append("()))")
                    else:
                        raise SyntaxError("%s: unknown function or array %s in: %s" % (r.name, t, s))
                    i += 1
                else:
                    out.append(t)
            elif t == "(":
                out.append("(")
                closers.append(")")
            elif t == ")":
                out.append(closers.pop())
            elif t.startswith("'") or t.startswith('"'):
                out.append(repr(t[1:-1]))
            elif t == "//":
                out.append("+")
            else:
                out.append(t)
            i += 1
        return _powers("".join(out))

    # ---- statements
    def translate(self, name):
        r = self.routines[name]
        if r.py is not None:
            return r.py
        # statement functions: NAME(ARGS) = EXPR with NAME a declared scalar
        lines = []
        ind = 1
        emit = lambda s: lines.append("    " * ind + s)
        outs = [a for a in r.args if a in r.decl and not r.decl[a]["dims"] and
                (r.decl[a]["intent"] in ("OUT", "INOUT") or a in self.no_intent_outs.get(name, ()))]
        r.outs = outs
        res = r.result if r.kind == "FUNCTION" else None
        ret = "return " + (res if res else ("(" + ", ".join(outs) + ("," if len(outs) == 1 else "") + ")" if outs else "None"))
        used = sorted({v for lst in r.uses.values() for v in lst if re.match(r"^[A-Z_][A-Z0-9_]*$", v)} - set(r.args) - set(r.decl))
        if used:
            emit("global " + ", ".join(used))
        rebase = []
        for nm in r.args:
            d = r.decl.get(nm)
            if d and d["dims"]:
                los = [(x.split(":")[0].strip() or "1") if ":" in x else "1" for x in d["dims"]]
                if any(lo != "1" for lo in los):
                    rebase.append((nm, los))
        # PARAMETERs and local arrays
        for nm, d in r.decl.items():
            if d["param"] and d["init"] is not None and not d["dims"]:
                emit("%s = %s" % (nm, self.expr(r, d["init"])))
        for nm, d in r.decl.items():      # Fortran leaves locals undefined; NaN makes a read-before-write visible (EPSILON(X) only asks for the kind)
            if not d["dims"] and nm not in r.args and not d["param"] and nm != getattr(r, "result", None) and d["type"] is float:
                emit("%s = float('nan')" % nm)
            elif not d["dims"] and nm not in r.args and not d["param"] and nm != getattr(r, "result", None) and d["type"] is bool:
                emit("%s = False" % nm)
        for nm, d in r.decl.items():
            if d["dims"] and nm not in r.args:
                b = []
                for x in d["dims"]:
                    if x.strip() == ":":
                        b = None
                        break
                    lo, hi = (x.split(":") + [None])[:2] if ":" in x else ("1", x)
                    b.append("(%s, %s)" % (self.expr(r, lo), self.expr(r, hi)))
                if b is not None:
                    emit("%s = FArr([%s], %s)" % (nm, ", ".join(b), {float: "float", int: "int", bool: "bool", str: "object"}[d["type"]]))
                else:
                    emit("%s = None" % nm)        # ALLOCATABLE: not allocated yet
        for nm, d in r.decl.items():      # 1-D PARAMETER arrays: (/ ... /)
            if d["dims"] and len(d["dims"]) == 1 and d["init"] and d["init"].strip().startswith("(/"):
                items = _split_top(d["init"].strip()[2:-2])
                emit("%s.a[:] = [%s]" % (nm, ", ".join(self.expr(r, x) for x in items)))
        for nm, los in rebase:
            emit("if %s is not None: %s = %s.rebase([%s])" % (nm, nm, nm, ", ".join(self.expr(r, lo) for lo in los)))
        stack = []
        body = r.body
        if any(ln.startswith("LABEL_SEG ") for ln in body):
            body = ["LABEL_SEG 0"] + list(body)
        for ln in body:
            self._stmt(r, ln, emit, lambda d: None, stack, ret, lines)
            # indentation is tracked through the `stack` list length
            ind = 1 + len(stack)
        while stack and stack[-1] in ("LABEL_LOOP", "SEG"):
            ind = 1 + len(stack)
            emit("break")
            stack.pop()
        ind = 1 + len(stack)
        emit(ret)
        sig = ", ".join(a + ("=None" if r.decl.get(a, {}).get("optional") else "") for a in r.args)
        src = "def %s(%s):\n" % (name, sig) + "\n".join(lines) + "\n"
        r.py = src
        return src

    def _stmt(self, r, ln, emit0, _unused, stack, ret, lines):
        ind = 1 + len(stack)

        def emit(s, extra=0):
            lines.append("    " * (ind + extra) + s)
        if re.search(r"\bCALL\s+IEEE_(GET|SET)_HALTING_MODE\b", ln):
            if ln.startswith("CALL"):
                return
            ln = re.sub(r"CALL\s+IEEE_\w+\s*\(.*\)$", "CONTINUE", ln)
        if re.match(r"^IF\s*\(\s*LHOOK\s*\)", ln) or re.match(r"^(WRITE|PRINT|FORMAT|CALL\s+FLUSH|CALL\s+GSTATS)\b", ln) or re.match(r"^\d+\s+FORMAT", ln):
            return
        if re.match(r"^CONTINUE$", ln) or re.match(r"^INCLUDE\b", ln):
            return
        m = re.match(r"^DATA\s+(\w+)\s*/(.*)/$", ln)
        if m:      # SAVE + DATA: initialised at every call here (the routines are called once)
            emit("%s = %s" % (m.group(1), self.expr(r, m.group(2))))
            return
        if ln.startswith("LABEL_SEG "):     # segments of a routine with forward GO TOs: `_pc` = 0 (running) or the label being skipped to
            n = ln.split()[1]
            if stack == ["SEG"]:
                emit("break"); stack.pop()
            if stack:
                raise NotImplementedError("%s: label %s inside a construct" % (r.name, n))
            if n == "0":
                lines.append("    _pc = 0")
            lines.append("    while _pc == 0 or _pc == %s:" % n); lines.append("        _pc = 0"); stack.append("SEG"); return
        if ln.startswith("GOTOSEG "):
            if "DO" in stack or "LABEL_LOOP" in stack or not stack or stack[0] != "SEG":
                raise NotImplementedError("%s: GO TO out of a loop" % r.name)
            emit("_pc = %s" % ln.split()[1]); emit("break"); return
        if ln == "LABEL_LOOP":      # loop until the end of the routine; the backward GO TO is its `continue`
            emit("while True:"); emit("pass", 1); stack.append("LABEL_LOOP"); return
        m = re.match(r"^DO\s+WHILE\s*\((.*)\)$", ln)
        if m:
            emit("while %s:" % self.expr(r, m.group(1))); emit("pass", 1); stack.append("DO"); return
        m = re.match(r"^DO\s+(\w+)\s*=\s*(.*)$", ln)
        if m:
            parts = _split_top(m.group(2))
            emit("for %s in frange(%s):" % (m.group(1), ", ".join(self.expr(r, p) for p in parts))); emit("pass", 1); stack.append("DO"); return
        if re.match(r"^END\s*DO$", ln):
            stack.pop(); return
        m = re.match(r"^IF\s*\((.*)\)\s*THEN$", ln)
        if m:
            emit("if %s:" % self.expr(r, m.group(1))); emit("pass", 1); stack.append("IF"); return
        m = re.match(r"^ELSE\s*IF\s*\((.*)\)\s*THEN$", ln)
        if m:
            lines.append("    " * (ind - 1) + "elif %s:" % self.expr(r, m.group(1))); lines.append("    " * ind + "pass"); return
        if re.match(r"^ELSE$", ln):
            lines.append("    " * (ind - 1) + "else:"); lines.append("    " * ind + "pass"); return
        if re.match(r"^END\s*IF$", ln):
            stack.pop(); return
        m = re.match(r"^SELECT\s+CASE\s*\((.*)\)$", ln)
        if m:
            emit("_sel = %s" % self.expr(r, m.group(1))); emit("if False:"); emit("pass", 1); stack.append("SELECT"); return
        m = re.match(r"^CASE\s*\((.*)\)$", ln)
        if m:
            vals = ", ".join(self.expr(r, v) for v in _split_top(m.group(1)))
            lines.append("    " * (ind - 1) + "elif _sel in (%s,):" % vals); lines.append("    " * ind + "pass"); return
        if re.match(r"^CASE\s+DEFAULT$", ln):
            lines.append("    " * (ind - 1) + "else:"); lines.append("    " * ind + "pass"); return
        if re.match(r"^END\s*SELECT$", ln):
            stack.pop(); return
        if re.match(r"^(END\s*)?WHERE\b", ln) or re.match(r"^ELSEWHERE\b", ln):
            if re.match(r"^WHERE\s*\(.*\)$", ln):
                emit("raise NotImplementedError('WHERE construct reached in %s')" % r.name)
                stack.append("WHERE-SKIP")
            elif re.match(r"^END\s*WHERE", ln):
                stack.pop()
            return
        if stack and stack[-1] == "WHERE-SKIP":
            return
        if re.match(r"^RETURN$", ln):
            emit(ret); return
        if re.match(r"^CYCLE$", ln):
            emit("continue"); return
        if re.match(r"^EXIT$", ln):
            emit("break"); return
        if re.match(r"^(CALL\s+)?(ABORT1|WAM_ABORT|ABOR1)\b", ln):
            emit("raise RuntimeError('ABORT in %s')" % r.name); return
        m = re.match(r"^IF\s*\(", ln)
        if m:   # one-line IF
            depth, j = 1, m.end()
            while depth:
                depth += {"(": 1, ")": -1}.get(ln[j], 0)
                j += 1
            cond, rest = ln[m.end(): j - 1], ln[j:].strip()
            emit("if %s:" % self.expr(r, cond))
            stack.append("IF1")
            n0 = len(lines)
            self._stmt(r, rest, None, None, stack, ret, lines)
            if len(lines) == n0:
                lines.append("    " * (ind + 1) + "pass")
            stack.pop()
            return
        m = re.match(r"^ALLOCATE\s*\((.*)\)$", ln)
        if m:
            for ent in _split_top(m.group(1)):
                em = re.match(r"^(\w+)\s*\((.*)\)$", ent)
                if not em:
                    continue          # STAT= and the like
                nm = em.group(1)
                b = []
                for x in _split_top(em.group(2)):
                    lo, hi = (x.split(":") + [None])[:2] if ":" in x else ("1", x)
                    b.append("(%s, %s)" % (self.expr(r, lo), self.expr(r, hi)))
                typ = r.decl[nm]["type"] if nm in r.decl else self.registry.get(nm, (float, True))[0]
                emit("%s = FArr([%s], %s)" % (nm, ", ".join(b), {float: "float", int: "int", bool: "bool"}[typ]))
            return
        if re.match(r"^DEALLOCATE\b", ln):
            return
        m = re.match(r"^CALL\s+(\w+)\s*(?:\((.*)\))?$", ln)
        if m:
            callee = m.group(1)
            args = _split_top(m.group(2)) if m.group(2) else []
            if callee in self.stubs:
                emit("pass")
                return
            if callee in self.externals:
                pa = []
                for a in args:
                    km = re.match(r"^(\w+)\s*=(?!=)\s*(.*)$", a)          # keyword argument (LDREPROD=..., CDSTRING=...)
                    pa.append("%s=%s" % (km.group(1), self.expr(r, km.group(2))) if km else self.expr(r, a))
                outs = [pa[k] for k in self.external_outs.get(callee, ())]
                emit("%s%s(%s)" % ((", ".join(outs) + ", = ") if outs else "", callee, ", ".join(pa)))
                return
            if callee not in self.routines:
                raise SyntaxError("%s: CALL of %s, which is not among the translated files" % (r.name, callee))
            c = self.routines[callee]
            self.translate(callee)
            pa = []
            for k, a in enumerate(args):
                am = re.match(r"^(\w+)\s*\(([^:]*)\)$", a)
                dummy = c.decl.get(c.args[k]) if k < len(c.args) else None
                sm = re.match(r"^(\w+)\s*\((.*:.*)\)$", a)
                if sm and (sm.group(1) in r.arrays or sm.group(1) in self.global_arrays) and dummy and dummy["dims"]:
                    pa.append("FArr.view(%s)" % self.expr(r, a))        # array section: the callee sees it with lower bounds 1
                elif am and (am.group(1) in r.arrays or am.group(1) in self.global_arrays) and dummy and dummy["dims"] and "," not in am.group(2):
                    pa.append("%s.elemview(%s)" % (am.group(1), self.expr(r, am.group(2))))      # A(i) passed to an array dummy
                else:
                    pa.append(self.expr(r, a))
            outs = [pa[c.args.index(o)] for o in c.outs]
            call = "%s(%s)" % (callee, ", ".join(pa))
            if outs:
                emit("%s = %s" % (", ".join(outs) + ("," if len(outs) == 1 else ""), call))
            else:
                emit(call)
            return
        # assignment
        depth, eq = 0, -1
        for j, ch in enumerate(ln):
            if ch in "([":
                depth += 1
            elif ch in ")]":
                depth -= 1
            elif ch == "=" and depth == 0 and ln[j - 1] not in "<>/=" and ln[j + 1: j + 2] != "=":
                eq = j
                break
        if eq < 0:
            raise SyntaxError("%s: cannot translate: %s" % (r.name, ln))
        lhs, rhs = ln[:eq].strip(), ln[eq + 1:].strip()
        lm = re.match(r"^(\w+)\s*(?:\((.*)\))?$", lhs)
        nm = lm.group(1)
        if lm.group(2) is not None and nm not in r.arrays and not (nm in self.global_arrays and nm not in r.decl):
            # statement function
            r.stmtfun.add(nm)
            emit("def SF_%s(%s): return %s" % (nm, lm.group(2), self.expr(r, rhs)))
            return
        if r.kind == "FUNCTION" and nm == r.result and lm.group(2) is None:
            emit("%s = %s" % (nm, self.expr(r, rhs)))
            return
        if lm.group(2) is None and (nm in r.arrays or (nm in self.global_arrays and nm not in r.decl)):
            emit("%s.assign(%s)" % (nm, self.expr(r, rhs)))
            return
        if lm.group(2) is None:      # scalar: Fortran converts to the declared type of the left-hand side
            typ = r.decl[nm]["type"] if nm in r.decl else self.registry.get(nm, (None, False))[0]
            conv = {int: "F_ICONV(%s)", float: "F_RCONV(%s)"}.get(typ, "%s")
            emit("%s = %s" % (nm, conv % self.expr(r, rhs)))
            return
        emit("%s = %s" % (self.expr(r, lhs), self.expr(r, rhs)))

    def compile(self, names, namespace):
        """Translate `names` (and everything they call) and exec the result in `namespace` (module variables must be in it)."""
        self.global_arrays = {k for k, v in namespace.items() if isinstance(v, FArr)} | {k for k, (t, a) in self.registry.items() if a}
        for n in names:
            self.translate(n)
        ns = dict(RUNTIME)
        ns.update(namespace)
        for nm, r in self.routines.items():
            if r.py is not None:
                exec(compile(r.py, "<f90:%s>" % nm, "exec"), ns)
        return ns

    def needed(self, names):
        """Module variables USEd by the translated routines (for building the namespace)."""
        need, seen, todo = {}, set(), list(names)
        while todo:
            n = todo.pop()
            if n in seen or n not in self.routines:
                continue
            seen.add(n)
            r = self.routines[n]
            for mod, lst in r.uses.items():
                if mod in ("PARKIND_WAVE", "YOMHOOK", "YOWTEST", "PARKIND1"):
                    continue
                for v in lst:
                    need.setdefault(v, set()).add(n)
            for ln in r.body:
                m = re.match(r"^(?:IF\s*\(.*\)\s*)?CALL\s+(\w+)", ln)
                if m:
                    todo.append(m.group(1))
                for f in re.findall(r"\b([A-Z_][A-Z0-9_]*)\s*\(", ln):
                    if f in self.routines:
                        todo.append(f)
        return need


def _powers(s):
    """a**b -> fpow(a, b) on the Python text (right-associative, binds tighter than unary minus on the left operand)."""
    while True:
        k = s.rfind("**")
        if k < 0:
            return s
        # right operand
        j = k + 2
        while j < len(s) and s[j] == " ":
            j += 1
        e = j
        if e < len(s) and s[e] in "+-":
            e += 1
        e = _operand_end(s, e)
        b = _operand_start(s, k)
        s = s[:b] + "fpow(" + s[b:k].strip() + ", " + s[j:e].strip() + ")" + s[e:]


def _operand_end(s, i):
    n = len(s)
    if i < n and s[i] == "(":
        return _match(s, i) + 1
    j = i
    while j < n and (s[j].isalnum() or s[j] in "_."):
        j += 1
    # scientific notation exponent sign (1.0e-05)
    if j < n and s[j] in "+-" and j > i and s[j - 1] in "eE" and re.match(r"^[\d.]+[eE]$", s[i:j]):
        j += 1
        while j < n and s[j].isdigit():
            j += 1
    while j < n and s[j] in "([":
        j = _match(s, j) + 1
    return j


def _operand_start(s, k):
    j = k - 1
    while j >= 0 and s[j] == " ":
        j -= 1
    if s[j] in ")]":
        depth = 0
        while True:
            if s[j] in ")]":
                depth += 1
            elif s[j] in "([":
                depth -= 1
                if depth == 0:
                    break
            j -= 1
        # function / array name in front of the bracket, possibly chained a[..](..)
        while True:
            i = j - 1
            while i >= 0 and (s[i].isalnum() or s[i] == "_"):
                i -= 1
            if i + 1 < j:
                j = i + 1
                break
            if j > 0 and s[j - 1] in ")]":
                j -= 1
                depth = 0
                while True:
                    if s[j] in ")]":
                        depth += 1
                    elif s[j] in "([":
                        depth -= 1
                        if depth == 0:
                            break
                    j -= 1
                continue
            break
        return j
    i = j
    while i >= 0 and (s[i].isalnum() or s[i] in "_."):
        i -= 1
    if i >= 1 and s[i] in "+-" and s[i - 1] in "eE" and re.match(r"^[\d.]+[eE]$", s[max(0, _numstart(s, i - 1)): i]):
        i = _numstart(s, i - 1) - 1
    return i + 1


def _numstart(s, i):
    while i >= 0 and (s[i].isdigit() or s[i] in ".eE"):
        i -= 1
    return i + 1


def _match(s, i):
    depth = 0
    for j in range(i, len(s)):
        if s[j] in "([":
            depth += 1
        elif s[j] in ")]":
            depth -= 1
            if depth == 0:
                return j
    raise SyntaxError("unbalanced: " + s)
