"""Clexulator source -> flat evaluation tables.

The reference runtime-compiles one CASM-generated C++ ``Clexulator`` per basis
set and calls it through virtual dispatch, one site at a time
(e.g. ``tests/unit/clexmonte/data/FCC_binary_vacancy/basis_sets/bset.default/
FCC_binary_vacancy_Clexulator_default.cc``; grammar census in SURVEY.md
Appendix A).  This module replaces the *compile + dlopen* step: it parses the
same generated source into flat tables that the sm_100a kernels evaluate.

What is extracted (reference file:line refer to the FCC default clexulator):

* sizes ``BaseClexulator(nlist_size, corr_size, n_point_corr)``          :306
* site basis functions ``m_occ_func_<b>_<f>[occ] = value``               :307-311
* function tables ``m_orbit_func_table_0`` / ``m_flower_func_table_0`` /
  ``m_delta_func_table_0`` (which (point, corr) pairs are zero)          :317-372
* neighbor-list metadata: weight matrix, sublattice set, neighborhood    :374-392
* per-corr ``m_orbit_site_neighborhood``                                 :394-439
* every ``eval_bfunc_*`` / ``site_eval_bfunc_*`` / ``site_deval_bfunc_*``
  body, kept as an expression in the *same association order* as the C++
  (so the device evaluator can reproduce the reference's rounding exactly):

      function := group (+ group)*
      group    := [dphi *] ( elem + elem + ... ) [/ div]      | dphi | term
      elem     := term | ( term + term + ... )
      term     := [coef *] occ_func(n) [* occ_func(n)]*        | number
      dphi     := (m_occ_func_b_f[occ_f] - m_occ_func_b_f[occ_i])

Literals are kept verbatim (the generated code carries 6-digit coefficients
such as 0.707107).  Anything outside this grammar
raises :class:`ClexulatorParseError` -- no silent approximation.

The prim neighbor list follows the convention verified against the generated
kernels: all unit cells with r^T W r <= R_max, sorted
by (r^T W r, lexicographic (i, j, k)); neighbor index
``n = cell_rank * n_nlist_sublat + position_of_sublattice``.
"""
from __future__ import annotations

import io
import re
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


class ClexulatorParseError(ValueError):
    """The generated source does not follow the recognised grammar."""


# ---------------------------------------------------------------------------
# expression parsing
# ---------------------------------------------------------------------------
_TOKEN = re.compile(
    r"\s*(?:"
    r"(?P<of>occ_func_(?P<ofb>\d+)_(?P<off>\d+)\(\s*(?P<ofn>\d+)\s*\))"
    r"|(?P<mo>m_occ_func_(?P<mob>\d+)_(?P<mof>\d+)\[\s*occ_(?P<mow>[if])\s*\])"
    r"|(?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?)"
    r"|(?P<op>[-+*/()])"
    r")"
)


def _tokenize(text: str):
    pos = 0
    out = []
    text = text.strip()
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m or m.end() == pos:
            raise ClexulatorParseError(f"unrecognised token at: {text[pos:pos + 60]!r}")
        pos = m.end()
        if m.group("of"):
            out.append(("of", int(m.group("ofb")), int(m.group("off")), int(m.group("ofn"))))
        elif m.group("mo"):
            out.append(("mo", int(m.group("mob")), int(m.group("mof")), m.group("mow")))
        elif m.group("num"):
            out.append(("num", float(m.group("num"))))
        else:
            out.append(("op", m.group("op")))
    return out


class _Parser:
    """Recursive descent with C++ precedence / left associativity."""

    def __init__(self, toks):
        self.t = toks
        self.i = 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def take(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def parse(self):
        e = self.sum()
        if self.peek() is not None:
            raise ClexulatorParseError(f"trailing tokens: {self.t[self.i:self.i + 5]}")
        return e

    def sum(self):
        left = self.mul()
        while self.peek() in (("op", "+"), ("op", "-")):
            op = self.take()[1]
            right = self.mul()
            left = (op, left, right)
        return left

    def mul(self):
        left = self.atom()
        while self.peek() in (("op", "*"), ("op", "/")):
            op = self.take()[1]
            right = self.atom()
            left = (op, left, right)
        return left

    def atom(self):
        tok = self.peek()
        if tok is None:
            raise ClexulatorParseError("unexpected end of expression")
        if tok == ("op", "("):
            self.take()
            e = self.sum()
            if self.peek() != ("op", ")"):
                raise ClexulatorParseError("missing ')'")
            self.take()
            return ("paren", e)
        if tok[0] in ("of", "mo", "num"):
            return self.take()
        raise ClexulatorParseError(f"unexpected token {tok}")


def parse_expression(text: str):
    """Parse a generated function body (the text between ``return`` and ``;``)."""
    return _Parser(_tokenize(text)).parse()


def eval_ast(ast, occ_func, m_occ_func, occ_i=None, occ_f=None) -> float:
    """Evaluate an AST with IEEE double arithmetic in the C++ order.

    ``occ_func(b, f, n)`` returns the site function value at neighbor ``n``;
    ``m_occ_func(b, f, occ)`` the table value.  Used by the tests as a third,
    independent statement of the arithmetic (besides the reference kernels and CUDA).
    """
    k = ast[0]
    if k == "num":
        return float(ast[1])
    if k == "of":
        return float(occ_func(ast[1], ast[2], ast[3]))
    if k == "mo":
        return float(m_occ_func(ast[1], ast[2], occ_f if ast[3] == "f" else occ_i))
    if k == "paren":
        return eval_ast(ast[1], occ_func, m_occ_func, occ_i, occ_f)
    a = eval_ast(ast[1], occ_func, m_occ_func, occ_i, occ_f)
    b = eval_ast(ast[2], occ_func, m_occ_func, occ_i, occ_f)
    if k == "+":
        return a + b
    if k == "-":
        return a - b
    if k == "*":
        return a * b
    if k == "/":
        return a / b
    raise ClexulatorParseError(f"bad node {k}")


# -- canonical form ----------------------------------------------------------
def _flatten(ast, op):
    """Left-associative chain  ((a op b) op c) -> [a, b, c]."""
    if ast[0] == op:
        return _flatten(ast[1], op) + [ast[2]]
    return [ast]


def _is_dphi(ast) -> Optional[Tuple[int, int]]:
    if ast[0] != "paren":
        return None
    e = ast[1]
    if e[0] == "-" and e[1][0] == "mo" and e[2][0] == "mo":
        a, b = e[1], e[2]
        if a[3] == "f" and b[3] == "i" and a[1:3] == b[1:3]:
            return (a[1], a[2])
    return None


def _term(ast):
    """term := [coef *] of [* of]*  | number   ->  (coef, [(b,f,n)...])"""
    parts = _flatten(ast, "*")
    coef = 1.0
    factors = []
    for k, p in enumerate(parts):
        if p[0] == "num":
            if k != 0:
                raise ClexulatorParseError("numeric literal not in leading position of a product")
            coef = float(p[1])
        elif p[0] == "of":
            factors.append((p[1], p[2], p[3]))
        else:
            raise ClexulatorParseError(f"unsupported factor in product: {p[0]}")
    return coef, factors


def _elem(ast):
    """elem := term | ( term + term ... )  -> list of terms"""
    if ast[0] == "paren":
        return [_term(t) for t in _flatten(ast[1], "+")]
    return [_term(ast)]


def _group(ast):
    """-> dict(dphi=(b,f)|None, has_sum, div, elems=[[term...]...])"""
    div = 0.0
    if ast[0] == "/":
        if ast[2][0] != "num":
            raise ClexulatorParseError("division by a non-literal")
        div = float(ast[2][1])
        ast = ast[1]
    parts = _flatten(ast, "*")
    dphi = _is_dphi(parts[0])
    if dphi is not None:
        rest = parts[1:]
        if not rest:
            return dict(dphi=dphi, has_sum=False, div=div, elems=[])
        if len(rest) == 1 and rest[0][0] == "paren":
            elems = [_elem(e) for e in _flatten(rest[0][1], "+")]
            return dict(dphi=dphi, has_sum=True, div=div, elems=elems)
        raise ClexulatorParseError("unsupported delta-function product shape")
    if len(parts) == 1 and parts[0][0] == "paren":
        elems = [_elem(e) for e in _flatten(parts[0][1], "+")]
        return dict(dphi=None, has_sum=True, div=div, elems=elems)
    # bare product / literal / single site function
    return dict(dphi=None, has_sum=True, div=div, elems=[[_term(ast)]])


def canonicalize(ast):
    """function := group (+ group)*   (see module docstring)."""
    return [_group(g) for g in _flatten(ast, "+")]


def eval_canonical(groups, occ_func, m_occ_func, occ_i=None, occ_f=None) -> float:
    """Evaluate the canonical form exactly the way the device evaluator does."""
    total = None
    for g in groups:
        s = None
        if g["has_sum"]:
            for elem in g["elems"]:
                ev = None
                for coef, factors in elem:
                    tv = coef
                    for (b, f, n) in factors:
                        tv = tv * occ_func(b, f, n)
                    ev = tv if ev is None else ev + tv
                s = ev if s is None else s + ev
        if g["dphi"] is not None:
            b, f = g["dphi"]
            d = m_occ_func(b, f, occ_f) - m_occ_func(b, f, occ_i)
            v = d * s if g["has_sum"] else d
        else:
            v = s
        if g["div"] != 0.0:
            v = v / g["div"]
        total = v if total is None else total + v
    return 0.0 if total is None else total


# ---------------------------------------------------------------------------
# neighbor list
# ---------------------------------------------------------------------------
def prim_neighbor_cells(weight_matrix: np.ndarray, neighborhood: Sequence[Sequence[int]],
                        rmax: Optional[int] = None) -> np.ndarray:
    """Ordered unit cells of the prim neighbor list."""
    W = np.asarray(weight_matrix, dtype=np.int64)
    nb = np.asarray(neighborhood, dtype=np.int64).reshape(-1, 3)

    def quad(r):
        return np.einsum("...a,ab,...b->...", r, W, r)

    r0 = int(quad(nb).max()) if len(nb) else 0
    if rmax is None or rmax < r0:
        rmax = r0
    B = int(np.abs(nb).max()) + 1 if len(nb) else 1
    while True:
        rng = np.arange(-B, B + 1)
        g = np.stack(np.meshgrid(rng, rng, rng, indexing="ij"), axis=-1).reshape(-1, 3)
        surf = g[np.abs(g).max(axis=1) == B]
        if (quad(surf) > rmax).all():
            break
        B += 1
    q = quad(g)
    keep = q <= rmax
    g, q = g[keep], q[keep]
    order = np.lexsort((g[:, 2], g[:, 1], g[:, 0], q))
    return g[order]


# ---------------------------------------------------------------------------
# tables
# ---------------------------------------------------------------------------
KIND_GLOBAL, KIND_POINT, KIND_DELTA = 0, 1, 2


@dataclass
class ClexulatorTables:
    """Flat tables for one basis set (the wire format of ``cmx_tables_create``)."""

    name: str
    nlist_size: int          # as declared by the generated source
    corr_size: int
    n_point_corr: int
    n_sublat: int
    max_occ: int
    n_func: int
    weight_matrix: np.ndarray            # (3,3) int64
    nlist_sublat: np.ndarray             # (n_nlist_sublat,) int32, sorted
    n_occ: np.ndarray                    # (n_sublat,) int32
    phi: np.ndarray                      # (n_sublat, n_func, max_occ) float64
    nbr: np.ndarray                      # (nlist_len, 4) int32: di, dj, dk, b
    # expression pool (CSR)
    factor_f: np.ndarray                 # int32
    factor_n: np.ndarray                 # int32
    term_coef: np.ndarray                # float64
    term_fbeg: np.ndarray                # int32 (n_terms+1)
    elem_tbeg: np.ndarray                # int32 (n_elems+1)
    group_ebeg: np.ndarray               # int32 (n_groups+1)
    group_dphi: np.ndarray               # int32: function index f' or -1
    group_has_sum: np.ndarray            # int32 0/1
    group_div: np.ndarray                # float64, 0 = no division
    # function -> group range; index = corr                     (global)
    #                          index = p * corr_size + corr      (point, delta)
    global_gbeg: np.ndarray              # int32 (corr_size+1)
    point_gbeg: np.ndarray               # int32 (n_point_corr*corr_size+1)
    delta_gbeg: np.ndarray               # int32 (n_point_corr*corr_size+1)
    # per-corr orbit site neighborhoods (b,i,j,k), CSR over corr
    orbit_nbhd_beg: np.ndarray           # int32 (corr_size+1)
    orbit_nbhd: np.ndarray               # (n,4) int32
    corr_orbit: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))   # corr -> orbit index
    corr_func: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))    # corr -> function index
    is_local: bool = False

    # -- convenience -------------------------------------------------------
    @property
    def nlist_len(self) -> int:
        return int(self.nbr.shape[0])

    @property
    def n_nlist_sublat(self) -> int:
        return int(self.nlist_sublat.shape[0])

    def point_sublat(self, p: int) -> int:
        """Sublattice of point-correlation index ``p``."""
        return int(self.nbr[p, 3])

    _ARRAYS = ("weight_matrix nlist_sublat n_occ phi nbr factor_f factor_n term_coef term_fbeg "
               "elem_tbeg group_ebeg group_dphi group_has_sum group_div global_gbeg point_gbeg "
               "delta_gbeg orbit_nbhd_beg orbit_nbhd corr_orbit corr_func").split()
    _SCALARS = "name nlist_size corr_size n_point_corr n_sublat max_occ n_func is_local".split()

    def save_flat(self, path) -> None:
        """The tables as one flat little-endian file, the form `cmx_tables_create_from_file`
        reads (a C++ plugin has no numpy): magic "CMXT1\0\0\0", the eleven int32 sizes of
        `cmx_table_desc` (n_sublat, max_occ, n_func, corr_size, n_point_corr, nlist_len,
        n_nlist_sublat, n_factors, n_terms, n_elems, n_groups), then its sixteen arrays in
        declaration order, each 8-byte aligned."""
        import struct
        sizes = [self.n_sublat, self.max_occ, self.n_func, self.corr_size, self.n_point_corr, self.nlist_len,
                 self.n_nlist_sublat, len(self.factor_f), len(self.term_coef), len(self.elem_tbeg) - 1,
                 len(self.group_div)]
        order = [("nlist_sublat", np.int32), ("n_occ", np.int32), ("phi", np.float64), ("nbr", np.int32),
                 ("factor_f", np.int32), ("factor_n", np.int32), ("term_coef", np.float64), ("term_fbeg", np.int32),
                 ("elem_tbeg", np.int32), ("group_ebeg", np.int32), ("group_dphi", np.int32),
                 ("group_has_sum", np.int32), ("group_div", np.float64), ("global_gbeg", np.int32),
                 ("point_gbeg", np.int32), ("delta_gbeg", np.int32)]
        with open(path, "wb") as f:
            f.write(b"CMXT1\0\0\0")
            f.write(struct.pack("<11i", *[int(x) for x in sizes]))
            f.write(b"\0" * 4)
            for name, dt in order:
                b = np.ascontiguousarray(getattr(self, name), dtype=dt).tobytes()
                f.write(b + b"\0" * (-len(b) % 8))

    def save(self, path) -> None:
        d = {k: getattr(self, k) for k in self._ARRAYS}
        for k in self._SCALARS:
            d["_" + k] = np.array(getattr(self, k))
        np.savez_compressed(path, **d)

    @classmethod
    def load(cls, path) -> "ClexulatorTables":
        z = np.load(path, allow_pickle=False)
        kw = {k: z[k] for k in cls._ARRAYS}
        for k in cls._SCALARS:
            v = z["_" + k]
            kw[k] = str(v) if k == "name" else (bool(v) if k == "is_local" else int(v))
        return cls(**kw)

    # -- statistics used for the roofline bookkeeping -----------------------
    def delta_work(self, p: int, corr_indices: Sequence[int]) -> Dict[str, int]:
        """Exact term / factor / distinct-neighbor counts of a restricted delta evaluation."""
        n_terms = n_factors = 0
        nbrs = set()
        for c in corr_indices:
            fi = p * self.corr_size + int(c)
            for g in range(self.delta_gbeg[fi], self.delta_gbeg[fi + 1]):
                for e in range(self.group_ebeg[g], self.group_ebeg[g + 1]):
                    for t in range(self.elem_tbeg[e], self.elem_tbeg[e + 1]):
                        n_terms += 1
                        lo, hi = self.term_fbeg[t], self.term_fbeg[t + 1]
                        n_factors += hi - lo
                        nbrs.update(int(x) for x in self.factor_n[lo:hi])
        return dict(terms=n_terms, factors=n_factors, neighbors=len(nbrs))


_FUNC_RE = re.compile(
    r"Scalar\s+\w+::(?P<name>eval_bfunc_(?P<o>\d+)_(?P<f>\d+)"
    r"|site_eval_bfunc_(?P<so>\d+)_(?P<sf>\d+)_at_(?P<sp>\d+)"
    r"|site_deval_bfunc_(?P<do>\d+)_(?P<df>\d+)_at_(?P<dp>\d+))"
    r"\s*\([^)]*\)\s*const\s*\{\s*return\s+(?P<body>.*?);\s*\}",
    re.S,
)


def _strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    return src


def parse_clexulator_source(source, name: Optional[str] = None,
                            rmax: Optional[int] = None) -> ClexulatorTables:
    """Parse a CASM-generated Clexulator ``.cc`` (path or text) into tables.

    ``rmax`` optionally grows the neighbor list beyond the clexulator's own
    neighborhood (the reference shares one expanding ``prim_neighbor_list``
    between basis sets, include/casm/clexmonte/system/System.hh:87-94; indices
    are stable under growth so this never changes the meaning of an index).
    """
    if isinstance(source, (str, Path)) and "\n" not in str(source) and Path(str(source)).exists():
        path = Path(str(source))
        text = path.read_text()
        if name is None:
            name = path.stem
    else:
        text = str(source)
    text = _strip_comments(text)
    if name is None:
        name = "clexulator"

    m = re.search(r":\s*BaseClexulator\(\s*(\d+)\s*,\s*(\d+)\s*,\s*(\d+)\s*\)", text)
    if not m:
        raise ClexulatorParseError("BaseClexulator(nlist, corr, npoint) constructor call not found")
    nlist_size, corr_size, n_point_corr = map(int, m.groups())

    # --- site basis function tables
    decl = {(int(b), int(f)): int(k)
            for b, f, k in re.findall(r"double\s+m_occ_func_(\d+)_(\d+)\[(\d+)\]\s*;", text)}
    vals = re.findall(r"m_occ_func_(\d+)_(\d+)\[(\d+)\]\s*=\s*([-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?)", text)
    if not decl or not vals:
        raise ClexulatorParseError("no m_occ_func tables found (not an occupation clexulator?)")

    m = re.search(r"m_n_sublattices\s*=\s*(\d+)\s*;", text)
    if not m:
        raise ClexulatorParseError("m_n_sublattices not found")
    n_sublat = int(m.group(1))
    m = re.search(r"m_sublat_indices\s*=\s*std::set<int>\s*\{([^}]*)\}", text)
    if not m:
        raise ClexulatorParseError("m_sublat_indices not found")
    nlist_sublat = np.array(sorted(int(x) for x in m.group(1).split(",") if x.strip()), dtype=np.int32)

    max_occ = max(decl.values())
    n_func = max(f for (_, f) in decl) + 1
    n_occ = np.ones(n_sublat, dtype=np.int32)
    phi = np.zeros((n_sublat, n_func, max_occ), dtype=np.float64)
    seen = set()
    for b, f, k, v in vals:
        b, f, k = int(b), int(f), int(k)
        if (b, f) not in decl or k >= decl[(b, f)]:
            raise ClexulatorParseError(f"m_occ_func_{b}_{f}[{k}] assigned but not declared")
        phi[b, f, k] = float(v)
        seen.add((b, f, k))
    for (b, f), k in decl.items():
        n_occ[b] = max(n_occ[b], k)
        for kk in range(k):
            if (b, f, kk) not in seen:
                raise ClexulatorParseError(f"m_occ_func_{b}_{f}[{kk}] never assigned")

    # --- neighbor list metadata
    W = np.zeros((3, 3), dtype=np.int64)
    rows = re.findall(r"m_weight_matrix\.row\((\d)\)\s*<<\s*(-?\d+)\s*,\s*(-?\d+)\s*,\s*(-?\d+)\s*;", text)
    if len(rows) != 3:
        raise ClexulatorParseError("m_weight_matrix rows not found")
    for r, a, b, c in rows:
        W[int(r)] = (int(a), int(b), int(c))
    m = re.search(r"m_neighborhood\s*=\s*std::set<xtal::UnitCell>\s*\{(.*?)\}\s*;", text, re.S)
    if not m:
        raise ClexulatorParseError("m_neighborhood not found")
    neighborhood = [tuple(map(int, t)) for t in
                    re.findall(r"UnitCell\(\s*(-?\d+)\s*,\s*(-?\d+)\s*,\s*(-?\d+)\s*\)", m.group(1))]
    cells = prim_neighbor_cells(W, neighborhood, rmax)
    nsub = len(nlist_sublat)
    nbr = np.zeros((len(cells) * nsub, 4), dtype=np.int32)
    for r, c in enumerate(cells):
        for p, b in enumerate(nlist_sublat):
            nbr[r * nsub + p] = (c[0], c[1], c[2], b)
    if len(nbr) < nlist_size:
        raise ClexulatorParseError(
            f"neighbor list rule yields {len(nbr)} sites < declared nlist_size {nlist_size}")

    # --- orbit site neighborhoods (with `= m_orbit_site_neighborhood[j]` aliases)
    osn: Dict[int, List[Tuple[int, int, int, int]]] = {}
    for mm in re.finditer(
            r"m_orbit_site_neighborhood\[(\d+)\]\s*=\s*(?:std::set<xtal::UnitCellCoord>\s*\{(.*?)\}|"
            r"m_orbit_site_neighborhood\[(\d+)\])\s*;", text, re.S):
        i = int(mm.group(1))
        if mm.group(3) is not None:
            osn[i] = osn[int(mm.group(3))]
        else:
            osn[i] = [tuple(map(int, t)) for t in re.findall(
                r"UnitCellCoord\(\s*(-?\d+)\s*,\s*(-?\d+)\s*,\s*(-?\d+)\s*,\s*(-?\d+)\s*\)", mm.group(2))]
    onb_beg = np.zeros(corr_size + 1, dtype=np.int32)
    onb = []
    for c in range(corr_size):
        onb.extend(osn.get(c, []))
        onb_beg[c + 1] = len(onb)
    orbit_nbhd = np.array(onb, dtype=np.int32).reshape(-1, 4)

    # --- function bodies
    bodies: Dict[str, str] = {}
    for mm in _FUNC_RE.finditer(text):
        bodies[mm.group("name")] = mm.group("body")

    # --- function tables
    def table(kind_name, two_d):
        out = {}
        if two_d:
            pat = rf"{kind_name}\[(\d+)\]\[(\d+)\]\s*=\s*&\w+::\s*(\w+?)\s*<\s*double\s*>\s*;"
            for p, i, fn in re.findall(pat, text):
                out[(int(p), int(i))] = fn
        else:
            pat = rf"{kind_name}\[(\d+)\]\s*=\s*&\w+::\s*(\w+?)\s*<\s*double\s*>\s*;"
            for i, fn in re.findall(pat, text):
                out[int(i)] = fn
        return out

    orbit_tab = table("m_orbit_func_table_0", False)
    flower_tab = table("m_flower_func_table_0", True)
    delta_tab = table("m_delta_func_table_0", True)
    if len(orbit_tab) != corr_size:
        raise ClexulatorParseError(f"m_orbit_func_table_0 has {len(orbit_tab)} entries, expected {corr_size}")
    if len(flower_tab) != n_point_corr * corr_size or len(delta_tab) != n_point_corr * corr_size:
        raise ClexulatorParseError("flower/delta function tables incomplete")

    _check_prepare_map(text, nbr)
    pool = _Pool(nbr, nlist_size)
    corr_orbit = np.zeros(corr_size, dtype=np.int32)
    corr_func = np.zeros(corr_size, dtype=np.int32)

    def emit(fn_name):
        if fn_name == "zero_func":
            return
        if fn_name not in bodies:
            raise ClexulatorParseError(f"function body for {fn_name} not found")
        try:
            pool.add_function(canonicalize(parse_expression(bodies[fn_name])))
        except ClexulatorParseError as e:
            raise ClexulatorParseError(f"{fn_name}: {e}") from None

    global_gbeg = [0]
    for c in range(corr_size):
        fn = orbit_tab[c]
        mm = re.fullmatch(r"eval_bfunc_(\d+)_(\d+)", fn)
        if not mm:
            raise ClexulatorParseError(f"unexpected orbit function name {fn}")
        corr_orbit[c], corr_func[c] = int(mm.group(1)), int(mm.group(2))
        emit(fn)
        global_gbeg.append(pool.n_groups)
    point_gbeg = [pool.n_groups]
    for p in range(n_point_corr):
        for c in range(corr_size):
            emit(flower_tab[(p, c)])
            point_gbeg.append(pool.n_groups)
    delta_gbeg = [pool.n_groups]
    for p in range(n_point_corr):
        for c in range(corr_size):
            fn = delta_tab[(p, c)]
            if fn != "zero_func":
                # the changing site's sublattice must be the point's sublattice
                g = canonicalize(parse_expression(bodies[fn]))
                for grp in g:
                    if grp["dphi"] is not None and grp["dphi"][0] != nbr[p, 3]:
                        raise ClexulatorParseError(f"{fn}: delta on sublattice {grp['dphi'][0]} "
                                                   f"but point {p} is on sublattice {nbr[p, 3]}")
            emit(fn)
            delta_gbeg.append(pool.n_groups)

    return ClexulatorTables(
        name=name, nlist_size=nlist_size, corr_size=corr_size, n_point_corr=n_point_corr,
        n_sublat=n_sublat, max_occ=max_occ, n_func=n_func, weight_matrix=W,
        nlist_sublat=nlist_sublat, n_occ=n_occ, phi=phi, nbr=nbr,
        factor_f=np.array(pool.factor_f, dtype=np.int32), factor_n=np.array(pool.factor_n, dtype=np.int32),
        term_coef=np.array(pool.term_coef, dtype=np.float64),
        term_fbeg=np.array(pool.term_fbeg, dtype=np.int32),
        elem_tbeg=np.array(pool.elem_tbeg, dtype=np.int32),
        group_ebeg=np.array(pool.group_ebeg, dtype=np.int32),
        group_dphi=np.array(pool.group_dphi, dtype=np.int32),
        group_has_sum=np.array(pool.group_has_sum, dtype=np.int32),
        group_div=np.array(pool.group_div, dtype=np.float64),
        global_gbeg=np.array(global_gbeg, dtype=np.int32),
        point_gbeg=np.array(point_gbeg, dtype=np.int32),
        delta_gbeg=np.array(delta_gbeg, dtype=np.int32),
        orbit_nbhd_beg=onb_beg, orbit_nbhd=orbit_nbhd,
        corr_orbit=corr_orbit, corr_func=corr_func,
        is_local=(n_point_corr == nlist_size and n_point_corr > len(nlist_sublat)),
    )


def _check_prepare_map(text: str, nbr: np.ndarray) -> None:
    """ParamPack slot (f, n) must be filled from table m_occ_func_<b(n)>_<f>.

    ``_global_prepare`` / ``_point_prepare`` contain one statement per slot:
    ``ParamPack::Val<Scalar>::set(m_params, key, f, n, eval_occ_func_<b>_<f'>(n'))``
    (FCC default clexulator :612-778).  The device tables index the site function
    as phi[sublattice of neighbor n][f], which is only right if b == b(n), f' == f
    and n' == n for every slot.
    """
    pat = re.compile(r"set\(\s*m_params\s*,\s*m_occ_site_func_param_key\s*,\s*(\d+)\s*,\s*(\d+)\s*,"
                     r"\s*eval_occ_func_(\d+)_(\d+)\(\s*(\d+)\s*\)\s*\)")
    found = False
    for f, n, b, f2, n2 in pat.findall(text):
        found = True
        f, n, b, f2, n2 = int(f), int(n), int(b), int(f2), int(n2)
        if n >= len(nbr):
            raise ClexulatorParseError(f"prepare: neighbor {n} outside the neighbor list")
        if n2 != n or f2 != f or nbr[n, 3] != b:
            raise ClexulatorParseError(
                f"prepare: slot (f={f}, n={n}) filled from eval_occ_func_{b}_{f2}({n2}) but neighbor "
                f"{n} is on sublattice {nbr[n, 3]}")
    if not found:
        raise ClexulatorParseError("no ParamPack prepare statements found")


class _Pool:
    def __init__(self, nbr, nlist_size):
        self.nbr = nbr
        self.nlist_size = nlist_size
        self.factor_f: List[int] = []
        self.factor_n: List[int] = []
        self.term_coef: List[float] = []
        self.term_fbeg: List[int] = [0]
        self.elem_tbeg: List[int] = [0]
        self.group_ebeg: List[int] = [0]
        self.group_dphi: List[int] = []
        self.group_has_sum: List[int] = []
        self.group_div: List[float] = []

    @property
    def n_groups(self):
        return len(self.group_dphi)

    def add_function(self, groups):
        for g in groups:
            for elem in g["elems"]:
                for coef, factors in elem:
                    for (b, f, n) in factors:
                        if n >= self.nlist_size and n >= len(self.nbr):
                            raise ClexulatorParseError(f"neighbor index {n} outside the neighbor list")
                        # NOTE: the <b> in the accessor name occ_func_<b>_<f> is the
                        # asymmetric unit's representative sublattice; the accessor
                        # only reads ParamPack row f, column n.  Which table filled
                        # that slot is checked against _global_prepare (see
                        # _check_prepare_map).
                        self.factor_f.append(f)
                        self.factor_n.append(n)
                    self.term_coef.append(coef)
                    self.term_fbeg.append(len(self.factor_f))
                self.elem_tbeg.append(len(self.term_coef))
            self.group_ebeg.append(len(self.elem_tbeg) - 1)
            self.group_dphi.append(-1 if g["dphi"] is None else g["dphi"][1])
            self.group_has_sum.append(1 if g["has_sum"] else 0)
            self.group_div.append(g["div"])


# ---------------------------------------------------------------------------
# ECI readers ("ECI inputs")
# ---------------------------------------------------------------------------
def read_eci(data, corr_size: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Return (index uint32[], value float64[]) from either reference ECI format.

    * sparse: ``[[index, value], ...]`` (formation_energy_sparse_eci.json)
    * dense : basis.json + ``"eci"`` inside ``orbits[*].cluster_functions[*]``
      keyed by ``linear_function_index`` (formation_energy_eci.json)
    Only non-zero... no: every *listed* coefficient is kept, in file order, as
    the reference's SparseCoefficients does [EXT].
    """
    if isinstance(data, (str, Path)):
        import json
        data = json.loads(Path(data).read_text())
    idx: List[int] = []
    val: List[float] = []
    if isinstance(data, list):
        for i, v in data:
            idx.append(int(i))
            val.append(float(v))
    elif isinstance(data, dict) and "orbits" in data:
        for orb in data["orbits"]:
            for cf in orb.get("cluster_functions", []):
                if "eci" in cf:
                    idx.append(int(cf["linear_function_index"]))
                    val.append(float(cf["eci"]))
    else:
        raise ValueError("unrecognised ECI format")
    if corr_size is not None and idx and max(idx) >= corr_size:
        raise ValueError(f"ECI index {max(idx)} out of range for corr_size {corr_size}")
    return np.array(idx, dtype=np.uint32), np.array(val, dtype=np.float64)


# ---------------------------------------------------------------------------
# basis.json reader: the reference's own description of the basis set, used to cross-check
# an exported table
# ---------------------------------------------------------------------------
def read_basis_json(data) -> Dict:
    """The parts of a CASM ``basis.json`` (written next to the generated Clexulator,
    e.g. tests/unit/clexmonte/data/FCC_binary_vacancy/basis_sets/bset.default/basis.json)
    that describe what the generated source computes:

    * ``phi[b][f][occ]``: occupation site basis functions per sublattice, occupants in the
      order of ``prim.basis[b].occupants`` (``site_functions[*].occ.basis``);
    * ``orbits``: ``linear_orbit_index``, multiplicity, prototype cluster sites ``(b, i, j, k)``
      and the ``linear_function_index`` of every cluster function of the orbit.
    """
    if isinstance(data, (str, Path)):
        import json
        data = json.loads(Path(data).read_text())
    occupants = [list(site["occupants"]) for site in data["prim"]["basis"]]
    phi: Dict[int, List[List[float]]] = {}
    for sf in data.get("site_functions", []):
        b = int(sf["sublat"])
        basis = (sf.get("occ") or {}).get("basis") or {}

        def f_index(key: str) -> int:       # "\\phi_{b,f}"
            m = re.search(r"\{\s*\d+\s*,\s*(\d+)\s*\}", key)
            if not m:
                raise ValueError(f"basis.json: unrecognised site function name {key!r}")
            return int(m.group(1))
        rows = sorted(((f_index(k), v) for k, v in basis.items()), key=lambda kv: kv[0])
        phi[b] = [[float(v[name]) for name in occupants[b]] for _, v in rows]
    orbits = []
    for orb in data["orbits"]:
        orbits.append(dict(
            index=int(orb["linear_orbit_index"]), mult=int(orb["mult"]),
            sites=[tuple(int(x) for x in site) for site in orb["prototype"]["sites"]],
            functions=([int(cf["linear_function_index"]) for cf in orb["cluster_functions"]]
                       if "cluster_functions" in orb else None)))     # None: a clust.json (orbits only)
    return dict(occupants=occupants, phi=phi, orbits=orbits)


def check_tables_against_basis(t: "ClexulatorTables", basis) -> None:
    """Cross-check exported tables against the project's ``basis.json`` (or ``clust.json``: the
    same orbits without the functions): number of functions,
    occupants per sublattice, the site basis functions (to the 6 digits the generated source
    prints), the orbit of every correlation, and per global function the cluster size (factors
    per term), the multiplicity (the divisor, and a whole number of terms per equivalent
    cluster).  Raises ValueError listing every mismatch."""
    if not isinstance(basis, dict) or "orbits" not in basis or "phi" not in basis:
        basis = read_basis_json(basis)
    bad: List[str] = []
    if any(o["functions"] is None for o in basis["orbits"]):
        # a clust.json: orbits and prototype clusters only -- the functions of an orbit are the
        # ones the generated source assigns to it
        if not len(t.corr_orbit):
            raise ValueError("tables without corr_orbit cannot be checked against a clust.json")
        for o in basis["orbits"]:
            o["functions"] = [c for c in range(t.corr_size) if int(t.corr_orbit[c]) == o["index"]]
        if sorted(set(int(x) for x in t.corr_orbit)) != sorted(o["index"] for o in basis["orbits"] if o["functions"]):
            bad.append("the orbits of the source's functions are not the orbits of clust.json")
    n_functions = sum(len(o["functions"]) for o in basis["orbits"])
    if n_functions != t.corr_size:
        bad.append(f"basis.json lists {n_functions} cluster functions, the source declares corr_size {t.corr_size}")
    for b, occ in enumerate(basis["occupants"]):
        if b < t.n_sublat and len(occ) != int(t.n_occ[b]):
            bad.append(f"sublattice {b}: {len(occ)} occupants in basis.json, {int(t.n_occ[b])} in the tables")
    for b, rows in basis["phi"].items():
        for f, row in enumerate(rows):
            got = t.phi[b, f, :len(row)] if b < t.n_sublat and f < t.n_func else None
            if got is None or not np.allclose(got, row, rtol=0.0, atol=5e-7):
                bad.append(f"site function phi_{{{b},{f}}}: basis.json {row}, tables {None if got is None else got.tolist()}")
    for orb in basis["orbits"]:
        for c in orb["functions"]:
            if c >= t.corr_size:
                continue
            if len(t.corr_orbit) and int(t.corr_orbit[c]) != orb["index"]:
                bad.append(f"function {c}: orbit {int(t.corr_orbit[c])} in the source, {orb['index']} in basis.json")
            n_terms = 0
            for g in range(int(t.global_gbeg[c]), int(t.global_gbeg[c + 1])):
                if orb["sites"] and t.group_div[g] not in (0.0, float(orb["mult"])):
                    bad.append(f"function {c}: divisor {t.group_div[g]} in the source, multiplicity {orb['mult']} in basis.json")
                for e in range(int(t.group_ebeg[g]), int(t.group_ebeg[g + 1])):
                    for tm in range(int(t.elem_tbeg[e]), int(t.elem_tbeg[e + 1])):
                        n_terms += 1
                        nf = int(t.term_fbeg[tm + 1] - t.term_fbeg[tm])
                        if nf != len(orb["sites"]):
                            bad.append(f"function {c}: a term with {nf} factors, the orbit's clusters have {len(orb['sites'])} sites")
            if orb["sites"] and (n_terms == 0 or n_terms % orb["mult"]):
                bad.append(f"function {c}: {n_terms} terms is not a multiple of the multiplicity {orb['mult']}")
            # the prototype cluster lies in the function's site neighborhood (m_orbit_site_neighborhood)
            if len(t.orbit_nbhd_beg) == t.corr_size + 1 and len(orb["sites"]) > 1:
                nb = {tuple(int(x) for x in r) for r in t.orbit_nbhd[int(t.orbit_nbhd_beg[c]):int(t.orbit_nbhd_beg[c + 1])]}
                o0 = orb["sites"][0]
                for site in orb["sites"]:
                    rel = (site[0], site[1] - o0[1], site[2] - o0[2], site[3] - o0[3])   # about the first site's cell
                    if rel not in nb:
                        bad.append(f"function {c}: prototype site {site} is not in the function's site neighborhood")
    if bad:
        raise ValueError("tables do not match basis.json:\n  " + "\n  ".join(bad))


def _main(argv=None) -> int:
    """python -m casmcode_clexmonte_b200.clexulator_tables <Clexulator.cc> <out.npz> [eci.json] [--flat out.cmxt]

    Export the flat tables of one CASM-generated Clexulator source (and print
    the work per single-site delta for the given coefficients).  --flat: also the
    single-file form cmx_tables_create_from_file reads (what a C++ plugin loads)."""
    import argparse
    ap = argparse.ArgumentParser(description=_main.__doc__)
    ap.add_argument("source")
    ap.add_argument("out")
    ap.add_argument("eci", nargs="?")
    ap.add_argument("--flat")
    a = ap.parse_args(argv)
    t = parse_clexulator_source(a.source)
    t.save(a.out)
    if a.flat:
        t.save_flat(a.flat)
    print(f"{a.out}: nlist {t.nlist_len}, corr {t.corr_size}, point corr {t.n_point_corr}, "
          f"sublattices on the neighbor list {t.n_nlist_sublat}")
    if a.eci:
        idx, val = read_eci(a.eci, t.corr_size)
        for p in range(t.n_nlist_sublat if t.n_point_corr == t.n_nlist_sublat else 0):
            print(f"  point {p}: {t.delta_work(p, idx)}")
    return 0


if __name__ == "__main__":
    raise SystemExit(_main())
