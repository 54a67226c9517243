"""Emit a SYNTHETIC FCC binary A-B pair + triplet Clexulator in the grammar of the CASM-generated
sources, so that BASELINE configs[0] ("FCC binary A-B ... pair+triplet
orbits") has a basis: the reference ships no such clexulator (SURVEY 8c "Gap": its FCC fixture is
A-B-Va with pairs only).

    python tests/golden/make_synthetic_clexulator.py [out.cc]     (default: oracle/_ref/...)

What is synthetic and what is not:
  * the FUNCTION BODIES, the site-function table, the neighborhood and the function tables are
    written in exactly the token shapes the reference's generated sources use (that is what the
    table exporter, casmcode_clexmonte_b200/clexulator_tables.py, parses) -- cluster orbits are
    enumerated here from the FCC geometry: point, 1NN pair, 2NN pair, 1NN equilateral triplet;
  * the class plumbing around them (parameter-pack filling, the _calc_* dispatch loops) is a
    compact stand-in written for this file: the oracle only needs the BaseClexulator entry
    points of oracle/shim to work.
The emitted source is compiled by `make -C oracle synthetic` into oracle/_ref/ (git-ignored, like
the reference's own clexulators) and is TEST INFRASTRUCTURE: golden vectors come from running it.

Site basis: one function, phi = (-1, +1) (A, B).  Correlation functions (corr index):
  0 constant; 1 point; 2 1NN pair / 6; 3 2NN pair / 3; 4 1NN triplet / 8
with the multiplicities per primitive cell as divisors, as CASM normalises.  Point functions sum
over every cluster that holds the site (12 / 6 / 24 clusters) with the same divisor, delta
functions are (phi[occ_f] - phi[occ_i]) * (sum over those clusters of the other sites) / mult.
"""
from __future__ import annotations

import itertools
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from casmcode_clexmonte_b200.clexulator_tables import ClexulatorTables, prim_neighbor_cells  # noqa: E402

NAME = "FCC_binary_Clexulator_synthetic"
LATTICE = np.array([[0, 2, 2], [2, 0, 2], [2, 2, 0]], dtype=np.int64)  # rows = lattice vectors (the FCC fixture's prim)


def dist2(c):
    r = np.asarray(c, dtype=np.int64) @ LATTICE
    return int(r @ r)


def geometry():
    """Neighbor-list cells (ordered by the prim neighbor list rule) and the clusters."""
    W = ClexulatorTables.load(ROOT / "tests/golden/tables/fcc_default.npz").weight_matrix  # the FCC prim's weight matrix
    rng = range(-2, 3)
    cells = [c for c in itertools.product(rng, rng, rng) if dist2(c) <= 16]
    assert len(cells) == 19
    order = [tuple(int(x) for x in c) for c in prim_neighbor_cells(W, cells)]
    assert sorted(order) == sorted(cells), "the neighbor list rule must enumerate exactly the neighborhood"
    index = {c: n for n, c in enumerate(order)}
    nn1 = [c for c in order if dist2(c) == 8]
    nn2 = [c for c in order if dist2(c) == 16]
    assert (len(nn1), len(nn2)) == (12, 6)
    fwd = lambda c: c > (0, 0, 0)  # noqa: E731  lexicographic half space
    tri_all = [(a, b) for a, b in itertools.combinations(nn1, 2) if dist2(np.subtract(a, b)) == 8]
    assert len(tri_all) == 24
    tri_fwd = [(a, b) for a, b in tri_all if fwd(a) and fwd(b)]
    assert len(tri_fwd) == 8
    return W, order, index, nn1, nn2, tri_all, tri_fwd, fwd


def of(n):
    return f"occ_func_0_0({n})"


def wrap_sum(terms, per_line=3):
    lines = [" + ".join(terms[i:i + per_line]) for i in range(0, len(terms), per_line)]
    return (" +\n          ").join(lines)


def emit() -> str:
    W, order, index, nn1, nn2, tri_all, tri_fwd, fwd = geometry()
    n_list = len(order)
    DPHI = "(m_occ_func_0_0[occ_f] - m_occ_func_0_0[occ_i])"
    bodies = {}
    # corr 0: constant
    bodies["eval_bfunc_0_0"] = "1"
    # corr 1: point
    bodies["eval_bfunc_1_0"] = of(0)
    bodies["site_eval_bfunc_1_0_at_0"] = of(0)
    bodies["site_deval_bfunc_1_0_at_0"] = DPHI
    # corr 2, 3: pairs
    for o, shell, mult in ((2, nn1, 6), (3, nn2, 3)):
        bodies[f"eval_bfunc_{o}_0"] = "(" + wrap_sum([f"{of(0)} * {of(index[c])}" for c in shell if fwd(c)]) + f") / {mult}."
        bodies[f"site_eval_bfunc_{o}_0_at_0"] = "(" + wrap_sum([f"{of(0)} * {of(index[c])}" for c in shell]) + f") / {mult}."
        bodies[f"site_deval_bfunc_{o}_0_at_0"] = DPHI + " * (" + wrap_sum([of(index[c]) for c in shell], 4) + f") / {mult}."
    # corr 4: 1NN triplet
    bodies["eval_bfunc_4_0"] = "(" + wrap_sum([f"{of(0)} * {of(index[a])} * {of(index[b])}" for a, b in tri_fwd], 2) + ") / 8."
    bodies["site_eval_bfunc_4_0_at_0"] = "(" + wrap_sum([f"{of(0)} * {of(index[a])} * {of(index[b])}" for a, b in tri_all], 2) + ") / 8."
    bodies["site_deval_bfunc_4_0_at_0"] = DPHI + " * (" + wrap_sum([f"{of(index[a])} * {of(index[b])}" for a, b in tri_all]) + ") / 8."

    def decl(name):
        args = "int occ_i, int occ_f" if "deval" in name else ""
        return f"  template <typename Scalar>\n  Scalar {name}({args}) const;\n"

    def defn(name, body):
        args = "int occ_i, int occ_f" if "deval" in name else ""
        return (f"template <typename Scalar>\nScalar {NAME}::{name}({args}) const {{\n"
                f"  return {body};\n}}\n")

    corr_size = 5
    orbit_cells = {0: [], 1: [(0, 0, 0)], 2: [(0, 0, 0)] + nn1, 3: [(0, 0, 0)] + nn2, 4: [(0, 0, 0)] + nn1}
    s = []
    s.append('#include <cstddef>\n\n#include "casm/clexulator/BaseClexulator.hh"\n'
             '#include "casm/clexulator/BasicClexParamPack.hh"\n#include "casm/global/eigen.hh"\n')
    s.append("/* SYNTHETIC basis set emitted by tests/golden/make_synthetic_clexulator.py (not a CASM project):\n"
             "   FCC binary A-B, site basis (-1, +1), point + 1NN pair + 2NN pair + 1NN triplet. */\n")
    s.append(f'extern "C" CASM::clexulator::BaseClexulator *make_{NAME}();\n')
    s.append("namespace CASM {\nnamespace clexulator {\n\ntypedef BasicClexParamPack ParamPack;\n")
    s.append(f"class {NAME} : public clexulator::BaseClexulator {{\n public:\n  {NAME}();\n")
    s.append("  ClexParamPack const &param_pack() const override { return m_params; }\n"
             "  ClexParamPack &param_pack() override { return m_params; }\n")
    for name in bodies:
        s.append(decl(name))
    s.append(f"""
 private:
  typedef double ({NAME}::*BasisFuncPtr)() const;
  typedef double ({NAME}::*DeltaBasisFuncPtr)(int, int) const;

  double m_occ_func_0_0[2];
  mutable ParamPack m_params;
  ParamPack::Key m_occ_site_func_param_key;
  BasisFuncPtr m_orbit_func_table_0[{corr_size}];
  BasisFuncPtr m_flower_func_table_0[1][{corr_size}];
  DeltaBasisFuncPtr m_delta_func_table_0[1][{corr_size}];
  mutable std::vector<double> m_corr_buffer;

  double eval_occ_func_0_0(const int &nlist_ind) const {{ return m_occ_func_0_0[_occ(nlist_ind)]; }}
  double occ_func_0_0(const int &nlist_ind) const {{
    return ParamPack::Val<double>::get(m_params, m_occ_site_func_param_key, 0, nlist_ind);
  }}
  template <typename Scalar>
  Scalar zero_func() const {{ return Scalar(0.0); }}
  template <typename Scalar>
  Scalar zero_func(int, int) const {{ return Scalar(0.0); }}

  // parameter pack: the site-function value of every neighbor-list site
  template <typename Scalar>
  void _prepare() const {{
""")
    for n in range(n_list):
        s.append(f"    ParamPack::Val<Scalar>::set(m_params, m_occ_site_func_param_key, 0, {n}, eval_occ_func_0_0({n}));\n")
    s.append(f"""  }}

  BaseClexulator *_clone() const override {{ return new {NAME}(*this); }}
  // (stand-in dispatch: bind was done by the caller, fill the pack, walk the function table)
  void _calc_global_corr_contribution() const override {{ _calc_global_corr_contribution(m_corr_buffer.data()); }}
  void _calc_global_corr_contribution(double *corr_begin) const override {{
    _prepare<double>();
    for (size_type i = 0; i < corr_size(); ++i) corr_begin[i] = (this->*m_orbit_func_table_0[i])();
  }}
  void _calc_restricted_global_corr_contribution(size_type const *b, size_type const *e) const override {{
    _calc_restricted_global_corr_contribution(m_corr_buffer.data(), b, e);
  }}
  void _calc_restricted_global_corr_contribution(double *corr_begin, size_type const *b, size_type const *e) const override {{
    _prepare<double>();
    for (; b < e; ++b) corr_begin[*b] = (this->*m_orbit_func_table_0[*b])();
  }}
  void _calc_point_corr(int nlist_ind) const override {{ _calc_point_corr(nlist_ind, m_corr_buffer.data()); }}
  void _calc_point_corr(int nlist_ind, double *corr_begin) const override {{
    _prepare<double>();
    for (size_type i = 0; i < corr_size(); ++i) corr_begin[i] = (this->*m_flower_func_table_0[nlist_ind][i])();
  }}
  void _calc_restricted_point_corr(int nlist_ind, size_type const *b, size_type const *e) const override {{
    _calc_restricted_point_corr(nlist_ind, m_corr_buffer.data(), b, e);
  }}
  void _calc_restricted_point_corr(int nlist_ind, double *corr_begin, size_type const *b, size_type const *e) const override {{
    _prepare<double>();
    for (; b < e; ++b) corr_begin[*b] = (this->*m_flower_func_table_0[nlist_ind][*b])();
  }}
  void _calc_delta_point_corr(int nlist_ind, int occ_i, int occ_f) const override {{
    _calc_delta_point_corr(nlist_ind, occ_i, occ_f, m_corr_buffer.data());
  }}
  void _calc_delta_point_corr(int nlist_ind, int occ_i, int occ_f, double *corr_begin) const override {{
    _prepare<double>();
    for (size_type i = 0; i < corr_size(); ++i) corr_begin[i] = (this->*m_delta_func_table_0[nlist_ind][i])(occ_i, occ_f);
  }}
  void _calc_restricted_delta_point_corr(int nlist_ind, int occ_i, int occ_f, size_type const *b, size_type const *e) const override {{
    _calc_restricted_delta_point_corr(nlist_ind, occ_i, occ_f, m_corr_buffer.data(), b, e);
  }}
  void _calc_restricted_delta_point_corr(int nlist_ind, int occ_i, int occ_f, double *corr_begin, size_type const *b,
                                         size_type const *e) const override {{
    _prepare<double>();
    for (; b < e; ++b) corr_begin[*b] = (this->*m_delta_func_table_0[nlist_ind][*b])(occ_i, occ_f);
  }}
}};

{NAME}::{NAME}() : BaseClexulator({n_list}, {corr_size}, 1) {{
  m_occ_func_0_0[0] = -1.000000000000, m_occ_func_0_0[1] = 1.000000000000;

  m_occ_site_func_param_key = m_params.allocate("occ_site_func", 1, {n_list}, true);
  m_corr_buffer.assign({corr_size}, 0.0);

""")
    for c in range(corr_size):
        s.append(f"  m_orbit_func_table_0[{c}] = &{NAME}::eval_bfunc_{c}_0<double>;\n")
    s.append("\n")
    for c in range(corr_size):
        fn = "zero_func" if c == 0 else f"site_eval_bfunc_{c}_0_at_0"
        s.append(f"  m_flower_func_table_0[0][{c}] = &{NAME}::{fn}<double>;\n")
    s.append("\n")
    for c in range(corr_size):
        fn = "zero_func" if c == 0 else f"site_deval_bfunc_{c}_0_at_0"
        s.append(f"  m_delta_func_table_0[0][{c}] = &{NAME}::{fn}<double>;\n")
    s.append("\n")
    for r in range(3):
        s.append(f"  m_weight_matrix.row({r}) << {int(W[r][0])}, {int(W[r][1])}, {int(W[r][2])};\n")
    s.append("\n  m_sublat_indices = std::set<int>{0};\n\n  m_n_sublattices = 1;\n\n")
    s.append("  m_neighborhood = std::set<xtal::UnitCell>{\n      " +
             ",\n      ".join(f"xtal::UnitCell({c[0]}, {c[1]}, {c[2]})" for c in sorted(order)) + "};\n\n")
    s.append(f"  m_orbit_neighborhood.resize(corr_size());\n  m_orbit_site_neighborhood.resize(corr_size());\n")
    for c in range(1, corr_size):
        cells = sorted(orbit_cells[c])
        s.append(f"  m_orbit_neighborhood[{c}] = std::set<xtal::UnitCell>{{\n      " +
                 ",\n      ".join(f"xtal::UnitCell({x[0]}, {x[1]}, {x[2]})" for x in cells) + "};\n")
        s.append(f"  m_orbit_site_neighborhood[{c}] = std::set<xtal::UnitCellCoord>{{\n      " +
                 ",\n      ".join(f"xtal::UnitCellCoord(0, {x[0]}, {x[1]}, {x[2]})" for x in cells) + "};\n\n")
    s.append("}\n\n")
    for name, body in bodies.items():
        s.append(defn(name, body))
        s.append("\n")
    s.append("}  // namespace clexulator\n}  // namespace CASM\n\n")
    s.append(f'extern "C" {{\nCASM::clexulator::BaseClexulator *make_{NAME}() {{\n'
             f"  return new CASM::clexulator::{NAME}();\n}}\n}}\n")
    return "".join(s)


if __name__ == "__main__":
    out = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "oracle/_ref" / f"{NAME}.cc"
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text(emit())
    print(out)
