// ORACLE / TEST INFRASTRUCTURE ONLY.
//
// This file is the CPU restatement of the reference's hot path *around* the
// reference's own CASM-generated Clexulator kernels (which are compiled
// unmodified into oracle/_ref/*.so by oracle/Makefile).  It is imported only
// by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl
// reference legs.  Nothing in the product (casmcode_clexmonte_b200/, include/)
// may link, import or execute it.
//
// Parity status: the *kernels* are the reference itself ("reference" kind).
// The call chain restated here lives in third-party libraries that are NOT
// under /root/reference (libcasm-clexulator v3.0a1, libcasm-monte v3.0a1,
// pyproject.toml:2-15); it is restated from their published behaviour and
// anchored on the reference's call sites.  Each function cites what it follows.
// "[EXT]" marks semantics that cannot be verified inside this container.
//
// Conventions:
//   linear site index   l = b * n_cells + cell          [EXT] Conversions
//   cell index          cell = i + N0 * (j + N1 * k)     (our choice; only
//                       affects site numbering, not physics)
//   occupation          int32, occupant index on the sublattice's allowed list
//   prim neighbor list  all unit cells with r^T W r <= R_max, sorted by
//                       (r^T W r, lexicographic (i,j,k)); neighbor index
//                       n = cell_rank * n_nlist_sublat + sublat_position
//                       (verified against the generated
//                       kernels by the identity  d(sum global) == delta corr).
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "casm/clexulator/BaseClexulator.hh"

using CASM::clexulator::BaseClexulator;
typedef BaseClexulator::size_type size_type;

namespace {

// CASM::KB, libcasm-global [EXT]; pinned by the documented event state in
// python/libcasm/clexmonte/_MonteCalculator.py:186-199.
const double KB = 8.6173303e-05;

struct Clex {
  void *dl = nullptr;
  BaseClexulator *clex = nullptr;
  std::vector<long> cells;   // 3 * n_nlist_cells, ordered prim neighbor list
  std::vector<int> sublats;  // sorted m_sublat_indices
  long rmax = 0;
};

long quad(Eigen::Matrix3l const &W, long i, long j, long k) {
  long r[3] = {i, j, k};
  long s = 0;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) s += r[a] * W(a, b) * r[b];
  return s;
}

// PrimNeighborList ordering [EXT], see header comment.  `rmax_override` < 0
// means "use the clexulator's own neighborhood"; a larger value reproduces the
// expansion of System::prim_neighbor_list shared by several basis sets
// (include/casm/clexmonte/system/System.hh:87-94).
void build_cells(Clex &c, long rmax_override) {
  Eigen::Matrix3l const &W = c.clex->weight_matrix();
  long rmax = 0, cmax = 0;
  for (auto const &u : c.clex->neighborhood()) {
    rmax = std::max(rmax, quad(W, u.c[0], u.c[1], u.c[2]));
    for (int a = 0; a < 3; ++a) cmax = std::max(cmax, std::labs(u.c[a]));
  }
  if (rmax_override > rmax) rmax = rmax_override;
  c.rmax = rmax;
  // find a box half-width B whose surface lies strictly outside the ellipsoid
  long B = cmax + 1;
  for (;; ++B) {
    bool outside = true;
    for (long i = -B; i <= B && outside; ++i)
      for (long j = -B; j <= B && outside; ++j)
        for (long k = -B; k <= B && outside; ++k) {
          if (std::max(std::labs(i), std::max(std::labs(j), std::labs(k))) != B)
            continue;
          if (quad(W, i, j, k) <= rmax) outside = false;
        }
    if (outside) break;
  }
  struct Item {
    long q, i, j, k;
  };
  std::vector<Item> items;
  for (long i = -B; i <= B; ++i)
    for (long j = -B; j <= B; ++j)
      for (long k = -B; k <= B; ++k) {
        long q = quad(W, i, j, k);
        if (q <= rmax) items.push_back({q, i, j, k});
      }
  std::sort(items.begin(), items.end(), [](Item const &a, Item const &b) {
    if (a.q != b.q) return a.q < b.q;
    if (a.i != b.i) return a.i < b.i;
    if (a.j != b.j) return a.j < b.j;
    return a.k < b.k;
  });
  c.cells.clear();
  for (auto const &it : items) {
    c.cells.push_back(it.i);
    c.cells.push_back(it.j);
    c.cells.push_back(it.k);
  }
  c.sublats.assign(c.clex->sublat_indices().begin(),
                   c.clex->sublat_indices().end());
}

// SuperNeighborList [EXT]: one row of linear site indices per unit cell.
struct Supercell {
  Clex *parent = nullptr;
  BaseClexulator *clex = nullptr;  // private clone (kernels keep mutable state)
  long N[3];
  long n_cells = 0;
  int n_sublat = 0;
  long row = 0;  // entries per row  (n_nlist_cells * n_nlist_sublat)
  std::vector<long> nlist;
  std::vector<int> sublat_to_pos;  // b -> position among nlist sublattices
  std::vector<double> tmp;
};

inline long wrap(long x, long n) {
  x %= n;
  return x < 0 ? x + n : x;
}

}  // namespace

extern "C" {

double orc_kb() { return KB; }

void *orc_open(const char *so_path, const char *factory, long rmax_override) {
  void *dl = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
  if (!dl) {
    std::fprintf(stderr, "orc_open: %s\n", dlerror());
    return nullptr;
  }
  typedef BaseClexulator *(*Factory)();
  Factory f = reinterpret_cast<Factory>(dlsym(dl, factory));
  if (!f) {
    std::fprintf(stderr, "orc_open: no symbol %s\n", factory);
    return nullptr;
  }
  Clex *c = new Clex;
  c->dl = dl;
  c->clex = f();
  build_cells(*c, rmax_override);
  return c;
}

void orc_close(void *h) {
  Clex *c = static_cast<Clex *>(h);
  delete c->clex;
  delete c;  // the .so stays mapped: other clones may still use its code
}

// out[0..7] = nlist_size, corr_size, n_point_corr, n_sublattices,
//             n_nlist_sublat, n_nlist_cells, rmax, |m_neighborhood|
void orc_info(void *h, long *out) {
  Clex *c = static_cast<Clex *>(h);
  out[0] = c->clex->nlist_size();
  out[1] = c->clex->corr_size();
  out[2] = c->clex->n_point_corr();
  out[3] = c->clex->n_sublattices();
  out[4] = static_cast<long>(c->sublats.size());
  out[5] = static_cast<long>(c->cells.size() / 3);
  out[6] = c->rmax;
  out[7] = static_cast<long>(c->clex->neighborhood().size());
}

void orc_nlist_cells(void *h, long *out) {
  Clex *c = static_cast<Clex *>(h);
  std::copy(c->cells.begin(), c->cells.end(), out);
}

void orc_sublat_indices(void *h, int *out) {
  Clex *c = static_cast<Clex *>(h);
  std::copy(c->sublats.begin(), c->sublats.end(), out);
}

void orc_weight_matrix(void *h, long *out) {
  Clex *c = static_cast<Clex *>(h);
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) out[3 * a + b] = c->clex->weight_matrix()(a, b);
}

// m_orbit_site_neighborhood[corr] as (b,i,j,k) rows; returns count
// (generated source, e.g. FCC_binary_vacancy_Clexulator_default.cc:394-439)
long orc_orbit_site_neighborhood(void *h, long corr, long *out, long cap) {
  Clex *c = static_cast<Clex *>(h);
  auto const &v = c->clex->orbit_site_neighborhood();
  if (corr < 0 || corr >= static_cast<long>(v.size())) return 0;
  long n = 0;
  for (auto const &u : v[corr]) {
    if (n < cap) {
      out[4 * n + 0] = u.b;
      out[4 * n + 1] = u.c[0];
      out[4 * n + 2] = u.c[1];
      out[4 * n + 3] = u.c[2];
    }
    ++n;
  }
  return n;
}

void *orc_supercell(void *h, long N0, long N1, long N2) {
  Clex *c = static_cast<Clex *>(h);
  Supercell *s = new Supercell;
  s->parent = c;
  s->clex = c->clex->clone();
  s->N[0] = N0;
  s->N[1] = N1;
  s->N[2] = N2;
  s->n_cells = N0 * N1 * N2;
  s->n_sublat = static_cast<int>(c->clex->n_sublattices());
  long ncell_nl = static_cast<long>(c->cells.size() / 3);
  long nsub_nl = static_cast<long>(c->sublats.size());
  s->row = ncell_nl * nsub_nl;
  s->sublat_to_pos.assign(s->n_sublat, -1);
  for (long p = 0; p < nsub_nl; ++p) s->sublat_to_pos[c->sublats[p]] = int(p);
  s->nlist.resize(static_cast<size_t>(s->n_cells) * s->row);
  for (long k = 0; k < N2; ++k)
    for (long j = 0; j < N1; ++j)
      for (long i = 0; i < N0; ++i) {
        long v = i + N0 * (j + N1 * k);
        long *dst = &s->nlist[static_cast<size_t>(v) * s->row];
        for (long r = 0; r < ncell_nl; ++r) {
          long ci = wrap(i + c->cells[3 * r + 0], N0);
          long cj = wrap(j + c->cells[3 * r + 1], N1);
          long ck = wrap(k + c->cells[3 * r + 2], N2);
          long cell = ci + N0 * (cj + N1 * ck);
          for (long p = 0; p < nsub_nl; ++p)
            dst[r * nsub_nl + p] = c->sublats[p] * s->n_cells + cell;
        }
      }
  s->tmp.assign(c->clex->corr_size(), 0.0);
  return s;
}

void orc_supercell_free(void *sh) {
  Supercell *s = static_cast<Supercell *>(sh);
  delete s->clex;
  delete s;
}

// [EXT] Correlations::occ_delta(l, new_occ), unrestricted:
//   v = l % n_cells, b = l / n_cells, p = position of b
//   -> generated _calc_delta_point_corr  (…default.cc:555-580)
void orc_delta_corr(void *sh, int const *occ, long l, int new_occ,
                    double *out) {
  Supercell *s = static_cast<Supercell *>(sh);
  long v = l % s->n_cells, b = l / s->n_cells;
  s->clex->calc_delta_point_corr(occ, &s->nlist[size_t(v) * s->row],
                                 s->sublat_to_pos[b], occ[l], new_occ, out);
}

// restricted form -> generated _calc_restricted_delta_point_corr (:582-610).
// Only the listed entries of `out` are written.
void orc_restricted_delta_corr(void *sh, int const *occ, long l, int new_occ,
                               unsigned const *idx, long nidx, double *out) {
  Supercell *s = static_cast<Supercell *>(sh);
  long v = l % s->n_cells, b = l / s->n_cells;
  s->clex->calc_restricted_delta_point_corr(
      occ, &s->nlist[size_t(v) * s->row], s->sublat_to_pos[b], occ[l], new_occ,
      out, idx, idx + nidx);
}

// generated _calc_point_corr (:500-553)
void orc_point_corr(void *sh, int const *occ, long l, double *out) {
  Supercell *s = static_cast<Supercell *>(sh);
  long v = l % s->n_cells, b = l / s->n_cells;
  s->clex->calc_point_corr(occ, &s->nlist[size_t(v) * s->row],
                           s->sublat_to_pos[b], out);
}

// one unit cell's contribution: generated _calc_global_corr_contribution
// (:446-470).  For local clexulators this is [EXT] LocalCorrelations::local.
void orc_cell_corr(void *sh, int const *occ, long cell, double *out) {
  Supercell *s = static_cast<Supercell *>(sh);
  s->clex->calc_global_corr_contribution(occ, &s->nlist[size_t(cell) * s->row],
                                         out);
}

// [EXT] Correlations::per_supercell(): sum over unit cells, ascending index.
void orc_global_corr(void *sh, int const *occ, double *out) {
  Supercell *s = static_cast<Supercell *>(sh);
  size_t nc = s->clex->corr_size();
  std::vector<double> tmp(nc);
  for (size_t i = 0; i < nc; ++i) out[i] = 0.0;
  for (long v = 0; v < s->n_cells; ++v) {
    s->clex->calc_global_corr_contribution(occ, &s->nlist[size_t(v) * s->row],
                                           tmp.data());
    for (size_t i = 0; i < nc; ++i) out[i] += tmp[i];
  }
}

// [EXT] Correlations::occ_delta({l_k},{new_k}) restricted to idx and
// ClusterExpansion::occ_delta_value = sum_i value_i * dcorr[index_i]
// (call sites: SemiGrandCanonicalCalculator.cc:198-199, CanonicalCalculator.cc:139,
//  BaseMonteEventData.cc:132-136).  Sites are applied one after another and
// restored at the end.  `dcorr` (corr_size, may be null) receives the summed
// delta correlations (only idx entries meaningful).
double orc_occ_delta_value(void *sh, int *occ, long nsites, long const *l,
                           int const *new_occ, unsigned const *idx,
                           double const *val, long nidx, double *dcorr) {
  Supercell *s = static_cast<Supercell *>(sh);
  size_t nc = s->clex->corr_size();
  std::vector<double> acc(nc, 0.0);
  std::vector<double> &tmp = s->tmp;
  int saved[8];
  for (long k = 0; k < nsites; ++k) {
    long v = l[k] % s->n_cells, b = l[k] / s->n_cells;
    s->clex->calc_restricted_delta_point_corr(
        occ, &s->nlist[size_t(v) * s->row], s->sublat_to_pos[b], occ[l[k]],
        new_occ[k], tmp.data(), idx, idx + nidx);
    if (k == 0)
      for (long i = 0; i < nidx; ++i) acc[idx[i]] = tmp[idx[i]];
    else
      for (long i = 0; i < nidx; ++i) acc[idx[i]] += tmp[idx[i]];
    saved[k] = occ[l[k]];
    occ[l[k]] = new_occ[k];
  }
  for (long k = 0; k < nsites; ++k) occ[l[k]] = saved[k];
  double e = 0.0;
  for (long i = 0; i < nidx; ++i) e += val[i] * acc[idx[i]];
  if (dcorr)
    for (size_t i = 0; i < nc; ++i) dcorr[i] = acc[i];
  return e;
}

// std::mt19937_64 + libstdc++ distributions, exposed so that the product's
// device-side restatement of the RNG stream can be checked draw by draw.
// [EXT] monte::RandomNumberGenerator: random_int(max) =
// uniform_int_distribution<long>(0,max), random_real(max) =
// uniform_real_distribution<double>(0,max).
void orc_rng_stream(uint64_t seed, long n, long const *int_max,
                    double const *real_max, int const *kind, long *out_int,
                    double *out_real, uint64_t *out_raw) {
  std::mt19937_64 eng(seed);
  for (long i = 0; i < n; ++i) {
    if (kind[i] == 0) {
      out_raw[i] = eng();
    } else if (kind[i] == 1) {
      out_int[i] = std::uniform_int_distribution<long>(0, int_max[i])(eng);
    } else {
      out_real[i] =
          std::uniform_real_distribution<double>(0.0, real_max[i])(eng);
    }
  }
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Sequential Metropolis loop  (methods/occupation_metropolis.hh:72-123)
// ---------------------------------------------------------------------------
namespace {

// [EXT] monte::OccCandidateList / OccLocation / OccSwap, restated from
// libcasm-monte v3.0a1.
struct Cand {
  int asym, species;
};
struct Swap {
  int a, b;  // candidate indices
};
struct Mol {
  long l;
  int asym, species;
  long loc;
};

struct Occupants {
  // inputs describing the prim
  int n_sublat, n_asym, n_species;
  std::vector<int> sublat_to_asym;          // [n_sublat]
  std::vector<std::vector<int>> occ_to_sp;  // [asym][occ] -> species index
  int occ_index(int asym, int species) const {
    for (size_t i = 0; i < occ_to_sp[asym].size(); ++i)
      if (occ_to_sp[asym][i] == species) return int(i);
    return -1;
  }
};

struct OccLoc {
  Occupants const *o;
  std::vector<Cand> cand;
  std::vector<std::vector<int>> cand_index;  // [asym][species] -> cand or -1
  std::vector<Mol> mol;
  std::vector<std::vector<long>> loc;  // [cand] -> mol ids

  void init(Occupants const &occs, long n_cells, int const *occ) {
    o = &occs;
    cand.clear();
    cand_index.assign(occs.n_asym, std::vector<int>(occs.n_species, -1));
    for (int a = 0; a < occs.n_asym; ++a) {
      if (occs.occ_to_sp[a].size() < 2) continue;
      for (size_t i = 0; i < occs.occ_to_sp[a].size(); ++i) {
        cand_index[a][occs.occ_to_sp[a][i]] = int(cand.size());
        cand.push_back({a, occs.occ_to_sp[a][i]});
      }
    }
    loc.assign(cand.size(), {});
    mol.clear();
    long n_sites = n_cells * occs.n_sublat;
    for (long l = 0; l < n_sites; ++l) {
      int a = occs.sublat_to_asym[l / n_cells];
      if (occs.occ_to_sp[a].size() < 2) continue;
      int sp = occs.occ_to_sp[a][occ[l]];
      int ci = cand_index[a][sp];
      Mol m{l, a, sp, long(loc[ci].size())};
      loc[ci].push_back(long(mol.size()));
      mol.push_back(m);
    }
  }
  // OccLocation::apply for one occ_transform
  void apply(long mol_id, int to_species, int *occ) {
    Mol &m = mol[mol_id];
    occ[m.l] = o->occ_index(m.asym, to_species);
    int ci = cand_index[m.asym][m.species];
    long back = loc[ci].back();
    loc[ci][m.loc] = back;
    mol[back].loc = m.loc;
    loc[ci].pop_back();
    m.species = to_species;
    ci = cand_index[m.asym][m.species];
    m.loc = long(loc[ci].size());
    loc[ci].push_back(mol_id);
  }
};

inline uint64_t fnv(uint64_t h, uint64_t x) {
  for (int i = 0; i < 8; ++i) {
    h ^= (x >> (8 * i)) & 0xffu;
    h *= 1099511628211ull;
  }
  return h;
}

Occupants make_occupants(int n_sublat, int const *sublat_to_asym,
                         int const *occ_to_species, int max_occ,
                         int n_species) {
  Occupants o;
  o.n_sublat = n_sublat;
  o.n_species = n_species;
  o.sublat_to_asym.assign(sublat_to_asym, sublat_to_asym + n_sublat);
  o.n_asym = 0;
  for (int b = 0; b < n_sublat; ++b)
    o.n_asym = std::max(o.n_asym, sublat_to_asym[b] + 1);
  o.occ_to_sp.assign(o.n_asym, {});
  for (int b = 0; b < n_sublat; ++b) {
    int a = sublat_to_asym[b];
    if (!o.occ_to_sp[a].empty()) continue;
    for (int i = 0; i < max_occ; ++i) {
      int sp = occ_to_species[b * max_occ + i];
      if (sp >= 0) o.occ_to_sp[a].push_back(sp);
    }
  }
  return o;
}

}  // namespace

extern "C" {

// Step log record written for the first `log_cap` steps.
struct OrcStep {
  long l0, l1;      // sites (l1 = -1 for semi-grand)
  int new0, new1;   // new occupant indices
  int accepted;
  int pad;
  double dE;        // delta potential energy (per supercell)
};

// Semi-grand canonical sequential loop.
//   propose : [EXT] propose_semigrand_canonical_event
//             (wrapper SemiGrandCanonicalCalculator.cc:104-115)
//   dE      : SemiGrandCanonicalPotential::occ_delta_per_supercell
//             (SemiGrandCanonicalCalculator.cc:186-213)
//   accept  : [EXT] metropolis_acceptance; loop body
//             methods/occupation_metropolis.hh:92-120, beta :86
//   apply   : [EXT] OccLocation::apply (:118-120)
// mode: 0 = semi-grand canonical, 1 = canonical
//   canonical: [EXT] propose_canonical_event (CanonicalCalculator.cc:78-88),
//   dE = two-site occ_delta_value (CanonicalCalculator.cc:137-140).
// exch_mu[n_species] = (R^T)^T mu, i.e. dE_pot = dE - sum_s exch_mu[s]*dN[s];
// the caller computes it the way :211-212 does (mu . (R^T dN)); here we keep
// the reference's evaluation order by passing R^T and mu separately.
long orc_metropolis_run(void *sh, int mode, int *occ, int n_sublat,
                        int const *sublat_to_asym, int const *occ_to_species,
                        int max_occ, int n_species, double const *param_chem_pot,
                        int n_param, double const *Rt /*[n_param][n_species]*/,
                        unsigned const *eci_idx, double const *eci_val,
                        long n_eci, double temperature, uint64_t seed,
                        long n_steps, OrcStep *log, long log_cap,
                        long *n_accept_out, uint64_t *hash_out,
                        double *seconds_out) {
  Supercell *s = static_cast<Supercell *>(sh);
  Occupants occs = make_occupants(n_sublat, sublat_to_asym, occ_to_species,
                                  max_occ, n_species);
  OccLoc ol;
  ol.init(occs, s->n_cells, occ);

  // swap tables (System.cc:55-58 -> [EXT] make_*_swaps)
  std::vector<Swap> swaps;
  int nc = int(ol.cand.size());
  if (mode == 0) {
    for (int a = 0; a < nc; ++a)
      for (int b = 0; b < nc; ++b)
        if (ol.cand[a].asym == ol.cand[b].asym &&
            ol.cand[a].species != ol.cand[b].species)
          swaps.push_back({a, b});
  } else {
    for (int a = 0; a < nc; ++a)
      for (int b = a + 1; b < nc; ++b)
        if (ol.cand[a].species != ol.cand[b].species &&
            ol.cand_index[ol.cand[a].asym][ol.cand[b].species] >= 0 &&
            ol.cand_index[ol.cand[b].asym][ol.cand[a].species] >= 0)
          swaps.push_back({a, b});
  }
  std::vector<double> tsum(swaps.size() + 1, 0.0);

  std::mt19937_64 eng(seed);
  double beta = 1.0 / (KB * temperature);
  std::vector<double> dN(n_species), x(n_param);
  long n_accept = 0;
  uint64_t h = 1469598103934665603ull;
  auto t0 = std::chrono::steady_clock::now();

  for (long step = 0; step < n_steps; ++step) {
    // ---- propose
    for (size_t i = 0; i < swaps.size(); ++i) {
      double w = double(ol.loc[swaps[i].a].size());
      if (mode == 1) w *= double(ol.loc[swaps[i].b].size());
      tsum[i + 1] = tsum[i] + w;
    }
    if (tsum.back() == 0.0) return -1;
    double r = std::uniform_real_distribution<double>(0.0, tsum.back())(eng);
    size_t si = 0;
    for (; si < swaps.size(); ++si)
      if (r < tsum[si + 1]) break;
    if (si == swaps.size()) return -2;
    Swap sw = swaps[si];
    long ls[2];
    int nw[2];
    long mols[2];
    int to_sp[2];
    long nsite = (mode == 0) ? 1 : 2;
    {
      long pos = std::uniform_int_distribution<long>(
          0, long(ol.loc[sw.a].size()) - 1)(eng);
      mols[0] = ol.loc[sw.a][pos];
      ls[0] = ol.mol[mols[0]].l;
      to_sp[0] = ol.cand[sw.b].species;
      nw[0] = occs.occ_index(ol.cand[sw.a].asym, to_sp[0]);
    }
    if (mode == 1) {
      long pos = std::uniform_int_distribution<long>(
          0, long(ol.loc[sw.b].size()) - 1)(eng);
      mols[1] = ol.loc[sw.b][pos];
      ls[1] = ol.mol[mols[1]].l;
      to_sp[1] = ol.cand[sw.a].species;
      nw[1] = occs.occ_index(ol.cand[sw.b].asym, to_sp[1]);
    }
    // ---- delta potential energy
    double dE = orc_occ_delta_value(s, occ, nsite, ls, nw, eci_idx, eci_val,
                                    n_eci, nullptr);
    if (mode == 0) {
      std::fill(dN.begin(), dN.end(), 0.0);
      for (long k = 0; k < nsite; ++k) {
        int asym = occs.sublat_to_asym[ls[k] / s->n_cells];
        dN[occs.occ_to_sp[asym][occ[ls[k]]]] += -1.0;
        dN[occs.occ_to_sp[asym][nw[k]]] += 1.0;
      }
      double dot = 0.0;
      for (int p = 0; p < n_param; ++p) {
        double xp = 0.0;
        for (int sp = 0; sp < n_species; ++sp)
          xp += Rt[p * n_species + sp] * dN[sp];
        dot += param_chem_pot[p] * xp;
      }
      dE = dE - dot;
    }
    // ---- accept / reject
    bool accept;
    if (dE < 0.0) {
      accept = true;
    } else {
      double u = std::uniform_real_distribution<double>(0.0, 1.0)(eng);
      accept = u < std::exp(-dE * beta);
    }
    if (step < log_cap) {
      log[step].l0 = ls[0];
      log[step].l1 = (mode == 1) ? ls[1] : -1;
      log[step].new0 = nw[0];
      log[step].new1 = (mode == 1) ? nw[1] : -1;
      log[step].accepted = accept ? 1 : 0;
      log[step].pad = 0;
      log[step].dE = dE;
    }
    h = fnv(h, uint64_t(ls[0]) * 2u + (accept ? 1u : 0u));
    if (accept) {
      ++n_accept;
      for (long k = 0; k < nsite; ++k) ol.apply(mols[k], to_sp[k], occ);
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  if (n_accept_out) *n_accept_out = n_accept;
  if (hash_out) *hash_out = h;
  if (seconds_out)
    *seconds_out = std::chrono::duration<double>(t1 - t0).count();
  return n_steps;
}

// SemiGrandCanonicalPotential::per_supercell
// (SemiGrandCanonicalCalculator.cc:171-179): E_clex - n_cells * mu . x,
// x = R^T (n - origin), n = mean number of each component per unit cell.
// CanonicalPotential::per_supercell (CanonicalCalculator.cc:126-133) is the
// n_param = 0 case.
double orc_potential_per_supercell(void *sh, int const *occ, int n_sublat,
                                   int const *sublat_to_asym,
                                   int const *occ_to_species, int max_occ,
                                   int n_species, double const *param_chem_pot,
                                   int n_param, double const *Rt,
                                   double const *origin,
                                   unsigned const *eci_idx,
                                   double const *eci_val, long n_eci,
                                   double *comp_n_out) {
  Supercell *s = static_cast<Supercell *>(sh);
  size_t nc = s->clex->corr_size();
  std::vector<double> corr(nc);
  orc_global_corr(s, occ, corr.data());
  double e = 0.0;
  for (long i = 0; i < n_eci; ++i) e += eci_val[i] * corr[eci_idx[i]];
  std::vector<double> n(n_species, 0.0);
  for (long l = 0; l < s->n_cells * n_sublat; ++l) {
    int b = int(l / s->n_cells);
    n[occ_to_species[b * max_occ + occ[l]]] += 1.0;
  }
  for (int sp = 0; sp < n_species; ++sp) n[sp] /= double(s->n_cells);
  if (comp_n_out)
    for (int sp = 0; sp < n_species; ++sp) comp_n_out[sp] = n[sp];
  double dot = 0.0;
  for (int p = 0; p < n_param; ++p) {
    double xp = 0.0;
    for (int sp = 0; sp < n_species; ++sp)
      xp += Rt[p * n_species + sp] * (n[sp] - origin[sp]);
    dot += param_chem_pot[p] * xp;
  }
  return e - double(s->n_cells) * dot;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// KMC event rates for the reference's own selector (oracle/kmc_lotto.cpp): the C++
// form of oracle.py::event_state, so that the CPU baseline of the KMC step has no
// Python in its loop.  EventStateCalculator::calculate_event_state +
// _default_event_state_calculation (BaseMonteEventData.cc:87-156), event_is_allowed
// (events/event_methods.cc:340-351); event id = unitcell * n_prim + prim_event
// (events/CompleteEventList.cc:76-91).
// ---------------------------------------------------------------------------
namespace {
struct KmcPrim {
  int n_sites;
  long site[4][4];  // (b, i, j, k)
  int occ_init[4], occ_final[4];
  Supercell *local;
  std::vector<unsigned> kra_idx, freq_idx;
  std::vector<double> kra_val, freq_val;
};
struct KmcCtx {
  Supercell *form;
  int *occ;
  std::vector<KmcPrim> prim;
  std::vector<unsigned> eci_idx;
  std::vector<double> eci_val;
  double beta;
  std::vector<double> corr;
};
inline void kmc_sites(KmcCtx const &c, KmcPrim const &p, long cell, long *l) {
  Supercell const *s = c.form;
  long i = cell % s->N[0], j = (cell / s->N[0]) % s->N[1], k = cell / (s->N[0] * s->N[1]);
  for (int q = 0; q < p.n_sites; ++q)
    l[q] = p.site[q][0] * s->n_cells + wrap(i + p.site[q][1], s->N[0]) +
           s->N[0] * (wrap(j + p.site[q][2], s->N[1]) + s->N[1] * wrap(k + p.site[q][3], s->N[2]));
}
}  // namespace

extern "C" {

// prim_sites[n_prim][4][4], prim_occ[n_prim][2][4] (init, final), local_sc[n_prim] supercell
// handles of each prim event's equivalent local clexulator, coefficient lists flattened
// with offsets kra_beg/freq_beg [n_prim + 1]
void *orc_kmc_ctx(void *form_sc, int *occ, long n_prim, int const *n_sites, long const *prim_sites,
                  int const *prim_occ, void *const *local_sc, long const *kra_beg, unsigned const *kra_idx,
                  double const *kra_val, long const *freq_beg, unsigned const *freq_idx, double const *freq_val,
                  unsigned const *eci_idx, double const *eci_val, long n_eci, double temperature) {
  KmcCtx *c = new KmcCtx;
  c->form = static_cast<Supercell *>(form_sc);
  c->occ = occ;
  c->eci_idx.assign(eci_idx, eci_idx + n_eci);
  c->eci_val.assign(eci_val, eci_val + n_eci);
  c->beta = 1.0 / (KB * temperature);
  for (long p = 0; p < n_prim; ++p) {
    KmcPrim e;
    e.n_sites = n_sites[p];
    for (int q = 0; q < 4; ++q) {
      for (int x = 0; x < 4; ++x) e.site[q][x] = prim_sites[(p * 4 + q) * 4 + x];
      e.occ_init[q] = prim_occ[(p * 2 + 0) * 4 + q];
      e.occ_final[q] = prim_occ[(p * 2 + 1) * 4 + q];
    }
    e.local = static_cast<Supercell *>(local_sc[p]);
    e.kra_idx.assign(kra_idx + kra_beg[p], kra_idx + kra_beg[p + 1]);
    e.kra_val.assign(kra_val + kra_beg[p], kra_val + kra_beg[p + 1]);
    e.freq_idx.assign(freq_idx + freq_beg[p], freq_idx + freq_beg[p + 1]);
    e.freq_val.assign(freq_val + freq_beg[p], freq_val + freq_beg[p + 1]);
    c->prim.push_back(e);
  }
  return c;
}

void orc_kmc_ctx_free(void *p) { delete static_cast<KmcCtx *>(p); }

double orc_kmc_rate(void *ctx, long event_id) {
  KmcCtx *c = static_cast<KmcCtx *>(ctx);
  long n_prim = (long)c->prim.size();
  long cell = event_id / n_prim;
  KmcPrim const &p = c->prim[event_id % n_prim];
  long l[4];
  kmc_sites(*c, p, cell, l);
  for (int q = 0; q < p.n_sites; ++q)
    if (c->occ[l[q]] != p.occ_init[q]) return 0.0;
  double dE = orc_occ_delta_value(c->form, c->occ, p.n_sites, l, p.occ_final, c->eci_idx.data(),
                                  c->eci_val.data(), (long)c->eci_idx.size(), nullptr);
  c->corr.resize(p.local->clex->corr_size());
  orc_cell_corr(p.local, c->occ, cell, c->corr.data());
  double ekra = 0.0, freq = 0.0;
  for (size_t i = 0; i < p.kra_idx.size(); ++i) ekra = ekra + p.kra_val[i] * c->corr[p.kra_idx[i]];
  for (size_t i = 0; i < p.freq_idx.size(); ++i) freq = freq + p.freq_val[i] * c->corr[p.freq_idx[i]];
  double dEa = dE * 0.5 + ekra;
  if (dEa < dE) dEa = dE;
  if (dEa < 0.0) dEa = 0.0;
  return freq * std::exp(-c->beta * dEa);
}

// OccLocation::apply for the complete event list: the sites take the final occupants
void orc_kmc_apply(void *ctx, long event_id) {
  KmcCtx *c = static_cast<KmcCtx *>(ctx);
  long n_prim = (long)c->prim.size();
  KmcPrim const &p = c->prim[event_id % n_prim];
  long l[4];
  kmc_sites(*c, p, event_id / n_prim, l);
  for (int q = 0; q < p.n_sites; ++q) c->occ[l[q]] = p.occ_final[q];
}
}
