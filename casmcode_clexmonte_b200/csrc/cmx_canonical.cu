// Canonical ensemble: parallel pair exchanges (see include/cmx_b200.h).
//
// Reference semantics reproduced per pair:
//   proposal   two sites with different species exchange them
//              (CanonicalCalculator.cc:78-88, propose_canonical_event [EXT])
//   delta E    ClusterExpansion::occ_delta_value({l_a, l_b}, {new_a, new_b}):
//              site a evaluated on the current configuration, site b with a
//              already changed (CanonicalCalculator.cc:137-140)
//   acceptance metropolis_acceptance [EXT]: dE < 0, else u < exp(-beta dE)
// What differs is WHICH pairs are visited and in what order: all pairs of one
// colour of one swap type at once (the pairs of a colour do not interact), then
// the next colour.  Every pair move satisfies detailed balance for the canonical
// distribution, so the composite chain keeps it invariant.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <set>

#include <cooperative_groups.h>

#include "cmx_internal.cuh"

namespace cg = cooperative_groups;

static int invalid(const std::string &msg) {
  cmx_set_error(msg);
  return CMX_ERR_INVALID;
}

struct SwapPlan {
  cmx_swap_type sw;
  int32_t pa, pb;       // point positions of the two sublattices
  int32_t S[3];         // colour strides
  int32_t n_colours;
  int32_t map_off_ab, map_off_ba;  // into the occupant maps
};

struct CanonicalPlan {
  std::vector<SwapPlan> swaps;
  int8_t *d_map = nullptr;  // per swap: map_ab[max_occ] (occ on b -> occ on a), map_ba[max_occ]
  long long *d_part = nullptr;  // [replica][blocks][2]: attempts, accepts
  double *d_part_dE = nullptr;  // [replica][blocks]
  int32_t blocks = 0;
  int coop_capacity = -1;  // co-resident blocks of k_canonical_pairs_warp (0: no cooperative launch)
  int lut_capacity = -1;   // co-resident blocks of k_canonical_pairs_lut
};

void cmx_canonical_free(cmx_state *s) {
  if (!s || !s->canon) return;
  cudaFree(s->canon->d_map);
  cudaFree(s->canon->d_part);
  cudaFree(s->canon->d_part_dE);
  delete s->canon;
  s->canon = nullptr;
}

struct CanonArgs {
  int8_t *occ;
  Geom g;
  DevTables T;
  int ba, bb, pa, pb;
  int t0, t1, t2;
  int S0, S1, S2, c0, c1, c2;
  FastDiv div0, div1;
  uint32_t items;
  const int32_t *gt_beg, *gt_fbeg, *gt_f, *gt_n;
  const double *gt_w;
  const double *beta;
  const int8_t *map_ab, *map_ba;
  long long *part;
  double *part_dE;
  uint32_t k0, k1, sweep_lo, ctr_hi;
};

// folded single-site delta E of point position p at cell (i,j,k), occupant
// oi -> of, with one overridden site (byte offset ov_off holds occupant ov_occ)
__device__ __forceinline__ double canon_site_delta(const CanonArgs &a, const int8_t *occ, int p,
                                                   int i, int j, int k, int oi, int of,
                                                   int64_t ov_off, int ov_occ) {
  const DevTables &T = a.T;
  const int mo = T.max_occ;
  double dE = 0.0;
  for (int t = a.gt_beg[p]; t < a.gt_beg[p + 1]; ++t) {
    double v = a.gt_w[((size_t)t * mo + oi) * mo + of];
    for (int q = a.gt_fbeg[t]; q < a.gt_fbeg[t + 1]; ++q) {
      const int n = a.gt_n[q];
      const int64_t no = cmx_nbr_offset(T, a.g, n, i, j, k, nullptr);
      const int o = (no == ov_off) ? ov_occ : cmx_dec(occ[no]);
      v *= T.phi[((size_t)T.nbr[n].w * T.n_func + a.gt_f[q]) * mo + o];
    }
    dE += v;
  }
  return dE;
}

__global__ void __launch_bounds__(128) k_canonical_pairs(CanonArgs a) {
  __shared__ long long sh_att[4], sh_acc[4];
  __shared__ double sh_sum[4];
  const int r = blockIdx.y;
  const Geom &g = a.g;
  int8_t *occ = a.occ + (size_t)r * g.rep_stride;
  const double beta = a.beta[r];
  long long n_att = 0, n_acc = 0;
  double e_sum = 0.0;
  for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < a.items;
       item += gridDim.x * blockDim.x) {
    uint32_t row, ii, kk, jj;
    fastdivmod(item, a.div0, row, ii);
    fastdivmod(row, a.div1, kk, jj);
    const int i = (int)ii * a.S0 + a.c0, j = (int)jj * a.S1 + a.c1, k = (int)kk * a.S2 + a.c2;
    int i2 = (i + a.t0) % g.N0, j2 = (j + a.t1) % g.N1, k2 = (k + a.t2) % g.N2;
    i2 += (i2 < 0) ? g.N0 : 0;
    j2 += (j2 < 0) ? g.N1 : 0;
    k2 += (k2 < 0) ? g.N2 : 0;
    const int64_t off_a = cmx_site_offset(g, a.ba, i, j, k);
    const int64_t off_b = cmx_site_offset(g, a.bb, i2, j2, k2);
    const int oa = cmx_dec(occ[off_a]), ob = cmx_dec(occ[off_b]);
    const int na = a.map_ab[ob], nb = a.map_ba[oa];  // species of b on a's sublattice, and v.v.
    if (na < 0 || nb < 0 || na == oa) continue;      // not allowed there / same species: no event
    ++n_att;
    double dE = canon_site_delta(a, occ, a.pa, i, j, k, oa, na, -1, 0);
    dE += canon_site_delta(a, occ, a.pb, i2, j2, k2, ob, nb, off_a, na);
    bool accept = dE < 0.0;
    if (!accept) {
      const uint32_t gid = (uint32_t)(((uint32_t)k * g.N1 + j) * g.N0 + i);
      const Philox ph = philox4x32_10(gid, (uint32_t)r, a.sweep_lo, a.ctr_hi, a.k0, a.k1);
      const unsigned long long u53 = ((unsigned long long)(ph.c[1] & 0x1FFFFFu) << 32) | ph.c[0];
      accept = (double)u53 * (1.0 / 9007199254740992.0) < exp(-dE * beta);
    }
    if (accept) {
      occ[off_a] = (int8_t)cmx_enc(g, na);
      occ[off_b] = (int8_t)cmx_enc(g, nb);
      ++n_acc;
      e_sum += dE;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_att += __shfl_down_sync(0xffffffffu, n_att, o);
    n_acc += __shfl_down_sync(0xffffffffu, n_acc, o);
    e_sum += __shfl_down_sync(0xffffffffu, e_sum, o);
  }
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    sh_att[wid] = n_att;
    sh_acc[wid] = n_acc;
    sh_sum[wid] = e_sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long A = 0, C = 0;
    double E = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      A += sh_att[w];
      C += sh_acc[w];
      E += sh_sum[w];
    }
    size_t slot = (size_t)r * gridDim.x + blockIdx.x;
    a.part[2 * slot] += A;
    a.part[2 * slot + 1] += C;
    a.part_dE[slot] += E;
  }
}

// Pair-LUT variant (the models the pair-LUT sweep covers: one sublattice, <= 3 occupants,
// point + pair functions of one neighbor shell): the single-site dE of the reference
// depends only on (occupant change, species counts over the shell) and is tabulated from
// the faithful evaluator (SweepPlan::d_pair_dE, k_build_pair_lut).  A swap is two
// lookups: site a on the current configuration, site b with a already changed (the
// reference's sequential occ_delta_value, CanonicalCalculator.cc:137-140) -- the shell of
// b is counted with a's new occupant when a belongs to it.  Same proposals and random
// bits as k_canonical_pairs; dE agrees with it to rounding (the reference sums the two
// sites' delta correlations before the ECI dot product, here the dot products are summed).
struct CanonLut {
  const double *pair_dE;  // [(oi * (nocc - 1) + alt) << 8 | n1 | n2 << 4]
  int nocc, z;
  int shell[48];
};

template <bool CG>
__device__ __forceinline__ double canon_lut_delta(const CanonArgs &a, const CanonLut &L, const int8_t *occ,
                                                  int i, int j, int k, int oi, int of, int64_t ov_off,
                                                  int ov_code) {
  const Geom &g = a.g;
  int n1 = 0, n2 = 0;
#pragma unroll 4
  for (int q = 0; q < L.z; ++q) {
    int ii = i + L.shell[3 * q], jj = j + L.shell[3 * q + 1], kk = k + L.shell[3 * q + 2];
    cmx_wrap_cell(g, ii, jj, kk);
    const int64_t no = cmx_site_offset(g, 0, ii, jj, kk);
    const int code = (no == ov_off) ? ov_code : (CG ? (int)__ldcg(occ + no) : (int)occ[no]);
    n1 += code & 1;   // storage codes 0 / 1 / 18
    n2 += code >> 4;
  }
  int alt = of - oi - 1;
  if (alt < 0) alt += L.nocc;
  return L.pair_dE[((oi * (L.nocc - 1) + alt) << 8) | n1 | (n2 << 4)];
}

// COOP: all colours [colour_begin, colour_end) of the swap type in one cooperative launch,
// a grid barrier between colours, occupations through L2 only
template <bool COOP>
__global__ void __launch_bounds__(128) k_canonical_pairs_lut(CanonArgs a, CanonLut L, int colour_begin,
                                                             int colour_end) {
  __shared__ long long sh_att[4], sh_acc[4];
  __shared__ double sh_sum[4];
  const int r = blockIdx.y;
  const Geom &g = a.g;
  int8_t *occ = a.occ + (size_t)r * g.rep_stride;
  const double beta = a.beta[r];
  long long n_att = 0, n_acc = 0;
  double e_sum = 0.0;
  for (int colour = colour_begin; colour < colour_end; ++colour) {
  const int c0 = colour % a.S0, c1 = (colour / a.S0) % a.S1, c2 = colour / (a.S0 * a.S1);
  const uint32_t ctr_hi = a.ctr_hi | (uint32_t)colour;
  for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < a.items;
       item += gridDim.x * blockDim.x) {
    uint32_t row, ii, kk, jj;
    fastdivmod(item, a.div0, row, ii);
    fastdivmod(row, a.div1, kk, jj);
    const int i = (int)ii * a.S0 + c0, j = (int)jj * a.S1 + c1, k = (int)kk * a.S2 + c2;
    int i2 = (i + a.t0) % g.N0, j2 = (j + a.t1) % g.N1, k2 = (k + a.t2) % g.N2;
    i2 += (i2 < 0) ? g.N0 : 0;
    j2 += (j2 < 0) ? g.N1 : 0;
    k2 += (k2 < 0) ? g.N2 : 0;
    const int64_t off_a = cmx_site_offset(g, 0, i, j, k);
    const int64_t off_b = cmx_site_offset(g, 0, i2, j2, k2);
    const int oa = cmx_dec(COOP ? (int)__ldcg(occ + off_a) : (int)occ[off_a]);
    const int ob = cmx_dec(COOP ? (int)__ldcg(occ + off_b) : (int)occ[off_b]);
    if (oa == ob) continue;  // same species: no event
    ++n_att;
    double dE = canon_lut_delta<COOP>(a, L, occ, i, j, k, oa, ob, -1, 0);
    dE += canon_lut_delta<COOP>(a, L, occ, i2, j2, k2, ob, oa, off_a, cmx_enc(g, ob));
    bool accept = dE < 0.0;
    if (!accept) {
      const uint32_t gid = (uint32_t)(((uint32_t)k * g.N1 + j) * g.N0 + i);
      const Philox ph = philox4x32_10(gid, (uint32_t)r, a.sweep_lo, ctr_hi, a.k0, a.k1);
      const unsigned long long u53 = ((unsigned long long)(ph.c[1] & 0x1FFFFFu) << 32) | ph.c[0];
      accept = (double)u53 * (1.0 / 9007199254740992.0) < exp(-dE * beta);
    }
    if (accept) {
      occ[off_a] = (int8_t)cmx_enc(g, ob);
      occ[off_b] = (int8_t)cmx_enc(g, oa);
      ++n_acc;
      e_sum += dE;
    }
  }
  if (COOP && colour + 1 < colour_end) cg::this_grid().sync();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_att += __shfl_down_sync(0xffffffffu, n_att, o);
    n_acc += __shfl_down_sync(0xffffffffu, n_acc, o);
    e_sum += __shfl_down_sync(0xffffffffu, e_sum, o);
  }
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    sh_att[wid] = n_att;
    sh_acc[wid] = n_acc;
    sh_sum[wid] = e_sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long A = 0, C = 0;
    double E = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      A += sh_att[w];
      C += sh_acc[w];
      E += sh_sum[w];
    }
    size_t slot = (size_t)r * gridDim.x + blockIdx.x;
    a.part[2 * slot] += A;
    a.part[2 * slot + 1] += C;
    a.part_dE[slot] += E;
  }
}

// Warp-cooperative variant: one pair per WARP (cmx_warp_site_delta), and ALL colours of a
// swap type in one cooperative launch with a grid barrier between colours -- a colour of
// a wide-orbit model holds a few hundred pairs (the stride must exceed the interaction
// range plus the translation), so one launch per colour is pure launch latency
// (ZrO 48^3: 13 800 launches per sweep).  Occupations are read through L2 only: other
// SMs wrote them before the barrier.  Same proposals, random bits and decision rule as
// k_canonical_pairs.
// SH: the term tables of the two point positions are staged in shared memory first
// (dynamic shared memory: [table a][table b, if another position][8 warps][stage_max + 1]).
template <bool SH>
__global__ void __launch_bounds__(256) k_canonical_pairs_warp(CanonArgs a, GenTerms G, int stage_max,
                                                              int colour_begin, int colour_end, int coop,
                                                              int table_bytes) {
  extern __shared__ __align__(16) unsigned char sh_dyn_c[];
  __shared__ long long sh_att[8], sh_acc[8];
  __shared__ double sh_sum[8];
  const int r = blockIdx.y;
  const Geom &g = a.g;
  int8_t *occ = a.occ + (size_t)r * g.rep_stride;
  const double beta = a.beta[r];
  const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  GenShared Sa = {}, Sb = {};
  int n_tab = 0;
  if (SH) {
    cmx_gen_stage_table(a.T, G, a.pa, sh_dyn_c, Sa);
    Sb = Sa;
    n_tab = 1;
    if (a.pb != a.pa) {
      cmx_gen_stage_table(a.T, G, a.pb, sh_dyn_c + table_bytes, Sb);
      n_tab = 2;
    }
  }
  double *sh_val = reinterpret_cast<double *>(sh_dyn_c + (size_t)n_tab * table_bytes) + (size_t)wib * (stage_max + 1);
  long long n_att = 0, n_acc = 0;
  double e_sum = 0.0;
  for (int colour = colour_begin; colour < colour_end; ++colour) {
    const int c0 = colour % a.S0, c1 = (colour / a.S0) % a.S1, c2 = colour / (a.S0 * a.S1);
    const uint32_t ctr_hi = a.ctr_hi | (uint32_t)colour;
    for (uint32_t item = blockIdx.x * 8u + wib; item < a.items; item += gridDim.x * 8u) {
      uint32_t row, ii, kk, jj;
      fastdivmod(item, a.div0, row, ii);
      fastdivmod(row, a.div1, kk, jj);
      const int i = (int)ii * a.S0 + c0, j = (int)jj * a.S1 + c1, k = (int)kk * a.S2 + c2;
      int i2 = (i + a.t0) % g.N0, j2 = (j + a.t1) % g.N1, k2 = (k + a.t2) % g.N2;
      i2 += (i2 < 0) ? g.N0 : 0;
      j2 += (j2 < 0) ? g.N1 : 0;
      k2 += (k2 < 0) ? g.N2 : 0;
      const int64_t off_a = cmx_site_offset(g, a.ba, i, j, k);
      const int64_t off_b = cmx_site_offset(g, a.bb, i2, j2, k2);
      const int oa = cmx_dec((int)__ldcg(occ + off_a)), ob = cmx_dec((int)__ldcg(occ + off_b));
      const int na = a.map_ab[ob], nb = a.map_ba[oa];
      if (na < 0 || nb < 0 || na == oa) continue;  // warp-uniform
      if (lane == 0) ++n_att;
      double dE;
      if (SH) {
        if (lane == 0) sh_val[Sa.n_act * a.T.n_func] = 1.0;  // the unused-factor slot
        dE = cmx_warp_site_delta_sh<true>(a.T, g, Sa, occ, sh_val, i, j, k, oa, na, -1, 0, lane);
        if (lane == 0) sh_val[Sb.n_act * a.T.n_func] = 1.0;
        dE += cmx_warp_site_delta_sh<true>(a.T, g, Sb, occ, sh_val, i2, j2, k2, ob, nb, off_a, na, lane);
      } else {
        dE = cmx_warp_site_delta<true>(a.T, g, G, occ, sh_val, a.pa, i, j, k, oa, na, -1, 0, lane);
        dE += cmx_warp_site_delta<true>(a.T, g, G, occ, sh_val, a.pb, i2, j2, k2, ob, nb, off_a, na, lane);
      }
      bool accept = dE < 0.0;
      if (!accept) {
        const uint32_t gid = (uint32_t)(((uint32_t)k * g.N1 + j) * g.N0 + i);
        const Philox ph = philox4x32_10(gid, (uint32_t)r, a.sweep_lo, ctr_hi, a.k0, a.k1);
        const unsigned long long u53 = ((unsigned long long)(ph.c[1] & 0x1FFFFFu) << 32) | ph.c[0];
        accept = (double)u53 * (1.0 / 9007199254740992.0) < exp(-dE * beta);
      }
      if (accept && lane == 0) {
        occ[off_a] = (int8_t)cmx_enc(g, na);
        occ[off_b] = (int8_t)cmx_enc(g, nb);
        ++n_acc;
        e_sum += dE;
      }
    }
    if (coop && colour + 1 < colour_end) cg::this_grid().sync();
  }
  if (lane == 0) {
    sh_att[wib] = n_att;
    sh_acc[wib] = n_acc;
    sh_sum[wib] = e_sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long A = 0, C = 0;
    double E = 0.0;
    for (int w = 0; w < 8; ++w) {
      A += sh_att[w];
      C += sh_acc[w];
      E += sh_sum[w];
    }
    size_t slot = (size_t)r * gridDim.x + blockIdx.x;
    a.part[2 * slot] += A;
    a.part[2 * slot + 1] += C;
    a.part_dE[slot] += E;
  }
}

__global__ void k_canonical_reduce(const long long *__restrict__ part,
                                   const double *__restrict__ part_dE, int nb,
                                   cmx_counters *out) {
  int r = blockIdx.x;
  if (threadIdx.x) return;
  long long A = 0, C = 0;
  double E = 0.0;
  for (int q = 0; q < nb; ++q) {
    A += part[2 * ((size_t)r * nb + q)];
    C += part[2 * ((size_t)r * nb + q) + 1];
    E += part_dE[(size_t)r * nb + q];
  }
  out[r].n_attempt = A;
  out[r].n_accept = C;
  out[r].dE_sum = E;
  out[r].reserved = 0;
}

static std::vector<int> divisors_of(int N) {
  std::vector<int> d;
  for (int s = 1; s <= N; ++s)
    if (N % s == 0) d.push_back(s);
  return d;
}

// Is the colouring with strides S valid for translation t?  Two pairs of one
// colour differ by a nonzero v in S*Z^3 (mod N).  They conflict when a site of
// one is identical to, or within the active neighborhood of, a site of the
// other:  v + e in Nbr u {0}  for some e in {0, +t, -t}.
static bool colouring_ok(const int N[3], const int S[3], const int t[3],
                         const std::set<std::array<int, 3>> &nbr) {
  for (auto const &o : nbr)
    for (int es = -1; es <= 1; ++es) {
      int v[3];
      bool lattice = true, zero = true;
      for (int ax = 0; ax < 3; ++ax) {
        int x = ((o[ax] - es * t[ax]) % N[ax] + N[ax]) % N[ax];
        v[ax] = x;
        if (x % S[ax]) lattice = false;
        if (x) zero = false;
      }
      (void)v;
      if (lattice && !zero) return false;
    }
  return true;
}

// The default swap table: the parallel counterpart of the reference's canonical swaps
// (make_canonical_swaps [EXT], built at src/casm/clexmonte/system/System.cc:55-58: every pair
// of candidates that share a species on one asymmetric unit).  For every ordered pair of
// mutable sublattices on the same asymmetric unit with a common species: the n_shell shortest
// translations of the prim neighbor list (one of +t / -t when the sublattices are equal: both
// give the same pairs) and, with long_range, one long translation that moves species across
// the whole box (t_i = 2 mod 4: a stride-4 colouring along i is conflict free for any
// short-ranged basis).  The host-language mirror is potential.canonical_swap_types.
extern "C" int cmx_canonical_default_swaps(const cmx_state *s, int32_t n_shell, int32_t long_range, int32_t cap,
                                           cmx_swap_type *swaps, int32_t *n) {
  if (!s || !n || n_shell <= 0 || cap < 0 || (cap && !swaps)) return invalid("cmx_canonical_default_swaps: bad argument");
  if (s->sublat_to_asym.empty()) {
    cmx_set_error("cmx_canonical_default_swaps: the occupants are not set (cmx_state_set_occupants)");
    return CMX_ERR_STATE;
  }
  const cmx_tables *t = s->t;
  const int nb = t->d.n_sublat, mo = t->d.max_occ;
  std::vector<cmx_swap_type> out;
  auto shares_species = [&](int ba, int bb) {
    for (int oa = 0; oa < t->n_occ[ba]; ++oa)
      for (int ob = 0; ob < t->n_occ[bb]; ++ob)
        if (s->occ_to_species[(size_t)ba * mo + oa] >= 0 &&
            s->occ_to_species[(size_t)ba * mo + oa] == s->occ_to_species[(size_t)bb * mo + ob])
          return true;
    return false;
  };
  for (int ba = 0; ba < nb; ++ba) {
    if (t->n_occ[ba] <= 1) continue;
    for (int bb = ba; bb < nb; ++bb) {
      if (t->n_occ[bb] <= 1 || s->sublat_to_asym[ba] != s->sublat_to_asym[bb] || !shares_species(ba, bb)) continue;
      const int want = (ba == bb) ? n_shell / 2 : n_shell;
      const size_t first = out.size();
      for (int q = 0; q < t->d.nlist_len && (int)(out.size() - first) < want; ++q) {
        const int32_t *o = &t->nbr[4 * q];
        if (o[3] != bb || (ba == bb && o[0] == 0 && o[1] == 0 && o[2] == 0)) continue;
        bool mirrored = false;
        for (size_t k = first; k < out.size() && ba == bb; ++k)
          mirrored = mirrored || (out[k].t[0] == -o[0] && out[k].t[1] == -o[1] && out[k].t[2] == -o[2]);
        if (mirrored) continue;
        cmx_swap_type w;
        w.b_a = ba;
        w.b_b = bb;
        w.t[0] = o[0];
        w.t[1] = o[1];
        w.t[2] = o[2];
        out.push_back(w);
      }
      if (long_range && s->g.N0 >= 8 && s->g.N0 % 4 == 0) {
        const int half = s->g.N0 / 2;
        cmx_swap_type w;
        w.b_a = ba;
        w.b_b = bb;
        w.t[0] = (half - (half % 4) + 2) % s->g.N0;
        w.t[1] = s->g.N1 / 3;
        w.t[2] = (s->g.N2 > 2) ? s->g.N2 / 2 + 1 : 0;
        out.push_back(w);
      }
    }
  }
  *n = (int32_t)out.size();
  if ((int32_t)out.size() > cap) {
    if (cap == 0) return CMX_OK;  // size query
    return invalid("cmx_canonical_default_swaps: more swap types than the caller's buffer holds");
  }
  for (size_t q = 0; q < out.size(); ++q) swaps[q] = out[q];
  return CMX_OK;
}

extern "C" int cmx_canonical_set_swaps(cmx_state *s, int32_t n, const cmx_swap_type *swaps) {
  if (!s || n <= 0 || !swaps) return invalid("cmx_canonical_set_swaps: bad argument");
  if (!s->plan.valid) {
    cmx_set_error("cmx_canonical_set_swaps: bind the ECI of a global clexulator first");
    return CMX_ERR_STATE;
  }
  if (s->g.halo) return invalid("cmx_canonical_set_swaps: slab states are not supported");
  if (s->g.s10 | s->g.s20 | s->g.s21) {
    cmx_set_error("cmx_canonical_set_swaps: pair-exchange colourings need a diag(N0, N1, N2) supercell "
                  "(general supercells: semi-grand sweeps, the reference-order mode and KMC)");
    return CMX_ERR_UNSUPPORTED;
  }
  const cmx_tables *t = s->t;
  const DevTables &T = t->d;
  const int mo = T.max_occ;
  const int N[3] = {s->g.N0, s->g.N1, s->g.N2};
  // active neighborhood (both directions) + the site itself
  std::set<std::array<int, 3>> nbr;
  nbr.insert(std::array<int, 3>{0, 0, 0});
  for (int nn : s->plan.active_nbr) {
    const int32_t *o = &t->nbr[4 * nn];
    nbr.insert(std::array<int, 3>{o[0], o[1], o[2]});
    nbr.insert(std::array<int, 3>{-o[0], -o[1], -o[2]});
  }
  CMX_CUDA(cudaSetDevice(t->device));
  cmx_canonical_free(s);
  CanonicalPlan *P = new CanonicalPlan;
  std::vector<int8_t> maps;
  for (int q = 0; q < n; ++q) {
    SwapPlan sp;
    sp.sw = swaps[q];
    const int ba = sp.sw.b_a, bb = sp.sw.b_b;
    auto fail = [&](const std::string &m) {
      delete P;
      return invalid("cmx_canonical_set_swaps: swap " + std::to_string(q) + ": " + m);
    };
    if (ba < 0 || ba >= T.n_sublat || bb < 0 || bb >= T.n_sublat) return fail("sublattice out of range");
    sp.pa = sp.pb = -1;
    for (int p = 0; p < T.n_nlist_sublat; ++p) {
      if (t->nlist_sublat[p] == ba) sp.pa = p;
      if (t->nlist_sublat[p] == bb) sp.pb = p;
    }
    if (sp.pa < 0 || sp.pb < 0) return fail("sublattice has no point functions");
    if (t->n_occ[ba] < 2 || t->n_occ[bb] < 2) return fail("sublattice is not mutable");
    int tt[3];
    bool tzero = true;
    for (int ax = 0; ax < 3; ++ax) {
      tt[ax] = ((sp.sw.t[ax] % N[ax]) + N[ax]) % N[ax];
      if (tt[ax]) tzero = false;
    }
    if (tzero && ba == bb) return fail("a site cannot be swapped with itself");
    // occupant maps through the species
    sp.map_off_ab = (int32_t)maps.size();
    maps.resize(maps.size() + 2 * mo, (int8_t)-1);
    sp.map_off_ba = sp.map_off_ab + mo;
    if (ba == bb) {
      for (int o = 0; o < t->n_occ[ba]; ++o) maps[sp.map_off_ab + o] = maps[sp.map_off_ba + o] = (int8_t)o;
    } else {
      if (s->occ_to_species.empty()) {
        delete P;
        cmx_set_error("cmx_canonical_set_swaps: cross-sublattice swaps need cmx_state_set_occupants");
        return CMX_ERR_STATE;
      }
      for (int ob = 0; ob < t->n_occ[bb]; ++ob)
        for (int oa = 0; oa < t->n_occ[ba]; ++oa)
          if (s->occ_to_species[bb * mo + ob] == s->occ_to_species[ba * mo + oa]) {
            maps[sp.map_off_ab + ob] = (int8_t)oa;
            maps[sp.map_off_ba + oa] = (int8_t)ob;
          }
    }
    // smallest valid colouring
    long best = -1;
    for (int s0 : divisors_of(N[0]))
      for (int s1 : divisors_of(N[1]))
        for (int s2 : divisors_of(N[2])) {
          long nc = (long)s0 * s1 * s2;
          if (nc > 4096 || (best >= 0 && nc >= best)) continue;
          const int S[3] = {s0, s1, s2};
          if (!colouring_ok(N, S, tt, nbr)) continue;
          best = nc;
          sp.S[0] = s0;
          sp.S[1] = s1;
          sp.S[2] = s2;
        }
    if (best < 0) return fail("no conflict-free colouring of this supercell (box too small for the translation)");
    sp.n_colours = (int32_t)best;
    P->swaps.push_back(sp);
  }
  cudaError_t e = cudaMalloc((void **)&P->d_map, maps.size());
  if (e == cudaSuccess) e = cudaMemcpy(P->d_map, maps.data(), maps.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(P->d_map);
    delete P;
    return cmx_cuda_fail(e, "cmx_canonical_set_swaps");
  }
  s->canon = P;
  return CMX_OK;
}

extern "C" int cmx_canonical_info(const cmx_state *s, int32_t i, int32_t *strides,
                                  int32_t *n_colours) {
  if (!s || !s->canon || i < 0 || i >= (int32_t)s->canon->swaps.size())
    return invalid("cmx_canonical_info: no such swap type");
  const SwapPlan &sp = s->canon->swaps[i];
  if (strides)
    for (int ax = 0; ax < 3; ++ax) strides[ax] = sp.S[ax];
  if (n_colours) *n_colours = sp.n_colours;
  return CMX_OK;
}

// sweeps enqueued on the state's stream; `reset` zeroes the per-block counters first
int cmx_canonical_enqueue(cmx_state *s, int64_t n_sweeps, uint64_t seed, int64_t first_sweep, bool reset) {
  if (!s) return invalid("cmx_canonical_sweep: null state");
  if (!s->canon) {
    cmx_set_error("cmx_canonical_sweep: no swap types (cmx_canonical_set_swaps)");
    return CMX_ERR_STATE;
  }
  if (n_sweeps < 0) return invalid("cmx_canonical_sweep: n_sweeps < 0");
  for (int r = 0; r < s->n_replicas; ++r)
    if (!(s->temperature[r] > 0.0)) {
      cmx_set_error("cmx_canonical_sweep: conditions not set for every replica");
      return CMX_ERR_STATE;
    }
  CMX_CUDA(cudaSetDevice(s->t->device));
  CanonicalPlan &P = *s->canon;
  SweepPlan &SP = s->plan;
  const Geom &g = s->g;
  // grid: sized for the coarsest colouring (most items per launch)
  uint32_t max_items = 1;
  for (auto const &sp : P.swaps)
    max_items = std::max<uint32_t>(max_items, (uint32_t)((g.n_cells) / sp.n_colours));
  const bool warp = cmx_use_warp_generic(s);
  const size_t table_bytes = (cmx_gen_shared_bytes(SP.pk_terms_max, SP.pk_act_max, s->t->d.max_occ) + 15) & ~(size_t)15;
  const bool staged = warp && SP.d_gt_pk && 2 * table_bytes + (size_t)(SP.stage_max + 1) * 64 <= 200 * 1024;
  const size_t stage_bytes = (size_t)(SP.stage_max + 1) * 8 * sizeof(double) + (staged ? 2 * table_bytes : 0);
  const void *kern = staged ? (const void *)k_canonical_pairs_warp<true> : (const void *)k_canonical_pairs_warp<false>;
  int blocks = (int)std::min<uint32_t>((max_items + 127) / 128,
                                       std::max(1, (148 * 16 + s->n_replicas - 1) / s->n_replicas));
  bool coop = false;
  const bool lut_path = !warp && SP.pair_lut && SP.z <= 16 && s->g.coded == (SP.nocc == 3) &&
                        !(s->sweep_flags & CMX_SWEEP_FORCE_GENERIC);
  if (lut_path) {
    if (P.lut_capacity < 0) {
      int per_sm = 0, dev = 0, sms = 0, can = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_canonical_pairs_lut<true>, 128, 0) != cudaSuccess)
        per_sm = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, dev);
      P.lut_capacity = can ? per_sm * sms : 0;
    }
    // a co-resident grid, so that all colours of a swap type run in one cooperative launch
    if (P.lut_capacity / std::max(1, s->n_replicas) >= 1)
      blocks = std::max(1, std::min(blocks, P.lut_capacity / s->n_replicas));
  }
  if (warp) {
    // one pair per warp, 8 warps per block; all blocks co-resident for the grid barrier
    if (P.coop_capacity < 0) {
      int per_sm = 0, dev = 0, sms = 0, can = 0;
      CMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, stage_bytes) != cudaSuccess)
        per_sm = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, dev);
      P.coop_capacity = can ? per_sm * sms : 0;
    }
    const int want = (int)((max_items + 7) / 8);
    const int cap = P.coop_capacity / std::max(1, s->n_replicas);
    static const bool no_coop = getenv("CMX_CANONICAL_NO_COOP") != nullptr;
    coop = cap >= 1 && !no_coop;
    blocks = std::max(1, std::min(want, coop ? cap : std::max(1, 148 * 8 / s->n_replicas)));
  }
  if (blocks != P.blocks || !P.d_part) {
    cudaFree(P.d_part);
    cudaFree(P.d_part_dE);
    P.d_part = nullptr;
    P.d_part_dE = nullptr;
    size_t nslots = (size_t)blocks * s->n_replicas;
    CMX_CUDA(cudaMalloc((void **)&P.d_part, sizeof(long long) * 2 * nslots));
    CMX_CUDA(cudaMalloc((void **)&P.d_part_dE, sizeof(double) * nslots));
    P.blocks = blocks;
    reset = true;
  }
  if (reset) {
    size_t nslots = (size_t)P.blocks * s->n_replicas;
    CMX_CUDA(cudaMemsetAsync(P.d_part, 0, sizeof(long long) * 2 * nslots, s->stream));
    CMX_CUDA(cudaMemsetAsync(P.d_part_dE, 0, sizeof(double) * nslots, s->stream));
  }
  CanonArgs a;
  a.occ = s->d_occ;
  a.g = g;
  a.T = s->t->d;
  a.gt_beg = SP.d_gt_beg;
  a.gt_fbeg = SP.d_gt_fbeg;
  a.gt_f = SP.d_gt_f;
  a.gt_n = SP.d_gt_n;
  a.gt_w = SP.d_gt_w;
  a.beta = s->d_beta;
  a.part = P.d_part;
  a.part_dE = P.d_part_dE;
  a.k0 = (uint32_t)seed;
  a.k1 = (uint32_t)(seed >> 32) ^ 0x43414E4Fu;  // "CANO": a stream of its own
  dim3 grid(P.blocks, s->n_replicas);
  for (int64_t w = 0; w < n_sweeps; ++w) {
    const int64_t sweep = first_sweep + w;
    a.sweep_lo = (uint32_t)sweep;
    for (size_t q = 0; q < P.swaps.size(); ++q) {
      const SwapPlan &sp = P.swaps[q];
      a.ba = sp.sw.b_a;
      a.bb = sp.sw.b_b;
      a.pa = sp.pa;
      a.pb = sp.pb;
      a.t0 = sp.sw.t[0];
      a.t1 = sp.sw.t[1];
      a.t2 = sp.sw.t[2];
      a.S0 = sp.S[0];
      a.S1 = sp.S[1];
      a.S2 = sp.S[2];
      uint32_t n0 = g.N0 / sp.S[0], n1 = g.N1 / sp.S[1], n2 = g.N2 / sp.S[2];
      a.div0 = make_fastdiv(n0);
      a.div1 = make_fastdiv(n1);
      a.items = n0 * n1 * n2;
      a.map_ab = P.d_map + sp.map_off_ab;
      a.map_ba = P.d_map + sp.map_off_ba;
      if (warp) {
        GenTerms G{SP.d_gt_beg, SP.d_gt_fbeg, SP.d_gt_vi, SP.d_act_beg, SP.d_act_n, SP.d_gt_w, SP.d_gt_pk};
        int stage_max = SP.stage_max, tb = (int)table_bytes;
        a.ctr_hi = ((uint32_t)((uint64_t)sweep >> 32) << 24) | ((uint32_t)q << 12);
        if (coop) {
          int c_begin = 0, c_end = sp.n_colours, one = 1;
          void *args[7] = {&a, &G, &stage_max, &c_begin, &c_end, &one, &tb};
          CMX_CUDA(cudaLaunchCooperativeKernel(kern, grid, dim3(256), args, stage_bytes, s->stream));
        } else {
          for (int c = 0; c < sp.n_colours; ++c) {
            if (staged) k_canonical_pairs_warp<true><<<grid, 256, stage_bytes, s->stream>>>(a, G, stage_max, c, c + 1, 0, tb);
            else k_canonical_pairs_warp<false><<<grid, 256, stage_bytes, s->stream>>>(a, G, stage_max, c, c + 1, 0, tb);
          }
        }
        continue;
      }
      const bool lut = lut_path;
      CanonLut CL;
      if (lut) {
        CL.pair_dE = SP.d_pair_dE;
        CL.nocc = SP.nocc;
        CL.z = SP.z;
        for (int x = 0; x < 48; ++x) CL.shell[x] = SP.shell[x];
      }
      if (lut) {
        a.ctr_hi = ((uint32_t)((uint64_t)sweep >> 32) << 24) | ((uint32_t)q << 12);
        static const bool no_coop = getenv("CMX_CANONICAL_NO_COOP") != nullptr;
        if (!no_coop && (long long)grid.x * grid.y <= P.lut_capacity) {
          int c_begin = 0, c_end = sp.n_colours;
          void *args[4] = {&a, &CL, &c_begin, &c_end};
          CMX_CUDA(cudaLaunchCooperativeKernel((const void *)k_canonical_pairs_lut<true>, grid, dim3(128), args, 0,
                                               s->stream));
        } else {
          for (int c = 0; c < sp.n_colours; ++c) k_canonical_pairs_lut<false><<<grid, 128, 0, s->stream>>>(a, CL, c, c + 1);
        }
        continue;
      }
      uint32_t colour = 0;
      for (int c2 = 0; c2 < sp.S[2]; ++c2)
        for (int c1 = 0; c1 < sp.S[1]; ++c1)
          for (int c0 = 0; c0 < sp.S[0]; ++c0) {
            a.c0 = c0;
            a.c1 = c1;
            a.c2 = c2;
            a.ctr_hi = ((uint32_t)((uint64_t)sweep >> 32) << 24) | ((uint32_t)q << 12) | colour;
            ++colour;
            k_canonical_pairs<<<grid, 128, 0, s->stream>>>(a);
          }
    }
    CMX_CUDA(cudaGetLastError());
  }
  return CMX_OK;
}

// counters since the last reset (null: only synchronise)
int cmx_canonical_counters(cmx_state *s, cmx_counters *counters) {
  CanonicalPlan &P = *s->canon;
  if (counters) {
    k_canonical_reduce<<<s->n_replicas, 32, 0, s->stream>>>(P.d_part, P.d_part_dE, P.blocks, s->d_counters);
    CMX_CUDA(cudaGetLastError());
    CMX_CUDA(cudaMemcpyAsync(counters, s->d_counters, sizeof(cmx_counters) * s->n_replicas,
                             cudaMemcpyDeviceToHost, s->stream));
  }
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  return CMX_OK;
}

extern "C" int cmx_canonical_sweep(cmx_state *s, int64_t n_sweeps, uint64_t seed,
                                   int64_t first_sweep, cmx_counters *counters) {
  int rc = cmx_canonical_enqueue(s, n_sweeps, seed, first_sweep, true);
  if (rc) return rc;
  return cmx_canonical_counters(s, counters);
}
