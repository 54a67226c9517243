// Internal definitions shared by the sm_100a translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "cmx_b200.h"

// CASM::KB [eV/K] -- the constant behind `beta = 1.0 / (CASM::KB * temperature)`
// (include/casm/clexmonte/methods/occupation_metropolis.hh:86); value pinned by
// the documented event state python/libcasm/clexmonte/_MonteCalculator.py:186-199.
#define CMX_KB 8.6173303e-05

void cmx_set_error(const std::string &msg);
int cmx_cuda_fail(cudaError_t e, const char *what);

#define CMX_CUDA(call)                                  \
  do {                                                  \
    cudaError_t _e = (call);                            \
    if (_e != cudaSuccess) return cmx_cuda_fail(_e, #call); \
  } while (0)

// Device view of one basis set (all pointers are device pointers).
struct DevTables {
  int32_t n_sublat, max_occ, n_func, corr_size, n_point_corr, nlist_len,
      n_nlist_sublat;
  const int32_t *nlist_sublat;
  const int32_t *n_occ;
  const double *phi;
  const int4 *nbr;  // (di, dj, dk, b)
  const int32_t *factor_f, *factor_n;
  const double *term_coef;
  const int32_t *term_fbeg, *elem_tbeg, *group_ebeg, *group_dphi,
      *group_has_sum;
  const double *group_div;
  const int32_t *global_gbeg, *point_gbeg, *delta_gbeg;
};

// Geometry of a (slab of a) supercell.
struct Geom {
  int32_t N0, N1, N2;  // local box (N2 = owned layers)
  int32_t halo;        // ghost layers on each k side (0 = periodic in k)
  int64_t n_cells;     // N0*N1*N2 (owned)
  int64_t layer;       // N0*N1
  int64_t sub_stride;  // bytes between sublattices: N0*N1*(N2+2*halo)
  int64_t rep_stride;  // bytes between replicas:   n_sublat*sub_stride
  // Storage code of the occupants.  coded == 0: the byte is the occupant index.
  // coded == 1 (single-sublattice ternary states): occupant 2 is stored as 18, so
  // that the byte-lane sum over a site's neighbors IS n1 + 18*n2, the index of
  // the pair-LUT sweep -- no per-neighbor masking -- and code & 3 is still the
  // occupant.  cmx_dec() reads both codes.
  int32_t coded;
  // Row layout.  xq_log == 0: site i of a row is byte i.  xq_log != 0 ("x4-interleaved
  // rows", N0 = 4 Q, Q = 1 << xq_log): site i is byte 4 (i mod Q) + i div Q, so the 32-bit
  // word w of a row holds the sites w, w + Q, w + 2Q, w + 3Q.  The x neighbors of a whole
  // word are then the words left and right of it (no byte shifts), and the sites of a word
  // share their x colour -- what the streaming pair-LUT sweep (cmx_sweep_stream.cuh) is
  // built on.  Chosen at creation for single-sublattice states whose N0 is a power of two
  // in [16, 512]; every kernel addresses sites through cmx_site_offset / cmx_xpos.
  int32_t xq_log;
  // General supercells (cmx_state_create_general): the supercell lattice in Hermite normal
  // form has the basis (N0, s10, s20), (0, N1, s21), (0, 0, N2) in prim coordinates, so the
  // unit cells are still the box 0 <= i < N0, 0 <= j < N1, 0 <= k < N2, but leaving it along
  // i also shifts j and k, leaving it along j shifts k (cmx_wrap_cell).  All zero: diag(N).
  int32_t s10, s20, s21;
};

__host__ __device__ __forceinline__ int cmx_xpos(const Geom &g, int i) {
  return g.xq_log ? (((i & ((1 << g.xq_log) - 1)) << 2) | (i >> g.xq_log)) : i;
}

// occupant index <-> stored byte (see Geom::coded)
__host__ __device__ __forceinline__ int cmx_dec(int raw) { return (raw & 7) | (raw >> 3); }
__host__ __device__ __forceinline__ int cmx_enc(const Geom &g, int v) {
  return (g.coded && v == 2) ? 18 : v;
}

struct cmx_tables {
  int device;
  DevTables d;               // device pointers
  std::vector<void *> allocs;  // for cleanup
  // host copies used when binding ECI / planning sweeps
  std::vector<int32_t> nlist_sublat, n_occ, nbr, factor_f, factor_n, term_fbeg,
      elem_tbeg, group_ebeg, group_dphi, group_has_sum, global_gbeg, point_gbeg,
      delta_gbeg;
  std::vector<double> phi, term_coef, group_div;
};

// exact division of 32-bit indices by a launch-constant divisor
struct FastDiv {
  uint32_t d, m;
};
__host__ __device__ static inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  f.m = (d <= 1) ? 0xFFFFFFFFu : (uint32_t)((1ull << 32) / d);
  return f;
}
__device__ __forceinline__ void fastdivmod(uint32_t n, FastDiv f, uint32_t &q,
                                           uint32_t &r) {
  q = __umulhi(n, f.m);
  r = n - q * f.d;
  if (r >= f.d) {
    r -= f.d;
    q += 1;
  }
}

// ---- checkerboard sweep plan (built by cmx_plan_sweep, see cmx_sweep.cu) ----
struct SweepPlan {
  bool valid = false;     // sweeps possible (global clexulator + ECI bound)
  bool pair_lut = false;  // pair-count LUT fast path usable
  bool rng16 = false;     // generic evaluator draws the pair16 kernel's random bits
  int32_t S[3] = {1, 1, 1};  // colour strides along i, j, k
  int32_t n_colours = 0;
  int32_t range_k = 0;  // max |dk| over active neighbors (halo depth needed)
  std::vector<int32_t> mut_points;  // point positions with > 1 occupant
  std::vector<int32_t> active_nbr;  // neighbor-list sites the bound ECI actually read
  // generic evaluator: ECI-folded, merged delta terms per point position
  int32_t n_gterms = 0;
  int32_t *d_gt_beg = nullptr;   // [n_point+1]
  double *d_gt_w = nullptr;      // [n_gterms][max_occ][max_occ]
  int32_t *d_gt_fbeg = nullptr;  // [n_gterms+1]
  int32_t *d_gt_f = nullptr, *d_gt_n = nullptr;
  // warp-cooperative evaluation: the neighbors a point position's terms read
  // (act_n[act_beg[p] .. act_beg[p+1])), and per factor the index of its staged
  // site-function value, slot * n_func + f
  int32_t *d_gt_vi = nullptr, *d_act_beg = nullptr, *d_act_n = nullptr;
  int32_t stage_max = 0;  // max over p of (active neighbors of p) * n_func
  // packed form for shared-memory staging: 4 x 16-bit value indices per term (unused
  // factors point at a slot that holds 1.0); 0 when a term has more than 4 factors
  uint2 *d_gt_pk = nullptr;
  int32_t pk_terms_max = 0, pk_act_max = 0;  // max over p of terms / active neighbors
  int generic_coop_capacity = -1;            // co-resident blocks of the cooperative generic kernel
  // pair-sum evaluator (k_sweep_pairsum): every folded term has at most one factor (point
  // and pair functions, any number of shells and sublattices), so dE is a sum of
  // per-neighbor tables V[slot][oi][of][neighbor occupant] (slots = act_n order) plus the
  // factor-free part P0[p][oi][of]
  bool pair_sum = false;
  double *d_ps_V = nullptr, *d_ps_P0 = nullptr;
  int32_t ps_range[3] = {0, 0, 0};  // max |offset| per axis over the active neighbors
  // pair LUT: one sublattice, <= 3 occupants, offsets in {-1,0,1}^3,
  // one class of symmetry-equivalent neighbors
  int32_t nocc = 0;
  int32_t z = 0;          // neighbors in the class
  int32_t shell[48] = {0};  // their offsets (di, dj, dk), at most 16
  uint32_t mask = 0;      // bit (dk+1)*9 + (dj+1)*3 + (di+1)
  int32_t n_lut = 0;      // nocc*(nocc-1)*256 entries
  double *d_pair_dE = nullptr;       // [n_lut] clex dE per (oi, alt, counts)
  int32_t n_tab = 0;              // entries of the pair16 acceptance table
  uint32_t *d_tab = nullptr;      // [replica][n_tab] threshold(15 bit)<<9 | 1<<8 | proposed code
  uint32_t *d_thr_lo = nullptr;   // [replica][n_tab] low 32 bits of the threshold (tie break)
  double *d_dEpot = nullptr;      // [replica][n_tab] dE - exch
  // two-class count table (k_sweep_pass16 with a second mask): point + pair bases whose active
  // neighbors form TWO classes (the reference's dense FCC ECI: 1NN + 2NN pairs) on
  // x4-interleaved rows.  Compact index: row(oi, alt) n1 n2 + i1 n2 + i2, i_c = index of the
  // class's (n_B, n_Va) among the combinations with n_B + n_Va <= z_c
  bool pair2 = false;
  uint32_t mask2 = 0;              // offsets of the second class (bit layout as `mask`)
  int32_t z2 = 0;
  int32_t shell2[48] = {0};
  int32_t n_tab2 = 0;              // 2 nocc x n1 x n2 entries
  double *d_pair2_dE = nullptr;    // [n_tab2] clex dE, summed in the pair-sum evaluator's order
  uint16_t *d_maps2 = nullptr;     // [544] index maps (S16Args::maps2)
  uint32_t *d_tab2 = nullptr;      // [replica][n_tab2] thr16 | proposed code << 16
  uint32_t *d_thr_lo2 = nullptr;   // [replica][n_tab2]
  double *d_dEpot2 = nullptr;      // [replica][n_tab2] dE - exch
  // streaming kernel (cmx_sweep_stream.cuh): x4-interleaved rows
  bool stream = false;
  int32_t n_tab24 = 0;            // entries of its acceptance table
  uint32_t *d_tab24 = nullptr;    // [replica][n_tab24] thr16 | proposed code << 16
  bool pdl_ok = false;            // (unused by the streaming kernel; kept for set_conditions)
  struct StreamList {             // the unit order of a call of (n sweeps, kgroup), built once
    struct StreamUnit *d_units = nullptr;
    uint32_t n_units = 0;
    uint32_t min_sep = 0;  // smallest distance (in units) between a unit and one it depends on
  };
  std::map<std::pair<int, int>, StreamList> stream_lists;
  uint32_t *d_ticket = nullptr;   // [replica] group tickets of the streaming kernel (dynamic assignment)
  uint32_t *d_gridbar = nullptr;  // [replica][32] arrivals at the colour-pass kernel's barrier (monotonic, wraps)
  uint32_t gridbar_count = 0;     // arrivals the launches so far have added (to every counter in use)
  int gridbar_mode = -1;          // 0: one counter per replica, 1: one for the whole grid (slabs)
  bool l2_window_set = false;     // persisting-L2 window of the lattice decided (colour passes on one GPU)
  int stream_capacity = -1;       // co-resident blocks of the kernel chosen for this state
  int stream_blocks = 0;          // blocks per replica
  uint32_t stream_gr = 1;         // row-steps per group
  uint32_t stream_gap = 0;        // units the host keeps between dependent units
  bool thr_dirty = true;
  // fast energy (cmx_energy.cu): per-cell energy as a function of (occupant,
  // species counts over the cell's "forward" neighbors), same byte-lane counting
  bool e_fast = false;
  uint32_t e_mask = 0;   // forward neighbor rows/offsets, bit layout as `mask`
  int32_t e_z = 0;
  double *d_e_lut = nullptr;  // [4][256]: index = cnt | occupant << 8
  bool e_lin = false;         // the table is linear in the counts: integer bond-count kernel
  double e_lin_c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // [occupant][c0, d1, d2]
  double *d_e_lin = nullptr;
  // streaming global correlations (cmx_energy.cu): functions grouped by forward-neighbor set
  int corr_lin_state = 0;  // 0: not planned yet, 1: usable, -1: not applicable
  int corr_n_masks = 0, corr_nocc = 0;
  uint32_t corr_mask[4] = {0, 0, 0, 0};
  int32_t corr_z[4] = {0, 0, 0, 0};
  int32_t *d_corr_func_mask = nullptr;  // [corr_size] index into corr_mask
  double *d_corr_lin = nullptr;         // [corr_size][9]: c0, d1, d2 per occupant
  // per-block partial counters of the current call
  long long *d_part_acc = nullptr;
  double *d_part_dE = nullptr;
  int32_t part_blocks = 0;
  long long attempts = 0;  // attempted steps per replica since the last reset
  // algorithmic work per attempted step, counted from the tables
  double bytes_per_step = 0, flops_per_step = 0;
};

struct cmx_state {
  const cmx_tables *t;
  Geom g;
  int32_t n_replicas;
  int32_t k_offset = 0;  // global k of local layer 0 (slab decomposition)
  int8_t *d_occ = nullptr;  // [replica][sublat][k(+halo)][j][i]
  // ECI
  int32_t n_eci = 0;
  uint32_t *d_eci_idx = nullptr;
  double *d_eci_val = nullptr;
  std::vector<uint32_t> eci_idx;
  std::vector<double> eci_val;
  // conditions
  std::vector<double> temperature;  // [replica]
  double *d_beta = nullptr;         // [replica]
  double *d_exch = nullptr;         // [replica][n_sublat][max_occ][max_occ]
  std::vector<double> exch;
  // occupants (sequential mode)
  std::vector<int32_t> sublat_to_asym, occ_to_species;
  int32_t n_species = 0;
  // sweeps
  SweepPlan plan;
  struct CanonicalPlan *canon = nullptr;  // cmx_canonical_set_swaps
  uint32_t sweep_flags = 0;            // CMX_SWEEP_* (cmx_state_set_sweep_flags)
  cmx_counters *d_counters = nullptr;  // [replica]
  int *d_flag = nullptr;               // device-side validation flag
  bool async_upload_pending = false;   // an asynchronous upload has not been checked yet
  // slab decomposition over NVLink peer memory (cmx_state_ipc_attach):
  // d_sig[0] / [1]: ring epochs (thin slabs), [3]: a dependency wait timed out
  unsigned long long *d_sig = nullptr;
  bool p2p = false;
  int8_t *peer_occ_dn = nullptr, *peer_occ_up = nullptr;
  unsigned long long *peer_sig_dn = nullptr, *peer_sig_up = nullptr;
  void *ipc_open[4] = {nullptr, nullptr, nullptr, nullptr};  // mappings to close
  // streaming sweeps: finished row-steps per layer, [replica][N2 + 2] ([0] and [N2+1]: the
  // ghost layers, counted up by the ring neighbours); lives behind d_sig in one allocation
  // so that one IPC handle exports both.  done_even / done_odd: the value every even / odd
  // layer's counter has when no sweep is in flight (all ranks agree: same call sequence)
  uint32_t *d_done = nullptr;
  uint32_t *peer_done_dn = nullptr, *peer_done_up = nullptr;
  uint32_t done_even = 0, done_odd = 0;
  // thin slabs (k_sweep_pass16): k-colour groups completed by this rank; d_sig[0] / [1]: the
  // epoch my lower / upper ring neighbour reached (written by them)
  unsigned long long epoch = 0;
  std::vector<int64_t> site_order;  // caller's l -> ours (upload / download), empty: identity
  long long seq_ties[17] = {0};  // tie report of the last cmx_metropolis_sequential call
  // scratch
  void *d_scratch = nullptr;
  size_t scratch_bytes = 0;
  cudaStream_t stream = nullptr;
};

int cmx_scratch(cmx_state *s, size_t bytes);
int cmx_plan_sweep(cmx_state *s);
int cmx_sgc_sweep_enqueue(cmx_state *s, uint64_t seed, int64_t first_sweep, int64_t n_sweeps);
bool cmx_use_warp_generic(const cmx_state *s);  // wide orbit sets: one site per warp
bool cmx_use_pair_sum(const cmx_state *s);      // point + pair bases outside the pair-LUT path
void cmx_canonical_free(cmx_state *s);
int cmx_canonical_enqueue(cmx_state *s, int64_t n_sweeps, uint64_t seed, int64_t first_sweep, bool reset);
int cmx_canonical_counters(cmx_state *s, cmx_counters *counters);
int cmx_plan_energy(cmx_state *s);                       // cmx_energy.cu
int cmx_energy_fast(cmx_state *s, int32_t replica, double *E);  // requires plan.e_fast
// integer bond counts of one block (k_energy_lin16); sl2 / sh2 carry a factor 16
struct LinSums {
  unsigned long long n1, n2, sl1, sh1, sl2, sh2;
};

// E and the occupant counts of one replica from the summed block counts
// v = (n1, n2, sl1, sh1, sl2, sh2):
//   N_o1 = sum n1 over cells with occupant o,  N_o2 = sum n2; every site is a forward
//   neighbor of exactly z cells, which gives the occupant-0 rows
__host__ __device__ inline double cmx_lin_energy(const unsigned long long *v, long long n_cells, int z,
                                                 const double *lin /*[3][3]: c0, d1, d2 per occupant*/,
                                                 unsigned long long *n_occ /*[3] or null*/) {
  const unsigned long long n1 = v[0], n2 = v[1], sl1 = v[2], sh1 = v[3], sl2 = v[4] >> 4, sh2 = v[5] >> 4;
  const long long N1 = (long long)n1, N2 = (long long)n2, N0 = n_cells - N1 - N2;
  const long long N12 = (long long)sh1, N11 = (long long)sl1 - 2 * N12;
  const long long N22 = (long long)sh2, N21 = (long long)sl2 - 2 * N22;
  const long long N01 = (long long)z * N1 - N11 - N21, N02 = (long long)z * N2 - N12 - N22;
  if (n_occ) {
    n_occ[0] = (unsigned long long)N0;
    n_occ[1] = n1;
    n_occ[2] = n2;
  }
  double E = (double)N0 * lin[0] + (double)N01 * lin[1] + (double)N02 * lin[2];
  E += (double)N1 * lin[3] + (double)N11 * lin[4] + (double)N12 * lin[5];
  E += (double)N2 * lin[6] + (double)N21 * lin[7] + (double)N22 * lin[8];
  return E;
}

#ifdef __CUDACC__
// sum of nb block records by one thread block (integers: any order); the result is
// valid in every thread after the call.  sh[6] is shared scratch.
__device__ __forceinline__ void cmx_lin_reduce(const LinSums *__restrict__ sums, int nb,
                                               unsigned long long *sh, unsigned long long (&v)[6]) {
  if (threadIdx.x < 6) sh[threadIdx.x] = 0;
  __syncthreads();
  unsigned long long t[6] = {0, 0, 0, 0, 0, 0};
  for (int b = threadIdx.x; b < nb; b += blockDim.x) {
    const LinSums r = sums[b];
    t[0] += r.n1;
    t[1] += r.n2;
    t[2] += r.sl1;
    t[3] += r.sh1;
    t[4] += r.sl2;
    t[5] += r.sh2;
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t[q] += __shfl_down_sync(0xffffffffu, t[q], o);
    if ((threadIdx.x & 31) == 0 && t[q]) atomicAdd(&sh[q], t[q]);
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 6; ++q) v[q] = sh[q];
}
#endif

int cmx_energy_fast_blocks(const cmx_state *s);
int cmx_energy_lin_batch(cmx_state *s, int nb, LinSums *d_sums);  // requires plan.e_lin
int cmx_global_corr_device(cmx_state *s, int32_t replica, double **d_out);  // cmx_faithful.cu; result in scratch
int cmx_plan_corr_lin(cmx_state *s);                                          // cmx_energy.cu
int cmx_global_corr_lin_device(cmx_state *s, int32_t replica, double **d_out);
int cmx_composition_device(cmx_state *s, int32_t replica, unsigned long long *d_counts, bool *bin0_missing);
void cmx_plan_free(SweepPlan &p);

// ---- device helpers ---------------------------------------------------------
__device__ __forceinline__ int cmx_wrap(int x, int n) {
  // x in [-n, 2n)
  x += (x < 0) ? n : 0;
  x -= (x >= n) ? n : 0;
  return x;
}

__host__ __device__ __forceinline__ int cmx_floor_div(int x, int n) {
  return (x >= 0) ? x / n : -((-x + n - 1) / n);
}
// unit cell (i, j, k) (any integers) -> its representative inside the box.  Slab states
// (halo) leave k alone.  diag(N) boxes take the single-wrap shortcut (|offset| <= N, checked
// at creation); skewed boxes reduce along i, then j, then k with the lattice basis above.
__host__ __device__ __forceinline__ void cmx_wrap_cell(const Geom &g, int &i, int &j, int &k) {
  if (g.s10 | g.s20 | g.s21) {
    const int q0 = cmx_floor_div(i, g.N0);
    i -= q0 * g.N0;
    j -= q0 * g.s10;
    k -= q0 * g.s20;
    const int q1 = cmx_floor_div(j, g.N1);
    j -= q1 * g.N1;
    k -= q1 * g.s21;
    k -= cmx_floor_div(k, g.N2) * g.N2;
    return;
  }
  i += (i < 0) ? g.N0 : 0;
  i -= (i >= g.N0) ? g.N0 : 0;
  j += (j < 0) ? g.N1 : 0;
  j -= (j >= g.N1) ? g.N1 : 0;
  if (!g.halo) {
    k += (k < 0) ? g.N2 : 0;
    k -= (k >= g.N2) ? g.N2 : 0;
  }
}

// byte offset of site (b; i,j,k) inside one replica
__device__ __forceinline__ int64_t cmx_site_offset(const Geom &g, int b, int i,
                                                   int j, int k) {
  return (int64_t)b * g.sub_stride + (int64_t)(k + g.halo) * g.layer +
         (int64_t)j * g.N0 + cmx_xpos(g, i);
}

// neighbor n of cell (i,j,k): byte offset + linear reference index l
__device__ __forceinline__ int64_t cmx_nbr_offset(const DevTables &T,
                                                  const Geom &g, int n, int i,
                                                  int j, int k, int64_t *l) {
  int4 o = T.nbr[n];
  int ii = i + o.x, jj = j + o.y, kk = k + o.z;
  cmx_wrap_cell(g, ii, jj, kk);
  if (l) {
    int kw = cmx_wrap(kk, g.N2);
    *l = (int64_t)o.w * g.n_cells + ((int64_t)kw * g.N1 + jj) * g.N0 + ii;
  }
  return cmx_site_offset(g, o.w, ii, jj, kk);
}

// Temporary occupant overrides (multi-site events are evaluated sequentially
// against a configuration in which the earlier sites already changed).
struct Override {
  int n;
  int64_t off[4];
  int occ[4];
};

__device__ __forceinline__ int cmx_load_occ(const int8_t *occ, int64_t off,
                                            const Override &ov) {
  int v = cmx_dec(occ[off]);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (q < ov.n && ov.off[q] == off) v = ov.occ[q];
  return v;
}

// Occupant fetchers for the evaluator below.
struct LatticeFetch {
  const DevTables &T;
  const Geom &g;
  const int8_t *occ;
  int i, j, k;
  const Override &ov;
  __device__ __forceinline__ int operator()(int n) const {
    int64_t off = cmx_nbr_offset(T, g, n, i, j, k, nullptr);
    return cmx_load_occ(occ, off, ov);
  }
};
struct LocalFetch {  // synthetic neighborhood (LUT construction)
  const int8_t *nbr_occ;
  __device__ __forceinline__ int operator()(int n) const { return nbr_occ[n]; }
};

// ---- warp-cooperative folded delta E (generic evaluator) -----------------------
struct GenTerms {
  const int32_t *gt_beg, *gt_fbeg, *gt_vi, *act_beg, *act_n;
  const double *gt_w;
  const uint2 *gt_pk;  // packed value indices, or null
};
// the tables of ONE point position staged in shared memory (cmx_gen_stage_table)
struct GenShared {
  const uint2 *vi;    // [n_terms] 4 x 16-bit indices into the staged values
  const double *w;    // [n_terms][max_occ][max_occ]
  const int4 *act;    // [n_act] neighbor offsets (di, dj, dk, sublattice)
  int n_terms, n_act;
};
__host__ __device__ inline size_t cmx_gen_shared_bytes(int n_terms, int n_act, int max_occ) {
  return (size_t)n_terms * (8 + 8 * max_occ * max_occ) + (size_t)n_act * 16;
}
// Single-site delta E of point position p at cell (i,j,k), occupant oi -> of, with one
// overridden site (byte offset ov_off holds occupant ov_occ), evaluated by ONE WARP:
//  1. the lanes stage phi_f(occupant) of every neighbor the terms of p read into the
//     warp's shared array (one gather + wrap arithmetic per neighbor instead of one per
//     factor: ZrO reads 225 neighbors through ~1400 factors);
//  2. the terms are dealt to the lanes round robin, each a product of staged values;
//  3. butterfly sum: every lane returns the same bits, whatever the grid.
// CG: occupations through L2 only (kernels that synchronise the grid between colours).
template <bool CG>
__device__ __forceinline__ double cmx_warp_site_delta(const DevTables &T, const Geom &g, const GenTerms &G,
                                                      const int8_t *occ, double *sh_val, int p, int i, int j,
                                                      int k, int oi, int of, int64_t ov_off, int ov_occ,
                                                      unsigned lane) {
  const int mo = T.max_occ, nf = T.n_func;
  const int ab = G.act_beg[p], ae = G.act_beg[p + 1];
#pragma unroll 4
  for (int s = ab + (int)lane; s < ae; s += 32) {
    const int n = G.act_n[s];
    const int64_t no = cmx_nbr_offset(T, g, n, i, j, k, nullptr);
    const int raw = CG ? (int)__ldcg(occ + no) : (int)occ[no];
    const int o = (no == ov_off) ? ov_occ : cmx_dec(raw);
    const double *ph = T.phi + ((size_t)T.nbr[n].w * nf) * mo + o;
    double *dst = sh_val + (size_t)(s - ab) * nf;
    for (int f = 0; f < nf; ++f) dst[f] = ph[(size_t)f * mo];
  }
  __syncwarp();
  double part = 0.0;
#pragma unroll 4
  for (int t = G.gt_beg[p] + (int)lane; t < G.gt_beg[p + 1]; t += 32) {
    double v = G.gt_w[((size_t)t * mo + oi) * mo + of];
    const int qb = G.gt_fbeg[t], qe = G.gt_fbeg[t + 1];
    for (int q = qb; q < qe; ++q) v *= sh_val[G.gt_vi[q]];
    part += v;
  }
  __syncwarp();  // the staged values may be overwritten by the caller's next evaluation
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  return part;
}

// Block-wide: copy the packed term table of point position p into shared memory at
// `smem` (16-byte aligned, cmx_gen_shared_bytes); ends with a barrier.
__device__ __forceinline__ void cmx_gen_stage_table(const DevTables &T, const GenTerms &G, int p,
                                                    unsigned char *smem, GenShared &S) {
  const int mo2 = T.max_occ * T.max_occ;
  const int tb = G.gt_beg[p], ab = G.act_beg[p];
  S.n_terms = G.gt_beg[p + 1] - tb;
  S.n_act = G.act_beg[p + 1] - ab;
  int4 *act = reinterpret_cast<int4 *>(smem);  // 16-byte entries first: alignment
  double *w = reinterpret_cast<double *>(act + S.n_act);
  uint2 *vi = reinterpret_cast<uint2 *>(w + (size_t)S.n_terms * mo2);
  for (int q = threadIdx.x; q < S.n_terms * mo2; q += blockDim.x) w[q] = G.gt_w[(size_t)tb * mo2 + q];
  for (int q = threadIdx.x; q < S.n_terms; q += blockDim.x) vi[q] = G.gt_pk[tb + q];
  for (int q = threadIdx.x; q < S.n_act; q += blockDim.x) act[q] = T.nbr[G.act_n[ab + q]];
  S.w = w;
  S.vi = vi;
  S.act = act;
  __syncthreads();
}

// cmx_warp_site_delta with the term table in shared memory: every read of the term loop
// (indices, weight, staged values) is a shared-memory read, the products of a lane's terms
// are independent and pipeline; same products in the same order as cmx_warp_site_delta,
// so the same bits.  sh_val[n_act * n_func] must hold 1.0 (the unused-factor slot).
template <bool CG>
__device__ __forceinline__ double cmx_warp_site_delta_sh(const DevTables &T, const Geom &g, const GenShared &S,
                                                         const int8_t *occ, double *sh_val, int i, int j, int k,
                                                         int oi, int of, int64_t ov_off, int ov_occ,
                                                         unsigned lane) {
  const int mo = T.max_occ, nf = T.n_func;
#pragma unroll 4
  for (int s = (int)lane; s < S.n_act; s += 32) {
    const int4 o = S.act[s];
    int ii = i + o.x, jj = j + o.y, kk = k + o.z;
    cmx_wrap_cell(g, ii, jj, kk);
    const int64_t no = cmx_site_offset(g, o.w, ii, jj, kk);
    const int raw = CG ? (int)__ldcg(occ + no) : (int)occ[no];
    const int oc = (no == ov_off) ? ov_occ : cmx_dec(raw);
    const double *ph = T.phi + ((size_t)o.w * nf) * mo + oc;
    double *dst = sh_val + (size_t)s * nf;
    for (int f = 0; f < nf; ++f) dst[f] = ph[(size_t)f * mo];
  }
  __syncwarp();
  const double *w = S.w + oi * mo + of;
  const int mo2 = mo * mo;
  double part = 0.0;
#pragma unroll 4
  for (int t = (int)lane; t < S.n_terms; t += 32) {
    const uint2 pk = S.vi[t];
    double v = w[(size_t)t * mo2];
    v *= sh_val[pk.x & 0xFFFFu];
    v *= sh_val[pk.x >> 16];
    v *= sh_val[pk.y & 0xFFFFu];
    v *= sh_val[pk.y >> 16];
    part += v;
  }
  __syncwarp();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  return part;
}

// Faithful evaluation of one generated function (see clexulator_tables.py for
// the canonical form): same association order as the C++ source, every
// operation individually rounded (no FMA contraction).
template <class Fetch>
__device__ inline double cmx_eval_function_t(const DevTables &T, int gbeg,
                                             int gend, const Fetch &fetch,
                                             int pb, int occ_i, int occ_f) {
  double total = 0.0;
  bool have_total = false;
  for (int gi = gbeg; gi < gend; ++gi) {
    double s = 0.0;
    bool has_sum = T.group_has_sum[gi] != 0;
    if (has_sum) {
      bool have_s = false;
      for (int e = T.group_ebeg[gi]; e < T.group_ebeg[gi + 1]; ++e) {
        double ev = 0.0;
        bool have_e = false;
        for (int t = T.elem_tbeg[e]; t < T.elem_tbeg[e + 1]; ++t) {
          double tv = T.term_coef[t];
          for (int q = T.term_fbeg[t]; q < T.term_fbeg[t + 1]; ++q) {
            int n = T.factor_n[q];
            int o = fetch(n);
            int b = T.nbr[n].w;
            double ph = T.phi[(b * T.n_func + T.factor_f[q]) * T.max_occ + o];
            tv = __dmul_rn(tv, ph);
          }
          ev = have_e ? __dadd_rn(ev, tv) : tv;
          have_e = true;
        }
        s = have_s ? __dadd_rn(s, ev) : ev;
        have_s = true;
      }
    }
    double v = s;
    int fd = T.group_dphi[gi];
    if (fd >= 0) {
      const double *ph = T.phi + (pb * T.n_func + fd) * T.max_occ;
      double d = __dsub_rn(ph[occ_f], ph[occ_i]);
      v = has_sum ? __dmul_rn(d, s) : d;
    }
    double dv = T.group_div[gi];
    if (dv != 0.0) v = __ddiv_rn(v, dv);
    total = have_total ? __dadd_rn(total, v) : v;
    have_total = true;
  }
  return total;
}

__device__ inline double cmx_eval_function(const DevTables &T, const Geom &g,
                                           const int8_t *occ, int gbeg,
                                           int gend, int i, int j, int k,
                                           const Override &ov, int pb,
                                           int occ_i, int occ_f) {
  LatticeFetch f{T, g, occ, i, j, k, ov};
  return cmx_eval_function_t(T, gbeg, gend, f, pb, occ_i, occ_f);
}

// ---- counter-based RNG: Philox4x32-10 (Salmon et al., SC'11) ---------------
struct Philox {
  uint32_t c[4];
};
template <int ROUNDS>
__device__ __forceinline__ Philox philox4x32(uint32_t c0, uint32_t c1,
                                             uint32_t c2, uint32_t c3,
                                             uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0;
    uint64_t p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += W0;
    k1 += W1;
  }
  Philox o;
  o.c[0] = c0;
  o.c[1] = c1;
  o.c[2] = c2;
  o.c[3] = c3;
  return o;
}

__device__ __forceinline__ Philox philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                uint32_t k0, uint32_t k1) {
  return philox4x32<10>(c0, c1, c2, c3, k0, k1);
}
// The semi-grand checkerboard sweeps draw from Philox4x32-7: seven rounds are the fewest
// that pass BigCrush ("Crush-resistant", Salmon et al., SC'11, table 2; ten add a safety
// margin).  The sweep kernels are issue bound and the generator is a third of their
// instructions; every evaluator of a sweep (streaming, block, generic) uses the same
// count, so they still agree bit for bit.
#define CMX_SWEEP_PHILOX_ROUNDS 7
__device__ __forceinline__ Philox philox_sweep(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
  return philox4x32<CMX_SWEEP_PHILOX_ROUNDS>(c0, c1, c2, c3, k0, k1);
}

// Same generator with the round keys (k0 + r*W0, k1 + r*W1) taken from a
// precomputed schedule: when `rk` lives in the kernel parameter (constant) bank
// the xors read it directly and the 20 key additions per call disappear.
struct PhiloxKeys {
  uint32_t k[20];
};
static inline PhiloxKeys philox_key_schedule(uint32_t k0, uint32_t k1) {
  PhiloxKeys s;
  for (int r = 0; r < 10; ++r) {
    s.k[2 * r] = k0 + (uint32_t)r * 0x9E3779B9u;
    s.k[2 * r + 1] = k1 + (uint32_t)r * 0xBB67AE85u;
  }
  return s;
}
template <int ROUNDS>
__device__ __forceinline__ Philox philox4x32_rk(uint32_t c0, uint32_t c1, uint32_t c2,
                                                uint32_t c3, const PhiloxKeys &rk) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0;
    uint64_t p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk.k[2 * r];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk.k[2 * r + 1];
    uint32_t n3 = (uint32_t)p0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
  }
  Philox o;
  o.c[0] = c0;
  o.c[1] = c1;
  o.c[2] = c2;
  o.c[3] = c3;
  return o;
}
__device__ __forceinline__ Philox philox_sweep_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                  const PhiloxKeys &rk) {
  return philox4x32_rk<CMX_SWEEP_PHILOX_ROUNDS>(c0, c1, c2, c3, rk);
}
