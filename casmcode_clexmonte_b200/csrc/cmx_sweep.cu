// Semi-grand canonical sublattice-checkerboard sweeps.
//
// Replaces the sequential loop of methods/occupation_metropolis.hh:92-120 +
// the semi-grand proposal (SemiGrandCanonicalCalculator.cc:104-120) + the
// potential delta (:186-213) by colour-by-colour simultaneous updates of
// non-interacting site sets.  Two evaluators:
//
//  * "pair_lut": the bound ECI select only point + pair functions on a single
//    sublattice with <= 3 occupants whose active neighbors all lie in
//    {-1,0,1}^3 and form one symmetry class (FCC/BCC/SC nearest-neighbor
//    models).  dE then depends only on (occ_i, occ_f, neighbor species
//    counts); it is tabulated ONCE by running the faithful evaluator on a
//    representative neighborhood per count combination, and the Metropolis
//    test  u < exp(-beta dE)  becomes a 53-bit integer compare against a
//    per-replica threshold table staged in shared memory.  Four sites per
//    thread, occupants gathered as 8-byte row chunks and counted bytewise.
//  * "generic": any table (multi-sublattice, triplets, quadruplets): ECI-folded
//    merged term lists, one site per thread, FP64 products, exp().
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include "cmx_internal.cuh"

static int invalid(const std::string &msg) {
  cmx_set_error(msg);
  return CMX_ERR_INVALID;
}

void cmx_plan_free(SweepPlan &p) {
  cudaFree(p.d_gt_beg);
  cudaFree(p.d_gt_w);
  cudaFree(p.d_gt_fbeg);
  cudaFree(p.d_gt_f);
  cudaFree(p.d_gt_n);
  cudaFree(p.d_pair_dE);
  cudaFree(p.d_thr);
  cudaFree(p.d_thr_lo);
  cudaFree(p.d_dEpot);
  cudaFree(p.d_part_acc);
  cudaFree(p.d_part_dE);
  p = SweepPlan();
}

// ---------------------------------------------------------------------------
// LUT construction
// ---------------------------------------------------------------------------
// entry e = (pair << 8) | counts,  pair = oi*(nocc-1)+alt,
// counts = nB | (nVa << 4)   (species 1 in the low nibble, species 2 high)
__global__ void k_build_pair_lut(DevTables T, int nocc, int z,
                                 const int32_t *__restrict__ class_nbr, int n_eci,
                                 const uint32_t *__restrict__ eci_idx,
                                 const double *__restrict__ eci_val,
                                 double *__restrict__ lut, int n_lut) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_lut) return;
  int pair = e >> 8, cnt = e & 255;
  int n1 = cnt & 15, n2 = cnt >> 4;
  int oi = pair / (nocc - 1), alt = pair - oi * (nocc - 1);
  int of = oi + 1 + alt;
  if (of >= nocc) of -= nocc;
  if (n1 + n2 > z || (nocc < 3 && n2 > 0)) {
    lut[e] = 0.0;  // unreachable combination
    return;
  }
  int8_t nb[64];
  for (int n = 0; n < T.nlist_len && n < 64; ++n) nb[n] = 0;
  for (int q = 0; q < z; ++q) nb[class_nbr[q]] = (q < n1) ? 1 : ((q < n1 + n2) ? 2 : 0);
  nb[0] = (int8_t)oi;
  LocalFetch f{nb};
  int b = T.nlist_sublat[0];
  double dE = 0.0;
  for (int q = 0; q < n_eci; ++q) {
    int fi = (int)eci_idx[q];  // point position 0
    double d = cmx_eval_function_t(T, T.delta_gbeg[fi], T.delta_gbeg[fi + 1], f, b, oi, of);
    dE = __dadd_rn(dE, __dmul_rn(eci_val[q], d));
  }
  lut[e] = dE;
}

// thresholds: accept  <=>  r53 < thr,  r53 uniform on [0, 2^53):
//   dE < 0            -> always            (metropolis_acceptance [EXT])
//   else u < exp(-dE*beta), u = r53 * 2^-53
// stored split: thr_hi = thr >> 22 (compared against the 31 random bits every
// site draws) and thr_lo = thr & (2^22-1) (22 more bits, drawn lazily only when
// the first 31 tie) -- together exactly the 53-bit comparison.
__global__ void k_build_thresholds(const double *__restrict__ lut, int n_lut,
                                   int nocc, int max_occ, int b,
                                   const double *__restrict__ beta,
                                   const double *__restrict__ exch, int exch_stride,
                                   uint32_t *__restrict__ thr_hi,
                                   uint32_t *__restrict__ thr_lo,
                                   double *__restrict__ dEpot) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  int r = blockIdx.y;
  if (e >= n_lut) return;
  int pair = e >> 8;
  int oi = pair / (nocc - 1), alt = pair - oi * (nocc - 1);
  int of = oi + 1 + alt;
  if (of >= nocc) of -= nocc;
  double x = exch[(size_t)r * exch_stride + (b * max_occ + oi) * max_occ + of];
  double dE = __dsub_rn(lut[e], x);
  unsigned long long t;
  const double two53 = 9007199254740992.0;
  if (dE < 0.0) {
    t = 1ull << 53;
  } else {
    double p = exp(-dE * beta[r]);
    double v = ceil(p * two53);
    t = (unsigned long long)v;
  }
  thr_hi[(size_t)r * n_lut + e] = (uint32_t)(t >> 22);
  thr_lo[(size_t)r * n_lut + e] = (uint32_t)(t & 0x3FFFFFull);
  dEpot[(size_t)r * n_lut + e] = dE;
}

// ---------------------------------------------------------------------------
// pair-LUT sweep kernel
// ---------------------------------------------------------------------------
struct FastDiv {
  uint32_t d, m;
};
static FastDiv make_fastdiv(uint32_t d) {
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

struct PairSweepArgs {
  int8_t *occ;  // replica 0 base (start of the low ghost layers)
  Geom g;
  int cy, cz;          // colour parities along j, k (i parity is a template arg)
  uint32_t mask;       // runtime neighbor mask
  FastDiv divW, divJ;  // chunks per row, rows per layer of this colour
  uint32_t W, J, K;    // item grid of one colour: chunk, row pair, layer pair
  uint32_t dc, djj, dkk;  // grid stride decomposed on (c, jj, kk)
  const uint32_t *thr_hi;  // [replica][n_lut]
  const uint32_t *thr_lo;  // [replica][n_lut]
  const double *dEpot;     // [replica][n_lut]
  int n_lut;
  long long *part_acc;  // [replica][gridDim.x]
  double *part_dE;
  uint32_t k0, k1;       // seed
  uint32_t sweep_lo;     // RNG counter words
  uint32_t ctr_hi;       // (sweep_hi << 16) | (colour << 8)
  int k_offset;          // global k of local layer 0 (slab decomposition)
};

template <int CX>
__device__ __forceinline__ uint32_t extract4(uint32_t lo, uint32_t hi,
                                             uint32_t side, int dx) {
  // the four neighbor bytes (one per target site) at x-offset dx
  if (CX == 0) {  // targets at bytes 0,2,4,6
    if (dx == 0) return __byte_perm(lo, hi, 0x6420);
    if (dx == 1) return __byte_perm(lo, hi, 0x7531);
    uint32_t v = __byte_perm(lo, hi, 0x5310);  // [b0,b1,b3,b5]
    return __byte_perm(v, side, 0x3217);       // [prev.3,b1,b3,b5]
  } else {  // targets at bytes 1,3,5,7
    if (dx == 0) return __byte_perm(lo, hi, 0x7531);
    if (dx == -1) return __byte_perm(lo, hi, 0x6420);
    uint32_t v = __byte_perm(lo, hi, 0x7642);  // [b2,b4,b6,b7]
    return __byte_perm(v, side, 0x4210);       // [b2,b4,b6,next.0]
  }
}

// One thread owns 4 same-colour sites of one 8-byte row chunk per iteration.
//  * occupants of the <= 9 neighbor rows arrive as 8-byte loads (+ one 4-byte
//    side word where the colour needs the byte just outside the chunk),
//    32-bit offsets from the replica base;
//  * the species counts of the 4 sites are accumulated bytewise in one
//    register (species 1 in the low nibble, species 2 in the high nibble);
//  * one Philox4x32-10 call yields the 4 x (31-bit uniform + 1 proposal bit);
//    the acceptance test is one integer compare against the shared-memory
//    threshold table, the remaining 22 bits of the 53-bit uniform are drawn
//    only on a tie of the first 31 (probability 2^-31);
//  * accepted dE (tabulated) is summed in FP64 per thread, reduced per block.
template <int CX, int NOCC, uint32_t MASK_CT>
__global__ void __launch_bounds__(256, 4)
    k_sweep_pair_lut(PairSweepArgs a) {
  __shared__ uint32_t sh_thr[NOCC * (NOCC - 1) * 256];
  __shared__ double sh_dE[NOCC * (NOCC - 1) * 256];
  __shared__ long long sh_acc[8];
  __shared__ double sh_sum[8];
  const int r = blockIdx.y;
  {
    const uint32_t *gt = a.thr_hi + (size_t)r * a.n_lut;
    const double *ge = a.dEpot + (size_t)r * a.n_lut;
    for (int q = threadIdx.x; q < a.n_lut; q += blockDim.x) {
      sh_thr[q] = gt[q];
      sh_dE[q] = ge[q];
    }
  }
  __syncthreads();
  const uint32_t mask = MASK_CT ? MASK_CT : a.mask;
  const Geom &g = a.g;
  int8_t *base = a.occ + (size_t)r * g.rep_stride;  // includes the ghost layers
  // keep the replica base in a register pair: without this the compiler
  // re-derives r * rep_stride (a 64-bit multiply-add) for every load
  asm volatile("" : "+l"(base));
  const uint32_t N0 = g.N0, N1 = g.N1, N2 = g.N2;
  const uint32_t layer = N0 * N1;
  const bool halo = g.halo != 0;
  uint32_t n_acc = 0;  // < 2^32 accepted sites per thread and launch
  double e_sum = 0.0;

  uint32_t c, jj, kk;
  {
    uint32_t item = blockIdx.x * blockDim.x + threadIdx.x, row;
    fastdivmod(item, a.divW, row, c);
    fastdivmod(row, a.divJ, kk, jj);
  }
  while (kk < a.K) {
    const uint32_t j = 2 * jj + a.cy, k = 2 * kk + a.cz;
    const uint32_t x0 = 8 * c;
    // 32-bit byte offsets of the neighbor rows relative to the replica base
    const uint32_t off_c = ((k + g.halo) * N1 + j) * N0 + x0;
    uint32_t dj[3], dk[3];
    dj[0] = (j == 0) ? (N1 - 1) * N0 : 0u - N0;
    dj[1] = 0;
    dj[2] = (j == N1 - 1) ? 0u - (N1 - 1) * N0 : N0;
    dk[0] = (!halo && k == 0) ? (N2 - 1) * layer : 0u - layer;
    dk[1] = 0;
    dk[2] = (!halo && k == N2 - 1) ? 0u - (N2 - 1) * layer : layer;
    // the side word holds the byte just outside the chunk on the side this
    // colour needs: x0-1 for CX == 0 (dx = -1), x0+8 for CX == 1 (dx = +1)
    const uint32_t dside = (CX == 0) ? ((x0 == 0) ? N0 - 4 : 0u - 4u)
                                     : ((x0 + 8 == N0) ? 8u - N0 : 8u);
    // bytewise sums over the neighbors of the 4 sites: s1 = sum of occupant
    // codes (n1 + 2 n2 per lane), s2 = sum of (code & 2) (2 n2 per lane)
    uint32_t s1 = 0, s2 = 0, self4 = 0;
    uint32_t lo_c = 0, hi_c = 0;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const uint32_t m3 = (mask >> ((dz + 1) * 9 + (dy + 1) * 3)) & 7u;
        const bool center = (dz == 0 && dy == 0);
        if (m3 == 0 && !center) continue;
        const uint32_t off = off_c + dk[dz + 1] + dj[dy + 1];
        const uint2 ch = *reinterpret_cast<const uint2 *>(base + off);
        uint32_t side = 0;
        const bool need_side = (CX == 0) ? (m3 & 1u) : (m3 & 4u);
        if (need_side) side = *reinterpret_cast<const uint32_t *>(base + (off + dside));
        if (center) {
          lo_c = ch.x;
          hi_c = ch.y;
          self4 = extract4<CX>(ch.x, ch.y, 0, 0);
        }
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          if (!(m3 & (1u << (dx + 1)))) continue;
          uint32_t w = extract4<CX>(ch.x, ch.y, side, dx);
          s1 += w;
          s2 += w & 0x02020202u;
        }
      }
    }
    // species-1 count in the low nibble, species-2 count in the high nibble of
    // each byte lane: (s1 - s2) + 8 * s2 = n1 + 16 n2
    const uint32_t acc = (s1 - s2) + (s2 << 3);
    // ---- random numbers: one 32-bit word per site
    const uint32_t gid = ((k + (uint32_t)a.k_offset) * N1 + j) * a.W + c;
    const Philox ph = philox4x32_10(gid, (uint32_t)r, a.sweep_lo, a.ctr_hi, a.k0, a.k1);
    uint32_t new4 = 0;
    uint32_t ties = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t w = ph.c[q];
      const uint32_t oi = (self4 >> (8 * q)) & 0xffu;
      const uint32_t alt = (NOCC == 3) ? (w & 1u) : 0u;
      const uint32_t pair = oi * (NOCC - 1) + alt;
      // of = (oi + 1 + alt) mod NOCC, two bits per pair packed in a constant
      const uint32_t of = (NOCC == 3) ? ((0x429u >> (2 * pair)) & 3u) : (oi ^ 1u);
      const uint32_t cnt = (acc >> (8 * q)) & 0xffu;
      const uint32_t idx = (pair << 8) | cnt;
      const uint32_t thr = sh_thr[idx];
      const uint32_t u31 = w >> 1;
      const bool ok = u31 < thr;
      ties |= (u31 == thr) ? (1u << q) : 0u;
      double d = 0.0;
      if (ok) d = sh_dE[idx];
      e_sum += d;
      n_acc += ok ? 1u : 0u;
      new4 |= (ok ? of : oi) << (8 * q);
    }
    if (ties) {  // probability 2^-31 per site: draw the low 22 bits of the uniform
      const Philox p2 = philox4x32_10(gid, (uint32_t)r, a.sweep_lo, a.ctr_hi | 1u, a.k0, a.k1);
      const uint32_t *gl = a.thr_lo + (size_t)r * a.n_lut;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (!(ties & (1u << q))) continue;
        const uint32_t w = ph.c[q];
        const uint32_t oi = (self4 >> (8 * q)) & 0xffu;
        const uint32_t alt = (NOCC == 3) ? (w & 1u) : 0u;
        uint32_t of = oi + 1u + alt;
        of -= (of >= (uint32_t)NOCC) ? (uint32_t)NOCC : 0u;
        const uint32_t idx = ((oi * (NOCC - 1) + alt) << 8) | ((acc >> (8 * q)) & 0xffu);
        if ((p2.c[q] & 0x3FFFFFu) < gl[idx]) {
          new4 = (new4 & ~(0xffu << (8 * q))) | (of << (8 * q));
          e_sum += sh_dE[idx];
          n_acc += 1;
        }
      }
    }
    if (new4 != self4) {
      uint2 out;
      if (CX == 0) {
        out.x = __byte_perm(lo_c, new4, 0x3514);
        out.y = __byte_perm(hi_c, new4, 0x3716);
      } else {
        out.x = __byte_perm(lo_c, new4, 0x5240);
        out.y = __byte_perm(hi_c, new4, 0x7260);
      }
      *reinterpret_cast<uint2 *>(base + off_c) = out;
    }
    // ---- next item: grid stride decomposed on (c, jj, kk)
    c += a.dc;
    if (c >= a.W) {
      c -= a.W;
      jj += 1;
    }
    jj += a.djj;
    if (jj >= a.J) {
      jj -= a.J;
      kk += 1;
    }
    kk += a.dkk;
  }
  // ---- block reduction of the counters (fixed order -> deterministic)
  long long n_acc64 = n_acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_acc64 += __shfl_down_sync(0xffffffffu, n_acc64, o);
    e_sum += __shfl_down_sync(0xffffffffu, e_sum, o);
  }
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    sh_acc[wid] = n_acc64;
    sh_sum[wid] = e_sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long A = 0;
    double E = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      A += sh_acc[w];
      E += sh_sum[w];
    }
    size_t slot = (size_t)r * gridDim.x + blockIdx.x;
    a.part_acc[slot] += A;
    a.part_dE[slot] += E;
  }
}

// ---------------------------------------------------------------------------
// generic sweep kernel: one site per thread
// ---------------------------------------------------------------------------
struct GenericSweepArgs {
  int8_t *occ;
  Geom g;
  DevTables T;
  int S0, S1, S2, c0, c1, c2, p;  // colour
  FastDiv div0, div1;             // N0/S0, N1/S1
  uint32_t items;
  const int32_t *gt_beg, *gt_fbeg, *gt_f, *gt_n;
  const double *gt_w;
  const double *beta, *exch;
  int exch_stride;
  long long *part_acc;
  double *part_dE;
  uint32_t k0, k1, sweep_lo, ctr_hi;
  int k_offset;
};

__global__ void __launch_bounds__(256) k_sweep_generic(GenericSweepArgs a) {
  __shared__ long long sh_acc[8];
  __shared__ double sh_sum[8];
  const int r = blockIdx.y;
  const Geom &g = a.g;
  const DevTables &T = a.T;
  int8_t *occ = a.occ + (size_t)r * g.rep_stride;
  const int b = T.nlist_sublat[a.p];
  const int nocc = T.n_occ[b];
  const int mo = T.max_occ;
  const double beta = a.beta[r];
  const double *exch = a.exch + (size_t)r * a.exch_stride + (size_t)b * mo * mo;
  const int tb = a.gt_beg[a.p], te = a.gt_beg[a.p + 1];
  long long n_acc = 0;
  double e_sum = 0.0;
  for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < a.items;
       item += gridDim.x * blockDim.x) {
    uint32_t row, ii, kk, jj;
    fastdivmod(item, a.div0, row, ii);
    fastdivmod(row, a.div1, kk, jj);
    const int i = (int)ii * a.S0 + a.c0, j = (int)jj * a.S1 + a.c1,
              k = (int)kk * a.S2 + a.c2;
    const int64_t off = cmx_site_offset(g, b, i, j, k);
    const int oi = occ[off];
    // global site id -> RNG counter
    const uint32_t gid_lo = (uint32_t)(((uint64_t)(k + a.k_offset) * g.N1 + j) * g.N0 + i);
    Philox ph = philox4x32_10(gid_lo, (uint32_t)r | ((uint32_t)b << 24), a.sweep_lo,
                              a.ctr_hi, a.k0, a.k1);
    const int alt = (int)__umulhi(ph.c[2], (uint32_t)(nocc - 1));
    int of = oi + 1 + alt;
    if (of >= nocc) of -= nocc;
    double dE = 0.0;
    for (int t = tb; t < te; ++t) {
      double v = a.gt_w[((size_t)t * mo + oi) * mo + of];
      for (int q = a.gt_fbeg[t]; q < a.gt_fbeg[t + 1]; ++q) {
        const int n = a.gt_n[q];
        const int64_t no = cmx_nbr_offset(T, g, n, i, j, k, nullptr);
        const int o = occ[no];
        v *= T.phi[((size_t)T.nbr[n].w * T.n_func + a.gt_f[q]) * mo + o];
      }
      dE += v;
    }
    dE -= exch[oi * mo + of];
    bool accept = dE < 0.0;
    if (!accept) {
      const unsigned long long u53 = ((unsigned long long)(ph.c[1] & 0x1FFFFFu) << 32) | ph.c[0];
      const double u = (double)u53 * (1.0 / 9007199254740992.0);
      accept = u < exp(-dE * beta);
    }
    if (accept) {
      occ[off] = (int8_t)of;
      ++n_acc;
      e_sum += dE;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_acc += __shfl_down_sync(0xffffffffu, n_acc, o);
    e_sum += __shfl_down_sync(0xffffffffu, e_sum, o);
  }
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    sh_acc[wid] = n_acc;
    sh_sum[wid] = e_sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long A = 0;
    double E = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      A += sh_acc[w];
      E += sh_sum[w];
    }
    size_t slot = (size_t)r * gridDim.x + blockIdx.x;
    a.part_acc[slot] += A;
    a.part_dE[slot] += E;
  }
}

__global__ void k_reduce_counters(const long long *__restrict__ part_acc,
                                  const double *__restrict__ part_dE, int nb,
                                  long long attempts, cmx_counters *out) {
  int r = blockIdx.x;
  if (threadIdx.x != 0) return;
  long long A = 0;
  double E = 0.0;
  for (int q = 0; q < nb; ++q) {
    A += part_acc[(size_t)r * nb + q];
    E += part_dE[(size_t)r * nb + q];
  }
  out[r].n_attempt = attempts;
  out[r].n_accept = A;
  out[r].dE_sum = E;
  out[r].reserved = 0;
}

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------
template <typename T>
static int to_device(const std::vector<T> &v, T **d) {
  size_t bytes = (v.size() ? v.size() : 1) * sizeof(T);
  CMX_CUDA(cudaMalloc((void **)d, bytes));
  if (!v.empty())
    CMX_CUDA(cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return CMX_OK;
}

static int smallest_divisor_above(int N, int R) {
  for (int s = R + 1; s <= N; ++s)
    if (N % s == 0) return s;
  return N;
}

int cmx_plan_sweep(cmx_state *s) {
  cmx_plan_free(s->plan);
  SweepPlan &P = s->plan;
  const cmx_tables *t = s->t;
  const DevTables &T = t->d;
  if (T.n_point_corr != T.n_nlist_sublat) return CMX_OK;  // local clexulator: no sweeps
  const int mo = T.max_occ, np = T.n_nlist_sublat;

  // ---- ECI-folded merged delta terms per point position
  typedef std::vector<std::pair<int, int>> Key;  // sorted (n, f)
  std::vector<int32_t> gt_beg(1, 0), gt_fbeg(1, 0), gt_f, gt_n;
  std::vector<double> gt_w;
  std::vector<std::map<Key, std::vector<double>>> merged(np);
  for (int p = 0; p < np; ++p) {
    int b = t->nlist_sublat[p];
    if (t->n_occ[b] > 1) P.mut_points.push_back(p);
    for (int q = 0; q < s->n_eci; ++q) {
      int fi = p * T.corr_size + (int)s->eci_idx[q];
      double e = s->eci_val[q];
      for (int g = t->delta_gbeg[fi]; g < t->delta_gbeg[fi + 1]; ++g) {
        int fd = t->group_dphi[g];
        if (fd < 0) return invalid("cmx_state_set_eci: delta function without a delta factor");
        double div = t->group_div[g] != 0.0 ? t->group_div[g] : 1.0;
        const double *ph = &t->phi[((size_t)b * T.n_func + fd) * mo];
        auto add = [&](const Key &key, double coef) {
          std::vector<double> &w = merged[p][key];
          if (w.empty()) w.assign((size_t)mo * mo, 0.0);
          for (int oi = 0; oi < mo; ++oi)
            for (int of = 0; of < mo; ++of)
              w[oi * mo + of] += e * coef * (ph[of] - ph[oi]) / div;
        };
        if (!t->group_has_sum[g]) {
          add(Key(), 1.0);
          continue;
        }
        for (int el = t->group_ebeg[g]; el < t->group_ebeg[g + 1]; ++el)
          for (int tm = t->elem_tbeg[el]; tm < t->elem_tbeg[el + 1]; ++tm) {
            Key key;
            for (int f = t->term_fbeg[tm]; f < t->term_fbeg[tm + 1]; ++f)
              key.push_back({t->factor_n[f], t->factor_f[f]});
            std::sort(key.begin(), key.end());
            add(key, t->term_coef[tm]);
          }
      }
    }
    for (auto const &kv : merged[p]) {
      bool nz = false;
      for (double w : kv.second) nz |= (w != 0.0);
      if (!nz) continue;
      for (auto const &nf : kv.first) {
        gt_n.push_back(nf.first);
        gt_f.push_back(nf.second);
      }
      gt_fbeg.push_back((int32_t)gt_n.size());
      gt_w.insert(gt_w.end(), kv.second.begin(), kv.second.end());
    }
    gt_beg.push_back((int32_t)gt_fbeg.size() - 1);
  }
  P.n_gterms = (int32_t)gt_fbeg.size() - 1;
  int rc;
  if ((rc = to_device(gt_beg, &P.d_gt_beg))) return rc;
  if ((rc = to_device(gt_fbeg, &P.d_gt_fbeg))) return rc;
  if ((rc = to_device(gt_f, &P.d_gt_f))) return rc;
  if ((rc = to_device(gt_n, &P.d_gt_n))) return rc;
  if ((rc = to_device(gt_w, &P.d_gt_w))) return rc;

  // ---- colouring: stride > interaction range along each axis.  Sites on
  // different sublattices of one cell interact through offset (0,0,0), so the
  // sublattice is part of the colour.
  int R[3] = {0, 0, 0};
  std::vector<char> active(T.nlist_len, 0);
  for (int n : gt_n) active[n] = 1;
  for (int n = 0; n < T.nlist_len; ++n)
    if (active[n])
      for (int a = 0; a < 3; ++a) R[a] = std::max(R[a], std::abs(t->nbr[4 * n + a]));
  int N[3] = {s->g.N0, s->g.N1, s->g.N2};
  for (int a = 0; a < 3; ++a) P.S[a] = smallest_divisor_above(N[a], R[a]);
  P.range_k = R[2];
  P.n_colours = P.S[0] * P.S[1] * P.S[2] * (int)P.mut_points.size();
  if (s->g.halo && s->g.halo < R[2])
    return invalid("cmx_state_set_eci: halo thinner than the interaction range along k");

  // work per attempted step (exact counts from the folded tables, averaged
  // over mutable points): distinct neighbor bytes + own byte + 1 write;
  // flops = products + adds (+ ~25 for exp/compare)
  {
    double bytes = 0, flops = 0;
    for (int p : P.mut_points) {
      std::vector<char> seen(T.nlist_len, 0);
      int nn = 0;
      for (int tt = gt_beg[p]; tt < gt_beg[p + 1]; ++tt) {
        int nf = gt_fbeg[tt + 1] - gt_fbeg[tt];
        flops += nf + 1;
        for (int f = gt_fbeg[tt]; f < gt_fbeg[tt + 1]; ++f)
          if (!seen[gt_n[f]]) {
            seen[gt_n[f]] = 1;
            ++nn;
          }
      }
      bytes += nn + 1 + 1;
      flops += 25;
    }
    int nm = std::max<int>(1, (int)P.mut_points.size());
    P.bytes_per_step = bytes / nm;
    P.flops_per_step = flops / nm;
  }
  P.valid = true;

  // ---- pair-LUT eligibility
  bool ok = (T.n_sublat == 1 && np == 1 && P.mut_points.size() == 1 &&
             t->n_occ[0] >= 2 && t->n_occ[0] <= 3 && T.nlist_len <= 64 &&
             s->g.N0 % 8 == 0 && s->g.N1 % 2 == 0 && s->g.N2 % 2 == 0 &&
             s->g.rep_stride < (int64_t)0xFFFFFFFFll &&
             P.S[0] == 2 && P.S[1] == 2 && P.S[2] == 2);
  std::map<int, std::vector<double>> V;  // neighbor -> V[on][oi][of]
  if (ok) {
    for (auto const &kv : merged[0]) {
      if (kv.first.size() > 1) {
        bool nz = false;
        for (double w : kv.second) nz |= (w != 0.0);
        if (nz) ok = false;
        continue;
      }
      if (kv.first.empty()) continue;
      int n = kv.first[0].first, f = kv.first[0].second;
      if (n == 0) {  // a self factor cannot appear in a delta function
        ok = false;
        continue;
      }
      std::vector<double> &v = V[n];
      if (v.empty()) v.assign((size_t)mo * mo * mo, 0.0);
      for (int on = 0; on < mo; ++on)
        for (int x = 0; x < mo * mo; ++x)
          v[(size_t)on * mo * mo + x] += kv.second[x] * t->phi[(size_t)f * mo + on];
    }
  }
  if (ok && !V.empty()) {
    // one class: all V tables equal (to rounding)
    auto const &v0 = V.begin()->second;
    double scale = 0;
    for (double x : v0) scale = std::max(scale, std::fabs(x));
    for (auto const &kv : V)
      for (size_t x = 0; x < v0.size(); ++x)
        if (std::fabs(kv.second[x] - v0[x]) > 1e-12 * scale) ok = false;
    if ((int)V.size() > 15) ok = false;  // nibble counters
    for (auto const &kv : V)
      for (int a = 0; a < 3; ++a)
        if (std::abs(t->nbr[4 * kv.first + a]) > 1) ok = false;
  }
  if (ok) {
    P.nocc = t->n_occ[0];
    P.z = (int)V.size();
    P.mask = 0;
    std::vector<int32_t> cls;
    for (auto const &kv : V) {
      const int32_t *o = &t->nbr[4 * kv.first];
      P.mask |= 1u << ((o[2] + 1) * 9 + (o[1] + 1) * 3 + (o[0] + 1));
      cls.push_back(kv.first);
    }
    P.n_lut = P.nocc * (P.nocc - 1) * 256;
    int32_t *d_cls = nullptr;
    if ((rc = to_device(cls, &d_cls))) return rc;
    CMX_CUDA(cudaMalloc((void **)&P.d_pair_dE, sizeof(double) * P.n_lut));
    k_build_pair_lut<<<(P.n_lut + 127) / 128, 128, 0, s->stream>>>(
        T, P.nocc, P.z, d_cls, s->n_eci, s->d_eci_idx, s->d_eci_val, P.d_pair_dE, P.n_lut);
    CMX_CUDA(cudaGetLastError());
    CMX_CUDA(cudaStreamSynchronize(s->stream));
    cudaFree(d_cls);
    CMX_CUDA(cudaMalloc((void **)&P.d_thr, sizeof(uint32_t) * P.n_lut * s->n_replicas));
    CMX_CUDA(cudaMalloc((void **)&P.d_thr_lo, sizeof(uint32_t) * P.n_lut * s->n_replicas));
    CMX_CUDA(cudaMalloc((void **)&P.d_dEpot, sizeof(double) * P.n_lut * s->n_replicas));
    P.pair_lut = true;
    // per step: z neighbor bytes + own byte read, 1 byte written; the FP64
    // work is the tabulated dE (2 site-function adds per neighbor, the ECI
    // dot, the exp folded into the threshold) -- report the table-free count
    P.bytes_per_step = P.z + 2;
  }
  P.thr_dirty = true;
  return CMX_OK;
}

static int ensure_partials(cmx_state *s, int blocks) {
  SweepPlan &P = s->plan;
  if (P.part_blocks == blocks && P.d_part_acc) return CMX_OK;
  cudaFree(P.d_part_acc);
  cudaFree(P.d_part_dE);
  P.d_part_acc = nullptr;
  P.d_part_dE = nullptr;
  size_t n = (size_t)blocks * s->n_replicas;
  CMX_CUDA(cudaMalloc((void **)&P.d_part_acc, sizeof(long long) * n));
  CMX_CUDA(cudaMalloc((void **)&P.d_part_dE, sizeof(double) * n));
  P.part_blocks = blocks;
  return CMX_OK;
}

// FCC nearest-neighbor shell in CASM's standard primitive cell
// (offsets of m_orbit_site_neighborhood[3], FCC default clexulator :404-414)
constexpr uint32_t mbit(int dx, int dy, int dz) {
  return 1u << ((dz + 1) * 9 + (dy + 1) * 3 + (dx + 1));
}
constexpr uint32_t kMaskFcc1NN =
    mbit(-1, 0, 0) | mbit(-1, 0, 1) | mbit(-1, 1, 0) | mbit(0, -1, 0) | mbit(0, -1, 1) |
    mbit(0, 0, -1) | mbit(0, 0, 1) | mbit(0, 1, -1) | mbit(0, 1, 0) | mbit(1, -1, 0) |
    mbit(1, 0, -1) | mbit(1, 0, 0);

template <int CX, int NOCC>
static void launch_pair(const PairSweepArgs &a, dim3 grid, cudaStream_t st, bool fcc) {
  if (fcc)
    k_sweep_pair_lut<CX, NOCC, kMaskFcc1NN><<<grid, 256, 0, st>>>(a);
  else
    k_sweep_pair_lut<CX, NOCC, 0u><<<grid, 256, 0, st>>>(a);
}

static int sweep_blocks_per_replica(uint32_t items, int n_replicas) {
  int want = (int)((items + 255) / 256);
  int cap = std::max(1, (148 * 8 + n_replicas - 1) / n_replicas);
  return std::max(1, std::min(want, cap));
}

// one pass over the colours whose k-colour equals kgroup (or all if < 0)
static int sweep_once(cmx_state *s, uint64_t seed, int64_t sweep, int kgroup,
                      int k_offset) {
  SweepPlan &P = s->plan;
  const DevTables &T = s->t->d;
  const Geom &g = s->g;
  size_t exs = (size_t)T.n_sublat * T.max_occ * T.max_occ;
  if (P.pair_lut) {
    if (P.thr_dirty) {
      dim3 grid((P.n_lut + 127) / 128, s->n_replicas);
      k_build_thresholds<<<grid, 128, 0, s->stream>>>(P.d_pair_dE, P.n_lut, P.nocc, T.max_occ, 0,
                                                      s->d_beta, s->d_exch, (int)exs, P.d_thr,
                                                      P.d_thr_lo, P.d_dEpot);
      CMX_CUDA(cudaGetLastError());
      P.thr_dirty = false;
    }
    PairSweepArgs a;
    a.occ = s->d_occ;
    a.g = g;
    a.mask = P.mask;
    uint32_t W = g.N0 / 8, J = g.N1 / 2, K = g.N2 / 2;
    a.divW = make_fastdiv(W);
    a.divJ = make_fastdiv(J);
    a.W = W;
    a.J = J;
    a.K = K;
    {
      uint32_t stride = (uint32_t)P.part_blocks * 256u;
      a.dc = stride % W;
      a.djj = (stride / W) % J;
      a.dkk = stride / (W * J);
    }
    a.thr_hi = P.d_thr;
    a.thr_lo = P.d_thr_lo;
    a.dEpot = P.d_dEpot;
    a.n_lut = P.n_lut;
    a.part_acc = P.d_part_acc;
    a.part_dE = P.d_part_dE;
    a.k0 = (uint32_t)seed;
    a.k1 = (uint32_t)(seed >> 32);
    a.sweep_lo = (uint32_t)sweep;
    a.k_offset = k_offset;
    dim3 grid(P.part_blocks, s->n_replicas);
    bool fcc = (P.mask == kMaskFcc1NN);
    for (int cz = 0; cz < 2; ++cz) {
      if (kgroup >= 0 && cz != kgroup) continue;
      for (int cy = 0; cy < 2; ++cy)
        for (int cx = 0; cx < 2; ++cx) {
          a.cy = cy;
          a.cz = cz;
          uint32_t colour = (cz * 2 + cy) * 2 + cx;
          a.ctr_hi = ((uint32_t)((uint64_t)sweep >> 32) << 16) | (colour << 8);
          if (P.nocc == 3) {
            if (cx == 0) launch_pair<0, 3>(a, grid, s->stream, fcc);
            else launch_pair<1, 3>(a, grid, s->stream, fcc);
          } else {
            if (cx == 0) launch_pair<0, 2>(a, grid, s->stream, fcc);
            else launch_pair<1, 2>(a, grid, s->stream, fcc);
          }
        }
    }
    CMX_CUDA(cudaGetLastError());
    return CMX_OK;
  }
  // generic
  GenericSweepArgs a;
  a.occ = s->d_occ;
  a.g = g;
  a.T = T;
  a.S0 = P.S[0];
  a.S1 = P.S[1];
  a.S2 = P.S[2];
  uint32_t n0 = g.N0 / P.S[0], n1 = g.N1 / P.S[1], n2 = g.N2 / P.S[2];
  a.div0 = make_fastdiv(n0);
  a.div1 = make_fastdiv(n1);
  a.items = n0 * n1 * n2;
  a.gt_beg = P.d_gt_beg;
  a.gt_fbeg = P.d_gt_fbeg;
  a.gt_f = P.d_gt_f;
  a.gt_n = P.d_gt_n;
  a.gt_w = P.d_gt_w;
  a.beta = s->d_beta;
  a.exch = s->d_exch;
  a.exch_stride = (int)exs;
  a.part_acc = P.d_part_acc;
  a.part_dE = P.d_part_dE;
  a.k0 = (uint32_t)seed;
  a.k1 = (uint32_t)(seed >> 32);
  a.sweep_lo = (uint32_t)sweep;
  a.k_offset = k_offset;
  dim3 grid(P.part_blocks, s->n_replicas);
  uint32_t colour = 0;
  for (int c2 = 0; c2 < P.S[2]; ++c2)
    for (int c1 = 0; c1 < P.S[1]; ++c1)
      for (int c0 = 0; c0 < P.S[0]; ++c0)
        for (int p : P.mut_points) {
          uint32_t col = colour++;
          if (kgroup >= 0 && c2 != kgroup) continue;
          a.c0 = c0;
          a.c1 = c1;
          a.c2 = c2;
          a.p = p;
          a.ctr_hi = ((uint32_t)((uint64_t)sweep >> 32) << 16) | (col & 0xffffu);
          k_sweep_generic<<<grid, 256, 0, s->stream>>>(a);
        }
  CMX_CUDA(cudaGetLastError());
  return CMX_OK;
}

static int sweep_prepare(cmx_state *s, const char *who) {
  if (!s) return invalid(std::string(who) + ": null state");
  if (!s->plan.valid) {
    cmx_set_error(std::string(who) + ": no sweep plan (bind ECI of a global clexulator first)");
    return CMX_ERR_STATE;
  }
  for (int r = 0; r < s->n_replicas; ++r)
    if (!(s->temperature[r] > 0.0)) {
      cmx_set_error(std::string(who) + ": conditions not set for every replica");
      return CMX_ERR_STATE;
    }
  CMX_CUDA(cudaSetDevice(s->t->device));
  SweepPlan &P = s->plan;
  uint32_t items;
  if (P.pair_lut)
    items = (uint32_t)(s->g.N0 / 8) * (s->g.N1 / 2) * (s->g.N2 / 2);
  else
    items = (uint32_t)(s->g.N0 / P.S[0]) * (s->g.N1 / P.S[1]) * (s->g.N2 / P.S[2]);
  int blocks = sweep_blocks_per_replica(items, s->n_replicas);
  if (P.part_blocks != blocks || !P.d_part_acc) {
    int rc = ensure_partials(s, blocks);
    if (rc) return rc;
    size_t n = (size_t)P.part_blocks * s->n_replicas;
    CMX_CUDA(cudaMemsetAsync(P.d_part_acc, 0, sizeof(long long) * n, s->stream));
    CMX_CUDA(cudaMemsetAsync(P.d_part_dE, 0, sizeof(double) * n, s->stream));
    P.attempts = 0;
  }
  return CMX_OK;
}

extern "C" int cmx_counters_reset(cmx_state *s) {
  int rc = sweep_prepare(s, "cmx_counters_reset");
  if (rc) return rc;
  SweepPlan &P = s->plan;
  size_t n = (size_t)P.part_blocks * s->n_replicas;
  CMX_CUDA(cudaMemsetAsync(P.d_part_acc, 0, sizeof(long long) * n, s->stream));
  CMX_CUDA(cudaMemsetAsync(P.d_part_dE, 0, sizeof(double) * n, s->stream));
  P.attempts = 0;
  return CMX_OK;
}

extern "C" int cmx_counters_read(cmx_state *s, cmx_counters *counters) {
  int rc = sweep_prepare(s, "cmx_counters_read");
  if (rc) return rc;
  if (!counters) return invalid("cmx_counters_read: null output");
  SweepPlan &P = s->plan;
  k_reduce_counters<<<s->n_replicas, 32, 0, s->stream>>>(P.d_part_acc, P.d_part_dE, P.part_blocks,
                                                        P.attempts, s->d_counters);
  CMX_CUDA(cudaGetLastError());
  CMX_CUDA(cudaMemcpyAsync(counters, s->d_counters, sizeof(cmx_counters) * s->n_replicas,
                           cudaMemcpyDeviceToHost, s->stream));
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  return CMX_OK;
}

extern "C" int cmx_sgc_sweep(cmx_state *s, int64_t n_sweeps, uint64_t seed,
                             int64_t first_sweep, cmx_counters *counters) {
  int rc = cmx_counters_reset(s);
  if (rc) return rc;
  if (n_sweeps < 0) return invalid("cmx_sgc_sweep: n_sweeps < 0");
  if (s->g.halo) return invalid("cmx_sgc_sweep: slab states are driven by cmx_sgc_sweep_kgroup");
  for (int64_t w = 0; w < n_sweeps; ++w) {
    rc = sweep_once(s, seed, first_sweep + w, -1, 0);
    if (rc) return rc;
  }
  s->plan.attempts += (long long)s->g.n_cells * (long long)s->plan.mut_points.size() * n_sweeps;
  if (counters) return cmx_counters_read(s, counters);
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  return CMX_OK;
}

extern "C" int cmx_sgc_sweep_kgroup(cmx_state *s, uint64_t seed, int64_t sweep,
                                    int32_t kgroup) {
  int rc = sweep_prepare(s, "cmx_sgc_sweep_kgroup");
  if (rc) return rc;
  if (kgroup < -1 || kgroup >= s->plan.S[2]) return invalid("cmx_sgc_sweep_kgroup: bad kgroup");
  rc = sweep_once(s, seed, sweep, kgroup, s->k_offset);
  if (rc) return rc;
  long long per = (long long)s->g.n_cells * (long long)s->plan.mut_points.size();
  if (kgroup >= 0) per /= s->plan.S[2];
  s->plan.attempts += per;
  return CMX_OK;  // asynchronous: enqueued on the state's stream
}

extern "C" int cmx_sweep_info(const cmx_state *s, char *name, size_t name_cap,
                              double *bytes_per_step, double *flops_per_step,
                              int32_t *n_colours, int32_t *colour_strides,
                              int32_t *range_k) {
  if (!s) return invalid("cmx_sweep_info: null state");
  if (!s->plan.valid) {
    cmx_set_error("cmx_sweep_info: no sweep plan");
    return CMX_ERR_STATE;
  }
  const char *nm = s->plan.pair_lut ? "pair_lut" : "generic";
  if (name && name_cap) {
    std::strncpy(name, nm, name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (bytes_per_step) *bytes_per_step = s->plan.bytes_per_step;
  if (flops_per_step) *flops_per_step = s->plan.flops_per_step;
  if (n_colours) *n_colours = s->plan.n_colours;
  if (colour_strides)
    for (int a = 0; a < 3; ++a) colour_strides[a] = s->plan.S[a];
  if (range_k) *range_k = s->plan.range_k;
  return CMX_OK;
}
