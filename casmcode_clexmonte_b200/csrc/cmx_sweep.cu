// Semi-grand canonical sublattice-checkerboard sweeps.
//
// Replaces the sequential loop of methods/occupation_metropolis.hh:92-120 +
// the semi-grand proposal (SemiGrandCanonicalCalculator.cc:104-120) + the
// potential delta (:186-213) by colour-by-colour simultaneous updates of
// non-interacting site sets.  Evaluators:
//
//  * "pair_lut": the bound ECI select only point + pair functions on a single
//    sublattice with <= 3 occupants whose active neighbors all lie in
//    {-1,0,1}^3 and form one symmetry class (FCC/BCC/SC nearest-neighbor
//    models).  dE then depends only on (occ_i, occ_f, neighbor species
//    counts); it is tabulated ONCE by running the faithful evaluator on a
//    representative neighborhood per count combination, and the Metropolis
//    test  u < exp(-beta dE)  becomes an integer compare against a
//    per-replica threshold table staged in shared memory.  Sixteen sites per
//    thread (both x colours of a 16-byte row chunk), neighbor species counted
//    for all 16 byte lanes at once (kernel k_sweep_pair16).
//  * "pair_lut2": the same with TWO neighbor classes (FCC 1NN + 2NN: the reference's dense
//    ECI) on x4-interleaved rows: a second byte-lane sum and a compact two-class table in
//    the colour-pass kernel (cmx_sweep_stream.cuh).
//  * "pair_sum": other point + pair bases: per-neighbor tables, one site per thread.
//  * "generic": any table (multi-sublattice, triplets, quadruplets): ECI-folded
//    merged term lists, one site per thread (wide orbit sets: one site per warp), FP64
//    products, exp().
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include <cooperative_groups.h>

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
  cudaFree(p.d_gt_vi);
  cudaFree(p.d_gt_pk);
  cudaFree(p.d_act_beg);
  cudaFree(p.d_act_n);
  cudaFree(p.d_ps_V);
  cudaFree(p.d_ps_P0);
  cudaFree(p.d_pair_dE);
  cudaFree(p.d_tab);
  cudaFree(p.d_thr_lo);
  cudaFree(p.d_dEpot);
  cudaFree(p.d_tab24);
  cudaFree(p.d_pair2_dE);
  cudaFree(p.d_maps2);
  cudaFree(p.d_tab2);
  cudaFree(p.d_thr_lo2);
  cudaFree(p.d_dEpot2);
  for (auto &kv : p.stream_lists) cudaFree(kv.second.d_units);
  cudaFree(p.d_ticket);
  cudaFree(p.d_gridbar);
  cudaFree(p.d_part_acc);
  cudaFree(p.d_part_dE);
  cudaFree(p.d_e_lut);
  cudaFree(p.d_e_lin);
  cudaFree(p.d_corr_func_mask);
  cudaFree(p.d_corr_lin);
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

// ---------------------------------------------------------------------------
// Acceptance tables of the pair16 kernel.
//
// The uniform of one attempted step is a 47-bit integer u47 = (u15 << 32) | w32:
// u15 comes from the 16-bit field every site draws (bit 0 of the field is the
// proposal bit `alt`), w32 is drawn lazily only when the first 15 bits tie.
//   accept  <=>  u47 < t47,   t47 = 2^47                        if dE < 0
//                              t47 = ceil(exp(-dE*beta) * 2^47)  otherwise
// (metropolis_acceptance [EXT]: `if (dE < 0) return true; return rng < exp(-dE*beta)`).
// Table index  idx = cnt | sa << 8:  cnt = n1 + CMX_VA_CODE*n2 is the byte-lane
// sum of the neighbors' storage codes (n1, n2 = neighbor counts of occupants 1
// and 2), sa = self occupant | alt << 2.  One 4-byte entry
//   e = t << 8 | code,  t = (t47 >> 32) << 1 | 1,  code = storage code proposed
// is compared with (field | 1) << 8 | 0xFF, so one unsigned compare decides
// u15 < t47 >> 32 and a tie is "the upper 24 bits are equal".  4-byte entries
// and an even Va code that is not a multiple of 16 spread the physically
// frequent (n1, n2) over the 32 shared-memory banks.
// ---------------------------------------------------------------------------
#define CMX_TAB16(NOCC) ((NOCC) == 3 ? 2048 : 512)
#define CMX_VA_CODE 18  // must match cmx_enc()

__global__ void k_build_tab16(const double *__restrict__ lut, int nocc, int z, int max_occ,
                              const double *__restrict__ beta,
                              const double *__restrict__ exch, int exch_stride,
                              int n_tab, uint32_t *__restrict__ tab,
                              uint32_t *__restrict__ thr_lo,
                              double *__restrict__ dEpot) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int r = blockIdx.y;
  if (idx >= n_tab) return;
  const int cnt = idx & 255, sa = idx >> 8;
  const int oi = sa & 3, alt = sa >> 2;
  const int n2 = (nocc == 3) ? cnt / CMX_VA_CODE : 0;
  const int n1 = cnt - n2 * CMX_VA_CODE;
  size_t o = (size_t)r * n_tab + idx;
  if (oi >= nocc || alt >= nocc - 1 || n1 + n2 > z) {
    tab[o] = 1u << 8;  // never accepted ((field | 1) >= 1); a tie finds thr_lo 0
    thr_lo[o] = 0;
    dEpot[o] = 0.0;
    return;
  }
  int of = oi + 1 + alt;
  if (of >= nocc) of -= nocc;
  const int pair = oi * (nocc - 1) + alt;
  double x = exch[(size_t)r * exch_stride + (0 * max_occ + oi) * max_occ + of];
  double dE = __dsub_rn(lut[(pair << 8) | n1 | (n2 << 4)], x);
  unsigned long long t;
  const double two47 = 140737488355328.0;
  if (dE < 0.0) {
    t = 1ull << 47;
  } else {
    double p = exp(-dE * beta[r]);
    t = (unsigned long long)ceil(p * two47);
  }
  const uint32_t t17 = (uint32_t)(t >> 32) << 1 | 1u;
  tab[o] = (t17 << 8) | (uint32_t)((of == 2) ? CMX_VA_CODE : of);
  thr_lo[o] = (uint32_t)(t & 0xFFFFFFFFull);
  dEpot[o] = dE;
}

// Acceptance tables of the two-class kernel (compact index, SweepPlan::pair2): the same
// threshold arithmetic, entry = thr16 | proposed code << 16 (the colour-pass kernel's format).
__global__ void k_build_tab2(const double *__restrict__ lut2, int n_tab2, int per_row, int nocc, int max_occ,
                             const double *__restrict__ beta, const double *__restrict__ exch, int exch_stride,
                             uint32_t *__restrict__ tab2, uint32_t *__restrict__ thr_lo2, double *__restrict__ dEpot2) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (idx >= n_tab2) return;
  const int row = idx / per_row;
  const int oi = row / (nocc - 1), alt = row - oi * (nocc - 1);
  int of = oi + 1 + alt;
  if (of >= nocc) of -= nocc;
  const double x = exch[(size_t)r * exch_stride + (0 * max_occ + oi) * max_occ + of];
  const double dE = __dsub_rn(lut2[idx], x);
  unsigned long long t;
  const double two47 = 140737488355328.0;
  if (dE < 0.0) {
    t = 1ull << 47;
  } else {
    const double p = exp(-dE * beta[r]);
    t = (unsigned long long)ceil(p * two47);
  }
  const size_t o = (size_t)r * n_tab2 + idx;
  tab2[o] = (uint32_t)(t >> 32) | ((uint32_t)((of == 2) ? CMX_VA_CODE : of) << 16);
  thr_lo2[o] = (uint32_t)(t & 0xFFFFFFFFull);
  dEpot2[o] = dE;
}

// ---------------------------------------------------------------------------
// pair16 sweep kernel
// ---------------------------------------------------------------------------

struct Pair16Args {
  int8_t *occ;  // replica 0 base (start of the low ghost layers)
  Geom g;
  int cy, cz;        // row colour (parities of j and k); both x colours are fused
  uint32_t mask;     // runtime neighbor mask, bit (dz+1)*9 + (dy+1)*3 + (dx+1)
  uint32_t W;        // 16-byte chunks per row
  uint32_t RB;       // whole rows per block and iteration (RB * W <= 256)
  FastDiv divW, divJ;
  uint32_t J;        // rows of this colour per layer (N1 / 2)
  uint32_t n_rows;   // J * (N2 / 2)
  const uint32_t *tab;     // [replica][n_tab]
  const uint32_t *thr_lo;  // [replica][n_tab]
  const double *dEpot;     // [replica][n_tab]
  long long *part_acc;     // [replica][part_stride]
  uint32_t part_stride;
  double *part_dE;
  uint32_t k0, k1;       // seed
  PhiloxKeys rk;         // its round-key schedule (read from the constant bank)
  uint32_t sweep_lo;     // RNG counter words
  uint32_t ctr_hi;       // (sweep_hi << 16) | (row colour << 9); bit 8 = x colour, low bits = draw
  int k_offset;          // global k of local layer 0 (slab decomposition)
};

// shared memory is addressed with explicit 32-bit addresses: the dE table sits at
// a compile-time offset from the acceptance table, so one address serves both
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// (x, 0xFF) byte permute with the selector as an immediate: `ff` is kept in a
// register (nvcc would otherwise put the selector in a register and move it
// from a uniform register before every use)
template <uint32_t SEL>
__device__ __forceinline__ uint32_t prmt_imm(uint32_t x, uint32_t ff) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(ff), "n"(SEL));
  return d;
}
template <int OFF>
__device__ __forceinline__ double lds_f64_at(uint32_t addr) {
  double v;
  asm("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
  return v;
}

// One x colour of one 16-site chunk.
//  cnt[i]  byte-lane neighbor sums of the chunk (only the lanes of colour CX are used)
//  C[i]    the chunk (storage codes); accepted sites are replaced in place
//  tab     shared-memory address of the acceptance table, dE table NTAB*4 bytes later
//          (entry idx at tab + 4*idx resp. tab + NTAB*4 + 8*idx)
template <int CX, int NOCC, bool ACCUM>
__device__ __forceinline__ void pair16_update(const uint32_t (&cnt)[4], uint32_t (&C)[4],
                                              const Philox &ph, uint32_t tab, uint32_t ff,
                                              uint32_t &n_acc, double &e_sum, bool &tie) {
  constexpr int NTAB = CMX_TAB16(NOCC);
  constexpr uint32_t lanes = CX ? 0x03000300u : 0x00030003u;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t R = ph.c[i];
    // occupant | alt << 2 in the target lanes, zero elsewhere (code & 3 = occupant)
    const uint32_t altw = (NOCC == 3) ? ((R >> (CX ? 5 : 13)) & (CX ? 0x04000400u : 0x00040004u)) : 0u;
    const uint32_t SA = (C[i] & lanes) | altw;
    // (u15 << 1 | 1) in both halves: bit 15 (alt) of the low field lands on the forced bit 16
    const uint32_t Rw = (R << 1) | 0x00010001u;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int b = 2 * h + CX;  // byte of the word
      // idx = cnt byte b | SA byte b << 8  (bytes 2,3 <- a zero byte of SA)
      const uint32_t sel = (uint32_t)b | ((uint32_t)(4 + b) << 4) | ((uint32_t)(4 + (b ^ 1)) << 8) |
                           ((uint32_t)(4 + (b ^ 1)) << 12);
      const uint32_t idx = __byte_perm(cnt[i], SA, sel);
      const uint32_t e = lds_u32(tab + 4u * idx);
      // (field | 1) << 8 | 0xFF
      const uint32_t f1s = h ? prmt_imm<0x5324u>(Rw, ff) : prmt_imm<0x5104u>(Rw, ff);
      const bool ok = f1s < e;
      tie |= ((f1s ^ e) < 256u);
      if (ok) {
        // byte b of C[i] <- proposed code (byte 0 of e)
        const uint32_t ins = (b == 0) ? 0x3214u : (b == 1) ? 0x3240u : (b == 2) ? 0x3410u : 0x4210u;
        C[i] = __byte_perm(C[i], e, ins);
        n_acc += 1;
        if (ACCUM) e_sum += lds_f64_at<NTAB * 4>(tab + 4u * idx + 4u * idx);
      }
    }
  }
}

// rare path (probability 2^-15 per site): the first 15 bits of the uniform equal
// the threshold's; draw 32 more bits per site and finish the 47-bit comparison.
// Re-derives everything from the chunk as it was BEFORE this colour's update.
template <int CX, int NOCC, bool ACCUM>
__device__ __forceinline__ void pair16_ties(const uint32_t (&cnt)[4], const uint4 *chunk0,
                                            uint32_t (&C)[4], const Philox &ph, uint32_t tab,
                                            const uint32_t *__restrict__ thr_lo, uint32_t gid,
                                            uint32_t r, uint32_t sweep_lo, uint32_t ctr, uint32_t k0,
                                            uint32_t k1, uint32_t &n_acc, double &e_sum) {
  constexpr int NTAB = CMX_TAB16(NOCC);
  const Philox lo0 = philox_sweep(gid, r, sweep_lo, ctr | 1u, k0, k1);
  const Philox lo1 = philox_sweep(gid, r, sweep_lo, ctr | 2u, k0, k1);
  // the chunk as stored in global memory: it is written back only after both
  // colours, and the occupants of THIS colour's lanes have not changed before
  const uint4 c0 = *chunk0;
  const uint32_t C0[4] = {c0.x, c0.y, c0.z, c0.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t R = ph.c[i];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int b = 2 * h + CX;
      const uint32_t field = h ? (R >> 16) : (R & 0xFFFFu);
      const uint32_t sa = ((C0[i] >> (8 * b)) & 3u) | ((NOCC == 3) ? ((field >> 15) << 2) : 0u);
      const uint32_t idx = ((cnt[i] >> (8 * b)) & 0xFFu) | (sa << 8);
      const uint32_t e = lds_u32(tab + 4u * idx);
      if ((((field & 0x7FFFu) << 1) | 1u) != (e >> 8)) continue;
      const int q = 2 * i + h;  // target site of the chunk, 0..7
      const uint32_t w32 = (q < 4) ? lo0.c[q & 3] : lo1.c[q & 3];
      if (w32 < thr_lo[idx]) {
        C[i] = (C[i] & ~(0xFFu << (8 * b))) | ((e & 0xFFu) << (8 * b));
        n_acc += 1;
        if (ACCUM) e_sum += lds_f64_at<NTAB * 4>(tab + 8u * idx);
      }
    }
  }
}

// One thread owns one 16-byte chunk (16 consecutive sites) of a row whose (j,k)
// parities are (cy,cz) and updates BOTH x colours of it:
//  * the <= 8 neighbor rows arrive as 16-byte loads (+ 4-byte side words for the
//    bytes just outside the chunk).  Because occupants are stored as 0/1/16, the
//    plain word-wise sum of the rows is, lane by lane, n1 + 16*n2: the rows are
//    summed per x-offset class first (A0/Am/Ap), the dx = -1/+1 classes are then
//    shifted by one byte lane with funnel shifts -- 16 sites at once;
//  * even lanes are updated first, the changed first byte goes to the left
//    neighbor chunk through shared memory (a block always owns whole rows), then
//    the odd lanes are updated against the new even lanes;
//  * per x colour one Philox4x32-10 call yields the eight 16-bit fields
//    (1 proposal bit + 15 uniform bits); the Metropolis test is one integer
//    compare against the shared-memory table, 32 more bits are drawn only on a tie;
//  * accepted dE (tabulated) is summed in FP64 per thread, reduced per block.
template <int NOCC, uint32_t MASK_CT, bool ACCUM, int MINB>
__global__ void __launch_bounds__(256, MINB) k_sweep_pair16(Pair16Args a) {
  constexpr int NTAB = CMX_TAB16(NOCC);
  // [acceptance table: NTAB x 4 B][dE table: NTAB x 8 B]
  __shared__ __align__(16) unsigned char sh_tables[NTAB * 4 + (ACCUM ? NTAB * 8 : 0)];
  __shared__ uint8_t sh_x[2][256];
  __shared__ long long sh_acc[8];
  __shared__ double sh_sum[8];
  const uint32_t r = blockIdx.y;
  {
    const uint32_t *gt = a.tab + (size_t)r * NTAB;
    const double *ge = a.dEpot + (size_t)r * NTAB;
    uint32_t *st = reinterpret_cast<uint32_t *>(sh_tables);
    double *se = reinterpret_cast<double *>(sh_tables + NTAB * 4);
    for (int q = threadIdx.x; q < NTAB; q += blockDim.x) {
      st[q] = gt[q];
      if (ACCUM) se[q] = ge[q];
    }
  }
  const uint32_t tab = (uint32_t)__cvta_generic_to_shared(sh_tables);
  uint32_t ff;
  asm volatile("mov.u32 %0, 0xFF;" : "=r"(ff));  // opaque to constant propagation, see prmt_imm
  __syncthreads();
  const uint32_t mask = MASK_CT ? MASK_CT : a.mask;
  const Geom &g = a.g;
  int8_t *base = a.occ + (size_t)r * g.rep_stride;  // includes the ghost layers
  // (no asm pinning of base: keep the global address space visible to the compiler)
  const uint32_t N0 = g.N0, N1 = g.N1, N2 = g.N2;
  const uint32_t layer = N0 * N1;
  const bool halo = g.halo != 0;
  uint32_t n_acc = 0;  // < 2^32 accepted sites per thread and launch
  double e_sum = 0.0;

  uint32_t rl, c;
  fastdivmod(threadIdx.x, a.divW, rl, c);
  const bool lane_on = rl < a.RB;
  const uint32_t nxt = (c == a.W - 1) ? threadIdx.x - (a.W - 1) : threadIdx.x + 1;
  const uint32_t x0 = 16 * c;
  const uint32_t dl = (x0 == 0) ? N0 - 4 : 0u - 4u;         // word holding byte x0-1
  const uint32_t dr = (x0 + 16 == N0) ? 16u - N0 : 16u;     // word holding byte x0+16
  const uint32_t mc = (mask >> 12) & 7u;                     // center row: dx = -1 / +1 bits
  uint32_t it = 0;
  for (uint32_t row0 = blockIdx.x * a.RB; row0 < a.n_rows; row0 += gridDim.x * a.RB, ++it) {
    const uint32_t row = row0 + rl;
    const bool on = lane_on && row < a.n_rows;
    uint32_t C[4] = {0, 0, 0, 0}, T[4] = {0, 0, 0, 0};
    uint32_t cl = 0, off_c = 0, gid = 0;
    if (on) {
      uint32_t kk, jj;
      fastdivmod(row, a.divJ, kk, jj);
      const uint32_t j = 2 * jj + a.cy, k = 2 * kk + a.cz;
      off_c = ((k + g.halo) * N1 + j) * N0 + x0;
      gid = ((k + (uint32_t)a.k_offset) * N1 + j) * a.W + c;
      uint32_t dj[3], dk[3];
      dj[0] = (j == 0) ? (N1 - 1) * N0 : 0u - N0;
      dj[1] = 0;
      dj[2] = (j == N1 - 1) ? 0u - (N1 - 1) * N0 : N0;
      dk[0] = (!halo && k == 0) ? (N2 - 1) * layer : 0u - layer;
      dk[1] = 0;
      dk[2] = (!halo && k == N2 - 1) ? 0u - (N2 - 1) * layer : layer;
      uint32_t A0[4] = {0, 0, 0, 0}, Am[4] = {0, 0, 0, 0}, Ap[4] = {0, 0, 0, 0};
      uint32_t sm = 0, sp = 0;
#pragma unroll
      for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          const uint32_t m3 = (mask >> ((dz + 1) * 9 + (dy + 1) * 3)) & 7u;
          const bool center = (dz == 0 && dy == 0);
          if (m3 == 0 && !center) continue;
          const uint32_t off = off_c + dk[dz + 1] + dj[dy + 1];
          const uint4 ch = *reinterpret_cast<const uint4 *>((base + off));
          if (center) {
            C[0] = ch.x;
            C[1] = ch.y;
            C[2] = ch.z;
            C[3] = ch.w;
            if (m3 & 1u) cl = *reinterpret_cast<const uint32_t *>((base + (off + dl)));
            continue;
          }
          if (m3 & 2u) {
            A0[0] += ch.x;
            A0[1] += ch.y;
            A0[2] += ch.z;
            A0[3] += ch.w;
          }
          if (m3 & 1u) {
            Am[0] += ch.x;
            Am[1] += ch.y;
            Am[2] += ch.z;
            Am[3] += ch.w;
            sm += *reinterpret_cast<const uint32_t *>((base + (off + dl)));
          }
          if (m3 & 4u) {
            Ap[0] += ch.x;
            Ap[1] += ch.y;
            Ap[2] += ch.z;
            Ap[3] += ch.w;
            sp += *reinterpret_cast<const uint32_t *>((base + (off + dr)));
          }
        }
      }
      // T[x] = A0[x] + Am[x-1] + Ap[x+1], byte lanes of the 16-byte chunk
      T[0] = A0[0] + __funnelshift_l(sm, Am[0], 8) + __funnelshift_r(Ap[0], Ap[1], 8);
      T[1] = A0[1] + __funnelshift_l(Am[0], Am[1], 8) + __funnelshift_r(Ap[1], Ap[2], 8);
      T[2] = A0[2] + __funnelshift_l(Am[1], Am[2], 8) + __funnelshift_r(Ap[2], Ap[3], 8);
      T[3] = A0[3] + __funnelshift_l(Am[2], Am[3], 8) + __funnelshift_r(Ap[3], sp, 8);
      // ---- x colour 0: even lanes; same-row neighbors are odd lanes (old values)
      uint32_t cnt[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        cnt[i] = T[i];
        if (mc & 1u) cnt[i] += __funnelshift_l(i ? C[i - 1] : cl, C[i], 8);
        if (mc & 4u) cnt[i] += __funnelshift_r(C[i], (i < 3) ? C[i + 1] : 0u, 8);
      }
      const uint32_t ctr0 = a.ctr_hi;
      const Philox ph = philox_sweep_rk(gid, r, a.sweep_lo, ctr0, a.rk);
      bool tie = false;
      pair16_update<0, NOCC, ACCUM>(cnt, C, ph, tab, ff, n_acc, e_sum, tie);
      if (tie)
        pair16_ties<0, NOCC, ACCUM>(cnt, reinterpret_cast<const uint4 *>(base + off_c), C, ph, tab,
                                    a.thr_lo + (size_t)r * NTAB, gid, r, a.sweep_lo, ctr0, a.k0, a.k1,
                                    n_acc, e_sum);
      sh_x[it & 1][threadIdx.x] = (uint8_t)(C[0] & 0xFFu);
    }
    __syncthreads();
    if (on) {
      // ---- x colour 1: odd lanes against the updated even lanes; byte 16 is the
      // (updated) first byte of the next chunk of the row
      const uint32_t nb = sh_x[it & 1][nxt];
      uint32_t cnt[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        cnt[i] = T[i];
        if (mc & 1u) cnt[i] += __funnelshift_l(i ? C[i - 1] : 0u, C[i], 8);
        if (mc & 4u) cnt[i] += __funnelshift_r(C[i], (i < 3) ? C[i + 1] : nb, 8);
      }
      const uint32_t ctr1 = a.ctr_hi | 0x100u;
      const Philox ph = philox_sweep_rk(gid, r, a.sweep_lo, ctr1, a.rk);
      bool tie = false;
      pair16_update<1, NOCC, ACCUM>(cnt, C, ph, tab, ff, n_acc, e_sum, tie);
      if (tie)
        pair16_ties<1, NOCC, ACCUM>(cnt, reinterpret_cast<const uint4 *>(base + off_c), C, ph, tab,
                                    a.thr_lo + (size_t)r * NTAB, gid, r, a.sweep_lo, ctr1, a.k0, a.k1,
                                    n_acc, e_sum);
      const uint4 out = make_uint4(C[0], C[1], C[2], C[3]);
      *reinterpret_cast<uint4 *>(base + off_c) = out;
    }
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
    size_t slot = (size_t)r * a.part_stride + blockIdx.x;
    a.part_acc[slot] += A;
    a.part_dE[slot] += E;
  }
}

#include "cmx_sweep_stream.cuh"

// ---------------------------------------------------------------------------
// generic sweep kernel: one site per thread
// ---------------------------------------------------------------------------
// Which random bits the pair-LUT kernels give site i of a row (rng16 mode: the generic
// evaluator draws the same bits).  One Philox call per (16-byte chunk, x colour) yields
// four words = eight 16-bit fields [alt:1 | u15:15]; the tie-break words come from the
// calls with counter | 1 (targets 0..3 of the colour) and | 2 (targets 4..7).
//   linear rows (k_sweep_pair16):       chunk = i / 16, x = i % 16, target q = x / 2,
//                                       field = half (x / 2) % 2 of word x / 4
//   x4-interleaved rows (stream16):     word w = i % Q, byte b = i / Q, chunk = w / 4,
//                                       h = (w % 4) / 2, target q = 4 h + b,
//                                       field = half b % 2 of word 2 h + b / 2
// (the x colour, i % 2 in both layouts, is part of the counter)
struct Rng16Site {
  uint32_t chunk, word, half, q;
};
__device__ __forceinline__ Rng16Site cmx_rng16_site(const Geom &g, int i) {
  Rng16Site s;
  if (g.xq_log) {
    const uint32_t w = (uint32_t)i & ((1u << g.xq_log) - 1u), b = (uint32_t)i >> g.xq_log;
    const uint32_t h = (w & 3u) >> 1;
    s.chunk = w >> 2;
    s.word = 2u * h + (b >> 1);
    s.half = b & 1u;
    s.q = 4u * h + b;
  } else {
    const uint32_t x = (uint32_t)i & 15u;
    s.chunk = (uint32_t)i >> 4;
    s.word = x >> 2;
    s.half = (x >> 1) & 1u;
    s.q = x >> 1;
  }
  return s;
}

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
  // rng16 != 0: draw exactly the random bits the pair16 kernel draws for this
  // site (same counters, same 47-bit comparison), so that the two evaluators
  // produce the same trajectory and can be compared bit for bit.
  int rng16;
  int accum;  // CMX_SWEEP_DE_SUM
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
    const int oi = cmx_dec(occ[off]);
    int alt;
    uint32_t u_hi, u_lo;  // rng16: 15 + 32 bit uniform; else 21 + 32 bit
    if (a.rng16) {
      const Rng16Site rs = cmx_rng16_site(g, i);
      const uint32_t gid = (uint32_t)(((uint32_t)(k + a.k_offset) * g.N1 + j) * (g.N0 >> 4) + rs.chunk);
      const uint32_t q = rs.q;
      const uint32_t ctr = a.ctr_hi;
      const Philox ph = philox_sweep(gid, (uint32_t)r, a.sweep_lo, ctr, a.k0, a.k1);
      const uint32_t R = ph.c[rs.word];
      const uint32_t field = rs.half ? (R >> 16) : (R & 0xFFFFu);
      alt = (nocc == 3) ? (int)(field >> 15) : 0;
      u_hi = field & 0x7FFFu;
      const Philox lo = philox_sweep(gid, (uint32_t)r, a.sweep_lo, ctr | ((q < 4) ? 1u : 2u), a.k0, a.k1);
      u_lo = lo.c[q & 3];
    } else {
      // global site id -> RNG counter
      const uint32_t gid_lo = (uint32_t)(((uint64_t)(k + a.k_offset) * g.N1 + j) * g.N0 + i);
      const Philox ph = philox4x32_10(gid_lo, (uint32_t)r | ((uint32_t)b << 24), a.sweep_lo,
                                      a.ctr_hi, a.k0, a.k1);
      alt = (int)__umulhi(ph.c[2], (uint32_t)(nocc - 1));
      u_hi = ph.c[1] & 0x1FFFFFu;
      u_lo = ph.c[0];
    }
    int of = oi + 1 + alt;
    if (of >= nocc) of -= nocc;
    double dE = 0.0;
    for (int t = tb; t < te; ++t) {
      double v = a.gt_w[((size_t)t * mo + oi) * mo + of];
      for (int q = a.gt_fbeg[t]; q < a.gt_fbeg[t + 1]; ++q) {
        const int n = a.gt_n[q];
        const int64_t no = cmx_nbr_offset(T, g, n, i, j, k, nullptr);
        const int o = cmx_dec(occ[no]);
        v *= T.phi[((size_t)T.nbr[n].w * T.n_func + a.gt_f[q]) * mo + o];
      }
      dE += v;
    }
    dE -= exch[oi * mo + of];
    bool accept = dE < 0.0;
    if (!accept) {
      const unsigned long long u = ((unsigned long long)u_hi << 32) | u_lo;
      if (a.rng16) {
        // the pair16 kernel's integer test: u47 < ceil(exp(-dE beta) 2^47)
        accept = u < (unsigned long long)ceil(exp(-dE * beta) * 140737488355328.0);
      } else {
        accept = (double)u * (1.0 / 9007199254740992.0) < exp(-dE * beta);
      }
    }
    if (accept) {
      occ[off] = (int8_t)cmx_enc(g, of);
      ++n_acc;
      if (a.accum) e_sum += dE;
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

// ---------------------------------------------------------------------------
// pair-sum sweep kernel: one site per thread, dE from per-neighbor tables
// ---------------------------------------------------------------------------
// Bases of point and pair functions only (every folded term has at most one factor):
//   dE(oi -> of) = P0[oi][of] + sum over the active neighbors s of V[s][oi][of][occupant_s]
// with V folded from the same term lists the generic evaluator walks (cmx_plan_sweep).  The
// tables of the colour's point position sit in shared memory; a site whose neighborhood does
// not cross a periodic seam (all but a thin shell of the box) reaches its neighbors by adding
// precomputed byte offsets.  Same random bits and the same decision rule as k_sweep_generic;
// dE differs from it in the last bits only (summation order), checked per proposal by
// cmx_sweep_debug_delta_e.  This is the evaluator of the reference's dense FCC ECI (points +
// 1NN + 2NN pairs, the 19-site neighbourhood of SURVEY 8d), of multi-sublattice pair models
// and of pair models on boxes the pair-LUT kernels do not take.
struct PairSumArgs {
  const double *V;    // [act slot][oi][of][on]
  const double *P0;   // [point position][oi][of]
  const int32_t *act_beg, *act_n;
  int32_t R[3];       // interaction range per axis
  int32_t fast;       // byte offsets of a replica fit 32 bits
};

__device__ __forceinline__ double pairsum_delta(const Geom &g, const int8_t *occ, const double *shV, const int4 *shN,
                                                const int32_t *shD, int n_act, int mo, const int32_t (&R)[3], int fast,
                                                double p0, int i, int j, int k, int64_t off, int oi, int of) {
  const int mo3 = mo * mo * mo;
  const double *Vrow = shV + (oi * mo + of) * mo;
  double dE = p0;
  const bool skewed = (g.s10 | g.s20 | g.s21) != 0;
  const bool in_jk = fast && j >= R[1] && j < g.N1 - R[1] && (g.halo || (k >= R[2] && k < g.N2 - R[2]));
  if (in_jk && R[0] <= 1 && (!skewed || (i >= 1 && i < g.N0 - 1))) {
    // Range 1 along i (every first-shell model): the x neighbors' byte offsets are computed per
    // site -- with x4-interleaved rows EVERY warp holds a site next to a word-lane seam, and
    // sending that one lane through the wrapped path would cost the warp the whole path
    const int im = (i == 0) ? g.N0 - 1 : i - 1, ip = (i == g.N0 - 1) ? 0 : i + 1;
    const int x0 = cmx_xpos(g, i);
    const int xm = cmx_xpos(g, im) - x0, xp = cmx_xpos(g, ip) - x0;
#pragma unroll 6
    for (int s = 0; s < n_act; ++s) {
      const int dx = shN[s].x;
      const int32_t d = shD[s] + (dx < 0 ? xm : (dx > 0 ? xp : 0));
      dE += Vrow[s * mo3 + cmx_dec(occ[off + d])];
    }
  } else if (in_jk && (g.xq_log ? ((i & ((1 << g.xq_log) - 1)) >= R[0] && (i & ((1 << g.xq_log) - 1)) < (1 << g.xq_log) - R[0])
                                : (i >= R[0] && i < g.N0 - R[0]))) {
    const int xs = g.xq_log ? 4 : 1;
#pragma unroll 6
    for (int s = 0; s < n_act; ++s) dE += Vrow[s * mo3 + cmx_dec(occ[off + shD[s] + shN[s].x * xs])];
  } else {
    for (int s = 0; s < n_act; ++s) {
      const int4 o = shN[s];
      int ii = i + o.x, jj = j + o.y, kk = k + o.z;
      cmx_wrap_cell(g, ii, jj, kk);
      dE += Vrow[s * mo3 + cmx_dec(occ[cmx_site_offset(g, o.w, ii, jj, kk)])];
    }
  }
  return dE;
}
// the tables of point position p into shared memory: [neighbor offsets][V][byte offsets of the (dj, dk, sublattice) part]
__device__ __forceinline__ void pairsum_stage(const DevTables &T, const Geom &g, const PairSumArgs &ps, int p,
                                              unsigned char *smem, double *&shV, int4 *&shN, int32_t *&shD, int &n_act) {
  const int mo3 = T.max_occ * T.max_occ * T.max_occ;
  const int ab = ps.act_beg[p];
  n_act = ps.act_beg[p + 1] - ab;
  const int b = T.nlist_sublat[p];
  shN = reinterpret_cast<int4 *>(smem);  // 16-byte entries first: alignment
  shV = reinterpret_cast<double *>(shN + n_act);
  shD = reinterpret_cast<int32_t *>(shV + (size_t)n_act * mo3);
  for (int q = threadIdx.x; q < n_act * mo3; q += blockDim.x) shV[q] = ps.V[(size_t)ab * mo3 + q];
  for (int q = threadIdx.x; q < n_act; q += blockDim.x) {
    const int4 o = T.nbr[ps.act_n[ab + q]];
    shN[q] = o;
    shD[q] = (int32_t)((int64_t)(o.w - b) * g.sub_stride + (int64_t)o.z * g.layer + (int64_t)o.y * g.N0);  // (x apart)
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_sweep_pairsum(GenericSweepArgs a, PairSumArgs ps) {
  extern __shared__ __align__(16) unsigned char sh_ps[];
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
  double *shV;
  int4 *shN;
  int32_t *shD;
  int n_act;
  pairsum_stage(T, g, ps, a.p, sh_ps, shV, shN, shD, n_act);
  const double *P0 = ps.P0 + (size_t)a.p * mo * mo;
  long long n_acc = 0;
  double e_sum = 0.0;
  for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < a.items; item += gridDim.x * blockDim.x) {
    uint32_t row, ii, kk, jj;
    fastdivmod(item, a.div0, row, ii);
    fastdivmod(row, a.div1, kk, jj);
    const int i = (int)ii * a.S0 + a.c0, j = (int)jj * a.S1 + a.c1, k = (int)kk * a.S2 + a.c2;
    const int64_t off = cmx_site_offset(g, b, i, j, k);
    const int oi = cmx_dec(occ[off]);
    int alt;
    uint32_t u_hi, u_lo;
    Rng16Site rs;
    uint32_t gid = 0;
    if (a.rng16) {
      rs = cmx_rng16_site(g, i);
      gid = (uint32_t)(((uint32_t)(k + a.k_offset) * g.N1 + j) * (g.N0 >> 4) + rs.chunk);
      const Philox ph = philox_sweep(gid, (uint32_t)r, a.sweep_lo, a.ctr_hi, a.k0, a.k1);
      const uint32_t R = ph.c[rs.word];
      const uint32_t field = rs.half ? (R >> 16) : (R & 0xFFFFu);
      alt = (nocc == 3) ? (int)(field >> 15) : 0;
      u_hi = field & 0x7FFFu;
      u_lo = 0;
    } else {
      const uint32_t gid_lo = (uint32_t)(((uint64_t)(k + a.k_offset) * g.N1 + j) * g.N0 + i);
      const Philox ph = philox4x32_10(gid_lo, (uint32_t)r | ((uint32_t)b << 24), a.sweep_lo, a.ctr_hi, a.k0, a.k1);
      alt = (int)__umulhi(ph.c[2], (uint32_t)(nocc - 1));
      u_hi = ph.c[1] & 0x1FFFFFu;
      u_lo = ph.c[0];
    }
    int of = oi + 1 + alt;
    if (of >= nocc) of -= nocc;
    double dE = pairsum_delta(g, occ, shV, shN, shD, n_act, mo, ps.R, ps.fast, P0[oi * mo + of], i, j, k, off, oi, of);
    dE -= exch[oi * mo + of];
    bool accept = dE < 0.0;
    if (!accept) {
      if (a.rng16) {
        // the pair16 kernel's integer test, u47 < ceil(exp(-dE beta) 2^47); the low 32 bits
        // only decide when the high 15 tie
        const unsigned long long thr = (unsigned long long)ceil(exp(-dE * beta) * 140737488355328.0);
        const uint32_t t_hi = (uint32_t)(thr >> 32);
        accept = u_hi < t_hi;
        if (u_hi == t_hi) {
          const Philox lo = philox_sweep(gid, (uint32_t)r, a.sweep_lo, a.ctr_hi | ((rs.q < 4) ? 1u : 2u), a.k0, a.k1);
          accept = lo.c[rs.q & 3] < (uint32_t)thr;
        }
      } else {
        const unsigned long long u = ((unsigned long long)u_hi << 32) | u_lo;
        accept = (double)u * (1.0 / 9007199254740992.0) < exp(-dE * beta);
      }
    }
    if (accept) {
      occ[off] = (int8_t)cmx_enc(g, of);
      ++n_acc;
      if (a.accum) e_sum += dE;
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

// ---------------------------------------------------------------------------
// generic sweep kernel, warp-cooperative: one site per WARP (cmx_warp_site_delta).
// Same random bits and the same decision rule as k_sweep_generic; dE differs from it
// in the last bits only (summation order).  Wide orbit sets (ZrO: ~700 merged terms,
// 225 neighbors per site) make the one-thread-per-site kernel latency bound and leave
// most of the chip idle when a colour holds a few thousand sites.
// ---------------------------------------------------------------------------
// SH: the term tables are staged in shared memory first (dynamic shared memory:
// [tables, table_bytes each][8 warps][stage_max + 1]).  The kernel works through the colours
// [col_begin, col_end) of a sweep in the order of the host loop (k colour outermost, point
// position innermost); COOP: the whole range in one cooperative launch with a grid barrier
// between colours (a colour of a wide-orbit model on a 48^3 box is 3 072 sites: ~5 us of
// work against ~4 us of launch), the tables of all mutable point positions staged once,
// occupations read through L2 only.
struct GenColours {
  int col_begin, col_end;
  int n_mut;
  int mut_p[8];
  uint32_t ctr_base;  // (sweep >> 32) << 16
};
template <bool SH, bool COOP>
__global__ void __launch_bounds__(256) k_sweep_generic_warp(GenericSweepArgs a, GenTerms G, int stage_max,
                                                            int table_bytes, GenColours C) {
  extern __shared__ __align__(16) unsigned char sh_dyn_g[];
  const int n_tab = SH ? (COOP ? C.n_mut : 1) : 0;
  double *sh_stage = reinterpret_cast<double *>(sh_dyn_g + (size_t)n_tab * table_bytes);
  __shared__ long long sh_acc[8];
  __shared__ double sh_sum[8];
  const int r = blockIdx.y;
  const Geom &g = a.g;
  const DevTables &T = a.T;
  GenShared Sq[8];
  if (SH) {
    if (COOP) {
      for (int q = 0; q < C.n_mut; ++q) cmx_gen_stage_table(T, G, C.mut_p[q], sh_dyn_g + (size_t)q * table_bytes, Sq[q]);
    } else {
      cmx_gen_stage_table(T, G, C.mut_p[C.col_begin % C.n_mut], sh_dyn_g, Sq[0]);
    }
  }
  int8_t *occ = a.occ + (size_t)r * g.rep_stride;
  const int mo = T.max_occ;
  const double beta = a.beta[r];
  const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  double *sh_val = sh_stage + (size_t)wib * (stage_max + 1);
  long long n_acc = 0;
  double e_sum = 0.0;
  for (int col = C.col_begin; col < C.col_end; ++col) {
    const int q = col % C.n_mut, p = C.mut_p[q];
    const int cc = col / C.n_mut;
    const int c0 = cc % a.S0, c1 = (cc / a.S0) % a.S1, c2 = cc / (a.S0 * a.S1);
    const uint32_t ctr_hi = C.ctr_base | (a.rng16 ? (((uint32_t)col & 0xffu) << 8) : ((uint32_t)col & 0xffffu));
    const GenShared &S = Sq[COOP ? q : 0];
    const int b = T.nlist_sublat[p];
    const int nocc = T.n_occ[b];
    const double *exch = a.exch + (size_t)r * a.exch_stride + (size_t)b * mo * mo;
    if (SH && lane == 0) sh_val[S.n_act * T.n_func] = 1.0;  // the unused-factor slot
    __syncwarp();
    for (uint32_t item = blockIdx.x * 8u + wib; item < a.items; item += gridDim.x * 8u) {
      uint32_t row, ii, kk, jj;
      fastdivmod(item, a.div0, row, ii);
      fastdivmod(row, a.div1, kk, jj);
      const int i = (int)ii * a.S0 + c0, j = (int)jj * a.S1 + c1, k = (int)kk * a.S2 + c2;
      const int64_t off = cmx_site_offset(g, b, i, j, k);
      const int oi = cmx_dec(COOP ? (int)__ldcg(occ + off) : (int)occ[off]);
      int alt;
      uint32_t u_hi, u_lo;
      if (a.rng16) {
        const Rng16Site rs = cmx_rng16_site(g, i);
        const uint32_t gid = (uint32_t)(((uint32_t)(k + a.k_offset) * g.N1 + j) * (g.N0 >> 4) + rs.chunk);
        const uint32_t qq = rs.q;
        const Philox ph = philox_sweep(gid, (uint32_t)r, a.sweep_lo, ctr_hi, a.k0, a.k1);
        const uint32_t R = ph.c[rs.word];
        const uint32_t field = rs.half ? (R >> 16) : (R & 0xFFFFu);
        alt = (nocc == 3) ? (int)(field >> 15) : 0;
        u_hi = field & 0x7FFFu;
        const Philox lo = philox_sweep(gid, (uint32_t)r, a.sweep_lo, ctr_hi | ((qq < 4) ? 1u : 2u), a.k0, a.k1);
        u_lo = lo.c[qq & 3];
      } else {
        const uint32_t gid_lo = (uint32_t)(((uint64_t)(k + a.k_offset) * g.N1 + j) * g.N0 + i);
        const Philox ph = philox4x32_10(gid_lo, (uint32_t)r | ((uint32_t)b << 24), a.sweep_lo, ctr_hi, a.k0, a.k1);
        alt = (int)__umulhi(ph.c[2], (uint32_t)(nocc - 1));
        u_hi = ph.c[1] & 0x1FFFFFu;
        u_lo = ph.c[0];
      }
      int of = oi + 1 + alt;
      if (of >= nocc) of -= nocc;
      double dE = SH ? cmx_warp_site_delta_sh<COOP>(T, g, S, occ, sh_val, i, j, k, oi, of, -1, 0, lane)
                     : cmx_warp_site_delta<COOP>(T, g, G, occ, sh_val, p, i, j, k, oi, of, -1, 0, lane);
      dE -= exch[oi * mo + of];
      bool accept = dE < 0.0;
      if (!accept) {
        const unsigned long long u = ((unsigned long long)u_hi << 32) | u_lo;
        if (a.rng16) accept = u < (unsigned long long)ceil(exp(-dE * beta) * 140737488355328.0);
        else accept = (double)u * (1.0 / 9007199254740992.0) < exp(-dE * beta);
      }
      if (accept && lane == 0) {
        occ[off] = (int8_t)cmx_enc(g, of);
        ++n_acc;
        if (a.accum) e_sum += dE;
      }
    }
    if (COOP && col + 1 < C.col_end) cooperative_groups::this_grid().sync();
  }
  if (lane == 0) {
    sh_acc[wib] = n_acc;
    sh_sum[wib] = e_sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long A = 0;
    double E = 0.0;
    for (int w = 0; w < 8; ++w) {
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

// FCC nearest-neighbor shell in CASM's standard primitive cell
// (offsets of m_orbit_site_neighborhood[3], FCC default clexulator :404-414)
constexpr uint32_t mbit(int dx, int dy, int dz) {
  return 1u << ((dz + 1) * 9 + (dy + 1) * 3 + (dx + 1));
}
constexpr uint32_t kMaskFcc1NN =
    mbit(-1, 0, 0) | mbit(-1, 0, 1) | mbit(-1, 1, 0) | mbit(0, -1, 0) | mbit(0, -1, 1) |
    mbit(0, 0, -1) | mbit(0, 0, 1) | mbit(0, 1, -1) | mbit(0, 1, 0) | mbit(1, -1, 0) |
    mbit(1, 0, -1) | mbit(1, 0, 0);

// ... and the second shell (neighbor-list sites 13-18 of the same clexulator): every offset has
// dx = +-1 and none lies in the site's own row
constexpr uint32_t kMaskFcc2NN =
    mbit(-1, -1, 1) | mbit(-1, 1, -1) | mbit(-1, 1, 1) | mbit(1, -1, -1) | mbit(1, -1, 1) | mbit(1, 1, -1);

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
  std::vector<double> psV, psP0;   // host copies of the pair-sum tables (two-class count table below)
  std::vector<int32_t> ps_act_n;
  {
    // warp-cooperative evaluation: per point position the neighbors its terms read,
    // per factor the index of the staged value
    std::vector<int32_t> act_beg(1, 0), act_n, gt_vi(gt_n.size(), 0);
    P.stage_max = 0;
    for (int p = 0; p < np; ++p) {
      std::vector<int32_t> slot(T.nlist_len, -1);
      for (int tt = gt_beg[p]; tt < gt_beg[p + 1]; ++tt)
        for (int f = gt_fbeg[tt]; f < gt_fbeg[tt + 1]; ++f) slot[gt_n[f]] = 0;
      int ns = 0;
      for (int n = 0; n < T.nlist_len; ++n)
        if (slot[n] == 0) {
          slot[n] = ns++;
          act_n.push_back(n);
        }
      for (int tt = gt_beg[p]; tt < gt_beg[p + 1]; ++tt)
        for (int f = gt_fbeg[tt]; f < gt_fbeg[tt + 1]; ++f) gt_vi[f] = slot[gt_n[f]] * T.n_func + gt_f[f];
      act_beg.push_back((int32_t)act_n.size());
      P.stage_max = std::max(P.stage_max, ns * T.n_func);
    }
    if ((rc = to_device(gt_vi, &P.d_gt_vi))) return rc;
    if ((rc = to_device(act_beg, &P.d_act_beg))) return rc;
    if ((rc = to_device(act_n, &P.d_act_n))) return rc;
    // packed form for the shared-memory path: <= 4 factors per term, 16-bit indices
    std::vector<uint2> pk(gt_fbeg.size() - 1);
    bool packable = true;
    P.pk_terms_max = P.pk_act_max = 0;
    for (int p = 0; p < np && packable; ++p) {
      const int ns = act_beg[p + 1] - act_beg[p];
      const uint32_t one = (uint32_t)(ns * T.n_func);  // the slot that holds 1.0
      if (one > 0xFFFFu) packable = false;
      P.pk_terms_max = std::max(P.pk_terms_max, gt_beg[p + 1] - gt_beg[p]);
      P.pk_act_max = std::max(P.pk_act_max, ns);
      for (int tt = gt_beg[p]; tt < gt_beg[p + 1] && packable; ++tt) {
        uint32_t v[4] = {one, one, one, one};
        const int nfac = gt_fbeg[tt + 1] - gt_fbeg[tt];
        if (nfac > 4) {
          packable = false;
          break;
        }
        for (int f = 0; f < nfac; ++f) v[f] = (uint32_t)gt_vi[gt_fbeg[tt] + f];
        pk[tt] = make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
      }
    }
    if (packable && (rc = to_device(pk, &P.d_gt_pk))) return rc;
    // pair-sum tables: V[slot][oi][of][on] = sum over the one-factor terms (slot, f) of
    // w[oi][of] phi_f(on), P0[p][oi][of] = the factor-free terms; both in term order
    bool single = mo <= 4 && !P.mut_points.empty();
    for (int tt = 0; tt + 1 < (int)gt_fbeg.size() && single; ++tt) single = gt_fbeg[tt + 1] - gt_fbeg[tt] <= 1;
    const int mo3 = mo * mo * mo;
    for (int p = 0; p < np && single; ++p)
      single = (size_t)(act_beg[p + 1] - act_beg[p]) * (mo3 * 8 + 20) <= 44 * 1024;
    if (single) {
      std::vector<double> V((size_t)act_n.size() * mo3, 0.0), P0((size_t)np * mo * mo, 0.0);
      for (int p = 0; p < np; ++p)
        for (int tt = gt_beg[p]; tt < gt_beg[p + 1]; ++tt) {
          const double *w = &gt_w[(size_t)tt * mo * mo];
          if (gt_fbeg[tt + 1] == gt_fbeg[tt]) {
            for (int x = 0; x < mo * mo; ++x) P0[(size_t)p * mo * mo + x] += w[x];
            continue;
          }
          const int f = gt_fbeg[tt], n = gt_n[f];
          const int slot = act_beg[p] + gt_vi[f] / T.n_func;
          const double *ph = &t->phi[((size_t)t->nbr[4 * n + 3] * T.n_func + gt_f[f]) * mo];
          for (int x = 0; x < mo * mo; ++x)
            for (int on = 0; on < mo; ++on) V[((size_t)slot * mo * mo + x) * mo + on] += w[x] * ph[on];
        }
      for (int a = 0; a < 3; ++a) P.ps_range[a] = 0;
      for (int n : act_n)
        for (int a = 0; a < 3; ++a) P.ps_range[a] = std::max(P.ps_range[a], std::abs(t->nbr[4 * n + a]));
      if ((rc = to_device(V, &P.d_ps_V))) return rc;
      if ((rc = to_device(P0, &P.d_ps_P0))) return rc;
      P.pair_sum = true;
      psV = V;
      psP0 = P0;
      ps_act_n = act_n;
    }
  }

  // ---- colouring: stride > interaction range along each axis.  Sites on
  // different sublattices of one cell interact through offset (0,0,0), so the
  // sublattice is part of the colour.
  int R[3] = {0, 0, 0};
  std::vector<char> active(T.nlist_len, 0);
  for (int n : gt_n) active[n] = 1;
  P.active_nbr.clear();
  for (int n = 0; n < T.nlist_len; ++n)
    if (active[n]) P.active_nbr.push_back(n);
  for (int n = 0; n < T.nlist_len; ++n)
    if (active[n])
      for (int a = 0; a < 3; ++a) R[a] = std::max(R[a], std::abs(t->nbr[4 * n + a]));
  int N[3] = {s->g.N0, s->g.N1, s->g.N2};
  for (int a = 0; a < 3; ++a) P.S[a] = smallest_divisor_above(N[a], R[a]);
  // skewed boxes: leaving the box along i shifts j by s10 and k by s20, along j shifts k by
  // s21 -- the colour (i mod S0, j mod S1, k mod S2) survives the wrap only if the strides
  // divide the shifts
  const bool skewed = (s->g.s10 | s->g.s20 | s->g.s21) != 0;
  bool colour_ok = true;
  if (skewed) {
    auto pick = [&](int n, int r, int d1, int d2) {
      for (int q = r + 1; q <= n; ++q)
        if (n % q == 0 && d1 % q == 0 && d2 % q == 0) return q;
      return 0;
    };
    P.S[1] = pick(N[1], R[1], s->g.s10, 0);
    P.S[2] = pick(N[2], R[2], s->g.s20, s->g.s21);
    colour_ok = P.S[1] > 0 && P.S[2] > 0;
  }
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
  P.valid = colour_ok;  // (no colouring of this skewed box: no checkerboard sweeps; everything else works)
  if (!colour_ok) return CMX_OK;

  // ---- pair-LUT eligibility
  bool ok = (!skewed && T.n_sublat == 1 && np == 1 && P.mut_points.size() == 1 &&
             t->n_occ[0] >= 2 && t->n_occ[0] <= 3 && T.nlist_len <= 64 &&
             s->g.N0 % 16 == 0 && s->g.N0 <= 4096 && s->g.N1 % 2 == 0 && s->g.N2 % 2 == 0 && s->g.N2 <= 65534 &&
             s->g.rep_stride < (int64_t)0x7FFFFFFFll &&
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
    // byte-lane sums n1 + 18 n2 (+ the column skew of the acceptance table, up to 22) must stay below 256
    if ((int)V.size() > (t->n_occ[0] == 3 ? 12 : 14)) ok = false;
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
      if (cls.size() < 16) {
        P.shell[3 * cls.size()] = o[0];
        P.shell[3 * cls.size() + 1] = o[1];
        P.shell[3 * cls.size() + 2] = o[2];
      }
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
    P.n_tab = CMX_TAB16(P.nocc);
    CMX_CUDA(cudaMalloc((void **)&P.d_tab, sizeof(uint32_t) * P.n_tab * s->n_replicas));
    CMX_CUDA(cudaMalloc((void **)&P.d_thr_lo, sizeof(uint32_t) * P.n_tab * s->n_replicas));
    CMX_CUDA(cudaMalloc((void **)&P.d_dEpot, sizeof(double) * P.n_tab * s->n_replicas));
    // x4-interleaved rows: the streaming kernel; linear rows: the block kernel
    P.stream = s->g.xq_log != 0;
    P.n_tab24 = CMX_TAB24(P.nocc);
    CMX_CUDA(cudaMalloc((void **)&P.d_tab24, sizeof(uint32_t) * P.n_tab24 * s->n_replicas));
    P.pair_lut = true;
    P.rng16 = true;  // the generic evaluator on this state mirrors the pair16 random bits
    // per step: z neighbor bytes + own byte read, 1 byte written; the FP64
    // work is the tabulated dE (2 site-function adds per neighbor, the ECI
    // dot, the exp folded into the threshold) -- report the table-free count
    P.bytes_per_step = P.z + 2;
  }
  // ---- two-class count table: a point + pair basis on one ternary sublattice whose active
  // neighbors are exactly the FCC 1NN and 2NN shells with one table per shell, on
  // x4-interleaved rows of one GPU's box.  dE per (oi, alt, class-1 counts, class-2 counts) is
  // summed HERE in the pair-sum evaluator's order (P0, then the neighbors in slot order) on a
  // representative arrangement, so the table holds values k_sweep_pairsum itself produces.
  if (!P.pair_lut && P.pair_sum && !skewed && T.n_sublat == 1 && np == 1 && P.mut_points.size() == 1 &&
      t->n_occ[0] == 3 && mo == 3 && s->g.coded && s->g.xq_log != 0 && s->g.N0 % 16 == 0 &&
      s->g.N1 % 2 == 0 && s->g.N2 % 2 == 0 && s->g.N2 <= 65534 && ((uint64_t)s->g.rep_stride >> 4) < (1ull << 32) &&
      P.S[0] == 2 && P.S[1] == 2 && P.S[2] == 2 && !ps_act_n.empty() && ps_act_n.size() <= 32) {
    const int n_act = (int)ps_act_n.size(), mo3 = mo * mo * mo;
    double scale = 0;
    for (double x : psV) scale = std::max(scale, std::fabs(x));
    auto same = [&](int a, int b) {
      for (int x = 0; x < mo3; ++x)
        if (std::fabs(psV[(size_t)a * mo3 + x] - psV[(size_t)b * mo3 + x]) > 1e-12 * scale) return false;
      return true;
    };
    std::vector<int> cls(n_act, -1);
    int rep[2] = {0, 0}, n_cls = 0;
    bool two = true;
    for (int a = 0; a < n_act && two; ++a) {
      for (int c = 0; c < n_cls; ++c)
        if (same(a, rep[c])) cls[a] = c;
      if (cls[a] < 0) {
        if (n_cls == 2) two = false;
        else {
          rep[n_cls] = a;
          cls[a] = n_cls++;
        }
      }
    }
    uint32_t m[2] = {0, 0};
    int zc[2] = {0, 0};
    if (two && n_cls == 2)
      for (int a = 0; a < n_act; ++a) {
        const int32_t *o = &t->nbr[4 * ps_act_n[a]];
        if (std::abs(o[0]) > 1 || std::abs(o[1]) > 1 || std::abs(o[2]) > 1 || (o[0] == 0 && o[1] == 0 && o[2] == 0)) two = false;
        else m[cls[a]] |= mbit(o[0], o[1], o[2]);
        ++zc[cls[a]];
      }
    int c1 = -1;
    if (two && n_cls == 2) {
      if (m[0] == kMaskFcc1NN && m[1] == kMaskFcc2NN) c1 = 0;
      else if (m[1] == kMaskFcc1NN && m[0] == kMaskFcc2NN) c1 = 1;
    }
    if (c1 >= 0 && zc[c1] == 12 && zc[1 - c1] == 6) {
      const int z1 = 12, z2 = 6, nocc = 3;
      const int n1 = (z1 + 1) * (z1 + 2) / 2, n2 = (z2 + 1) * (z2 + 2) / 2, per_row = n1 * n2;
      const int n_tab2 = nocc * (nocc - 1) * per_row;
      auto tri = [](int z, int nB, int nV) { return nV * (z + 1) - nV * (nV - 1) / 2 + nB; };
      std::vector<uint16_t> maps(CMX_S16_MAPS, 0);
      for (int nV = 0; nV <= z1; ++nV)
        for (int nB = 0; nB + nV <= z1; ++nB) maps[nB + CMX_VA_CODE * nV] = (uint16_t)(tri(z1, nB, nV) * n2);
      for (int nV = 0; nV <= z2; ++nV)
        for (int nB = 0; nB + nV <= z2; ++nB) maps[256 + nB + CMX_VA_CODE * nV] = (uint16_t)tri(z2, nB, nV);
      for (int oi = 0; oi < nocc; ++oi)
        for (int alt = 0; alt < nocc - 1; ++alt)
          maps[512 + ((oi == 2 ? CMX_VA_CODE : oi) | (alt << 2))] = (uint16_t)((oi * (nocc - 1) + alt) * per_row);
      std::vector<double> lut2((size_t)n_tab2, 0.0);
      std::vector<int> occ_n(n_act);
      for (int oi = 0; oi < nocc; ++oi)
        for (int alt = 0; alt < nocc - 1; ++alt) {
          int of = oi + 1 + alt;
          if (of >= nocc) of -= nocc;
          const size_t row = (size_t)(oi * (nocc - 1) + alt) * per_row;
          for (int nV1 = 0; nV1 <= z1; ++nV1)
            for (int nB1 = 0; nB1 + nV1 <= z1; ++nB1)
              for (int nV2 = 0; nV2 <= z2; ++nV2)
                for (int nB2 = 0; nB2 + nV2 <= z2; ++nB2) {
                  int seen[2] = {0, 0};
                  for (int a = 0; a < n_act; ++a) {
                    const bool first = (cls[a] == c1);
                    const int nB = first ? nB1 : nB2, nV = first ? nV1 : nV2;
                    const int q = seen[first ? 0 : 1]++;
                    occ_n[a] = (q < nB) ? 1 : ((q < nB + nV) ? 2 : 0);
                  }
                  volatile double dE = psP0[(size_t)oi * mo + of];
                  for (int a = 0; a < n_act; ++a) dE = dE + psV[(((size_t)a * mo + oi) * mo + of) * mo + occ_n[a]];
                  lut2[row + (size_t)tri(z1, nB1, nV1) * n2 + tri(z2, nB2, nV2)] = dE;
                }
        }
      if ((rc = to_device(lut2, &P.d_pair2_dE))) return rc;
      if ((rc = to_device(maps, &P.d_maps2))) return rc;
      CMX_CUDA(cudaMalloc((void **)&P.d_tab2, sizeof(uint32_t) * n_tab2 * s->n_replicas));
      CMX_CUDA(cudaMalloc((void **)&P.d_thr_lo2, sizeof(uint32_t) * n_tab2 * s->n_replicas));
      CMX_CUDA(cudaMalloc((void **)&P.d_dEpot2, sizeof(double) * n_tab2 * s->n_replicas));
      P.nocc = nocc;
      P.z = z1;
      P.z2 = z2;
      P.mask = kMaskFcc1NN;
      P.mask2 = kMaskFcc2NN;
      int q1 = 0, q2 = 0;
      for (int a = 0; a < n_act; ++a) {
        const int32_t *o = &t->nbr[4 * ps_act_n[a]];
        int32_t *dst = (cls[a] == c1) ? &P.shell[3 * q1++] : &P.shell2[3 * q2++];
        dst[0] = o[0];
        dst[1] = o[1];
        dst[2] = o[2];
      }
      P.n_tab2 = n_tab2;
      // slabs (ghost layers) keep the pair-sum kernel, but draw the SAME random bits: the
      // trajectory of a box does not depend on its decomposition
      P.pair2 = s->g.halo == 0;
      P.rng16 = true;  // the pair-sum and term-list evaluators of this state mirror the kernel's random bits
    }
  }
  P.thr_dirty = true;
  if (P.pair_lut) return cmx_plan_energy(s);
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

template <int NOCC, int MINB>
static void launch_pair16(const Pair16Args &a, dim3 grid, cudaStream_t st, bool fcc, bool accum) {
  if (fcc) {
    if (accum) k_sweep_pair16<NOCC, kMaskFcc1NN, true, MINB><<<grid, 256, 0, st>>>(a);
    else k_sweep_pair16<NOCC, kMaskFcc1NN, false, MINB><<<grid, 256, 0, st>>>(a);
  } else {
    if (accum) k_sweep_pair16<NOCC, 0u, true, MINB><<<grid, 256, 0, st>>>(a);
    else k_sweep_pair16<NOCC, 0u, false, MINB><<<grid, 256, 0, st>>>(a);
  }
}

// tuning knobs (environment, read once)
static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
static int pair16_minb() {  // resident blocks per SM k_sweep_pair16 is compiled for
  static int v = env_int("CMX_PAIR16_MINB", 3);
  return v;
}
static int sweep_grid_per_sm() {  // blocks per SM in the grid
  static int v = env_int("CMX_SWEEP_BLOCKS_PER_SM", 3);
  return v;
}

// wide orbit sets: one site per warp (k_sweep_generic_warp)
bool cmx_use_warp_generic(const cmx_state *s) {
  const SweepPlan &P = s->plan;
  if ((s->sweep_flags & CMX_SWEEP_THREAD_GENERIC) || P.stage_max <= 0 || P.mut_points.empty()) return false;
  if ((size_t)(P.stage_max + 1) * 8 * sizeof(double) > 96 * 1024) return false;
  return P.n_gterms / (int)P.mut_points.size() >= 96;
}

static bool use_pair(const cmx_state *s) {
  return s->plan.pair_lut && !(s->sweep_flags & CMX_SWEEP_FORCE_GENERIC);
}
// point + pair bases the pair-LUT path does not take (several neighbor classes, several
// sublattices, box shapes): per-neighbor tables, one site per thread.  Either generic flag
// selects the term-list evaluators instead (the cross-check of this one).
static bool pair_sum_allowed(const cmx_state *s) {
  return s->plan.pair_sum && !use_pair(s) &&
         !(s->sweep_flags & (CMX_SWEEP_FORCE_GENERIC | CMX_SWEEP_THREAD_GENERIC));
}
// two neighbor classes on x4-interleaved rows: the colour-pass kernel with the two-class count
// table, unless a flag asks for one of the evaluators it is cross-checked against
static bool use_pair2(const cmx_state *s) {
  return s->plan.pair2 && pair_sum_allowed(s) && !(s->sweep_flags & CMX_SWEEP_PAIR_SUM);
}
bool cmx_use_pair_sum(const cmx_state *s) { return pair_sum_allowed(s) && !use_pair2(s); }
static bool use_stream(const cmx_state *s) { return use_pair(s) && s->plan.stream; }

static int sweep_blocks_per_replica(uint32_t items, int n_replicas) {
  int want = (int)((items + 255) / 256);
  // one co-resident wave: rounding the per-replica share UP would leave a few blocks of
  // the last replicas for a second wave
  int cap = std::max(1, (148 * sweep_grid_per_sm()) / n_replicas);
  return std::max(1, std::min(want, cap));
}

// acceptance tables of the pair-LUT kernels, rebuilt when ECI or conditions changed
static int pair_tables(cmx_state *s) {
  SweepPlan &P = s->plan;
  const DevTables &T = s->t->d;
  if (!P.thr_dirty) return CMX_OK;
  const size_t exs = (size_t)T.n_sublat * T.max_occ * T.max_occ;
  dim3 grid((P.n_tab + 127) / 128, s->n_replicas);
  k_build_tab16<<<grid, 128, 0, s->stream>>>(P.d_pair_dE, P.nocc, P.z, T.max_occ, s->d_beta, s->d_exch,
                                             (int)exs, P.n_tab, P.d_tab, P.d_thr_lo, P.d_dEpot);
  dim3 grid24((P.n_tab24 + 127) / 128, s->n_replicas);
  k_build_tab24<<<grid24, 128, 0, s->stream>>>(P.d_tab, P.nocc, P.n_tab, P.n_tab24, P.d_tab24);
  CMX_CUDA(cudaGetLastError());
  P.thr_dirty = false;
  return CMX_OK;
}

static int pair2_tables(cmx_state *s) {
  SweepPlan &P = s->plan;
  const DevTables &T = s->t->d;
  if (!P.thr_dirty) return CMX_OK;
  const size_t exs = (size_t)T.n_sublat * T.max_occ * T.max_occ;
  dim3 grid((P.n_tab2 + 127) / 128, s->n_replicas);
  k_build_tab2<<<grid, 128, 0, s->stream>>>(P.d_pair2_dE, P.n_tab2, P.n_tab2 / (P.nocc * (P.nocc - 1)), P.nocc, T.max_occ,
                                            s->d_beta, s->d_exch, (int)exs, P.d_tab2, P.d_thr_lo2, P.d_dEpot2);
  CMX_CUDA(cudaGetLastError());
  P.thr_dirty = false;
  return CMX_OK;
}

// arguments of the block kernel
static int pair_args(cmx_state *s, uint64_t seed, int64_t sweep, int k_offset, Pair16Args &a) {
  SweepPlan &P = s->plan;
  const Geom &g = s->g;
  int rc = pair_tables(s);
  if (rc) return rc;
  a.occ = s->d_occ;
  a.g = g;
  a.mask = P.mask;
  a.W = g.N0 / 16;
  a.RB = 256 / a.W;
  a.divW = make_fastdiv(a.W);
  a.J = g.N1 / 2;
  a.divJ = make_fastdiv(a.J);
  a.n_rows = a.J * (uint32_t)(g.N2 / 2);
  a.tab = P.d_tab;
  a.thr_lo = P.d_thr_lo;
  a.dEpot = P.d_dEpot;
  a.part_acc = P.d_part_acc;
  a.part_dE = P.d_part_dE;
  a.part_stride = (uint32_t)P.part_blocks;
  a.k0 = (uint32_t)seed;
  a.k1 = (uint32_t)(seed >> 32);
  a.rk = philox_key_schedule(a.k0, a.k1);
  a.sweep_lo = (uint32_t)sweep;
  a.ctr_hi = 0;
  a.cy = a.cz = 0;
  a.k_offset = k_offset;
  return CMX_OK;
}

// ---------------------------------------------------------------------------
// streaming kernel: schedule and launch
// ---------------------------------------------------------------------------
typedef void (*StreamKernel)(S16Args);
template <int NOCC, uint32_t MASK, bool FULL>
static StreamKernel stream_kernel_nmf(bool accum, bool slab) {
  if (accum) return slab ? k_sweep_stream16<NOCC, MASK, true, true, FULL> : k_sweep_stream16<NOCC, MASK, true, false, FULL>;
  return slab ? k_sweep_stream16<NOCC, MASK, false, true, FULL> : k_sweep_stream16<NOCC, MASK, false, false, FULL>;
}
template <int NOCC, uint32_t MASK>
static StreamKernel stream_kernel_nm(bool accum, bool slab, bool full) {
  return full ? stream_kernel_nmf<NOCC, MASK, true>(accum, slab) : stream_kernel_nmf<NOCC, MASK, false>(accum, slab);
}
static StreamKernel stream_kernel(const cmx_state *s, bool accum, bool slab, size_t *smem) {
  const SweepPlan &P = s->plan;
  const bool fcc = (P.mask == kMaskFcc1NN);
  // the rows of one colour of a layer divide evenly into row-steps: no partial row-step
  const uint32_t rpw = 32u / ((uint32_t)s->g.N0 / 16u);
  const bool full = ((uint32_t)s->g.N1 / 2u) % rpw == 0;
  if (P.nocc == 3) {
    *smem = s16_smem_bytes<3>(fcc ? kMaskFcc1NN : 0u);
    return fcc ? stream_kernel_nm<3, kMaskFcc1NN>(accum, slab, full) : stream_kernel_nm<3, 0u>(accum, slab, full);
  }
  *smem = s16_smem_bytes<2>(fcc ? kMaskFcc1NN : 0u);
  return fcc ? stream_kernel_nm<2, kMaskFcc1NN>(accum, slab, full) : stream_kernel_nm<2, 0u>(accum, slab, full);
}

typedef void (*PassKernel)(S16Args, PassArgs);
template <int NOCC, uint32_t MASK, bool FULL>
static PassKernel pass_kernel_nmf(bool accum, bool slab) {
  if (slab) return accum ? k_sweep_pass16<NOCC, MASK, true, true, FULL> : k_sweep_pass16<NOCC, MASK, false, true, FULL>;
  return accum ? k_sweep_pass16<NOCC, MASK, true, false, FULL> : k_sweep_pass16<NOCC, MASK, false, false, FULL>;
}
// (slab: the instantiation that carries the ring protocol and the peer stores)
static PassKernel pass_kernel(const cmx_state *s, bool accum, size_t *smem) {
  const SweepPlan &P = s->plan;
  const bool fcc = (P.mask == kMaskFcc1NN);
  const bool slab = s->p2p && s->g.halo;
  const uint32_t rpw = 32u / ((uint32_t)s->g.N0 / 16u);
  const bool full = ((uint32_t)s->g.N1 / 2u) % rpw == 0;
  if (P.nocc == 3) {
    *smem = s16_smem_bytes<3>(fcc ? kMaskFcc1NN : 0u);
    if (fcc) return full ? pass_kernel_nmf<3, kMaskFcc1NN, true>(accum, slab) : pass_kernel_nmf<3, kMaskFcc1NN, false>(accum, slab);
    return full ? pass_kernel_nmf<3, 0u, true>(accum, slab) : pass_kernel_nmf<3, 0u, false>(accum, slab);
  }
  *smem = s16_smem_bytes<2>(fcc ? kMaskFcc1NN : 0u);
  if (fcc) return full ? pass_kernel_nmf<2, kMaskFcc1NN, true>(accum, slab) : pass_kernel_nmf<2, kMaskFcc1NN, false>(accum, slab);
  return full ? pass_kernel_nmf<2, 0u, true>(accum, slab) : pass_kernel_nmf<2, 0u, false>(accum, slab);
}

// two neighbor classes (FCC 1NN + 2NN, ternary)
static PassKernel pass2_kernel(const cmx_state *s, bool accum, size_t *smem) {
  const uint32_t rpw = 32u / ((uint32_t)s->g.N0 / 16u);
  const bool full = ((uint32_t)s->g.N1 / 2u) % rpw == 0;
  *smem = (size_t)s16_tab2_bytes((uint32_t)s->plan.n_tab2) + 8u * s16_n_slots(kMaskFcc1NN | kMaskFcc2NN) * CMX_S16_SLOT;
  if (full)
    return accum ? k_sweep_pass16<3, kMaskFcc1NN, true, false, true, kMaskFcc2NN>
                 : k_sweep_pass16<3, kMaskFcc1NN, false, false, true, kMaskFcc2NN>;
  return accum ? k_sweep_pass16<3, kMaskFcc1NN, true, false, false, kMaskFcc2NN>
               : k_sweep_pass16<3, kMaskFcc1NN, false, false, false, kMaskFcc2NN>;
}

// x4-interleaved rows: colour passes with grid barriers (k_sweep_pass16, the default: measured
// faster than the streaming kernel on every configuration of this round) unless
// CMX_SWEEP_STREAM asks for the barrier-free streaming kernel.  Slab states decide once (they
// must not mix the two ring protocols, layer counters / epochs): the flag as it is when the
// first sweep runs.
static bool use_pass(const cmx_state *s) {
  return use_stream(s) && !(s->sweep_flags & CMX_SWEEP_STREAM);
}

// geometry of the schedule: row-steps per unit, blocks per replica, row-steps per group,
// and the distance (in units) the list keeps between dependent units
static int stream_geometry(cmx_state *s) {
  SweepPlan &P = s->plan;
  const Geom &g = s->g;
  const bool accum = (s->sweep_flags & CMX_SWEEP_DE_SUM) != 0;
  const bool slab = s->p2p && g.halo;
  size_t smem = 0;
  const void *kern = use_pair2(s)  ? (const void *)pass2_kernel(s, accum, &smem)
                     : use_pass(s) ? (const void *)pass_kernel(s, accum, &smem)
                                   : (const void *)stream_kernel(s, accum, slab, &smem);
  CMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0, dev = 0, sms = 0, can = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem) != cudaSuccess) per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, dev);
  static const int stream_per_sm = env_int("CMX_STREAM_BLOCKS_PER_SM", CMX_S16_MINB);
  per_sm = std::min(per_sm, stream_per_sm);
  P.stream_capacity = can ? per_sm * sms : 0;
  if (P.stream_capacity < s->n_replicas) {
    cmx_set_error("streaming sweep: the replicas do not fit a co-resident grid on this device");
    return CMX_ERR_UNSUPPORTED;
  }
  const uint32_t W = (uint32_t)g.N0 / 16, rpw = 32u / W, J = (uint32_t)g.N1 / 2, H = (uint32_t)g.N2 / 2;
  const uint32_t tpu = (J + rpw - 1) / rpw;
  // at least ~4 row-steps per warp and sweep, at most the co-resident share of a replica
  const uint32_t want = std::max(1u, (4u * H * tpu + 31u) / 32u);
  P.stream_blocks = (int)std::min<uint32_t>((uint32_t)(P.stream_capacity / s->n_replicas), want);
  static const int gr_env = env_int("CMX_STREAM_GR", 0);
  auto gap_for = [&](uint32_t gr) {
    // units in flight.  The warps advance in rounds (group g, then g + n_warps): a unit is
    // checked while its warp still works on the previous round, so what it depends on must
    // lie two rounds back to be complete at that time
    const uint32_t window = ((uint32_t)P.stream_blocks * 8u * gr + tpu - 1) / tpu + 1u;
    // (measured, 512^3: 3 windows cost 1.6x the algorithmic DRAM traffic -- the wavefront no
    // longer fits L2 -- for no gain in time; 2 windows: 1.1x)
    return 2u * window + 6u;
  };
  uint32_t gr = (tpu % 4 == 0 && gap_for(4) <= H) ? 4u : ((tpu % 2 == 0 && gap_for(2) <= H) ? 2u : 1u);
  if (gr_env > 0 && tpu % (uint32_t)gr_env == 0) gr = (uint32_t)gr_env;
  P.stream_gr = gr;
  static const int gap_env = env_int("CMX_STREAM_GAP", 0);
  P.stream_gap = gap_env > 0 ? (uint32_t)gap_env : gap_for(gr);
  return CMX_OK;
}

// Order of the units of a call: n sweeps, k colour group kg (-1: both).  Phases of a
// sweep: 0/1 = even layers 2t, row colour 0/1; 2/3 = odd layers 2t+1, row colour 0/1.
// True dependencies (they make the order a topological one; on the device the layer
// counters enforce them):
//   (0, t, s) <- (3, t-1, s-1), (3, t, s-1)      (1, t, s) <- (0, t, s)
//   (2, t, s) <- (1, t, s), (1, t+1, s)          (3, t, s) <- (2, t, s)          (t mod H)
// List scheduling: units are emitted in the order of a helix key (the wavefront: phase p
// of layer pair t at step s H + t + p D), but never sooner than `gap` positions after the
// units they depend on -- then the warps that take a unit find its dependencies complete.
// Where the key order would violate that (the seam of the periodic box: layer 0 of sweep
// s+1 follows the last odd layer of sweep s) other ready units fill in.
struct HostUnit {
  int s, ph, t;
};
static void stream_order(int H, int n_sweeps, int kg, int gap, std::vector<HostUnit> &out, uint32_t *min_sep) {
  const int nph = 4;
  const long long total = (long long)n_sweeps * nph * H;
  auto id_of = [&](int s, int ph, int t) { return ((long long)s * nph + ph) * H + t; };
  auto present = [&](int ph) { return kg < 0 || (ph >> 1) == kg; };
  const int D = std::max(1, gap / (kg < 0 ? 4 : 2) + 1);
  const int lag[4] = {0, D, kg < 0 ? 2 * D + 1 : 0, kg < 0 ? 3 * D + 1 : D};
  std::vector<int> emit(total, -1), ndep(total, 0);
  std::vector<long long> ready(total, 0);
  auto deps = [&](int s, int ph, int t, long long (&d)[2]) -> int {
    int n = 0;
    if (ph == 0) {
      if (s > 0 && present(3)) {
        d[n++] = id_of(s - 1, 3, (t + H - 1) % H);
        if (H > 1) d[n++] = id_of(s - 1, 3, t);
      }
    } else if (ph == 1) {
      d[n++] = id_of(s, 0, t);
    } else if (ph == 2) {
      if (present(1)) {
        d[n++] = id_of(s, 1, t);
        if (H > 1) d[n++] = id_of(s, 1, (t + 1) % H);
      }
    } else {
      d[n++] = id_of(s, 2, t);
    }
    return n;
  };
  std::vector<long long> cand;
  long long n_present = 0;
  for (int s = 0; s < n_sweeps; ++s)
    for (int ph = 0; ph < nph; ++ph) {
      if (!present(ph)) continue;
      for (int t = 0; t < H; ++t) {
        long long d[2];
        const long long id = id_of(s, ph, t);
        ndep[id] = deps(s, ph, t, d);
        if (ndep[id] == 0) cand.push_back(id);
        ++n_present;
      }
    }
  std::vector<long long> keys(total);
  for (int s = 0; s < n_sweeps; ++s)
    for (int ph = 0; ph < nph; ++ph)
      for (int t = 0; t < H; ++t) keys[id_of(s, ph, t)] = (long long)s * H + t + lag[ph];
  auto key_of = [&](long long id) { return keys[id]; };
  out.clear();
  out.reserve(n_present);
  long long sep = n_present + 1;
  for (long long pos = 0; pos < n_present; ++pos) {
    // ready units: smallest key; none ready: the one that becomes ready first
    size_t best = 0;
    bool best_ready = false;
    long long best_key = 0, best_when = 0;
    for (size_t c = 0; c < cand.size(); ++c) {
      const long long id = cand[c], when = ready[id], key = key_of(id);
      const bool is_ready = when <= pos;
      bool better;
      if (c == 0) better = true;
      else if (is_ready != best_ready) better = is_ready;
      else if (is_ready) better = key < best_key || (key == best_key && id < cand[best]);
      else better = when < best_when || (when == best_when && key < best_key);
      if (better) {
        best = c;
        best_ready = is_ready;
        best_key = key;
        best_when = when;
      }
    }
    const long long id = cand[best];
    cand[best] = cand.back();
    cand.pop_back();
    emit[id] = (int)pos;
    const int t = (int)(id % H), ph = (int)((id / H) % nph), s = (int)(id / ((long long)H * nph));
    out.push_back({s, ph, t});
    {
      long long d[2];
      const int nd = deps(s, ph, t, d);
      for (int q = 0; q < nd; ++q) sep = std::min(sep, pos - (long long)emit[d[q]]);
    }
    // dependents
    auto release = [&](int s2, int ph2, int t2) {
      if (s2 >= n_sweeps || !present(ph2)) return;
      const long long id2 = id_of(s2, ph2, t2);
      ready[id2] = std::max(ready[id2], pos + gap);
      if (--ndep[id2] == 0) cand.push_back(id2);
    };
    if (ph == 0) release(s, 1, t);
    else if (ph == 1) {
      release(s, 2, t);
      if (H > 1) release(s, 2, (t + H - 1) % H);
    } else if (ph == 2) release(s, 3, t);
    else {
      release(s + 1, 0, t);
      if (H > 1) release(s + 1, 0, (t + 1) % H);
    }
  }
  *min_sep = (uint32_t)std::min<long long>(sep, 0x7fffffff);
}

#define CMX_STREAM_MAX_SWEEPS 32  // sweeps per launch (bounds the cached unit lists)

static int stream_list(cmx_state *s, int n_sweeps, int kg, const SweepPlan::StreamList **out) {
  SweepPlan &P = s->plan;
  const Geom &g = s->g;
  const bool push = s->p2p && g.halo;
  const std::pair<int, int> key(n_sweeps, (kg + 1) * 2 + (push ? 1 : 0));
  auto it = P.stream_lists.find(key);
  if (it != P.stream_lists.end()) {
    *out = &it->second;
    return CMX_OK;
  }
  const int H = g.N2 / 2;
  const uint32_t W = (uint32_t)g.N0 / 16, rpw = 32u / W, J = (uint32_t)g.N1 / 2;
  const uint32_t tpu = (J + rpw - 1) / rpw;
  std::vector<HostUnit> order;
  uint32_t min_sep = 0;
  stream_order(H, n_sweeps, kg, (int)P.stream_gap, order, &min_sep);
  std::vector<StreamUnit> units(order.size());
  for (size_t q = 0; q < order.size(); ++q) {
    const HostUnit &u = order[q];
    const int cz = u.ph >> 1, cy = u.ph & 1;
    const int k = 2 * u.t + cz;
    StreamUnit &d = units[q];
    uint32_t flags = 0;
    if (cy == 0) {
      flags = CMX_S16_DEP_LO | CMX_S16_DEP_HI;
      // ghost layers nobody counts up (slabs whose halo is exchanged by the host)
      if (g.halo && !push) {
        if (k == 0) flags &= ~CMX_S16_DEP_LO;
        if (k == g.N2 - 1) flags &= ~CMX_S16_DEP_HI;
      }
      // the other k colour: all of the previous sweep (even layers) / of this sweep (odd)
      d.need_rel = (kg < 0) ? (uint32_t)(2 * u.s + 2 * cz) * tpu : 0u;
      d.base_sel = cz ? 0u : 1u;
    } else {
      flags = CMX_S16_DEP_OWN;
      d.need_rel = (uint32_t)(2 * u.s + 1) * tpu;
      d.base_sel = cz ? 1u : 0u;
    }
    d.kf = (uint32_t)k | flags | ((uint32_t)(cz * 2 + cy) << CMX_S16_COLOUR_SHIFT);
    d.sweep_rel = (uint32_t)u.s;
  }
  SweepPlan::StreamList L;
  L.n_units = (uint32_t)units.size();
  L.min_sep = min_sep;
  CMX_CUDA(cudaMalloc((void **)&L.d_units, sizeof(StreamUnit) * std::max<size_t>(1, units.size())));
  CMX_CUDA(cudaMemcpy(L.d_units, units.data(), sizeof(StreamUnit) * units.size(), cudaMemcpyHostToDevice));
  auto ins = P.stream_lists.emplace(key, L);
  *out = &ins.first->second;
  return CMX_OK;
}

// n_sweeps sweeps (k colour group kg, -1 = whole sweeps) as cooperative launches of the
// streaming kernel on the state's stream
static int sweep_stream(cmx_state *s, uint64_t seed, int64_t first_sweep, int64_t n_sweeps, int kg) {
  SweepPlan &P = s->plan;
  const Geom &g = s->g;
  int rc = pair_tables(s);
  if (rc) return rc;
  const bool accum = (s->sweep_flags & CMX_SWEEP_DE_SUM) != 0;
  const bool slab = s->p2p && g.halo;
  size_t smem = 0;
  StreamKernel kern = stream_kernel(s, accum, slab, &smem);
  S16Args a;
  a.occ = s->d_occ;
  a.g = g;
  a.mask = P.mask;
  a.W = (uint32_t)g.N0 / 16;
  a.logW = 0;
  while ((1u << a.logW) < a.W) ++a.logW;
  a.J = (uint32_t)g.N1 / 2;
  const uint32_t rpw = 32u / a.W;
  a.tpu = (a.J + rpw - 1) / rpw;
  a.gr = P.stream_gr;
  const uint32_t gpu = (a.tpu + a.gr - 1) / a.gr;
  a.div_gpu = make_fastdiv(gpu);
  a.tab24 = P.d_tab24;
  a.thr_lo = P.d_thr_lo;
  a.dEpot = P.d_dEpot;
  a.part_acc = P.d_part_acc;
  a.part_dE = P.d_part_dE;
  a.part_stride = (uint32_t)P.part_blocks;
  a.k0 = (uint32_t)seed;
  a.k1 = (uint32_t)(seed >> 32);
  a.rk = philox_key_schedule(a.k0, a.k1);
  a.k_offset = s->k_offset;
  a.wrap_j = (g.N1 - 1) * g.N0;
  a.layer = g.N0 * g.N1;
  a.wrap_k = (g.N2 - 1) * a.layer;
  a.step_pc = (int32_t)(2u * rpw) * g.N0;
  a.step_j = (int32_t)(2u * rpw);
  a.step_gid = 2u * rpw * a.W;
  a.done = s->d_done;
  a.done_stride = (uint32_t)g.N2 + 2u;
  a.peer_dn = s->peer_occ_dn;
  a.peer_up = s->peer_occ_up;
  a.peer_done_dn = s->peer_done_dn;
  a.peer_done_up = s->peer_done_up;
  a.push = slab ? 1 : 0;
  a.fail = s->d_sig + 3;
  a.n_tab2 = 0;
  a.maps2 = nullptr;
  static const int dbg_env = env_int("CMX_STREAM_DEBUG", 0);
  a.dbg = (uint32_t)dbg_env;
  dim3 grid((unsigned)P.stream_blocks, (unsigned)s->n_replicas);
  for (int64_t done = 0; done < n_sweeps;) {
    const int n = (int)std::min<int64_t>(n_sweeps - done, CMX_STREAM_MAX_SWEEPS);
    const SweepPlan::StreamList *L = nullptr;
    if ((rc = stream_list(s, n, kg, &L))) return rc;
    a.units = L->d_units;
    a.n_groups = L->n_units * gpu;
    // block-combined completion counts need dependent groups at least a block-iteration apart
    a.agg = ((unsigned long long)L->min_sep * gpu >= 8ull && !(a.dbg & 8u)) ? 1u : 0u;
    // dynamic assignment by default (measured +4 % at 512^3, and robust against SMs of
    // unequal speed); CMX_STREAM_TICKET=0: static round robin with block-combined counts
    static const int ticket_env = env_int("CMX_STREAM_TICKET", 1);
    a.ticket = nullptr;
    if (ticket_env) {
      if (!P.d_ticket) CMX_CUDA(cudaMalloc((void **)&P.d_ticket, sizeof(uint32_t) * s->n_replicas));
      CMX_CUDA(cudaMemsetAsync(P.d_ticket, 0, sizeof(uint32_t) * s->n_replicas, s->stream));
      a.ticket = P.d_ticket;
      a.agg = 0;
    }
    a.first_sweep = (unsigned long long)(first_sweep + done);
    a.base[0] = s->done_even;
    a.base[1] = s->done_odd;
    static const int trace_env = env_int("CMX_STREAM_TRACE", 0);
    a.trace = nullptr;
    if (trace_env) {
      CMX_CUDA(cudaMalloc((void **)&a.trace, sizeof(uint32_t) * 2 * L->n_units));
      CMX_CUDA(cudaMemsetAsync(a.trace, 0, sizeof(uint32_t) * 2 * L->n_units, s->stream));
    }
    void *args[1] = {&a};
    CMX_CUDA(cudaLaunchCooperativeKernel((const void *)kern, grid, dim3(256), args, smem, s->stream));
    if (trace_env) {  // diagnostics: which units made their groups wait
      std::vector<uint32_t> tr(2 * (size_t)L->n_units);
      std::vector<StreamUnit> un(L->n_units);
      CMX_CUDA(cudaStreamSynchronize(s->stream));
      CMX_CUDA(cudaMemcpy(tr.data(), a.trace, sizeof(uint32_t) * tr.size(), cudaMemcpyDeviceToHost));
      CMX_CUDA(cudaMemcpy(un.data(), L->d_units, sizeof(StreamUnit) * un.size(), cudaMemcpyDeviceToHost));
      cudaFree(a.trace);
      unsigned long long blocked = 0, polls = 0, by_ph[4] = {0, 0, 0, 0};
      for (uint32_t u = 0; u < L->n_units; ++u) {
        blocked += tr[2 * u];
        polls += tr[2 * u + 1];
        by_ph[(un[u].kf >> CMX_S16_COLOUR_SHIFT) & 3u] += tr[2 * u];
      }
      fprintf(stderr, "[stream trace] units %u groups/unit %u: blocked groups %llu (of %u), polls %llu; by colour %llu %llu %llu %llu\n",
              L->n_units, gpu, blocked, a.n_groups, polls, by_ph[0], by_ph[1], by_ph[2], by_ph[3]);
      if (trace_env > 1)
        for (uint32_t u = 0; u < L->n_units; ++u)
          if (tr[2 * u])
            fprintf(stderr, "  pos %u sweep %u colour %u layer %u: blocked %u polls %u\n", u, un[u].sweep_rel,
                    (un[u].kf >> CMX_S16_COLOUR_SHIFT) & 3u, un[u].kf & 0xFFFFu, tr[2 * u], tr[2 * u + 1]);
    }
    if (kg != 1) s->done_even += 2u * (uint32_t)n * a.tpu;
    if (kg != 0) s->done_odd += 2u * (uint32_t)n * a.tpu;
    done += n;
  }
  return CMX_OK;
}

// A lattice a little larger than L2 (512^3: 134 MB against 126 MB): pin as much of it as the
// device allows as PERSISTING L2 lines for the state's stream; those layers then never go
// back to DRAM between the four colour passes of a sweep, only the remainder streams.
// (Together with the alternating pass direction: measured DRAM traffic per sweep 2.4x ->
// see profiles/.)  No effect on lattices that fit L2 anyway, or on slabs.
static int l2_persist_lattice(cmx_state *s) {
  SweepPlan &P = s->plan;
  if (P.l2_window_set) return CMX_OK;
  P.l2_window_set = true;
  static const int off = env_int("CMX_NO_L2_PERSIST", 0);
  const size_t bytes = (size_t)s->g.rep_stride * s->n_replicas;
  int dev = 0, l2 = 0, max_persist = 0, max_window = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
  cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
  if (off || max_persist <= 0 || bytes <= (size_t)l2 / 2) return CMX_OK;
  static const int frac_env = env_int("CMX_L2_PERSIST_PCT", 100);
  size_t want = std::min<size_t>((size_t)max_persist * (size_t)frac_env / 100, (size_t)max_window);
  want = std::min(want, bytes);
  if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) {
    cudaGetLastError();
    return CMX_OK;
  }
  cudaStreamAttrValue av;
  memset(&av, 0, sizeof(av));
  av.accessPolicyWindow.base_ptr = s->d_occ;
  av.accessPolicyWindow.num_bytes = want;
  av.accessPolicyWindow.hitRatio = 1.0f;
  av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute(s->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
  return CMX_OK;
}

// thin peer-attached slabs: n_sweeps whole sweeps as colour passes in cooperative launches
static int sweep_pass(cmx_state *s, uint64_t seed, int64_t first_sweep, int64_t n_sweeps, int kgroup) {
  SweepPlan &P = s->plan;
  const Geom &g = s->g;
  const bool two = use_pair2(s);
  int rc = two ? pair2_tables(s) : pair_tables(s);
  if (rc) return rc;
  const bool accum = (s->sweep_flags & CMX_SWEEP_DE_SUM) != 0;
  size_t smem = 0;
  PassKernel kern = two ? pass2_kernel(s, accum, &smem) : pass_kernel(s, accum, &smem);
  S16Args a;
  memset(&a, 0, sizeof(a));
  a.occ = s->d_occ;
  a.g = g;
  a.mask = P.mask;
  a.W = (uint32_t)g.N0 / 16;
  a.logW = 0;
  while ((1u << a.logW) < a.W) ++a.logW;
  a.J = (uint32_t)g.N1 / 2;
  const uint32_t rpw = 32u / a.W;
  a.tpu = (a.J + rpw - 1) / rpw;
  a.tab24 = two ? P.d_tab2 : P.d_tab24;
  a.thr_lo = two ? P.d_thr_lo2 : P.d_thr_lo;
  a.dEpot = two ? P.d_dEpot2 : P.d_dEpot;
  a.n_tab2 = two ? (uint32_t)P.n_tab2 : 0u;
  a.maps2 = two ? P.d_maps2 : nullptr;
  a.part_acc = P.d_part_acc;
  a.part_dE = P.d_part_dE;
  a.part_stride = (uint32_t)P.part_blocks;
  a.k0 = (uint32_t)seed;
  a.k1 = (uint32_t)(seed >> 32);
  a.rk = philox_key_schedule(a.k0, a.k1);
  a.k_offset = s->k_offset;
  a.wrap_j = (g.N1 - 1) * g.N0;
  a.layer = g.N0 * g.N1;
  a.wrap_k = (g.N2 - 1) * a.layer;
  a.peer_dn = s->peer_occ_dn;
  a.peer_up = s->peer_occ_up;
  a.push = (s->p2p && g.halo) ? 1 : 0;
  a.fail = s->d_sig + 3;
  PassArgs c;
  c.div_tpu = make_fastdiv(a.tpu);
  c.H = (uint32_t)g.N2 / 2;
  c.kgroup = kgroup;
  c.my_sig = s->d_sig;
  c.peer_sig_dn = s->peer_sig_dn;
  c.peer_sig_up = s->peer_sig_up;
  const uint32_t n_rs = a.tpu * c.H;
  if (!g.halo && (rc = l2_persist_lattice(s))) return rc;
  dim3 grid(std::min<uint32_t>((uint32_t)P.stream_blocks, (n_rs + 7) / 8), (unsigned)s->n_replicas);
  c.dq_k = (grid.x * 8u) / a.tpu;
  c.dq_r = (grid.x * 8u) % a.tpu;
  c.ch_row = a.W;
  c.ch_layer = (uint32_t)g.N1 * a.W;
  c.ch_pair = 2u * c.ch_layer;
  c.ch_wrap_j = (int32_t)(((uint32_t)g.N1 - 1u) * a.W);
  c.ch_wrap_k = (int32_t)(((uint32_t)g.N2 - 1u) * c.ch_layer);
  c.gid_off = ((uint32_t)s->k_offset - (uint32_t)g.halo) * c.ch_layer;
  if (((uint64_t)g.rep_stride >> 4) >= (1ull << 32)) {
    cmx_set_error("colour-pass sweep: a replica exceeds 2^32 16-byte chunks");
    return CMX_ERR_UNSUPPORTED;
  }
  // one barrier counter per replica (independent chains wait for their own blocks only); the
  // ring protocol of peer-attached slabs publishes an epoch for all replicas: one counter
  const int bar_mode = a.push ? 1 : 0;
  const size_t bar_bytes = sizeof(uint32_t) * 32 * (size_t)s->n_replicas;
  if (!P.d_gridbar) CMX_CUDA(cudaMalloc((void **)&P.d_gridbar, bar_bytes));
  if (P.gridbar_mode != bar_mode) {
    CMX_CUDA(cudaMemsetAsync(P.d_gridbar, 0, bar_bytes, s->stream));
    P.gridbar_count = 0;
    P.gridbar_mode = bar_mode;
  }
  c.bar = P.d_gridbar;
  c.bar_stride = bar_mode ? 0u : 32u;
  c.n_blocks = bar_mode ? grid.x * grid.y : grid.x;
  const uint32_t passes_per_sweep = kgroup < 0 ? 4u : 2u;
  // (the barrier counter is compared modulo 2^32: a launch adds less than 2^30 arrivals)
  const int64_t max_sweeps = std::max<int64_t>(1, (int64_t)((1u << 30) / (passes_per_sweep * c.n_blocks)));
  for (int64_t done = 0; done < n_sweeps;) {
    const int64_t n = std::min<int64_t>(n_sweeps - done, std::min<int64_t>(1 << 20, max_sweeps));
    c.n_sweeps = (uint32_t)n;
    c.first_sweep = (unsigned long long)(first_sweep + done);
    c.epoch0 = s->epoch;
    c.bar_base = P.gridbar_count;
    P.gridbar_count += (uint32_t)n * passes_per_sweep * c.n_blocks;
    void *args[2] = {&a, &c};
    CMX_CUDA(cudaLaunchCooperativeKernel((const void *)kern, grid, dim3(256), args, smem, s->stream));
    if (a.push) s->epoch += 2ull * (unsigned long long)n;
    done += n;
  }
  return CMX_OK;
}

// one pass over the colours whose k-colour equals kgroup (or all if < 0): block pair-LUT
// kernel (states with linear rows) or the generic evaluators
static int sweep_once(cmx_state *s, uint64_t seed, int64_t sweep, int kgroup,
                      int k_offset) {
  SweepPlan &P = s->plan;
  const DevTables &T = s->t->d;
  const Geom &g = s->g;
  size_t exs = (size_t)T.n_sublat * T.max_occ * T.max_occ;
  if (use_pair(s)) {
    Pair16Args a;
    int rc = pair_args(s, seed, sweep, k_offset, a);
    if (rc) return rc;
    dim3 grid(P.part_blocks, s->n_replicas);
    const bool fcc = (P.mask == kMaskFcc1NN);
    const bool accum = (s->sweep_flags & CMX_SWEEP_DE_SUM) != 0;
    for (int cz = 0; cz < 2; ++cz) {
      if (kgroup >= 0 && cz != kgroup) continue;
      for (int cy = 0; cy < 2; ++cy) {
        a.cy = cy;
        a.cz = cz;
        a.ctr_hi = ((uint32_t)((uint64_t)sweep >> 32) << 16) | ((uint32_t)(cz * 2 + cy) << 9);
        if (P.nocc == 3) {
          if (pair16_minb() >= 4) launch_pair16<3, 4>(a, grid, s->stream, fcc, accum);
          else launch_pair16<3, 3>(a, grid, s->stream, fcc, accum);
        } else {
          launch_pair16<2, 3>(a, grid, s->stream, fcc, accum);
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
  a.rng16 = P.rng16 ? 1 : 0;
  a.accum = (s->sweep_flags & CMX_SWEEP_DE_SUM) ? 1 : 0;
  dim3 grid(P.part_blocks, s->n_replicas);
  const bool pair_sum = pair_sum_allowed(s);  // (a two-class state reaches this function through cmx_sgc_sweep_kgroup only)
  const bool warp = cmx_use_warp_generic(s) && !use_pair(s) && !pair_sum;
  const size_t stage_bytes = (size_t)(P.stage_max + 1) * 8 * sizeof(double);
  PairSumArgs ps;
  ps.V = P.d_ps_V;
  ps.P0 = P.d_ps_P0;
  ps.act_beg = P.d_act_beg;
  ps.act_n = P.d_act_n;
  for (int q = 0; q < 3; ++q) ps.R[q] = P.ps_range[q];
  ps.fast = (g.rep_stride < (int64_t)0x7FFFFFFFll) ? 1 : 0;
  const size_t ps_bytes = (size_t)P.pk_act_max * (T.max_occ * T.max_occ * T.max_occ * 8 + 20) + 16;
  GenTerms G{P.d_gt_beg, P.d_gt_fbeg, P.d_gt_vi, P.d_act_beg, P.d_act_n, P.d_gt_w, P.d_gt_pk};
  const size_t table_bytes = (cmx_gen_shared_bytes(P.pk_terms_max, P.pk_act_max, T.max_occ) + 15) & ~(size_t)15;
  const bool staged = warp && P.d_gt_pk && table_bytes + stage_bytes <= 160 * 1024;
  const int n_mut = (int)P.mut_points.size();
  GenColours C;
  C.n_mut = n_mut;
  for (int q = 0; q < 8; ++q) C.mut_p[q] = (q < n_mut) ? P.mut_points[q] : 0;
  C.ctr_base = (uint32_t)((uint64_t)sweep >> 32) << 16;
  const int n_col = P.S[0] * P.S[1] * P.S[2] * n_mut;
  if (warp) {
    static bool attr_set = false;
    if (!attr_set) {
      CMX_CUDA(cudaFuncSetAttribute(k_sweep_generic_warp<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      CMX_CUDA(cudaFuncSetAttribute(k_sweep_generic_warp<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      CMX_CUDA(cudaFuncSetAttribute(k_sweep_generic_warp<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
  }
  // all colours of the sweep in one cooperative launch, if the grid can be co-resident
  static const bool no_coop = getenv("CMX_GENERIC_NO_COOP") != nullptr;
  const size_t coop_bytes = (size_t)n_mut * table_bytes + stage_bytes;
  // (pays for small colours only: with the tables of every point position resident fewer
  // blocks fit an SM -- measured on ZrO: 24^3 +26 %, 48^3 -9 %, 96^3 -27 %)
  if (staged && !no_coop && kgroup < 0 && !g.halo && n_mut <= 8 && coop_bytes <= 200 * 1024 && a.items <= 1024) {
    if (P.generic_coop_capacity < 0) {
      int per_sm = 0, dev = 0, sms = 0, can = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_generic_warp<true, true>, 256, coop_bytes) != cudaSuccess)
        per_sm = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, dev);
      P.generic_coop_capacity = can ? per_sm * sms : 0;
    }
    const int cap = P.generic_coop_capacity / std::max(1, s->n_replicas);
    if (cap >= 1) {
      dim3 gc(std::min<uint32_t>(grid.x, (uint32_t)cap), grid.y);
      C.col_begin = 0;
      C.col_end = n_col;
      int stage_max = P.stage_max, tb = (int)table_bytes;
      void *args[5] = {&a, &G, &stage_max, &tb, &C};
      CMX_CUDA(cudaLaunchCooperativeKernel((const void *)k_sweep_generic_warp<true, true>, gc, dim3(256), args,
                                           coop_bytes, s->stream));
      return CMX_OK;
    }
  }
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
          a.ctr_hi = C.ctr_base | (a.rng16 ? ((col & 0xffu) << 8) : (col & 0xffffu));
          C.col_begin = (int)col;
          C.col_end = (int)col + 1;
          if (pair_sum)
            k_sweep_pairsum<<<grid, 256, ps_bytes, s->stream>>>(a, ps);
          else if (staged)
            k_sweep_generic_warp<true, false><<<grid, 256, table_bytes + stage_bytes, s->stream>>>(a, G, P.stage_max, (int)table_bytes, C);
          else if (warp)
            k_sweep_generic_warp<false, false><<<grid, 256, stage_bytes, s->stream>>>(a, G, P.stage_max, 0, C);
          else
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
  int blocks;
  bool coop = use_stream(s) || use_pair2(s);
  if (coop && (P.stream_capacity < 0 || P.stream_blocks == 0)) {
    int rc = stream_geometry(s);
    if (rc == CMX_ERR_UNSUPPORTED && use_pair2(s)) {
      // more replicas than co-resident blocks of the two-class kernel (2 per SM): this state
      // sweeps with the per-neighbor-table kernel -- same random bits, same decisions
      P.pair2 = false;
      coop = false;
    } else if (rc) {
      return rc;
    }
  }
  if (coop) {
    blocks = P.stream_blocks;
  } else {
    uint32_t items;
    if (use_pair(s)) {
      // one block iteration = RB whole rows of one (cy,cz) colour
      uint32_t W = s->g.N0 / 16, RB = 256 / W;
      uint32_t n_rows = (uint32_t)(s->g.N1 / 2) * (uint32_t)(s->g.N2 / 2);
      items = (n_rows + RB - 1) / RB * 256;
    } else {
      items = (uint32_t)(s->g.N0 / P.S[0]) * (s->g.N1 / P.S[1]) * (s->g.N2 / P.S[2]);
      if (cmx_use_warp_generic(s) && !pair_sum_allowed(s))
        items = (uint32_t)std::min<uint64_t>((uint64_t)items * 32u, 0xFFFFFF00u);  // a warp per site
    }
    blocks = sweep_blocks_per_replica(items, s->n_replicas);
  }
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

extern "C" int cmx_state_set_sweep_flags(cmx_state *s, uint32_t flags) {
  if (!s) return invalid("cmx_state_set_sweep_flags: null state");
  if (flags & ~(uint32_t)(CMX_SWEEP_DE_SUM | CMX_SWEEP_FORCE_GENERIC | CMX_SWEEP_THREAD_GENERIC | CMX_SWEEP_STREAM |
                          CMX_SWEEP_PAIR_SUM))
    return invalid("cmx_state_set_sweep_flags: unknown flag");
  s->sweep_flags = flags;
  s->plan.part_blocks = 0;      // the grid may change with the evaluator
  s->plan.stream_capacity = -1;  // and with the kernel variant
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
  unsigned long long timed_out = 0;
  CMX_CUDA(cudaMemcpyAsync(&timed_out, s->d_sig + 3, sizeof(timed_out), cudaMemcpyDeviceToHost, s->stream));
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  if (timed_out) {
    cmx_set_error("cmx_counters_read: a streaming sweep timed out waiting for a layer counter "
                  "(grid not co-resident, or a ring neighbour did not run the same sweeps)");
    return CMX_ERR_CUDA;
  }
  return CMX_OK;
}

extern "C" int cmx_sgc_sweep(cmx_state *s, int64_t n_sweeps, uint64_t seed,
                             int64_t first_sweep, cmx_counters *counters) {
  int rc = cmx_counters_reset(s);
  if (rc) return rc;
  if (n_sweeps < 0) return invalid("cmx_sgc_sweep: n_sweeps < 0");
  if (s->g.halo) return invalid("cmx_sgc_sweep: slab states are driven by cmx_sgc_sweep_kgroup");
  rc = cmx_sgc_sweep_enqueue(s, seed, first_sweep, n_sweeps);
  if (rc) return rc;
  if (counters) return cmx_counters_read(s, counters);
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  return CMX_OK;
}

// n_sweeps whole sweeps enqueued on the state's stream (no reset, no synchronisation):
// the loop body of cmx_sgc_sweep, shared with cmx_sweep_run
int cmx_sgc_sweep_enqueue(cmx_state *s, uint64_t seed, int64_t first_sweep, int64_t n_sweeps) {
  int rc = sweep_prepare(s, "cmx_sgc_sweep");
  if (rc) return rc;
  if (s->g.halo) return invalid("cmx_sgc_sweep: slab states are driven by cmx_sgc_sweep_kgroup / cmx_sgc_sweep_slab");
  if (use_pair2(s)) {
    if (n_sweeps > 0 && (rc = sweep_pass(s, seed, first_sweep, n_sweeps, -1))) return rc;
  } else if (use_stream(s)) {
    if (n_sweeps > 0) {
      rc = use_pass(s) ? sweep_pass(s, seed, first_sweep, n_sweeps, -1) : sweep_stream(s, seed, first_sweep, n_sweeps, -1);
      if (rc) return rc;
    }
  } else {
    for (int64_t w = 0; w < n_sweeps; ++w) {
      rc = sweep_once(s, seed, first_sweep + w, -1, 0);
      if (rc) return rc;
    }
  }
  s->plan.attempts += (long long)s->g.n_cells * (long long)s->plan.mut_points.size() * n_sweeps;
  return CMX_OK;
}

// cmx_sgc_sweep without the synchronisation: counters are reset, n_sweeps sweeps are
// enqueued; read the counters (cmx_counters_read) or call cmx_state_synchronize later
extern "C" int cmx_sgc_sweep_async(cmx_state *s, int64_t n_sweeps, uint64_t seed, int64_t first_sweep) {
  int rc = cmx_counters_reset(s);
  if (rc) return rc;
  if (n_sweeps < 0) return invalid("cmx_sgc_sweep_async: n_sweeps < 0");
  return cmx_sgc_sweep_enqueue(s, seed, first_sweep, n_sweeps);
}

// more sweeps of the same accounting period: no counter reset, no synchronisation
extern "C" int cmx_sgc_sweep_continue(cmx_state *s, int64_t n_sweeps, uint64_t seed, int64_t first_sweep) {
  if (!s) return invalid("cmx_sgc_sweep_continue: null state");
  if (n_sweeps < 0) return invalid("cmx_sgc_sweep_continue: n_sweeps < 0");
  return cmx_sgc_sweep_enqueue(s, seed, first_sweep, n_sweeps);
}

extern "C" int cmx_sgc_sweep_kgroup(cmx_state *s, uint64_t seed, int64_t sweep,
                                    int32_t kgroup) {
  int rc = sweep_prepare(s, "cmx_sgc_sweep_kgroup");
  if (rc) return rc;
  if (kgroup < -1 || kgroup >= s->plan.S[2]) return invalid("cmx_sgc_sweep_kgroup: bad kgroup");
  if (use_pair2(s)) {
    rc = sweep_pass(s, seed, sweep, 1, kgroup);
  } else if (use_stream(s)) {
    if (s->p2p && s->g.halo && kgroup >= 0)
      return invalid("cmx_sgc_sweep_kgroup: peer-attached slabs sweep with cmx_sgc_sweep_slab");
    rc = use_pass(s) ? sweep_pass(s, seed, sweep, 1, kgroup) : sweep_stream(s, seed, sweep, 1, kgroup);
  } else {
    rc = sweep_once(s, seed, sweep, kgroup, s->k_offset);
  }
  if (rc) return rc;
  long long per = (long long)s->g.n_cells * (long long)s->plan.mut_points.size();
  if (kgroup >= 0) per /= s->plan.S[2];
  s->plan.attempts += per;
  return CMX_OK;  // asynchronous: enqueued on the state's stream
}

// Slab states over peer memory: n_sweeps whole sweeps of the streaming kernel; boundary
// rows go to the ring neighbours' ghost layers and layer counters.  Asynchronous.
// CMX_ERR_UNSUPPORTED when the state is not a peer-attached slab on the streaming kernel:
// drive it with cmx_sgc_sweep_kgroup then.
extern "C" int cmx_sgc_sweep_slab(cmx_state *s, int64_t n_sweeps, uint64_t seed, int64_t first_sweep) {
  int rc = sweep_prepare(s, "cmx_sgc_sweep_slab");
  if (rc) return rc;
  if (n_sweeps < 0) return invalid("cmx_sgc_sweep_slab: n_sweeps < 0");
  if (!s->g.halo || !s->p2p || !use_stream(s)) {
    cmx_set_error("cmx_sgc_sweep_slab: not a peer-attached slab state on the streaming kernel");
    return CMX_ERR_UNSUPPORTED;
  }
  if (n_sweeps == 0) return CMX_OK;
  rc = use_pass(s) ? sweep_pass(s, seed, first_sweep, n_sweeps, -1) : sweep_stream(s, seed, first_sweep, n_sweeps, -1);
  if (rc) return rc;
  s->plan.attempts += (long long)s->g.n_cells * (long long)s->plan.mut_points.size() * n_sweeps;
  return CMX_OK;
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
  const char *nm = use_pair(s) ? "pair_lut" : use_pair2(s) ? "pair_lut2" : cmx_use_pair_sum(s) ? "pair_sum" : "generic";
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

extern "C" int cmx_sweep_stream_info(cmx_state *s, int32_t *stream, int32_t *blocks, int32_t *group_rowsteps,
                                     int32_t *gap_units) {
  int rc = sweep_prepare(s, "cmx_sweep_stream_info");
  if (rc) return rc;
  const bool on = use_stream(s) && !use_pass(s);
  if (stream) *stream = on ? 1 : 0;
  if (blocks) *blocks = on ? s->plan.stream_blocks : 0;
  if (group_rowsteps) *group_rowsteps = on ? (int32_t)s->plan.stream_gr : 0;
  if (gap_units) *gap_units = on ? (int32_t)s->plan.stream_gap : 0;
  return CMX_OK;
}

extern "C" int cmx_sweep_launches(const cmx_state *s, int32_t *per_sweep) {
  if (!s || !per_sweep) return invalid("cmx_sweep_launches: null argument");
  if (!s->plan.valid) {
    cmx_set_error("cmx_sweep_launches: no sweep plan");
    return CMX_ERR_STATE;
  }
  *per_sweep = (use_stream(s) || use_pair2(s)) ? 0 : (use_pair(s) ? 4 : s->plan.n_colours);  // 0: one launch per call
  return CMX_OK;
}

extern "C" int cmx_sweep_term_counts(const cmx_state *s, double *terms_per_step, double *neighbors_per_step) {
  if (!s) return invalid("cmx_sweep_term_counts: null state");
  if (!s->plan.valid || s->plan.mut_points.empty()) {
    cmx_set_error("cmx_sweep_term_counts: no sweep plan");
    return CMX_ERR_STATE;
  }
  const double nm = (double)s->plan.mut_points.size();
  if (terms_per_step) *terms_per_step = s->plan.n_gterms / nm;
  if (neighbors_per_step) *neighbors_per_step = s->plan.bytes_per_step - 2.0;  // (bytes = neighbors + own + write)
  return CMX_OK;
}

// ---------------------------------------------------------------------------
// per-proposal dE of the sweep's evaluator (parity/debug entry)
// ---------------------------------------------------------------------------
struct DebugDeArgs {
  const int8_t *occ;  // the replica
  Geom g;
  DevTables T;
  long long n;
  const long long *l;
  const int32_t *new_occ;
  double *out;
  // pair LUT
  int nocc, z;
  int32_t shell[48];
  const double *dEpot;  // [CMX_TAB16] of the replica
  // generic
  const int32_t *gt_beg, *gt_fbeg, *gt_f, *gt_n;
  const double *gt_w, *exch;
  // pair sum
  const double *ps_V, *ps_P0;
  const int32_t *act_beg, *act_n;
  // two-class count table
  int z2;
  int32_t shell2[48];
  const uint16_t *maps2;
};
__device__ __forceinline__ bool debug_site(const DebugDeArgs &a, long long q, int &b, int &i, int &j, int &k) {
  const Geom &g = a.g;
  const long long l = a.l[q];
  if (l < 0 || l >= g.n_cells * a.T.n_sublat) return false;
  b = (int)(l / g.n_cells);
  const long long cell = l - (long long)b * g.n_cells;
  i = (int)(cell % g.N0);
  const long long rest = cell / g.N0;
  j = (int)(rest % g.N1);
  k = (int)(rest / g.N1);
  return true;
}
// mode 0: pair-LUT table entry; 1: folded term lists, one thread per proposal; 2: pair-sum tables;
// 3: entry of the two-class count table (a.dEpot = the replica's dEpot2)
__global__ void k_sweep_debug_de(DebugDeArgs a, int mode) {
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= a.n) return;
  const Geom &g = a.g;
  const DevTables &T = a.T;
  int b, i, j, k;
  if (!debug_site(a, q, b, i, j, k)) {
    a.out[q] = nan("");
    return;
  }
  const int oi = cmx_dec(a.occ[cmx_site_offset(g, b, i, j, k)]), of = a.new_occ[q];
  if (mode == 0) {
    int cnt = 0;
    for (int n = 0; n < a.z; ++n) {
      int ii = i + a.shell[3 * n], jj = j + a.shell[3 * n + 1], kk = k + a.shell[3 * n + 2];
      cmx_wrap_cell(g, ii, jj, kk);
      cnt += (int)(uint8_t)a.occ[cmx_site_offset(g, 0, ii, jj, kk)];  // storage codes: n1 + 18 n2
    }
    int alt = of - oi - 1;
    if (alt < 0) alt += a.nocc;
    a.out[q] = (of == oi || alt >= a.nocc - 1) ? nan("") : a.dEpot[cnt | ((oi | (alt << 2)) << 8)];
    return;
  }
  if (mode == 3) {
    int cnt1 = 0, cnt2 = 0;
    for (int n = 0; n < a.z + a.z2; ++n) {
      const int32_t *o = (n < a.z) ? &a.shell[3 * n] : &a.shell2[3 * (n - a.z)];
      int ii = i + o[0], jj = j + o[1], kk = k + o[2];
      cmx_wrap_cell(g, ii, jj, kk);
      const int code = (int)(uint8_t)a.occ[cmx_site_offset(g, 0, ii, jj, kk)];
      if (n < a.z) cnt1 += code;
      else cnt2 += code;
    }
    int alt = of - oi - 1;
    if (alt < 0) alt += a.nocc;
    const uint32_t sab = (uint32_t)((oi == 2 ? CMX_VA_CODE : oi) | (alt << 2));
    a.out[q] = (of == oi || alt >= a.nocc - 1)
                   ? nan("")
                   : a.dEpot[(uint32_t)a.maps2[cnt1] + (uint32_t)a.maps2[256 + cnt2] + (uint32_t)a.maps2[512 + sab]];
    return;
  }
  const int mo = T.max_occ;
  int p = -1;
  for (int pp = 0; pp < T.n_nlist_sublat; ++pp)
    if (T.nlist_sublat[pp] == b) p = pp;
  double dE = 0.0;
  if (mode == 2) {  // per-neighbor tables of the pair-sum evaluator, in its summation order
    const int ab = a.act_beg[p], ae = a.act_beg[p + 1];
    dE = a.ps_P0[((size_t)p * mo + oi) * mo + of];
    for (int s = ab; s < ae; ++s) {
      const int64_t no = cmx_nbr_offset(T, g, a.act_n[s], i, j, k, nullptr);
      dE += a.ps_V[(((size_t)s * mo + oi) * mo + of) * mo + cmx_dec(a.occ[no])];
    }
    a.out[q] = dE - a.exch[((size_t)b * mo + oi) * mo + of];
    return;
  }
  for (int t = a.gt_beg[p]; t < a.gt_beg[p + 1]; ++t) {
    double v = a.gt_w[((size_t)t * mo + oi) * mo + of];
    for (int f = a.gt_fbeg[t]; f < a.gt_fbeg[t + 1]; ++f) {
      const int n = a.gt_n[f];
      const int64_t no = cmx_nbr_offset(T, g, n, i, j, k, nullptr);
      v *= T.phi[((size_t)T.nbr[n].w * T.n_func + a.gt_f[f]) * mo + cmx_dec(a.occ[no])];
    }
    dE += v;
  }
  a.out[q] = dE - a.exch[((size_t)b * mo + oi) * mo + of];
}
// folded term lists, one WARP per proposal (the evaluator of wide orbit sets)
__global__ void __launch_bounds__(256) k_sweep_debug_de_warp(DebugDeArgs a, GenTerms G, int stage_max) {
  extern __shared__ __align__(16) unsigned char sh_dbg[];
  const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  double *sh_val = reinterpret_cast<double *>(sh_dbg) + (size_t)wib * (stage_max + 1);
  const long long q = blockIdx.x * 8ll + wib;
  if (q >= a.n) return;
  const Geom &g = a.g;
  const DevTables &T = a.T;
  int b, i, j, k;
  if (!debug_site(a, q, b, i, j, k)) {
    if (lane == 0) a.out[q] = nan("");
    return;
  }
  int p = -1;
  for (int pp = 0; pp < T.n_nlist_sublat; ++pp)
    if (T.nlist_sublat[pp] == b) p = pp;
  const int oi = cmx_dec(a.occ[cmx_site_offset(g, b, i, j, k)]), of = a.new_occ[q];
  const double dE = cmx_warp_site_delta<false>(T, g, G, a.occ, sh_val, p, i, j, k, oi, of, -1, 0, lane);
  if (lane == 0) a.out[q] = dE - a.exch[((size_t)b * T.max_occ + oi) * T.max_occ + of];
}

extern "C" int cmx_sweep_debug_delta_e(cmx_state *s, int32_t replica, int64_t n, const int64_t *l,
                                       const int32_t *new_occ, double *out) {
  int rc = sweep_prepare(s, "cmx_sweep_debug_delta_e");
  if (rc) return rc;
  if (replica < 0 || replica >= s->n_replicas || n < 0 || (n && (!l || !new_occ || !out)))
    return invalid("cmx_sweep_debug_delta_e: bad argument");
  if (n == 0) return CMX_OK;
  SweepPlan &P = s->plan;
  const DevTables &T = s->t->d;
  for (int64_t q = 0; q < n; ++q) {
    if (l[q] < 0 || l[q] >= s->g.n_cells * T.n_sublat) return invalid("cmx_sweep_debug_delta_e: site out of range");
    const int b = (int)(l[q] / s->g.n_cells);
    if (new_occ[q] < 0 || new_occ[q] >= s->t->n_occ[b]) return invalid("cmx_sweep_debug_delta_e: occupant out of range");
  }
  const size_t bytes = (size_t)n * (sizeof(long long) + sizeof(int32_t) + sizeof(double));
  if ((rc = cmx_scratch(s, bytes + 64))) return rc;
  double *d_out = (double *)s->d_scratch;
  long long *d_l = (long long *)(d_out + n);
  int32_t *d_new = (int32_t *)(d_l + n);
  CMX_CUDA(cudaMemcpyAsync(d_l, l, sizeof(long long) * n, cudaMemcpyHostToDevice, s->stream));
  CMX_CUDA(cudaMemcpyAsync(d_new, new_occ, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s->stream));
  DebugDeArgs a;
  a.occ = s->d_occ + (size_t)replica * s->g.rep_stride;
  a.g = s->g;
  a.T = T;
  a.n = n;
  a.l = d_l;
  a.new_occ = d_new;
  a.out = d_out;
  a.nocc = P.nocc;
  a.z = std::min(P.z, 16);
  for (int q = 0; q < 48; ++q) a.shell[q] = P.shell[q];
  a.dEpot = P.d_dEpot ? P.d_dEpot + (size_t)replica * P.n_tab : nullptr;
  a.gt_beg = P.d_gt_beg;
  a.gt_fbeg = P.d_gt_fbeg;
  a.gt_f = P.d_gt_f;
  a.gt_n = P.d_gt_n;
  a.gt_w = P.d_gt_w;
  const size_t exs = (size_t)T.n_sublat * T.max_occ * T.max_occ;
  a.exch = s->d_exch + (size_t)replica * exs;
  a.ps_V = P.d_ps_V;
  a.ps_P0 = P.d_ps_P0;
  a.act_beg = P.d_act_beg;
  a.act_n = P.d_act_n;
  a.z2 = P.z2;
  for (int q = 0; q < 48; ++q) a.shell2[q] = P.shell2[q];
  a.maps2 = P.d_maps2;
  if (use_pair2(s)) {
    if ((rc = pair2_tables(s))) return rc;
    a.dEpot = P.d_dEpot2 + (size_t)replica * P.n_tab2;
    k_sweep_debug_de<<<(unsigned)((n + 127) / 128), 128, 0, s->stream>>>(a, 3);
  } else if (cmx_use_pair_sum(s)) {
    k_sweep_debug_de<<<(unsigned)((n + 127) / 128), 128, 0, s->stream>>>(a, 2);
  } else if (use_pair(s)) {
    if (P.z > 16) return invalid("cmx_sweep_debug_delta_e: neighbor class larger than 16");
    if ((rc = pair_tables(s))) return rc;
    k_sweep_debug_de<<<(unsigned)((n + 127) / 128), 128, 0, s->stream>>>(a, 0);
  } else if (cmx_use_warp_generic(s)) {
    GenTerms G{P.d_gt_beg, P.d_gt_fbeg, P.d_gt_vi, P.d_act_beg, P.d_act_n, P.d_gt_w, P.d_gt_pk};
    const size_t stage_bytes = (size_t)(P.stage_max + 1) * 8 * sizeof(double);
    static bool attr_set = false;
    if (!attr_set) {
      CMX_CUDA(cudaFuncSetAttribute(k_sweep_debug_de_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr_set = true;
    }
    k_sweep_debug_de_warp<<<(unsigned)((n + 7) / 8), 256, stage_bytes, s->stream>>>(a, G, P.stage_max);
  } else {
    k_sweep_debug_de<<<(unsigned)((n + 127) / 128), 128, 0, s->stream>>>(a, 1);
  }
  CMX_CUDA(cudaGetLastError());
  CMX_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream));
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  return CMX_OK;
}
