// Warp-row variant of the pair-LUT sweep: k_sweep_row16 (included by cmx_sweep.cu).
//
// Same update rule, same random bits and same decisions as k_sweep_pair16 -- one
// thread owns a 16-site chunk of a row of the launch's (cy,cz) colour and updates
// both x colours of it -- but organised around one WARP per group of whole rows
// (W = N0/16 chunks per row, W a power of two <= 32, 32/W rows per warp):
//
//  * everything a chunk needs from the chunks left and right of it (the bytes
//    just outside the 16-byte window of the neighbor rows, the first byte of the
//    next chunk after the even-lane update) lives in a lane of the same warp and
//    arrives by warp shuffle: no side-word loads, no shared-memory exchange, and
//    NO block barrier inside the loop -- warps run free and overlap each other's
//    load latency;
//  * the Metropolis test runs on two sites per instruction.  A Philox word holds
//    two 16-bit fields [alt:1 | u15:15]; with U = R | 0x80008000 and T = the two
//    thresholds packed the same way (thr <= 0x8000), D = U - T never borrows
//    across the halves and bit 15 of each half of D is "rejected" (u15 >= thr),
//    D_half == 0x8000 is "the 15 bits tie" (min.s16x2 accumulates it).  One byte
//    permute in sign-replicate mode turns D into the byte mask that merges the
//    proposed codes into the chunk; acceptances are counted from the masks with
//    one dot-product instruction per chunk;
//  * the table index takes the storage code as stored (no masking):
//    idx = cnt | (code | alt << 2) << 8, 23 x 256 four-byte entries
//    e = thr16 | proposed code << 16.
//
// The kernel processes the target rows of the colour layers [kk_begin, kk_begin +
// n_tiles * rows_per_warp / J): the host slices a sweep into k-slabs that fit L2
// and orders the colour passes so that every layer is read from DRAM once and
// written once per sweep (see sweep_once).  Launched with programmatic stream
// serialisation: the table load and the address set-up of launch n+1 overlap the
// tail of launch n; griddepcontrol.wait orders the lattice accesses.
#pragma once
#include <cooperative_groups.h>

#define CMX_TAB24(NOCC) ((NOCC) == 3 ? 23 * 256 : 512)

// compact index (k_build_tab16 layout) of a tab24 index
__host__ __device__ __forceinline__ uint32_t cmx_tab24_to_16(uint32_t idx24) {
  const uint32_t sab = idx24 >> 8;
  return (idx24 & 255u) | (((sab & 3u) | (((sab >> 2) & 1u) << 2)) << 8);
}

__global__ void k_build_tab24(const uint32_t *__restrict__ tab16, int nocc, int n_tab16, int n_tab24,
                              uint32_t *__restrict__ tab24) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (idx >= n_tab24) return;
  const uint32_t sab = (uint32_t)idx >> 8;
  const uint32_t code = (nocc == 3) ? (sab & ~4u) : sab;
  uint32_t e = 0;  // thr 0: never accepted, a tie finds thr_lo 0
  if (code == 0 || code == 1 || (nocc == 3 && code == CMX_VA_CODE)) {
    const uint32_t o = tab16[(size_t)r * n_tab16 + cmx_tab24_to_16((uint32_t)idx)];
    // tab16 entry: ((t47 >> 32) << 1 | 1) << 8 | proposed code
    e = ((o >> 9) & 0xFFFFu) | ((o & 0xFFu) << 16);
  }
  tab24[(size_t)r * n_tab24 + idx] = e;
}

// byte permute with the full 4-bit selector nibbles (bit 3: replicate the sign of the
// selected byte); __byte_perm only honours 3 bits
template <uint32_t SEL>
__device__ __forceinline__ uint32_t prmt_s(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "n"(SEL));
  return d;
}
__device__ __forceinline__ uint32_t min_s16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("min.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

// One x colour of one chunk, two sites per step (see the file comment).
//  rej[i]  out: byte mask of word i, 0x00 in the lanes accepted by this colour
template <int CX, int NOCC, bool ACCUM>
__device__ __forceinline__ void row16_update(const uint32_t (&cnt)[4], uint32_t (&C)[4],
                                             const Philox &ph, uint32_t tab, uint32_t (&rej)[4],
                                             uint32_t &tmin, const double *__restrict__ dEpot,
                                             double &e_sum) {
  tmin = 0x7FFF7FFFu;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t R = ph.c[i];
    const uint32_t U = R | 0x80008000u;
    uint32_t SA = C[i];
    if (NOCC == 3) SA |= (R >> (CX ? 5 : 13)) & (CX ? 0x04000400u : 0x00040004u);
    // idx = cnt byte b | SA byte b << 8; bytes 2,3 <- sign of an SA byte (codes < 128: zero)
    constexpr uint32_t sel_a = (uint32_t)CX | ((uint32_t)(4 + CX) << 4) | (0xCu << 8) | (0xCu << 12);
    constexpr uint32_t sel_b = (uint32_t)(CX + 2) | ((uint32_t)(6 + CX) << 4) | (0xCu << 8) | (0xCu << 12);
    const uint32_t ia = prmt_s<sel_a>(cnt[i], SA);
    const uint32_t ib = prmt_s<sel_b>(cnt[i], SA);
    const uint32_t ea = lds_u32(tab + 4u * ia);
    const uint32_t eb = lds_u32(tab + 4u * ib);
    const uint32_t T = __byte_perm(ea, eb, 0x5410u);
    const uint32_t D = U - T;
    tmin = min_s16x2(tmin, D);
    // lanes of this colour <- sign of the matching half of D, other lanes <- 0xFF
    const uint32_t rj = prmt_s<CX ? 0xBD9Du : 0xDBD9u>(D, U);
    const uint32_t P = __byte_perm(ea, eb, CX ? 0x6226u : 0x6622u);
    if (ACCUM) {
      if (!((rj >> (8 * CX)) & 1u)) e_sum += dEpot[cmx_tab24_to_16(ia)];
      if (!((rj >> (8 * (CX + 2))) & 1u)) e_sum += dEpot[cmx_tab24_to_16(ib)];
    }
    C[i] = (C[i] & rj) | (P & ~rj);
    rej[i] = rj;
  }
}

// rare path (2^-15 per site): the 15 bits of a site equal its threshold's; the
// main path left the site rejected.  Draw 32 more bits and finish the 47-bit test.
// Out of line, its operands pass through local memory: the hot loop keeps its
// registers.
struct Row16Tie {
  uint32_t cnt[4], C[4], rej[4], R[4];
  double e_sum;
};
template <int CX, int NOCC, bool ACCUM>
__device__ __noinline__ void row16_ties(Row16Tie *t, uint32_t tab, const uint32_t *__restrict__ thr_lo,
                                        const double *__restrict__ dEpot, uint32_t gid, uint32_t r,
                                        uint32_t sweep_lo, uint32_t ctr, uint32_t k0, uint32_t k1) {
  const Philox lo0 = philox4x32_10(gid, r, sweep_lo, ctr | 1u, k0, k1);
  const Philox lo1 = philox4x32_10(gid, r, sweep_lo, ctr | 2u, k0, k1);
  for (int i = 0; i < 4; ++i) {
    const uint32_t R = t->R[i];
    for (int h = 0; h < 2; ++h) {
      const int b = 2 * h + CX;
      if (!((t->rej[i] >> (8 * b)) & 1u)) continue;  // accepted by the main path: the lane holds the NEW code
      const uint32_t field = h ? (R >> 16) : (R & 0xFFFFu);
      const uint32_t sab = ((t->C[i] >> (8 * b)) & 0xFFu) | ((NOCC == 3) ? ((field >> 15) << 2) : 0u);
      const uint32_t idx = ((t->cnt[i] >> (8 * b)) & 0xFFu) | (sab << 8);
      const uint32_t e = lds_u32(tab + 4u * idx);
      if ((field & 0x7FFFu) != (e & 0xFFFFu)) continue;
      const int q = 2 * i + h;  // target site of the chunk, 0..7
      const uint32_t w32 = (q < 4) ? lo0.c[q & 3] : lo1.c[q & 3];
      const uint32_t i16 = cmx_tab24_to_16(idx);
      if (w32 < thr_lo[i16]) {
        t->C[i] = (t->C[i] & ~(0xFFu << (8 * b))) | (((e >> 16) & 0xFFu) << (8 * b));
        t->rej[i] &= ~(0xFFu << (8 * b));
        if (ACCUM) t->e_sum += dEpot[i16];
      }
    }
  }
}
#define CMX_ROW16_TIES(CX, REJ, CTR)                                                              \
  do {                                                                                            \
    const uint32_t tz = tmin ^ 0x80008000u;                                                       \
    if ((tz - 0x00010001u) & ~tz & 0x80008000u) {                                                 \
      Row16Tie t;                                                                                 \
      _Pragma("unroll") for (int i = 0; i < 4; ++i) {                                             \
        t.cnt[i] = cnt[i];                                                                        \
        t.C[i] = C[i];                                                                            \
        t.rej[i] = REJ[i];                                                                        \
        t.R[i] = ph.c[i];                                                                         \
      }                                                                                           \
      t.e_sum = 0.0;                                                                              \
      row16_ties<CX, NOCC, ACCUM>(&t, tab, a.thr_lo + (size_t)r * NTAB16, dEpot, gid, r, sweep_lo, CTR,   \
                                  a.k0, a.k1);                                                    \
      _Pragma("unroll") for (int i = 0; i < 4; ++i) {                                             \
        C[i] = t.C[i];                                                                            \
        REJ[i] = t.rej[i];                                                                        \
      }                                                                                           \
      if (ACCUM) e_sum += t.e_sum;                                                                \
    }                                                                                             \
  } while (0)

// per-thread constants of the tile body
struct Row16Lane {
  uint32_t tab;          // shared-memory address of the acceptance table
  const double *dEpot;   // [CMX_TAB16] of the replica (ACCUM only)
  const uint32_t *thr_lo;
  int8_t *base;          // replica base (includes the ghost layers)
  uint32_t r, c, lane_l, lane_r, mask, mc;
  bool any_m, any_p;
};

template <bool CG>
__device__ __forceinline__ uint4 row16_load(const int8_t *p) {
  // CG: rows other SMs wrote during the same launch (fused kernel) or during a
  // launch this one overlapped (programmatic dependent launch: a block may start,
  // and its SM's L1 be invalidated, before older launches stopped filling it): L2 only
  return CG ? __ldcg(reinterpret_cast<const uint4 *>(p)) : *reinterpret_cast<const uint4 *>(p);
}

// One tile: every lane updates both x colours of its chunk of row (j, k) and stores
// it (lanes with on == false redo a valid row without storing).  Returns the
// chunk's address.
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}
__device__ __forceinline__ uint4 lds_u128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
#define CMX_ROW16_SLOT 512u  // bytes between the staged rows of a warp (32 lanes x 16 B)

// address of this lane's chunk of row (j, k) and the byte offsets to the rows around it
struct Row16Addr {
  int8_t *pc;
  int32_t dj[3], dk[3];
};
__device__ __forceinline__ Row16Addr row16_addr(const Pair16Args &a, const Row16Lane &L, int32_t j, int32_t k) {
  const Geom &g = a.g;
  const int32_t N0 = g.N0, N1 = g.N1, N2 = g.N2;
  const int32_t layer = N0 * N1;
  const bool halo = g.halo != 0;
  Row16Addr A;
  A.pc = L.base + ((size_t)((uint32_t)(k + g.halo) * (uint32_t)N1 + (uint32_t)j) * (uint32_t)N0 + 16u * L.c);
  A.dj[0] = (j == 0) ? (N1 - 1) * N0 : -N0;
  A.dj[1] = 0;
  A.dj[2] = (j == N1 - 1) ? -(N1 - 1) * N0 : N0;
  A.dk[0] = (!halo && k == 0) ? (N2 - 1) * layer : -layer;
  A.dk[1] = 0;
  A.dk[2] = (!halo && k == N2 - 1) ? -(N2 - 1) * layer : layer;
  return A;
}

// stage the rows tile (j, k) reads into this lane's shared-memory slots (asynchronous
// 16-byte copies, L2 only): slot n = n-th row of the mask in (dz, dy) order
template <uint32_t MASK_CT>
__device__ __forceinline__ void row16_issue(const Pair16Args &a, const Row16Lane &L, int32_t j, int32_t k,
                                            uint32_t slots) {
  const uint32_t mask = MASK_CT ? MASK_CT : L.mask;
  const Row16Addr A = row16_addr(a, L, j, k);
  uint32_t n = 0;
#pragma unroll
  for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const uint32_t m3 = (mask >> ((dz + 1) * 9 + (dy + 1) * 3)) & 7u;
      if (m3 == 0 && !(dz == 0 && dy == 0)) continue;
      cp_async16(slots + n * CMX_ROW16_SLOT, A.pc + (ptrdiff_t)(A.dk[dz + 1] + A.dj[dy + 1]));
      ++n;
    }
  }
}

struct Row16NoHook {
  __device__ __forceinline__ void operator()() const {}
};
// SRC = 0: rows are loaded here (CG: through L2 only); SRC = 1: rows were staged by
// row16_issue into `slots`; after_loads() runs once this lane has read its slots.
// PUSH: slab states also store boundary rows into the ring neighbours' ghost layers
// (compile-time: the mere branch costs the plain kernel 1.5 %)
template <int NOCC, uint32_t MASK_CT, bool ACCUM, bool CG, int SRC = 0, typename Hook = Row16NoHook, bool PUSH = true>
__device__ __forceinline__ int8_t *row16_tile(const Pair16Args &a, const Row16Lane &L, int32_t j, int32_t k,
                                              uint32_t sweep_lo, uint32_t ctr_hi, bool on, uint32_t &n_acc,
                                              double &e_tot, uint32_t slots = 0, Hook after_loads = Hook()) {
  constexpr int NTAB16 = CMX_TAB16(NOCC);
  const Geom &g = a.g;
  const int32_t N0 = g.N0, N1 = g.N1, N2 = g.N2;
  const int32_t layer = N0 * N1;
  const uint32_t mask = MASK_CT ? MASK_CT : L.mask;
  const uint32_t mc = (mask >> 12) & 7u;  // center row: dx = -1 / +1 bits
  const uint32_t tab = L.tab, r = L.r;
  const double *dEpot = L.dEpot;
  double e_sum = 0.0;
  const uint32_t gid = ((uint32_t)(k + a.k_offset) * (uint32_t)N1 + (uint32_t)j) * a.W + L.c;
  const Row16Addr A = row16_addr(a, L, j, k);
  int8_t *pc = A.pc;
  // both colours' random fields up front: two independent 10-round chains interleave
  const Philox ph_c0 = philox4x32_10_rk(gid, r, sweep_lo, ctr_hi, a.rk);
  const Philox ph_c1 = philox4x32_10_rk(gid, r, sweep_lo, ctr_hi | 0x100u, a.rk);
  uint32_t C[4], T[4];
  {
    uint32_t A0[4] = {0, 0, 0, 0}, Am[4] = {0, 0, 0, 0}, Ap[4] = {0, 0, 0, 0};
    uint32_t n_slot = 0;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const uint32_t m3 = (mask >> ((dz + 1) * 9 + (dy + 1) * 3)) & 7u;
        const bool center = (dz == 0 && dy == 0);
        if (m3 == 0 && !center) continue;
        uint4 ch;
        if (SRC == 1) ch = lds_u128(slots + (n_slot++) * CMX_ROW16_SLOT);
        else ch = row16_load<CG>(pc + (ptrdiff_t)(A.dk[dz + 1] + A.dj[dy + 1]));
        if (center) {
          C[0] = ch.x;
          C[1] = ch.y;
          C[2] = ch.z;
          C[3] = ch.w;
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
        }
        if (m3 & 4u) {
          Ap[0] += ch.x;
          Ap[1] += ch.y;
          Ap[2] += ch.z;
          Ap[3] += ch.w;
        }
      }
    }
    after_loads();
    // words just outside the chunk: the summed classes of the neighbor chunks
    const uint32_t sm = L.any_m ? __shfl_sync(0xffffffffu, Am[3], L.lane_l) : 0u;
    const uint32_t sp = L.any_p ? __shfl_sync(0xffffffffu, Ap[0], L.lane_r) : 0u;
    // T[x] = A0[x] + Am[x-1] + Ap[x+1], byte lanes of the 16-byte chunk
    T[0] = A0[0] + __funnelshift_l(sm, Am[0], 8) + __funnelshift_r(Ap[0], Ap[1], 8);
    T[1] = A0[1] + __funnelshift_l(Am[0], Am[1], 8) + __funnelshift_r(Ap[1], Ap[2], 8);
    T[2] = A0[2] + __funnelshift_l(Am[1], Am[2], 8) + __funnelshift_r(Ap[2], Ap[3], 8);
    T[3] = A0[3] + __funnelshift_l(Am[2], Am[3], 8) + __funnelshift_r(Ap[3], sp, 8);
  }
  const uint32_t cl = (mc & 1u) ? __shfl_sync(0xffffffffu, C[3], L.lane_l) : 0u;
  uint32_t rej0[4], rej1[4], cnt[4], tmin;
  // ---- x colour 0: even lanes; same-row neighbors are odd lanes (old values)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    cnt[i] = T[i];
    if (mc & 1u) cnt[i] += __funnelshift_l(i ? C[i - 1] : cl, C[i], 8);
    if (mc & 4u) cnt[i] += __funnelshift_r(C[i], (i < 3) ? C[i + 1] : 0u, 8);
  }
  {
    const uint32_t ctr0 = ctr_hi;
    const Philox &ph = ph_c0;
    row16_update<0, NOCC, ACCUM>(cnt, C, ph, tab, rej0, tmin, dEpot, e_sum);
    CMX_ROW16_TIES(0, rej0, ctr0);
  }
  // ---- x colour 1: odd lanes against the updated even lanes; byte 16 is the
  // (updated) first byte of the next chunk of the row
  const uint32_t nb = (mc & 4u) ? __shfl_sync(0xffffffffu, C[0], L.lane_r) : 0u;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    cnt[i] = T[i];
    if (mc & 1u) cnt[i] += __funnelshift_l(i ? C[i - 1] : 0u, C[i], 8);
    if (mc & 4u) cnt[i] += __funnelshift_r(C[i], (i < 3) ? C[i + 1] : nb, 8);
  }
  {
    const uint32_t ctr1 = ctr_hi | 0x100u;
    const Philox &ph = ph_c1;
    row16_update<1, NOCC, ACCUM>(cnt, C, ph, tab, rej1, tmin, dEpot, e_sum);
    CMX_ROW16_TIES(1, rej1, ctr1);
  }
  if (on) {
    if (ACCUM) e_tot += e_sum;
    *reinterpret_cast<uint4 *>(pc) = make_uint4(C[0], C[1], C[2], C[3]);
    // accepted lanes: 0x00 in rej0 & rej1 (the colours own disjoint lanes)
    const uint32_t A = (~(rej0[0] & rej1[0]) & 0x01010101u) + (~(rej0[1] & rej1[1]) & 0x01010101u) +
                       (~(rej0[2] & rej1[2]) & 0x01010101u) + (~(rej0[3] & rej1[3]) & 0x01010101u);
    n_acc = __dp4a(A, 0x01010101u, n_acc);
    if (PUSH && a.push) {
      // my layer 0 is the lower neighbour's upper ghost, my last layer the
      // upper neighbour's lower ghost (same slab geometry on every rank)
      const uint4 out = make_uint4(C[0], C[1], C[2], C[3]);
      const ptrdiff_t o = pc - a.occ;
      if (k == 0) *reinterpret_cast<uint4 *>(a.peer_dn + o + (ptrdiff_t)N2 * layer) = out;
      if (k == N2 - 1) *reinterpret_cast<uint4 *>(a.peer_up + o - (ptrdiff_t)N2 * layer) = out;
    }
  }
  return pc;
}

template <int NOCC, uint32_t MASK_CT>
__device__ __forceinline__ Row16Lane row16_lane(const Pair16Args &a, const uint32_t *sh_tab) {
  constexpr int NTAB16 = CMX_TAB16(NOCC);
  Row16Lane L;
  L.r = blockIdx.y;
  L.tab = (uint32_t)__cvta_generic_to_shared(sh_tab);
  L.dEpot = a.dEpot + (size_t)L.r * NTAB16;
  L.thr_lo = a.thr_lo + (size_t)L.r * NTAB16;
  L.mask = MASK_CT ? MASK_CT : a.mask;
  L.mc = (L.mask >> 12) & 7u;
  L.base = a.occ + (size_t)L.r * a.g.rep_stride;
  const uint32_t lane = threadIdx.x & 31u, Wm = a.W - 1u;
  L.c = lane & Wm;
  L.lane_l = (lane & ~Wm) | ((L.c - 1u) & Wm);
  L.lane_r = (lane & ~Wm) | ((L.c + 1u) & Wm);
  L.any_m = L.any_p = false;
#pragma unroll
  for (int q = 0; q < 9; ++q) {
    if (q == 4) continue;
    L.any_m |= ((L.mask >> (3 * q)) & 1u) != 0;
    L.any_p |= ((L.mask >> (3 * q)) & 4u) != 0;
  }
  return L;
}

// block reduction of the counters (fixed order -> deterministic) into the block's slot
__device__ __forceinline__ bool row16_reduce(const Pair16Args &a, uint32_t r, uint32_t n_acc, double e_tot,
                                             long long *sh_acc, double *sh_sum) {
  long long n_acc64 = n_acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_acc64 += __shfl_down_sync(0xffffffffu, n_acc64, o);
    e_tot += __shfl_down_sync(0xffffffffu, e_tot, o);
  }
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    sh_acc[wid] = n_acc64;
    sh_sum[wid] = e_tot;
  }
  __syncthreads();
  if (threadIdx.x != 0) return false;
  long long A = 0;
  double E = 0.0;
  for (int w = 0; w < 8; ++w) {
    A += sh_acc[w];
    E += sh_sum[w];
  }
  const size_t slot = (size_t)r * a.part_stride + blockIdx.x;
  a.part_acc[slot] += A;
  a.part_dE[slot] += E;
  return true;
}

// rows a tile reads (center included) for a compile-time mask; 9 when the mask is a runtime value
__host__ __device__ constexpr uint32_t row16_n_slots(uint32_t mask_ct) {
  if (mask_ct == 0) return 9;
  uint32_t n = 0;
  for (int q = 0; q < 9; ++q) n += (q == 4 || ((mask_ct >> (3 * q)) & 7u)) ? 1u : 0u;
  return n;
}
template <int NOCC>
__host__ __device__ constexpr size_t row16_smem_bytes(uint32_t mask_ct) {
  return (size_t)CMX_TAB24(NOCC) * 4 + 8u * row16_n_slots(mask_ct) * CMX_ROW16_SLOT;
}

// the rows of the acceptance table that exist: (code | alt << 2) in {0,1,18,4,5,22}
template <int NOCC>
__device__ __forceinline__ void row16_load_table(uint32_t *sh_tab, const uint32_t *__restrict__ gt) {
  if (NOCC == 3) {
    for (int q = threadIdx.x; q < 6 * 256; q += 256) {
      const int row = q >> 8;
      const int sab = (row < 3 ? 0 : 4) | ((row % 3) == 2 ? CMX_VA_CODE : (row % 3));
      sh_tab[(sab << 8) | (q & 255)] = gt[(sab << 8) | (q & 255)];
    }
  } else {
    for (int q = threadIdx.x; q < CMX_TAB24(NOCC); q += 256) sh_tab[q] = gt[q];
  }
}

// One colour pass (a.cy, a.cz) of this block's share of the rows: the tile loop of
// k_sweep_row16 (and of every pass of k_sweep_row16_coop).
template <int NOCC, uint32_t MASK_CT, bool ACCUM, bool SLAB, bool FULL = false>
__device__ __forceinline__ void row16_pass(const Pair16Args &a, const Row16Lane &L, uint32_t slots, uint32_t lane,
                                           uint32_t wib, uint32_t rl, uint32_t rpw_log, uint32_t &n_acc,
                                           double &e_tot) {
  const uint32_t warp0 = blockIdx.x * 8u + wib, n_warps = gridDim.x * 8u;
  // rows advance by a fixed stride per iteration: decode (kk, jj) once, then step
  uint32_t row = (warp0 << rpw_log) + rl;  // relative to row_begin
  uint32_t kk, jj, kk_last, jj_last, step_k, step_j;
  fastdivmod(min(row, a.n_rows - 1u) + a.row_begin, a.divJ, kk, jj);
  fastdivmod(a.n_rows - 1u + a.row_begin, a.divJ, kk_last, jj_last);
  fastdivmod(n_warps << rpw_log, a.divJ, step_k, step_j);
  // Only the layer next to a ghost layer (k = 0 for cz = 0, k = N2-1 for cz = 1) reads
  // the neighbours' data and writes into their ghost layers.  The layers are visited in
  // rotated order so that this layer comes LAST, and only its tiles wait for the
  // neighbours' epoch: the interior of the slab is updated while the neighbours finish
  // their previous step and their flag travels.
  const int32_t kk_end = (int32_t)((a.row_begin + a.n_rows) / a.J), n_layers = (int32_t)(a.n_rows / a.J);
  // (SLAB is a compile-time switch: the vote and the rotated order cost the plain
  // single-GPU kernel 4.6 % when they are merely predicated off)
  const int32_t rot = (SLAB && a.push && a.cz == 0) ? 1 : 0;
  auto wait_neighbours = [&](int32_t k) {
    if (!SLAB) return;
    const bool need = a.wait_epoch && (k == 0 || k == a.g.N2 - 1);
    if (!__any_sync(0xffffffffu, need)) return;
    if (lane == 0) {
      // acquire: the neighbours' pushes into my ghost layers precede their flag
      const long long t0 = clock64();
      while (ld_sys(a.my_sig + 0) < a.wait_epoch || ld_sys(a.my_sig + 1) < a.wait_epoch) {
        if (clock64() - t0 > 8000000000ll) {  // ~4 s: a neighbour is gone
          a.my_sig[3] = 1ull;
          break;
        }
        __nanosleep(100);
      }
    }
    __syncwarp();
  };
  // (j, k, on) of the tile at the current position, then advance the position
  auto take = [&](int32_t &j, int32_t &k, bool &on) {
    // a partial last tile: the idle lanes redo the last row, unstored (FULL: the host saw
    // that the rows divide evenly into tiles -- the predicate and its selects are gone, +1.2 %)
    on = FULL ? true : (row < a.n_rows);
    j = 2 * (int32_t)(on ? jj : jj_last) + a.cy;
    if (SLAB) {
      int32_t kq = (int32_t)(on ? kk : kk_last) + rot;
      if (kq >= kk_end) kq -= n_layers;
      k = 2 * kq + a.cz;
    } else {
      k = 2 * (int32_t)(on ? kk : kk_last) + a.cz;
    }
    row += n_warps << rpw_log;
    jj += step_j;
    kk += step_k;
    if (jj >= a.J) {
      jj -= a.J;
      kk += 1;
    }
  };
  int32_t j = 0, k = 0, jn = 0, kn = 0;
  bool on = false, on_n = false;
  uint32_t tile = warp0;
  if (tile < a.n_tiles) {
    take(j, k, on);
    wait_neighbours(k);
    row16_issue<MASK_CT>(a, L, j, k, slots);
  }
  for (; tile < a.n_tiles; tile += n_warps) {
    const bool more = tile + n_warps < a.n_tiles;
    if (more) take(jn, kn, on_n);
    cp_async_wait_all();
    auto stage_next = [&]() {
      if (more) {
        wait_neighbours(kn);
        row16_issue<MASK_CT>(a, L, jn, kn, slots);
      }
    };
    row16_tile<NOCC, MASK_CT, ACCUM, true, 1, decltype(stage_next), SLAB>(a, L, j, k, a.sweep_lo, a.ctr_hi, on, n_acc, e_tot,
                                                                         slots, stage_next);
    j = jn;
    k = kn;
    on = on_n;
  }
}

// ---- one colour pass (cy,cz) over the colour layers [row_begin/J, ...) --------------
// Software pipelined: while a tile is updated, the rows of the warp's NEXT tile are
// already on their way into the warp's shared-memory slots (cp.async, 16 B per lane
// and row) -- the global latency is hidden behind ~350 instructions of compute and
// costs no registers.  A lane only ever reads the slots it filled itself: no
// barrier, not even a warp one.
template <int NOCC, uint32_t MASK_CT, bool ACCUM, bool SLAB, bool FULL>
__global__ void __launch_bounds__(256, 3) k_sweep_row16(Pair16Args a) {
  constexpr int NTAB = CMX_TAB24(NOCC);
  constexpr uint32_t NSLOT = row16_n_slots(MASK_CT);
  // dynamic shared memory: [acceptance table][8 warps x NSLOT row slots x (32 lanes x 16 B)]
  extern __shared__ __align__(16) unsigned char sh_dyn[];
  uint32_t *sh_tab = reinterpret_cast<uint32_t *>(sh_dyn);
  unsigned char *sh_rows = sh_dyn + NTAB * 4;
  __shared__ long long sh_acc[8];
  __shared__ double sh_sum[8];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  row16_load_table<NOCC>(sh_tab, a.tab24 + (size_t)blockIdx.y * NTAB);
  const Row16Lane L = row16_lane<NOCC, MASK_CT>(a, sh_tab);
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  const uint32_t rl = lane >> a.logW;
  const uint32_t rpw_log = 5u - a.logW;  // log2(rows per warp)
  const uint32_t slots = (uint32_t)__cvta_generic_to_shared(sh_rows) + wib * (NSLOT * CMX_ROW16_SLOT) + 16u * lane;
  uint32_t n_acc = 0;
  double e_tot = 0.0;
  // the lattice may only be touched once the previous launch has completed
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // Slab ring protocol (see sweep_once).  The previous launches of this rank are complete,
  // their stores into the neighbours' ghost layers included: publish the epoch they
  // reached.  Publishing here instead of at the end of the previous launch keeps
  // system-scope fences and remote round trips out of every launch's tail.
  if (SLAB && a.signal_epoch && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    __threadfence_system();
    st_sys(a.peer_sig_dn + 1, a.signal_epoch);  // I am their upper neighbour
    st_sys(a.peer_sig_up + 0, a.signal_epoch);  // I am their lower neighbour
  }
  __syncthreads();
  row16_pass<NOCC, MASK_CT, ACCUM, SLAB, FULL>(a, L, slots, lane, wib, rl, rpw_log, n_acc, e_tot);
  row16_reduce(a, L.r, n_acc, e_tot, sh_acc, sh_sum);
}

// ---- whole sweeps in one cooperative launch, grid barriers between colour passes ------
// For launches too short to amortise their fixed cost (slabs of a domain-decomposed
// supercell: a 64-layer slab is 5 us of work per colour pass against ~5 us of launch,
// ramp and tail): n_sweeps x 4 passes in ONE launch, cooperative-groups grid barrier
// between passes.  The slab ring protocol lives inside: a k-colour group waits for the
// neighbours' epoch in its boundary tiles only (visited last), and the epoch a rank has
// completed is published by block 0 right after the barrier that ends the group.
struct CoopArgs {
  uint32_t n_sweeps;
  unsigned long long first_sweep;
  unsigned long long epoch0;  // k-colour groups this rank completed before the launch
};
template <int NOCC, uint32_t MASK_CT, bool ACCUM, bool SLAB>
__global__ void __launch_bounds__(256, 3) k_sweep_row16_coop(Pair16Args a, CoopArgs c) {
  constexpr int NTAB = CMX_TAB24(NOCC);
  constexpr uint32_t NSLOT = row16_n_slots(MASK_CT);
  extern __shared__ __align__(16) unsigned char sh_dyn[];
  uint32_t *sh_tab = reinterpret_cast<uint32_t *>(sh_dyn);
  unsigned char *sh_rows = sh_dyn + NTAB * 4;
  __shared__ long long sh_acc[8];
  __shared__ double sh_sum[8];
  namespace cgr = cooperative_groups;
  cgr::grid_group grid = cgr::this_grid();
  row16_load_table<NOCC>(sh_tab, a.tab24 + (size_t)blockIdx.y * NTAB);
  const Row16Lane L = row16_lane<NOCC, MASK_CT>(a, sh_tab);
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  const uint32_t rl = lane >> a.logW;
  const uint32_t rpw_log = 5u - a.logW;
  const uint32_t slots = (uint32_t)__cvta_generic_to_shared(sh_rows) + wib * (NSLOT * CMX_ROW16_SLOT) + 16u * lane;
  uint32_t n_acc = 0;
  double e_tot = 0.0;
  __syncthreads();
  unsigned long long epoch = c.epoch0;
  const bool publisher = SLAB && a.push && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  if (publisher && epoch) {  // what the previous launches completed (idempotent)
    __threadfence_system();
    st_sys(a.peer_sig_dn + 1, epoch);
    st_sys(a.peer_sig_up + 0, epoch);
  }
  for (uint32_t s = 0; s < c.n_sweeps; ++s) {
    const unsigned long long sweep = c.first_sweep + s;
    a.sweep_lo = (uint32_t)sweep;
    for (int cz = 0; cz < 2; ++cz) {
      for (int cy = 0; cy < 2; ++cy) {
        a.cy = cy;
        a.cz = cz;
        a.ctr_hi = ((uint32_t)(sweep >> 32) << 16) | ((uint32_t)(cz * 2 + cy) << 9);
        a.wait_epoch = (SLAB && a.push) ? epoch : 0ull;
        a.signal_epoch = 0;
        row16_pass<NOCC, MASK_CT, ACCUM, SLAB>(a, L, slots, lane, wib, rl, rpw_log, n_acc, e_tot);
        grid.sync();  // every store of the pass, the ones into the neighbours' ghost layers included
      }
      ++epoch;
      if (publisher) {
        __threadfence_system();
        st_sys(a.peer_sig_dn + 1, epoch);
        st_sys(a.peer_sig_up + 0, epoch);
      }
    }
  }
  row16_reduce(a, L.r, n_acc, e_tot, sh_acc, sh_sum);
}

// ---- replica grids: one flat tile space ---------------------------------------------
// With one grid row per replica (blockIdx.y) every replica gets the same whole number of
// blocks: 64 replicas on 444 co-resident blocks use 6 x 64 = 384 of them (86 %).  Here the
// tiles of all replicas form ONE index space, cut into gridDim.x equal contiguous ranges;
// a range touches at most two replicas (the host checks), whose acceptance tables both
// sit in shared memory, and a tile looks up its replica.  Same tiles, same counters (a
// block adds to the slots of the replicas it touched), same random numbers.
template <int NOCC, uint32_t MASK_CT, bool ACCUM>
__global__ void __launch_bounds__(256, 3) k_sweep_row16_flat(Pair16Args a, uint32_t n_replicas) {
  constexpr int NTAB = CMX_TAB24(NOCC);
  constexpr int NTAB16 = CMX_TAB16(NOCC);
  constexpr uint32_t NSLOT = row16_n_slots(MASK_CT);
  // dynamic shared memory: [2 acceptance tables][8 warps x NSLOT row slots x (32 lanes x 16 B)]
  extern __shared__ __align__(16) unsigned char sh_dyn[];
  uint32_t *sh_tab = reinterpret_cast<uint32_t *>(sh_dyn);
  unsigned char *sh_rows = sh_dyn + 2 * NTAB * 4;
  __shared__ long long sh_acc[8];
  __shared__ double sh_sum[8];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const unsigned long long T = (unsigned long long)a.n_tiles * n_replicas;
  const uint32_t tb = (uint32_t)(T * blockIdx.x / gridDim.x), te = (uint32_t)(T * (blockIdx.x + 1) / gridDim.x);
  const uint32_t r_lo = tb / a.n_tiles, r_hi = (te > tb) ? (te - 1u) / a.n_tiles : r_lo;  // r_hi <= r_lo + 1
  row16_load_table<NOCC>(sh_tab, a.tab24 + (size_t)r_lo * NTAB);
  if (r_hi != r_lo) row16_load_table<NOCC>(sh_tab + NTAB, a.tab24 + (size_t)r_hi * NTAB);
  Row16Lane L = row16_lane<NOCC, MASK_CT>(a, sh_tab);
  const uint32_t tab0 = L.tab;
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  const uint32_t rl = lane >> a.logW;
  const uint32_t rpw_log = 5u - a.logW;
  const uint32_t slots = (uint32_t)__cvta_generic_to_shared(sh_rows) + wib * (NSLOT * CMX_ROW16_SLOT) + 16u * lane;
  uint32_t n_acc0 = 0, n_acc1 = 0;
  double e_tot0 = 0.0, e_tot1 = 0.0;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __syncthreads();
  // replica, (j, k) and `on` of tile t
  auto locate = [&](uint32_t t, uint32_t &rep, int32_t &j, int32_t &k, bool &on) {
    rep = (r_hi != r_lo && t >= r_hi * a.n_tiles) ? r_hi : r_lo;
    const uint32_t row = ((t - rep * a.n_tiles) << rpw_log) + rl;
    on = row < a.n_rows;  // a partial last tile: the idle lanes redo the last row, unstored
    uint32_t kk, jj;
    fastdivmod(min(row, a.n_rows - 1u) + a.row_begin, a.divJ, kk, jj);
    j = 2 * (int32_t)jj + a.cy;
    k = 2 * (int32_t)kk + a.cz;
  };
  auto set_replica = [&](Row16Lane &X, uint32_t rep) {
    X.r = rep;
    X.tab = tab0 + ((rep != r_lo) ? (uint32_t)(NTAB * 4) : 0u);
    X.base = a.occ + (size_t)rep * a.g.rep_stride;
    X.dEpot = a.dEpot + (size_t)rep * NTAB16;
    X.thr_lo = a.thr_lo + (size_t)rep * NTAB16;
  };
  uint32_t rep = r_lo, rep_n = r_lo;
  int32_t j = 0, k = 0, jn = 0, kn = 0;
  bool on = false, on_n = false;
  uint32_t tile = tb + wib;
  if (tile < te) {
    locate(tile, rep, j, k, on);
    set_replica(L, rep);
    row16_issue<MASK_CT>(a, L, j, k, slots);
  }
  for (; tile < te; tile += 8u) {
    const bool more = tile + 8u < te;
    Row16Lane Ln = L;
    if (more) {
      locate(tile + 8u, rep_n, jn, kn, on_n);
      set_replica(Ln, rep_n);
    }
    cp_async_wait_all();
    auto stage_next = [&]() {
      if (more) row16_issue<MASK_CT>(a, Ln, jn, kn, slots);
    };
    uint32_t n_acc = 0;
    double e_tot = 0.0;
    row16_tile<NOCC, MASK_CT, ACCUM, true, 1, decltype(stage_next), false>(a, L, j, k, a.sweep_lo, a.ctr_hi, on, n_acc, e_tot,
                                                                          slots, stage_next);
    if (rep == r_lo) {
      n_acc0 += n_acc;
      e_tot0 += e_tot;
    } else {
      n_acc1 += n_acc;
      e_tot1 += e_tot;
    }
    L = Ln;
    rep = rep_n;
    j = jn;
    k = kn;
    on = on_n;
  }
  row16_reduce(a, r_lo, n_acc0, e_tot0, sh_acc, sh_sum);
  if (r_hi != r_lo) {
    __syncthreads();
    row16_reduce(a, r_hi, n_acc1, e_tot1, sh_acc, sh_sum);
  }
}

// ---- whole sweeps in ONE launch ---------------------------------------------------
// Persistent co-resident grid (cooperative launch); n_sweeps sweeps of a periodic
// (halo-free) lattice.  Tiles are enumerated in the k-sliced colour order
//   slice q:  even layers [qL, (q+1)L) row colour 0, then 1;
//             odd  layers [qL-1, (q+1)L-1) row colour 0, then 1      (q = 0 .. n_s)
// (slots whose layer falls outside [0, n_kk) are skipped) and dealt round-robin to
// the warps.  Instead of kernel boundaries between colour passes every row carries a
// stamp = number of updates it has received: a tile starts when each neighbor row
// shows the stamp the colour order requires (row colour q' earlier than mine in the
// sweep: one more update than rows of later colours), and publishes its own rows'
// stamps after its stores are visible.  The enumeration is a topological order of
// these dependencies, so the round-robin deal cannot deadlock while all warps are
// resident.  A slice spans ~1.5x the tiles the grid runs concurrently: dependencies
// are normally long satisfied, and the 2L+1 layers a slice touches stay in L2, so a
// sweep reads every layer from DRAM once and writes it once.
struct FusedArgs {
  uint32_t *stamps;      // [replica][N2][N1]
  uint32_t stamp_base;   // every row's stamp when the call starts
  uint32_t n_sweeps;
  uint64_t first_sweep;
  uint32_t L, n_kk, n_slices;  // colour layers per slice, N2/2, slices incl. the closing odd one
  uint32_t tpl;                // tiles per (layer, row colour)
  uint32_t n_tiles_sweep;      // n_slices * 4 * L * tpl slots
  unsigned long long *timeout; // set when a dependency never arrives
};

__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu(uint32_t *p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int NOCC, uint32_t MASK_CT, bool ACCUM>
__global__ void __launch_bounds__(256, 4) k_sweep_row16_fused(Pair16Args a, FusedArgs f) {
  constexpr int NTAB = CMX_TAB24(NOCC);
  __shared__ __align__(16) uint32_t sh_tab[NTAB];
  __shared__ long long sh_acc[8];
  __shared__ double sh_sum[8];
  row16_load_table<NOCC>(sh_tab, a.tab24 + (size_t)blockIdx.y * NTAB);
  const Row16Lane L = row16_lane<NOCC, MASK_CT>(a, sh_tab);
  const uint32_t mask = MASK_CT ? MASK_CT : a.mask;
  const int32_t N1 = a.g.N1, N2 = a.g.N2;
  uint32_t *stamps = f.stamps + (size_t)blockIdx.y * (size_t)N1 * (size_t)N2;
  const uint32_t rl = (threadIdx.x & 31u) >> a.logW;
  const uint32_t rpw = 32u >> a.logW;
  uint32_t n_acc = 0;
  double e_tot = 0.0;
  __syncthreads();
  const uint32_t warp0 = blockIdx.x * 8u + (threadIdx.x >> 5), n_warps = gridDim.x * 8u;
  const uint32_t Tp = f.L * f.tpl;  // slots per phase of a slice
  // position of the warp's current slot: sweep t, slice q, phase ph, offset w in the phase
  uint32_t t = 0, q = 0, ph = 0, w = warp0;
  const FastDiv divTpl = make_fastdiv(f.tpl);
  bool dead = false;
  for (;;) {
    while (w >= Tp) {
      w -= Tp;
      if (++ph == 4) {
        ph = 0;
        if (++q == f.n_slices) {
          q = 0;
          ++t;
        }
      }
    }
    if (t >= f.n_sweeps) break;
    const uint32_t cz = ph >> 1, cy = ph & 1u;
    uint32_t kk_rel, tl;
    fastdivmod(w, divTpl, kk_rel, tl);
    const int32_t kk = (int32_t)(q * f.L + kk_rel) - (int32_t)cz;
    w += n_warps;
    if (kk < 0 || kk >= (int32_t)f.n_kk) continue;  // slot outside the lattice
    const uint32_t jj = tl * rpw + rl;
    const bool on = jj < a.J;
    const int32_t j = 2 * (int32_t)(on ? jj : a.J - 1u) + (int32_t)cy, k = 2 * kk + (int32_t)cz;
    // ---- wait for the neighbor rows: stamp >= base + t + (its colour precedes mine)
    {
      const uint32_t need_y = f.stamp_base + t + cy;  // rows j+-1 of my layer (other row colour)
      const uint32_t need_z = f.stamp_base + t + cz;  // rows of layers k+-1 (other layer colour)
      const int32_t jm = (j == 0) ? N1 - 1 : j - 1, jp = (j == N1 - 1) ? 0 : j + 1;
      const int32_t km = (k == 0) ? N2 - 1 : k - 1, kp = (k == N2 - 1) ? 0 : k + 1;
      const int32_t js[3] = {jm, j, jp}, ks[3] = {km, k, kp};
      long long t0 = 0;
      for (;;) {
        bool ok = true;
#pragma unroll
        for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy) {
            if ((dz == 0 && dy == 0) || !((mask >> ((dz + 1) * 9 + (dy + 1) * 3)) & 7u)) continue;
            const uint32_t v = ld_relaxed_gpu(stamps + (size_t)ks[dz + 1] * N1 + js[dy + 1]);
            ok &= (int32_t)(v - (dz == 0 ? need_y : need_z)) >= 0;
          }
        }
        if (__all_sync(0xffffffffu, ok)) break;
        if (t0 == 0) t0 = clock64();
        if (clock64() - t0 > 20000000000ll) {  // ~10 s: the grid is not co-resident
          *f.timeout = 1ull;
          dead = true;
          break;
        }
        __nanosleep(200);
      }
      if (dead) break;
    }
    const uint64_t sweep = f.first_sweep + t;
    const uint32_t ctr_hi = ((uint32_t)(sweep >> 32) << 16) | ((cz * 2u + cy) << 9);
    row16_tile<NOCC, MASK_CT, ACCUM, true>(a, L, j, k, (uint32_t)sweep, ctr_hi, on, n_acc, e_tot);
    // ---- publish: the stores of every lane, then the stamps of the tile's rows
    __threadfence();
    __syncwarp();
    if (on && L.c == 0) st_relaxed_gpu(stamps + (size_t)k * N1 + j, f.stamp_base + t + 1u);
  }
  row16_reduce(a, L.r, n_acc, e_tot, sh_acc, sh_sum);
}
