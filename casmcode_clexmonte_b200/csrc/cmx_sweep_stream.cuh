// Streaming pair-LUT sweep: k_sweep_stream16 (included by cmx_sweep.cu).
//
// The semi-grand checkerboard update of cmx_sweep.cu for models whose dE depends only on
// (occupant, proposed occupant, neighbor species counts), on x4-interleaved rows
// (Geom::xq_log): the 32-bit word w of a row holds the sites w, w+Q, w+2Q, w+3Q
// (Q = N0/4), so
//   * the x neighbors of a whole word are the words left and right of it: the byte-lane
//     neighbor sums need no byte shifts, only word adds (and one shuffle per row end);
//   * the four sites of a word share their x colour: every instruction of the Metropolis
//     test works on four (compare: two) useful sites.
// One lane owns a 16-byte chunk (words 4c..4c+3 of a row), a warp owns 32/W whole rows
// ("row-step", W = N0/16 chunks per row), and updates the even words (x colour 0), then
// the odd words against the updated even ones.
//
// A whole call (n sweeps x 4 row colours x all layers) is ONE cooperative launch with no
// grid barrier.  The host lists the (layer, row colour) UNITS of the call in an order that
// (i) is a topological order of the colour dependencies and (ii) walks the lattice as a
// narrow wavefront: layer 2t is updated (row colour 0, then 1) a few steps before layer
// 2t-1 (row colour 0, then 1), so that all four colour passes of a sweep touch a layer
// while it is still in L2 -- every byte of the lattice crosses HBM once per sweep in each
// direction instead of four times.  Groups of row-steps are dealt round robin to the
// persistent warps; a per-layer counter of finished row-steps (release/acquire) tells a
// group when the units it depends on are complete.  Dependencies are normally satisfied
// long before they are checked (the host keeps dependent units further apart than the
// warps in flight), so the check is one acquire load per group.
//
// Slab decomposition (SLAB): the rows of the layers next to a ghost layer are also stored
// into the ring neighbour's ghost layer over NVLink peer memory, and the group's count is
// added to the neighbour's ghost-layer counter: the same dependency test covers the halo
// exchange -- no collective, no barrier, no host round trip.
#pragma once
#include <cooperative_groups.h>
#include <type_traits>

#define CMX_TAB24(NOCC) ((NOCC) == 3 ? 23 * 256 : 512)

// tab24 index of (cnt, sab = code | alt << 2): the column is SKEWED by sab.  Sites with equal
// neighbor counts and different (occupant, proposal) are common in a warp; unskewed they
// would read different words of the same shared-memory bank (rows are 256 words apart).
// cnt + sab <= 18 * 12 + 22 < 256 for ternary models (cmx_plan_sweep: z <= 12).
__host__ __device__ __forceinline__ uint32_t cmx_tab24_index(uint32_t cnt, uint32_t sab) {
  return ((cnt + sab) & 255u) | (sab << 8);
}
// compact index (k_build_tab16 layout) of a tab24 index
__host__ __device__ __forceinline__ uint32_t cmx_tab24_to_16(uint32_t idx24) {
  const uint32_t sab = idx24 >> 8;
  return ((idx24 - sab) & 255u) | (((sab & 3u) | (((sab >> 2) & 1u) << 2)) << 8);
}

// tab24[cmx_tab24_index(cnt, code | alt << 2)] = thr16 | proposed code << 16, thr16 in [0, 0x8000]
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
// Poll of a layer counter.  Relaxed (no fence, no L1 invalidation -- an acquire load costs
// a CCTL.IVALL and was half of the kernel's stall samples): what the counter guards is read
// afterwards through L2 only (cp.async.cg / LDGSTS.BYPASS never allocates in L1, so there
// is no stale line to invalidate), and the reads are control dependent on the polled value.
template <bool SYS>
__device__ __forceinline__ uint32_t ld_poll_u32(const uint32_t *p) {
  uint32_t v;
  if (SYS) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  else asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_relaxed_add_u32(uint32_t *p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_sys_add_u32(uint32_t *p, uint32_t v) {
  asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// release at gpu / system scope: MEMBAR + ERRBAR, then the reduction (no L1 invalidation)
template <bool SYS>
__device__ __forceinline__ void red_release_add_u32(uint32_t *p, uint32_t v) {
  if (SYS) asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  else asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// arrival on a shared counter, release + acquire at CTA scope (MEMBAR.ALL.CTA + ATOMS)
__device__ __forceinline__ uint32_t atom_acq_rel_cta_add(uint32_t *sh, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], %2;"
               : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(sh)), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void st_release_cta(uint32_t *sh, uint32_t v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(sh)), "r"(v) : "memory");
}
#ifndef CMX_S16_MINB
#define CMX_S16_MINB 3  // resident blocks per SM the kernel is compiled for
#endif
#ifndef CMX_S16_POLL_NS
#define CMX_S16_POLL_NS 512  // sleep between two polls of a blocked group
#endif
#define CMX_S16_SLOT 512u  // bytes between the staged rows of a warp (32 lanes x 16 B)

// One (layer, row colour) of one sweep of a call.  kf = local layer | flags:
//   bit 16 / 17: wait for the layer below / above (row colour 0: the other k colour must
//   have finished what precedes this unit); bit 18: wait for the own layer (row colour 1
//   follows row colour 0); bits 19-20: the colour 2 cz + cy.  Every awaited counter must
//   have reached base[base_sel] + need_rel finished row-steps (wrap-around compare), base =
//   the counters of the even / odd layers when the call starts.  The list of a call depends
//   only on (sweeps, k colour group, geometry): it is built once and cached.
struct StreamUnit {
  uint32_t kf, sweep_rel, need_rel, base_sel;
};
#define CMX_S16_DEP_LO 0x10000u
#define CMX_S16_DEP_HI 0x20000u
#define CMX_S16_DEP_OWN 0x40000u
#define CMX_S16_COLOUR_SHIFT 19

struct S16Args {
  int8_t *occ;  // replica 0 base (start of the low ghost layers)
  Geom g;
  uint32_t mask;       // runtime neighbor mask, bit (dz+1)*9 + (dy+1)*3 + (dx+1)
  uint32_t W, logW;    // 16-byte chunks per row
  uint32_t J;          // rows of one row colour per layer (N1 / 2)
  uint32_t tpu;        // row-steps per unit: ceil(J / (32 / W))
  uint32_t gr;         // row-steps per group
  FastDiv div_gpu;     // groups per unit: ceil(tpu / gr)
  uint32_t n_groups;   // n_units * groups per unit
  const StreamUnit *units;
  unsigned long long first_sweep;  // RNG counter of the call's first sweep
  uint32_t base[2];                // counters of the even / odd layers when the call starts
  const uint32_t *tab24;   // [replica][CMX_TAB24]
  const uint32_t *thr_lo;  // [replica][CMX_TAB16]
  const double *dEpot;     // [replica][CMX_TAB16]
  long long *part_acc;     // [replica][part_stride] accepted steps
  double *part_dE;
  uint32_t part_stride;
  PhiloxKeys rk;           // round keys of the seed (constant bank)
  uint32_t k0, k1;         // the seed (tie-break draws)
  int32_t k_offset;        // global k of local layer 0 (slab decomposition)
  int32_t wrap_j;          // (N1 - 1) * N0: byte offset from row 0 to row N1 - 1 of a layer
  int32_t layer, wrap_k;   // N0 * N1; (N2 - 1) * layer
  int32_t step_pc, step_j; // bytes / rows between consecutive row-steps of a unit: 2 * rpw rows
  uint32_t step_gid;       // 2 * rpw * W
  uint32_t *done;          // [replica][N2 + 2] finished row-steps per layer, [0] / [N2+1]: ghost layers
  uint32_t done_stride;
  // slabs over peer memory: the ring neighbours' lattices and layer counters
  int8_t *peer_dn, *peer_up;
  uint32_t *peer_done_dn, *peer_done_up;
  int32_t push;
  uint32_t agg;  // the warps of a block count their groups together (see k_sweep_stream16)
  uint32_t *trace;   // CMX_STREAM_TRACE: [n_units][2] groups that had to block / polls they spent, or null
  uint32_t *ticket;  // [replica] next group to hand out (dynamic assignment), or null: warp w takes w, w + n_warps, ...
  unsigned long long *fail;  // set when a dependency never arrives
  uint32_t dbg;  // CMX_STREAM_DEBUG (diagnostics only): 2 no release fence (WRONG), 4 no dependency checks (WRONG)
  // two neighbor classes (k_sweep_pass16 with MASK2_CT): tab24 / thr_lo / dEpot then hold n_tab2
  // entries per replica in the COMPACT index of s16_update_word2, maps2 = its three index maps
  uint32_t n_tab2;
  const uint16_t *maps2;  // [256] class-1 sum -> i1 * n2 | [256] class-2 sum -> i2 | [32] code | alt << 2 -> row offset
};
#define CMX_S16_MAPS 544u  // uint16 entries of the index maps

// per-thread constants
struct S16Lane {
  uint32_t tab;  // shared-memory address of the acceptance table
  uint32_t maps; // shared-memory address of the index maps (two neighbor classes)
  const double *dEpot;
  const uint32_t *thr_lo;
  int8_t *base;  // replica base (includes the ghost layers)
  uint32_t r, c, lane_l, lane_r;
  uint32_t rot_l, rot_r;  // 8 at the row ends (the word beyond the end is the row's other end, one byte lane over)
};

// ---- Metropolis test of one word: four sites of one x colour ---------------------------
//  cnt   byte-lane neighbor sums (n1 + 18 n2)
//  C     the four storage codes; accepted sites are replaced
//  R0/R1 random fields [alt:1 | u15:15] of bytes 0,1 / 2,3 (low half = lower byte)
//  rj    out: 0xFF in the byte lanes that were NOT accepted
//  tmin  running min over the 16-bit halves of D = (R | 0x8000) - thr: a half equal to
//        0x8000 is "the 15 bits tie" (the site was left rejected; s16_ties finishes it)
template <int NOCC>
__device__ __forceinline__ void s16_update_word(uint32_t cnt, uint32_t &C, uint32_t R0, uint32_t R1, uint32_t tab,
                                                uint32_t &rj, uint32_t &tmin, uint32_t (&idx)[4]) {
  uint32_t SA = C;
  if (NOCC == 3) {
    // alt = bit 15 of every field -> bit 2 of the site's byte (sign-replicating permute)
    const uint32_t am = prmt_s<0xFDB9u>(R0, R1);
    SA = C | (am & 0x04040404u);
  }
  // idx = (cnt + SA) byte b | SA byte b << 8 (cmx_tab24_index; no carry between the byte lanes);
  // bytes 2,3 <- sign of an SA byte (codes < 128: zero)
  const uint32_t cs = cnt + SA;
  idx[0] = prmt_s<0xCC40u>(cs, SA);
  idx[1] = prmt_s<0xDD51u>(cs, SA);
  idx[2] = prmt_s<0xEE62u>(cs, SA);
  idx[3] = prmt_s<0xFF73u>(cs, SA);
  const uint32_t e0 = lds_u32(tab + 4u * idx[0]);
  const uint32_t e1 = lds_u32(tab + 4u * idx[1]);
  const uint32_t e2 = lds_u32(tab + 4u * idx[2]);
  const uint32_t e3 = lds_u32(tab + 4u * idx[3]);
  const uint32_t T01 = __byte_perm(e0, e1, 0x5410u);
  const uint32_t T23 = __byte_perm(e2, e3, 0x5410u);
  const uint32_t D0 = (R0 | 0x80008000u) - T01;
  const uint32_t D1 = (R1 | 0x80008000u) - T23;
  tmin = min_s16x2(tmin, min_s16x2(D0, D1));
  rj = prmt_s<0xFDB9u>(D0, D1);  // bit 15 of a half: u15 >= thr
  const uint32_t P = __byte_perm(__byte_perm(e0, e1, 0x0062u), __byte_perm(e2, e3, 0x0062u), 0x5410u);
  C = (C & rj) | (P & ~rj);
}

// rare path (2^-15 per site): the 15 bits of a site equal its threshold's; the main path
// left the site rejected.  Draw 32 more bits and finish the 47-bit test.  Out of line, its
// operands pass through local memory: the hot loop keeps its registers.
struct S16Tie {
  uint32_t cnt[2], C[2], rj[2], R[4];
  double e_sum;
};
template <int NOCC, bool ACCUM>
__device__ __noinline__ void s16_ties(S16Tie *t, uint32_t tab, const uint32_t *__restrict__ thr_lo,
                                      const double *__restrict__ dEpot, uint32_t gid, uint32_t r,
                                      uint32_t sweep_lo, uint32_t ctr, uint32_t k0, uint32_t k1) {
  const Philox lo0 = philox_sweep(gid, r, sweep_lo, ctr | 1u, k0, k1);
  const Philox lo1 = philox_sweep(gid, r, sweep_lo, ctr | 2u, k0, k1);
  for (int h = 0; h < 2; ++h) {
    for (int b = 0; b < 4; ++b) {
      if (!((t->rj[h] >> (8 * b)) & 1u)) continue;  // accepted by the main path: the lane holds the NEW code
      const uint32_t R = t->R[2 * h + (b >> 1)];
      const uint32_t field = (b & 1) ? (R >> 16) : (R & 0xFFFFu);
      const uint32_t sab = ((t->C[h] >> (8 * b)) & 0xFFu) | ((NOCC == 3) ? ((field >> 15) << 2) : 0u);
      const uint32_t idx = cmx_tab24_index((t->cnt[h] >> (8 * b)) & 0xFFu, sab);
      const uint32_t e = lds_u32(tab + 4u * idx);
      if ((field & 0x7FFFu) != (e & 0xFFFFu)) continue;
      const int q = 4 * h + b;  // target site of the chunk's colour, 0..7
      const uint32_t w32 = (q < 4) ? lo0.c[q & 3] : lo1.c[q & 3];
      const uint32_t i16 = cmx_tab24_to_16(idx);
      if (w32 < thr_lo[i16]) {
        t->C[h] = (t->C[h] & ~(0xFFu << (8 * b))) | (((e >> 16) & 0xFFu) << (8 * b));
        t->rj[h] &= ~(0xFFu << (8 * b));
        if (ACCUM) t->e_sum += dEpot[i16];
      }
    }
  }
}

// ---- two neighbor classes (the reference's dense FCC ECI: 1NN + 2NN pairs) -------------------
// dE depends on (occupant, proposal, class-1 species counts, class-2 species counts).  The two
// byte-lane sums cnt1, cnt2 (n_B + 18 n_Va each) go through small index maps to the compact
// entry  idx = row(code, alt) + i1(cnt1) n2 + i2(cnt2)  of a table of 6 x n1 x n2 entries
// (FCC: 6 x 91 x 28 = 15 288, 61 KB of shared memory); entry, compare and merge are the
// one-class path's.
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
template <int NOCC>
__device__ __forceinline__ void s16_update_word2(uint32_t cnt1, uint32_t cnt2, uint32_t &C, uint32_t R0, uint32_t R1,
                                                 uint32_t tab, uint32_t maps, uint32_t &rj, uint32_t &tmin,
                                                 uint32_t (&idx)[4]) {
  uint32_t SA = C;
  if (NOCC == 3) {
    const uint32_t am = prmt_s<0xFDB9u>(R0, R1);
    SA = C | (am & 0x04040404u);
  }
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const uint32_t o1 = (b == 0) ? ((cnt1 << 1) & 0x1FEu) : ((cnt1 >> (8 * b - 1)) & 0x1FEu);
    const uint32_t o2 = (b == 0) ? ((cnt2 << 1) & 0x1FEu) : ((cnt2 >> (8 * b - 1)) & 0x1FEu);
    const uint32_t os = (b == 0) ? ((SA << 1) & 0x3Eu) : ((SA >> (8 * b - 1)) & 0x3Eu);
    idx[b] = lds_u16(maps + o1) + lds_u16(maps + 512u + o2) + lds_u16(maps + 1024u + os);
  }
  const uint32_t e0 = lds_u32(tab + 4u * idx[0]);
  const uint32_t e1 = lds_u32(tab + 4u * idx[1]);
  const uint32_t e2 = lds_u32(tab + 4u * idx[2]);
  const uint32_t e3 = lds_u32(tab + 4u * idx[3]);
  const uint32_t T01 = __byte_perm(e0, e1, 0x5410u);
  const uint32_t T23 = __byte_perm(e2, e3, 0x5410u);
  const uint32_t D0 = (R0 | 0x80008000u) - T01;
  const uint32_t D1 = (R1 | 0x80008000u) - T23;
  tmin = min_s16x2(tmin, min_s16x2(D0, D1));
  rj = prmt_s<0xFDB9u>(D0, D1);  // bit 15 of a half: u15 >= thr
  const uint32_t P = __byte_perm(__byte_perm(e0, e1, 0x0062u), __byte_perm(e2, e3, 0x0062u), 0x5410u);
  C = (C & rj) | (P & ~rj);
}
struct S16Tie2 {
  uint32_t cnt1[2], cnt2[2], C[2], rj[2], R[4];
  double e_sum;
};
template <int NOCC, bool ACCUM>
__device__ __noinline__ void s16_ties2(S16Tie2 *t, uint32_t tab, const uint16_t *__restrict__ maps,
                                       const uint32_t *__restrict__ thr_lo, const double *__restrict__ dEpot,
                                       uint32_t gid, uint32_t r, uint32_t sweep_lo, uint32_t ctr, uint32_t k0,
                                       uint32_t k1) {
  const Philox lo0 = philox_sweep(gid, r, sweep_lo, ctr | 1u, k0, k1);
  const Philox lo1 = philox_sweep(gid, r, sweep_lo, ctr | 2u, k0, k1);
  for (int h = 0; h < 2; ++h) {
    for (int b = 0; b < 4; ++b) {
      if (!((t->rj[h] >> (8 * b)) & 1u)) continue;  // accepted by the main path: the lane holds the NEW code
      const uint32_t R = t->R[2 * h + (b >> 1)];
      const uint32_t field = (b & 1) ? (R >> 16) : (R & 0xFFFFu);
      const uint32_t sab = ((t->C[h] >> (8 * b)) & 0xFFu) | ((NOCC == 3) ? ((field >> 15) << 2) : 0u);
      const uint32_t idx = (uint32_t)maps[(t->cnt1[h] >> (8 * b)) & 0xFFu] +
                           (uint32_t)maps[256u + ((t->cnt2[h] >> (8 * b)) & 0xFFu)] + (uint32_t)maps[512u + (sab & 31u)];
      const uint32_t e = lds_u32(tab + 4u * idx);
      if ((field & 0x7FFFu) != (e & 0xFFFFu)) continue;
      const int q = 4 * h + b;  // target site of the chunk's colour, 0..7
      const uint32_t w32 = (q < 4) ? lo0.c[q & 3] : lo1.c[q & 3];
      if (w32 < thr_lo[idx]) {
        t->C[h] = (t->C[h] & ~(0xFFu << (8 * b))) | (((e >> 16) & 0xFFu) << (8 * b));
        t->rj[h] &= ~(0xFFu << (8 * b));
        if (ACCUM) t->e_sum += dEpot[idx];
      }
    }
  }
}

// stage the rows a row-step reads into this lane's shared-memory slots (asynchronous
// 16-byte copies, L2 only): slot n = n-th row of the mask in (dz, dy) order.  pc = this
// lane's chunk of the target row j; dkm / dkp = byte offsets to the layers below / above
// (periodic wrap folded in by the caller, once per group)
template <uint32_t MASK_CT>
__device__ __forceinline__ void s16_issue(const S16Args &a, const int8_t *pc, int32_t j, int32_t dkm, int32_t dkp,
                                          uint32_t slots) {
  const uint32_t mask = MASK_CT ? MASK_CT : a.mask;
  const int32_t N0 = a.g.N0;
  const int32_t dj[3] = {(j == 0) ? a.wrap_j : -N0, 0, (j == a.g.N1 - 1) ? -a.wrap_j : N0};
  const int32_t dk[3] = {dkm, 0, dkp};
  uint32_t n = 0;
#pragma unroll
  for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const uint32_t m3 = (mask >> ((dz + 1) * 9 + (dy + 1) * 3)) & 7u;
      if (m3 == 0 && !(dz == 0 && dy == 0)) continue;
      cp_async16(slots + n * CMX_S16_SLOT, pc + (ptrdiff_t)(dk[dz + 1] + dj[dy + 1]));
      ++n;
    }
  }
}

// One row-step: every lane updates both x colours of its chunk of row (j, k) and stores it
// (lanes with on == false redo a valid row without storing).  The rows were staged by
// s16_issue into `slots`; after_loads() runs once this lane has read its slots.
// base + 16 t: one multiply-add (the chunk index t is what the callers carry)
__device__ __forceinline__ int8_t *s16_chunk_ptr(const int8_t *base, uint32_t t) {
  unsigned long long p;
  asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(p) : "r"(t), "l"((unsigned long long)base));
  return reinterpret_cast<int8_t *>(p);
}

// MASK2_CT != 0: a second neighbor class (compile-time offsets, none in the site's own row)
// summed separately; the table index is then s16_update_word2's.
template <int NOCC, uint32_t MASK_CT, bool ACCUM, bool SLAB, uint32_t MASK2_CT = 0u, typename Hook>
__device__ __forceinline__ void s16_rowstep(const S16Args &a, const S16Lane &L, const int8_t *pbase, uint32_t pt,
                                            uint32_t gid, int32_t k, uint32_t sweep_lo, uint32_t ctr_hi, bool on,
                                            int32_t &n_acc, double &e_tot, uint32_t slots, Hook after_loads) {
  constexpr bool TWO = MASK2_CT != 0u;
  static_assert(!TWO || (MASK_CT != 0u && ((MASK2_CT >> 12) & 7u) == 0u), "second class: compile-time masks, no same-row neighbors");
  const Geom &g = a.g;
  const uint32_t mask = MASK_CT ? MASK_CT : a.mask;
  const uint32_t mc = (mask >> 12) & 7u;  // center row: dx = -1 / +1 bits
  const uint32_t tab = L.tab, r = L.r;
  // both colours' random fields up front: two independent chains interleave
  const Philox ph0 = philox_sweep_rk(gid, r, sweep_lo, ctr_hi, a.rk);
  const Philox ph1 = philox_sweep_rk(gid, r, sweep_lo, ctr_hi | 0x100u, a.rk);
  uint32_t C[4], T[4], T2[4] = {0, 0, 0, 0};
  {
    uint32_t A0[4] = {0, 0, 0, 0}, Am[4] = {0, 0, 0, 0}, Ap[4] = {0, 0, 0, 0};
    uint32_t B0[4] = {0, 0, 0, 0}, Bm[4] = {0, 0, 0, 0}, Bp[4] = {0, 0, 0, 0};
    bool any_m = false, any_p = false;
    uint32_t n_slot = 0;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const uint32_t m3 = (mask >> ((dz + 1) * 9 + (dy + 1) * 3)) & 7u;
        const uint32_t m3b = (MASK2_CT >> ((dz + 1) * 9 + (dy + 1) * 3)) & 7u;
        const bool center = (dz == 0 && dy == 0);
        if ((m3 | m3b) == 0 && !center) continue;
        const uint4 ch = lds_u128(slots + (n_slot++) * CMX_S16_SLOT);
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
          any_m = true;
          Am[0] += ch.x;
          Am[1] += ch.y;
          Am[2] += ch.z;
          Am[3] += ch.w;
        }
        if (m3 & 4u) {
          any_p = true;
          Ap[0] += ch.x;
          Ap[1] += ch.y;
          Ap[2] += ch.z;
          Ap[3] += ch.w;
        }
        if (TWO) {
          if (m3b & 2u) {
            B0[0] += ch.x;
            B0[1] += ch.y;
            B0[2] += ch.z;
            B0[3] += ch.w;
          }
          if (m3b & 1u) {
            Bm[0] += ch.x;
            Bm[1] += ch.y;
            Bm[2] += ch.z;
            Bm[3] += ch.w;
          }
          if (m3b & 4u) {
            Bp[0] += ch.x;
            Bp[1] += ch.y;
            Bp[2] += ch.z;
            Bp[3] += ch.w;
          }
        }
      }
    }
    after_loads();
    // the words just outside the chunk: the summed classes of the neighbor chunks (at a
    // row end: the row's other end, one byte lane over)
    uint32_t sm = 0, sp = 0;
    if (any_m) {
      sm = __shfl_sync(0xffffffffu, Am[3], L.lane_l);
      sm = __funnelshift_l(sm, sm, L.rot_l);
    }
    if (any_p) {
      sp = __shfl_sync(0xffffffffu, Ap[0], L.lane_r);
      sp = __funnelshift_r(sp, sp, L.rot_r);
    }
    // T[w] = A0[w] + Am[w-1] + Ap[w+1]
    T[0] = A0[0] + sm + Ap[1];
    T[1] = A0[1] + Am[0] + Ap[2];
    T[2] = A0[2] + Am[1] + Ap[3];
    T[3] = A0[3] + Am[2] + sp;
    if (TWO) {
      uint32_t sm2 = __shfl_sync(0xffffffffu, Bm[3], L.lane_l);
      sm2 = __funnelshift_l(sm2, sm2, L.rot_l);
      uint32_t sp2 = __shfl_sync(0xffffffffu, Bp[0], L.lane_r);
      sp2 = __funnelshift_r(sp2, sp2, L.rot_r);
      T2[0] = B0[0] + sm2 + Bp[1];
      T2[1] = B0[1] + Bm[0] + Bp[2];
      T2[2] = B0[2] + Bm[1] + Bp[3];
      T2[3] = B0[3] + Bm[2] + sp2;
    }
  }
  uint32_t rj[4], tmin, idx[4];
  double e_sum = 0.0;
  auto accum = [&](uint32_t rjw, const uint32_t (&ix)[4]) {
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (!((rjw >> (8 * b)) & 1u)) e_sum += L.dEpot[TWO ? ix[b] : cmx_tab24_to_16(ix[b])];
  };
  // Metropolis test of one word / the rare tie path of a colour (words wa, wb), either table
  auto update = [&](uint32_t cnt, uint32_t cnt2, uint32_t &Cw, uint32_t R0, uint32_t R1, uint32_t &rjw) {
    if (TWO) s16_update_word2<NOCC>(cnt, cnt2, Cw, R0, R1, tab, L.maps, rjw, tmin, idx);
    else s16_update_word<NOCC>(cnt, Cw, R0, R1, tab, rjw, tmin, idx);
  };
  auto ties = [&](uint32_t ca, uint32_t cb, int wa, int wb, const Philox &ph, uint32_t ctr) {
    if (TWO) {
      S16Tie2 t;
      t.cnt1[0] = ca;
      t.cnt1[1] = cb;
      t.cnt2[0] = T2[wa];
      t.cnt2[1] = T2[wb];
      t.C[0] = C[wa];
      t.C[1] = C[wb];
      t.rj[0] = rj[wa];
      t.rj[1] = rj[wb];
#pragma unroll
      for (int i = 0; i < 4; ++i) t.R[i] = ph.c[i];
      t.e_sum = 0.0;
      s16_ties2<NOCC, ACCUM>(&t, tab, a.maps2, L.thr_lo, L.dEpot, gid, r, sweep_lo, ctr, a.k0, a.k1);
      C[wa] = t.C[0];
      C[wb] = t.C[1];
      rj[wa] = t.rj[0];
      rj[wb] = t.rj[1];
      if (ACCUM) e_sum += t.e_sum;
    } else {
      S16Tie t;
      t.cnt[0] = ca;
      t.cnt[1] = cb;
      t.C[0] = C[wa];
      t.C[1] = C[wb];
      t.rj[0] = rj[wa];
      t.rj[1] = rj[wb];
#pragma unroll
      for (int i = 0; i < 4; ++i) t.R[i] = ph.c[i];
      t.e_sum = 0.0;
      s16_ties<NOCC, ACCUM>(&t, tab, L.thr_lo, L.dEpot, gid, r, sweep_lo, ctr, a.k0, a.k1);
      C[wa] = t.C[0];
      C[wb] = t.C[1];
      rj[wa] = t.rj[0];
      rj[wb] = t.rj[1];
      if (ACCUM) e_sum += t.e_sum;
    }
  };
  // ---- x colour 0: words 0 and 2; their same-row neighbors are the odd words (old values)
  {
    uint32_t cl = 0;
    if (mc & 1u) {
      cl = __shfl_sync(0xffffffffu, C[3], L.lane_l);
      cl = __funnelshift_l(cl, cl, L.rot_l);
    }
    uint32_t c0 = T[0], c2 = T[2];
    if (mc & 1u) {
      c0 += cl;
      c2 += C[1];
    }
    if (mc & 4u) {
      c0 += C[1];
      c2 += C[3];
    }
    tmin = 0x7FFF7FFFu;
    update(c0, T2[0], C[0], ph0.c[0], ph0.c[1], rj[0]);
    if (ACCUM) accum(rj[0], idx);
    update(c2, T2[2], C[2], ph0.c[2], ph0.c[3], rj[2]);
    if (ACCUM) accum(rj[2], idx);
    const uint32_t tz = tmin ^ 0x80008000u;
    if ((tz - 0x00010001u) & ~tz & 0x80008000u) ties(c0, c2, 0, 2, ph0, ctr_hi);
  }
  // ---- x colour 1: words 1 and 3 against the updated even words; word 4 is the (updated)
  // first word of the next chunk of the row
  {
    uint32_t nb = 0;
    if (mc & 4u) {
      nb = __shfl_sync(0xffffffffu, C[0], L.lane_r);
      nb = __funnelshift_r(nb, nb, L.rot_r);
    }
    uint32_t c1 = T[1], c3 = T[3];
    if (mc & 1u) {
      c1 += C[0];
      c3 += C[2];
    }
    if (mc & 4u) {
      c1 += C[2];
      c3 += nb;
    }
    tmin = 0x7FFF7FFFu;
    update(c1, T2[1], C[1], ph1.c[0], ph1.c[1], rj[1]);
    if (ACCUM) accum(rj[1], idx);
    update(c3, T2[3], C[3], ph1.c[2], ph1.c[3], rj[3]);
    if (ACCUM) accum(rj[3], idx);
    const uint32_t tz = tmin ^ 0x80008000u;
    if ((tz - 0x00010001u) & ~tz & 0x80008000u) ties(c1, c3, 1, 3, ph1, ctr_hi | 0x100u);
  }
  if (on) {
    if (ACCUM) e_tot += e_sum;
    const uint4 out = make_uint4(C[0], C[1], C[2], C[3]);
    int8_t *const pc = s16_chunk_ptr(pbase, pt);
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(pc), "r"(out.x), "r"(out.y), "r"(out.z), "r"(out.w) : "memory");
    // rejected lanes hold 0xFF = -1: four signed dot products count them
    n_acc = __dp4a((int)rj[0], 0x01010101, n_acc);
    n_acc = __dp4a((int)rj[1], 0x01010101, n_acc);
    n_acc = __dp4a((int)rj[2], 0x01010101, n_acc);
    n_acc = __dp4a((int)rj[3], 0x01010101, n_acc);
    n_acc += 16;  // attempted - rejected = accepted
    if (SLAB && a.push) {
      // my layer 0 is the lower neighbour's upper ghost, my last layer the upper
      // neighbour's lower ghost (same slab geometry on every rank)
      const int32_t N2 = g.N2;
      const ptrdiff_t layer = (ptrdiff_t)g.N0 * g.N1;
      const ptrdiff_t o = pc - a.occ;
      if (k == 0) *reinterpret_cast<uint4 *>(a.peer_dn + o + (ptrdiff_t)N2 * layer) = out;
      if (k == N2 - 1) *reinterpret_cast<uint4 *>(a.peer_up + o - (ptrdiff_t)N2 * layer) = out;
    }
  }
}

// rows a row-step reads (center included) for a compile-time mask; 9 when the mask is a runtime value
__host__ __device__ constexpr uint32_t s16_n_slots(uint32_t mask_ct) {
  if (mask_ct == 0) return 9;
  uint32_t n = 0;
  for (int q = 0; q < 9; ++q) n += (q == 4 || ((mask_ct >> (3 * q)) & 7u)) ? 1u : 0u;
  return n;
}
template <int NOCC>
__host__ __device__ constexpr size_t s16_smem_bytes(uint32_t mask_ct) {
  return (size_t)CMX_TAB24(NOCC) * 4 + 8u * s16_n_slots(mask_ct) * CMX_S16_SLOT;
}

// the rows of the acceptance table that exist: (code | alt << 2) in {0,1,18,4,5,22}
template <int NOCC>
__device__ __forceinline__ void s16_load_table(uint32_t *sh_tab, const uint32_t *__restrict__ gt) {
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

// position of a warp in the call's sequence of row-steps
struct S16Pos {
  uint32_t g;            // group index
  uint32_t rs, rs_end;   // current row-step and end of the group's range within the unit
  uint32_t rs0;          // first row-step of the group
  uint32_t kf, sweep_lo, ctr_hi, need;
  // running address state of the current row-step: consecutive row-steps of a group are
  // 2 * rpw rows apart in the same layer, so chunk address and RNG counter advance by
  // constants and the layer offsets are computed once per group
  int8_t *pc;            // this lane's chunk of its target row
  uint32_t gid;          // global chunk id (RNG counter word)
  int32_t j;             // the row (only compared with the layer's edges)
  int32_t dkm, dkp;      // byte offsets to the layers below / above
  bool on;               // the lane's row exists (partial last row-step of a unit: false)
};

// FULL: the rows of a unit divide evenly into row-steps (no partial row-step anywhere)
template <int NOCC, uint32_t MASK_CT, bool ACCUM, bool SLAB, bool FULL>
__global__ void __launch_bounds__(256, CMX_S16_MINB) k_sweep_stream16(S16Args a) {
  constexpr int NTAB = CMX_TAB24(NOCC);
  constexpr int NTAB16 = CMX_TAB16(NOCC);
  constexpr uint32_t NSLOT = s16_n_slots(MASK_CT);
  // dynamic shared memory: [acceptance table][8 warps x NSLOT row slots x (32 lanes x 16 B)]
  extern __shared__ __align__(16) unsigned char sh_dyn[];
  uint32_t *sh_tab = reinterpret_cast<uint32_t *>(sh_dyn);
  unsigned char *sh_rows = sh_dyn + NTAB * 4;
  __shared__ long long sh_acc[8];
  __shared__ double sh_sum[8];
  __shared__ uint32_t sh_arr[8];  // arrivals per iteration (mod 8) of the block's warps
  __shared__ uint32_t sh_comp;    // iterations whose eight groups are all finished
  if (threadIdx.x < 8) sh_arr[threadIdx.x] = 0;
  if (threadIdx.x == 8) sh_comp = 0;
  const uint32_t r = blockIdx.y;
  s16_load_table<NOCC>(sh_tab, a.tab24 + (size_t)r * NTAB);
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  S16Lane L;
  {
    L.r = r;
    // (opaque: the compiler would otherwise rematerialise the shared-window address at every use)
    asm volatile("mov.u32 %0, %1;" : "=r"(L.tab) : "r"((uint32_t)__cvta_generic_to_shared(sh_tab)));
    L.maps = 0u;
    L.dEpot = a.dEpot + (size_t)r * NTAB16;
    L.thr_lo = a.thr_lo + (size_t)r * NTAB16;
    L.base = a.occ + (size_t)r * a.g.rep_stride;
    const uint32_t Wm = a.W - 1u;
    L.c = lane & Wm;
    L.lane_l = (lane & ~Wm) | ((L.c - 1u) & Wm);
    L.lane_r = (lane & ~Wm) | ((L.c + 1u) & Wm);
    L.rot_l = (L.c == 0) ? 8u : 0u;
    L.rot_r = (L.c == Wm) ? 8u : 0u;
  }
  const uint32_t rl = lane >> a.logW;       // row of the row-step this lane works on
  const uint32_t rpw_log = 5u - a.logW;     // log2(rows per row-step)
  const uint32_t slots = (uint32_t)__cvta_generic_to_shared(sh_rows) + wib * (NSLOT * CMX_S16_SLOT) + 16u * lane;
  uint32_t *done = a.done + (size_t)r * a.done_stride;
  const int32_t N2 = a.g.N2;
  const bool halo = a.g.halo != 0;
  int32_t n_acc = 0;
  long long n_acc64 = 0;
  double e_tot = 0.0;
  __syncthreads();

  const uint32_t warp0 = blockIdx.x * 8u + wib, n_warps = gridDim.x * 8u;
  auto decode = [&](uint32_t g, S16Pos &p) {
    uint32_t u, gi;
    fastdivmod(g, a.div_gpu, u, gi);
    const uint4 un = __ldg(reinterpret_cast<const uint4 *>(a.units) + u);
    const unsigned long long sweep = a.first_sweep + un.y;
    p.g = g;
    p.kf = un.x;
    p.sweep_lo = (uint32_t)sweep;
    p.ctr_hi = ((uint32_t)(sweep >> 32) << 16) | (((un.x >> CMX_S16_COLOUR_SHIFT) & 3u) << 9);
    p.need = a.base[un.w & 1u] + un.z;
    p.rs0 = p.rs = gi * a.gr;
    p.rs_end = min(p.rs + a.gr, a.tpu);
    const int32_t k = (int32_t)(un.x & 0xFFFFu);
    const uint32_t cy = (un.x >> CMX_S16_COLOUR_SHIFT) & 1u;
    const uint32_t jj = (p.rs << rpw_log) + rl;
    p.on = FULL ? true : (jj < a.J);
    p.j = 2 * (int32_t)(p.on ? jj : a.J - 1u) + (int32_t)cy;
    const uint32_t row = (uint32_t)(k + a.g.halo) * (uint32_t)a.g.N1 + (uint32_t)p.j;
    p.pc = L.base + ((size_t)row * (uint32_t)a.g.N0 + 16u * L.c);
    p.gid = ((uint32_t)(k + a.k_offset) * (uint32_t)a.g.N1 + (uint32_t)p.j) * a.W + L.c;
    p.dkm = (!halo && k == 0) ? a.wrap_k : -a.layer;
    p.dkp = (!halo && k == N2 - 1) ? -a.wrap_k : a.layer;
  };
  // the next row-step of the same group
  auto advance = [&](S16Pos &p) {
    p.rs += 1u;
    if (FULL) {
      p.pc += a.step_pc;
      p.gid += a.step_gid;
      p.j += a.step_j;
    } else {
      const uint32_t jj = (p.rs << rpw_log) + rl;
      if (jj < a.J) {
        p.pc += a.step_pc;
        p.gid += a.step_gid;
        p.j += a.step_j;
      } else {
        p.on = false;  // stays on its last row, unstored
      }
    }
  };
  // counters the unit waits for: index k + 1 of the layer; the layer below 0 / above N2-1
  // is the ghost layer (slabs) or the other end of the box
  auto dep_index = [&](uint32_t kf, uint32_t &i0, uint32_t &i1) {
    const int32_t k = (int32_t)(kf & 0xFFFFu);
    const uint32_t own = (uint32_t)k + 1u;
    const uint32_t lo = (k == 0) ? (halo ? 0u : (uint32_t)N2) : (uint32_t)k;
    const uint32_t hi = (k == N2 - 1) ? (halo ? (uint32_t)N2 + 1u : 1u) : (uint32_t)k + 2u;
    // (a neighbour nobody counts -- a ghost layer filled by the host -- is replaced by the
    // other one; `own` only for the row colour 1 units, whose threshold is the own layer's)
    if (kf & CMX_S16_DEP_OWN) {
      i0 = i1 = own;
    } else {
      i0 = (kf & CMX_S16_DEP_LO) ? lo : hi;
      i1 = (kf & CMX_S16_DEP_HI) ? hi : lo;
    }
  };
  // dependency test on polled counter values
  auto has_deps = [&](uint32_t kf) { return (kf & (CMX_S16_DEP_LO | CMX_S16_DEP_HI | CMX_S16_DEP_OWN)) && !(a.dbg & 4u); };
  // (system scope only for the ghost-layer counters, which the ring neighbours write)
  auto poll1 = [&](uint32_t i) {
    if (SLAB && (i == 0u || i == (uint32_t)N2 + 1u)) return ld_poll_u32<true>(done + i);
    return ld_poll_u32<false>(done + i);
  };
  auto poll = [&](const S16Pos &p, uint32_t &v0, uint32_t &v1) {
    uint32_t i0, i1;
    dep_index(p.kf, i0, i1);
    v0 = poll1(i0);
    v1 = (i1 != i0) ? poll1(i1) : v0;
  };
  auto reached = [&](const S16Pos &p, uint32_t v0, uint32_t v1) {
    return (int32_t)(v0 - p.need) >= 0 && (int32_t)(v1 - p.need) >= 0;
  };
  bool dead = false;
  auto wait_deps = [&](const S16Pos &p) {
    if (!has_deps(p.kf)) return;
    uint32_t v0, v1;
    poll(p, v0, v1);
    if (reached(p, v0, v1)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    if (a.trace && lane == 0) atomicAdd(a.trace + 2 * (p.g / a.div_gpu.d), 1u);
    for (;;) {
      __nanosleep(CMX_S16_POLL_NS);
      poll(p, v0, v1);
      if (a.trace && lane == 0) atomicAdd(a.trace + 2 * (p.g / a.div_gpu.d) + 1, 1u);
      if (reached(p, v0, v1)) return;
      if ((++spins & 1023u) == 0) {
        if (*reinterpret_cast<volatile unsigned long long *>(a.fail) != 0ull || clock64() - t0 > 20000000000ll) {
          // ~10 s: a producer is gone (grid not co-resident, a ring neighbour died)
          *reinterpret_cast<volatile unsigned long long *>(a.fail) = 1ull;
          dead = true;
          return;
        }
      }
    }
  };
  // ---- completion signals -------------------------------------------------------------
  // A finished group must be counted on its layer's counter with RELEASE semantics at GPU
  // scope (every store of the group visible before the count).  On this part a gpu-scope
  // release fence (MEMBAR.ALL.GPU + ERRBAR) stalls the warp for ~2 us, whoever issues it and
  // whenever: one per group made the kernel 2-4x slower.  So the eight warps of a block
  // (they always hold eight CONSECUTIVE groups: iteration n of warp w is group
  // (blockIdx.x * 8 + w) + n * n_warps) combine: each warp ends its group with a cheap
  // CTA-scope release on a shared arrival counter, and the LAST warp to arrive counts the
  // whole iteration of the block on the layer counter(s) behind ONE gpu-scope fence
  // (cta release -> cta acquire -> gpu release: cumulative).  That warp defers the fence
  // into its next row-step, after it has read its staged rows and before it stages the
  // next ones -- the other seven never wait at all.  The arrival itself is deferred to the
  // same point: ANY fence, CTA scope included, waits for the asynchronous copies the warp
  // has in flight, and right after a row-step's store those are the next rows on their way
  // from DRAM.  A warp that has to block on a dependency, or ends, settles first.  Arrival slots are reused every 8 iterations;
  // sh_comp (iterations completed, they complete in order) guards the reuse.
  const uint32_t gpu_groups = a.div_gpu.d;
  uint32_t pend_g0 = 0, pend_cnt = 0;
  auto flush = [&]() {
    if (!pend_cnt) return;
    __syncwarp();
    if (lane == 0) {
      uint32_t g = pend_g0;
      const uint32_t g_end = pend_g0 + pend_cnt;
      bool first = !(a.dbg & 2u);  // the first count carries the release fence
      // system scope only when a counted layer was also pushed to a ring neighbour
      bool sys = false;
      if (SLAB && a.push) {
        uint32_t u0, u1, q;
        fastdivmod(g, a.div_gpu, u0, q);
        fastdivmod(g_end - 1u, a.div_gpu, u1, q);
        for (uint32_t u = u0; u <= u1; ++u) {
          const int32_t k = (int32_t)(__ldg(&a.units[u].kf) & 0xFFFFu);
          sys |= (k == 0 || k == N2 - 1);
        }
      }
      while (g < g_end) {
        uint32_t u, gi0;
        fastdivmod(g, a.div_gpu, u, gi0);
        const uint32_t gi1 = min(gi0 + (g_end - g), gpu_groups);
        const uint32_t n_rs = min(gi1 * a.gr, a.tpu) - gi0 * a.gr;
        const int32_t k = (int32_t)(__ldg(&a.units[u].kf) & 0xFFFFu);
        if (first) {
          if (sys) red_release_add_u32<true>(done + k + 1, n_rs);
          else red_release_add_u32<false>(done + k + 1, n_rs);
        } else {
          red_relaxed_add_u32(done + k + 1, n_rs);
        }
        first = false;
        if (SLAB && a.push) {
          if (k == 0) red_relaxed_sys_add_u32(a.peer_done_dn + (size_t)r * a.done_stride + (uint32_t)N2 + 1u, n_rs);
          if (k == N2 - 1) red_relaxed_sys_add_u32(a.peer_done_up + (size_t)r * a.done_stride, n_rs);
        }
        g += gi1 - gi0;
      }
    }
    pend_cnt = 0;
  };
  // the warp finished group g, its iteration n.  Small boxes, whose dependent units cannot
  // be kept eight groups apart (a.agg == 0: a group could wait for a count that needs its
  // own warp's arrival), count every group by itself.
  auto arrive = [&](uint32_t n, uint32_t g) {
    if (!a.agg) {
      flush();
      pend_g0 = g;
      pend_cnt = 1u;
      return;
    }
    __syncwarp();
    uint32_t last = 0;
    const uint32_t g0 = blockIdx.x * 8u + n * n_warps;
    const uint32_t n_valid = min(a.n_groups - g0, 8u);
    if (lane == 0) {
      while ((int32_t)(n - *reinterpret_cast<volatile uint32_t *>(&sh_comp)) >= 8) __nanosleep(64);
      const uint32_t old = atom_acq_rel_cta_add(&sh_arr[n & 7u], 1u);
      if (old + 1u == n_valid) {
        sh_arr[n & 7u] = 0u;
        st_release_cta(&sh_comp, n + 1u);
        last = 1u;
      }
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      flush();  // (an older pending iteration of this warp: cannot normally exist)
      pend_g0 = g0;
      pend_cnt = n_valid;
    }
  };

  bool arr_pending = false;
  uint32_t arr_n = 0, arr_g = 0;
  auto settle = [&]() {
    if (arr_pending) {
      arrive(arr_n, arr_g);
      arr_pending = false;
    }
    flush();
  };
  // group assignment: static round robin, or tickets (a free warp takes the oldest group
  // nobody has started: slow warps then simply take fewer, and the groups in flight stay
  // the oldest n_warps -- with the static deal every warp must keep the pace of the slowest)
  // (the ticket stays in lane 0 until it is needed: broadcasting it right away would make
  // the warp wait for the atomic's round trip)
  auto take_ticket = [&]() -> uint32_t { return (lane == 0) ? atomicAdd(a.ticket + r, 1u) : 0u; };
  S16Pos nxt;
  uint32_t g_first = warp0, tk_next = 0;
  if (a.ticket) {
    g_first = __shfl_sync(0xffffffffu, take_ticket(), 0);
    tk_next = take_ticket();
  }
  bool have = g_first < a.n_groups;
  if (have) {
    decode(g_first, nxt);
    wait_deps(nxt);
    if (!dead) s16_issue<MASK_CT>(a, nxt.pc, nxt.j, nxt.dkm, nxt.dkp, slots);
  }
  uint32_t it = 0, itg = 0;
  while (have && !dead) {
    // the row-step to compute now; `nxt` moves on to the one after it
    int8_t *const pc = nxt.pc;
    const uint32_t gid = nxt.gid, sweep_lo = nxt.sweep_lo, ctr_hi = nxt.ctr_hi, g_cur = nxt.g;
    const int32_t k = (int32_t)(nxt.kf & 0xFFFFu);
    const bool on = nxt.on;
    bool more = true;
    const bool new_group = nxt.rs + 1u >= nxt.rs_end;
    if (!new_group) {
      advance(nxt);
    } else {
      const uint32_t g2 = a.ticket ? __shfl_sync(0xffffffffu, tk_next, 0) : nxt.g + n_warps;
      more = g2 < a.n_groups;
      if (more) {
        decode(g2, nxt);
        if (a.ticket) tk_next = take_ticket();  // used a whole group later: its latency is hidden
      }
    }
    // the next group's dependencies are polled now and tested when this row-step stages
    // the next rows: the poll's latency hides behind the row-step's arithmetic
    const bool check = more && new_group && has_deps(nxt.kf);
    uint32_t pv0 = 0, pv1 = 0;
    if (check) poll(nxt, pv0, pv1);
    bool issued = false;
    cp_async_wait_all();
    auto stage_next = [&]() {
      settle();
      // across a group boundary only when the next group's dependencies are already
      // satisfied (the normal case); otherwise after this group has been signalled --
      // the next group may depend on it
      if (more && (!check || reached(nxt, pv0, pv1))) {
        s16_issue<MASK_CT>(a, nxt.pc, nxt.j, nxt.dkm, nxt.dkp, slots);
        issued = true;
      }
    };
    s16_rowstep<NOCC, MASK_CT, ACCUM, SLAB>(a, L, pc, 0u, gid, k, sweep_lo, ctr_hi, on, n_acc, e_tot, slots, stage_next);
    if (new_group) {
      arr_pending = true;
      arr_n = itg++;
      arr_g = g_cur;
    }
    if (more && !issued) {
      settle();
      wait_deps(nxt);
      if (dead) break;
      s16_issue<MASK_CT>(a, nxt.pc, nxt.j, nxt.dkm, nxt.dkp, slots);
    }
    have = more;
    if ((++it & 0xFFFFFu) == 0) {  // a row-step adds at most 16 to the 32-bit count
      n_acc64 += n_acc;
      n_acc = 0;
    }
  }
  cp_async_wait_all();
  settle();
  n_acc64 += n_acc;

  // block reduction of the counters (fixed order -> deterministic) into the block's slot
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_acc64 += __shfl_down_sync(0xffffffffu, n_acc64, o);
    e_tot += __shfl_down_sync(0xffffffffu, e_tot, o);
  }
  if (lane == 0) {
    sh_acc[wib] = n_acc64;
    sh_sum[wib] = e_tot;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long A = 0;
    double E = 0.0;
    for (int w = 0; w < 8; ++w) {
      A += sh_acc[w];
      E += sh_sum[w];
    }
    const size_t slot = (size_t)r * a.part_stride + blockIdx.x;
    a.part_acc[slot] += A;
    a.part_dE[slot] += E;
  }
}

// ---- thin slabs: colour passes with grid barriers ----------------------------------------
// The streaming schedule needs dependent units a few hundred microseconds of work apart;
// a slab of a few dozen layers (a 512^3 box over 4 or 8 GPUs) is a sweep of ~20-40 us, in
// which every link of the colour chain (row colour 0 -> 1 -> other k colour ...) has only a
// quarter of that: the completion latency of a unit (~5 us: row-step, release fence, poll)
// cannot hide and the dataflow degenerates into waiting.  Such slabs run the same row-step
// as plain colour passes: n_sweeps x 4 passes in ONE cooperative launch, a grid barrier
// after every pass, the rows of a pass dealt round robin to the warps.  Halo exchange over
// peer memory as in the streaming kernel (boundary rows are also stored into the ring
// neighbour's ghost layer); the ring is ordered by epochs = completed k-colour groups: a
// pass waits for both neighbours' epoch in its boundary row-steps only, which come LAST
// (the layers are visited in rotated order), and block 0 publishes the epoch after the
// barrier that ends a group.
struct PassArgs {
  uint32_t n_sweeps;
  unsigned long long first_sweep;
  unsigned long long epoch0;   // k-colour groups this rank completed before the launch
  FastDiv div_tpu;
  uint32_t H;                  // layers of one k colour
  uint32_t dq_k, dq_r;         // (layers, row-steps) a warp moves on between its row-steps: divmod(warps, tpu)
  // the lattice in 16-byte chunks (kernel parameters are constant-bank operands: no registers)
  uint32_t ch_row, ch_layer, ch_pair;  // chunks per row (W), per layer, per two layers
  int32_t ch_wrap_j, ch_wrap_k;        // (N1 - 1) rows, (N2 - 1) layers
  uint32_t gid_off;                    // global chunk id (RNG counter) minus local chunk index
  int32_t kgroup;              // only this k colour (slabs whose halo the host exchanges), -1: both
  unsigned long long *my_sig, *peer_sig_dn, *peer_sig_up;
  // barrier between colour passes: arrivals so far (monotonic, modulo 2^32).  Replicas are
  // independent Markov chains, so on one GPU every replica (grid row) has its OWN counter
  // (bar + blockIdx.y * bar_stride, one 128-byte line each) and waits for its own blocks only;
  // the ring protocol of slabs needs all replicas behind one barrier (bar_stride = 0)
  uint32_t *bar;
  uint32_t bar_stride;
  uint32_t bar_base, n_blocks; // the counters' value when the launch starts; arrivals per barrier
};

// Grid barrier of the colour-pass kernel (cooperative launch: all blocks are co-resident).
// Every block adds one arrival (release at GPU scope: the block's stores of the pass,
// ordered before it by the block barrier, are visible to whoever sees the count), and waits
// until the count has reached `target`.  WARP 0 polls as a whole: cooperative_groups'
// grid.sync() polls with thread 0 alone, and that thread did not rejoin its warp afterwards
// (measured: warp 0 of every block ran the next pass split 1 + 31 lanes, every instruction
// twice and every shuffle through the divergent path, and the grid waited for it at the
// next barrier).  The poll is a relaxed load: what the barrier guards is read through L2
// only (cp.async.cg), control dependent on the polled value -- no L1 invalidation per poll.
__device__ __forceinline__ void s16_grid_barrier(uint32_t *bar, uint32_t target) {
  __syncthreads();
  if (threadIdx.x < 32u) {
    // (fence by the whole warp, the arrival predicated on lane 0: no divergent branch at all)
    asm volatile(
        "{ .reg .pred p; setp.eq.u32 p, %1, 0; fence.acq_rel.gpu; @p red.relaxed.gpu.global.add.u32 [%0], 1; }" ::"l"(bar),
        "r"(threadIdx.x)
        : "memory");
    while ((int32_t)(ld_poll_u32<false>(bar) - target) < 0) {
    }
  }
  __syncthreads();
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// stage the rows of a row-step, addressed by CHUNK INDEX: t = this lane's chunk of the
// target row counted from the replica base (16-byte units; a replica of an x4-interleaved
// state is < 2^32 chunks), djm / djp / dkm / dkp = chunks to the rows below / above in j and
// k (periodic wrap folded in by the caller).  One add and one multiply-add per row.
template <uint32_t MASK_CT>
__device__ __forceinline__ void s16_issue_t(const S16Args &a, const int8_t *base, uint32_t t, int32_t djm, int32_t djp,
                                            int32_t dkm, int32_t dkp, uint32_t slots) {
  const uint32_t mask = MASK_CT ? MASK_CT : a.mask;
  const int32_t dj[3] = {djm, 0, djp};
  const int32_t dk[3] = {dkm, 0, dkp};
  uint32_t n = 0;
#pragma unroll
  for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const uint32_t m3 = (mask >> ((dz + 1) * 9 + (dy + 1) * 3)) & 7u;
      if (m3 == 0 && !(dz == 0 && dy == 0)) continue;
      cp_async16(slots + n * CMX_S16_SLOT, s16_chunk_ptr(base, t + (uint32_t)(dk[dz + 1] + dj[dy + 1])));
      ++n;
    }
  }
}

// dynamic shared memory of the two-class instantiation: [table n_tab2 x 4 B][index maps][row slots]
__host__ __device__ constexpr uint32_t s16_tab2_bytes(uint32_t n_tab2) {
  return (n_tab2 * 4u + CMX_S16_MAPS * 2u + 15u) & ~15u;
}
// MASK2_CT != 0: two neighbor classes (s16_update_word2; 2 blocks per SM: the second set of
// byte-lane sums does not fit the 80-register budget of 3 blocks)
template <int NOCC, uint32_t MASK_CT, bool ACCUM, bool SLAB, bool FULL, uint32_t MASK2_CT = 0u>
__global__ void __launch_bounds__(256, MASK2_CT ? 2 : CMX_S16_MINB) k_sweep_pass16(S16Args a, PassArgs c) {
  constexpr bool TWO = MASK2_CT != 0u;
  static_assert(!TWO || !SLAB, "two neighbor classes: single-GPU boxes only");
  constexpr int NTAB = CMX_TAB24(NOCC);
  constexpr int NTAB16 = CMX_TAB16(NOCC);
  constexpr uint32_t NSLOT = s16_n_slots(MASK_CT | MASK2_CT);
  extern __shared__ __align__(16) unsigned char sh_dyn[];
  uint32_t *sh_tab = reinterpret_cast<uint32_t *>(sh_dyn);
  unsigned char *sh_rows = sh_dyn + (TWO ? s16_tab2_bytes(a.n_tab2) : (uint32_t)NTAB * 4u);
  __shared__ long long sh_acc[8];
  __shared__ double sh_sum[8];
  const uint32_t r = blockIdx.y;
  if (TWO) {
    const uint32_t *gt = a.tab24 + (size_t)r * a.n_tab2;
    for (uint32_t q = threadIdx.x; q < a.n_tab2; q += 256u) sh_tab[q] = gt[q];
    uint16_t *sh_maps = reinterpret_cast<uint16_t *>(sh_tab + a.n_tab2);
    for (uint32_t q = threadIdx.x; q < CMX_S16_MAPS; q += 256u) sh_maps[q] = a.maps2[q];
  } else {
    s16_load_table<NOCC>(sh_tab, a.tab24 + (size_t)r * NTAB);
  }
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  S16Lane L;
  {
    L.r = r;
    asm volatile("mov.u32 %0, %1;" : "=r"(L.tab) : "r"((uint32_t)__cvta_generic_to_shared(sh_tab)));
    L.maps = TWO ? L.tab + 4u * a.n_tab2 : 0u;
    L.dEpot = a.dEpot + (size_t)r * (TWO ? a.n_tab2 : (uint32_t)NTAB16);
    L.thr_lo = a.thr_lo + (size_t)r * (TWO ? a.n_tab2 : (uint32_t)NTAB16);
    L.base = a.occ + (size_t)r * a.g.rep_stride;
    const uint32_t Wm = a.W - 1u;
    L.c = lane & Wm;
    L.lane_l = (lane & ~Wm) | ((L.c - 1u) & Wm);
    L.lane_r = (lane & ~Wm) | ((L.c + 1u) & Wm);
    L.rot_l = (L.c == 0) ? 8u : 0u;
    L.rot_r = (L.c == Wm) ? 8u : 0u;
  }
  uint32_t bar_target = c.bar_base;
  const uint32_t rl = lane >> a.logW, rpw_log = 5u - a.logW;
  const uint32_t slots = (uint32_t)__cvta_generic_to_shared(sh_rows) + wib * (NSLOT * CMX_S16_SLOT) + 16u * lane;
  const uint32_t halo = (uint32_t)a.g.halo;
  int32_t n_acc = 0;
  long long n_acc64 = 0;
  double e_tot = 0.0;
  __syncthreads();
  const uint32_t warp0 = blockIdx.x * 8u + wib, n_warps = gridDim.x * 8u;
  const uint32_t n_rs = a.tpu * c.H;  // row-steps of one colour pass
  // Everything is addressed in 16-byte chunks from the replica base: row (k, j) starts at
  // chunk ((k + halo) N1 + j) W, a row-step is 64 chunks (2 rpw rows x W) after the previous
  // one of its layer, a layer pair 2 N1 W chunks.  The global chunk id (RNG counter) is the
  // local one plus a constant.
  unsigned long long epoch = c.epoch0;
  const bool ring = SLAB && a.push;
  const bool publisher = ring && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  auto publish = [&]() {
    // every store of the finished groups (the ones into the neighbours' ghost layers
    // included: the grid barrier ordered them before this thread) precedes the epoch
    __threadfence_system();
    st_release_sys_u64(c.peer_sig_dn + 1, epoch);  // I am their upper neighbour
    st_release_sys_u64(c.peer_sig_up + 0, epoch);  // I am their lower neighbour
  };
  if (publisher && epoch) publish();  // what the previous launches completed (idempotent)
  // this warp's first row-step of a pass: (layer of the colour, row-step of the layer)
  uint32_t kk0, rs0;
  fastdivmod(warp0, c.div_tpu, kk0, rs0);
  // the row-step in which this lane meets the periodic seam in j: row 0 (row colour 0, the
  // row below wraps) / row N1-1 (row colour 1, the row above wraps); none: never
  const uint32_t jl = a.J - 1u;
  const uint32_t rs_edge0 = (rl == 0u) ? 0u : 0xFFFFFFFFu;
  const uint32_t rs_edge1 = ((jl & ((1u << rpw_log) - 1u)) == rl) ? (jl >> rpw_log) : 0xFFFFFFFFu;
  const int8_t *const base = L.base;

  // One colour pass (cy, cz compile-time: the seam tests and the constant row offsets fold)
  auto pass = [&](auto cy_tag, auto cz_tag, uint32_t sweep_lo, uint32_t sweep_hi) {
    constexpr int CY = decltype(cy_tag)::value, CZ = decltype(cz_tag)::value;
    const uint32_t ctr_hi = (sweep_hi << 16) | ((uint32_t)(CZ * 2 + CY) << 9);
    // the layer next to a ghost layer (k = 0 for cz = 0, k = N2-1 for cz = 1) comes last
    const uint32_t rot = (ring && CZ == 0) ? 1u : 0u;
    // One box on one GPU: consecutive passes walk the layers in OPPOSITE directions (row
    // colour 1 backwards).  A pass ends with the layers it touched last still in L2 (126 MB
    // against a 134 MB lattice at 512^3), and that is where the next pass begins: most of
    // its reads hit, only the far end comes from DRAM again -- no extra synchronisation,
    // just the order.
    const bool back = !ring && CY != 0;
    // chunk of this lane in row-step 0 of layer pair 0 of the pass's colour
    const uint32_t t00 = (((uint32_t)CZ + halo) * (uint32_t)a.g.N1 + 2u * rl + (uint32_t)CY) * c.ch_row + L.c;
    const uint32_t rs_edge = CY ? rs_edge1 : rs_edge0;
    const uint32_t kk_edge = CZ ? c.H - 1u : 0u;
    // row offsets towards the seam side, inside the box / across the seam (slabs do not
    // wrap in k: ghost layers)
    const int32_t dj_in = CY ? (int32_t)c.ch_row : -(int32_t)c.ch_row, dj_wrap = CY ? -c.ch_wrap_j : c.ch_wrap_j;
    const int32_t dk_in = CZ ? (int32_t)c.ch_layer : -(int32_t)c.ch_layer;
    const int32_t dk_wrap = halo ? dk_in : (CZ ? -c.ch_wrap_k : c.ch_wrap_k);
    struct Pos {
      uint32_t t;   // this lane's chunk of its target row
      uint32_t kk;  // layer of the colour (after rotation / reversal): k = 2 kk + cz
      bool on;
    };
    // position of row-step (kq, rs) and the offsets of its seam-side rows
    auto locate = [&](uint32_t kq, uint32_t rs, Pos &p, int32_t &dj_e, int32_t &dk_e) {
      uint32_t kk = kq;
      if (ring) {
        kk += rot;
        if (kk >= c.H) kk -= c.H;
      } else if (back) {
        kk = c.H - 1u - kk;
      }
      p.kk = kk;
      p.on = true;
      uint32_t tr = rs * 64u;
      if (!FULL) {
        const uint32_t jj = (rs << rpw_log) + rl;
        if (jj >= a.J) {  // partial last row-step of a layer: redo the last row, unstored
          p.on = false;
          tr = 2u * (jl - rl) * c.ch_row;
        }
      }
      p.t = kk * c.ch_pair + tr + t00;
      bool edge = (rs == rs_edge);
      if (!FULL && !p.on) edge = CY ? true : (jl == 0u);  // (the redone row is the last one of its colour)
      dj_e = edge ? dj_wrap : dj_in;
      dk_e = (kk == kk_edge) ? dk_wrap : dk_in;
    };
    auto issue = [&](const Pos &p, int32_t dj_e, int32_t dk_e) {
      // (cy = 0: the row below may wrap, the row above never does; cy = 1 the other way)
      const int32_t djm = CY ? -(int32_t)c.ch_row : dj_e, djp = CY ? dj_e : (int32_t)c.ch_row;
      const int32_t dkm = CZ ? -(int32_t)c.ch_layer : dk_e, dkp = CZ ? dk_e : (int32_t)c.ch_layer;
      s16_issue_t<MASK_CT | MASK2_CT>(a, base, p.t, djm, djp, dkm, dkp, slots);
    };
    auto wait_neighbours = [&](uint32_t kk) {
      if (!ring || !epoch || kk != kk_edge) return;  // (warp-uniform: a row-step lies in one layer)
      if (lane == 0) {
        const long long t0 = clock64();
        while (ld_acquire_sys_u64(c.my_sig + 0) < epoch || ld_acquire_sys_u64(c.my_sig + 1) < epoch) {
          if (clock64() - t0 > 20000000000ll) {  // ~10 s: a neighbour is gone
            *reinterpret_cast<volatile unsigned long long *>(a.fail) = 1ull;
            break;
          }
          __nanosleep(100);
        }
      }
      __syncwarp();
    };
    Pos cur, nxt;
    uint32_t q = warp0, kq = kk0, rs = rs0;
    if (q < n_rs) {
      int32_t dj_e, dk_e;
      locate(kq, rs, nxt, dj_e, dk_e);
      wait_neighbours(nxt.kk);
      issue(nxt, dj_e, dk_e);
    }
    for (; q < n_rs; q += n_warps) {
      cur = nxt;
      const bool more = q + n_warps < n_rs;
      // the warp's next row-step: n_warps further = (dq_k layers, dq_r row-steps) with carry
      rs += c.dq_r;
      kq += c.dq_k;
      if (rs >= a.tpu) {
        rs -= a.tpu;
        kq += 1u;
      }
      int32_t dj_e = 0, dk_e = 0;
      if (more) locate(kq, rs, nxt, dj_e, dk_e);
      cp_async_wait_all();
      auto stage_next = [&]() {
        if (more) {
          wait_neighbours(nxt.kk);
          issue(nxt, dj_e, dk_e);
        }
      };
      s16_rowstep<NOCC, MASK_CT, ACCUM, SLAB, MASK2_CT>(a, L, base, cur.t, cur.t + c.gid_off, 2 * (int32_t)cur.kk + CZ,
                                                        sweep_lo, ctr_hi, FULL ? true : cur.on, n_acc, e_tot, slots,
                                                        stage_next);
    }
    n_acc64 += n_acc;  // (a pass adds at most 16 per row-step to the 32-bit count)
    n_acc = 0;
    // every store of the pass, the ones into the neighbours' ghost layers included
    bar_target += c.n_blocks;
    s16_grid_barrier(c.bar + (size_t)blockIdx.y * c.bar_stride, bar_target);
  };
  using T0 = std::integral_constant<int, 0>;
  using T1 = std::integral_constant<int, 1>;
  for (uint32_t s = 0; s < c.n_sweeps; ++s) {
    const unsigned long long sweep = c.first_sweep + s;
    const uint32_t sweep_lo = (uint32_t)sweep, sweep_hi = (uint32_t)(sweep >> 32);
    if (c.kgroup != 1) {
      pass(T0{}, T0{}, sweep_lo, sweep_hi);
      pass(T1{}, T0{}, sweep_lo, sweep_hi);
      ++epoch;
      if (publisher) publish();
    }
    if (c.kgroup != 0) {
      pass(T0{}, T1{}, sweep_lo, sweep_hi);
      pass(T1{}, T1{}, sweep_lo, sweep_hi);
      ++epoch;
      if (publisher) publish();
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_acc64 += __shfl_down_sync(0xffffffffu, n_acc64, o);
    e_tot += __shfl_down_sync(0xffffffffu, e_tot, o);
  }
  if (lane == 0) {
    sh_acc[wib] = n_acc64;
    sh_sum[wib] = e_tot;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long A = 0;
    double E = 0.0;
    for (int w = 0; w < 8; ++w) {
      A += sh_acc[w];
      E += sh_sum[w];
    }
    const size_t slot = (size_t)r * a.part_stride + blockIdx.x;
    a.part_acc[slot] += A;
    a.part_dE[slot] += E;
  }
}
