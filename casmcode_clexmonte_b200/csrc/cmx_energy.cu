// Fast ClusterExpansion::per_supercell() (reference row a6, call sites
// sampling_functions.cc:148,205-217,282 -- the samplers clex.formation_energy and
// potential_energy recompute it at every sample) for the models the pair-LUT sweep
// covers: point + pair functions of one neighbor class.
//
// The reference sums, over all unit cells, the generated
// _calc_restricted_global_corr_contribution (…default.cc:472-498), which counts every
// cluster once through the prototype-orbit equivalents that contain the origin
// cell: for a pair orbit that is HALF of the neighbor shell (the "forward"
// neighbors).  The per-cell energy  sum_c eci_c * contribution_c  therefore
// depends only on the cell's occupant and the species counts over its forward
// neighbors.  It is tabulated once by the faithful evaluator on representative
// neighborhoods (two different arrangements must agree), and the supercell sum is
// one streaming pass: 16 sites per thread, forward-neighbor counts for all 16 byte
// lanes at once (the storage code makes the byte sum n1 + 18 n2), one table
// lookup and one FP64 add per site, fixed-order block/grid reduction.
//
// k_energy_lin16 (the default): with point and pair functions only the per-cell
// energy is LINEAR in the forward-neighbor counts,
//   E_cell(o; n1, n2) = c0(o) + n1 d1(o) + n2 d2(o)        (verified on the table),
// so the supercell energy is a dot product of 9 coefficients with INTEGER bond counts
//   N_o = #cells with occupant o,  N_o1 = sum_{cells: o} n1,  N_o2 = sum_{cells: o} n2.
// With <= 7 forward neighbors the byte-lane sum n1 + 18 n2 splits into nibbles
// (low: n1 + 2 n2 <= 14, high: n2), and four dot-product instructions per 4 sites
// (occupant indicator lanes x count lanes) accumulate the bond counts: no table
// gather, no FP64 in the streaming pass, ~6 instructions per site; the FP64 work is
// 9 multiply-adds per replica in the finishing kernel.  Exact integer counting also
// makes the sum independent of the grid.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "cmx_internal.cuh"

static int invalid(const std::string &msg) {
  cmx_set_error(msg);
  return CMX_ERR_INVALID;
}

#define CMX_VA_CODE 18

// lut[occ << 8 | cnt], cnt = n1 + 18 n2 over the z forward neighbors
__global__ void k_build_cell_lut(DevTables T, int nocc, int z, const int32_t *__restrict__ fwd,
                                 int n_eci, const uint32_t *__restrict__ eci_idx,
                                 const double *__restrict__ eci_val, int reversed,
                                 double *__restrict__ lut) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 4 * 256) return;
  const int cnt = e & 255, oi = e >> 8;
  const int n2 = (nocc == 3) ? cnt / CMX_VA_CODE : 0;
  const int n1 = cnt - n2 * CMX_VA_CODE;
  if (oi >= nocc || n1 + n2 > z) {
    lut[e] = 0.0;
    return;
  }
  int8_t nb[64];
  for (int n = 0; n < T.nlist_len && n < 64; ++n) nb[n] = 0;
  for (int q = 0; q < z; ++q) {
    int qq = reversed ? z - 1 - q : q;
    nb[fwd[qq]] = (q < n1) ? 1 : ((q < n1 + n2) ? 2 : 0);
  }
  nb[0] = (int8_t)oi;
  LocalFetch f{nb};
  double E = 0.0;
  for (int q = 0; q < n_eci; ++q) {
    int c = (int)eci_idx[q];
    double v = cmx_eval_function_t(T, T.global_gbeg[c], T.global_gbeg[c + 1], f, 0, 0, 0);
    E = __dadd_rn(E, __dmul_rn(eci_val[q], v));
  }
  lut[e] = E;
}

struct EnergyArgs {
  const int8_t *occ;  // base of replica blockIdx.y = 0
  size_t rep_stride;  // bytes between replicas (blockIdx.y)
  Geom g;
  uint32_t mask;
  uint32_t W;       // 16-byte chunks per row
  FastDiv divW, divJ;
  uint32_t n_items;  // W * N1 * N2
  const double *lut;
  double *partial;  // [gridDim.y][gridDim.x]
  LinSums *sums;    // k_energy_lin16: [gridDim.y][gridDim.x]
  LinSums *sums2;   // second mask of the fused FCC pass
};

// one 16-site chunk: its codes C[4] and the byte-lane sums cnt[4] = n1 + 18 n2 over the
// forward neighbors of every site
__device__ __forceinline__ void energy_chunk(const EnergyArgs &a, const int8_t *base, uint32_t item,
                                             uint32_t (&C)[4], uint32_t (&cnt)[4]) {
  const Geom &g = a.g;
  const uint32_t N0 = g.N0, N1 = g.N1, N2 = g.N2;
  const uint32_t layer = N0 * N1;
  const bool halo = g.halo != 0;
  const uint32_t mask = a.mask;
    uint32_t row, c, k, j;
    fastdivmod(item, a.divW, row, c);
    fastdivmod(row, a.divJ, k, j);
    const uint32_t x0 = 16 * c;
    const uint32_t off_c = ((k + g.halo) * N1 + j) * N0 + x0;
    const uint32_t dl = (x0 == 0) ? N0 - 4 : 0u - 4u;
    const uint32_t dr = (x0 + 16 == N0) ? 16u - N0 : 16u;
    uint32_t dj[3], dk[3];
    dj[0] = (j == 0) ? (N1 - 1) * N0 : 0u - N0;
    dj[1] = 0;
    dj[2] = (j == N1 - 1) ? 0u - (N1 - 1) * N0 : N0;
    dk[0] = (!halo && k == 0) ? (N2 - 1) * layer : 0u - layer;
    dk[1] = 0;
    dk[2] = (!halo && k == N2 - 1) ? 0u - (N2 - 1) * layer : layer;
    uint32_t A0[4] = {0, 0, 0, 0}, Am[4] = {0, 0, 0, 0}, Ap[4] = {0, 0, 0, 0};
    C[0] = C[1] = C[2] = C[3] = 0;
    uint32_t sm = 0, sp = 0;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const uint32_t m3 = (mask >> ((dz + 1) * 9 + (dy + 1) * 3)) & 7u;
        const bool center = (dz == 0 && dy == 0);
        if (m3 == 0 && !center) continue;
        const uint32_t off = off_c + dk[dz + 1] + dj[dy + 1];
        const uint4 ch = *reinterpret_cast<const uint4 *>(base + off);
        if (center) {
          C[0] = ch.x;
          C[1] = ch.y;
          C[2] = ch.z;
          C[3] = ch.w;
        }
        if (m3 & 2u && !center) {
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
          sm += *reinterpret_cast<const uint32_t *>(base + (off + dl));
        }
        if (m3 & 4u) {
          Ap[0] += ch.x;
          Ap[1] += ch.y;
          Ap[2] += ch.z;
          Ap[3] += ch.w;
          sp += *reinterpret_cast<const uint32_t *>(base + (off + dr));
        }
      }
    }
    if (g.xq_log) {
      // x4-interleaved rows: the x neighbors of a word are the words beside it; beyond a
      // row end lies the row's other end, one byte lane over
      if (x0 == 0) sm = __funnelshift_l(sm, sm, 8);
      if (x0 + 16 == N0) sp = __funnelshift_r(sp, sp, 8);
      cnt[0] = A0[0] + sm + Ap[1];
      cnt[1] = A0[1] + Am[0] + Ap[2];
      cnt[2] = A0[2] + Am[1] + Ap[3];
      cnt[3] = A0[3] + Am[2] + sp;
      return;
    }
    cnt[0] = A0[0] + __funnelshift_l(sm, Am[0], 8) + __funnelshift_r(Ap[0], Ap[1], 8);
    cnt[1] = A0[1] + __funnelshift_l(Am[0], Am[1], 8) + __funnelshift_r(Ap[1], Ap[2], 8);
    cnt[2] = A0[2] + __funnelshift_l(Am[1], Am[2], 8) + __funnelshift_r(Ap[2], Ap[3], 8);
    cnt[3] = A0[3] + __funnelshift_l(Am[2], Am[3], 8) + __funnelshift_r(Ap[3], sp, 8);
}

template <int NOCC>
__global__ void __launch_bounds__(256) k_energy_pair16(EnergyArgs a) {
  __shared__ double sh_lut[NOCC * 256];
  __shared__ double sh_red[256];
  for (int q = threadIdx.x; q < NOCC * 256; q += blockDim.x) sh_lut[q] = a.lut[q];
  __syncthreads();
  const int8_t *base = a.occ + (size_t)blockIdx.y * a.rep_stride;
  double acc = 0.0;
  for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < a.n_items;
       item += gridDim.x * blockDim.x) {
    uint32_t C[4], cnt[4];
    energy_chunk(a, base, item, C, cnt);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t S = C[i] & 0x03030303u;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint32_t idx = ((cnt[i] >> (8 * b)) & 0xFFu) | (((S >> (8 * b)) & 0xFFu) << 8);
        acc += sh_lut[idx];
      }
    }
  }
  sh_red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh_red[threadIdx.x] += sh_red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) a.partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = sh_red[0];
}

// bond counts by dot-product instructions (see the file comment); NOCC == 2: codes 0 / 1 only
template <int NOCC>
__global__ void __launch_bounds__(256) k_energy_lin16(EnergyArgs a) {
  __shared__ unsigned long long sh_sum[6];
  if (threadIdx.x < 6) sh_sum[threadIdx.x] = 0;
  __syncthreads();
  const int8_t *base = a.occ + (size_t)blockIdx.y * a.rep_stride;
  unsigned long long tot[6] = {0, 0, 0, 0, 0, 0};
  uint32_t acc[6] = {0, 0, 0, 0, 0, 0};
  uint32_t it = 0;
  for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < a.n_items;
       item += gridDim.x * blockDim.x) {
    uint32_t C[4], cnt[4];
    energy_chunk(a, base, item, C, cnt);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t L = cnt[i] & 0x0F0F0F0Fu;         // n1 + 2 n2 per lane
      const uint32_t H = (cnt[i] >> 4) & 0x0F0F0F0Fu;  // n2 per lane
      const uint32_t p1 = C[i] & 0x01010101u;          // occupant 1 lanes (weight 1)
      acc[0] = __dp4a(p1, 0x01010101u, acc[0]);
      acc[2] = __dp4a(p1, L, acc[2]);
      acc[3] = __dp4a(p1, H, acc[3]);
      if (NOCC == 3) {
        const uint32_t p2 = C[i] & 0x10101010u;        // occupant 2 lanes (code 18; weight 16)
        acc[1] += __popc(p2);
        acc[4] = __dp4a(p2, L, acc[4]);
        acc[5] = __dp4a(p2, H, acc[5]);
      }
    }
    if ((++it & 1023u) == 0) {  // a chunk adds < 2^12 to a 32-bit accumulator
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        tot[q] += acc[q];
        acc[q] = 0;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    tot[q] += acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot[q] += __shfl_down_sync(0xffffffffu, tot[q], o);
    if ((threadIdx.x & 31) == 0 && tot[q]) atomicAdd(&sh_sum[q], tot[q]);  // integers: order-free
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    LinSums r;
    r.n1 = sh_sum[0];
    r.n2 = sh_sum[1];
    r.sl1 = sh_sum[2];
    r.sh1 = sh_sum[3];
    r.sl2 = sh_sum[4];
    r.sh2 = sh_sum[5];
    a.sums[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = r;
  }
}

// Warp-row variant of k_energy_lin16 (N0/16 a power of two <= 32, the shapes the
// warp-row sweep kernel covers): one warp owns 32 / W whole rows per iteration, the
// bytes just outside a lane's 16-byte window come from the neighbor lane by shuffle
// (no side-word loads), rows advance with a fixed stride (one division per tile).
// forward half of the FCC first shell as the generated basis orders it:
// (0,0,1) (0,1,-1) (0,1,0) (1,-1,0) (1,0,-1) (1,0,0); MASK_CT = 0: mask from the arguments
constexpr uint32_t kMaskFccFwd = 0x4148a0u;
// forward half of the FCC second shell: (1,-1,-1) (1,-1,1) (1,1,-1)
constexpr uint32_t kMaskFcc2nnFwd = 0x100104u;
// MASK2_CT != 0: the bond counts of a SECOND forward-neighbor set in the same pass (the
// global correlations of both FCC pair shells read the lattice once), into a.sums2
template <int NOCC, uint32_t MASK_CT, uint32_t MASK2_CT>
__global__ void __launch_bounds__(256) k_energy_row16(EnergyArgs a, uint32_t logW, uint32_t n_rows, uint32_t n_tiles,
                                                      uint32_t any_m_rt, uint32_t any_p_rt) {
  constexpr int NM = MASK2_CT ? 2 : 1;
  __shared__ unsigned long long sh_sum[NM][6];
  if (threadIdx.x < 6 * NM) (&sh_sum[0][0])[threadIdx.x] = 0;
  __syncthreads();
  const Geom &g = a.g;
  const int8_t *base = a.occ + (size_t)blockIdx.y * a.rep_stride;
  const int32_t N0 = g.N0, N1 = g.N1, N2 = g.N2, layer = N0 * N1;
  const bool halo = g.halo != 0;
  const uint32_t mask[2] = {MASK_CT ? MASK_CT : a.mask, MASK2_CT};
  const bool any_m[2] = {MASK_CT ? ((MASK_CT & 0x1249249u) != 0) : (any_m_rt != 0), (MASK2_CT & 0x1249249u) != 0};
  const bool any_p[2] = {MASK_CT ? ((MASK_CT & 0x4924924u) != 0) : (any_p_rt != 0), (MASK2_CT & 0x4924924u) != 0};
  const uint32_t lane = threadIdx.x & 31u, Wm = a.W - 1u, c = lane & Wm, rl = lane >> logW;
  const uint32_t lane_l = (lane & ~Wm) | ((c - 1u) & Wm), lane_r = (lane & ~Wm) | ((c + 1u) & Wm);
  const uint32_t rpw_log = 5u - logW;
  const uint32_t warp0 = blockIdx.x * 8u + (threadIdx.x >> 5), n_warps = gridDim.x * 8u;
  // 32-bit partial sums per lane, handed to the block's 64-bit shared sums every 1024 tiles and at
  // the end (no 64-bit totals in registers).  Scaled: [1] = 16 n2; [2] = sum over occupant-1 cells
  // of the whole count byte L + 16 H instead of L, [4] the same over occupant-2 cells (x 16 like
  // [5]) -- one dot product with the byte sums as they are, undone in `flush`.  The occupant
  // counts do not depend on the neighbor set: only set 0 counts them.
  uint32_t acc[NM][6];
#pragma unroll
  for (int m = 0; m < NM; ++m)
#pragma unroll
    for (int q = 0; q < 6; ++q) acc[m][q] = 0;
  auto flush = [&]() {
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      unsigned long long t[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) t[q] = acc[m][q];
      t[0] = acc[0][0];
      t[1] = acc[0][1] >> 4;
      t[2] -= 16ull * t[3];
      t[4] -= 16ull * t[5];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t[q] += __shfl_down_sync(0xffffffffu, t[q], o);
        if ((threadIdx.x & 31) == 0 && t[q]) atomicAdd(&sh_sum[m][q], t[q]);  // integers: order-free
      }
    }
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
      for (int q = 0; q < 6; ++q) acc[m][q] = 0;
  };
  uint32_t it = 0;
  for (uint32_t tile = warp0; tile < n_tiles; tile += n_warps) {
    const uint32_t row_raw = (tile << rpw_log) + rl;
    const bool on = row_raw < n_rows;  // a partial last tile: the idle lanes redo the last row, uncounted
    uint32_t k, j;
    fastdivmod(on ? row_raw : n_rows - 1u, a.divJ, k, j);
    const int8_t *pc = base + ((size_t)((k + (uint32_t)g.halo) * (uint32_t)N1 + j) * (uint32_t)N0 + 16u * c);
    int32_t dj[3], dk[3];
    dj[0] = (j == 0) ? (N1 - 1) * N0 : -N0;
    dj[1] = 0;
    dj[2] = ((int32_t)j == N1 - 1) ? -(N1 - 1) * N0 : N0;
    dk[0] = (!halo && k == 0) ? (N2 - 1) * layer : -layer;
    dk[1] = 0;
    dk[2] = (!halo && (int32_t)k == N2 - 1) ? -(N2 - 1) * layer : layer;
    uint32_t A0[NM][4], Am[NM][4], Ap[NM][4], C[4] = {0, 0, 0, 0};
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
      for (int i = 0; i < 4; ++i) A0[m][i] = Am[m][i] = Ap[m][i] = 0;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int sh = (dz + 1) * 9 + (dy + 1) * 3;
        const uint32_t m3[2] = {(mask[0] >> sh) & 7u, (mask[1] >> sh) & 7u};
        const bool center = (dz == 0 && dy == 0);
        if (m3[0] == 0 && m3[NM - 1] == 0 && !center) continue;
        const uint4 ch = *reinterpret_cast<const uint4 *>(pc + (ptrdiff_t)(dk[dz + 1] + dj[dy + 1]));
        if (center) {
          C[0] = ch.x;
          C[1] = ch.y;
          C[2] = ch.z;
          C[3] = ch.w;
        }
#pragma unroll
        for (int m = 0; m < NM; ++m) {
          if ((m3[m] & 2u) && !center) {
            A0[m][0] += ch.x;
            A0[m][1] += ch.y;
            A0[m][2] += ch.z;
            A0[m][3] += ch.w;
          }
          if (m3[m] & 1u) {
            Am[m][0] += ch.x;
            Am[m][1] += ch.y;
            Am[m][2] += ch.z;
            Am[m][3] += ch.w;
          }
          if (m3[m] & 4u) {
            Ap[m][0] += ch.x;
            Ap[m][1] += ch.y;
            Ap[m][2] += ch.z;
            Ap[m][3] += ch.w;
          }
        }
      }
    }
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      uint32_t sm = any_m[m] ? __shfl_sync(0xffffffffu, Am[m][3], lane_l) : 0u;
      uint32_t sp = any_p[m] ? __shfl_sync(0xffffffffu, Ap[m][0], lane_r) : 0u;
      uint32_t cnt[4];
      if (g.xq_log) {  // x4-interleaved rows, see energy_chunk
        if (c == 0) sm = __funnelshift_l(sm, sm, 8);
        if (c == Wm) sp = __funnelshift_r(sp, sp, 8);
        cnt[0] = A0[m][0] + sm + Ap[m][1];
        cnt[1] = A0[m][1] + Am[m][0] + Ap[m][2];
        cnt[2] = A0[m][2] + Am[m][1] + Ap[m][3];
        cnt[3] = A0[m][3] + Am[m][2] + sp;
      } else {
        cnt[0] = A0[m][0] + __funnelshift_l(sm, Am[m][0], 8) + __funnelshift_r(Ap[m][0], Ap[m][1], 8);
        cnt[1] = A0[m][1] + __funnelshift_l(Am[m][0], Am[m][1], 8) + __funnelshift_r(Ap[m][1], Ap[m][2], 8);
        cnt[2] = A0[m][2] + __funnelshift_l(Am[m][1], Am[m][2], 8) + __funnelshift_r(Ap[m][2], Ap[m][3], 8);
        cnt[3] = A0[m][3] + __funnelshift_l(Am[m][2], Am[m][3], 8) + __funnelshift_r(Ap[m][3], sp, 8);
      }
      if (on) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t H = (cnt[i] >> 4) & 0x0F0F0F0Fu;
          const uint32_t p1 = C[i] & 0x01010101u;
          if (m == 0) acc[0][0] = __dp4a(p1, 0x01010101u, acc[0][0]);
          acc[m][2] = __dp4a(p1, cnt[i], acc[m][2]);
          acc[m][3] = __dp4a(p1, H, acc[m][3]);
          if (NOCC == 3) {
            const uint32_t p2 = C[i] & 0x10101010u;
            if (m == 0) acc[0][1] = __dp4a(p2, 0x01010101u, acc[0][1]);
            acc[m][4] = __dp4a(p2, cnt[i], acc[m][4]);
            acc[m][5] = __dp4a(p2, H, acc[m][5]);
          }
        }
      }
    }
    if ((++it & 1023u) == 0) flush();  // a tile adds < 2^15 to a 32-bit sum
  }
  flush();
  __syncthreads();
  if (threadIdx.x < NM) {
    const int m = threadIdx.x;
    LinSums r;
    r.n1 = sh_sum[m][0];
    r.n2 = sh_sum[m][1];
    r.sl1 = sh_sum[m][2];
    r.sh1 = sh_sum[m][3];
    r.sl2 = sh_sum[m][4];
    r.sh2 = sh_sum[m][5];
    (m ? a.sums2 : a.sums)[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = r;
  }
}

__global__ void k_energy_lin_final(const LinSums *__restrict__ sums, int nb, long long n_cells, int z,
                                   const double *__restrict__ lin, double *out) {
  __shared__ unsigned long long sh[6];
  unsigned long long v[6];
  cmx_lin_reduce(sums, nb, sh, v);
  if (threadIdx.x == 0) *out = cmx_lin_energy(v, n_cells, z, lin, nullptr);
}

__global__ void k_energy_final(const double *__restrict__ partial, int nb, double *out) {
  if (threadIdx.x || blockIdx.x) return;
  double e = 0.0;
  for (int q = 0; q < nb; ++q) e += partial[q];
  *out = e;
}

int cmx_plan_energy(cmx_state *s) {
  SweepPlan &P = s->plan;
  const cmx_tables *t = s->t;
  const DevTables &T = t->d;
  P.e_fast = false;
  if (s->g.s10 | s->g.s20 | s->g.s21) return CMX_OK;  // row kernels assume a diag(N) box
  // forward neighbors = the neighbor-list sites the selected GLOBAL functions read
  std::vector<char> used(T.nlist_len, 0);
  for (int q = 0; q < s->n_eci; ++q) {
    int c = (int)s->eci_idx[q];
    for (int g = t->global_gbeg[c]; g < t->global_gbeg[c + 1]; ++g) {
      if (t->group_dphi[g] >= 0) return CMX_OK;  // not a plain global function
      for (int el = t->group_ebeg[g]; el < t->group_ebeg[g + 1]; ++el)
        for (int tm = t->elem_tbeg[el]; tm < t->elem_tbeg[el + 1]; ++tm) {
          int nf = t->term_fbeg[tm + 1] - t->term_fbeg[tm];
          if (nf > 2) return CMX_OK;  // beyond pairs
          bool has_self = false;
          for (int f = t->term_fbeg[tm]; f < t->term_fbeg[tm + 1]; ++f) {
            used[t->factor_n[f]] = 1;
            if (t->factor_n[f] == 0) has_self = true;
          }
          if (nf == 2 && !has_self) return CMX_OK;  // a pair that does not contain the origin
        }
    }
  }
  std::vector<int32_t> fwd;
  uint32_t mask = 0;
  for (int n = 1; n < T.nlist_len; ++n)
    if (used[n]) {
      const int32_t *o = &t->nbr[4 * n];
      if (std::abs(o[0]) > 1 || std::abs(o[1]) > 1 || std::abs(o[2]) > 1) return CMX_OK;
      mask |= 1u << ((o[2] + 1) * 9 + (o[1] + 1) * 3 + (o[0] + 1));
      fwd.push_back(n);
    }
  if (fwd.size() > 14 || T.nlist_len > 64) return CMX_OK;
  // tabulate with two different arrangements of the species over the forward
  // neighbors: the table is only valid if the energy depends on the counts alone
  int32_t *d_fwd = nullptr;
  double *d_lut2 = nullptr;
  CMX_CUDA(cudaMalloc((void **)&d_fwd, sizeof(int32_t) * std::max<size_t>(1, fwd.size())));
  if (!fwd.empty())
    CMX_CUDA(cudaMemcpy(d_fwd, fwd.data(), sizeof(int32_t) * fwd.size(), cudaMemcpyHostToDevice));
  CMX_CUDA(cudaMalloc((void **)&P.d_e_lut, sizeof(double) * 1024));
  CMX_CUDA(cudaMalloc((void **)&d_lut2, sizeof(double) * 1024));
  for (int rev = 0; rev < 2; ++rev)
    k_build_cell_lut<<<8, 128, 0, s->stream>>>(T, P.nocc, (int)fwd.size(), d_fwd, s->n_eci, s->d_eci_idx,
                                               s->d_eci_val, rev, rev ? d_lut2 : P.d_e_lut);
  CMX_CUDA(cudaGetLastError());
  std::vector<double> l1(1024), l2(1024);
  CMX_CUDA(cudaMemcpyAsync(l1.data(), P.d_e_lut, sizeof(double) * 1024, cudaMemcpyDeviceToHost, s->stream));
  CMX_CUDA(cudaMemcpyAsync(l2.data(), d_lut2, sizeof(double) * 1024, cudaMemcpyDeviceToHost, s->stream));
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  cudaFree(d_fwd);
  cudaFree(d_lut2);
  double scale = 0.0;
  for (double x : l1) scale = std::max(scale, std::fabs(x));
  for (int q = 0; q < 1024; ++q)
    if (std::fabs(l1[q] - l2[q]) > 1e-13 * std::max(scale, 1e-300)) return CMX_OK;  // arrangement matters
  P.e_mask = mask;
  P.e_z = (int32_t)fwd.size();
  P.e_fast = true;
  // linear in the counts (always true for point + pair functions; checked, not assumed)
  P.e_lin = false;
  if (P.e_z <= 7 && (P.nocc == 2 || P.nocc == 3)) {
    bool ok = true;
    for (int o = 0; o < 3; ++o) {
      double c0 = 0.0, d1 = 0.0, d2 = 0.0;
      if (o < P.nocc) {
        c0 = l1[o << 8];
        d1 = (P.e_z >= 1) ? l1[(o << 8) | 1] - c0 : 0.0;
        d2 = (P.nocc == 3 && P.e_z >= 1) ? l1[(o << 8) | CMX_VA_CODE] - c0 : 0.0;
        for (int n2 = 0; n2 <= (P.nocc == 3 ? P.e_z : 0); ++n2)
          for (int n1 = 0; n1 + n2 <= P.e_z; ++n1) {
            const double want = l1[(o << 8) | (n1 + CMX_VA_CODE * n2)];
            if (std::fabs(c0 + n1 * d1 + n2 * d2 - want) > 1e-12 * std::max(scale, 1e-300)) ok = false;
          }
      }
      P.e_lin_c[3 * o] = c0;
      P.e_lin_c[3 * o + 1] = d1;
      P.e_lin_c[3 * o + 2] = d2;
    }
    if (ok) {
      if (!P.d_e_lin) CMX_CUDA(cudaMalloc((void **)&P.d_e_lin, sizeof(double) * 9));
      CMX_CUDA(cudaMemcpy(P.d_e_lin, P.e_lin_c, sizeof(double) * 9, cudaMemcpyHostToDevice));
      P.e_lin = true;
    }
  }
  return CMX_OK;
}

static EnergyArgs energy_args(const cmx_state *s, int32_t first_replica) {
  const SweepPlan &P = s->plan;
  EnergyArgs a;
  a.occ = s->d_occ + (size_t)first_replica * s->g.rep_stride;
  a.rep_stride = (size_t)s->g.rep_stride;
  a.g = s->g;
  a.mask = P.e_mask;
  a.W = s->g.N0 / 16;
  a.divW = make_fastdiv(a.W);
  a.divJ = make_fastdiv((uint32_t)s->g.N1);
  a.n_items = a.W * (uint32_t)s->g.N1 * (uint32_t)s->g.N2;
  a.lut = P.d_e_lut;
  a.partial = nullptr;
  a.sums = nullptr;
  a.sums2 = nullptr;
  return a;
}

// launch the bond-count pass over `n_rep` replicas starting at a.occ
static bool energy_warp_rows(const EnergyArgs &a) {
  static const bool force_block = getenv("CMX_ENERGY_BLOCK") != nullptr;  // cross-check of the two kernels
  return a.W <= 32 && (a.W & (a.W - 1)) == 0 && !force_block;  // a warp owns whole rows
}
static int launch_energy_lin(const cmx_state *s, EnergyArgs &a, int nb, int n_rep, int nocc_override = 0,
                             uint32_t mask2 = 0) {
  const int nocc = nocc_override ? nocc_override : s->plan.nocc;
  dim3 grid(nb, n_rep);
  if (energy_warp_rows(a)) {
    uint32_t logW = 0;
    while ((1u << logW) < a.W) ++logW;
    const uint32_t n_rows = (uint32_t)s->g.N1 * (uint32_t)s->g.N2, rpw = 32u >> logW;
    const uint32_t n_tiles = (n_rows + rpw - 1) / rpw;
    uint32_t any_m = 0, any_p = 0;
    for (int q = 0; q < 9; ++q) {
      any_m |= (a.mask >> (3 * q)) & 1u;
      any_p |= (a.mask >> (3 * q + 2)) & 1u;
    }
    if (a.mask == kMaskFccFwd && mask2 == kMaskFcc2nnFwd) {
      if (nocc == 3) k_energy_row16<3, kMaskFccFwd, kMaskFcc2nnFwd><<<grid, 256, 0, s->stream>>>(a, logW, n_rows, n_tiles, any_m, any_p);
      else k_energy_row16<2, kMaskFccFwd, kMaskFcc2nnFwd><<<grid, 256, 0, s->stream>>>(a, logW, n_rows, n_tiles, any_m, any_p);
    } else if (mask2) {
      return CMX_ERR_UNSUPPORTED;  // (callers ask cmx_energy_lin_fusable first)
    } else if (a.mask == kMaskFccFwd) {
      if (nocc == 3) k_energy_row16<3, kMaskFccFwd, 0u><<<grid, 256, 0, s->stream>>>(a, logW, n_rows, n_tiles, any_m, any_p);
      else k_energy_row16<2, kMaskFccFwd, 0u><<<grid, 256, 0, s->stream>>>(a, logW, n_rows, n_tiles, any_m, any_p);
    } else {
      if (nocc == 3) k_energy_row16<3, 0u, 0u><<<grid, 256, 0, s->stream>>>(a, logW, n_rows, n_tiles, any_m, any_p);
      else k_energy_row16<2, 0u, 0u><<<grid, 256, 0, s->stream>>>(a, logW, n_rows, n_tiles, any_m, any_p);
    }
  } else {
    if (mask2) return CMX_ERR_UNSUPPORTED;
    if (nocc == 3) k_energy_lin16<3><<<grid, 256, 0, s->stream>>>(a);
    else k_energy_lin16<2><<<grid, 256, 0, s->stream>>>(a);
  }
  CMX_CUDA(cudaGetLastError());
  return CMX_OK;
}

int cmx_energy_fast_blocks(const cmx_state *s) {
  const uint32_t n_items = (uint32_t)(s->g.N0 / 16) * (uint32_t)s->g.N1 * (uint32_t)s->g.N2;
  // one co-resident wave over all replicas (4 blocks of 256 threads per SM): every warp
  // works through tens of tiles, the per-block reduction is amortised
  static const int per_sm = getenv("CMX_ENERGY_BLOCKS_PER_SM") ? atoi(getenv("CMX_ENERGY_BLOCKS_PER_SM")) : 4;
  const int cap = std::max(1, 148 * per_sm / std::max(1, s->n_replicas));
  return (int)std::min<uint32_t>((n_items + 255) / 256, (uint32_t)cap);
}

// every replica in one launch (asynchronous, on the state's stream): per-block bond
// counts d_sums[replica][nb]; requires plan.e_lin
int cmx_energy_lin_batch(cmx_state *s, int nb, LinSums *d_sums) {
  EnergyArgs a = energy_args(s, 0);
  a.sums = d_sums;
  return launch_energy_lin(s, a, nb, s->n_replicas);
}

int cmx_energy_fast(cmx_state *s, int32_t replica, double *E) {
  SweepPlan &P = s->plan;
  CMX_CUDA(cudaSetDevice(s->t->device));
  EnergyArgs a = energy_args(s, replica);
  int nb = (int)std::min<uint32_t>((a.n_items + 255) / 256, 148 * 8);
  static const bool force_lut = getenv("CMX_ENERGY_LUT") != nullptr;  // cross-check of the two kernels
  if (P.e_lin && !force_lut) {
    int rc = cmx_scratch(s, sizeof(LinSums) * nb + sizeof(double));
    if (rc) return rc;
    a.sums = (LinSums *)s->d_scratch;
    double *d_E = (double *)(a.sums + nb);
    if ((rc = launch_energy_lin(s, a, nb, 1))) return rc;
    k_energy_lin_final<<<1, 256, 0, s->stream>>>(a.sums, nb, (long long)s->g.n_cells, P.e_z, P.d_e_lin, d_E);
    CMX_CUDA(cudaGetLastError());
    CMX_CUDA(cudaMemcpyAsync(E, d_E, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CMX_CUDA(cudaStreamSynchronize(s->stream));
    return CMX_OK;
  }
  int rc = cmx_scratch(s, sizeof(double) * (nb + 1));
  if (rc) return rc;
  a.partial = (double *)s->d_scratch;
  if (P.nocc == 3) k_energy_pair16<3><<<nb, 256, 0, s->stream>>>(a);
  else k_energy_pair16<2><<<nb, 256, 0, s->stream>>>(a);
  CMX_CUDA(cudaGetLastError());
  k_energy_final<<<1, 32, 0, s->stream>>>(a.partial, nb, a.partial + nb);
  CMX_CUDA(cudaGetLastError());
  CMX_CUDA(cudaMemcpyAsync(E, a.partial + nb, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  return CMX_OK;
}

// ---------------------------------------------------------------------------
// Streaming Correlations::per_supercell() (reference row a6: the `corr.<bset>` sampler,
// monte_calculator/sampling_functions.cc:121-139, sums the generated
// _calc_global_corr_contribution over all unit cells).
//
// For basis sets of point and pair functions on one sublattice the per-cell value of EVERY
// correlation function is linear in the species counts over the function's forward
// neighbors (its orbit's half shell),  f_c(o; n1, n2) = a_c(o) + n1 b_c(o) + n2 d_c(o)  --
// tabulated by the faithful evaluator and checked, as for the energy above.  The functions
// are grouped by forward-neighbor set (FCC pairs <= 2NN: none, the 1NN half shell, the 2NN
// half shell); per set ONE pass of the integer bond-count kernel of the energy (the same
// launch, another mask) counts N_o, N_o1, N_o2, and every function of the set is a dot
// product of nine coefficients with these integers.  512^3: two passes of ~70 us instead of
// 39 ms of term-by-term evaluation; exact bond counts, so independent of the grid, and
// within 1e-12 relative of the faithful sum (summation order).
// ---------------------------------------------------------------------------
struct CorrLinFinalArgs {
  const LinSums *sums;  // [n_masks][nb]
  int nb, n_masks, corr_size, mask0_from;
  long long n_cells;
  int z[4];
  const int32_t *func_mask;  // [corr_size]
  const double *lin;         // [corr_size][9]
  double *out;               // [corr_size]
};
__global__ void __launch_bounds__(256) k_corr_lin_final(CorrLinFinalArgs a) {  // one block
  __shared__ unsigned long long sh[4][6];
  if (threadIdx.x < 24) (&sh[0][0])[threadIdx.x] = 0;
  __syncthreads();
  for (int m = 0; m < a.n_masks; ++m) {
    const LinSums *sm = a.sums + (size_t)(m ? m : a.mask0_from) * a.nb;
    unsigned long long v[6] = {0, 0, 0, 0, 0, 0};
    for (int b = threadIdx.x; b < a.nb; b += blockDim.x) {  // integers: any order
      v[0] += sm[b].n1;
      v[1] += sm[b].n2;
      v[2] += sm[b].sl1;
      v[3] += sm[b].sh1;
      v[4] += sm[b].sl2;
      v[5] += sm[b].sh2;
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
      if ((threadIdx.x & 31) == 0 && v[q]) atomicAdd(&sh[m][q], v[q]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.corr_size; c += blockDim.x) {
    const int m = a.func_mask[c];
    unsigned long long v[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) v[q] = sh[m][q];
    a.out[c] = cmx_lin_energy(v, a.n_cells, a.z[m], a.lin + 9 * c, nullptr);
  }
}

int cmx_plan_corr_lin(cmx_state *s) {
  SweepPlan &P = s->plan;
  if (P.corr_lin_state != 0) return CMX_OK;
  P.corr_lin_state = -1;
  static const bool off = getenv("CMX_NO_CORR_LIN") != nullptr;
  const cmx_tables *t = s->t;
  const DevTables &T = t->d;
  const int nocc = t->n_occ.empty() ? 0 : t->n_occ[0];
  if (off || (s->g.s10 | s->g.s20 | s->g.s21) || T.n_sublat != 1 || nocc < 2 || nocc > 3 || T.nlist_len > 64 || s->g.N0 % 16 != 0 ||
      s->g.n_cells % 16 != 0 || s->g.coded != (nocc == 3) || T.n_point_corr != T.n_nlist_sublat)
    return CMX_OK;
  const int nc = T.corr_size;
  std::vector<std::vector<int32_t>> fwd(nc);
  std::vector<uint32_t> masks;  // distinct, masks[0] = 0 (no neighbors)
  masks.push_back(0u);
  std::vector<int32_t> z_of(1, 0), func_mask(nc, 0);
  for (int c = 0; c < nc; ++c) {
    std::vector<char> used(T.nlist_len, 0);
    for (int g = t->global_gbeg[c]; g < t->global_gbeg[c + 1]; ++g) {
      if (t->group_dphi[g] >= 0) return CMX_OK;
      for (int el = t->group_ebeg[g]; el < t->group_ebeg[g + 1]; ++el)
        for (int tm = t->elem_tbeg[el]; tm < t->elem_tbeg[el + 1]; ++tm) {
          const int nf = t->term_fbeg[tm + 1] - t->term_fbeg[tm];
          if (nf > 2) return CMX_OK;  // beyond pairs
          bool has_self = false;
          for (int f = t->term_fbeg[tm]; f < t->term_fbeg[tm + 1]; ++f) {
            used[t->factor_n[f]] = 1;
            has_self |= (t->factor_n[f] == 0);
          }
          if (nf == 2 && !has_self) return CMX_OK;
        }
    }
    uint32_t mask = 0;
    for (int n = 1; n < T.nlist_len; ++n)
      if (used[n]) {
        const int32_t *o = &t->nbr[4 * n];
        if (std::abs(o[0]) > 1 || std::abs(o[1]) > 1 || std::abs(o[2]) > 1) return CMX_OK;
        mask |= 1u << ((o[2] + 1) * 9 + (o[1] + 1) * 3 + (o[0] + 1));
        fwd[c].push_back(n);
      }
    if (fwd[c].size() > 7) return CMX_OK;
    size_t m = 0;
    while (m < masks.size() && masks[m] != mask) ++m;
    if (m == masks.size()) {
      if (masks.size() == 4) return CMX_OK;
      masks.push_back(mask);
      z_of.push_back((int32_t)fwd[c].size());
    }
    func_mask[c] = (int32_t)m;
  }
  // per function: tabulate with two arrangements, check that the counts alone matter and that
  // the table is linear in them
  std::vector<double> lin((size_t)nc * 9, 0.0);
  int32_t *d_fwd = nullptr;
  uint32_t *d_idx = nullptr;
  double *d_one = nullptr, *d_lut = nullptr;
  CMX_CUDA(cudaMalloc((void **)&d_fwd, sizeof(int32_t) * 8));
  CMX_CUDA(cudaMalloc((void **)&d_idx, sizeof(uint32_t)));
  CMX_CUDA(cudaMalloc((void **)&d_one, sizeof(double)));
  CMX_CUDA(cudaMalloc((void **)&d_lut, sizeof(double) * 2048));
  const double one = 1.0;
  CMX_CUDA(cudaMemcpy(d_one, &one, sizeof(double), cudaMemcpyHostToDevice));
  bool ok = true;
  std::vector<double> l1(1024), l2(1024);
  for (int c = 0; c < nc && ok; ++c) {
    const int z = (int)fwd[c].size();
    const uint32_t cu = (uint32_t)c;
    if (z) CMX_CUDA(cudaMemcpy(d_fwd, fwd[c].data(), sizeof(int32_t) * z, cudaMemcpyHostToDevice));
    CMX_CUDA(cudaMemcpy(d_idx, &cu, sizeof(uint32_t), cudaMemcpyHostToDevice));
    for (int rev = 0; rev < 2; ++rev)
      k_build_cell_lut<<<8, 128, 0, s->stream>>>(T, nocc, z, d_fwd, 1, d_idx, d_one, rev, d_lut + 1024 * rev);
    CMX_CUDA(cudaGetLastError());
    CMX_CUDA(cudaMemcpyAsync(l1.data(), d_lut, sizeof(double) * 1024, cudaMemcpyDeviceToHost, s->stream));
    CMX_CUDA(cudaMemcpyAsync(l2.data(), d_lut + 1024, sizeof(double) * 1024, cudaMemcpyDeviceToHost, s->stream));
    CMX_CUDA(cudaStreamSynchronize(s->stream));
    double scale = 0.0;
    for (double x : l1) scale = std::max(scale, std::fabs(x));
    for (int q = 0; q < 1024; ++q)
      if (std::fabs(l1[q] - l2[q]) > 1e-13 * std::max(scale, 1e-300)) ok = false;
    for (int o = 0; o < nocc && ok; ++o) {
      const double c0 = l1[o << 8];
      const double d1 = (z >= 1) ? l1[(o << 8) | 1] - c0 : 0.0;
      const double d2 = (nocc == 3 && z >= 1) ? l1[(o << 8) | CMX_VA_CODE] - c0 : 0.0;
      for (int n2 = 0; n2 <= (nocc == 3 ? z : 0); ++n2)
        for (int n1 = 0; n1 + n2 <= z; ++n1)
          if (std::fabs(c0 + n1 * d1 + n2 * d2 - l1[(o << 8) | (n1 + CMX_VA_CODE * n2)]) > 1e-12 * std::max(scale, 1e-300))
            ok = false;
      lin[(size_t)c * 9 + 3 * o] = c0;
      lin[(size_t)c * 9 + 3 * o + 1] = d1;
      lin[(size_t)c * 9 + 3 * o + 2] = d2;
    }
  }
  cudaFree(d_fwd);
  cudaFree(d_idx);
  cudaFree(d_one);
  cudaFree(d_lut);
  if (!ok) return CMX_OK;
  P.corr_n_masks = (int)masks.size();
  for (size_t m = 0; m < masks.size(); ++m) {
    P.corr_mask[m] = masks[m];
    P.corr_z[m] = z_of[m];
  }
  CMX_CUDA(cudaMalloc((void **)&P.d_corr_func_mask, sizeof(int32_t) * nc));
  CMX_CUDA(cudaMalloc((void **)&P.d_corr_lin, sizeof(double) * 9 * nc));
  CMX_CUDA(cudaMemcpy(P.d_corr_func_mask, func_mask.data(), sizeof(int32_t) * nc, cudaMemcpyHostToDevice));
  CMX_CUDA(cudaMemcpy(P.d_corr_lin, lin.data(), sizeof(double) * 9 * nc, cudaMemcpyHostToDevice));
  P.corr_nocc = nocc;
  P.corr_lin_state = 1;
  return CMX_OK;
}

// result in scratch (valid until the next call that uses the scratch), asynchronous
int cmx_global_corr_lin_device(cmx_state *s, int32_t replica, double **d_out) {
  SweepPlan &P = s->plan;
  const int nc = s->t->d.corr_size, M = P.corr_n_masks;
  EnergyArgs a = energy_args(s, replica);
  const int nb = (int)std::min<uint32_t>((a.n_items + 255) / 256, 148 * 4);
  const size_t b_sums = sizeof(LinSums) * (size_t)nb * M;
  const size_t off_out = (b_sums + 255) & ~(size_t)255;
  int rc = cmx_scratch(s, off_out + sizeof(double) * nc);
  if (rc) return rc;
  char *base = static_cast<char *>(s->d_scratch);
  // masks[0] (no neighbors) needs the occupant counts only: they are part of every pass, so
  // it shares the pass of masks[1] when there is one
  if (M == 3 && P.corr_mask[1] == kMaskFccFwd && P.corr_mask[2] == kMaskFcc2nnFwd && energy_warp_rows(a)) {
    // both FCC pair shells in one pass over the lattice
    a.mask = kMaskFccFwd;
    a.sums = reinterpret_cast<LinSums *>(base) + (size_t)nb;
    a.sums2 = reinterpret_cast<LinSums *>(base) + (size_t)2 * nb;
    if ((rc = launch_energy_lin(s, a, nb, 1, P.corr_nocc, kMaskFcc2nnFwd))) return rc;
  } else {
    for (int m = (M > 1 ? 1 : 0); m < M; ++m) {
      a.mask = P.corr_mask[m];
      a.sums = reinterpret_cast<LinSums *>(base) + (size_t)m * nb;
      if ((rc = launch_energy_lin(s, a, nb, 1, P.corr_nocc))) return rc;
    }
  }
  CorrLinFinalArgs f;
  f.sums = reinterpret_cast<const LinSums *>(base);
  f.nb = nb;
  f.n_masks = M;
  f.corr_size = nc;
  f.n_cells = (long long)s->g.n_cells;
  for (int m = 0; m < 4; ++m) f.z[m] = (m < M) ? P.corr_z[m] : 0;
  f.func_mask = P.d_corr_func_mask;
  f.lin = P.d_corr_lin;
  f.out = reinterpret_cast<double *>(base + off_out);
  // functions without neighbors (mask 0) read the counts of pass 1: their d1 = d2 = 0, only
  // N_o enters, which every pass counts
  f.mask0_from = (M > 1) ? 1 : 0;
  k_corr_lin_final<<<1, 256, 0, s->stream>>>(f);
  CMX_CUDA(cudaGetLastError());
  *d_out = f.out;
  return CMX_OK;
}
