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
#include <algorithm>
#include <cmath>

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
  const int8_t *occ;  // replica base
  Geom g;
  uint32_t mask;
  uint32_t W;       // 16-byte chunks per row
  FastDiv divW, divJ;
  uint32_t n_items;  // W * N1 * N2
  const double *lut;
  double *partial;  // [gridDim.x]
};

template <int NOCC>
__global__ void __launch_bounds__(256) k_energy_pair16(EnergyArgs a) {
  __shared__ double sh_lut[NOCC * 256];
  __shared__ double sh_red[256];
  for (int q = threadIdx.x; q < NOCC * 256; q += blockDim.x) sh_lut[q] = a.lut[q];
  __syncthreads();
  const Geom &g = a.g;
  const int8_t *base = a.occ;
  const uint32_t N0 = g.N0, N1 = g.N1, N2 = g.N2;
  const uint32_t layer = N0 * N1;
  const bool halo = g.halo != 0;
  const uint32_t mask = a.mask;
  const uint32_t mc = (mask >> 12) & 7u;
  double acc = 0.0;
  for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < a.n_items;
       item += gridDim.x * blockDim.x) {
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
    uint32_t A0[4] = {0, 0, 0, 0}, Am[4] = {0, 0, 0, 0}, Ap[4] = {0, 0, 0, 0}, C[4] = {0, 0, 0, 0};
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
    (void)mc;
    uint32_t cnt[4];
    cnt[0] = A0[0] + __funnelshift_l(sm, Am[0], 8) + __funnelshift_r(Ap[0], Ap[1], 8);
    cnt[1] = A0[1] + __funnelshift_l(Am[0], Am[1], 8) + __funnelshift_r(Ap[1], Ap[2], 8);
    cnt[2] = A0[2] + __funnelshift_l(Am[1], Am[2], 8) + __funnelshift_r(Ap[2], Ap[3], 8);
    cnt[3] = A0[3] + __funnelshift_l(Am[2], Am[3], 8) + __funnelshift_r(Ap[3], sp, 8);
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
  if (threadIdx.x == 0) a.partial[blockIdx.x] = sh_red[0];
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
  return CMX_OK;
}

int cmx_energy_fast(cmx_state *s, int32_t replica, double *E) {
  SweepPlan &P = s->plan;
  CMX_CUDA(cudaSetDevice(s->t->device));
  EnergyArgs a;
  a.occ = s->d_occ + (size_t)replica * s->g.rep_stride;
  a.g = s->g;
  a.mask = P.e_mask;
  a.W = s->g.N0 / 16;
  a.divW = make_fastdiv(a.W);
  a.divJ = make_fastdiv((uint32_t)s->g.N1);
  a.n_items = a.W * (uint32_t)s->g.N1 * (uint32_t)s->g.N2;
  a.lut = P.d_e_lut;
  int nb = (int)std::min<uint32_t>((a.n_items + 255) / 256, 148 * 8);
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
