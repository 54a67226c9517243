// Faithful (reference-order, individually rounded) correlation / delta-E
// kernels: the table-driven replacement of the generated Clexulator methods
//   _calc_delta_point_corr / _calc_restricted_delta_point_corr
//       (FCC_binary_vacancy_Clexulator_default.cc:555-610)
//   _calc_point_corr (:500-553), _calc_global_corr_contribution (:446-498)
// and of the [EXT] wrappers ClusterExpansion::occ_delta_value / per_supercell
// (call sites SemiGrandCanonicalCalculator.cc:171-213).
#include <algorithm>

#include "cmx_internal.cuh"

static int invalid(const std::string &msg) {
  cmx_set_error(msg);
  return CMX_ERR_INVALID;
}

__device__ __forceinline__ void cell_ijk(const Geom &g, int64_t cell, int &i,
                                         int &j, int &k) {
  i = (int)(cell % g.N0);
  int64_t r = cell / g.N0;
  j = (int)(r % g.N1);
  k = (int)(r / g.N1);
}

// position of sublattice b among the neighbor-list sublattices
__device__ __forceinline__ int point_index(const DevTables &T, int b) {
  for (int p = 0; p < T.n_nlist_sublat; ++p)
    if (T.nlist_sublat[p] == b) return p;
  return -1;
}

// kind 0: delta corr (needs new_occ), kind 1: point corr, kind 2: cell corr.
// One thread per (item, corr index).
__global__ void k_corr_batch(DevTables T, Geom g, const int8_t *__restrict__ occ,
                             int kind, int64_t n, const int64_t *__restrict__ l,
                             const int32_t *__restrict__ new_occ,
                             double *__restrict__ out) {
  int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (x >= n * T.corr_size) return;
  int64_t item = x / T.corr_size;
  int c = (int)(x - item * T.corr_size);
  Override ov;
  ov.n = 0;
  int i, j, k;
  if (kind == 2) {
    cell_ijk(g, l[item], i, j, k);
    out[x] = cmx_eval_function(T, g, occ, T.global_gbeg[c], T.global_gbeg[c + 1],
                               i, j, k, ov, 0, 0, 0);
    return;
  }
  int64_t li = l[item];
  int b = (int)(li / g.n_cells);
  cell_ijk(g, li - (int64_t)b * g.n_cells, i, j, k);
  int p = point_index(T, b);
  if (p < 0) {
    out[x] = 0.0;
    return;
  }
  int fi = p * T.corr_size + c;
  if (kind == 1) {
    out[x] = cmx_eval_function(T, g, occ, T.point_gbeg[fi], T.point_gbeg[fi + 1],
                               i, j, k, ov, b, 0, 0);
  } else {
    int oi = cmx_dec(occ[cmx_site_offset(g, b, i, j, k)]);
    out[x] = cmx_eval_function(T, g, occ, T.delta_gbeg[fi], T.delta_gbeg[fi + 1],
                               i, j, k, ov, b, oi, new_occ[item]);
  }
}

static int run_corr_batch(const cmx_state *cs, int32_t replica, int kind,
                          int64_t n, const int64_t *l, const int32_t *new_occ,
                          double *out, const char *who) {
  cmx_state *s = const_cast<cmx_state *>(cs);
  if (!s) return invalid(std::string(who) + ": null state");
  if (replica < 0 || replica >= s->n_replicas)
    return invalid(std::string(who) + ": replica out of range");
  if (n < 0 || (n && (!l || !out)) || (kind == 0 && n && !new_occ))
    return invalid(std::string(who) + ": bad argument");
  if (n == 0) return CMX_OK;
  const DevTables &T = s->t->d;
  int64_t n_sites = s->g.n_cells * T.n_sublat;
  for (int64_t q = 0; q < n; ++q) {
    if (kind == 2) {
      if (l[q] < 0 || l[q] >= s->g.n_cells)
        return invalid(std::string(who) + ": unit cell index out of range");
    } else {
      if (l[q] < 0 || l[q] >= n_sites)
        return invalid(std::string(who) + ": linear site index out of range");
      if (kind == 0) {
        int b = (int)(l[q] / s->g.n_cells);
        if (new_occ[q] < 0 || new_occ[q] >= s->t->n_occ[b])
          return invalid(std::string(who) + ": new occupant index out of range");
      }
    }
  }
  CMX_CUDA(cudaSetDevice(s->t->device));
  size_t b_l = sizeof(int64_t) * n, b_o = sizeof(int32_t) * n;
  size_t b_out = sizeof(double) * n * T.corr_size;
  size_t off_o = (b_l + 255) & ~(size_t)255;
  size_t off_out = (off_o + b_o + 255) & ~(size_t)255;
  int rc = cmx_scratch(s, off_out + b_out);
  if (rc) return rc;
  char *base = static_cast<char *>(s->d_scratch);
  CMX_CUDA(cudaMemcpyAsync(base, l, b_l, cudaMemcpyHostToDevice, s->stream));
  if (kind == 0)
    CMX_CUDA(cudaMemcpyAsync(base + off_o, new_occ, b_o, cudaMemcpyHostToDevice, s->stream));
  int64_t total = n * T.corr_size;
  int threads = 128;
  int64_t blocks = (total + threads - 1) / threads;
  k_corr_batch<<<(unsigned)blocks, threads, 0, s->stream>>>(
      T, s->g, s->d_occ + (size_t)replica * s->g.rep_stride, kind, n,
      (const int64_t *)base, (const int32_t *)(base + off_o),
      (double *)(base + off_out));
  CMX_CUDA(cudaGetLastError());
  CMX_CUDA(cudaMemcpyAsync(out, base + off_out, b_out, cudaMemcpyDeviceToHost, s->stream));
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  return CMX_OK;
}

extern "C" int cmx_delta_corr(const cmx_state *s, int32_t replica, int64_t n,
                              const int64_t *l, const int32_t *new_occ, double *out) {
  return run_corr_batch(s, replica, 0, n, l, new_occ, out, "cmx_delta_corr");
}
extern "C" int cmx_point_corr(const cmx_state *s, int32_t replica, int64_t n,
                              const int64_t *l, double *out) {
  return run_corr_batch(s, replica, 1, n, l, nullptr, out, "cmx_point_corr");
}
extern "C" int cmx_cell_corr(const cmx_state *s, int32_t replica, int64_t n,
                             const int64_t *cell, double *out) {
  return run_corr_batch(s, replica, 2, n, cell, nullptr, out, "cmx_cell_corr");
}

// ---------------------------------------------------------------------------
// ClusterExpansion::occ_delta_value for batches of (multi-site) events.
// Stage 1: one thread per (event, eci slot): sum over the event's sites of the
// restricted delta corr, site k evaluated with sites 0..k-1 already changed.
// Stage 2: one thread per event: ordered dot product with the coefficients,
// then the semi-grand exchange term.
// ---------------------------------------------------------------------------
__global__ void k_event_dcorr(DevTables T, Geom g, const int8_t *__restrict__ occ,
                              int64_t n, int sites_per_event,
                              const int64_t *__restrict__ l,
                              const int32_t *__restrict__ new_occ, int n_eci,
                              const uint32_t *__restrict__ eci_idx,
                              double *__restrict__ dcorr) {
  int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (x >= n * n_eci) return;
  int64_t ev = x / n_eci;
  int c = (int)eci_idx[x - ev * n_eci];
  Override ov;
  ov.n = 0;
  double acc = 0.0;
  for (int q = 0; q < sites_per_event; ++q) {
    int64_t li = l[ev * sites_per_event + q];
    int no = new_occ[ev * sites_per_event + q];
    int b = (int)(li / g.n_cells);
    int i, j, k;
    cell_ijk(g, li - (int64_t)b * g.n_cells, i, j, k);
    int64_t off = cmx_site_offset(g, b, i, j, k);
    int oi = cmx_load_occ(occ, off, ov);
    int p = point_index(T, b);
    double d = 0.0;
    if (p >= 0) {
      int fi = p * T.corr_size + c;
      d = cmx_eval_function(T, g, occ, T.delta_gbeg[fi], T.delta_gbeg[fi + 1], i,
                            j, k, ov, b, oi, no);
    }
    acc = (q == 0) ? d : __dadd_rn(acc, d);
    if (q < 4) {
      ov.off[ov.n] = off;
      ov.occ[ov.n] = no;
      ov.n++;
    }
  }
  dcorr[x] = acc;
}

__global__ void k_event_energy(Geom g, const int8_t *__restrict__ occ, int64_t n,
                               int sites_per_event, const int64_t *__restrict__ l,
                               const int32_t *__restrict__ new_occ, int n_eci,
                               const double *__restrict__ eci_val,
                               const double *__restrict__ dcorr,
                               const double *__restrict__ exch, int max_occ,
                               double *__restrict__ dE) {
  int64_t ev = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (ev >= n) return;
  double e = 0.0;
  for (int q = 0; q < n_eci; ++q)
    e = __dadd_rn(e, __dmul_rn(eci_val[q], dcorr[ev * n_eci + q]));
  if (exch) {
    // mu . (R^T dN): each site contributes exch[b][occ_i][occ_f]
    double x = 0.0;
    for (int q = 0; q < sites_per_event; ++q) {
      int64_t li = l[ev * sites_per_event + q];
      int b = (int)(li / g.n_cells);
      int i, j, k;
      cell_ijk(g, li - (int64_t)b * g.n_cells, i, j, k);
      int oi = cmx_dec(occ[cmx_site_offset(g, b, i, j, k)]);
      double t = exch[(b * max_occ + oi) * max_occ + new_occ[ev * sites_per_event + q]];
      x = (q == 0) ? t : __dadd_rn(x, t);
    }
    e = __dsub_rn(e, x);
  }
  dE[ev] = e;
}

extern "C" int cmx_delta_e(const cmx_state *cs, int32_t replica, int64_t n,
                           int32_t spe, const int64_t *l, const int32_t *new_occ,
                           int32_t potential, double *dE) {
  cmx_state *s = const_cast<cmx_state *>(cs);
  if (!s) return invalid("cmx_delta_e: null state");
  if (replica < 0 || replica >= s->n_replicas)
    return invalid("cmx_delta_e: replica out of range");
  if (spe < 1 || spe > 4) return invalid("cmx_delta_e: sites_per_event must be 1..4");
  if (n < 0 || (n && (!l || !new_occ || !dE))) return invalid("cmx_delta_e: bad argument");
  if (!s->d_eci_idx) {
    cmx_set_error("cmx_delta_e: no ECI bound (call cmx_state_set_eci)");
    return CMX_ERR_STATE;
  }
  if (n == 0) return CMX_OK;
  const DevTables &T = s->t->d;
  int64_t n_sites = s->g.n_cells * T.n_sublat;
  for (int64_t q = 0; q < n * spe; ++q) {
    if (l[q] < 0 || l[q] >= n_sites)
      return invalid("cmx_delta_e: linear site index out of range");
    int b = (int)(l[q] / s->g.n_cells);
    if (new_occ[q] < 0 || new_occ[q] >= s->t->n_occ[b])
      return invalid("cmx_delta_e: new occupant index out of range");
  }
  CMX_CUDA(cudaSetDevice(s->t->device));
  int ne = s->n_eci;
  size_t b_l = sizeof(int64_t) * n * spe, b_o = sizeof(int32_t) * n * spe;
  size_t b_dc = sizeof(double) * n * (ne ? ne : 1), b_e = sizeof(double) * n;
  size_t off_o = (b_l + 255) & ~(size_t)255;
  size_t off_dc = (off_o + b_o + 255) & ~(size_t)255;
  size_t off_e = (off_dc + b_dc + 255) & ~(size_t)255;
  int rc = cmx_scratch(s, off_e + b_e);
  if (rc) return rc;
  char *base = static_cast<char *>(s->d_scratch);
  const int8_t *occ = s->d_occ + (size_t)replica * s->g.rep_stride;
  CMX_CUDA(cudaMemcpyAsync(base, l, b_l, cudaMemcpyHostToDevice, s->stream));
  CMX_CUDA(cudaMemcpyAsync(base + off_o, new_occ, b_o, cudaMemcpyHostToDevice, s->stream));
  int threads = 128;
  if (ne) {
    int64_t total = n * ne;
    k_event_dcorr<<<(unsigned)((total + threads - 1) / threads), threads, 0, s->stream>>>(
        T, s->g, occ, n, spe, (const int64_t *)base, (const int32_t *)(base + off_o),
        ne, s->d_eci_idx, (double *)(base + off_dc));
    CMX_CUDA(cudaGetLastError());
  }
  size_t ex = (size_t)T.n_sublat * T.max_occ * T.max_occ;
  k_event_energy<<<(unsigned)((n + threads - 1) / threads), threads, 0, s->stream>>>(
      s->g, occ, n, spe, (const int64_t *)base, (const int32_t *)(base + off_o), ne,
      s->d_eci_val, (const double *)(base + off_dc),
      potential ? s->d_exch + replica * ex : nullptr, T.max_occ,
      (double *)(base + off_e));
  CMX_CUDA(cudaGetLastError());
  CMX_CUDA(cudaMemcpyAsync(dE, base + off_e, b_e, cudaMemcpyDeviceToHost, s->stream));
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  return CMX_OK;
}

// ---------------------------------------------------------------------------
// Correlations::per_supercell(): sum of the per-cell global contribution over
// all unit cells.  Each thread owns a strided set of cells and one correlation
// function; block partials are reduced in a fixed order (deterministic).
// grid = (blocks_x, corr_size)
// ---------------------------------------------------------------------------
__global__ void k_global_corr(DevTables T, Geom g, const int8_t *__restrict__ occ,
                              double *__restrict__ partial) {
  int c = blockIdx.y;
  int gb = T.global_gbeg[c], ge = T.global_gbeg[c + 1];
  Override ov;
  ov.n = 0;
  double acc = 0.0;
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
       cell < g.n_cells; cell += (int64_t)gridDim.x * blockDim.x) {
    int i, j, k;
    cell_ijk(g, cell, i, j, k);
    acc += cmx_eval_function(T, g, occ, gb, ge, i, j, k, ov, 0, 0, 0);
  }
  // fixed-order block reduction
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(int64_t)c * gridDim.x + blockIdx.x] = sh[0];
}

__global__ void k_reduce_partials(const double *__restrict__ partial, int nb,
                                  int corr_size, double *__restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= corr_size) return;
  double a = 0.0;
  for (int q = 0; q < nb; ++q) a += partial[(int64_t)c * nb + q];
  out[c] = a;
}

int cmx_global_corr_device(cmx_state *s, int32_t replica, double **d_out) {
  const DevTables &T = s->t->d;
  // point + pair bases on one sublattice: integer bond counts per forward-neighbor set
  if (!(s->sweep_flags & CMX_SWEEP_FORCE_GENERIC)) {
    int rc = cmx_plan_corr_lin(s);
    if (rc) return rc;
    if (s->plan.corr_lin_state == 1) return cmx_global_corr_lin_device(s, replica, d_out);
  }
  int nb = (int)((s->g.n_cells + 255) / 256);
  if (nb > 592) nb = 592;
  size_t b_part = sizeof(double) * (size_t)nb * T.corr_size;
  size_t off_out = (b_part + 255) & ~(size_t)255;
  int rc = cmx_scratch(s, off_out + sizeof(double) * T.corr_size);
  if (rc) return rc;
  char *base = static_cast<char *>(s->d_scratch);
  dim3 grid(nb, T.corr_size);
  k_global_corr<<<grid, 256, 0, s->stream>>>(
      T, s->g, s->d_occ + (size_t)replica * s->g.rep_stride, (double *)base);
  CMX_CUDA(cudaGetLastError());
  k_reduce_partials<<<(T.corr_size + 127) / 128, 128, 0, s->stream>>>(
      (const double *)base, nb, T.corr_size, (double *)(base + off_out));
  CMX_CUDA(cudaGetLastError());
  *d_out = (double *)(base + off_out);
  return CMX_OK;
}

extern "C" int cmx_global_corr(const cmx_state *cs, int32_t replica, double *out) {
  cmx_state *s = const_cast<cmx_state *>(cs);
  if (!s || !out) return invalid("cmx_global_corr: null argument");
  if (replica < 0 || replica >= s->n_replicas)
    return invalid("cmx_global_corr: replica out of range");
  CMX_CUDA(cudaSetDevice(s->t->device));
  double *d_out = nullptr;
  int rc = cmx_global_corr_device(s, replica, &d_out);
  if (rc) return rc;
  CMX_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * s->t->d.corr_size,
                           cudaMemcpyDeviceToHost, s->stream));
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  return CMX_OK;
}

extern "C" int cmx_energy(const cmx_state *cs, int32_t replica, double *E) {
  cmx_state *s = const_cast<cmx_state *>(cs);
  if (!s || !E) return invalid("cmx_energy: null argument");
  if (!s->d_eci_idx) {
    cmx_set_error("cmx_energy: no ECI bound (call cmx_state_set_eci)");
    return CMX_ERR_STATE;
  }
  if (replica < 0 || replica >= s->n_replicas) return invalid("cmx_energy: replica out of range");
  if (s->plan.e_fast && !(s->sweep_flags & CMX_SWEEP_FORCE_GENERIC)) return cmx_energy_fast(s, replica, E);
  std::vector<double> corr(s->t->d.corr_size);
  int rc = cmx_global_corr(s, replica, corr.data());
  if (rc) return rc;
  double e = 0.0;
  for (int i = 0; i < s->n_eci; ++i) e += s->eci_val[i] * corr[s->eci_idx[i]];
  *E = e;
  return CMX_OK;
}

// occupant histogram per sublattice
__global__ void k_composition(Geom g, int n_sublat, int max_occ,
                              const int8_t *__restrict__ occ,
                              unsigned long long *__restrict__ counts) {
  extern __shared__ unsigned int sh_cnt[];
  int nbins = n_sublat * max_occ;
  for (int q = threadIdx.x; q < nbins; q += blockDim.x) sh_cnt[q] = 0;
  __syncthreads();
  int64_t total = g.n_cells * n_sublat;
  for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < total;
       l += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = l / g.n_cells, cell = l - b * g.n_cells;
    int v = cmx_dec(occ[b * g.sub_stride + g.halo * g.layer + cell]);
    atomicAdd(&sh_cnt[b * max_occ + v], 1u);
  }
  __syncthreads();
  for (int q = threadIdx.x; q < nbins; q += blockDim.x)
    if (sh_cnt[q]) atomicAdd(&counts[q], (unsigned long long)sh_cnt[q]);
}

// single-sublattice states with <= 3 occupants (storage codes 0 / 1 / 18, or 0 / 1):
// 16 sites per load, occupants counted with two population counts per word
__global__ void k_composition_coded(const uint4 *__restrict__ occ, int64_t n16,
                                    unsigned long long *__restrict__ counts) {
  unsigned int n1 = 0, n2 = 0;
  for (int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; x < n16;
       x += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = occ[x];
    n1 += __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u) + __popc(v.z & 0x01010101u) +
          __popc(v.w & 0x01010101u);
    n2 += __popc(v.x & 0x10101010u) + __popc(v.y & 0x10101010u) + __popc(v.z & 0x10101010u) +
          __popc(v.w & 0x10101010u);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n1 += __shfl_down_sync(0xffffffffu, n1, o);
    n2 += __shfl_down_sync(0xffffffffu, n2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (n1) atomicAdd(&counts[1], (unsigned long long)n1);
    if (n2) atomicAdd(&counts[2], (unsigned long long)n2);
  }
}

// occupant counts of one replica into d_counts[n_sublat * max_occ] (asynchronous, on the
// state's stream); the coded path leaves bin 0 to the caller (n_cells - the others)
int cmx_composition_device(cmx_state *s, int32_t replica, unsigned long long *d_counts, bool *bin0_missing) {
  const DevTables &T = s->t->d;
  int nbins = T.n_sublat * T.max_occ;
  CMX_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * nbins, s->stream));
  int64_t total = s->g.n_cells * T.n_sublat;
  int nb = (int)((total + 255) / 256);
  if (nb > 1184) nb = 1184;
  if (T.n_sublat == 1 && T.max_occ <= 3 && s->g.n_cells % 16 == 0 && (s->g.coded || T.max_occ <= 2)) {
    // owned layers are contiguous (ghost layers sit before and after them)
    const int8_t *first = s->d_occ + (size_t)replica * s->g.rep_stride + (size_t)s->g.halo * s->g.layer;
    int64_t n16 = s->g.n_cells / 16;
    int nbc = (int)std::min<int64_t>((n16 + 255) / 256, 1184);
    k_composition_coded<<<nbc, 256, 0, s->stream>>>((const uint4 *)first, n16, d_counts);
    CMX_CUDA(cudaGetLastError());
    *bin0_missing = true;
    return CMX_OK;
  }
  k_composition<<<nb, 256, sizeof(unsigned int) * nbins, s->stream>>>(
      s->g, T.n_sublat, T.max_occ, s->d_occ + (size_t)replica * s->g.rep_stride, d_counts);
  CMX_CUDA(cudaGetLastError());
  *bin0_missing = false;
  return CMX_OK;
}

extern "C" int cmx_composition(const cmx_state *cs, int32_t replica, int64_t *counts) {
  cmx_state *s = const_cast<cmx_state *>(cs);
  if (!s || !counts) return invalid("cmx_composition: null argument");
  if (replica < 0 || replica >= s->n_replicas)
    return invalid("cmx_composition: replica out of range");
  CMX_CUDA(cudaSetDevice(s->t->device));
  const DevTables &T = s->t->d;
  int nbins = T.n_sublat * T.max_occ;
  int rc = cmx_scratch(s, sizeof(unsigned long long) * nbins);
  if (rc) return rc;
  bool bin0_missing = false;
  rc = cmx_composition_device(s, replica, (unsigned long long *)s->d_scratch, &bin0_missing);
  if (rc) return rc;
  CMX_CUDA(cudaMemcpyAsync(counts, s->d_scratch, sizeof(int64_t) * nbins,
                           cudaMemcpyDeviceToHost, s->stream));
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  if (bin0_missing) {
    int64_t rest = s->g.n_cells;
    for (int q = 1; q < nbins; ++q) rest -= counts[q];
    counts[0] = rest;
  }
  return CMX_OK;
}
