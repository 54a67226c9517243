// Device-side samplers: the quantities the reference
// samples every `period` passes -- clex.formation_energy, potential_energy,
// mol_composition, param_composition, corr
// (src/casm/clexmonte/monte_calculator/sampling_functions.cc:37-54,57-90,121-139,
// 141-155,274-288; all per unit cell) -- are evaluated for EVERY replica on the
// device and appended to a device-resident series.  A run (cmx_sweep_run) is then
// one stream of sweep and sample launches with a single synchronisation at the
// end: no N-byte occupation download per sample (the reference copies nothing
// because it lives on the host; a naive port would move 134 MB per sample).
//
// The analysis functions (heat_capacity, mol_susc, param_susc, *_thermochem_susc:
// monte_calculator/analysis_functions.cc:43-173) are variances / covariances of
// these series times n_unitcells / (kB T^2) or / (kB T); they are evaluated from
// the series by the host layer, as the reference does after the run.
//
// One sample =
//   pair-LUT models:  ONE streaming pass (k_energy_lin16 over all replicas: integer
//                     bond and occupant counts from the same 16-byte chunks)
//   other models:     faithful global correlations per replica (k_global_corr) and
//                     the occupant histogram (k_composition)
//   then k_sample_finish: fixed-order reduction of the block partials,
//   n = counts -> species per unit cell, x = R^T (n - origin),
//   potential = E - n_cells * mu . x  (SemiGrandCanonicalCalculator.cc:171-179;
//   canonical: mu = 0, CanonicalCalculator.cc:126-133), one series row per replica.
#include <algorithm>
#include <cmath>
#include <vector>

#include "cmx_internal.cuh"

static int invalid(const std::string &msg) {
  cmx_set_error(msg);
  return CMX_ERR_INVALID;
}

struct cmx_sampler {
  cmx_state *s = nullptr;
  int32_t capacity = 0, n_samples = 0;
  int32_t n_species = 0, n_param = 0, corr_size = 0;  // corr_size = 0: corr not sampled
  int32_t nq = 0;
  int32_t nb = 0;  // blocks per replica of the fast pass
  bool fast = false;
  double *d_series = nullptr;  // [replica][capacity][nq]
  double *d_origin = nullptr;  // [n_species]
  double *d_Rt = nullptr;      // [n_param][n_species]
  double *d_mu = nullptr;      // [replica][n_param]
  int32_t *d_o2s = nullptr;    // [n_sublat][max_occ] species index or -1
  LinSums *d_sums = nullptr;   // [replica][nb] bond / occupant counts of the fast pass
  unsigned long long *d_counts = nullptr;  // [replica][n_sublat*max_occ]
  double *d_corr = nullptr;                // [replica][corr_size of the tables] (generic path / corr sampling)
};

struct FinishArgs {
  int32_t n_replicas, nb, nbins, max_occ, n_species, n_param, corr_size, t_corr_size, nq, capacity, sample;
  int32_t fast, bin0_missing, n_eci, z;
  long long n_cells;
  const LinSums *sums;
  const double *lin;
  unsigned long long *counts;
  const double *corr;
  const uint32_t *eci_idx;
  const double *eci_val;
  const int32_t *o2s;
  const double *origin, *Rt, *mu;
  double *series;
};

// one block per replica: the block sums the integer block records (any order gives the
// same result), thread 0 does the O(n_species * n_param) arithmetic in a fixed order
__global__ void k_sample_finish(FinishArgs a) {
  const int r = blockIdx.x;
  unsigned long long *cnt = a.counts + (size_t)r * a.nbins;
  __shared__ unsigned long long sh[6];
  unsigned long long v[6];
  if (a.fast) cmx_lin_reduce(a.sums + (size_t)r * a.nb, a.nb, sh, v);
  if (threadIdx.x) return;
  double E = 0.0;
  if (a.fast) {
    unsigned long long n_occ[3];
    E = cmx_lin_energy(v, a.n_cells, a.z, a.lin, n_occ);
    for (int q = 0; q < a.nbins && q < 3; ++q) cnt[q] = n_occ[q];
  } else {
    const double *corr = a.corr + (size_t)r * a.t_corr_size;
    for (int q = 0; q < a.n_eci; ++q) E += a.eci_val[q] * corr[a.eci_idx[q]];
    if (a.bin0_missing) {
      unsigned long long rest = (unsigned long long)a.n_cells;
      for (int q = 1; q < a.nbins; ++q) rest -= cnt[q];
      cnt[0] = rest;
    }
  }
  double *row = a.series + ((size_t)r * a.capacity + a.sample) * a.nq;
  const double nc = (double)a.n_cells;
  // CompositionCalculator::mean_num_each_component
  double *n = row + 2;
  for (int sp = 0; sp < a.n_species; ++sp) n[sp] = 0.0;
  for (int q = 0; q < a.nbins; ++q) {
    const int sp = a.o2s[q];
    if (sp >= 0) n[sp] += (double)cnt[q];
  }
  for (int sp = 0; sp < a.n_species; ++sp) n[sp] /= nc;
  // CompositionConverter::param_composition: x = R^T (n - origin)
  double *x = row + 2 + a.n_species;
  double dot = 0.0;
  for (int p = 0; p < a.n_param; ++p) {
    double xp = 0.0;
    for (int sp = 0; sp < a.n_species; ++sp) xp += a.Rt[p * a.n_species + sp] * (n[sp] - a.origin[sp]);
    x[p] = xp;
    dot += a.mu[(size_t)r * a.n_param + p] * xp;
  }
  const double pot = E - nc * dot;
  row[0] = E / nc;
  row[1] = pot / nc;
  if (a.corr_size) {
    const double *corr = a.corr + (size_t)r * a.t_corr_size;
    double *out = row + 2 + a.n_species + a.n_param;
    for (int c = 0; c < a.corr_size; ++c) out[c] = corr[c] / nc;
  }
}

extern "C" int cmx_sampler_create(cmx_state *s, int32_t capacity, int32_t n_param, const double *origin,
                                  const double *Rt, int32_t with_corr, cmx_sampler **out) {
  if (!s || !out) return invalid("cmx_sampler_create: null argument");
  if (capacity < 1) return invalid("cmx_sampler_create: capacity < 1");
  if (!s->d_eci_idx) {
    cmx_set_error("cmx_sampler_create: no ECI bound (call cmx_state_set_eci)");
    return CMX_ERR_STATE;
  }
  if (s->n_species < 1 || s->occ_to_species.empty()) {
    cmx_set_error("cmx_sampler_create: occupants unknown (call cmx_state_set_occupants)");
    return CMX_ERR_STATE;
  }
  if (n_param < 0 || (n_param > 0 && (!origin || !Rt))) return invalid("cmx_sampler_create: composition axes missing");
  if (s->g.halo) return invalid("cmx_sampler_create: slab states are sampled by the host layer (per-rank partial sums)");
  CMX_CUDA(cudaSetDevice(s->t->device));
  const DevTables &T = s->t->d;
  cmx_sampler *m = new cmx_sampler();
  m->s = s;
  m->capacity = capacity;
  m->n_species = s->n_species;
  m->n_param = n_param;
  m->corr_size = with_corr ? T.corr_size : 0;
  m->nq = 2 + m->n_species + m->n_param + m->corr_size;
  const int nbins = T.n_sublat * T.max_occ;
  const int R = s->n_replicas;
  auto fail = [&](cudaError_t e) {
    cmx_set_error(std::string("cmx_sampler_create: ") + cudaGetErrorString(e));
    cmx_sampler_destroy(m);
    return CMX_ERR_CUDA;
  };
  cudaError_t e;
  if ((e = cudaMalloc((void **)&m->d_series, sizeof(double) * (size_t)R * capacity * m->nq)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void **)&m->d_origin, sizeof(double) * std::max(1, m->n_species))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void **)&m->d_Rt, sizeof(double) * std::max(1, n_param * m->n_species))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void **)&m->d_mu, sizeof(double) * std::max(1, R * n_param))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void **)&m->d_o2s, sizeof(int32_t) * nbins)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void **)&m->d_counts, sizeof(unsigned long long) * (size_t)R * nbins)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void **)&m->d_corr, sizeof(double) * (size_t)R * T.corr_size)) != cudaSuccess) return fail(e);
  std::vector<double> org(std::max(1, m->n_species), 0.0);
  if (origin) std::copy(origin, origin + m->n_species, org.begin());
  CMX_CUDA(cudaMemcpy(m->d_origin, org.data(), sizeof(double) * m->n_species, cudaMemcpyHostToDevice));
  if (n_param) CMX_CUDA(cudaMemcpy(m->d_Rt, Rt, sizeof(double) * n_param * m->n_species, cudaMemcpyHostToDevice));
  CMX_CUDA(cudaMemset(m->d_mu, 0, sizeof(double) * std::max(1, R * n_param)));
  CMX_CUDA(cudaMemcpy(m->d_o2s, s->occ_to_species.data(), sizeof(int32_t) * nbins, cudaMemcpyHostToDevice));
  *out = m;
  return CMX_OK;
}

extern "C" int cmx_sampler_destroy(cmx_sampler *m) {
  if (!m) return CMX_OK;
  cudaFree(m->d_series);
  cudaFree(m->d_origin);
  cudaFree(m->d_Rt);
  cudaFree(m->d_mu);
  cudaFree(m->d_o2s);
  cudaFree(m->d_sums);
  cudaFree(m->d_counts);
  cudaFree(m->d_corr);
  delete m;
  return CMX_OK;
}

extern "C" int cmx_sampler_set_param_chem_pot(cmx_sampler *m, int32_t replica, const double *param_chem_pot) {
  if (!m) return invalid("cmx_sampler_set_param_chem_pot: null sampler");
  if (replica < 0 || replica >= m->s->n_replicas) return invalid("cmx_sampler_set_param_chem_pot: replica out of range");
  if (m->n_param == 0) return CMX_OK;
  std::vector<double> mu(m->n_param, 0.0);  // null: canonical potential (no mu . x term)
  if (param_chem_pot) std::copy(param_chem_pot, param_chem_pot + m->n_param, mu.begin());
  CMX_CUDA(cudaSetDevice(m->s->t->device));
  // ordered after the samples already enqueued
  CMX_CUDA(cudaMemcpyAsync(m->d_mu + (size_t)replica * m->n_param, mu.data(), sizeof(double) * m->n_param,
                           cudaMemcpyHostToDevice, m->s->stream));
  CMX_CUDA(cudaStreamSynchronize(m->s->stream));
  return CMX_OK;
}

extern "C" int cmx_sampler_info(const cmx_sampler *m, int32_t *n_quantities, int32_t *n_species, int32_t *n_param,
                                int32_t *corr_size, int32_t *n_samples, int32_t *capacity) {
  if (!m) return invalid("cmx_sampler_info: null sampler");
  if (n_quantities) *n_quantities = m->nq;
  if (n_species) *n_species = m->n_species;
  if (n_param) *n_param = m->n_param;
  if (corr_size) *corr_size = m->corr_size;
  if (n_samples) *n_samples = m->n_samples;
  if (capacity) *capacity = m->capacity;
  return CMX_OK;
}

extern "C" int cmx_sampler_reset(cmx_sampler *m) {
  if (!m) return invalid("cmx_sampler_reset: null sampler");
  m->n_samples = 0;
  return CMX_OK;
}

// asynchronous: one series row per replica, on the state's stream
extern "C" int cmx_sampler_sample(cmx_sampler *m) {
  if (!m) return invalid("cmx_sampler_sample: null sampler");
  cmx_state *s = m->s;
  if (m->n_samples >= m->capacity) {
    cmx_set_error("cmx_sampler_sample: series full (read and reset the sampler)");
    return CMX_ERR_STATE;
  }
  if (!s->d_eci_idx) {
    cmx_set_error("cmx_sampler_sample: no ECI bound");
    return CMX_ERR_STATE;
  }
  CMX_CUDA(cudaSetDevice(s->t->device));
  const DevTables &T = s->t->d;
  const int R = s->n_replicas;
  const int nbins = T.n_sublat * T.max_occ;
  const bool fast = s->plan.valid && s->plan.e_fast && s->plan.e_lin && T.n_sublat == 1 && T.max_occ <= 3 &&
                    !(s->sweep_flags & CMX_SWEEP_FORCE_GENERIC) && (s->g.coded || T.max_occ <= 2);
  FinishArgs a = {};
  a.bin0_missing = 0;
  if (fast) {
    const int nb = cmx_energy_fast_blocks(s);
    if (nb != m->nb || !m->d_sums) {
      cudaFree(m->d_sums);
      m->d_sums = nullptr;
      CMX_CUDA(cudaMalloc((void **)&m->d_sums, sizeof(LinSums) * (size_t)R * nb));
      m->nb = nb;
    }
    int rc = cmx_energy_lin_batch(s, nb, m->d_sums);
    if (rc) return rc;
  }
  if (!fast || m->corr_size) {
    for (int r = 0; r < R; ++r) {
      double *d_out = nullptr;
      int rc = cmx_global_corr_device(s, r, &d_out);
      if (rc) return rc;
      CMX_CUDA(cudaMemcpyAsync(m->d_corr + (size_t)r * T.corr_size, d_out, sizeof(double) * T.corr_size,
                               cudaMemcpyDeviceToDevice, s->stream));
    }
  }
  if (!fast) {
    for (int r = 0; r < R; ++r) {
      bool missing = false;
      int rc = cmx_composition_device(s, r, m->d_counts + (size_t)r * nbins, &missing);
      if (rc) return rc;
      a.bin0_missing = missing ? 1 : 0;
    }
  }
  a.n_replicas = R;
  a.nb = m->nb;
  a.nbins = nbins;
  a.max_occ = T.max_occ;
  a.n_species = m->n_species;
  a.n_param = m->n_param;
  a.corr_size = m->corr_size;
  a.t_corr_size = T.corr_size;
  a.nq = m->nq;
  a.capacity = m->capacity;
  a.sample = m->n_samples;
  a.fast = fast ? 1 : 0;
  a.n_eci = s->n_eci;
  a.n_cells = (long long)s->g.n_cells;
  a.sums = m->d_sums;
  a.lin = s->plan.d_e_lin;
  a.z = s->plan.e_z;
  a.counts = m->d_counts;
  a.corr = m->d_corr;
  a.eci_idx = s->d_eci_idx;
  a.eci_val = s->d_eci_val;
  a.o2s = m->d_o2s;
  a.origin = m->d_origin;
  a.Rt = m->d_Rt;
  a.mu = m->d_mu;
  a.series = m->d_series;
  k_sample_finish<<<R, 128, 0, s->stream>>>(a);
  CMX_CUDA(cudaGetLastError());
  m->fast = fast;
  m->n_samples += 1;
  return CMX_OK;
}

// rows [first, first + n) of one replica's series; synchronises the stream
extern "C" int cmx_sampler_read(cmx_sampler *m, int32_t replica, int32_t first, int32_t n, double *out) {
  if (!m || !out) return invalid("cmx_sampler_read: null argument");
  if (replica < 0 || replica >= m->s->n_replicas) return invalid("cmx_sampler_read: replica out of range");
  if (first < 0 || n < 0 || first + n > m->n_samples) return invalid("cmx_sampler_read: rows out of range");
  CMX_CUDA(cudaSetDevice(m->s->t->device));
  if (n)
    CMX_CUDA(cudaMemcpyAsync(out, m->d_series + ((size_t)replica * m->capacity + first) * m->nq,
                             sizeof(double) * (size_t)n * m->nq, cudaMemcpyDeviceToHost, m->s->stream));
  CMX_CUDA(cudaStreamSynchronize(m->s->stream));
  return CMX_OK;
}

// occupation_metropolis_v2 with a sampling fixture of period `sweeps_per_sample`
// passes (methods/occupation_metropolis.hh:92-120: sample_data_by_count_if_due after
// every step; one pass = one attempted step per mutable site = one sweep), as ONE stream
// of launches: n_samples x (sweeps_per_sample sweeps, one sample), one synchronisation.
//   ensemble 0: semi-grand canonical (cmx_sgc_sweep), 1: canonical (cmx_canonical_sweep)
extern "C" int cmx_sweep_run(cmx_state *s, cmx_sampler *m, int32_t ensemble, int64_t n_samples,
                             int64_t sweeps_per_sample, uint64_t seed, int64_t first_sweep,
                             cmx_counters *counters) {
  if (!s || !m) return invalid("cmx_sweep_run: null argument");
  if (m->s != s) return invalid("cmx_sweep_run: the sampler belongs to another state");
  if (ensemble != 0 && ensemble != 1) return invalid("cmx_sweep_run: ensemble must be 0 (semi-grand) or 1 (canonical)");
  if (n_samples < 0 || sweeps_per_sample < 0) return invalid("cmx_sweep_run: negative count");
  if (m->n_samples + n_samples > m->capacity) {
    cmx_set_error("cmx_sweep_run: the series does not hold that many samples");
    return CMX_ERR_STATE;
  }
  if (s->g.halo) return invalid("cmx_sweep_run: slab states are driven by cmx_sgc_sweep_kgroup");
  int rc;
  if (ensemble == 0) {
    if ((rc = cmx_counters_reset(s))) return rc;
  } else if (!s->canon) {
    cmx_set_error("cmx_sweep_run: no swap types (cmx_canonical_set_swaps)");
    return CMX_ERR_STATE;
  }
  int64_t sweep = first_sweep;
  for (int64_t k = 0; k < n_samples; ++k) {
    if (ensemble == 0) {
      if ((rc = cmx_sgc_sweep_enqueue(s, seed, sweep, sweeps_per_sample))) return rc;
    } else {
      if ((rc = cmx_canonical_enqueue(s, sweeps_per_sample, seed, sweep, k == 0))) return rc;
    }
    sweep += sweeps_per_sample;
    if ((rc = cmx_sampler_sample(m))) return rc;
  }
  if (ensemble == 0) {
    if (counters) return cmx_counters_read(s, counters);
    CMX_CUDA(cudaStreamSynchronize(s->stream));
    return CMX_OK;
  }
  if (n_samples == 0) {
    if (counters)
      for (int r = 0; r < s->n_replicas; ++r) counters[r] = cmx_counters{0, 0, 0.0, 0};
    return CMX_OK;
  }
  return cmx_canonical_counters(s, counters);
}

// ---- moments of the sampled series (replica grids: statistics are reduced, not series) ----
// Per replica, over the samples [first, n_samples) and the Q = 2 + n_species + n_param scalar
// quantities q = (clex.formation_energy, potential_energy, mol_composition, param_composition):
//   out[replica] = { n, sum_a q_a (Q), sum q_a q_b (Q x Q) }          M = 1 + Q + Q^2 doubles
// -- everything heat_capacity, mol_susc, param_susc and the thermochemical susceptibilities
// (analysis_functions.cc:43-173, covariances per run/io/convariance_functions.cc:29-143) need,
// and ADDITIVE over disjoint sets of samples, so that ranks holding different replicas (or
// different sample ranges of one replica) combine them with one all-reduce.  One thread per
// entry sums its samples in ascending order: deterministic.
__global__ void k_sampler_moments(const double *__restrict__ series, int capacity, int nq, int Q, int first, int n_samples,
                                  double *__restrict__ out) {
  const int r = blockIdx.x, M = 1 + Q + Q * Q;
  const double *ser = series + (size_t)r * capacity * nq;
  for (int e = threadIdx.x; e < M; e += blockDim.x) {
    double acc = 0.0;
    if (e == 0) {
      acc = (double)(n_samples - first);
    } else if (e <= Q) {
      for (int t = first; t < n_samples; ++t) acc += ser[(size_t)t * nq + (e - 1)];
    } else {
      const int a = (e - 1 - Q) / Q, b = (e - 1 - Q) % Q;
      for (int t = first; t < n_samples; ++t) acc += ser[(size_t)t * nq + a] * ser[(size_t)t * nq + b];
    }
    out[(size_t)r * M + e] = acc;
  }
}

extern "C" int cmx_sampler_moments(cmx_sampler *m, int32_t first, double *out, int32_t out_on_device) {
  if (!m || !out) return invalid("cmx_sampler_moments: null argument");
  if (first < 0 || first > m->n_samples) return invalid("cmx_sampler_moments: first sample out of range");
  cmx_state *s = m->s;
  CMX_CUDA(cudaSetDevice(s->t->device));
  const int Q = 2 + m->n_species + m->n_param, M = 1 + Q + Q * Q, R = s->n_replicas;
  double *d_out = out;
  if (!out_on_device) {
    int rc = cmx_scratch(s, sizeof(double) * (size_t)R * M);
    if (rc) return rc;
    d_out = (double *)s->d_scratch;
  }
  k_sampler_moments<<<R, 128, 0, s->stream>>>(m->d_series, m->capacity, m->nq, Q, first, m->n_samples, d_out);
  CMX_CUDA(cudaGetLastError());
  if (!out_on_device) {
    CMX_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)R * M, cudaMemcpyDeviceToHost, s->stream));
    CMX_CUDA(cudaStreamSynchronize(s->stream));
  }
  return CMX_OK;  // out_on_device: asynchronous on the state's stream
}
