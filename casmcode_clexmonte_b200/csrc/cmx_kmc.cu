// Kinetic Monte Carlo event states, batched over (trajectory, unit cell, prim
// event) triples.  One thread evaluates one event exactly as
//   EventStateCalculator::calculate_event_state       (BaseMonteEventData.cc:87-119)
//   EventStateCalculator::_default_event_state_calculation (:122-156)
//   event_is_allowed                                  (events/event_methods.cc:340-351)
// do: allowed? -> dE_final = ECI . delta_corr(sites applied one after the other)
// -> local corr of the event's equivalent local clexulator -> E_kra, freq ->
// dE_activated = dE_final/2 + E_kra, clamped -> rate = freq exp(-beta dE_activated).
// All correlation arithmetic goes through the faithful evaluator, so dE_final,
// E_kra and freq are bit-identical to the reference's generated kernels.
#include <algorithm>
#include <cstring>

#include "cmx_internal.cuh"
#include "cmx_mt64.cuh"

static int invalid(const std::string &msg) {
  cmx_set_error(msg);
  return CMX_ERR_INVALID;
}

struct DevEventType {
  int32_t n_equivalents;
  int32_t table0;  // index of equivalent 0 in the DevTables array
  int32_t kra_beg, kra_end, freq_beg, freq_end;  // into coef_idx / coef_val
};

struct cmx_kmc {
  cmx_state *s;
  int32_t n_types, n_prim;
  std::vector<cmx_prim_event> prim;
  std::vector<DevEventType> types;
  cmx_prim_event *d_prim = nullptr;
  DevEventType *d_types = nullptr;
  DevTables *d_tables = nullptr;  // local clexulators, all equivalents of all types
  uint32_t *d_coef_idx = nullptr;
  double *d_coef_val = nullptr;
  double *d_rates = nullptr;  // [replica][cell][prim]
  double *d_total = nullptr;  // [replica]
  // rejection-free run (cmx_kmc_run_*)
  int32_t *d_imp_beg = nullptr;  // [n_prim + 1]
  int4 *d_imp = nullptr;         // (prim event, dx, dy, dz) relative to the event's unit cell
  int32_t max_imp = 0;
  std::vector<int64_t> level_size, level_off;  // sum tree above the leaves (= d_rates)
  double *d_tree = nullptr;      // [replica][upper nodes]
  Mt64 *d_mt = nullptr;          // [replica]
  double *d_time = nullptr;      // [replica]
  long long *d_pending = nullptr;  // [replica] event whose impact list is not yet applied, or -1
  long long *d_steps = nullptr;    // [replica] steps done
  bool run_ready = false;
};

struct KmcArgs {
  DevTables T;  // formation energy basis set
  Geom g;
  const int8_t *occ;  // replica 0
  int n_eci;
  const uint32_t *eci_idx;
  const double *eci_val;
  const double *beta;  // [replica]
  const cmx_prim_event *prim;
  const DevEventType *types;
  const DevTables *local;
  const uint32_t *coef_idx;
  const double *coef_val;
};

__device__ __forceinline__ int kmc_point_index(const DevTables &T, int b) {
  for (int p = 0; p < T.n_nlist_sublat; ++p)
    if (T.nlist_sublat[p] == b) return p;
  return -1;
}

// ---- the event state in three pieces, so that a thread block can spread the function
// evaluations of an allowed event over its threads (k_kmc_run) while a single thread
// (k_kmc_event_states, k_kmc_all_rates) runs the same pieces one after the other: the
// arithmetic, and therefore every bit of the result, is the same either way.
struct KmcEventSites {
  int ci, cj, ck;
  int si[CMX_EVENT_MAX_SITES], sj[CMX_EVENT_MAX_SITES], sk[CMX_EVENT_MAX_SITES];
};

// event_is_allowed (events/event_methods.cc:340-351): every site holds the event's
// initial occupant
__device__ __forceinline__ bool kmc_event_sites(const KmcArgs &a, const int8_t *occ, int64_t cell,
                                                const cmx_prim_event &E, KmcEventSites &S) {
  const Geom &g = a.g;
  S.ci = (int)(cell % g.N0);
  const int64_t rest = cell / g.N0;
  S.cj = (int)(rest % g.N1);
  S.ck = (int)(rest / g.N1);
  bool allowed = true;
  for (int q = 0; q < E.n_sites; ++q) {
    S.si[q] = S.ci + E.site[q][1];
    S.sj[q] = S.cj + E.site[q][2];
    S.sk[q] = S.ck + E.site[q][3];
    cmx_wrap_cell(g, S.si[q], S.sj[q], S.sk[q]);
    const int o = cmx_dec(occ[cmx_site_offset(g, E.site[q][0], S.si[q], S.sj[q], S.sk[q])]);
    if (o != E.occ_init[q]) allowed = false;
  }
  return allowed;
}

// tasks of an event: [0, n_eci * n_sites): delta corr of ECI function c = task / n_sites at
// site q = task % n_sites, evaluated with sites 0..q-1 already changed
// (Correlations::occ_delta [EXT]); then the kra functions, then the freq
// functions of the event's equivalent local clexulator about the unit cell
__device__ __forceinline__ int kmc_n_tasks(const KmcArgs &a, const cmx_prim_event &E) {
  const DevEventType &Y = a.types[E.event_type];
  return a.n_eci * E.n_sites + (Y.kra_end - Y.kra_beg) + (Y.freq_end - Y.freq_beg);
}
__device__ double kmc_task_value(const KmcArgs &a, const int8_t *occ, const cmx_prim_event &E,
                                 const KmcEventSites &S, int task) {
  const Geom &g = a.g;
  const DevTables &T = a.T;
  const int n_delta = a.n_eci * E.n_sites;
  if (task < n_delta) {
    const int c = task / E.n_sites, q = task - c * E.n_sites;
    Override ov;
    ov.n = 0;
    for (int q2 = 0; q2 < q; ++q2) {
      ov.off[ov.n] = cmx_site_offset(g, E.site[q2][0], S.si[q2], S.sj[q2], S.sk[q2]);
      ov.occ[ov.n] = E.occ_final[q2];
      ov.n++;
    }
    const int b = E.site[q][0];
    const int p = kmc_point_index(T, b);
    if (p < 0) return 0.0;
    const int fi = p * T.corr_size + (int)a.eci_idx[c];
    return cmx_eval_function(T, g, occ, T.delta_gbeg[fi], T.delta_gbeg[fi + 1], S.si[q], S.sj[q], S.sk[q], ov, b,
                             E.occ_init[q], E.occ_final[q]);
  }
  const DevEventType &Y = a.types[E.event_type];
  const DevTables &L = a.local[Y.table0 + E.equivalent_index];
  Override none;
  none.n = 0;
  const int c = (int)a.coef_idx[Y.kra_beg + (task - n_delta)];  // kra and freq entries are contiguous
  return cmx_eval_function(L, g, occ, L.global_gbeg[c], L.global_gbeg[c + 1], S.ci, S.cj, S.ck, none, 0, 0, 0);
}

// EventStateCalculator::_default_event_state_calculation (BaseMonteEventData.cc:122-156)
// from the task values v[kmc_n_tasks]
template <typename V>
__device__ __forceinline__ void kmc_event_combine(const KmcArgs &a, int r, const cmx_prim_event &E, const V &v,
                                                  cmx_event_state &out) {
  double dE = 0.0;
  for (int c = 0; c < a.n_eci; ++c) {
    double acc = 0.0;
    for (int q = 0; q < E.n_sites; ++q) {
      const double d = v[c * E.n_sites + q];
      acc = (q == 0) ? d : __dadd_rn(acc, d);
    }
    dE = __dadd_rn(dE, __dmul_rn(a.eci_val[c], acc));
  }
  out.dE_final = dE;
  const DevEventType &Y = a.types[E.event_type];
  int t = a.n_eci * E.n_sites;
  double ekra = 0.0, freq = 0.0;
  for (int q = Y.kra_beg; q < Y.kra_end; ++q) ekra = __dadd_rn(ekra, __dmul_rn(a.coef_val[q], v[t++]));
  for (int q = Y.freq_beg; q < Y.freq_end; ++q) freq = __dadd_rn(freq, __dmul_rn(a.coef_val[q], v[t++]));
  out.Ekra = ekra;
  out.freq = freq;
  // BaseMonteEventData.cc:148-155
  double dEa = __dadd_rn(__dmul_rn(dE, 0.5), ekra);
  out.is_normal = (dEa > 0.0) && (dEa > dE);
  if (dEa < dE) dEa = dE;
  if (dEa < 0.0) dEa = 0.0;
  out.dE_activated = dEa;
  out.rate = freq * exp(-a.beta[r] * dEa);
}

struct KmcLazyValues {  // single-thread evaluation: a task is computed when it is read
  const KmcArgs &a;
  const int8_t *occ;
  const cmx_prim_event &E;
  const KmcEventSites &S;
  __device__ __forceinline__ double operator[](int task) const { return kmc_task_value(a, occ, E, S, task); }
};

__device__ void kmc_event_state(const KmcArgs &a, int r, int64_t cell, int pe,
                                cmx_event_state &out) {
  const int8_t *occ = a.occ + (size_t)r * a.g.rep_stride;
  const cmx_prim_event &E = a.prim[pe];
  out.is_allowed = 1;
  out.is_normal = 0;
  out.dE_final = out.Ekra = out.dE_activated = out.freq = out.rate = 0.0;
  KmcEventSites S;
  if (!kmc_event_sites(a, occ, cell, E, S)) {
    out.is_allowed = 0;
    return;
  }
  const KmcLazyValues v{a, occ, E, S};
  kmc_event_combine(a, r, E, v, out);
}

__global__ void k_kmc_event_states(KmcArgs a, int64_t n, const int32_t *__restrict__ replica,
                                   const int64_t *__restrict__ cell,
                                   const int32_t *__restrict__ prim,
                                   cmx_event_state *__restrict__ out) {
  int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (x >= n) return;
  cmx_event_state st;
  kmc_event_state(a, replica[x], cell[x], prim[x], st);
  out[x] = st;
}

// every event of every replica; thread x = (replica, cell, prim) with prim fastest
__global__ void k_kmc_all_rates(KmcArgs a, int n_prim, int64_t total,
                                double *__restrict__ rates) {
  for (int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; x < total;
       x += (int64_t)gridDim.x * blockDim.x) {
    int pe = (int)(x % n_prim);
    int64_t rc = x / n_prim;
    int64_t cell = rc % a.g.n_cells;
    int r = (int)(rc / a.g.n_cells);
    cmx_event_state st;
    kmc_event_state(a, r, cell, pe, st);
    rates[x] = st.rate;
  }
}

// per-replica total rate: fixed-order two-stage sum (deterministic)
__global__ void k_kmc_sum(const double *__restrict__ rates, int64_t per_replica,
                          double *__restrict__ partial) {
  __shared__ double sh[256];
  int r = blockIdx.y;
  const double *p = rates + (size_t)r * per_replica;
  double acc = 0.0;
  for (int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; x < per_replica;
       x += (int64_t)gridDim.x * blockDim.x)
    acc += p[x];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(size_t)r * gridDim.x + blockIdx.x] = sh[0];
}
__global__ void k_kmc_sum2(const double *__restrict__ partial, int nb, double *__restrict__ total) {
  int r = blockIdx.x;
  if (threadIdx.x) return;
  double a = 0.0;
  for (int q = 0; q < nb; ++q) a += partial[(size_t)r * nb + q];
  total[r] = a;
}

template <typename T>
static int kmc_to_device(const std::vector<T> &v, T **d) {
  size_t bytes = (v.size() ? v.size() : 1) * sizeof(T);
  CMX_CUDA(cudaMalloc((void **)d, bytes));
  if (!v.empty())
    CMX_CUDA(cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return CMX_OK;
}

extern "C" void cmx_kmc_destroy(cmx_kmc *k) {
  if (!k) return;
  cudaSetDevice(k->s->t->device);
  cudaFree(k->d_prim);
  cudaFree(k->d_types);
  cudaFree(k->d_tables);
  cudaFree(k->d_coef_idx);
  cudaFree(k->d_coef_val);
  cudaFree(k->d_rates);
  cudaFree(k->d_total);
  cudaFree(k->d_imp_beg);
  cudaFree(k->d_imp);
  cudaFree(k->d_tree);
  cudaFree(k->d_mt);
  cudaFree(k->d_time);
  cudaFree(k->d_pending);
  cudaFree(k->d_steps);
  delete k;
}

extern "C" int cmx_kmc_create(cmx_state *s, int32_t n_types, const cmx_event_type *types,
                              int32_t n_prim, const cmx_prim_event *prim, cmx_kmc **out) {
  if (!s || !types || !prim || !out || n_types <= 0 || n_prim <= 0)
    return invalid("cmx_kmc_create: bad argument");
  if (!s->d_eci_idx) {
    cmx_set_error("cmx_kmc_create: bind the formation_energy ECI first (cmx_state_set_eci)");
    return CMX_ERR_STATE;
  }
  const DevTables &T = s->t->d;
  if (T.n_point_corr != T.n_nlist_sublat)
    return invalid("cmx_kmc_create: the state's basis set must be a global (periodic) clexulator");
  std::vector<DevTables> tabs;
  std::vector<DevEventType> dtypes;
  std::vector<uint32_t> cidx;
  std::vector<double> cval;
  for (int y = 0; y < n_types; ++y) {
    const cmx_event_type &Y = types[y];
    if (Y.n_equivalents <= 0 || !Y.local_tables || Y.n_kra < 0 || Y.n_freq < 0 ||
        (Y.n_kra && (!Y.kra_index || !Y.kra_value)) || (Y.n_freq && (!Y.freq_index || !Y.freq_value)))
      return invalid("cmx_kmc_create: bad event type");
    DevEventType d;
    d.n_equivalents = Y.n_equivalents;
    d.table0 = (int32_t)tabs.size();
    for (int e = 0; e < Y.n_equivalents; ++e) {
      const cmx_tables *lt = Y.local_tables[e];
      if (!lt) return invalid("cmx_kmc_create: null local clexulator");
      if (lt->device != s->t->device) return invalid("cmx_kmc_create: local clexulator on another device");
      if (lt->d.n_sublat != T.n_sublat || lt->d.max_occ > 8)
        return invalid("cmx_kmc_create: local clexulator does not match the prim");
      for (int n = 0; n < lt->d.nlist_len; ++n) {
        const int32_t *o = &lt->nbr[4 * n];
        if (std::abs(o[0]) > s->g.N0 || std::abs(o[1]) > s->g.N1 || std::abs(o[2]) > s->g.N2)
          return invalid("cmx_kmc_create: supercell smaller than the local neighborhood");
      }
      for (int q = 0; q < Y.n_kra; ++q)
        if (Y.kra_index[q] >= (uint32_t)lt->d.corr_size) return invalid("cmx_kmc_create: kra index out of range");
      for (int q = 0; q < Y.n_freq; ++q)
        if (Y.freq_index[q] >= (uint32_t)lt->d.corr_size) return invalid("cmx_kmc_create: freq index out of range");
      tabs.push_back(lt->d);
    }
    d.kra_beg = (int32_t)cidx.size();
    cidx.insert(cidx.end(), Y.kra_index, Y.kra_index + Y.n_kra);
    cval.insert(cval.end(), Y.kra_value, Y.kra_value + Y.n_kra);
    d.kra_end = d.freq_beg = (int32_t)cidx.size();
    cidx.insert(cidx.end(), Y.freq_index, Y.freq_index + Y.n_freq);
    cval.insert(cval.end(), Y.freq_value, Y.freq_value + Y.n_freq);
    d.freq_end = (int32_t)cidx.size();
    dtypes.push_back(d);
  }
  if (s->g.halo) return invalid("cmx_kmc_create: slab states are not supported");
  for (int p = 0; p < n_prim; ++p) {
    const cmx_prim_event &E = prim[p];
    if (E.n_sites < 1 || E.n_sites > CMX_EVENT_MAX_SITES) return invalid("cmx_kmc_create: bad number of event sites");
    if (E.event_type < 0 || E.event_type >= n_types) return invalid("cmx_kmc_create: event type out of range");
    if (E.equivalent_index < 0 || E.equivalent_index >= types[E.event_type].n_equivalents)
      return invalid("cmx_kmc_create: equivalent index out of range");
    for (int q = 0; q < E.n_sites; ++q) {
      int b = E.site[q][0];
      if (b < 0 || b >= T.n_sublat) return invalid("cmx_kmc_create: event sublattice out of range");
      if (std::abs(E.site[q][1]) > s->g.N0 || std::abs(E.site[q][2]) > s->g.N1 || std::abs(E.site[q][3]) > s->g.N2)
        return invalid("cmx_kmc_create: supercell smaller than the event");
      if (E.occ_init[q] < 0 || E.occ_init[q] >= s->t->n_occ[b] || E.occ_final[q] < 0 ||
          E.occ_final[q] >= s->t->n_occ[b])
        return invalid("cmx_kmc_create: event occupant index out of range");
    }
  }
  CMX_CUDA(cudaSetDevice(s->t->device));
  cmx_kmc *k = new cmx_kmc;
  k->s = s;
  k->n_types = n_types;
  k->n_prim = n_prim;
  k->prim.assign(prim, prim + n_prim);
  k->types = dtypes;
  int rc;
  if ((rc = kmc_to_device(k->prim, &k->d_prim)) || (rc = kmc_to_device(dtypes, &k->d_types)) ||
      (rc = kmc_to_device(tabs, &k->d_tables)) || (rc = kmc_to_device(cidx, &k->d_coef_idx)) ||
      (rc = kmc_to_device(cval, &k->d_coef_val))) {
    cmx_kmc_destroy(k);
    return rc;
  }
  *out = k;
  return CMX_OK;
}

static int kmc_args(cmx_kmc *k, KmcArgs &a, const char *who) {
  cmx_state *s = k->s;
  for (int r = 0; r < s->n_replicas; ++r)
    if (!(s->temperature[r] > 0.0)) {
      cmx_set_error(std::string(who) + ": temperature not set for every replica");
      return CMX_ERR_STATE;
    }
  a.T = s->t->d;
  a.g = s->g;
  a.occ = s->d_occ;
  a.n_eci = s->n_eci;
  a.eci_idx = s->d_eci_idx;
  a.eci_val = s->d_eci_val;
  a.beta = s->d_beta;
  a.prim = k->d_prim;
  a.types = k->d_types;
  a.local = k->d_tables;
  a.coef_idx = k->d_coef_idx;
  a.coef_val = k->d_coef_val;
  return CMX_OK;
}

extern "C" int cmx_kmc_event_states(cmx_kmc *k, int64_t n, const int32_t *replica,
                                    const int64_t *unitcell, const int32_t *prim_event,
                                    cmx_event_state *out) {
  if (!k) return invalid("cmx_kmc_event_states: null handle");
  if (n < 0 || (n && (!replica || !unitcell || !prim_event || !out)))
    return invalid("cmx_kmc_event_states: bad argument");
  if (n == 0) return CMX_OK;
  cmx_state *s = k->s;
  for (int64_t q = 0; q < n; ++q) {
    if (replica[q] < 0 || replica[q] >= s->n_replicas)
      return invalid("cmx_kmc_event_states: replica out of range");
    if (unitcell[q] < 0 || unitcell[q] >= s->g.n_cells)
      return invalid("cmx_kmc_event_states: unit cell index out of range");
    if (prim_event[q] < 0 || prim_event[q] >= k->n_prim)
      return invalid("cmx_kmc_event_states: prim event index out of range");
  }
  KmcArgs a;
  int rc = kmc_args(k, a, "cmx_kmc_event_states");
  if (rc) return rc;
  CMX_CUDA(cudaSetDevice(s->t->device));
  size_t b_r = sizeof(int32_t) * n, b_c = sizeof(int64_t) * n, b_o = sizeof(cmx_event_state) * n;
  size_t off_c = (b_r + 255) & ~(size_t)255;
  size_t off_p = (off_c + b_c + 255) & ~(size_t)255;
  size_t off_o = (off_p + b_r + 255) & ~(size_t)255;
  rc = cmx_scratch(s, off_o + b_o);
  if (rc) return rc;
  char *base = static_cast<char *>(s->d_scratch);
  CMX_CUDA(cudaMemcpyAsync(base, replica, b_r, cudaMemcpyHostToDevice, s->stream));
  CMX_CUDA(cudaMemcpyAsync(base + off_c, unitcell, b_c, cudaMemcpyHostToDevice, s->stream));
  CMX_CUDA(cudaMemcpyAsync(base + off_p, prim_event, b_r, cudaMemcpyHostToDevice, s->stream));
  k_kmc_event_states<<<(unsigned)((n + 127) / 128), 128, 0, s->stream>>>(
      a, n, (const int32_t *)base, (const int64_t *)(base + off_c), (const int32_t *)(base + off_p),
      (cmx_event_state *)(base + off_o));
  CMX_CUDA(cudaGetLastError());
  CMX_CUDA(cudaMemcpyAsync(out, base + off_o, b_o, cudaMemcpyDeviceToHost, s->stream));
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  return CMX_OK;
}

extern "C" int cmx_kmc_all_rates(cmx_kmc *k, double *rates, double *total, void **d_rates) {
  if (!k) return invalid("cmx_kmc_all_rates: null handle");
  cmx_state *s = k->s;
  KmcArgs a;
  int rc = kmc_args(k, a, "cmx_kmc_all_rates");
  if (rc) return rc;
  CMX_CUDA(cudaSetDevice(s->t->device));
  int64_t per = s->g.n_cells * k->n_prim, tot = per * s->n_replicas;
  if (!k->d_rates) {
    CMX_CUDA(cudaMalloc((void **)&k->d_rates, sizeof(double) * tot));
    CMX_CUDA(cudaMalloc((void **)&k->d_total, sizeof(double) * s->n_replicas));
  }
  int64_t nb = std::min<int64_t>((tot + 127) / 128, 148 * 16);
  k_kmc_all_rates<<<(unsigned)nb, 128, 0, s->stream>>>(a, k->n_prim, tot, k->d_rates);
  CMX_CUDA(cudaGetLastError());
  if (total) {
    int nbs = (int)std::min<int64_t>((per + 255) / 256, 64);
    rc = cmx_scratch(s, sizeof(double) * (size_t)nbs * s->n_replicas);
    if (rc) return rc;
    dim3 grid(nbs, s->n_replicas);
    k_kmc_sum<<<grid, 256, 0, s->stream>>>(k->d_rates, per, (double *)s->d_scratch);
    k_kmc_sum2<<<s->n_replicas, 32, 0, s->stream>>>((const double *)s->d_scratch, nbs, k->d_total);
    CMX_CUDA(cudaGetLastError());
    CMX_CUDA(cudaMemcpyAsync(total, k->d_total, sizeof(double) * s->n_replicas,
                             cudaMemcpyDeviceToHost, s->stream));
  }
  if (rates)
    CMX_CUDA(cudaMemcpyAsync(rates, k->d_rates, sizeof(double) * tot, cudaMemcpyDeviceToHost, s->stream));
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  if (d_rates) *d_rates = k->d_rates;
  return CMX_OK;
}


// ---------------------------------------------------------------------------
// Rejection-free KMC, whole steps on the device.
//
// Reference: kinetic_monte_carlo_v2 (include/casm/clexmonte/methods/kinetic_monte_carlo.hh:
// 396-507) driving lotto::RejectionFreeEventSelector over the complete event list
// (CompleteKineticEventData, monte_calculator/kinetic_events.hh:73-133; event ids in the
// order of make_complete_event_id_list, events/CompleteEventList.cc:76-91: unit cell major):
//   select_event (submodules/kmc-lotto/include/lotto/rejection_free.hpp:93-111):
//     update the rates of the events impacted by the previous event (:151-159),
//     total = root of the sum tree, dt = -log(u1) / total (event_selector.hpp:65-70),
//     query = total * u2, descend the tree (event_rate_tree_impl.hpp:63-75,135-150:
//     left if query <= left.rate, else subtract and go right),
//   time += dt, apply the event (OccLocation::apply), remember its impact list.
// The tree is lotto's InvertedBinarySumTree (sum_tree_impl.hpp:190-240): leaves joined
// pairwise level by level, an odd node out is carried up unchanged; here as arrays per
// level, parent = left + right in that order, so every partial sum has the reference's
// bits.  u1, u2 come from std::mt19937_64 through lotto's (0, 1] distribution
// (cmx_mt64.cuh).  One thread block per trajectory: the impact list (708 events for the
// FCC A-B-Va system) is re-evaluated by the block's threads with the faithful
// event-state function, then the touched tree paths are re-summed level by level.
// ---------------------------------------------------------------------------
#define CMX_KMC_MAX_LEVELS 40
struct KmcTree {
  int n_levels;  // levels above the leaves
  long long size[CMX_KMC_MAX_LEVELS + 1];  // size[0] = leaves
  long long off[CMX_KMC_MAX_LEVELS + 1];   // offset of level l >= 1 inside a replica's d_tree
  long long upper;                         // nodes above the leaves per replica
};

__device__ __forceinline__ double *kmc_level(double *leaves, double *tree, const KmcTree &t, int l) {
  return l == 0 ? leaves : tree + t.off[l];
}

// level l (>= 1) of every replica from level l - 1
__global__ void k_kmc_tree_level(double *rates, double *tree, KmcTree t, int l, long long per) {
  const int r = blockIdx.y;
  double *leaves = rates + (size_t)r * per, *up = tree + (size_t)r * t.upper;
  const double *child = kmc_level(leaves, up, t, l - 1);
  double *node = kmc_level(leaves, up, t, l);
  const long long nc = t.size[l - 1];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < t.size[l];
       i += (long long)gridDim.x * blockDim.x)
    node[i] = (2 * i + 1 < nc) ? __dadd_rn(child[2 * i], child[2 * i + 1]) : child[2 * i];
}

__global__ void k_kmc_seed(Mt64 *mt, const unsigned long long *seeds, double *time, long long *pending,
                           long long *steps, int n) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  mt_seed(mt[r], seeds[r]);
  time[r] = 0.0;
  pending[r] = -1;
  steps[r] = 0;
}

struct KmcRunArgs {
  KmcArgs a;
  KmcTree t;
  int n_prim;
  long long per;  // leaves per replica
  double *rates, *tree;
  const int32_t *imp_beg;
  const int4 *imp;
  Mt64 *mt;
  double *time;
  long long *pending, *steps;
  cmx_kmc_step *log;
  long long log_cap;
  long long n_steps;
  int max_imp, t_max;
  int l_sh;  // first tree level kept in shared memory
};

#define CMX_KMC_CHUNK 16  // allowed events evaluated per round of the block

// shared scratch of k_kmc_run: ids[max_imp] | allowed[max_imp] | values[CHUNK][t_max]
struct KmcShared {
  long long *ids;
  int *allowed;
  double *val;
  int t_max;
  // the upper levels of the sum tree (from level l_sh on: <= 1024 nodes) live in shared
  // memory while the kernel runs: re-summing and descending them costs shared-memory
  // latency instead of an L2 round trip per level
  double *tree;
  int l_sh;
};
__device__ __forceinline__ double *kmc_level_sh(double *leaves, double *up, const KmcTree &t, const KmcShared &sh,
                                                int l) {
  if (l >= sh.l_sh) return sh.tree + (t.off[l] - t.off[sh.l_sh]);
  return l == 0 ? leaves : up + t.off[l];
}

// re-evaluate the events impacted by event `ev` and re-sum their tree paths
__device__ void kmc_apply_impact(const KmcRunArgs &A, int r, long long ev, double *leaves, double *up,
                                 const KmcShared &sh) {
  __shared__ int sh_n_allowed;
  const Geom &g = A.a.g;
  const int8_t *occ = A.a.occ + (size_t)r * g.rep_stride;
  const int pe = (int)(ev % A.n_prim);
  const long long cell = ev / A.n_prim;
  const int ci = (int)(cell % g.N0);
  const long long rest = cell / g.N0;
  const int cj = (int)(rest % g.N1), ck = (int)(rest / g.N1);
  const int ib = A.imp_beg[pe], n_imp = A.imp_beg[pe + 1] - ib;
  if (threadIdx.x == 0) sh_n_allowed = 0;
  __syncthreads();
  // 1. events that are not allowed get rate 0 at once; the allowed ones are listed
  for (int q = threadIdx.x; q < n_imp; q += blockDim.x) {
    const int4 e = A.imp[ib + q];
    int i2 = ci + e.y, j2 = cj + e.z, k2 = ck + e.w;
    cmx_wrap_cell(g, i2, j2, k2);
    const long long c2 = ((long long)k2 * g.N1 + j2) * g.N0 + i2;
    const long long id = c2 * A.n_prim + e.x;
    sh.ids[q] = id;
    KmcEventSites S;
    if (kmc_event_sites(A.a, occ, c2, A.a.prim[e.x], S)) sh.allowed[atomicAdd(&sh_n_allowed, 1)] = q;
    else leaves[id] = 0.0;
  }
  __syncthreads();
  // 2. allowed events, CMX_KMC_CHUNK at a time: their function evaluations (tasks) are
  // spread over the block, then one thread per event combines them in reference order.
  // (the order of `allowed` depends on the atomics; every event is evaluated
  // independently, so the results do not)
  const int n_allowed = sh_n_allowed;
  for (int base = 0; base < n_allowed; base += CMX_KMC_CHUNK) {
    const int n_ev = min(CMX_KMC_CHUNK, n_allowed - base);
    for (int w = threadIdx.x; w < n_ev * sh.t_max; w += blockDim.x) {
      // neighbouring threads evaluate the SAME function (task) for different events: the
      // loops over groups, terms and factors stay convergent within a warp
      const int task = w / n_ev, x = w - task * n_ev;
      const long long id = sh.ids[sh.allowed[base + x]];
      const cmx_prim_event &E = A.a.prim[(int)(id % A.n_prim)];
      if (task >= kmc_n_tasks(A.a, E)) continue;
      KmcEventSites S;
      kmc_event_sites(A.a, occ, id / A.n_prim, E, S);
      sh.val[x * sh.t_max + task] = kmc_task_value(A.a, occ, E, S, task);
    }
    __syncthreads();
    for (int x = threadIdx.x; x < n_ev; x += blockDim.x) {
      const long long id = sh.ids[sh.allowed[base + x]];
      const cmx_prim_event &E = A.a.prim[(int)(id % A.n_prim)];
      cmx_event_state st;
      kmc_event_combine(A.a, r, E, sh.val + (size_t)x * sh.t_max, st);
      leaves[id] = st.rate;
    }
    __syncthreads();
  }
  // 3. re-sum the touched paths, level by level
  for (int l = 1; l <= A.t.n_levels; ++l) {
    const double *child = kmc_level_sh(leaves, up, A.t, sh, l - 1);
    double *node = kmc_level_sh(leaves, up, A.t, sh, l);
    const long long nc = A.t.size[l - 1];
    for (int q = threadIdx.x; q < n_imp; q += blockDim.x) {
      const long long i = sh.ids[q] >> l;
      // ids mostly ascend (cell-major impact table): the previous entry usually has the
      // same parent; a repeated parent that slips through writes the same value again
      if (q > 0 && (sh.ids[q - 1] >> l) == i) continue;
      node[i] = (2 * i + 1 < nc) ? __dadd_rn(child[2 * i], child[2 * i + 1]) : child[2 * i];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_kmc_run(KmcRunArgs A) {
  extern __shared__ long long sh_dyn[];
  __shared__ long long sh_ev;
  KmcShared sh_ids;
  sh_ids.ids = sh_dyn;
  sh_ids.val = reinterpret_cast<double *>(sh_dyn + A.max_imp);
  sh_ids.allowed = reinterpret_cast<int *>(sh_ids.val + CMX_KMC_CHUNK * A.t_max);
  sh_ids.t_max = A.t_max;
  sh_ids.tree = sh_ids.val + CMX_KMC_CHUNK * A.t_max + (A.max_imp + 1) / 2;  // after `allowed` (ints)
  sh_ids.l_sh = A.l_sh;
  const int r = blockIdx.x;
  const Geom &g = A.a.g;
  double *leaves = A.rates + (size_t)r * A.per, *up = A.tree + (size_t)r * A.t.upper;
  int8_t *occ = const_cast<int8_t *>(A.a.occ) + (size_t)r * g.rep_stride;
  long long pending = A.pending[r];
  double time = A.time[r];
  long long steps = A.steps[r];
  const long long n_sh = A.t.upper - A.t.off[A.l_sh];  // nodes of the shared levels (contiguous at the end)
  for (long long q = threadIdx.x; q < n_sh; q += blockDim.x) sh_ids.tree[q] = up[A.t.off[A.l_sh] + q];
  __syncthreads();
  for (long long step = 0; step < A.n_steps; ++step) {
    if (pending >= 0) kmc_apply_impact(A, r, pending, leaves, up, sh_ids);
    if (threadIdx.x == 0) {
      const double total = *kmc_level_sh(leaves, up, A.t, sh_ids, A.t.n_levels);
      long long ev = -1;
      if (total > 0.0) {
        Mt64 &mt = A.mt[r];
        const double dt = __ddiv_rn(-log(mt_unit_interval(mt)), total);
        double q = __dmul_rn(total, mt_unit_interval(mt));
        long long idx = 0;
        for (int l = A.t.n_levels; l > 0; --l) {
          const double *child = kmc_level_sh(leaves, up, A.t, sh_ids, l - 1);
          const double left = child[2 * idx];
          if (q <= left || 2 * idx + 1 >= A.t.size[l - 1]) {
            idx = 2 * idx;
          } else {
            q = __dsub_rn(q, left);
            idx = 2 * idx + 1;
          }
        }
        // Rounding in the residual query can end the descent on a leaf of rate 0 (a
        // disallowed event; lotto's bifurcate returns the internal node there).  Never
        // apply such an event: take the nearest leaf that can happen and mark the step
        // (cmx_kmc_step::pad = 1).
        int32_t fallback = 0;
        if (!(leaves[idx] > 0.0)) {
          fallback = 1;
          long long lo = idx - 1, hi = idx + 1;
          const long long n_leaf = A.t.size[0];
          for (;;) {
            if (lo >= 0 && leaves[lo] > 0.0) { idx = lo; break; }
            if (hi < n_leaf && leaves[hi] > 0.0) { idx = hi; break; }
            --lo;
            ++hi;
            if (lo < 0 && hi >= n_leaf) break;  // unreachable: total > 0
          }
        }
        ev = idx;
        time = __dadd_rn(time, dt);
        // OccLocation::apply: every site of the event takes its final occupant
        const cmx_prim_event &E = A.a.prim[(int)(ev % A.n_prim)];
        const long long cell = ev / A.n_prim;
        const int ci = (int)(cell % g.N0);
        const long long rest = cell / g.N0;
        const int cj = (int)(rest % g.N1), ck = (int)(rest / g.N1);
        for (int s = 0; s < E.n_sites; ++s) {
          int i2 = ci + E.site[s][1], j2 = cj + E.site[s][2], k2 = ck + E.site[s][3];
          cmx_wrap_cell(g, i2, j2, k2);
          occ[cmx_site_offset(g, E.site[s][0], i2, j2, k2)] = (int8_t)cmx_enc(g, E.occ_final[s]);
        }
        if (A.log && steps < A.log_cap) {
          cmx_kmc_step &L = A.log[(size_t)r * A.log_cap + steps];
          L.unitcell = cell;
          L.prim_event = (int32_t)(ev % A.n_prim);
          L.pad = fallback;
          L.time_increment = dt;
          L.total_rate = total;
        }
        ++steps;
      }
      sh_ev = ev;
    }
    __syncthreads();
    pending = sh_ev;
    if (pending < 0) break;  // no event can happen (total rate 0)
  }
  // leave rates and tree consistent with the occupation (the reference does this at the
  // start of the next select_event; the update is idempotent)
  if (pending >= 0) kmc_apply_impact(A, r, pending, leaves, up, sh_ids);
  __syncthreads();
  for (long long q = threadIdx.x; q < n_sh; q += blockDim.x) up[A.t.off[A.l_sh] + q] = sh_ids.tree[q];
  if (threadIdx.x == 0) {
    A.pending[r] = -1;
    A.time[r] = time;
    A.steps[r] = steps;
  }
}

extern "C" int cmx_kmc_set_impact_table(cmx_kmc *k, int32_t n_entries, const int32_t *beg,
                                        const int32_t *entries) {
  if (!k || !beg || (n_entries && !entries) || n_entries < 0) return invalid("cmx_kmc_set_impact_table: bad argument");
  if (beg[0] != 0 || beg[k->n_prim] != n_entries) return invalid("cmx_kmc_set_impact_table: inconsistent offsets");
  const Geom &g = k->s->g;
  int max_imp = 0;
  for (int p = 0; p < k->n_prim; ++p) {
    if (beg[p + 1] < beg[p]) return invalid("cmx_kmc_set_impact_table: offsets not ascending");
    max_imp = std::max(max_imp, beg[p + 1] - beg[p]);
  }
  std::vector<int4> imp(n_entries);
  for (int q = 0; q < n_entries; ++q) {
    const int32_t *e = entries + 4 * q;
    if (e[0] < 0 || e[0] >= k->n_prim) return invalid("cmx_kmc_set_impact_table: prim event out of range");
    if (std::abs(e[1]) > g.N0 || std::abs(e[2]) > g.N1 || std::abs(e[3]) > g.N2)
      return invalid("cmx_kmc_set_impact_table: supercell smaller than the impact neighborhood");
    imp[q] = make_int4(e[0], e[1], e[2], e[3]);
  }
  CMX_CUDA(cudaSetDevice(k->s->t->device));
  cudaFree(k->d_imp_beg);
  cudaFree(k->d_imp);
  k->d_imp_beg = nullptr;
  k->d_imp = nullptr;
  std::vector<int32_t> b(beg, beg + k->n_prim + 1);
  int rc;
  if ((rc = kmc_to_device(b, &k->d_imp_beg)) || (rc = kmc_to_device(imp, &k->d_imp))) return rc;
  k->max_imp = max_imp;
  k->run_ready = false;
  return CMX_OK;
}

static KmcTree kmc_tree(const cmx_kmc *k) {
  KmcTree t;
  t.n_levels = (int)k->level_size.size() - 1;
  for (size_t l = 0; l < k->level_size.size(); ++l) {
    t.size[l] = k->level_size[l];
    t.off[l] = k->level_off[l];
  }
  t.upper = k->level_off.back() + k->level_size.back();
  return t;
}

// all rates from the current occupation, the sum tree, seeds, time = 0
extern "C" int cmx_kmc_run_begin(cmx_kmc *k, const uint64_t *seeds) {
  if (!k || !seeds) return invalid("cmx_kmc_run_begin: null argument");
  if (!k->d_imp_beg) {
    cmx_set_error("cmx_kmc_run_begin: no impact table (cmx_kmc_set_impact_table)");
    return CMX_ERR_STATE;
  }
  cmx_state *s = k->s;
  const int R = s->n_replicas;
  int rc = cmx_kmc_all_rates(k, nullptr, nullptr, nullptr);
  if (rc) return rc;
  const long long per = (long long)s->g.n_cells * k->n_prim;
  k->level_size.assign(1, per);
  k->level_off.assign(1, 0);
  long long off = 0;
  while (k->level_size.back() > 1) {
    const long long n = (k->level_size.back() + 1) / 2;
    k->level_off.push_back(off);
    k->level_size.push_back(n);
    off += n;
  }
  if ((int)k->level_size.size() - 1 > CMX_KMC_MAX_LEVELS) return invalid("cmx_kmc_run_begin: event list too large");
  if (k->level_size.size() == 1) {  // a single event: the root is the leaf
    k->level_off.push_back(0);
    k->level_size.push_back(1);
    off = 1;
  }
  const KmcTree t = kmc_tree(k);
  cudaFree(k->d_tree);
  cudaFree(k->d_mt);
  cudaFree(k->d_time);
  cudaFree(k->d_pending);
  cudaFree(k->d_steps);
  k->d_tree = nullptr;
  k->d_mt = nullptr;
  k->d_time = nullptr;
  k->d_pending = nullptr;
  k->d_steps = nullptr;
  CMX_CUDA(cudaMalloc((void **)&k->d_tree, sizeof(double) * (size_t)t.upper * R));
  CMX_CUDA(cudaMalloc((void **)&k->d_mt, sizeof(Mt64) * R));
  CMX_CUDA(cudaMalloc((void **)&k->d_time, sizeof(double) * R));
  CMX_CUDA(cudaMalloc((void **)&k->d_pending, sizeof(long long) * R));
  CMX_CUDA(cudaMalloc((void **)&k->d_steps, sizeof(long long) * R));
  for (int l = 1; l <= t.n_levels; ++l) {
    dim3 grid((unsigned)std::min<long long>((t.size[l] + 255) / 256, 1024), R);
    k_kmc_tree_level<<<grid, 256, 0, s->stream>>>(k->d_rates, k->d_tree, t, l, per);
  }
  CMX_CUDA(cudaGetLastError());
  rc = cmx_scratch(s, sizeof(unsigned long long) * R);
  if (rc) return rc;
  CMX_CUDA(cudaMemcpyAsync(s->d_scratch, seeds, sizeof(unsigned long long) * R, cudaMemcpyHostToDevice, s->stream));
  k_kmc_seed<<<(R + 63) / 64, 64, 0, s->stream>>>(k->d_mt, (const unsigned long long *)s->d_scratch, k->d_time,
                                                 k->d_pending, k->d_steps, R);
  CMX_CUDA(cudaGetLastError());
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  k->run_ready = true;
  return CMX_OK;
}

// n_steps events of every trajectory.  log[n_replicas][log_cap] (may be NULL) receives the
// first log_cap steps since cmx_kmc_run_begin; time / total_rate / n_steps_done [n_replicas]
// (each may be NULL): simulated time, current total rate, steps done so far.
extern "C" int cmx_kmc_run(cmx_kmc *k, int64_t n_steps, cmx_kmc_step *log, int64_t log_cap, double *time,
                           double *total_rate, int64_t *n_steps_done) {
  if (!k) return invalid("cmx_kmc_run: null handle");
  if (!k->run_ready) {
    cmx_set_error("cmx_kmc_run: call cmx_kmc_run_begin first");
    return CMX_ERR_STATE;
  }
  if (n_steps < 0 || log_cap < 0 || (log_cap && !log)) return invalid("cmx_kmc_run: bad argument");
  cmx_state *s = k->s;
  const int R = s->n_replicas;
  KmcRunArgs A;
  int rc = kmc_args(k, A.a, "cmx_kmc_run");
  if (rc) return rc;
  CMX_CUDA(cudaSetDevice(s->t->device));
  A.t = kmc_tree(k);
  A.n_prim = k->n_prim;
  A.per = (long long)s->g.n_cells * k->n_prim;
  A.rates = k->d_rates;
  A.tree = k->d_tree;
  A.imp_beg = k->d_imp_beg;
  A.imp = k->d_imp;
  A.mt = k->d_mt;
  A.time = k->d_time;
  A.pending = k->d_pending;
  A.steps = k->d_steps;
  A.log = nullptr;
  A.log_cap = log_cap;
  A.n_steps = n_steps;
  cmx_kmc_step *d_log = nullptr;
  if (log_cap) {
    CMX_CUDA(cudaMalloc((void **)&d_log, sizeof(cmx_kmc_step) * (size_t)R * log_cap));
    CMX_CUDA(cudaMemcpyAsync(d_log, log, sizeof(cmx_kmc_step) * (size_t)R * log_cap, cudaMemcpyHostToDevice, s->stream));
    A.log = d_log;
  }
  int t_max = 1;
  for (const cmx_prim_event &E : k->prim) {
    const DevEventType &Y = k->types[E.event_type];
    t_max = std::max(t_max, s->n_eci * E.n_sites + (Y.kra_end - Y.kra_beg) + (Y.freq_end - Y.freq_beg));
  }
  A.max_imp = std::max(1, k->max_imp);
  A.t_max = t_max;
  A.l_sh = A.t.n_levels;
  while (A.l_sh > 1 && A.t.upper - A.t.off[A.l_sh - 1] <= 2048) --A.l_sh;  // <= 16 KB of upper levels
  const size_t n_sh = (size_t)(A.t.upper - A.t.off[A.l_sh]);
  const size_t smem = sizeof(long long) * A.max_imp + sizeof(double) * CMX_KMC_CHUNK * t_max +
                      sizeof(double) * ((A.max_imp + 1) / 2) + sizeof(double) * n_sh;
  if (smem > 48 * 1024) {
    cudaFree(d_log);
    return invalid("cmx_kmc_run: impact lists longer than 6144 events are not supported");
  }
  k_kmc_run<<<R, 256, smem, s->stream>>>(A);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    cudaFree(d_log);
    cmx_set_error(std::string("cmx_kmc_run: ") + cudaGetErrorString(e));
    return CMX_ERR_CUDA;
  }
  if (log_cap)
    CMX_CUDA(cudaMemcpyAsync(log, d_log, sizeof(cmx_kmc_step) * (size_t)R * log_cap, cudaMemcpyDeviceToHost, s->stream));
  if (time) CMX_CUDA(cudaMemcpyAsync(time, k->d_time, sizeof(double) * R, cudaMemcpyDeviceToHost, s->stream));
  if (n_steps_done)
    CMX_CUDA(cudaMemcpyAsync(n_steps_done, k->d_steps, sizeof(long long) * R, cudaMemcpyDeviceToHost, s->stream));
  std::vector<double> roots;
  if (total_rate) {
    // the root of replica r sits at the end of its block of upper levels
    const KmcTree &t = A.t;
    CMX_CUDA(cudaMemcpy2DAsync(total_rate, sizeof(double), k->d_tree + t.off[t.n_levels], sizeof(double) * t.upper,
                               sizeof(double), R, cudaMemcpyDeviceToHost, s->stream));
  }
  e = cudaStreamSynchronize(s->stream);
  cudaFree(d_log);
  if (e != cudaSuccess) {
    cmx_set_error(std::string("cmx_kmc_run: ") + cudaGetErrorString(e));
    return CMX_ERR_CUDA;
  }
  return CMX_OK;
}

// rates and total rates as the selector currently holds them (no re-evaluation):
// rates[n_replicas][n_cells][n_prim], total[n_replicas]; either may be NULL
extern "C" int cmx_kmc_current_rates(cmx_kmc *k, double *rates, double *total) {
  if (!k) return invalid("cmx_kmc_current_rates: null handle");
  if (!k->run_ready) {
    cmx_set_error("cmx_kmc_current_rates: call cmx_kmc_run_begin first");
    return CMX_ERR_STATE;
  }
  cmx_state *s = k->s;
  CMX_CUDA(cudaSetDevice(s->t->device));
  const KmcTree t = kmc_tree(k);
  const size_t per = (size_t)s->g.n_cells * k->n_prim;
  if (rates)
    CMX_CUDA(cudaMemcpyAsync(rates, k->d_rates, sizeof(double) * per * s->n_replicas, cudaMemcpyDeviceToHost, s->stream));
  if (total)
    CMX_CUDA(cudaMemcpy2DAsync(total, sizeof(double), k->d_tree + t.off[t.n_levels], sizeof(double) * t.upper,
                               sizeof(double), s->n_replicas, cudaMemcpyDeviceToHost, s->stream));
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  return CMX_OK;
}
