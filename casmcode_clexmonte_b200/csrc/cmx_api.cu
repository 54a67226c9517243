// C ABI: handles, tables, state, host<->device movement.
#include <cstring>

#include "cmx_internal.cuh"

static thread_local std::string g_last_error;

void cmx_set_error(const std::string &msg) { g_last_error = msg; }

int cmx_cuda_fail(cudaError_t e, const char *what) {
  g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  return CMX_ERR_CUDA;
}

static int invalid(const std::string &msg) {
  cmx_set_error(msg);
  return CMX_ERR_INVALID;
}

extern "C" const char *cmx_last_error(void) { return g_last_error.c_str(); }
extern "C" int cmx_version(void) { return 100; }
extern "C" int cmx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

template <typename T>
static int upload(cmx_tables *t, const T *src, size_t n, const T **dst,
                  std::vector<T> *host) {
  if (host) host->assign(src, src + n);
  void *p = nullptr;
  size_t bytes = (n ? n : 1) * sizeof(T);
  CMX_CUDA(cudaMalloc(&p, bytes));
  t->allocs.push_back(p);
  if (n) CMX_CUDA(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
  *dst = static_cast<const T *>(p);
  return CMX_OK;
}

extern "C" int cmx_tables_create(const cmx_table_desc *d, int device,
                                 cmx_tables **out) {
  if (!d || !out) return invalid("cmx_tables_create: null argument");
  if (d->n_sublat <= 0 || d->max_occ <= 0 || d->n_func <= 0 ||
      d->corr_size <= 0 || d->nlist_len <= 0 || d->n_point_corr <= 0)
    return invalid("cmx_tables_create: non-positive size");
  if (d->max_occ > 8) return invalid("cmx_tables_create: max_occ > 8 unsupported");
  // structural validation of the CSR
  for (int i = 0; i < d->n_terms; ++i)
    if (d->term_fbeg[i] > d->term_fbeg[i + 1])
      return invalid("cmx_tables_create: term_fbeg not monotone");
  if (d->term_fbeg[d->n_terms] != d->n_factors ||
      d->elem_tbeg[d->n_elems] != d->n_terms ||
      d->group_ebeg[d->n_groups] != d->n_elems)
    return invalid("cmx_tables_create: CSR sizes inconsistent");
  for (int i = 0; i < d->n_factors; ++i) {
    if (d->factor_n[i] < 0 || d->factor_n[i] >= d->nlist_len)
      return invalid("cmx_tables_create: factor neighbor index out of range");
    if (d->factor_f[i] < 0 || d->factor_f[i] >= d->n_func)
      return invalid("cmx_tables_create: factor function index out of range");
  }
  for (int i = 0; i < d->nlist_len; ++i) {
    int b = d->nbr[4 * i + 3];
    if (b < 0 || b >= d->n_sublat)
      return invalid("cmx_tables_create: neighbor sublattice out of range");
  }
  if (d->global_gbeg[d->corr_size] > d->n_groups ||
      d->delta_gbeg[d->n_point_corr * d->corr_size] > d->n_groups)
    return invalid("cmx_tables_create: function table out of range");
  if (cmx_device_count() == 0) {
    cmx_set_error("cmx_tables_create: no CUDA device (there is no CPU path)");
    return CMX_ERR_CUDA;
  }
  CMX_CUDA(cudaSetDevice(device));
  cmx_tables *t = new cmx_tables;
  t->device = device;
  DevTables &D = t->d;
  D.n_sublat = d->n_sublat;
  D.max_occ = d->max_occ;
  D.n_func = d->n_func;
  D.corr_size = d->corr_size;
  D.n_point_corr = d->n_point_corr;
  D.nlist_len = d->nlist_len;
  D.n_nlist_sublat = d->n_nlist_sublat;
  int rc = CMX_OK;
  const int32_t *nbr_dev = nullptr;
#define UP(field, n, host)                                               \
  if (rc == CMX_OK) rc = upload(t, d->field, (size_t)(n), &D.field, host)
  UP(nlist_sublat, d->n_nlist_sublat, &t->nlist_sublat);
  UP(n_occ, d->n_sublat, &t->n_occ);
  UP(phi, (size_t)d->n_sublat * d->n_func * d->max_occ, &t->phi);
  if (rc == CMX_OK)
    rc = upload(t, d->nbr, (size_t)d->nlist_len * 4, &nbr_dev, &t->nbr);
  D.nbr = reinterpret_cast<const int4 *>(nbr_dev);
  UP(factor_f, d->n_factors, &t->factor_f);
  UP(factor_n, d->n_factors, &t->factor_n);
  UP(term_coef, d->n_terms, &t->term_coef);
  UP(term_fbeg, d->n_terms + 1, &t->term_fbeg);
  UP(elem_tbeg, d->n_elems + 1, &t->elem_tbeg);
  UP(group_ebeg, d->n_groups + 1, &t->group_ebeg);
  UP(group_dphi, d->n_groups, &t->group_dphi);
  UP(group_has_sum, d->n_groups, &t->group_has_sum);
  UP(group_div, d->n_groups, &t->group_div);
  UP(global_gbeg, d->corr_size + 1, &t->global_gbeg);
  UP(point_gbeg, d->n_point_corr * d->corr_size + 1, &t->point_gbeg);
  UP(delta_gbeg, d->n_point_corr * d->corr_size + 1, &t->delta_gbeg);
#undef UP
  if (rc != CMX_OK) {
    cmx_tables_destroy(t);
    return rc;
  }
  *out = t;
  return CMX_OK;
}

// the flat file ClexulatorTables.save_flat writes (clexulator_tables.py): a plugin written in
// C++ loads the export of a basis set without numpy
extern "C" int cmx_tables_create_from_file(const char *path, int device, cmx_tables **out) {
  if (!path || !out) return invalid("cmx_tables_create_from_file: null argument");
  FILE *f = fopen(path, "rb");
  if (!f) return invalid(std::string("cmx_tables_create_from_file: cannot open ") + path);
  std::vector<unsigned char> buf;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf.resize(n > 0 ? (size_t)n : 0);
  const size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), f);
  fclose(f);
  if (got != buf.size() || buf.size() < 8 + 48 || memcmp(buf.data(), "CMXT1\0\0\0", 8) != 0)
    return invalid(std::string("cmx_tables_create_from_file: not a flat tables file: ") + path);
  int32_t sz[11];
  memcpy(sz, buf.data() + 8, sizeof(sz));
  for (int q = 0; q < 11; ++q)
    if (sz[q] < 0) return invalid("cmx_tables_create_from_file: negative size");
  cmx_table_desc d;
  memset(&d, 0, sizeof(d));
  d.n_sublat = sz[0];
  d.max_occ = sz[1];
  d.n_func = sz[2];
  d.corr_size = sz[3];
  d.n_point_corr = sz[4];
  d.nlist_len = sz[5];
  d.n_nlist_sublat = sz[6];
  d.n_factors = sz[7];
  d.n_terms = sz[8];
  d.n_elems = sz[9];
  d.n_groups = sz[10];
  size_t off = 8 + 48;
  bool ok = true;
  auto take = [&](size_t count, size_t elem) -> const void * {
    const size_t bytes = count * elem;
    if (off + bytes > buf.size()) {
      ok = false;
      return nullptr;
    }
    const void *p = buf.data() + off;
    off += bytes + ((8 - bytes % 8) % 8);
    return p;
  };
  const size_t pc = (size_t)d.n_point_corr * d.corr_size + 1;
  d.nlist_sublat = (const int32_t *)take(d.n_nlist_sublat, 4);
  d.n_occ = (const int32_t *)take(d.n_sublat, 4);
  d.phi = (const double *)take((size_t)d.n_sublat * d.n_func * d.max_occ, 8);
  d.nbr = (const int32_t *)take((size_t)d.nlist_len * 4, 4);
  d.factor_f = (const int32_t *)take(d.n_factors, 4);
  d.factor_n = (const int32_t *)take(d.n_factors, 4);
  d.term_coef = (const double *)take(d.n_terms, 8);
  d.term_fbeg = (const int32_t *)take((size_t)d.n_terms + 1, 4);
  d.elem_tbeg = (const int32_t *)take((size_t)d.n_elems + 1, 4);
  d.group_ebeg = (const int32_t *)take((size_t)d.n_groups + 1, 4);
  d.group_dphi = (const int32_t *)take(d.n_groups, 4);
  d.group_has_sum = (const int32_t *)take(d.n_groups, 4);
  d.group_div = (const double *)take(d.n_groups, 8);
  d.global_gbeg = (const int32_t *)take((size_t)d.corr_size + 1, 4);
  d.point_gbeg = (const int32_t *)take(pc, 4);
  d.delta_gbeg = (const int32_t *)take(pc, 4);
  if (!ok) return invalid(std::string("cmx_tables_create_from_file: truncated file: ") + path);
  return cmx_tables_create(&d, device, out);
}

extern "C" void cmx_tables_destroy(cmx_tables *t) {
  if (!t) return;
  for (void *p : t->allocs) cudaFree(p);
  delete t;
}

// ---------------------------------------------------------------------------
extern "C" int cmx_state_create(const cmx_tables *t, int32_t N0, int32_t N1,
                                int32_t N2, int32_t n_replicas, int32_t halo,
                                cmx_state **out) {
  return cmx_state_create_opts(t, N0, N1, N2, n_replicas, halo, 0u, out);
}

// Hermite normal form of a transformation matrix T (columns = supercell lattice vectors in
// prim coordinates): the same lattice with the basis (h00, h10, h20), (0, h11, h21),
// (0, 0, h22), 0 <= h10 < h11, 0 <= h20, h21 < h22.  Integer column operations only.
static bool hermite_normal_form(const int32_t *T9, long long (&H)[3][3]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) H[r][c] = T9[3 * r + c];
  for (int r = 0; r < 3; ++r) {
    for (;;) {  // Euclid on row r over the columns r..2
      int n_nz = 0, c0 = -1;
      for (int c = r; c < 3; ++c)
        if (H[r][c] != 0) {
          ++n_nz;
          if (c0 < 0 || std::llabs(H[r][c]) < std::llabs(H[r][c0])) c0 = c;
        }
      if (n_nz == 0) return false;  // singular
      if (n_nz == 1) {
        if (c0 != r)
          for (int q = 0; q < 3; ++q) std::swap(H[q][r], H[q][c0]);
        break;
      }
      for (int c = r; c < 3; ++c)
        if (c != c0 && H[r][c] != 0) {
          const long long f = H[r][c] / H[r][c0];
          for (int q = 0; q < 3; ++q) H[q][c] -= f * H[q][c0];
        }
    }
    if (H[r][r] < 0)
      for (int q = 0; q < 3; ++q) H[q][r] = -H[q][r];
    for (int c = 0; c < r; ++c) {  // reduce the entries left of the diagonal into [0, h_rr)
      long long f = H[r][c] / H[r][r];
      if (H[r][c] - f * H[r][r] < 0) --f;
      for (int q = 0; q < 3; ++q) H[q][c] -= f * H[q][r];
    }
  }
  return true;
}

// host only: the box a transformation matrix gets (no device needed)
extern "C" int cmx_supercell_box(const int32_t *T9, int32_t *box6) {
  if (!T9 || !box6) return invalid("cmx_supercell_box: null argument");
  long long H[3][3];
  if (!hermite_normal_form(T9, H)) return invalid("cmx_supercell_box: singular transformation matrix");
  const long long v[6] = {H[0][0], H[1][1], H[2][2], H[1][0], H[2][0], H[2][1]};
  for (int q = 0; q < 6; ++q) {
    if (v[q] > 0x7fffffffll) return invalid("cmx_supercell_box: supercell too large");
    box6[q] = (int32_t)v[q];
  }
  return CMX_OK;
}

extern "C" int cmx_state_create_general(const cmx_tables *t, const int32_t *T9, int32_t n_replicas, uint32_t options,
                                        cmx_state **out) {
  if (!t || !T9 || !out) return invalid("cmx_state_create_general: null argument");
  long long H[3][3];
  if (!hermite_normal_form(T9, H)) return invalid("cmx_state_create_general: singular transformation matrix");
  for (int a = 0; a < 3; ++a)
    if (H[a][a] <= 0 || H[a][a] > 0x7fffffffll) return invalid("cmx_state_create_general: supercell too large");
  const bool skew = H[1][0] != 0 || H[2][0] != 0 || H[2][1] != 0;
  int rc = cmx_state_create_opts(t, (int32_t)H[0][0], (int32_t)H[1][1], (int32_t)H[2][2], n_replicas, 0,
                                 options | (skew ? CMX_STATE_LINEAR_ROWS | 0x80000000u : 0u), out);
  if (rc) return rc;
  (*out)->g.s10 = (int32_t)H[1][0];
  (*out)->g.s20 = (int32_t)H[2][0];
  (*out)->g.s21 = (int32_t)H[2][1];
  return CMX_OK;
}

// unit cells given by prim-lattice coordinates (any integers) -> cell index inside the box
extern "C" int cmx_state_cell_index(const cmx_state *s, int64_t n, const int32_t *ijk, int64_t *cell) {
  if (!s || n < 0 || (n && (!ijk || !cell))) return invalid("cmx_state_cell_index: bad argument");
  Geom g = s->g;
  g.halo = 0;
  for (int64_t q = 0; q < n; ++q) {
    int i = ijk[3 * q], j = ijk[3 * q + 1], k = ijk[3 * q + 2];
    if (!(g.s10 | g.s20 | g.s21)) {  // (the diag shortcut of cmx_wrap_cell wraps once only)
      i -= cmx_floor_div(i, g.N0) * g.N0;
      j -= cmx_floor_div(j, g.N1) * g.N1;
      k -= cmx_floor_div(k, g.N2) * g.N2;
    } else {
      cmx_wrap_cell(g, i, j, k);
    }
    cell[q] = (int64_t)i + (int64_t)g.N0 * ((int64_t)j + (int64_t)g.N1 * (int64_t)k);
  }
  return CMX_OK;
}

extern "C" int cmx_state_box(const cmx_state *s, int32_t *box6) {
  if (!s || !box6) return invalid("cmx_state_box: null argument");
  box6[0] = s->g.N0;
  box6[1] = s->g.N1;
  box6[2] = s->g.N2;
  box6[3] = s->g.s10;
  box6[4] = s->g.s20;
  box6[5] = s->g.s21;
  return CMX_OK;
}

// Site order of the caller (upload / download only): order[l_caller] = l of this library
// (l = b * n_cells + cell).  NULL restores the identity.
extern "C" int cmx_state_set_site_order(cmx_state *s, const int64_t *order) {
  if (!s) return invalid("cmx_state_set_site_order: null state");
  const int64_t n = s->g.n_cells * s->t->d.n_sublat;
  s->site_order.clear();
  if (!order) return CMX_OK;
  std::vector<char> seen((size_t)n, 0);
  for (int64_t l = 0; l < n; ++l) {
    if (order[l] < 0 || order[l] >= n || seen[(size_t)order[l]])
      return invalid("cmx_state_set_site_order: not a permutation of the sites");
    seen[(size_t)order[l]] = 1;
  }
  s->site_order.assign(order, order + n);
  return CMX_OK;
}

extern "C" int cmx_state_create_opts(const cmx_tables *t, int32_t N0, int32_t N1,
                                     int32_t N2, int32_t n_replicas, int32_t halo,
                                     uint32_t options, cmx_state **out) {
  if (!t || !out) return invalid("cmx_state_create: null argument");
  const bool general = (options & 0x80000000u) != 0;  // internal: a skewed box follows (no range check)
  options &= ~0x80000000u;
  if (options & ~(uint32_t)CMX_STATE_LINEAR_ROWS) return invalid("cmx_state_create_opts: unknown option");
  if (N0 <= 0 || N1 <= 0 || N2 <= 0 || n_replicas <= 0 || halo < 0)
    return invalid("cmx_state_create: non-positive dimension");
  // the neighbor arithmetic wraps once: every |offset| must be <= N
  for (int n = 0; n < t->d.nlist_len && !general; ++n) {
    const int32_t *o = &t->nbr[4 * n];
    if (std::abs(o[0]) > N0 || std::abs(o[1]) > N1 ||
        (!halo && std::abs(o[2]) > N2))
      return invalid("cmx_state_create: supercell smaller than the neighbor list range");
  }
  CMX_CUDA(cudaSetDevice(t->device));
  cmx_state *s = new cmx_state;
  s->t = t;
  s->n_replicas = n_replicas;
  Geom &g = s->g;
  g.N0 = N0;
  g.N1 = N1;
  g.N2 = N2;
  g.halo = halo;
  g.layer = (int64_t)N0 * N1;
  g.n_cells = g.layer * N2;
  g.sub_stride = g.layer * (N2 + 2 * halo);
  g.rep_stride = g.sub_stride * t->d.n_sublat;
  g.coded = (t->d.n_sublat == 1 && t->n_occ[0] == 3) ? 1 : 0;
  // x4-interleaved rows (see Geom::xq_log) for the shapes the streaming pair-LUT sweep covers
  g.s10 = g.s20 = g.s21 = 0;
  g.xq_log = 0;
  if (!(options & CMX_STATE_LINEAR_ROWS) && t->d.n_sublat == 1 && t->d.max_occ <= 3 && N0 >= 16 && N0 <= 512 &&
      (N0 & (N0 - 1)) == 0 && N1 % 2 == 0 && N2 % 2 == 0 && N2 <= 65534)
    while ((4 << g.xq_log) < N0) ++g.xq_log;
  size_t bytes = (size_t)g.rep_stride * n_replicas;
  cudaError_t e = cudaMalloc(&s->d_occ, bytes);
  if (e != cudaSuccess) {
    delete s;
    return cmx_cuda_fail(e, "cudaMalloc(occupation)");
  }
  cudaMemset(s->d_occ, 0, bytes);
  s->temperature.assign(n_replicas, 0.0);
  size_t ex = (size_t)t->d.n_sublat * t->d.max_occ * t->d.max_occ;
  s->exch.assign(ex * n_replicas, 0.0);
  cudaMalloc(&s->d_beta, sizeof(double) * n_replicas);
  cudaMalloc(&s->d_exch, sizeof(double) * ex * n_replicas);
  cudaMemset(s->d_beta, 0, sizeof(double) * n_replicas);
  cudaMemset(s->d_exch, 0, sizeof(double) * ex * n_replicas);
  cudaMalloc(&s->d_flag, 2 * sizeof(int));  // [0]: synchronous uploads, [1]: asynchronous uploads
  cudaMemset(s->d_flag, 0, 2 * sizeof(int));
  {
    const size_t sig_bytes = sizeof(unsigned long long) * 4 + sizeof(uint32_t) * (size_t)n_replicas * (N2 + 2);
    cudaMalloc((void **)&s->d_sig, sig_bytes);
    cudaMemset(s->d_sig, 0, sig_bytes);
    s->d_done = reinterpret_cast<uint32_t *>(s->d_sig + 4);
  }
  cudaMalloc(&s->d_counters, sizeof(cmx_counters) * n_replicas);
  cudaMemset(s->d_counters, 0, sizeof(cmx_counters) * n_replicas);
  cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    cmx_state_destroy(s);
    return cmx_cuda_fail(e, "cmx_state_create");
  }
  *out = s;
  return CMX_OK;
}

extern "C" void cmx_state_destroy(cmx_state *s) {
  if (!s) return;
  cudaSetDevice(s->t->device);
  cmx_canonical_free(s);
  cmx_plan_free(s->plan);
  cudaFree(s->d_occ);
  cudaFree(s->d_eci_idx);
  cudaFree(s->d_eci_val);
  cudaFree(s->d_beta);
  cudaFree(s->d_exch);
  cudaFree(s->d_counters);
  cudaFree(s->d_flag);
  for (void *m : s->ipc_open)
    if (m) cudaIpcCloseMemHandle(m);
  cudaFree(s->d_sig);
  cudaFree(s->d_scratch);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

int cmx_scratch(cmx_state *s, size_t bytes) {
  if (bytes <= s->scratch_bytes) return CMX_OK;
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  cudaFree(s->d_scratch);
  s->d_scratch = nullptr;
  s->scratch_bytes = 0;
  CMX_CUDA(cudaMalloc(&s->d_scratch, bytes));
  s->scratch_bytes = bytes;
  return CMX_OK;
}

// byte position of unit cell `cell` (reference order) inside the owned layers of a sublattice
__device__ __forceinline__ int64_t cmx_cell_pos(const Geom &g, int64_t cell) {
  if (!g.xq_log) return cell;
  const int64_t row = cell / g.N0;
  return row * g.N0 + cmx_xpos(g, (int)(cell - row * g.N0));
}

// int8 image in reference order <-> x4-interleaved rows, 16 bytes of a row per thread: the
// four words x = 4c + bQ .. +3 (b = 0..3) of the image are the 4x4 byte transpose of chunk c
// of the stored row.  TO_DEVICE: validate and encode; else decode.
__device__ __forceinline__ void cmx_transpose4x4(uint32_t (&w)[4]) {
  const uint32_t t0 = __byte_perm(w[0], w[1], 0x5140u), t1 = __byte_perm(w[2], w[3], 0x5140u);
  const uint32_t t2 = __byte_perm(w[0], w[1], 0x7362u), t3 = __byte_perm(w[2], w[3], 0x7362u);
  w[0] = __byte_perm(t0, t1, 0x5410u);
  w[1] = __byte_perm(t0, t1, 0x7632u);
  w[2] = __byte_perm(t2, t3, 0x5410u);
  w[3] = __byte_perm(t2, t3, 0x7632u);
}
template <bool TO_DEVICE>
__global__ void k_transfer_x4(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int32_t N0, int32_t xq_log,
                              int64_t n_rows, int n_occ, int coded, int *bad) {
  const uint32_t W = (uint32_t)N0 >> 4, q4 = (1u << xq_log) >> 2, wpr = (uint32_t)N0 >> 2;  // chunks, Q/4, words per row
  const int64_t n_chunks = n_rows * W;
  int flag = 0;
  for (int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; x < n_chunks; x += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = x / W;
    const uint32_t c = (uint32_t)(x - row * W);
    uint32_t w[4];
    if (TO_DEVICE) {
      const uint32_t *sr = src + row * wpr;
#pragma unroll
      for (int b = 0; b < 4; ++b) w[b] = sr[c + b * q4];
      cmx_transpose4x4(w);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int k = 0; k < 4; ++k) flag |= (int)((w[q] >> (8 * k)) & 0xffu) >= n_occ;
        if (coded) w[q] = w[q] | ((w[q] & 0x02020202u) << 3);  // 2 -> 18 per byte lane
      }
      *reinterpret_cast<uint4 *>(dst + row * wpr + 4 * c) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      const uint4 v = *reinterpret_cast<const uint4 *>(src + row * wpr + 4 * c);
      w[0] = v.x;
      w[1] = v.y;
      w[2] = v.z;
      w[3] = v.w;
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = (w[q] & 0x07070707u) | ((w[q] >> 3) & 0x03030303u);  // cmx_dec per lane
      cmx_transpose4x4(w);
      uint32_t *dr = dst + row * wpr;
#pragma unroll
      for (int b = 0; b < 4; ++b) dr[c + b * q4] = w[b];
    }
  }
  if (TO_DEVICE && flag) *bad = 1;
}

// reference layout (no ghost layers, int32 or int8) <-> device layout (int8).
// Occupant indices are validated on the device (a bad index would read outside
// the site-function tables): *bad is set when any is out of range.
template <typename SrcT>
__global__ void k_scatter_occ(const SrcT *__restrict__ src, int8_t *dst, Geom g,
                              int n_sublat, const int32_t *__restrict__ n_occ,
                              int *bad) {
  int64_t total = g.n_cells * n_sublat;
  int flag = 0;
  for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < total;
       l += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = l / g.n_cells, cell = l - b * g.n_cells;
    int v = (int)src[l];
    flag |= (v < 0 || v >= n_occ[b]);
    dst[b * g.sub_stride + g.halo * g.layer + cmx_cell_pos(g, cell)] = (int8_t)cmx_enc(g, v);
  }
  if (flag) *bad = 1;
}
// in-place validation of an int8 image copied straight into the state (16 sites
// per thread-iteration; n_cells is a multiple of 16 on this path)
// (and re-coded in place when the state stores occupant 2 as 18, Geom::coded)
__global__ void k_validate_occ16(int4 *occ, int64_t n16, int64_t cells16,
                                 const int32_t *__restrict__ n_occ, int coded,
                                 int *bad) {
  int flag = 0;
  for (int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; x < n16;
       x += (int64_t)gridDim.x * blockDim.x) {
    int no = n_occ[x / cells16];
    int4 v = occ[x];
    uint32_t w[4] = {(uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int k = 0; k < 4; ++k) flag |= (int)((w[q] >> (8 * k)) & 0xffu) >= no;
      // 2 -> 18 per byte lane: copy bit 1 to bit 4
      w[q] = w[q] | ((w[q] & 0x02020202u) << 3);
    }
    if (coded) occ[x] = make_int4((int)w[0], (int)w[1], (int)w[2], (int)w[3]);
  }
  if (flag) *bad = 1;
}
template <typename DstT>
__global__ void k_gather_occ(const int8_t *__restrict__ src, DstT *dst, Geom g,
                             int n_sublat) {
  int64_t total = g.n_cells * n_sublat;
  for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < total;
       l += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = l / g.n_cells, cell = l - b * g.n_cells;
    dst[l] = (DstT)cmx_dec(src[b * g.sub_stride + g.halo * g.layer + cmx_cell_pos(g, cell)]);
  }
}

static int check_replica(const cmx_state *s, int32_t r, const char *who) {
  if (!s) return invalid(std::string(who) + ": null state");
  if (r < 0 || r >= s->n_replicas)
    return invalid(std::string(who) + ": replica out of range");
  return CMX_OK;
}

template <typename T>
static int upload_occ(cmx_state *s, int32_t replica, const T *occ) {
  int rc = check_replica(s, replica, "cmx_state_upload_occ");
  if (rc) return rc;
  if (!occ) return invalid("cmx_state_upload_occ: null occupation");
  CMX_CUDA(cudaSetDevice(s->t->device));
  size_t n = (size_t)s->g.n_cells * s->t->d.n_sublat;
  int8_t *dst = s->d_occ + (size_t)replica * s->g.rep_stride;
  std::vector<T> permuted;
  if (!s->site_order.empty()) {  // the caller's site order -> ours
    permuted.resize(n);
    for (size_t l = 0; l < n; ++l) permuted[(size_t)s->site_order[l]] = occ[l];
    occ = permuted.data();
  }
  CMX_CUDA(cudaMemsetAsync(s->d_flag, 0, sizeof(int), s->stream));
  if (sizeof(T) == 1 && s->g.xq_log) {
    // x4-interleaved rows: the image goes to scratch and is transposed into place
    rc = cmx_scratch(s, n);
    if (rc) return rc;
    CMX_CUDA(cudaMemcpyAsync(s->d_scratch, occ, n, cudaMemcpyHostToDevice, s->stream));
    k_transfer_x4<true><<<1184, 256, 0, s->stream>>>((const uint32_t *)s->d_scratch,
                                                     (uint32_t *)(dst + (size_t)s->g.halo * s->g.layer), s->g.N0,
                                                     s->g.xq_log, (int64_t)s->g.N1 * s->g.N2, s->t->n_occ[0],
                                                     s->g.coded, s->d_flag);
  } else if (sizeof(T) == 1 && s->g.halo == 0 && s->g.n_cells % 16 == 0) {
    // device layout == reference layout: one DMA, validated in place.  On a
    // bad index the previous occupation is lost -- the caller gets an error.
    CMX_CUDA(cudaMemcpyAsync(dst, occ, n, cudaMemcpyHostToDevice, s->stream));
    k_validate_occ16<<<1184, 256, 0, s->stream>>>((int4 *)dst, (int64_t)(n / 16),
                                                  s->g.n_cells / 16, s->t->d.n_occ,
                                                  s->g.coded, s->d_flag);
  } else {
    rc = cmx_scratch(s, n * sizeof(T));
    if (rc) return rc;
    CMX_CUDA(cudaMemcpyAsync(s->d_scratch, occ, n * sizeof(T), cudaMemcpyHostToDevice, s->stream));
    k_scatter_occ<T><<<1184, 256, 0, s->stream>>>((const T *)s->d_scratch, dst, s->g,
                                                  s->t->d.n_sublat, s->t->d.n_occ, s->d_flag);
  }
  CMX_CUDA(cudaGetLastError());
  int bad = 0;
  CMX_CUDA(cudaMemcpyAsync(&bad, s->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  if (bad) return invalid("cmx_state_upload_occ: occupant index out of range");
  return CMX_OK;
}

template <typename T>
static int download_occ(const cmx_state *cs, int32_t replica, T *occ) {
  cmx_state *s = const_cast<cmx_state *>(cs);
  int rc = check_replica(s, replica, "cmx_state_download_occ");
  if (rc) return rc;
  if (!occ) return invalid("cmx_state_download_occ: null occupation");
  CMX_CUDA(cudaSetDevice(s->t->device));
  size_t n = (size_t)s->g.n_cells * s->t->d.n_sublat;
  const int8_t *src = s->d_occ + (size_t)replica * s->g.rep_stride;
  if (sizeof(T) == 1 && s->g.halo == 0 && !s->g.coded && !s->g.xq_log && s->site_order.empty()) {
    CMX_CUDA(cudaMemcpyAsync(occ, src, n, cudaMemcpyDeviceToHost, s->stream));
    CMX_CUDA(cudaStreamSynchronize(s->stream));
    return CMX_OK;
  }
  rc = cmx_scratch(s, n * sizeof(T));
  if (rc) return rc;
  if (sizeof(T) == 1 && s->g.xq_log)
    k_transfer_x4<false><<<1184, 256, 0, s->stream>>>((const uint32_t *)(src + (size_t)s->g.halo * s->g.layer),
                                                      (uint32_t *)s->d_scratch, s->g.N0, s->g.xq_log,
                                                      (int64_t)s->g.N1 * s->g.N2, 0, 0, nullptr);
  else
    k_gather_occ<T><<<1184, 256, 0, s->stream>>>(src, (T *)s->d_scratch, s->g,
                                                 s->t->d.n_sublat);
  CMX_CUDA(cudaGetLastError());
  CMX_CUDA(cudaMemcpyAsync(occ, s->d_scratch, n * sizeof(T),
                           cudaMemcpyDeviceToHost, s->stream));
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  if (!s->site_order.empty()) {  // ours -> the caller's site order
    std::vector<T> tmp(occ, occ + n);
    for (size_t l = 0; l < n; ++l) occ[l] = tmp[(size_t)s->site_order[l]];
  }
  return CMX_OK;
}

// ---- asynchronous transfers: a pipeline of independent states (one stream each) overlaps
// the upload of one job, the sweeps of another and the download of a third.  Nothing here
// synchronises; cmx_state_synchronize waits and reports what went wrong meanwhile.
extern "C" int cmx_state_upload_occ_i8_async(cmx_state *s, int32_t replica, const int8_t *occ) {
  int rc = check_replica(s, replica, "cmx_state_upload_occ_i8_async");
  if (rc) return rc;
  if (!occ) return invalid("cmx_state_upload_occ_i8_async: null occupation");
  if (!s->site_order.empty()) {
    cmx_set_error("cmx_state_upload_occ_i8_async: a caller site order is set (use the synchronous transfer)");
    return CMX_ERR_UNSUPPORTED;
  }
  CMX_CUDA(cudaSetDevice(s->t->device));
  size_t n = (size_t)s->g.n_cells * s->t->d.n_sublat;
  int8_t *dst = s->d_occ + (size_t)replica * s->g.rep_stride;
  if (s->g.xq_log) {
    rc = cmx_scratch(s, n);
    if (rc) return rc;
    CMX_CUDA(cudaMemcpyAsync(s->d_scratch, occ, n, cudaMemcpyHostToDevice, s->stream));
    k_transfer_x4<true><<<1184, 256, 0, s->stream>>>((const uint32_t *)s->d_scratch,
                                                     (uint32_t *)(dst + (size_t)s->g.halo * s->g.layer), s->g.N0,
                                                     s->g.xq_log, (int64_t)s->g.N1 * s->g.N2, s->t->n_occ[0],
                                                     s->g.coded, s->d_flag + 1);
  } else if (s->g.halo == 0 && s->g.n_cells % 16 == 0) {
    CMX_CUDA(cudaMemcpyAsync(dst, occ, n, cudaMemcpyHostToDevice, s->stream));
    k_validate_occ16<<<1184, 256, 0, s->stream>>>((int4 *)dst, (int64_t)(n / 16), s->g.n_cells / 16,
                                                  s->t->d.n_occ, s->g.coded, s->d_flag + 1);
  } else {
    rc = cmx_scratch(s, n);
    if (rc) return rc;
    CMX_CUDA(cudaMemcpyAsync(s->d_scratch, occ, n, cudaMemcpyHostToDevice, s->stream));
    k_scatter_occ<int8_t><<<1184, 256, 0, s->stream>>>((const int8_t *)s->d_scratch, dst, s->g, s->t->d.n_sublat,
                                                       s->t->d.n_occ, s->d_flag + 1);
  }
  CMX_CUDA(cudaGetLastError());
  s->async_upload_pending = true;
  return CMX_OK;
}

extern "C" int cmx_state_download_occ_i8_async(cmx_state *s, int32_t replica, int8_t *occ) {
  int rc = check_replica(s, replica, "cmx_state_download_occ_i8_async");
  if (rc) return rc;
  if (!occ) return invalid("cmx_state_download_occ_i8_async: null occupation");
  if (!s->site_order.empty()) {
    cmx_set_error("cmx_state_download_occ_i8_async: a caller site order is set (use the synchronous transfer)");
    return CMX_ERR_UNSUPPORTED;
  }
  CMX_CUDA(cudaSetDevice(s->t->device));
  size_t n = (size_t)s->g.n_cells * s->t->d.n_sublat;
  const int8_t *src = s->d_occ + (size_t)replica * s->g.rep_stride;
  if (s->g.halo == 0 && !s->g.coded && !s->g.xq_log) {
    CMX_CUDA(cudaMemcpyAsync(occ, src, n, cudaMemcpyDeviceToHost, s->stream));
    return CMX_OK;
  }
  rc = cmx_scratch(s, n);
  if (rc) return rc;
  if (s->g.xq_log)
    k_transfer_x4<false><<<1184, 256, 0, s->stream>>>((const uint32_t *)(src + (size_t)s->g.halo * s->g.layer),
                                                      (uint32_t *)s->d_scratch, s->g.N0, s->g.xq_log,
                                                      (int64_t)s->g.N1 * s->g.N2, 0, 0, nullptr);
  else
    k_gather_occ<int8_t><<<1184, 256, 0, s->stream>>>(src, (int8_t *)s->d_scratch, s->g, s->t->d.n_sublat);
  CMX_CUDA(cudaGetLastError());
  CMX_CUDA(cudaMemcpyAsync(occ, s->d_scratch, n, cudaMemcpyDeviceToHost, s->stream));
  return CMX_OK;
}

// wait for everything enqueued on the state's stream; an occupant index out of range in
// an asynchronous upload since the last call is reported here
extern "C" int cmx_state_synchronize(cmx_state *s) {
  if (!s) return invalid("cmx_state_synchronize: null state");
  CMX_CUDA(cudaSetDevice(s->t->device));
  int bad = 0;
  if (s->async_upload_pending) {
    CMX_CUDA(cudaMemcpyAsync(&bad, s->d_flag + 1, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CMX_CUDA(cudaMemsetAsync(s->d_flag + 1, 0, sizeof(int), s->stream));
  }
  // (a sweep that gave up waiting for a ring neighbour or a layer counter set this word and
  // went on: callers that never read the counters learn it here)
  unsigned long long timed_out = 0;
  CMX_CUDA(cudaMemcpyAsync(&timed_out, s->d_sig + 3, sizeof(timed_out), cudaMemcpyDeviceToHost, s->stream));
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  s->async_upload_pending = false;
  if (bad) return invalid("cmx_state_synchronize: occupant index out of range in an asynchronous upload");
  if (timed_out) {
    cmx_set_error("cmx_state_synchronize: a sweep timed out waiting for a ring neighbour's epoch or a layer counter "
                  "(grid not co-resident, or a neighbour did not run the same sweeps); the occupation is not valid");
    return CMX_ERR_CUDA;
  }
  return CMX_OK;
}

extern "C" int cmx_state_upload_occ(cmx_state *s, int32_t r, const int32_t *o) {
  return upload_occ<int32_t>(s, r, o);
}
extern "C" int cmx_state_upload_occ_i8(cmx_state *s, int32_t r, const int8_t *o) {
  return upload_occ<int8_t>(s, r, o);
}
extern "C" int cmx_state_download_occ(const cmx_state *s, int32_t r, int32_t *o) {
  return download_occ<int32_t>(s, r, o);
}
extern "C" int cmx_state_download_occ_i8(const cmx_state *s, int32_t r, int8_t *o) {
  return download_occ<int8_t>(s, r, o);
}

// i.i.d. uniform occupation; counter = (site, replica), key = seed
__global__ void k_randomize(int8_t *occ, Geom g, int n_sublat, int n_replicas,
                            const int32_t *__restrict__ n_occ, uint32_t k0,
                            uint32_t k1) {
  int64_t per = g.n_cells * n_sublat;
  int64_t total = per * n_replicas;
  for (int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; x < total;
       x += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = x / per, l = x - r * per;
    int64_t b = l / g.n_cells, cell = l - b * g.n_cells;
    int no = n_occ[b];
    int v = 0;
    if (no > 1) {
      Philox p = philox4x32_10((uint32_t)l, (uint32_t)(l >> 32), (uint32_t)r,
                               0x52414e44u, k0, k1);
      v = (int)__umulhi(p.c[0], (uint32_t)no);
    }
    occ[r * g.rep_stride + b * g.sub_stride + g.halo * g.layer + cmx_cell_pos(g, cell)] = (int8_t)cmx_enc(g, v);
  }
}

extern "C" int cmx_state_randomize(cmx_state *s, uint64_t seed) {
  if (!s) return invalid("cmx_state_randomize: null state");
  CMX_CUDA(cudaSetDevice(s->t->device));
  k_randomize<<<1184, 256, 0, s->stream>>>(s->d_occ, s->g, s->t->d.n_sublat,
                                           s->n_replicas, s->t->d.n_occ,
                                           (uint32_t)seed, (uint32_t)(seed >> 32));
  CMX_CUDA(cudaGetLastError());
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  return CMX_OK;
}

extern "C" int cmx_state_set_k_offset(cmx_state *s, int32_t k_offset) {
  if (!s || k_offset < 0) return invalid("cmx_state_set_k_offset: bad argument");
  s->k_offset = k_offset;
  return CMX_OK;
}

extern "C" int cmx_state_stream(cmx_state *s, void **stream) {
  if (!s || !stream) return invalid("cmx_state_stream: null argument");
  *stream = (void *)s->stream;
  return CMX_OK;
}

extern "C" int cmx_state_device_ptr(cmx_state *s, void **d_ptr, size_t *n_bytes) {
  if (!s || !d_ptr) return invalid("cmx_state_device_ptr: null argument");
  *d_ptr = s->d_occ;
  if (n_bytes) *n_bytes = (size_t)s->g.rep_stride * s->n_replicas;
  return CMX_OK;
}

// ---- NVLink peer memory -------------------------------------------------------
struct IpcBlob {
  cudaIpcMemHandle_t occ, sig;
};
static_assert(sizeof(IpcBlob) <= CMX_IPC_HANDLE_BYTES, "handle size");

extern "C" int cmx_state_ipc_export(cmx_state *s, void *handle) {
  if (!s || !handle) return invalid("cmx_state_ipc_export: null argument");
  CMX_CUDA(cudaSetDevice(s->t->device));
  IpcBlob b;
  memset(&b, 0, sizeof(b));
  CMX_CUDA(cudaIpcGetMemHandle(&b.occ, s->d_occ));
  CMX_CUDA(cudaIpcGetMemHandle(&b.sig, s->d_sig));
  memset(handle, 0, CMX_IPC_HANDLE_BYTES);
  memcpy(handle, &b, sizeof(b));
  return CMX_OK;
}

extern "C" int cmx_state_ipc_attach(cmx_state *s, const void *handle_dn, const void *handle_up) {
  if (!s) return invalid("cmx_state_ipc_attach: null state");
  if (!s->g.halo) return invalid("cmx_state_ipc_attach: not a slab state (halo = 0)");
  if (s->p2p) return invalid("cmx_state_ipc_attach: already attached");
  CMX_CUDA(cudaSetDevice(s->t->device));
  auto open_one = [&](const void *h, int slot, int8_t **occ, unsigned long long **sig) -> int {
    if (!h) {  // the neighbour is this state
      *occ = s->d_occ;
      *sig = s->d_sig;
      return CMX_OK;
    }
    IpcBlob b;
    memcpy(&b, h, sizeof(b));
    void *po = nullptr, *ps = nullptr;
    CMX_CUDA(cudaIpcOpenMemHandle(&po, b.occ, cudaIpcMemLazyEnablePeerAccess));
    s->ipc_open[slot] = po;
    CMX_CUDA(cudaIpcOpenMemHandle(&ps, b.sig, cudaIpcMemLazyEnablePeerAccess));
    s->ipc_open[slot + 1] = ps;
    *occ = (int8_t *)po;
    *sig = (unsigned long long *)ps;
    return CMX_OK;
  };
  int rc = open_one(handle_dn, 0, &s->peer_occ_dn, &s->peer_sig_dn);
  if (rc) return rc;
  if (handle_dn && handle_up && memcmp(handle_dn, handle_up, sizeof(IpcBlob)) == 0) {
    s->peer_occ_up = s->peer_occ_dn;  // two slabs: both neighbours are the same rank
    s->peer_sig_up = s->peer_sig_dn;
  } else {
    rc = open_one(handle_up, 2, &s->peer_occ_up, &s->peer_sig_up);
    if (rc) return rc;
  }
  s->peer_done_dn = reinterpret_cast<uint32_t *>(s->peer_sig_dn + 4);
  s->peer_done_up = reinterpret_cast<uint32_t *>(s->peer_sig_up + 4);
  // (the layer counters are NOT reset here: a neighbour that attached first may already
  // count on them; they are zero from creation and only ever advance in step with
  // done_even / done_odd)
  s->p2p = true;
  // the sweep kernel variant (and with it the co-resident grid) changes: recompute the schedule
  s->plan.stream_capacity = -1;
  s->plan.part_blocks = 0;
  for (auto &kv : s->plan.stream_lists) cudaFree(kv.second.d_units);
  s->plan.stream_lists.clear();
  return CMX_OK;
}

extern "C" int cmx_state_p2p_active(const cmx_state *s, int32_t *active) {
  if (!s || !active) return invalid("cmx_state_p2p_active: null argument");
  *active = (s->p2p && s->plan.valid && s->plan.pair_lut && s->plan.stream &&
             !(s->sweep_flags & CMX_SWEEP_FORCE_GENERIC)) ? 1 : 0;
  return CMX_OK;
}

extern "C" int cmx_state_set_eci(cmx_state *s, int32_t n, const uint32_t *index,
                                 const double *value) {
  if (!s || n < 0 || (n && (!index || !value)))
    return invalid("cmx_state_set_eci: bad argument");
  for (int i = 0; i < n; ++i)
    if (index[i] >= (uint32_t)s->t->d.corr_size)
      return invalid("cmx_state_set_eci: coefficient index out of range");
  CMX_CUDA(cudaSetDevice(s->t->device));
  CMX_CUDA(cudaStreamSynchronize(s->stream));  // sweeps in flight still read the old coefficients
  cudaFree(s->d_eci_idx);
  cudaFree(s->d_eci_val);
  s->d_eci_idx = nullptr;
  s->d_eci_val = nullptr;
  s->n_eci = n;
  s->eci_idx.assign(index, index + n);
  s->eci_val.assign(value, value + n);
  CMX_CUDA(cudaMalloc(&s->d_eci_idx, sizeof(uint32_t) * (n ? n : 1)));
  CMX_CUDA(cudaMalloc(&s->d_eci_val, sizeof(double) * (n ? n : 1)));
  if (n) {
    CMX_CUDA(cudaMemcpy(s->d_eci_idx, index, sizeof(uint32_t) * n, cudaMemcpyHostToDevice));
    CMX_CUDA(cudaMemcpy(s->d_eci_val, value, sizeof(double) * n, cudaMemcpyHostToDevice));
  }
  cmx_canonical_free(s);  // swap colourings depend on the active neighborhood
  return cmx_plan_sweep(s);
}

extern "C" int cmx_state_set_conditions(cmx_state *s, int32_t replica,
                                        double temperature, const double *exch) {
  int rc = check_replica(s, replica, "cmx_state_set_conditions");
  if (rc) return rc;
  if (!(temperature > 0.0)) return invalid("cmx_state_set_conditions: temperature must be > 0");
  CMX_CUDA(cudaSetDevice(s->t->device));
  size_t ex = (size_t)s->t->d.n_sublat * s->t->d.max_occ * s->t->d.max_occ;
  s->temperature[replica] = temperature;
  double beta = 1.0 / (CMX_KB * temperature);
  for (size_t i = 0; i < ex; ++i) s->exch[replica * ex + i] = exch ? exch[i] : 0.0;
  // s->stream is non-blocking: order the update against sweeps already enqueued on it
  // (sources are pageable host memory: the copies are staged before they return)
  CMX_CUDA(cudaMemcpyAsync(s->d_beta + replica, &beta, sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CMX_CUDA(cudaMemcpyAsync(s->d_exch + replica * ex, &s->exch[replica * ex],
                           sizeof(double) * ex, cudaMemcpyHostToDevice, s->stream));
  CMX_CUDA(cudaStreamSynchronize(s->stream));
  // the pair-LUT acceptance tables (and the dE - exch table) are functions of (beta, exch)
  s->plan.thr_dirty = true;
  s->plan.pdl_ok = false;
  return CMX_OK;
}

extern "C" int cmx_state_set_occupants(cmx_state *s, const int32_t *sublat_to_asym,
                                       const int32_t *occ_to_species,
                                       int32_t n_species) {
  if (!s || !sublat_to_asym || !occ_to_species || n_species <= 0)
    return invalid("cmx_state_set_occupants: bad argument");
  int nb = s->t->d.n_sublat, mo = s->t->d.max_occ;
  for (int b = 0; b < nb; ++b) {
    if (sublat_to_asym[b] < 0 || sublat_to_asym[b] >= nb)
      return invalid("cmx_state_set_occupants: asym index out of range");
    for (int o = 0; o < s->t->n_occ[b]; ++o)
      if (occ_to_species[b * mo + o] < 0 || occ_to_species[b * mo + o] >= n_species)
        return invalid("cmx_state_set_occupants: species index out of range");
  }
  s->sublat_to_asym.assign(sublat_to_asym, sublat_to_asym + nb);
  s->occ_to_species.assign(occ_to_species, occ_to_species + (size_t)nb * mo);
  s->n_species = n_species;
  return CMX_OK;
}
