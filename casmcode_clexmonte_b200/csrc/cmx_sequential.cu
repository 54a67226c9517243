// Reference-order (sequential) Metropolis on the device.
//
// One warp walks the exact sequence of
//   methods/occupation_metropolis.hh:92-120
// with the reference's random stream (std::mt19937_64 + libstdc++
// uniform_int_distribution / uniform_real_distribution, consumed in the same
// order as [EXT] monte::propose_semigrand_canonical_event /
// propose_canonical_event / metropolis_acceptance), the OccLocation
// swap-and-pop bookkeeping ([EXT] OccLocation::apply), and the faithful
// (individually rounded, reference-order) delta-E evaluator -- so the
// occupation trajectory is reproduced bit for bit.  Lane 0 owns the RNG and
// the lists; the lanes share the restricted delta-correlation evaluation
// (one ECI slot per lane), combined by lane 0 in coefficient order.
//
// "bit-exact" is defined relative to libstdc++ (gcc 13) on x86-64, see
// SURVEY.md Appendix B.  The only operation not under our control is exp():
// CUDA's and glibc's results can differ in the last ulp, which changes an
// accept/reject decision only if the uniform draw falls between them
// (probability ~1e-16 per step).
#include <algorithm>

#include "cmx_internal.cuh"

static int invalid(const std::string &msg) {
  cmx_set_error(msg);
  return CMX_ERR_INVALID;
}

#include "cmx_mt64.cuh"

// test hook: replay a stream of draws on the device (kind 0 raw, 1 int, 2 real)
__global__ void k_rng_stream(unsigned long long seed, long long n,
                             const long long *int_max, const double *real_max,
                             const int *kind, long long *out_int, double *out_real,
                             unsigned long long *out_raw) {
  __shared__ Mt64 g;
  if (threadIdx.x != 0) return;
  mt_seed(g, seed);
  for (long long i = 0; i < n; ++i) {
    if (kind[i] == 0) out_raw[i] = mt_next(g);
    else if (kind[i] == 1) out_int[i] = mt_int(g, int_max[i]);
    else out_real[i] = mt_real(g, real_max[i]);
  }
}

extern "C" int cmx_rng_stream_test(uint64_t seed, int64_t n, const int64_t *int_max,
                                   const double *real_max, const int32_t *kind,
                                   int64_t *out_int, double *out_real, uint64_t *out_raw) {
  if (n <= 0 || !int_max || !real_max || !kind || !out_int || !out_real || !out_raw)
    return invalid("cmx_rng_stream_test: bad argument");
  void *d[6];
  size_t sz[6] = {sizeof(int64_t) * n, sizeof(double) * n, sizeof(int32_t) * n,
                  sizeof(int64_t) * n, sizeof(double) * n, sizeof(uint64_t) * n};
  for (int i = 0; i < 6; ++i) CMX_CUDA(cudaMalloc(&d[i], sz[i]));
  CMX_CUDA(cudaMemcpy(d[0], int_max, sz[0], cudaMemcpyHostToDevice));
  CMX_CUDA(cudaMemcpy(d[1], real_max, sz[1], cudaMemcpyHostToDevice));
  CMX_CUDA(cudaMemcpy(d[2], kind, sz[2], cudaMemcpyHostToDevice));
  for (int i = 3; i < 6; ++i) CMX_CUDA(cudaMemset(d[i], 0, sz[i]));
  k_rng_stream<<<1, 32>>>(seed, n, (const long long *)d[0], (const double *)d[1],
                          (const int *)d[2], (long long *)d[3], (double *)d[4],
                          (unsigned long long *)d[5]);
  CMX_CUDA(cudaGetLastError());
  CMX_CUDA(cudaMemcpy(out_int, d[3], sz[3], cudaMemcpyDeviceToHost));
  CMX_CUDA(cudaMemcpy(out_real, d[4], sz[4], cudaMemcpyDeviceToHost));
  CMX_CUDA(cudaMemcpy(out_raw, d[5], sz[5], cudaMemcpyDeviceToHost));
  for (int i = 0; i < 6; ++i) cudaFree(d[i]);
  return CMX_OK;
}

// ---------------------------------------------------------------------------
struct SeqArgs {
  DevTables T;
  Geom g;
  int8_t *occ;
  int mode;  // 0 semi-grand, 1 canonical
  long long n_steps;
  unsigned long long seed;
  double beta;
  const double *exch;  // [n_sublat][max_occ][max_occ]
  int n_eci;
  const uint32_t *eci_idx;
  const double *eci_val;
  // OccLocation
  int n_cand, n_swap;
  const int *cand_asym, *cand_species;  // [n_cand]
  const int *swap_a, *swap_b;           // [n_swap]
  const int *occ_index;                 // [n_asym][n_species] -> occupant index or -1
  int n_species;
  const int *sublat_to_asym;            // [n_sublat]
  long long *cand_size;                 // [n_cand]
  long long *loc;                       // [n_cand][cap]
  long long cap;
  long long *mol_l, *mol_loc;           // [n_mol]
  int *mol_species;                     // [n_mol]
  // output
  cmx_step_record *log;
  long long log_cap;
  long long *n_accept;
  unsigned long long *hash;
  int *status;
  // tie report: [0] = number of near ties, [1 + q] = step of the q-th (q < CMX_TIE_CAP)
  long long *ties;
};
#define CMX_TIE_CAP 16

__device__ __forceinline__ unsigned long long fnv_step(unsigned long long h, unsigned long long x) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h ^= (x >> (8 * i)) & 0xffull;
    h *= 1099511628211ull;
  }
  return h;
}

__global__ void __launch_bounds__(32) k_metropolis_sequential(SeqArgs a) {
  __shared__ Mt64 rng;
  __shared__ double sh_dcorr[64];
  __shared__ double sh_tsum[65];
  const int lane = threadIdx.x;
  const DevTables &T = a.T;
  const Geom &g = a.g;
  if (lane == 0) mt_seed(rng, a.seed);
  __syncwarp();
  long long n_accept = 0;
  unsigned long long h = 1469598103934665603ull;
  const int nsite = a.mode == 0 ? 1 : 2;

  for (long long step = 0; step < a.n_steps; ++step) {
    long long ls[2] = {0, 0}, mols[2] = {0, 0};
    int nw[2] = {0, 0}, to_sp[2] = {0, 0};
    int fail = 0;
    if (lane == 0) {
      // ---- choose swap: weights cand_size(a) [* cand_size(b)]
      sh_tsum[0] = 0.0;
      for (int i = 0; i < a.n_swap; ++i) {
        double w = (double)a.cand_size[a.swap_a[i]];
        if (a.mode == 1) w = __dmul_rn(w, (double)a.cand_size[a.swap_b[i]]);
        sh_tsum[i + 1] = __dadd_rn(sh_tsum[i], w);
      }
      double tot = sh_tsum[a.n_swap];
      if (tot == 0.0) {
        fail = 1;
      } else {
        double r = mt_real(rng, tot);
        int si = 0;
        for (; si < a.n_swap; ++si)
          if (r < sh_tsum[si + 1]) break;
        if (si == a.n_swap) {
          fail = 2;
        } else {
          int ca = a.swap_a[si], cb = a.swap_b[si];
          long long pos = mt_int(rng, a.cand_size[ca] - 1);
          mols[0] = a.loc[(long long)ca * a.cap + pos];
          ls[0] = a.mol_l[mols[0]];
          to_sp[0] = a.cand_species[cb];
          nw[0] = a.occ_index[a.cand_asym[ca] * a.n_species + to_sp[0]];
          if (a.mode == 1) {
            long long pos2 = mt_int(rng, a.cand_size[cb] - 1);
            mols[1] = a.loc[(long long)cb * a.cap + pos2];
            ls[1] = a.mol_l[mols[1]];
            to_sp[1] = a.cand_species[ca];
            nw[1] = a.occ_index[a.cand_asym[cb] * a.n_species + to_sp[1]];
          }
        }
      }
    }
    fail = __shfl_sync(0xffffffffu, fail, 0);
    if (fail) {
      if (lane == 0) *a.status = fail;
      return;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      ls[q] = __shfl_sync(0xffffffffu, ls[q], 0);
      nw[q] = __shfl_sync(0xffffffffu, nw[q], 0);
    }
    // ---- restricted delta correlations, one ECI slot per lane
    for (int e0 = 0; e0 < a.n_eci; e0 += 32) {
      int e = e0 + lane;
      if (e < a.n_eci) {
        int c = (int)a.eci_idx[e];
        Override ov;
        ov.n = 0;
        double acc = 0.0;
        for (int q = 0; q < nsite; ++q) {
          int b = (int)(ls[q] / g.n_cells);
          long long cell = ls[q] - (long long)b * g.n_cells;
          int i = (int)(cell % g.N0);
          long long rr = cell / g.N0;
          int j = (int)(rr % g.N1), k = (int)(rr / g.N1);
          int64_t off = cmx_site_offset(g, b, i, j, k);
          int oi = cmx_load_occ(a.occ, off, ov);
          int p = -1;
          for (int pp = 0; pp < T.n_nlist_sublat; ++pp)
            if (T.nlist_sublat[pp] == b) p = pp;
          double d = 0.0;
          if (p >= 0) {
            int fi = p * T.corr_size + c;
            d = cmx_eval_function(T, g, a.occ, T.delta_gbeg[fi], T.delta_gbeg[fi + 1], i, j, k,
                                  ov, b, oi, nw[q]);
          }
          acc = (q == 0) ? d : __dadd_rn(acc, d);
          ov.off[ov.n] = off;
          ov.occ[ov.n] = nw[q];
          ov.n++;
        }
        sh_dcorr[e & 63] = acc;
      }
      // (n_eci <= 64 enforced by the host)
    }
    __syncwarp();
    int accept = 0;
    double dE = 0.0;
    if (lane == 0) {
      for (int e = 0; e < a.n_eci; ++e)
        dE = __dadd_rn(dE, __dmul_rn(a.eci_val[e], sh_dcorr[e]));
      if (a.mode == 0) {
        int b = (int)(ls[0] / g.n_cells);
        long long cell = ls[0] - (long long)b * g.n_cells;
        int i = (int)(cell % g.N0);
        long long rr = cell / g.N0;
        int j = (int)(rr % g.N1), k = (int)(rr / g.N1);
        int oi = cmx_dec(a.occ[cmx_site_offset(g, b, i, j, k)]);
        dE = __dsub_rn(dE, a.exch[(b * T.max_occ + oi) * T.max_occ + nw[0]]);
      }
      if (dE < 0.0) {
        accept = 1;
      } else {
        double u = mt_real(rng, 1.0);
        const double p = exp(__dmul_rn(-dE, a.beta));
        accept = u < p;
        // near tie: the decision would flip if exp() were one ulp off (CUDA's and glibc's
        // exp may differ in the last place) -- the only way this mode can leave the
        // reference's trajectory; counted and reported, never hidden
        if (u == p || u == nextafter(p, 0.0) || u == nextafter(p, 2.0)) {
          const long long q = a.ties[0]++;
          if (q < CMX_TIE_CAP) a.ties[1 + q] = step;
        }
      }
      if (step < a.log_cap) {
        cmx_step_record &rec = a.log[step];
        rec.l0 = ls[0];
        rec.l1 = a.mode == 1 ? ls[1] : -1;
        rec.new0 = nw[0];
        rec.new1 = a.mode == 1 ? nw[1] : -1;
        rec.accepted = accept;
        rec.pad = 0;
        rec.dE = dE;
      }
      h = fnv_step(h, (unsigned long long)ls[0] * 2ull + (accept ? 1ull : 0ull));
      if (accept) {
        ++n_accept;
        for (int q = 0; q < nsite; ++q) {
          long long m = mols[q];
          long long l = a.mol_l[m];
          int b = (int)(l / g.n_cells);
          long long cell = l - (long long)b * g.n_cells;
          int i = (int)(cell % g.N0);
          long long rr = cell / g.N0;
          int j = (int)(rr % g.N1), k = (int)(rr / g.N1);
          int asym = a.sublat_to_asym[b];
          a.occ[cmx_site_offset(g, b, i, j, k)] =
              (int8_t)cmx_enc(g, a.occ_index[asym * a.n_species + to_sp[q]]);
          // remove from the old candidate list: swap with last, pop
          int ci = -1;
          for (int cc = 0; cc < a.n_cand; ++cc)
            if (a.cand_asym[cc] == asym && a.cand_species[cc] == a.mol_species[m]) ci = cc;
          long long back = a.loc[(long long)ci * a.cap + a.cand_size[ci] - 1];
          a.loc[(long long)ci * a.cap + a.mol_loc[m]] = back;
          a.mol_loc[back] = a.mol_loc[m];
          a.cand_size[ci] -= 1;
          a.mol_species[m] = to_sp[q];
          int cj = -1;
          for (int cc = 0; cc < a.n_cand; ++cc)
            if (a.cand_asym[cc] == asym && a.cand_species[cc] == to_sp[q]) cj = cc;
          a.mol_loc[m] = a.cand_size[cj];
          a.loc[(long long)cj * a.cap + a.cand_size[cj]] = m;
          a.cand_size[cj] += 1;
        }
        __threadfence_block();
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    *a.n_accept = n_accept;
    *a.hash = h;
    *a.status = 0;
  }
}

template <typename T>
static int dev_copy(const std::vector<T> &v, T **d, std::vector<void *> &allocs) {
  size_t bytes = (v.size() ? v.size() : 1) * sizeof(T);
  CMX_CUDA(cudaMalloc((void **)d, bytes));
  allocs.push_back(*d);
  if (!v.empty()) CMX_CUDA(cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return CMX_OK;
}

extern "C" int cmx_metropolis_sequential(cmx_state *s, int32_t replica, int32_t mode,
                                         int64_t n_steps, uint64_t seed,
                                         cmx_step_record *log, int64_t log_cap,
                                         int64_t *n_accept, uint64_t *hash) {
  if (!s) return invalid("cmx_metropolis_sequential: null state");
  if (replica < 0 || replica >= s->n_replicas)
    return invalid("cmx_metropolis_sequential: replica out of range");
  if (mode != 0 && mode != 1) return invalid("cmx_metropolis_sequential: mode must be 0 or 1");
  if (n_steps < 0 || log_cap < 0) return invalid("cmx_metropolis_sequential: negative count");
  if (s->g.halo) return invalid("cmx_metropolis_sequential: not available on slab states");
  if (!s->d_eci_idx || s->sublat_to_asym.empty() || !(s->temperature[replica] > 0.0)) {
    cmx_set_error("cmx_metropolis_sequential: ECI, occupants and conditions must be set first");
    return CMX_ERR_STATE;
  }
  if (s->n_eci > 64) {
    cmx_set_error("cmx_metropolis_sequential: more than 64 coefficients unsupported");
    return CMX_ERR_UNSUPPORTED;
  }
  CMX_CUDA(cudaSetDevice(s->t->device));
  const DevTables &T = s->t->d;
  const int nb = T.n_sublat, mo = T.max_occ, nsp = s->n_species;
  const int64_t n_cells = s->g.n_cells;
  // ---- OccCandidateList / OccLocation::initialize [EXT], SURVEY.md App. B
  int n_asym = 0;
  for (int b = 0; b < nb; ++b) n_asym = std::max(n_asym, s->sublat_to_asym[b] + 1);
  std::vector<std::vector<int>> occ_to_sp(n_asym);
  for (int b = 0; b < nb; ++b) {
    int a = s->sublat_to_asym[b];
    if (!occ_to_sp[a].empty()) continue;
    for (int o = 0; o < s->t->n_occ[b]; ++o) occ_to_sp[a].push_back(s->occ_to_species[b * mo + o]);
  }
  std::vector<int> cand_asym, cand_species, occ_index((size_t)n_asym * nsp, -1);
  std::vector<std::vector<int>> cand_of(n_asym, std::vector<int>(nsp, -1));
  for (int a = 0; a < n_asym; ++a) {
    for (size_t o = 0; o < occ_to_sp[a].size(); ++o) occ_index[a * nsp + occ_to_sp[a][o]] = (int)o;
    if (occ_to_sp[a].size() < 2) continue;
    for (size_t o = 0; o < occ_to_sp[a].size(); ++o) {
      cand_of[a][occ_to_sp[a][o]] = (int)cand_asym.size();
      cand_asym.push_back(a);
      cand_species.push_back(occ_to_sp[a][o]);
    }
  }
  const int nc = (int)cand_asym.size();
  std::vector<int> swap_a, swap_b;
  if (mode == 0) {  // make_semigrand_canonical_swaps
    for (int a = 0; a < nc; ++a)
      for (int b = 0; b < nc; ++b)
        if (cand_asym[a] == cand_asym[b] && cand_species[a] != cand_species[b]) {
          swap_a.push_back(a);
          swap_b.push_back(b);
        }
  } else {  // make_canonical_swaps
    for (int a = 0; a < nc; ++a)
      for (int b = a + 1; b < nc; ++b)
        if (cand_species[a] != cand_species[b] && cand_of[cand_asym[a]][cand_species[b]] >= 0 &&
            cand_of[cand_asym[b]][cand_species[a]] >= 0) {
          swap_a.push_back(a);
          swap_b.push_back(b);
        }
  }
  if (swap_a.empty() || swap_a.size() > 64) {
    cmx_set_error("cmx_metropolis_sequential: no (or more than 64) swap types");
    return CMX_ERR_UNSUPPORTED;
  }
  std::vector<int32_t> occ((size_t)n_cells * nb);
  int rc = cmx_state_download_occ(s, replica, occ.data());
  if (rc) return rc;
  std::vector<long long> mol_l, mol_loc, cand_size(nc, 0);
  std::vector<int> mol_species;
  for (int64_t l = 0; l < n_cells * nb; ++l) {
    int a = s->sublat_to_asym[l / n_cells];
    if (occ_to_sp[a].size() < 2) continue;
    mol_l.push_back(l);
    int sp = occ_to_sp[a][occ[l]];
    mol_species.push_back(sp);
    mol_loc.push_back(cand_size[cand_of[a][sp]]++);
  }
  const long long cap = (long long)mol_l.size();
  std::vector<long long> loc((size_t)nc * std::max<long long>(cap, 1), -1);
  {
    std::vector<long long> fill(nc, 0);
    for (long long m = 0; m < cap; ++m) {
      int a = s->sublat_to_asym[mol_l[m] / n_cells];
      int ci = cand_of[a][mol_species[m]];
      loc[(size_t)ci * cap + fill[ci]++] = m;
    }
  }
  std::vector<void *> allocs;
  SeqArgs a;
  a.T = T;
  a.g = s->g;
  a.occ = s->d_occ + (size_t)replica * s->g.rep_stride;
  a.mode = mode;
  a.n_steps = n_steps;
  a.seed = seed;
  a.beta = 1.0 / (CMX_KB * s->temperature[replica]);
  a.exch = s->d_exch + (size_t)replica * nb * mo * mo;
  a.n_eci = s->n_eci;
  a.eci_idx = s->d_eci_idx;
  a.eci_val = s->d_eci_val;
  a.n_cand = nc;
  a.n_swap = (int)swap_a.size();
  a.n_species = nsp;
  a.cap = cap;
  a.log_cap = std::min<int64_t>(log_cap, n_steps);
  int *d_ca, *d_cs, *d_sa, *d_sb, *d_oi, *d_s2a, *d_ms, *d_status;
  long long *d_csz, *d_loc, *d_ml, *d_mloc, *d_nacc;
  unsigned long long *d_hash;
  cmx_step_record *d_log = nullptr;
#define DC(v, p)                                        \
  if ((rc = dev_copy(v, &p, allocs))) {                 \
    for (void *q : allocs) cudaFree(q);                 \
    return rc;                                          \
  }
  DC(cand_asym, d_ca);
  DC(cand_species, d_cs);
  DC(swap_a, d_sa);
  DC(swap_b, d_sb);
  DC(occ_index, d_oi);
  std::vector<int> s2a(s->sublat_to_asym.begin(), s->sublat_to_asym.end());
  DC(s2a, d_s2a);
  DC(mol_species, d_ms);
  DC(cand_size, d_csz);
  DC(loc, d_loc);
  DC(mol_l, d_ml);
  DC(mol_loc, d_mloc);
  std::vector<int> st(1, -1);
  DC(st, d_status);
  std::vector<long long> z1(1, 0);
  DC(z1, d_nacc);
  std::vector<unsigned long long> z2(1, 0);
  DC(z2, d_hash);
  std::vector<long long> z3(1 + CMX_TIE_CAP, 0);
  long long *d_ties;
  DC(z3, d_ties);
#undef DC
  if (a.log_cap > 0) {
    cudaError_t e = cudaMalloc((void **)&d_log, sizeof(cmx_step_record) * a.log_cap);
    if (e != cudaSuccess) {
      for (void *q : allocs) cudaFree(q);
      return cmx_cuda_fail(e, "cudaMalloc(log)");
    }
    allocs.push_back(d_log);
  }
  a.cand_asym = d_ca;
  a.cand_species = d_cs;
  a.swap_a = d_sa;
  a.swap_b = d_sb;
  a.occ_index = d_oi;
  a.sublat_to_asym = d_s2a;
  a.mol_species = d_ms;
  a.cand_size = d_csz;
  a.loc = d_loc;
  a.mol_l = d_ml;
  a.mol_loc = d_mloc;
  a.log = d_log;
  a.n_accept = d_nacc;
  a.hash = d_hash;
  a.status = d_status;
  a.ties = d_ties;
  k_metropolis_sequential<<<1, 32, 0, s->stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  int status = -1;
  long long nacc = 0;
  unsigned long long hh = 0;
  if (e == cudaSuccess) {
    cudaMemcpy(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost);
    cudaMemcpy(&nacc, d_nacc, sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(&hh, d_hash, sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(s->seq_ties, d_ties, sizeof(long long) * (1 + CMX_TIE_CAP), cudaMemcpyDeviceToHost);
    if (a.log_cap > 0 && log)
      cudaMemcpy(log, d_log, sizeof(cmx_step_record) * a.log_cap, cudaMemcpyDeviceToHost);
  }
  for (void *q : allocs) cudaFree(q);
  if (e != cudaSuccess) return cmx_cuda_fail(e, "k_metropolis_sequential");
  if (status != 0) {
    cmx_set_error("cmx_metropolis_sequential: proposal failed (empty candidate lists)");
    return CMX_ERR_STATE;
  }
  if (n_accept) *n_accept = nacc;
  if (hash) *hash = hh;
  return CMX_OK;
}

// Tie report of the last cmx_metropolis_sequential call on this state: the number of steps
// whose uniform draw was within one ulp of exp(-beta dE) -- the only steps at which the
// device (CUDA exp) and the reference (glibc exp) could decide differently -- and the first
// `cap` (<= 16) of their step indices.
extern "C" int cmx_metropolis_sequential_ties(const cmx_state *s, int64_t *n_near_ties, int64_t *steps, int32_t cap) {
  if (!s || !n_near_ties) return invalid("cmx_metropolis_sequential_ties: null argument");
  *n_near_ties = s->seq_ties[0];
  for (int q = 0; q < cap && q < CMX_TIE_CAP && steps; ++q) steps[q] = (q < s->seq_ties[0]) ? s->seq_ties[1 + q] : -1;
  return CMX_OK;
}
