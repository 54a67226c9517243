/*
 * cmx_b200.h -- thin C ABI of the B200-native clexmonte hot path.
 *
 * This is the boundary a `BaseMonteCalculator` subclass (the reference's
 * plugin interface, include/casm/clexmonte/monte_calculator/BaseMonteCalculator.hh:47-264,
 * obtained through `extern "C" make_<Name>()`, e.g.
 * src/casm/clexmonte/monte_calculator/SemiGrandCanonicalCalculator.cc:517-523)
 * binds to: opaque handles, plain pointers and sizes, `int` status codes
 * (0 = ok), `cmx_last_error()` for the message.  No C++ or torch types.
 * INTEGRATION.md shows the reference-side subclass that calls it.
 *
 * All pointer arguments are HOST pointers unless the name starts with `d_`.
 * Every entry point is implemented by hand-written sm_100a CUDA kernels in
 * casmcode_clexmonte_b200/csrc/; there is no CPU implementation behind it --
 * without a CUDA device every compute call fails with CMX_ERR_CUDA.
 *
 * Conventions:
 *   linear site index  l = b * n_cells + cell  (sublattice-major, as the reference's, SURVEY.md
 *                      Appendix B), cell = i + N0 * (j + N1 * k) over the unit cells of the
 *                      supercell box.  The reference's unit-cell ORDER (UnitCellIndexConverter,
 *                      [EXT] libcasm-crystallography) is not reproduced: a caller that needs
 *                      its own numbering binds through cmx_state_cell_index /
 *                      cmx_state_set_site_order.
 *   occupation value   occupant index on the sublattice's allowed list
 */
#ifndef CMX_B200_H
#define CMX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMX_OK 0
#define CMX_ERR_INVALID 1  /* bad argument / inconsistent tables */
#define CMX_ERR_CUDA 2     /* CUDA runtime error or no device */
#define CMX_ERR_UNSUPPORTED 3
#define CMX_ERR_STATE 4    /* call sequence error (e.g. no ECI bound) */

typedef struct cmx_tables cmx_tables; /* one basis set (one Clexulator) */
typedef struct cmx_state cmx_state;   /* device-resident supercell(s)   */

/* Message of the last error raised on the calling thread. */
const char *cmx_last_error(void);
/* Library/ABI version: major*10000 + minor*100 + patch. */
int cmx_version(void);
/* Number of visible CUDA devices (0 when there is none); never fails. */
int cmx_device_count(void);

/* -------------------------------------------------------------------------
 * Tables: the flat export of one CASM-generated Clexulator.
 * Replaces the runtime-compiled object behind
 *   clexulator::BaseClexulator  (generated source, e.g.
 *   tests/unit/clexmonte/data/FCC_binary_vacancy/basis_sets/bset.default/
 *   FCC_binary_vacancy_Clexulator_default.cc:306-439 for the data,
 *   :780-1062 for the expressions).
 * Layout documented in casmcode_clexmonte_b200/clexulator_tables.py.
 * ------------------------------------------------------------------------- */
typedef struct cmx_table_desc {
  int32_t n_sublat;       /* sublattices in the prim                        */
  int32_t max_occ;        /* max allowed occupants on any sublattice        */
  int32_t n_func;         /* max site basis functions per sublattice        */
  int32_t corr_size;      /* number of correlation (basis) functions        */
  int32_t n_point_corr;   /* point-corr positions (global: #nlist sublat;   */
                          /*   local clexulators: nlist_size)               */
  int32_t nlist_len;      /* neighbor-list sites                            */
  int32_t n_nlist_sublat; /* sublattices on the neighbor list               */
  int32_t n_factors, n_terms, n_elems, n_groups;
  const int32_t *nlist_sublat;  /* [n_nlist_sublat] sorted                   */
  const int32_t *n_occ;         /* [n_sublat] allowed occupants (1 = fixed)  */
  const double *phi;            /* [n_sublat][n_func][max_occ]               */
  const int32_t *nbr;           /* [nlist_len][4] = di, dj, dk, b            */
  const int32_t *factor_f;      /* [n_factors] site function index           */
  const int32_t *factor_n;      /* [n_factors] neighbor index                */
  const double *term_coef;      /* [n_terms]                                 */
  const int32_t *term_fbeg;     /* [n_terms+1]                               */
  const int32_t *elem_tbeg;     /* [n_elems+1]                               */
  const int32_t *group_ebeg;    /* [n_groups+1]                              */
  const int32_t *group_dphi;    /* [n_groups] f' of the delta factor or -1   */
  const int32_t *group_has_sum; /* [n_groups]                                */
  const double *group_div;      /* [n_groups] divisor, 0 = none              */
  const int32_t *global_gbeg;   /* [corr_size+1]                             */
  const int32_t *point_gbeg;    /* [n_point_corr*corr_size+1]                */
  const int32_t *delta_gbeg;    /* [n_point_corr*corr_size+1]                */
} cmx_table_desc;

int cmx_tables_create(const cmx_table_desc *desc, int device, cmx_tables **out);
/* The same from the flat file `ClexulatorTables.save_flat(path)` writes
 * (casmcode_clexmonte_b200/clexulator_tables.py: the export of one generated Clexulator
 * source): what a C++ plugin loads in its _reset(). */
int cmx_tables_create_from_file(const char *path, int device, cmx_tables **out);
void cmx_tables_destroy(cmx_tables *t);

/* -------------------------------------------------------------------------
 * State: `n_replicas` independent periodic supercells of N0 x N1 x N2 unit
 * cells (transformation matrix diag(N0,N1,N2)), occupation stored as int8 on
 * the device.  Stands in for the (Configuration, SuperNeighborList,
 * ClusterExpansion, OccLocation) bundle the reference builds in
 * StateData (src/casm/clexmonte/monte_calculator/StateData.cc:9-79) -- no
 * neighbor list is materialised; neighbors are (di,dj,dk,b) arithmetic.
 *
 * `halo` > 0 creates a slab for domain decomposition along k: the local box
 * holds N2 layers plus `halo` ghost layers on each side, k is NOT wrapped and
 * the ghost layers are filled by cmx_state_halo_* (BASELINE config 3).
 * ------------------------------------------------------------------------- */
int cmx_state_create(const cmx_tables *t, int32_t N0, int32_t N1, int32_t N2,
                     int32_t n_replicas, int32_t halo, cmx_state **out);
/* Same with options.  The device layout of a row is private to the library (every
 * transfer converts): single-sublattice states with <= 3 occupants whose N0 is a power of
 * two in [16, 512] store "x4-interleaved" rows (word w of a row holds the sites w, w+Q,
 * w+2Q, w+3Q, Q = N0/4), the layout of the streaming pair-LUT sweep (k_sweep_stream16).
 * CMX_STATE_LINEAR_ROWS keeps site i at byte i: such states sweep with the block kernel
 * (k_sweep_pair16) -- a cross-check of the streaming kernel, not a faster path. */
#define CMX_STATE_LINEAR_ROWS 1u
int cmx_state_create_opts(const cmx_tables *t, int32_t N0, int32_t N1, int32_t N2,
                          int32_t n_replicas, int32_t halo, uint32_t options, cmx_state **out);
/* General supercells: T[9] (row major) is the transformation matrix to the supercell, columns
 * = supercell lattice vectors in prim coordinates (StateData::transformation_matrix_to_super,
 * e.g. 10 * fcc_conventional, tests/unit/teststructures.hh:12-16).  The library brings it to
 * Hermite normal form: the unit cells are the box 0 <= i < box[0], j < box[1], k < box[2]
 * of prim-lattice coordinates, periodic with the skewed images (box[3..5] = s10, s20, s21:
 * leaving the box along i shifts j and k, along j shifts k).  cell = i + box[0] (j + box[1] k),
 * l = b * n_cells + cell.  The reference numbers unit cells differently (its
 * UnitCellIndexConverter walks the Smith normal form); cmx_state_cell_index converts unit-cell
 * coordinates, and cmx_state_set_site_order makes upload / download speak the caller's order.
 * Skewed boxes run the faithful evaluators, the reference-order mode, KMC and the generic
 * checkerboard sweeps; the pair-LUT kernels and pair exchanges need diag(N0, N1, N2). */
int cmx_state_create_general(const cmx_tables *t, const int32_t *T, int32_t n_replicas, uint32_t options,
                             cmx_state **out);
int cmx_state_box(const cmx_state *s, int32_t *box /*[6]*/);
int cmx_supercell_box(const int32_t *T, int32_t *box /*[6]*/); /* host only: the box of T */
int cmx_state_cell_index(const cmx_state *s, int64_t n, const int32_t *ijk /*[n][3]*/, int64_t *cell /*[n]*/);
int cmx_state_set_site_order(cmx_state *s, const int64_t *order /*[n_sites] caller l -> library l, or NULL*/);
void cmx_state_destroy(cmx_state *s);

/* occupation in the reference's layout (int32, Eigen::VectorXi order
 * l = b*n_cells + cell); converted to/from int8 on the device. */
int cmx_state_upload_occ(cmx_state *s, int32_t replica, const int32_t *occ);
int cmx_state_download_occ(const cmx_state *s, int32_t replica, int32_t *occ);
/* same, int8 host buffers (1 byte/site over PCIe) */
int cmx_state_upload_occ_i8(cmx_state *s, int32_t replica, const int8_t *occ);
int cmx_state_download_occ_i8(const cmx_state *s, int32_t replica, int8_t *occ);
/* i.i.d. uniform occupation over each sublattice's allowed occupants from the
 * counter-based generator, keyed by (seed, replica, site). */
int cmx_state_randomize(cmx_state *s, uint64_t seed);
/* slab decomposition: global k index of this slab's first owned layer (enters
 * the RNG counters so that results do not depend on the decomposition). */
int cmx_state_set_k_offset(cmx_state *s, int32_t k_offset);
/* the CUDA stream (cudaStream_t) every kernel of this state is launched on,
 * so that callers can bracket the work with their own events. */
int cmx_state_stream(cmx_state *s, void **stream);
/* raw device pointer to replica 0's int8 occupation including ghost layers,
 * and its size in bytes (for NCCL halo plumbing by the host). */
int cmx_state_device_ptr(cmx_state *s, void **d_ptr, size_t *n_bytes);

/* Slab decomposition over NVLink peer memory (one process per GPU, one box).
 * cmx_state_ipc_export writes an opaque handle (CUDA IPC) of a slab state;
 * the host exchanges the handles of ring neighbours (e.g. an all-gather) and
 * calls cmx_state_ipc_attach(handle of the lower neighbour, of the upper
 * neighbour; NULL = the neighbour is this state itself, i.e. one slab).  From
 * then on the pair-LUT sweep kernel stores every boundary row it changes
 * straight into the neighbour's ghost layer and the ranks synchronise through
 * epoch flags in peer memory (release/acquire at system scope): the halo
 * exchange is fused into the sweep kernel and cmx_sgc_sweep_kgroup needs no
 * separate exchange between k-colour groups.  Every rank must issue the same
 * sequence of sweep calls. */
#define CMX_IPC_HANDLE_BYTES 128
int cmx_state_ipc_export(cmx_state *s, void *handle);
int cmx_state_ipc_attach(cmx_state *s, const void *handle_dn, const void *handle_up);
/* 1 when sweeps of this state push their halos themselves */
int cmx_state_p2p_active(const cmx_state *s, int32_t *active);

/* Bind the cluster expansion coefficients: sparse (index, value) pairs as in
 * the reference's SparseCoefficients; ClexData parsed at
 * src/casm/clexmonte/system/io/json/System_json_io.cc:491-539. */
int cmx_state_set_eci(cmx_state *s, int32_t n, const uint32_t *index,
                      const double *value);

/* Thermodynamic conditions of one replica.
 *   temperature [K]; beta = 1/(KB*T), KB = 8.6173303e-05 eV/K
 *     (include/casm/clexmonte/methods/occupation_metropolis.hh:86)
 *   exch[b][occ_i][occ_f] (n_sublat*max_occ*max_occ doubles, may be NULL = 0):
 *     the semi-grand term  mu_x . (R^T dN)  of
 *     SemiGrandCanonicalPotential::occ_delta_per_supercell
 *     (SemiGrandCanonicalCalculator.cc:201-212) tabulated per occupant change;
 *     dE_potential = dE_clex - exch[b][occ_i][occ_f]. */
int cmx_state_set_conditions(cmx_state *s, int32_t replica, double temperature,
                             const double *exch);

/* -------------------------------------------------------------------------
 * Potential / correlations (reference rows a1-a8 of SURVEY.md section 8a).
 * "Faithful" evaluation: the reference's operation order, no FMA contraction,
 * so results are bit-identical to the generated C++ on x86-64.
 * ------------------------------------------------------------------------- */

/* Clexulator::calc_delta_point_corr for `n` independent single-site changes
 * (…default.cc:555-580).  out[n][corr_size]. */
int cmx_delta_corr(const cmx_state *s, int32_t replica, int64_t n,
                   const int64_t *l, const int32_t *new_occ, double *out);
/* Clexulator::calc_point_corr (:500-553).  out[n][corr_size]. */
int cmx_point_corr(const cmx_state *s, int32_t replica, int64_t n,
                   const int64_t *l, double *out);
/* Clexulator::calc_global_corr_contribution for `n` unit cells (:446-470);
 * for local clexulators this is LocalCorrelations::local.  out[n][corr_size] */
int cmx_cell_corr(const cmx_state *s, int32_t replica, int64_t n,
                  const int64_t *cell, double *out);
/* ClusterExpansion::occ_delta_value for `n` independent events of
 * `sites_per_event` sites each (1 = semi-grand flip, 2 = canonical swap / KMC
 * hop): sites are applied sequentially and restored, restricted to the bound
 * ECI indices (call sites SemiGrandCanonicalCalculator.cc:198-199,
 * CanonicalCalculator.cc:139).  With `potential` != 0 the replica's exch term
 * is subtracted (occ_delta_per_supercell, :186-213).  dE[n]. */
int cmx_delta_e(const cmx_state *s, int32_t replica, int64_t n,
                int32_t sites_per_event, const int64_t *l,
                const int32_t *new_occ, int32_t potential, double *dE);
/* Correlations::per_supercell(): sum over all unit cells of the global
 * contribution (call sites sampling_functions.cc:131-133).  out[corr_size]. */
int cmx_global_corr(const cmx_state *s, int32_t replica, double *out);
/* ClusterExpansion::per_supercell() = sum_i value_i * corr[index_i]. */
int cmx_energy(const cmx_state *s, int32_t replica, double *E);
/* occupant counts counts[b][occ] (n_sublat*max_occ int64) -- the input of
 * CompositionCalculator::mean_num_each_component (sampling_functions.cc:37-54). */
int cmx_composition(const cmx_state *s, int32_t replica, int64_t *counts);

/* -------------------------------------------------------------------------
 * Metropolis drivers (reference rows a9/a10).
 * ------------------------------------------------------------------------- */
typedef struct cmx_counters {
  int64_t n_attempt;   /* attempted steps                                    */
  int64_t n_accept;    /* accepted steps                                     */
  double dE_sum;       /* sum of accepted delta potential energies           */
                       /*   (only with CMX_SWEEP_DE_SUM, else 0)             */
  int64_t reserved;
} cmx_counters;

/* Semi-grand canonical sublattice-checkerboard sweeps over ALL replicas:
 * one sweep attempts one single-site change at every mutable site, colour by
 * colour (non-interacting site sets updated simultaneously), counter-based
 * RNG keyed by (seed, replica, sweep, site).  Replaces the loop
 * methods/occupation_metropolis.hh:92-120 with the semi-grand proposal
 * (SemiGrandCanonicalCalculator.cc:104-120).  counters[n_replicas] (may be
 * NULL) receive the totals of this call.  `first_sweep` is the index of the
 * first sweep (the RNG counter), so consecutive calls continue one stream. */
int cmx_sgc_sweep(cmx_state *s, int64_t n_sweeps, uint64_t seed,
                  int64_t first_sweep, cmx_counters *counters);
/* The unit between two halo exchanges of a slab-decomposed run: one pass over
 * the colours whose k-colour is `kgroup` (0 .. S_k-1; -1 = all).  ASYNCHRONOUS:
 * the kernels are enqueued on the state's stream (cmx_state_stream) and the
 * call returns; counters accumulate until cmx_counters_reset. */
int cmx_sgc_sweep_kgroup(cmx_state *s, uint64_t seed, int64_t sweep,
                         int32_t kgroup);
int cmx_counters_reset(cmx_state *s);
/* Sweep options (default 0).
 *   CMX_SWEEP_DE_SUM        also accumulate cmx_counters.dE_sum, the sum of the
 *                           accepted delta potential energies (a diagnostic:
 *                           it must equal E(after) - E(before)).  Off by
 *                           default -- the reference loop counts acceptances
 *                           only (methods/occupation_metropolis.hh:109-116)
 *                           and samples energies from scratch -- and dE_sum
 *                           then reads 0.
 *   CMX_SWEEP_FORCE_GENERIC use the generic term-list evaluator even where the
 *                           pair-LUT kernel applies (same random bits, same
 *                           decisions: a cross-check of the fast path) */
#define CMX_SWEEP_DE_SUM 1u
#define CMX_SWEEP_FORCE_GENERIC 2u
/* pair-LUT kernels (the trajectory does not depend on the kernel).  States with
 * x4-interleaved rows (cmx_state_create_opts) run a whole cmx_sgc_sweep call -- all its
 * sweeps, all colours -- as ONE cooperative launch:
 *   default           colour passes with a grid barrier after each (k_sweep_pass16);
 *                     consecutive passes walk the layers in opposite directions and a
 *                     lattice slightly larger than L2 is pinned there as far as the device
 *                     allows, so a sweep reads most of it from L2
 *   CMX_SWEEP_STREAM  the barrier-free streaming kernel (k_sweep_stream16): the units of the
 *                     call in wavefront order, per-layer completion counters instead of
 *                     barriers; every lattice byte crosses HBM once per sweep and direction
 *                     whatever the lattice size.
 * States with linear rows launch the block kernel once per colour pass (k_sweep_pair16). */
#define CMX_SWEEP_STREAM 8u
/* generic (term-list) evaluator variants: models with wide orbit sets (>= 96 merged
 * terms per site, e.g. ZrO with triplets and quadruplets) evaluate one site per WARP
 * (k_sweep_generic_warp, k_canonical_pairs_warp: neighborhood staged once, terms dealt
 * to the lanes); THREAD_GENERIC forces one site per thread (same random bits and
 * decisions; dE differs in the last bits through the summation order). */
#define CMX_SWEEP_THREAD_GENERIC 16u
/* Point + pair bases whose active neighbors form TWO symmetry classes (the reference's dense
 * FCC formation_energy_eci.json: points, 1NN and 2NN pairs -- SURVEY 8d's 19-site
 * neighbourhood) run the colour-pass kernel with a two-class count table on x4-interleaved
 * rows (evaluator "pair_lut2"); CMX_SWEEP_PAIR_SUM selects the per-neighbor-table kernel
 * (k_sweep_pairsum, evaluator "pair_sum") instead: same random bits, same decisions. */
#define CMX_SWEEP_PAIR_SUM 32u
int cmx_state_set_sweep_flags(cmx_state *s, uint32_t flags);
/* Asynchronous forms for pipelines of independent states (each state owns a stream): the
 * upload of one job overlaps the sweeps of another and the download of a third.  Host
 * buffers must be pinned for the copies to be asynchronous.  cmx_state_synchronize waits
 * for the state's stream and reports an occupant index out of range found by an
 * asynchronous upload; cmx_counters_read (synchronising) returns the sweep counters. */
int cmx_state_upload_occ_i8_async(cmx_state *s, int32_t replica, const int8_t *occ);
int cmx_state_download_occ_i8_async(cmx_state *s, int32_t replica, int8_t *occ);
int cmx_sgc_sweep_async(cmx_state *s, int64_t n_sweeps, uint64_t seed, int64_t first_sweep);
/* more sweeps of the same accounting period: like cmx_sgc_sweep_async, but the counters
 * keep accumulating (no reset) */
int cmx_sgc_sweep_continue(cmx_state *s, int64_t n_sweeps, uint64_t seed, int64_t first_sweep);
int cmx_state_synchronize(cmx_state *s);
/* Slab states attached over peer memory (cmx_state_ipc_attach): n_sweeps whole sweeps in
 * one cooperative launch of the streaming kernel; boundary rows are stored into the ring
 * neighbours' ghost layers and counted on their layer counters (no collective, no barrier).
 * Asynchronous (enqueued on the state's stream).  CMX_ERR_UNSUPPORTED when the state is
 * not such a slab: use cmx_sgc_sweep_kgroup (+ the host-side halo exchange) then. */
int cmx_sgc_sweep_slab(cmx_state *s, int64_t n_sweeps, uint64_t seed, int64_t first_sweep);
/* synchronises the state's stream; counters[n_replicas] */
int cmx_counters_read(cmx_state *s, cmx_counters *counters);
/* Name of the evaluator the sweep uses for the bound ECI ("pair_lut",
 * "generic"); algorithmic work per attempted step for the roofline
 * bookkeeping: neighbor bytes read and FP64 flops, counted from the tables. */
int cmx_sweep_info(const cmx_state *s, char *name, size_t name_cap,
                   double *bytes_per_step, double *flops_per_step,
                   int32_t *n_colours, int32_t *colour_strides /*[3]*/,
                   int32_t *range_k);

/* -------------------------------------------------------------------------
 * Canonical ensemble: parallel pair exchanges.
 * Replaces the loop of methods/occupation_metropolis.hh:92-120 with the
 * canonical proposal (CanonicalCalculator.cc:78-88: two unlike-species sites
 * swap their species) and CanonicalPotential::occ_delta_per_supercell
 * (CanonicalCalculator.cc:137-140: the two-site, sequentially evaluated
 * delta E of the formation energy).
 *
 * A swap type pairs site (b_a, cell) with site (b_b, cell + t) for every unit
 * cell.  Per type the cells are coloured so that no site of one pair is within
 * the interaction range of (or identical to) a site of another pair of the
 * same colour; the colours are then updated one after the other, all pairs of
 * a colour simultaneously.  t may be a nearest-neighbour vector (Kawasaki
 * exchange) or long (global mixing, like the reference's any-two-sites swaps).
 * Composition is conserved exactly.  Like-species pairs are skipped and not
 * counted as attempts (the reference never proposes them).
 * ------------------------------------------------------------------------- */
typedef struct cmx_swap_type {
  int32_t b_a, b_b; /* sublattices of the two sites                         */
  int32_t t[3];     /* unit-cell translation from site a to site b          */
} cmx_swap_type;
/* Needs the ECI (cmx_state_set_eci); cross-sublattice swaps need the species
 * of the occupants (cmx_state_set_occupants).  Fails with CMX_ERR_INVALID when
 * a type admits no valid colouring of this supercell. */
int cmx_canonical_set_swaps(cmx_state *s, int32_t n, const cmx_swap_type *swaps);
/* The default swap table of this state, the parallel counterpart of the reference's
 * canonical swaps (make_canonical_swaps [EXT], built at system/System.cc:55-58): for every
 * pair of mutable sublattices on one asymmetric unit that share a species, the `n_shell`
 * shortest translations of the prim neighbor list (half of them when the sublattices are
 * equal: +t and -t give the same pairs) and, with `long_range`, one translation across the
 * box.  Needs cmx_state_set_occupants.  *n receives the number of types; cap = 0 only asks
 * for it. */
int cmx_canonical_default_swaps(const cmx_state *s, int32_t n_shell, int32_t long_range, int32_t cap,
                                cmx_swap_type *swaps, int32_t *n);
/* One sweep = every swap type, every colour, once: n_swap_types * n_cells pairs
 * visited.  counters[n_replicas]: n_attempt = unlike-species pairs evaluated,
 * n_accept, dE_sum (always accumulated here). */
int cmx_canonical_sweep(cmx_state *s, int64_t n_sweeps, uint64_t seed,
                        int64_t first_sweep, cmx_counters *counters);
/* colour strides chosen for swap type `i` and its number of colours */
int cmx_canonical_info(const cmx_state *s, int32_t i, int32_t *strides /*[3]*/,
                       int32_t *n_colours);

/* Kernel launches one full sweep takes with the current evaluator (block pair-LUT kernel:
 * 4 launches for 8 colours; generic: one per colour); 0 = the kernels on x4-interleaved rows,
 * which run a whole cmx_sgc_sweep call as one launch. */
int cmx_sweep_launches(const cmx_state *s, int32_t *per_sweep);
/* Size of the ECI-folded term lists the term-list evaluators walk, averaged over the mutable
 * point positions: merged terms per attempted step and distinct neighbor sites they read
 * (for the shared-memory / FP64 roofline bookkeeping of wide orbit sets). */
int cmx_sweep_term_counts(const cmx_state *s, double *terms_per_step, double *neighbors_per_step);
/* Schedule of the streaming kernel on this state: *stream = 1 when it is the evaluator
 * (CMX_SWEEP_STREAM on a state with x4-interleaved rows),
 * *blocks = co-resident blocks per replica, *group_rowsteps = row-steps a warp takes at a
 * time, *gap_units = (layer, row colour) units the host keeps between dependent units. */
int cmx_sweep_stream_info(cmx_state *s, int32_t *stream, int32_t *blocks, int32_t *group_rowsteps,
                          int32_t *gap_units);
/* Debug / parity entry: delta potential energy of the proposal (l, new_occ) on the current
 * occupation of `replica` as the SWEEP's evaluator computes it -- the pair-LUT table entry
 * (dE - exchange term) or the folded term lists of the generic evaluator (thread or warp
 * variant, as the sweep would choose) -- to compare with cmx_delta_e (the faithful
 * evaluator) proposal by proposal.  out[n]. */
int cmx_sweep_debug_delta_e(cmx_state *s, int32_t replica, int64_t n, const int64_t *l,
                            const int32_t *new_occ, double *out);

/* Occupant bookkeeping needed by the reference-order mode:
 * sublat_to_asym[n_sublat], occ_to_species[n_sublat][max_occ] (-1 padded). */
int cmx_state_set_occupants(cmx_state *s, const int32_t *sublat_to_asym,
                            const int32_t *occ_to_species, int32_t n_species);

typedef struct cmx_step_record {
  int64_t l0, l1;   /* l1 = -1 for semi-grand */
  int32_t new0, new1;
  int32_t accepted, pad;
  double dE;
} cmx_step_record;

/* Reference-order (sequential) Metropolis on the device, one replica:
 * std::mt19937_64(seed) stream, libstdc++ uniform_int/uniform_real draws,
 * OccLocation candidate lists with swap-and-pop, proposals as
 * propose_semigrand_canonical_event (mode 0) / propose_canonical_event
 * (mode 1), acceptance as metropolis_acceptance -- the exact sequence of
 * methods/occupation_metropolis.hh:92-120.  Reproduces the reference
 * trajectory bit for bit; `hash` is FNV-1a over (l0*2+accepted) per step.
 * log[log_cap] (may be NULL) receives the first steps. */
int cmx_metropolis_sequential(cmx_state *s, int32_t replica, int32_t mode,
                              int64_t n_steps, uint64_t seed,
                              cmx_step_record *log, int64_t log_cap,
                              int64_t *n_accept, uint64_t *hash);
/* Tie report of the last cmx_metropolis_sequential call on this state.  The one operation of
 * that mode not under our control is exp(): CUDA's and glibc's results may differ in the last
 * place, which flips an accept/reject decision only if the uniform draw lies within one ulp of
 * exp(-beta dE).  Such steps are counted (*n_near_ties) and the first `cap` (<= 16) step
 * indices returned (steps[cap], -1 padded): a trajectory that leaves the reference's can be
 * traced to them, and a report of 0 near ties means none could have. */
int cmx_metropolis_sequential_ties(const cmx_state *s, int64_t *n_near_ties, int64_t *steps, int32_t cap);

/* -------------------------------------------------------------------------
 * Device-side samplers.  Replaces, for every replica
 * at once and without moving the occupation to the host, the state sampling
 * functions of src/casm/clexmonte/monte_calculator/sampling_functions.cc:
 *   "clex.formation_energy" (:141-155)  ClusterExpansion::per_unitcell
 *   "potential_energy"      (:274-288)  potential().per_unitcell()
 *                                       = (E - n_cells mu.x) / n_cells
 *                                       (SemiGrandCanonicalCalculator.cc:171-179;
 *                                        canonical: mu = 0, CanonicalCalculator.cc:126-133)
 *   "mol_composition"       (:37-54)    mean_num_each_component
 *   "param_composition"     (:57-90)    x = R^T (n - origin)
 *   "corr"                  (:121-139)  Correlations::per_unitcell  (optional)
 * One series row per replica and sample:
 *   [formation_energy, potential_energy, mol_composition[n_species],
 *    param_composition[n_param], corr[corr_size] (with_corr)]
 * The analysis functions heat_capacity / mol_susc / param_susc / *_thermochem_susc
 * (monte_calculator/analysis_functions.cc:43-173) are variances and covariances of
 * these series (host layer: Sampler.analysis()).
 * Requires bound ECI and cmx_state_set_occupants.  origin[n_species],
 * Rt[n_param][n_species] as in the composition axes of the system.
 * ------------------------------------------------------------------------- */
typedef struct cmx_sampler cmx_sampler;
int cmx_sampler_create(cmx_state *s, int32_t capacity, int32_t n_param, const double *origin,
                       const double *Rt, int32_t with_corr, cmx_sampler **out);
int cmx_sampler_destroy(cmx_sampler *m);
/* param_chem_pot[n_param] of one replica (NULL: canonical potential, no mu.x term) */
int cmx_sampler_set_param_chem_pot(cmx_sampler *m, int32_t replica, const double *param_chem_pot);
int cmx_sampler_info(const cmx_sampler *m, int32_t *n_quantities, int32_t *n_species, int32_t *n_param,
                     int32_t *corr_size, int32_t *n_samples, int32_t *capacity);
/* forget the recorded samples */
int cmx_sampler_reset(cmx_sampler *m);
/* append one sample of every replica (asynchronous, on the state's stream) */
int cmx_sampler_sample(cmx_sampler *m);
/* rows [first, first+n) of one replica: out[n][n_quantities]; synchronises */
int cmx_sampler_read(cmx_sampler *m, int32_t replica, int32_t first, int32_t n, double *out);
/* occupation_metropolis_v2 (methods/occupation_metropolis.hh:72-123) with a
 * sampling period of `sweeps_per_sample` passes: n_samples x (sweeps, sample) as
 * one stream of launches, one synchronisation at the end.
 * ensemble 0 = semi-grand canonical sweeps, 1 = canonical pair exchanges.
 * counters[n_replicas] may be NULL. */
int cmx_sweep_run(cmx_state *s, cmx_sampler *m, int32_t ensemble, int64_t n_samples,
                  int64_t sweeps_per_sample, uint64_t seed, int64_t first_sweep,
                  cmx_counters *counters);
/* Moments of the sampled series of every replica over the samples [first, n_samples):
 * out[replica][M], M = 1 + Q + Q*Q, Q = 2 + n_species + n_param scalar quantities
 * q = (clex.formation_energy, potential_energy, mol_composition..., param_composition...):
 * { n, sum q_a, sum q_a q_b }.  Additive over disjoint samples / replicas: ranks that hold
 * different replicas of a (mu, T) grid combine them with ONE all-reduce, and the analysis
 * functions (heat_capacity, *_susc: analysis_functions.cc:43-173) follow from the sums.
 * out_on_device != 0: `out` is a device pointer (e.g. a torch tensor handed to NCCL) and the
 * call is asynchronous on the state's stream. */
int cmx_sampler_moments(cmx_sampler *m, int32_t first, double *out, int32_t out_on_device);

/* -------------------------------------------------------------------------
 * Kinetic Monte Carlo: event-state / event-rate evaluation (reference rows
 * a11/a12).  Replaces EventStateCalculator::calculate_event_state +
 * _default_event_state_calculation
 * (src/casm/clexmonte/monte_calculator/BaseMonteEventData.cc:87-156), batched
 * over (trajectory, unit cell, prim event) triples -- the unit of work of
 * lotto's update_impacted_event_rates (events/lotto/rejection_free.hpp:322-330)
 * -- and over many independent trajectories (= replicas of the state).
 * ------------------------------------------------------------------------- */
#define CMX_EVENT_MAX_SITES 4

/* PrimEventData (include/casm/clexmonte/events/event_data.hh:47-72): one event
 * of the origin unit cell.  site[q] = (b, i, j, k) as xtal::UnitCellCoord. */
typedef struct cmx_prim_event {
  int32_t n_sites;
  int32_t site[CMX_EVENT_MAX_SITES][4];
  int32_t occ_init[CMX_EVENT_MAX_SITES];
  int32_t occ_final[CMX_EVENT_MAX_SITES];
  int32_t event_type;       /* index into the event types below              */
  int32_t equivalent_index; /* which equivalent local clexulator to use      */
} cmx_prim_event;

/* One event type: its local multi-cluster-expansion (LocalMultiClexData with the
 * "kra" and "freq" coefficients, BaseMonteEventData.cc:36-61) over one local
 * clexulator per equivalent (system/io/json/System_json_io.cc:546-602). */
typedef struct cmx_event_type {
  int32_t n_equivalents;
  const cmx_tables *const *local_tables; /* [n_equivalents], same device     */
  int32_t n_kra;
  const uint32_t *kra_index;
  const double *kra_value;
  int32_t n_freq;
  const uint32_t *freq_index;
  const double *freq_value;
} cmx_event_type;

/* EventState (events/event_data.hh:19-30) without the two pointers. */
typedef struct cmx_event_state {
  int32_t is_allowed;
  int32_t is_normal;
  double dE_final;
  double Ekra;
  double dE_activated;
  double freq;
  double rate;
} cmx_event_state;

typedef struct cmx_kmc cmx_kmc;

/* The state's bound ECI are the "formation_energy" cluster expansion; the
 * temperature of each replica (cmx_state_set_conditions) gives beta. */
int cmx_kmc_create(cmx_state *s, int32_t n_event_types, const cmx_event_type *types,
                   int32_t n_prim_events, const cmx_prim_event *prim_events,
                   cmx_kmc **out);
void cmx_kmc_destroy(cmx_kmc *k);
/* Event states of `n` (replica, unit cell, prim event) triples.  out[n]. */
int cmx_kmc_event_states(cmx_kmc *k, int64_t n, const int32_t *replica,
                         const int64_t *unitcell, const int32_t *prim_event,
                         cmx_event_state *out);
/* Rates of EVERY event of every replica -- what the event selector computes
 * once at construction (lotto/rejection_free.hpp:231-234):
 * rates[replica][unitcell][prim_event] (host, may be NULL) and the per-replica
 * total rate total[n_replicas] (may be NULL).  The rates stay on the device
 * (d_rates) for the caller's selector. */
int cmx_kmc_all_rates(cmx_kmc *k, double *rates, double *total, void **d_rates);

/* -------------------------------------------------------------------------
 * Rejection-free KMC, whole steps on the device: every
 * replica of the state is an independent trajectory, one thread block each.
 * Replaces kinetic_monte_carlo_v2 (methods/kinetic_monte_carlo.hh:396-507) with
 * lotto::RejectionFreeEventSelector over the complete event list
 * (submodules/kmc-lotto/include/lotto/rejection_free.hpp:50-198; tree:
 * sum_tree_impl.hpp:190-240, event_rate_tree_impl.hpp:63-75,135-150; random numbers:
 * std::mt19937_64 + lotto::RandomGeneratorT, random.hpp:35-91): same tree shape, same
 * summation order, same draws -> the same event sequence as the reference selector
 * given the same rates and seed.  Event id = unitcell * n_prim_events + prim_event
 * (make_complete_event_id_list, events/CompleteEventList.cc:76-91).
 * ------------------------------------------------------------------------- */
typedef struct cmx_kmc_step {
  int64_t unitcell;
  int32_t prim_event, pad;
  double time_increment; /* -log(u) / total_rate                      */
  double total_rate;     /* total rate the event was selected from    */
} cmx_kmc_step;
/* Relative impact table (make_relative_impact_table, events/ImpactTable.cc:150-185):
 * entries[beg[j] .. beg[j+1]) = (prim event i, dx, dy, dz): after prim event j happened in
 * unit cell c the rate of event i in cell c + (dx,dy,dz) must be recomputed.
 * casmcode_clexmonte_b200.kmc.make_relative_impact_table builds it from the tables. */
int cmx_kmc_set_impact_table(cmx_kmc *k, int32_t n_entries, const int32_t *beg, const int32_t *entries);
/* (Re)start: all rates from the current occupation, the sum tree, one seed per
 * trajectory, time = 0. */
int cmx_kmc_run_begin(cmx_kmc *k, const uint64_t *seeds);
/* n_steps events of every trajectory; see csrc/cmx_kmc.cu for the arguments. */
int cmx_kmc_run(cmx_kmc *k, int64_t n_steps, cmx_kmc_step *log, int64_t log_cap, double *time,
                double *total_rate, int64_t *n_steps_done);
/* The selector's current leaves and roots, not re-evaluated:
 * rates[n_replicas][n_cells][n_prim], total[n_replicas] (either may be NULL). */
int cmx_kmc_current_rates(cmx_kmc *k, double *rates, double *total);

/* Test hook: replay `n` draws of the device-side restatement of
 * std::mt19937_64(seed) + libstdc++ distributions (kind 0 = raw 64-bit,
 * 1 = uniform_int_distribution<long>(0,int_max), 2 =
 * uniform_real_distribution<double>(0,real_max)) so the stream can be compared
 * draw by draw with the host C++ library. */
int cmx_rng_stream_test(uint64_t seed, int64_t n, const int64_t *int_max,
                        const double *real_max, const int32_t *kind,
                        int64_t *out_int, double *out_real, uint64_t *out_raw);

#ifdef __cplusplus
}
#endif
#endif /* CMX_B200_H */
