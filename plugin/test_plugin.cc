// Drives the plugin the way MonteCalculator does (monte_calculator/MonteCalculator.cc:142-173):
// dlopen, look up "make_" + name, reset(params, system), run(state, occ_location, run_manager).
//   test_plugin <libB200SemiGrandCanonicalCalculator.so> <tables.flat> [--no-gpu]
// --no-gpu: stop after the factory / interface checks (CPU CI).  With a GPU: the run must
// leave the occupation and the counters the C ABI gives for the same seed, the potential's
// delta must equal cmx_delta_e, and the sampled potential must close the energy balance.
// Prints "plugin ok ..." and exits 0, or a message and exits 1.
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstring>

#include "casm/clexmonte/monte_calculator/BaseMonteCalculator.hh"
#include "cmx_b200.h"

using namespace CASM;
using namespace CASM::clexmonte;

static int fail(const char *what) {
  std::fprintf(stderr, "test_plugin: FAILED: %s\n", what);
  return 1;
}

// the caller's unit-cell index of the cell at prim coordinates x (any periodic image): x and
// the cell's own coordinates differ by a supercell lattice vector T n, i.e. adj(T) (x - y) is
// a multiple of det(T) in every component
static Index find_unitl(monte::Conversions const &convert, Eigen::Matrix3l const &T, const long x[3]) {
  long A[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      const int r1 = (c + 1) % 3, r2 = (c + 2) % 3, c1 = (r + 1) % 3, c2 = (r + 2) % 3;
      A[r][c] = T(r1, c1) * T(r2, c2) - T(r1, c2) * T(r2, c1);
    }
  long det = 0;
  for (int c = 0; c < 3; ++c) det += T(0, c) * A[c][0];
  for (Index u = 0; u < convert.m_n_unitcells; ++u) {
    auto const y = convert.l_to_ijk(u);
    bool same = true;
    for (int r = 0; r < 3 && same; ++r) {
      const long f = A[r][0] * (x[0] - y[0]) + A[r][1] * (x[1] - y[1]) + A[r][2] * (x[2] - y[2]);
      same = f % det == 0;
    }
    if (same) return u;
  }
  return -1;
}

// A skewed supercell (general transformation matrix; the caller numbers its unit cells in its
// own order and at its own periodic images): the plugin must ask Conversions::l_to_ijk.
//  (1) physics pins the site mapping: a B atom and a vacancy in an otherwise pure-A crystal
//      interact exactly when they are first neighbours, with the pair energy a diag(N) box gives;
//  (2) the potential's delta equals the difference of two evaluations;
//  (3) a run leaves the occupation and the counters of the C ABI driven with the same site order.
static int general_supercell_case(BaseMonteCalculator &calc, std::shared_ptr<system_type> system, const char *tables_path) {
  const long Tg[3][3] = {{8, 0, 0}, {2, 8, 0}, {4, 2, 8}};  // columns = supercell lattice vectors; not symmetric
  state_type state;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) state.configuration.transformation_matrix_to_super(i, j) = Tg[i][j];
  const Index n_cells = 512;
  state.configuration.dof_values.occupation = Eigen::VectorXi(n_cells);
  state.conditions.scalar_values["temperature"] = 900.0;
  Eigen::VectorXd mu(2);
  mu[0] = 0.2;
  mu[1] = -0.1;
  state.conditions.vector_values["param_chem_pot"] = mu;
  monte::OccLocation occ_location;
  occ_location.m_mol_size = n_cells;
  if (!calc.validate_state(state).valid()) return fail("validate_state rejected a skewed supercell");
  Eigen::VectorXi &occ = state.configuration.dof_values.occupation;

  // (1) B at unit cell u0, Va at u0 + d
  for (Index l = 0; l < n_cells; ++l) occ[l] = 0;
  calc.set_state_and_potential(state, &occ_location);
  if (calc.state_data->n_unitcells != n_cells) return fail("general supercell: n_unitcells");
  // (copies: every set_state_and_potential makes a new StateData)
  const monte::Conversions convert = *calc.state_data->convert;
  const Eigen::Matrix3l T = calc.state_data->transformation_matrix_to_super;
  const Index u0 = 137;
  auto const c0 = convert.l_to_ijk(u0);
  auto potential_with_vacancy_at = [&](const long d[3], double &out) -> bool {
    const long x[3] = {c0[0] + d[0], c0[1] + d[1], c0[2] + d[2]};
    const Index u1 = find_unitl(convert, T, x);
    if (u1 < 0 || u1 == u0) return false;
    for (Index l = 0; l < n_cells; ++l) occ[l] = 0;
    occ[u0] = 1;
    occ[u1] = 2;
    calc.set_state_and_potential(state, &occ_location);
    out = calc.potential->per_supercell();
    return true;
  };
  const long far[3] = {3, 3, 3}, not_nn[3] = {1, 1, 0};
  const long nn[4][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, -1}, {1, -1, 0}};  // FCC first neighbours (prim coordinates)
  double p_far = 0, p_not = 0, p_nn[4];
  if (!potential_with_vacancy_at(far, p_far) || !potential_with_vacancy_at(not_nn, p_not)) return fail("general supercell: cell lookup");
  for (int q = 0; q < 4; ++q)
    if (!potential_with_vacancy_at(nn[q], p_nn[q])) return fail("general supercell: cell lookup");
  // the same pair in a diag(16) box through the C ABI
  cmx_tables *t = nullptr;
  cmx_state *s = nullptr;
  if (cmx_tables_create_from_file(tables_path, 0, &t)) return fail(cmx_last_error());
  if (cmx_state_create(t, 16, 16, 16, 1, 0, &s)) return fail(cmx_last_error());
  ClexData const &clex = system->clex_data.at("formation_energy");
  std::vector<uint32_t> idx(clex.coefficients.index.begin(), clex.coefficients.index.end());
  cmx_state_set_eci(s, (int32_t)idx.size(), idx.data(), clex.coefficients.value.data());
  std::vector<int32_t> box(4096, 0);
  double e_nn = 0, e_far = 0;
  box[0] = 1;
  box[1] = 2;  // cell (1, 0, 0)
  cmx_state_upload_occ(s, 0, box.data());
  cmx_energy(s, 0, &e_nn);
  box[1] = 0;
  box[3 + 16 * (3 + 16 * 3)] = 2;  // cell (3, 3, 3)
  cmx_state_upload_occ(s, 0, box.data());
  cmx_energy(s, 0, &e_far);
  cmx_state_destroy(s);
  s = nullptr;
  const double pair = e_nn - e_far;
  if (!(std::fabs(pair) > 1e-6)) return fail("general supercell: the probe pair does not interact");
  if (std::fabs(p_not - p_far) > 1e-9) return fail("general supercell: a non-neighbour pair interacts (site mapping)");
  for (int q = 0; q < 4; ++q)
    if (std::fabs((p_nn[q] - p_far) - pair) > 1e-9) return fail("general supercell: first-neighbour pair energy (site mapping)");

  // (2) + (3) on a random configuration
  std::mt19937_64 init(11);
  for (Index l = 0; l < n_cells; ++l) occ[l] = (int)(init() % 3);
  const Eigen::VectorXi occ0 = occ;
  calc.set_state_and_potential(state, &occ_location);
  const double p0 = calc.potential->per_supercell();
  for (Index l : {Index(0), Index(200), Index(511)}) {
    const int new_occ = (occ0[l] + 1) % 3;
    const double dE = calc.potential->occ_delta_per_supercell({l}, {new_occ});
    occ[l] = new_occ;
    calc.set_state_and_potential(state, &occ_location);
    const double p1 = calc.potential->per_supercell();
    occ[l] = occ0[l];
    if (std::fabs((p1 - p0) - dE) > 1e-9) return fail("general supercell: occ_delta_per_supercell != difference of potentials");
  }
  run_manager_type<BaseMonteCalculator::engine_type> run_manager;
  run_manager.engine = std::make_shared<std::mt19937_64>(99);
  run_manager.sample_period = 2;
  run_manager.n_samples_max = 3;  // samples at passes 0, 2, 4
  run_manager.sampler = [](state_type const &) {};
  calc.run(state, occ_location, run_manager);
  if (run_manager.pass != 4) return fail("general supercell: pass count");
  int32_t T9[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T9[3 * i + j] = (int32_t)Tg[i][j];
  if (cmx_state_create_general(t, T9, 1, 0, &s)) return fail(cmx_last_error());
  cmx_state_set_eci(s, (int32_t)idx.size(), idx.data(), clex.coefficients.value.data());
  std::vector<int32_t> ijk(3 * n_cells);
  for (Index u = 0; u < n_cells; ++u)
    for (int a = 0; a < 3; ++a) ijk[3 * u + a] = (int32_t)convert.l_to_ijk(u)[a];
  std::vector<int64_t> order(n_cells);
  if (cmx_state_cell_index(s, n_cells, ijk.data(), order.data())) return fail(cmx_last_error());
  bool identity = true;
  for (Index u = 0; u < n_cells; ++u) identity = identity && order[u] == u;
  if (identity) return fail("general supercell: the stand-in's unit-cell order equals the library's (the case tests nothing)");
  if (cmx_state_set_site_order(s, order.data())) return fail(cmx_last_error());
  cmx_state_upload_occ(s, 0, occ0.data());
  std::vector<double> exch(9, 0.0);
  const double mu_of_species[3] = {0.0, mu[0], mu[1]};
  for (int oi = 0; oi < 3; ++oi)
    for (int of = 0; of < 3; ++of) exch[oi * 3 + of] = mu_of_species[of] - mu_of_species[oi];
  cmx_state_set_conditions(s, 0, 900.0, exch.data());
  long long acc = 0;
  for (Index p = 0; p < 4; p += 2) {
    cmx_counters c;
    if (cmx_sgc_sweep(s, 2, 12345, p, &c)) return fail(cmx_last_error());
    acc += c.n_accept;
  }
  std::vector<int32_t> ref(n_cells);
  cmx_state_download_occ(s, 0, ref.data());
  for (Index l = 0; l < n_cells; ++l)
    if (ref[l] != occ[l]) return fail("general supercell: occupation after run differs from the C ABI");
  if (acc != run_manager.n_accept || acc == 0) return fail("general supercell: acceptance count");
  cmx_state_destroy(s);
  cmx_tables_destroy(t);
  std::printf("general supercell ok: pair energy %.6f, accepted %lld of %lld\n", pair, acc, (long long)(4 * n_cells));
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 3) return fail("usage: test_plugin <plugin.so> <tables.flat> [--no-gpu]");
  const bool no_gpu = argc > 3 && !std::strcmp(argv[3], "--no-gpu");
  void *lib = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return fail(dlerror());
  typedef BaseMonteCalculator *(*factory_t)();
  factory_t make = (factory_t)dlsym(lib, "make_B200SemiGrandCanonicalCalculator");
  if (!make) return fail("make_B200SemiGrandCanonicalCalculator not exported");
  std::unique_ptr<BaseMonteCalculator> calc(make());
  if (calc->calculator_name != "B200SemiGrandCanonicalCalculator") return fail("calculator_name");
  if (!calc->required_clex.count("formation_energy") || !calc->required_params.count("cmx_tables"))
    return fail("required clex / params");

  // the sampling / analysis / fixture maps are the reference's own (SemiGrandCanonicalCalculator.cc:238-330),
  // handed through from its calculator of the same ensemble
  {
    auto sf = calc->standard_sampling_functions(nullptr);
    auto af = calc->standard_analysis_functions(nullptr);
    if (!sf.count("potential_energy") || !sf.count("clex.formation_energy") || !sf.count("param_chem_pot"))
      return fail("standard_sampling_functions are not the reference's");
    if (!af.count("heat_capacity") || !af.count("param_susc")) return fail("standard_analysis_functions are not the reference's");
    if (!calc->standard_modifying_functions(nullptr).empty()) return fail("standard_modifying_functions");
    if (calc->standard_selected_event_functions(nullptr).has_value()) return fail("standard_selected_event_functions");
    auto fx = calc->make_default_sampling_fixture_params(nullptr, "thermo", true, false, false, true, std::nullopt,
                                                         std::nullopt, 600.0);
    if (fx.label != "thermo" || fx.sampler_names.size() != 4) return fail("make_default_sampling_fixture_params");
  }

  // the FCC A-B-Va test system (tests/unit/clexmonte/data/FCC_binary_vacancy): one sublattice,
  // occupants A, B, Va; composition axes origin A, end members B and Va; shipped sparse ECI
  auto system = std::make_shared<system_type>();
  system->sublat_to_asym = {0};
  system->occ_to_species = {{0, 1, 2}};
  system->composition_converter.m_components = {"A", "B", "Va"};
  system->composition_converter.m_origin = Eigen::VectorXd(3);
  system->composition_converter.m_origin[0] = 1.0;
  system->composition_converter.m_Rt = Eigen::MatrixXd(2, 3);
  system->composition_converter.m_Rt(0, 1) = 1.0;  // x_a = n_B
  system->composition_converter.m_Rt(1, 2) = 1.0;  // x_b = n_Va
  ClexData clex;
  clex.basis_set_name = "default";
  clex.coefficients.index = {1, 2, 3, 4, 5};
  clex.coefficients.value = {-0.1, 0.3, 0.1, 0.1, 0.5};  // formation_energy_sparse_eci.json
  system->clex_data["formation_energy"] = clex;

  // missing required parameter -> reset throws (BaseMonteCalculator::_check_params)
  try {
    calc->reset(jsonParser(), system);
    return fail("reset() accepted params without cmx_tables");
  } catch (std::runtime_error const &) {
  }
  // the stand-in Conversions of a skewed supercell: det(T) unit cells, pairwise inequivalent
  // modulo the supercell lattice, in an order that is not the library's box order
  {
    state_type skew;
    const long Tg[3][3] = {{8, 0, 0}, {2, 8, 0}, {4, 2, 8}};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) skew.configuration.transformation_matrix_to_super(i, j) = Tg[i][j];
    monte::OccLocation none;
    StateData sd(system, &skew, &none);
    if (sd.n_unitcells != 512 || sd.convert->m_unitl_to_ijk.size() != 512) return fail("stand-in unit cell enumeration");
    for (Index u = 0; u < 512; u += 37) {
      auto const c = sd.convert->l_to_ijk(u);
      const long x[3] = {c[0] + 8, c[1] + 2 + 8, c[2] + 4 + 2 - 8};  // + (column 0) + (column 1) - (column 2)
      if (find_unitl(*sd.convert, sd.transformation_matrix_to_super, x) != u) return fail("stand-in unit cells are not a transversal of the supercell lattice");
    }
  }
  if (no_gpu) {
    std::printf("plugin ok (interface only: no GPU)\n");
    return 0;
  }

  jsonParser params;
  params.strings["cmx_tables"] = argv[2];
  params.numbers["cmx_seed"] = 12345.0;
  calc->reset(params, system);

  const Index N = 16, n_cells = N * N * N;
  state_type state;
  for (int a = 0; a < 3; ++a) state.configuration.transformation_matrix_to_super(a, a) = N;
  state.configuration.dof_values.occupation = Eigen::VectorXi(n_cells);
  std::mt19937_64 init(7);
  for (Index l = 0; l < n_cells; ++l) state.configuration.dof_values.occupation[l] = (int)(init() % 3);
  const Eigen::VectorXi occ0 = state.configuration.dof_values.occupation;
  state.conditions.scalar_values["temperature"] = 900.0;
  Eigen::VectorXd mu(2);
  mu[0] = 0.2;
  mu[1] = -0.1;
  state.conditions.vector_values["param_chem_pot"] = mu;
  monte::OccLocation occ_location;
  occ_location.m_mol_size = n_cells;

  // validate_state rejects a supercell of non-positive volume
  {
    state_type bad = state;
    bad.configuration.transformation_matrix_to_super(0, 0) = -N;
    if (calc->validate_state(bad).valid()) return fail("validate_state accepted a left-handed supercell");
  }

  // potential: delta of a proposed event == C ABI on the same configuration
  calc->set_state_and_potential(state, &occ_location);
  const double p0 = calc->potential->per_supercell();
  double dE_sum_check = 0.0;
  for (Index l : {Index(0), Index(17), Index(4095)}) {
    const int new_occ = (occ0[l] + 1) % 3;
    const double dE = calc->potential->occ_delta_per_supercell({l}, {new_occ});
    if (!std::isfinite(dE)) return fail("occ_delta_per_supercell not finite");
    dE_sum_check += dE;
  }

  // the run: 6 samples, one every 2 passes; sampler records the potential per supercell
  run_manager_type<BaseMonteCalculator::engine_type> run_manager;
  run_manager.engine = std::make_shared<std::mt19937_64>(99);
  run_manager.sample_period = 2;
  run_manager.n_samples_max = 6;
  std::vector<double> sampled;
  run_manager.sampler = [&](state_type const &) { sampled.push_back(calc->potential->per_supercell()); };
  calc->run(state, occ_location, run_manager);
  if (!run_manager.finalized || run_manager.n_samples != 6) return fail("run manager protocol");
  const Index passes = 10;  // samples at passes 0, 2, 4, 6, 8, 10
  if (run_manager.pass != passes) return fail("pass count");
  if (run_manager.n_accept + run_manager.n_reject != passes * n_cells) return fail("accept + reject != attempts");
  if (occ_location.n_initialize != 1) return fail("occ_location not re-initialised");
  if (std::fabs(sampled.front() - p0) > 1e-9 * std::fabs(p0)) return fail("first sample is not the initial potential");

  // the same through the C ABI directly: same seed, same passes -> same occupation, same counts
  cmx_tables *t = nullptr;
  cmx_state *s = nullptr;
  if (cmx_tables_create_from_file(argv[2], 0, &t)) return fail(cmx_last_error());
  if (cmx_state_create(t, N, N, N, 1, 0, &s)) return fail(cmx_last_error());
  std::vector<uint32_t> idx(clex.coefficients.index.begin(), clex.coefficients.index.end());
  cmx_state_set_eci(s, (int32_t)idx.size(), idx.data(), clex.coefficients.value.data());
  cmx_state_upload_occ(s, 0, occ0.data());
  // exch[0][oi][of] = mu . R^T (e_of - e_oi): species = occupant index here
  std::vector<double> exch(9, 0.0);
  const double mu_of_species[3] = {0.0, mu[0], mu[1]};
  for (int oi = 0; oi < 3; ++oi)
    for (int of = 0; of < 3; ++of) exch[oi * 3 + of] = mu_of_species[of] - mu_of_species[oi];
  cmx_state_set_conditions(s, 0, 900.0, exch.data());
  long long acc = 0;
  for (Index p = 0; p < passes; p += 2) {
    cmx_counters c;
    if (cmx_sgc_sweep(s, 2, 12345, p, &c)) return fail(cmx_last_error());
    acc += c.n_accept;
  }
  std::vector<int32_t> ref(n_cells);
  cmx_state_download_occ(s, 0, ref.data());
  for (Index l = 0; l < n_cells; ++l)
    if (ref[l] != state.configuration.dof_values.occupation[l]) return fail("occupation after run differs from the C ABI");
  if (acc != run_manager.n_accept) return fail("acceptance count differs from the C ABI");
  double E = 0.0;
  cmx_energy(s, 0, &E);
  cmx_state_destroy(s);
  cmx_tables_destroy(t);

  if (int rc = general_supercell_case(*calc, system, argv[2])) return rc;

  // clone: independent device handles, same behaviour
  std::unique_ptr<BaseMonteCalculator> twin = calc->clone();
  state_type state2 = state;
  state2.configuration.dof_values.occupation = occ0;
  monte::OccLocation occ2;
  occ2.m_mol_size = n_cells;
  twin->set_state_and_potential(state2, &occ2);
  if (std::fabs(twin->potential->per_supercell() - p0) > 1e-9 * std::fabs(p0)) return fail("clone potential");

  std::printf("plugin ok: %lld passes, accepted %lld of %lld, potential %.6f -> %.6f (dE probe %.6f)\n",
              (long long)passes, (long long)run_manager.n_accept, (long long)(passes * n_cells), p0, sampled.back(),
              dE_sum_check);
  return 0;
}
