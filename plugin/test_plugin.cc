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

  // validate_state rejects a non-diagonal supercell
  {
    state_type bad = state;
    bad.configuration.transformation_matrix_to_super(0, 1) = 1;
    if (calc->validate_state(bad).valid()) return fail("validate_state accepted a skewed supercell");
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
