// Drives plugin/B200CanonicalCalculator.cc the way MonteCalculator does
// (monte_calculator/MonteCalculator.cc:142-173): dlopen, look up "make_" + name,
// reset(params, system), run(state, occ_location, run_manager).
//   test_plugin_canonical <libB200CanonicalCalculator.so> <tables.flat> [--no-gpu]
// --no-gpu: stop after the factory / interface checks (CPU CI).  With a GPU: the potential is
// the formation energy of the C ABI, a two-site delta equals the difference of two
// evaluations, a run conserves the composition and leaves the occupation the C ABI gives for
// the same seed and swap table.  Prints "canonical plugin ok ..." and exits 0.
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstring>

#include "casm/clexmonte/monte_calculator/BaseMonteCalculator.hh"
#include "cmx_b200.h"

using namespace CASM;
using namespace CASM::clexmonte;

static int fail(const char *what) {
  std::fprintf(stderr, "test_plugin_canonical: FAILED: %s\n", what);
  return 1;
}

int main(int argc, char **argv) {
  if (argc < 3) return fail("usage: test_plugin_canonical <plugin.so> <tables.flat> [--no-gpu]");
  const bool no_gpu = argc > 3 && !std::strcmp(argv[3], "--no-gpu");
  void *lib = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return fail(dlerror());
  typedef BaseMonteCalculator *(*factory_t)();
  factory_t make = (factory_t)dlsym(lib, "make_B200CanonicalCalculator");
  if (!make) return fail("make_B200CanonicalCalculator not exported");
  std::unique_ptr<BaseMonteCalculator> calc(make());
  if (calc->calculator_name != "B200CanonicalCalculator") return fail("calculator_name");
  if (!calc->required_clex.count("formation_energy") || !calc->required_params.count("cmx_tables"))
    return fail("required clex / params");

  // the sampling / analysis / state-modifying maps are the reference's own (CanonicalCalculator.cc:171-280)
  {
    auto sf = calc->standard_sampling_functions(nullptr);
    if (!sf.count("potential_energy") || !sf.count("mol_composition") || sf.count("param_chem_pot"))
      return fail("standard_sampling_functions are not the reference's");
    if (!calc->standard_analysis_functions(nullptr).count("heat_capacity")) return fail("standard_analysis_functions");
    if (!calc->standard_modifying_functions(nullptr).count("enforce.composition")) return fail("standard_modifying_functions");
    if (!calc->standard_json_sampling_functions(nullptr).count("config")) return fail("standard_json_sampling_functions");
  }

  // the FCC A-B-Va test system (tests/unit/clexmonte/data/FCC_binary_vacancy), shipped sparse ECI
  auto system = std::make_shared<system_type>();
  system->sublat_to_asym = {0};
  system->occ_to_species = {{0, 1, 2}};
  system->composition_converter.m_components = {"A", "B", "Va"};
  system->composition_converter.m_origin = Eigen::VectorXd(3);
  system->composition_converter.m_origin[0] = 1.0;
  system->composition_converter.m_Rt = Eigen::MatrixXd(2, 3);
  system->composition_converter.m_Rt(0, 1) = 1.0;
  system->composition_converter.m_Rt(1, 2) = 1.0;
  ClexData clex;
  clex.basis_set_name = "default";
  clex.coefficients.index = {1, 2, 3, 4, 5};
  clex.coefficients.value = {-0.1, 0.3, 0.1, 0.1, 0.5};
  system->clex_data["formation_energy"] = clex;

  try {
    calc->reset(jsonParser(), system);
    return fail("reset() accepted params without cmx_tables");
  } catch (std::runtime_error const &) {
  }
  if (no_gpu) {
    std::printf("canonical plugin ok (interface only: no GPU)\n");
    return 0;
  }

  jsonParser params;
  params.strings["cmx_tables"] = argv[2];
  params.numbers["cmx_seed"] = 4242.0;
  calc->reset(params, system);

  const Index N = 16, n_cells = N * N * N;
  state_type state;
  for (int a = 0; a < 3; ++a) state.configuration.transformation_matrix_to_super(a, a) = N;
  state.configuration.dof_values.occupation = Eigen::VectorXi(n_cells);
  Eigen::VectorXi &occ = state.configuration.dof_values.occupation;
  std::mt19937_64 init(7);
  long count0[3] = {0, 0, 0};
  for (Index l = 0; l < n_cells; ++l) {
    occ[l] = (int)(init() % 3);
    ++count0[occ[l]];
  }
  const Eigen::VectorXi occ0 = occ;
  state.conditions.scalar_values["temperature"] = 900.0;
  Eigen::VectorXd mol(3);
  for (int q = 0; q < 3; ++q) mol[q] = (double)count0[q] / (double)n_cells;
  state.conditions.vector_values["mol_composition"] = mol;
  monte::OccLocation occ_location;
  occ_location.m_mol_size = n_cells;

  // validate_state: a configuration off the conditions' composition is rejected (CanonicalCalculator.cc:318-360)
  if (!calc->validate_state(state).valid()) return fail("validate_state rejected a consistent state");
  {
    state_type bad = state;
    bad.conditions.vector_values["mol_composition"][0] += 0.1;
    if (calc->validate_state(bad).valid()) return fail("validate_state accepted an inconsistent composition");
    state_type bad2 = state;
    bad2.conditions.scalar_values.clear();
    if (calc->validate_state(bad2).valid()) return fail("validate_state accepted conditions without temperature");
  }

  // the C ABI side of every comparison
  cmx_tables *t = nullptr;
  cmx_state *s = nullptr;
  if (cmx_tables_create_from_file(argv[2], 0, &t)) return fail(cmx_last_error());
  if (cmx_state_create(t, N, N, N, 1, 0, &s)) return fail(cmx_last_error());
  std::vector<uint32_t> idx(clex.coefficients.index.begin(), clex.coefficients.index.end());
  cmx_state_set_eci(s, (int32_t)idx.size(), idx.data(), clex.coefficients.value.data());
  const int32_t asym[1] = {0}, species[3] = {0, 1, 2};
  if (cmx_state_set_occupants(s, asym, species, 3)) return fail(cmx_last_error());
  cmx_state_set_conditions(s, 0, 900.0, nullptr);
  cmx_state_upload_occ(s, 0, occ0.data());
  double e_ref = 0.0;
  cmx_energy(s, 0, &e_ref);

  // potential: the formation energy; a two-site delta = the difference of two evaluations
  calc->set_state_and_potential(state, &occ_location);
  const double p0 = calc->potential->per_supercell();
  if (std::fabs(p0 - e_ref) > 1e-9 * std::fabs(e_ref)) return fail("per_supercell is not the formation energy of the C ABI");
  Index la = 5, lb = 6;
  while (occ0[lb] == occ0[la]) ++lb;
  const double dE = calc->potential->occ_delta_per_supercell({la, lb}, {occ0[lb], occ0[la]});
  std::swap(occ[la], occ[lb]);
  calc->set_state_and_potential(state, &occ_location);
  const double p1 = calc->potential->per_supercell();
  std::swap(occ[la], occ[lb]);
  if (std::fabs((p1 - p0) - dE) > 1e-9) return fail("occ_delta_per_supercell != difference of potentials");

  // the run: samples at passes 0, 2, 4, 6 (a pass = one sweep over the swap table)
  run_manager_type<BaseMonteCalculator::engine_type> run_manager;
  run_manager.engine = std::make_shared<std::mt19937_64>(99);
  run_manager.sample_period = 2;
  run_manager.n_samples_max = 4;
  std::vector<double> sampled;
  run_manager.sampler = [&](state_type const &) { sampled.push_back(calc->potential->per_supercell()); };
  calc->run(state, occ_location, run_manager);
  const Index passes = 6;
  if (!run_manager.finalized || run_manager.n_samples != 4 || run_manager.pass != passes) return fail("run manager protocol");
  if (run_manager.n_accept + run_manager.n_reject != passes * n_cells) return fail("accept + reject != steps");
  if (run_manager.n_accept <= 0 || run_manager.n_accept >= passes * n_cells) return fail("acceptance count");
  if (occ_location.n_initialize != 1) return fail("occ_location not re-initialised");
  if (std::fabs(sampled.front() - p0) > 1e-9 * std::fabs(p0)) return fail("first sample is not the initial potential");
  long count1[3] = {0, 0, 0};
  for (Index l = 0; l < n_cells; ++l) ++count1[occ[l]];
  for (int q = 0; q < 3; ++q)
    if (count1[q] != count0[q]) return fail("composition not conserved");

  // the same through the C ABI: default swap table, same seed, same passes
  int32_t n_swaps = 0;
  if (cmx_canonical_default_swaps(s, 12, 1, 0, nullptr, &n_swaps) || n_swaps <= 0) return fail("default swap table");
  std::vector<cmx_swap_type> swaps((size_t)n_swaps);
  if (cmx_canonical_default_swaps(s, 12, 1, n_swaps, swaps.data(), &n_swaps)) return fail(cmx_last_error());
  if (cmx_canonical_set_swaps(s, n_swaps, swaps.data())) return fail(cmx_last_error());
  long long acc = 0, att = 0;
  for (Index p = 0; p < passes; p += 2) {
    cmx_counters c;
    if (cmx_canonical_sweep(s, 2, 4242, p, &c)) return fail(cmx_last_error());
    acc += c.n_accept;
    att += c.n_attempt;
  }
  std::vector<int32_t> ref(n_cells);
  cmx_state_download_occ(s, 0, ref.data());
  for (Index l = 0; l < n_cells; ++l)
    if (ref[l] != occ[l]) return fail("occupation after run differs from the C ABI");
  double e_end = 0.0;
  cmx_energy(s, 0, &e_end);
  if (std::fabs(sampled.back() - e_end) > 1e-9 * std::fabs(e_end)) return fail("last sample is not the final formation energy");
  // the RunManager saw the acceptance rate of the sweeps (scaled to steps_per_pass steps per pass)
  const double rate = (double)acc / (double)att, seen = (double)run_manager.n_accept / (double)(passes * n_cells);
  if (std::fabs(rate - seen) > 1e-3) return fail("acceptance rate handed to the RunManager");
  cmx_state_destroy(s);
  cmx_tables_destroy(t);

  std::printf("canonical plugin ok: %lld passes, %d swap types, %lld of %lld unlike pairs exchanged, E %.6f -> %.6f\n",
              (long long)passes, (int)n_swaps, acc, att, p0, sampled.back());
  return 0;
}
