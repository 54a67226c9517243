// Shared by the B200 calculators (plugin/B200SemiGrandCanonicalCalculator.cc,
// plugin/B200CanonicalCalculator.cc): the device handles of one (system, supercell) and the
// binding of a reference state to them through the C ABI of include/cmx_b200.h.
#ifndef B200_PLUGIN_COMMON_HH
#define B200_PLUGIN_COMMON_HH

#include "casm/clexmonte/monte_calculator/BaseMonteCalculator.hh"
#include "casm/clexmonte/monte_calculator/StateData.hh"
#include "casm/clexmonte/system/System.hh"
#include "cmx_b200.h"

namespace CASM {
namespace clexmonte {
namespace b200 {

inline void cmx_check(int rc, const char *who) {
  if (rc != CMX_OK) throw std::runtime_error(std::string(who) + ": " + cmx_last_error());
}

// device handles of one (system, supercell): never shared between calculator clones
struct DeviceState {
  cmx_tables *tables = nullptr;
  cmx_state *state = nullptr;
  long T[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // the transformation matrix the state was created for
  // the reference's linear site index l -> the library's (empty: the same).  The order of
  // the unit cells in a supercell is xtal::UnitCellIndexConverter's [EXT]; it is ASKED
  // (Conversions::l_to_ijk), never assumed.
  std::vector<int64_t> site_order;
  int64_t to_library(Index l) const { return site_order.empty() ? (int64_t)l : site_order[(size_t)l]; }
  ~DeviceState() {
    if (state) cmx_state_destroy(state);
    if (tables) cmx_tables_destroy(tables);
  }
};

inline int max_occupants(system_type const &system) {
  int m = 0;
  for (auto const &sp : system.occ_to_species) m = std::max<int>(m, (int)sp.size());
  return m;
}

inline long determinant(Eigen::Matrix3l const &T) {
  return T(0, 0) * (T(1, 1) * T(2, 2) - T(1, 2) * T(2, 1)) - T(0, 1) * (T(1, 0) * T(2, 2) - T(1, 2) * T(2, 0)) +
         T(0, 2) * (T(1, 0) * T(2, 1) - T(1, 1) * T(2, 0));
}

/// tables named by the `cmx_tables` param (a calculator's _reset)
inline void load_tables(DeviceState &dev, jsonParser const &params, const char *who) {
  const int device = params.contains("cmx_device") ? (int)params.get_number("cmx_device") : 0;
  cmx_check(cmx_tables_create_from_file(params.get_string("cmx_tables").c_str(), device, &dev.tables), who);
}

/// The device state of `state_data`'s supercell: created when the transformation matrix
/// changed (diag(N0, N1, N2): the row kernels; otherwise the general-supercell path), with the
/// reference's site order, the formation_energy ECI and the occupant bookkeeping; then the
/// host occupation is uploaded.  Returns true when the state was (re)created.
inline bool bind_state(DeviceState &dev, StateData const &state_data, system_type const &system, int max_occ,
                       const char *who) {
  Eigen::Matrix3l const &T = state_data.transformation_matrix_to_super;
  bool same = dev.state != nullptr, diagonal = true;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      same = same && dev.T[3 * i + j] == T(i, j);
      diagonal = diagonal && (i == j || T(i, j) == 0);
    }
  const Index n_cells = state_data.n_unitcells;
  const size_t n_sublat = system.occ_to_species.size();
  if (!same) {
    if (dev.state) cmx_state_destroy(dev.state);
    dev.state = nullptr;
    int32_t T9[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) T9[3 * i + j] = (int32_t)(dev.T[3 * i + j] = T(i, j));
    if (diagonal)
      cmx_check(cmx_state_create(dev.tables, T9[0], T9[4], T9[8], 1, 0, &dev.state), who);
    else
      cmx_check(cmx_state_create_general(dev.tables, T9, 1, 0, &dev.state), who);
    // the reference's site order: l = b * n_unitcells + unitl, unit cell `unitl` at l_to_ijk(l)
    monte::Conversions const &convert = *state_data.convert;
    const Index n_sites = n_cells * (Index)n_sublat;
    std::vector<int32_t> ijk(3 * (size_t)n_cells);
    for (Index u = 0; u < n_cells; ++u) {
      auto const cell = convert.l_to_ijk(u);
      for (int a = 0; a < 3; ++a) ijk[3 * (size_t)u + a] = (int32_t)cell[a];
    }
    std::vector<int64_t> cell_index((size_t)n_cells);
    cmx_check(cmx_state_cell_index(dev.state, n_cells, ijk.data(), cell_index.data()), who);
    dev.site_order.resize((size_t)n_sites);
    bool identity = true;
    for (Index l = 0; l < n_sites; ++l) {
      dev.site_order[(size_t)l] = (int64_t)convert.l_to_b(l) * n_cells + cell_index[(size_t)(l % n_cells)];
      identity = identity && dev.site_order[(size_t)l] == (int64_t)l;
    }
    if (identity) dev.site_order.clear();
    cmx_check(cmx_state_set_site_order(dev.state, identity ? nullptr : dev.site_order.data()), who);
    clexulator::SparseCoefficients const &eci = get_clex_data(system, "formation_energy").coefficients;
    std::vector<uint32_t> index(eci.index.begin(), eci.index.end());
    cmx_check(cmx_state_set_eci(dev.state, (int32_t)index.size(), index.data(), eci.value.data()), who);
    // occupant bookkeeping (Conversions: asym unit, species): the reference-order mode, swaps
    std::vector<int32_t> asym(system.sublat_to_asym.begin(), system.sublat_to_asym.end());
    std::vector<int32_t> species(n_sublat * max_occ, -1);
    for (size_t b = 0; b < n_sublat; ++b)
      for (size_t o = 0; o < system.occ_to_species[b].size(); ++o)
        species[b * max_occ + o] = (int32_t)system.occ_to_species[b][o];
    cmx_check(cmx_state_set_occupants(dev.state, asym.data(), species.data(),
                                      (int32_t)get_composition_converter(system).components().size()),
              who);
  }
  Eigen::VectorXi const &occupation = get_occupation(*state_data.state);
  if (occupation.size() != n_cells * (Index)n_sublat) throw std::runtime_error(std::string(who) + ": occupation size mismatch");
  cmx_check(cmx_state_upload_occ(dev.state, 0, occupation.data()), who);
  return !same;
}

/// What a B200 calculator does NOT re-implement: the sampling functions, analysis functions,
/// state-modifying functions and default sampling fixture of its ensemble are the reference's
/// own.  They are obtained from the reference's calculator of the same ensemble (the C-linkage
/// factories libcasm_clexmonte exports, forward-declared the way
/// python/src/clexmonte_monte_calculator.cpp:34-45 does) and read the state through
/// `calculation->state_data()` / `calculation->potential()` -- i.e. through THIS calculator's
/// StateData and device potential.  Sampler names, JSON output and post-processing stay unchanged.
class DelegatingCalculator : public BaseMonteCalculator {
 public:
  typedef BaseMonteCalculator *(*reference_factory)();
  DelegatingCalculator(reference_factory make_reference, std::string _calculator_name,
                       std::set<std::string> _required_basis_set, std::set<std::string> _required_local_basis_set,
                       std::set<std::string> _required_clex, std::set<std::string> _required_multiclex,
                       std::set<std::string> _required_local_clex, std::set<std::string> _required_local_multiclex,
                       std::set<std::string> _required_dof_spaces, std::set<std::string> _required_params,
                       std::set<std::string> _optional_params, bool _time_sampling_allowed, bool _update_atoms,
                       bool _save_atom_info, bool _is_multistate_method)
      : BaseMonteCalculator(_calculator_name, _required_basis_set, _required_local_basis_set, _required_clex,
                            _required_multiclex, _required_local_clex, _required_local_multiclex, _required_dof_spaces,
                            _required_params, _optional_params, _time_sampling_allowed, _update_atoms, _save_atom_info,
                            _is_multistate_method),
        m_reference(make_reference()) {}

  std::map<std::string, state_sampling_function_type> standard_sampling_functions(
      std::shared_ptr<MonteCalculator> const &calculation) const override {
    return m_reference->standard_sampling_functions(calculation);
  }
  std::map<std::string, json_state_sampling_function_type> standard_json_sampling_functions(
      std::shared_ptr<MonteCalculator> const &calculation) const override {
    return m_reference->standard_json_sampling_functions(calculation);
  }
  std::map<std::string, results_analysis_function_type> standard_analysis_functions(
      std::shared_ptr<MonteCalculator> const &calculation) const override {
    return m_reference->standard_analysis_functions(calculation);
  }
  StateModifyingFunctionMap standard_modifying_functions(
      std::shared_ptr<MonteCalculator> const &calculation) const override {
    return m_reference->standard_modifying_functions(calculation);
  }
  std::optional<monte::SelectedEventFunctions> standard_selected_event_functions(
      std::shared_ptr<MonteCalculator> const &calculation) const override {
    return m_reference->standard_selected_event_functions(calculation);
  }
  sampling_fixture_params_type make_default_sampling_fixture_params(
      std::shared_ptr<MonteCalculator> const &calculation, std::string label, bool write_results,
      bool write_trajectory, bool write_observations, bool write_status, std::optional<std::string> output_dir,
      std::optional<std::string> log_file, double log_frequency_in_s) const override {
    return m_reference->make_default_sampling_fixture_params(calculation, label, write_results, write_trajectory,
                                                             write_observations, write_status, output_dir, log_file,
                                                             log_frequency_in_s);
  }

 private:
  std::unique_ptr<BaseMonteCalculator> m_reference;
};

/// RunManager counters of n_attempt steps (occupation_metropolis.hh:109-116)
template <typename RunManagerType>
inline void count_steps(RunManagerType &run_manager, Index n_passes, int64_t n_attempt, int64_t n_accept) {
#ifdef CMX_HAVE_RUNMANAGER_BULK
  run_manager.add_passes(n_passes, n_accept, n_attempt - n_accept);
#else
  (void)n_passes;
  for (int64_t q = 0; q < n_accept; ++q) run_manager.increment_n_accept();
  for (int64_t q = n_accept; q < n_attempt; ++q) run_manager.increment_n_reject();
  for (int64_t q = 0; q < n_attempt; ++q) run_manager.increment_step();
#endif
}

}  // namespace b200
}  // namespace clexmonte
}  // namespace CASM

#endif
