// Stand-in for the two factories libcasm_clexmonte exports (SemiGrandCanonicalCalculator.cc:517-523,
// CanonicalCalculator.cc:472-478): the B200 calculators obtain the standard sampling / analysis /
// state-modifying functions and the default sampling fixture of their ensemble from the
// reference's own calculator.  Here a minimal calculator that returns the reference's NAMES
// (the functions themselves are [EXT]-typed, opaque in plugin/shim): enough for the drivers to
// check that the plugins hand the reference's maps through.  Not part of the product.
#include "casm/clexmonte/monte_calculator/BaseMonteCalculator.hh"

namespace CASM {
namespace clexmonte {
namespace {

class ReferenceStandIn : public BaseMonteCalculator {
 public:
  ReferenceStandIn(std::string name, bool semigrand)
      : BaseMonteCalculator(name, {}, {}, {"formation_energy"}, {}, {}, {}, {}, {}, {}, false, false, false, false),
        m_semigrand(semigrand) {}

  std::map<std::string, state_sampling_function_type> standard_sampling_functions(
      std::shared_ptr<MonteCalculator> const &) const override {
    // monte_calculator::common_sampling_functions (sampling_functions.cc:383-425) + the ensemble's own
    std::map<std::string, state_sampling_function_type> m;
    for (const char *n : {"temperature", "mol_composition", "param_composition", "clex.formation_energy",
                          "potential_energy", "corr.default"})
      m[n] = state_sampling_function_type{n};
    if (m_semigrand) m["param_chem_pot"] = state_sampling_function_type{"param_chem_pot"};
    return m;
  }
  std::map<std::string, json_state_sampling_function_type> standard_json_sampling_functions(
      std::shared_ptr<MonteCalculator> const &) const override {
    return {{"config", json_state_sampling_function_type{"config"}}};
  }
  std::map<std::string, results_analysis_function_type> standard_analysis_functions(
      std::shared_ptr<MonteCalculator> const &) const override {
    std::map<std::string, results_analysis_function_type> m;
    m["heat_capacity"] = results_analysis_function_type{"heat_capacity"};
    if (m_semigrand)
      for (const char *n : {"mol_susc", "param_susc", "mol_thermochem_susc", "param_thermochem_susc"})
        m[n] = results_analysis_function_type{n};
    return m;
  }
  StateModifyingFunctionMap standard_modifying_functions(std::shared_ptr<MonteCalculator> const &) const override {
    StateModifyingFunctionMap m;
    if (!m_semigrand)
      for (const char *n : {"match.mol_composition", "enforce.composition"}) m[n] = StateModifyingFunction{n};
    return m;
  }
  std::optional<monte::SelectedEventFunctions> standard_selected_event_functions(
      std::shared_ptr<MonteCalculator> const &) const override {
    return std::nullopt;
  }
  sampling_fixture_params_type make_default_sampling_fixture_params(std::shared_ptr<MonteCalculator> const &,
                                                                    std::string label, bool, bool, bool, bool,
                                                                    std::optional<std::string>,
                                                                    std::optional<std::string>, double) const override {
    return sampling_fixture_params_type{
        label, {"clex.formation_energy", "potential_energy", "mol_composition", "param_composition"}};
  }
  Validator validate_configuration(state_type &) const override { return Validator(); }
  Validator validate_conditions(state_type &) const override { return Validator(); }
  Validator validate_state(state_type &) const override { return Validator(); }
  void set_state_and_potential(state_type &, monte::OccLocation *) override {
    throw std::runtime_error("reference stand-in: not a calculator");
  }
  void set_event_data() override {}
  void run(state_type &, monte::OccLocation &, run_manager_type<engine_type> &) override {
    throw std::runtime_error("reference stand-in: not a calculator");
  }
  void run(int, std::vector<state_type> &, std::vector<monte::OccLocation> &, run_manager_type<engine_type> &) override {
    throw std::runtime_error("reference stand-in: not a calculator");
  }

 private:
  bool m_semigrand;
  void _reset() override {}
  BaseMonteCalculator *_clone() const override { return new ReferenceStandIn(calculator_name, m_semigrand); }
};

}  // namespace
}  // namespace clexmonte
}  // namespace CASM

extern "C" {
CASM::clexmonte::BaseMonteCalculator *make_SemiGrandCanonicalCalculator() {
  return new CASM::clexmonte::ReferenceStandIn("SemiGrandCanonicalCalculator", true);
}
CASM::clexmonte::BaseMonteCalculator *make_CanonicalCalculator() {
  return new CASM::clexmonte::ReferenceStandIn("CanonicalCalculator", false);
}
}
