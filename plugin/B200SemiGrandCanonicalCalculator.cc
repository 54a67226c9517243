// B200SemiGrandCanonicalCalculator -- the reference-side plugin of the B200 hot path.
//
// A Monte Carlo method of CASM clexmonte is a shared library that exports
//   extern "C" CASM::clexmonte::BaseMonteCalculator *make_<Name>();
// (src/casm/clexmonte/monte_calculator/SemiGrandCanonicalCalculator.cc:517-523; loaded by
// make_monte_calculator_from_source, monte_calculator/MonteCalculator.cc:142-173).  This file
// is that library for the semi-grand canonical ensemble: the same calculator as the
// reference's (SemiGrandCanonicalCalculator.cc:216-515) whose potential and whose Metropolis
// loop run on the GPU through the C ABI of include/cmx_b200.h -- no other CASM type crosses
// that boundary.
//
// Build (with libcasm installed: drop -Iplugin/shim, add the libcasm include / link flags):
//   make -C plugin          (g++ -std=c++17 -O2 -fPIC -shared -Iplugin/shim -Iinclude ...
//                            -Lcasmcode_clexmonte_b200 -lcmx_b200)
//
// Calculator params (the `params` json of MonteCalculator):
//   "cmx_tables"      path of the flat table export of the formation_energy basis set
//                     (python -m casmcode_clexmonte_b200.clexulator_tables <Clexulator.cc> <out.npz> --flat <out.cmxt>)
//   "cmx_device"      CUDA device (default 0)
//   "cmx_seed"        seed of the counter-based generator; default: one draw of run_manager.engine
//   "cmx_reference_order"  != 0: reproduce the reference's proposal order bit for bit
//                     (cmx_metropolis_sequential) instead of the checkerboard sweeps
#include "b200_common.hh"

extern "C" {
/// the reference's calculator of this ensemble (libcasm_clexmonte,
/// SemiGrandCanonicalCalculator.cc:517-523): source of the standard sampling / analysis functions
CASM::clexmonte::BaseMonteCalculator *make_SemiGrandCanonicalCalculator();
}

namespace CASM {
namespace clexmonte {

namespace {

using b200::DeviceState;
void cmx_check(int rc) { b200::cmx_check(rc, "B200SemiGrandCanonicalCalculator"); }

// exch[b][occ_i][occ_f] = mu_x . R^T (e_species(occ_f) - e_species(occ_i)): the exchange term
// of SemiGrandCanonicalPotential::occ_delta_per_supercell (SemiGrandCanonicalCalculator.cc:186-213)
std::vector<double> exchange_table(system_type const &system, Eigen::VectorXd const &param_chem_pot, int max_occ) {
  composition::CompositionConverter const &cc = get_composition_converter(system);
  Eigen::MatrixXd Rt = cc.dparam_dmol();
  const size_t n_sublat = system.occ_to_species.size();
  std::vector<double> exch(n_sublat * max_occ * max_occ, 0.0);
  Eigen::VectorXd delta_N((long)cc.components().size());
  for (size_t b = 0; b < n_sublat; ++b) {
    auto const &sp = system.occ_to_species[b];
    for (size_t oi = 0; oi < sp.size(); ++oi)
      for (size_t of = 0; of < sp.size(); ++of) {
        delta_N.setZero();
        delta_N[sp[oi]] += -1.0;
        delta_N[sp[of]] += 1.0;
        exch[(b * max_occ + oi) * max_occ + of] = param_chem_pot.dot(Rt * delta_N);
      }
  }
  return exch;
}

}  // namespace

/// SemiGrandCanonicalPotential (SemiGrandCanonicalCalculator.cc:125-213) on the device
class B200SemiGrandCanonicalPotential : public BaseMontePotential {
 public:
  B200SemiGrandCanonicalPotential(std::shared_ptr<StateData> _state_data, std::shared_ptr<DeviceState> _dev,
                                  int _max_occ)
      : BaseMontePotential(_state_data),
        dev(_dev),
        max_occ(_max_occ),
        n_unitcells(_state_data->n_unitcells),
        composition_converter(get_composition_converter(*_state_data->system)),
        param_chem_pot(_state_data->state->conditions.vector_values.at("param_chem_pot")) {
    if (param_chem_pot.size() != composition_converter.independent_compositions())
      throw std::runtime_error("Error in B200SemiGrandCanonicalPotential: param_chem_pot size error");
  }

  std::shared_ptr<DeviceState> dev;
  int max_occ;
  Index n_unitcells;
  composition::CompositionConverter const &composition_converter;
  Eigen::VectorXd const &param_chem_pot;

  /// E - N_unit mu_x . x  (:171-179); E and the occupant counts come from the device
  double per_supercell() override {
    double E = 0.0;
    cmx_check(cmx_energy(dev->state, 0, &E));
    system_type const &system = *state_data->system;
    const size_t n_sublat = system.occ_to_species.size();
    std::vector<int64_t> counts(n_sublat * max_occ, 0);
    cmx_check(cmx_composition(dev->state, 0, counts.data()));
    Eigen::VectorXd mol((long)composition_converter.components().size());
    for (size_t b = 0; b < n_sublat; ++b)
      for (size_t o = 0; o < system.occ_to_species[b].size(); ++o)
        mol[system.occ_to_species[b][o]] += (double)counts[b * max_occ + o] / (double)n_unitcells;
    return E - n_unitcells * param_chem_pot.dot(composition_converter.param_composition(mol));
  }
  double per_unitcell() override { return this->per_supercell() / n_unitcells; }

  /// dE_clex - mu_x . R^T dN of a proposed event (:186-213), on the device-resident occupation
  double occ_delta_per_supercell(std::vector<Index> const &linear_site_index,
                                 std::vector<int> const &new_occ) override {
    std::vector<int64_t> l(linear_site_index.size());
    for (size_t q = 0; q < l.size(); ++q) l[q] = dev->to_library(linear_site_index[q]);
    std::vector<int32_t> occ(new_occ.begin(), new_occ.end());
    double dE = 0.0;
    cmx_check(cmx_delta_e(dev->state, 0, 1, (int32_t)l.size(), l.data(), occ.data(), /*potential=*/1, &dE));
    return dE;
  }
};

class B200SemiGrandCanonicalCalculator : public b200::DelegatingCalculator {
 public:
  B200SemiGrandCanonicalCalculator()
      : DelegatingCalculator(&make_SemiGrandCanonicalCalculator, "B200SemiGrandCanonicalCalculator",
                            {},                      // required_basis_set
                            {},                      // required_local_basis_set
                            {"formation_energy"},    // required_clex
                            {}, {}, {}, {},          // multiclex, local clex, local multiclex, dof spaces
                            {"cmx_tables"},          // required_params
                            {"cmx_device", "cmx_seed", "cmx_reference_order"},  // optional_params
                            /*time_sampling_allowed=*/false, /*update_atoms=*/false, /*save_atom_info=*/false,
                            /*is_multistate_method=*/false) {}

  /// any integer transformation matrix of positive volume (diag(N0, N1, N2) boxes run the
  /// row kernels, skewed ones the general-supercell path of the library)
  Validator validate_configuration(state_type &state) const override {
    Validator v;
    Eigen::Matrix3l const &T = get_transformation_matrix_to_super(state);
    const long det = b200::determinant(T);
    if (det <= 0) v.error.insert("B200SemiGrandCanonicalCalculator: transformation_matrix_to_super must have a positive determinant");
    return v;
  }
  Validator validate_conditions(state_type &state) const override {  // SemiGrandCanonicalCalculator.cc:362-383
    Validator v;
    if (!state.conditions.scalar_values.count("temperature")) v.error.insert("Missing required condition: temperature");
    if (!state.conditions.vector_values.count("param_chem_pot"))
      v.error.insert("Missing required condition: param_chem_pot");
    return v;
  }
  Validator validate_state(state_type &state) const override {
    Validator v = this->validate_configuration(state);
    Validator c = this->validate_conditions(state);
    v.error.insert(c.error.begin(), c.error.end());
    return v;
  }

  /// StateData + potential for `state` (SemiGrandCanonicalCalculator.cc:386-414): here also the
  /// device state -- created once per supercell shape, refreshed from the host occupation
  void set_state_and_potential(state_type &state, monte::OccLocation *occ_location) override {
    if (this->system == nullptr)
      throw std::runtime_error("Error in B200SemiGrandCanonicalCalculator::run: system==nullptr");
    Validator v = this->validate_state(state);
    if (!v.valid()) throw std::runtime_error("Error in B200SemiGrandCanonicalCalculator::run: " + *v.error.begin());
    this->state_data = std::make_shared<StateData>(this->system, &state, occ_location);
    b200::bind_state(*m_dev, *this->state_data, *this->system, m_max_occ, "B200SemiGrandCanonicalCalculator");
    std::vector<double> exch =
        exchange_table(*this->system, state.conditions.vector_values.at("param_chem_pot"), m_max_occ);
    cmx_check(cmx_state_set_conditions(m_dev->state, 0, state.conditions.scalar_values.at("temperature"), exch.data()));
    this->potential = std::make_shared<B200SemiGrandCanonicalPotential>(this->state_data, m_dev, m_max_occ);
  }

  void set_event_data() override {
    throw std::runtime_error("Error in B200SemiGrandCanonicalCalculator::set_event_data: not valid");
  }

  /// One run at fixed conditions: occupation_metropolis_v2 (methods/occupation_metropolis.hh:72-123)
  /// with the steps between two sampling points executed on the device.  A pass is
  /// `steps_per_pass = occ_location.mol_size()` attempted steps (:83) = one checkerboard sweep.
  void run(state_type &state, monte::OccLocation &occ_location,
           run_manager_type<engine_type> &run_manager) override {
    this->set_state_and_potential(state, &occ_location);
    if (run_manager.engine == nullptr)
      throw std::runtime_error("Error in B200SemiGrandCanonicalCalculator::run: run_manager.engine==nullptr");
    this->engine = run_manager.engine;
    // one draw of the run's engine seeds the device generator (the reference draws every
    // random number of the run from it, CanonicalCalculator.cc:419-425)
    const uint64_t seed = params.contains("cmx_seed") ? (uint64_t)params.get_number("cmx_seed") : (uint64_t)(*this->engine)();
    const bool reference_order = params.contains("cmx_reference_order") && params.get_number("cmx_reference_order") != 0.0;
    const Index steps_per_pass = occ_location.mol_size();
    Eigen::VectorXi &occupation = get_occupation(state);
    int64_t pass = 0;

    run_manager.initialize(steps_per_pass);
    run_manager.sample_data_by_count_if_due(state);
    while (!run_manager.is_complete()) {
      run_manager.write_status_if_due();
      const Index n_passes = std::max<Index>(1, run_manager.passes_until_sample_due());
      int64_t n_attempt = 0, n_accept = 0;
      if (reference_order) {
        // the reference's proposals, random numbers and acceptance, step by step, on the device
        int64_t acc = 0;
        uint64_t hash = 0;
        cmx_check(cmx_metropolis_sequential(m_dev->state, 0, /*mode: semi-grand*/ 0, n_passes * steps_per_pass,
                                            seed + (uint64_t)pass, nullptr, 0, &acc, &hash));
        n_attempt = n_passes * steps_per_pass;
        n_accept = acc;
      } else {
        cmx_counters c;
        cmx_check(cmx_sgc_sweep(m_dev->state, n_passes, seed, pass, &c));  // the hot loop
        n_attempt = c.n_attempt;
        n_accept = c.n_accept;
      }
      pass += n_passes;
      // RunManager counters (occupation_metropolis.hh:109-116)
      b200::count_steps(run_manager, n_passes, n_attempt, n_accept);
      // the sampling functions read the host state (sampling_functions.cc): bring it up to date
      cmx_check(cmx_state_download_occ(m_dev->state, 0, occupation.data()));
      run_manager.sample_data_by_count_if_due(state);
    }
    // the occupation changed behind OccLocation's back: rebuild its lists (BaseMonteCalculator.hh:222-228)
    occ_location.initialize(occupation);
    run_manager.finalize(state);
  }

  void run(int, std::vector<state_type> &, std::vector<monte::OccLocation> &,
           run_manager_type<engine_type> &) override {
    throw std::runtime_error("Error: B200SemiGrandCanonicalCalculator does not allow multi-state runs");
  }

 private:
  std::shared_ptr<DeviceState> m_dev = std::make_shared<DeviceState>();
  int m_max_occ = 0;

  /// after `params` / `system` changed (BaseMonteCalculator.hh:107-113): load the tables
  void _reset() override {
    m_dev = std::make_shared<DeviceState>();
    b200::load_tables(*m_dev, params, "B200SemiGrandCanonicalCalculator");
    m_max_occ = b200::max_occupants(*this->system);
  }

  /// deep copy: the clone gets its own device handles (re-created by reset / set_state_and_potential)
  B200SemiGrandCanonicalCalculator *_clone() const override {
    auto *c = new B200SemiGrandCanonicalCalculator();
    c->params = this->params;
    c->system = this->system;
    c->engine = this->engine;
    if (this->system) c->_reset();
    return c;
  }
};

}  // namespace clexmonte
}  // namespace CASM

extern "C" {
/// \brief Returns a clexmonte::BaseMonteCalculator* owning a B200SemiGrandCanonicalCalculator
CASM::clexmonte::BaseMonteCalculator *make_B200SemiGrandCanonicalCalculator() {
  return new CASM::clexmonte::B200SemiGrandCanonicalCalculator();
}
}
