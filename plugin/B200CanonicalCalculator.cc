// B200CanonicalCalculator -- the reference-side plugin of the B200 hot path for the canonical
// ensemble (BASELINE configs[0], [3]): the same calculator as the reference's
// (src/casm/clexmonte/monte_calculator/CanonicalCalculator.cc:171-470; factory :472-478)
// whose potential and whose Metropolis loop run on the GPU through the C ABI of
// include/cmx_b200.h.  Loaded like every method of CASM clexmonte: a shared library exporting
//   extern "C" CASM::clexmonte::BaseMonteCalculator *make_B200CanonicalCalculator();
// (make_monte_calculator_from_source, monte_calculator/MonteCalculator.cc:142-173).
//
// The reference proposes one swap of two unlike-species sites anywhere in the supercell per
// step (monte::propose_canonical_event [EXT] over the canonical swap table of
// system/System.cc:55-58).  Here a pass is one sweep of parallel pair exchanges over the
// library's default swap table (cmx_canonical_default_swaps: the shortest translations of the
// prim neighbor list plus one translation across the box per sublattice pair, conflict-free
// colours updated simultaneously): composition is conserved exactly, like-species pairs are
// skipped and not counted, the same ensemble is sampled (3-sigma tests against the reference
// loop: tests/test_gpu_canonical.py).  `cmx_reference_order` != 0 runs the reference's own
// proposal order step by step instead (cmx_metropolis_sequential, mode 1).
//
// Calculator params: "cmx_tables" (flat table export of the formation_energy basis set),
// "cmx_device", "cmx_seed", "cmx_reference_order", "cmx_swap_shell" (translations per
// sublattice pair, default 12), "cmx_swap_long_range" (default 1).
#include "b200_common.hh"

extern "C" {
/// the reference's calculator of this ensemble (libcasm_clexmonte, CanonicalCalculator.cc:472-478):
/// source of the standard sampling / analysis / state-modifying functions
CASM::clexmonte::BaseMonteCalculator *make_CanonicalCalculator();
}

namespace CASM {
namespace clexmonte {

namespace {
using b200::DeviceState;
const char *const kName = "B200CanonicalCalculator";
void cmx_check(int rc) { b200::cmx_check(rc, kName); }
}  // namespace

/// CanonicalPotential (CanonicalCalculator.cc:96-141) on the device
class B200CanonicalPotential : public BaseMontePotential {
 public:
  B200CanonicalPotential(std::shared_ptr<StateData> _state_data, std::shared_ptr<DeviceState> _dev)
      : BaseMontePotential(_state_data), dev(_dev), n_unitcells(_state_data->n_unitcells) {}

  std::shared_ptr<DeviceState> dev;
  Index n_unitcells;

  /// formation_energy_clex->per_supercell() (:126-128)
  double per_supercell() override {
    double E = 0.0;
    cmx_check(cmx_energy(dev->state, 0, &E));
    return E;
  }
  double per_unitcell() override { return this->per_supercell() / n_unitcells; }

  /// formation_energy_clex->occ_delta_value(linear_site_index, new_occ) (:137-140): the sites
  /// of the event are changed one after the other
  double occ_delta_per_supercell(std::vector<Index> const &linear_site_index,
                                 std::vector<int> const &new_occ) override {
    std::vector<int64_t> l(linear_site_index.size());
    for (size_t q = 0; q < l.size(); ++q) l[q] = dev->to_library(linear_site_index[q]);
    std::vector<int32_t> occ(new_occ.begin(), new_occ.end());
    double dE = 0.0;
    cmx_check(cmx_delta_e(dev->state, 0, 1, (int32_t)l.size(), l.data(), occ.data(), /*potential=*/0, &dE));
    return dE;
  }
};

class B200CanonicalCalculator : public b200::DelegatingCalculator {
 public:
  B200CanonicalCalculator()
      : DelegatingCalculator(&make_CanonicalCalculator, kName,
                            {},                    // required_basis_set
                            {},                    // required_local_basis_set
                            {"formation_energy"},  // required_clex
                            {}, {}, {}, {},        // multiclex, local clex, local multiclex, dof spaces
                            {"cmx_tables"},        // required_params
                            {"cmx_device", "cmx_seed", "cmx_reference_order", "cmx_swap_shell",
                             "cmx_swap_long_range"},  // optional_params
                            /*time_sampling_allowed=*/false, /*update_atoms=*/false, /*save_atom_info=*/false,
                            /*is_multistate_method=*/false) {}

  /// CanonicalCalculator.cc:282-287: every configuration is valid; here the supercell must
  /// have positive volume
  Validator validate_configuration(state_type &state) const override {
    Validator v;
    if (b200::determinant(get_transformation_matrix_to_super(state)) <= 0)
      v.error.insert(std::string(kName) + ": transformation_matrix_to_super must have a positive determinant");
    return v;
  }
  /// :296-315: scalar temperature required; mol_composition / param_composition optional
  Validator validate_conditions(state_type &state) const override {
    Validator v;
    if (!state.conditions.scalar_values.count("temperature")) v.error.insert("Missing required condition: temperature");
    for (auto const &kv : state.conditions.vector_values)
      if (kv.first != "param_composition" && kv.first != "mol_composition")
        v.error.insert("Unknown vector condition: " + kv.first);
    return v;
  }
  /// :318-360: the configuration's composition must be the conditions' (when they name one)
  Validator validate_state(state_type &state) const override {
    Validator v = this->validate_configuration(state);
    Validator c = this->validate_conditions(state);
    v.error.insert(c.error.begin(), c.error.end());
    if (!v.valid() || this->system == nullptr || !state.conditions.vector_values.count("mol_composition")) return v;
    Eigen::VectorXd const &want = state.conditions.vector_values.at("mol_composition");
    Eigen::VectorXi const &occupation = get_occupation(state);
    const size_t n_sublat = this->system->occ_to_species.size();
    const Index n_cells = occupation.size() / (Index)n_sublat;
    std::vector<double> mol((size_t)want.size(), 0.0);
    if (n_cells == 0 || occupation.size() != n_cells * (Index)n_sublat) {
      v.error.insert("Error: the occupation does not fit the system's sublattices");
      return v;
    }
    for (Index l = 0; l < occupation.size(); ++l) {
      auto const &allowed = this->system->occ_to_species[(size_t)(l / n_cells)];
      if (occupation[l] < 0 || (size_t)occupation[l] >= allowed.size()) {
        v.error.insert("Error: occupant index out of range");
        return v;
      }
      const Index species = allowed[(size_t)occupation[l]];
      if (species >= 0 && species < (Index)mol.size()) mol[(size_t)species] += 1.0 / (double)n_cells;
    }
    for (long q = 0; q < want.size(); ++q)
      if (std::fabs(mol[(size_t)q] - want[q]) > this->mol_composition_tol)
        v.error.insert("Error: the configuration's composition is not the conditions' mol_composition");
    return v;
  }
  double mol_composition_tol = 1e-4;
  /// unlike-species pairs the device evaluated / accepted in the last run (exact counts)
  int64_t n_pairs_attempted = 0, n_pairs_accepted = 0;

  /// StateData + potential for `state` (CanonicalCalculator.cc:363-392), and the device state
  void set_state_and_potential(state_type &state, monte::OccLocation *occ_location) override {
    if (this->system == nullptr) throw std::runtime_error(std::string("Error in ") + kName + "::run: system==nullptr");
    Validator v = this->validate_state(state);
    if (!v.valid()) throw std::runtime_error(std::string("Error in ") + kName + "::run: " + *v.error.begin());
    this->state_data = std::make_shared<StateData>(this->system, &state, occ_location);
    const bool fresh = b200::bind_state(*m_dev, *this->state_data, *this->system, m_max_occ, kName);
    cmx_check(cmx_state_set_conditions(m_dev->state, 0, state.conditions.scalar_values.at("temperature"), nullptr));
    if (fresh) m_swaps_set = false;
    this->potential = std::make_shared<B200CanonicalPotential>(this->state_data, m_dev);
  }

  void set_event_data() override {
    throw std::runtime_error(std::string("Error in ") + kName + "::set_event_data: not valid");
  }

  /// One run at fixed conditions: occupation_metropolis_v2 (methods/occupation_metropolis.hh:72-123)
  /// with the steps between two sampling points executed on the device (CanonicalCalculator.cc:394-440).
  void run(state_type &state, monte::OccLocation &occ_location,
           run_manager_type<engine_type> &run_manager) override {
    this->set_state_and_potential(state, &occ_location);
    if (run_manager.engine == nullptr)
      throw std::runtime_error(std::string("Error in ") + kName + "::run: run_manager.engine==nullptr");
    this->engine = run_manager.engine;
    const uint64_t seed = params.contains("cmx_seed") ? (uint64_t)params.get_number("cmx_seed") : (uint64_t)(*this->engine)();
    const bool reference_order = params.contains("cmx_reference_order") && params.get_number("cmx_reference_order") != 0.0;
    if (!reference_order && !m_swaps_set) {
      const int32_t n_shell = params.contains("cmx_swap_shell") ? (int32_t)params.get_number("cmx_swap_shell") : 12;
      const int32_t long_range = params.contains("cmx_swap_long_range") ? (int32_t)params.get_number("cmx_swap_long_range") : 1;
      int32_t n = 0;
      cmx_check(cmx_canonical_default_swaps(m_dev->state, n_shell, long_range, 0, nullptr, &n));
      std::vector<cmx_swap_type> swaps((size_t)std::max<int32_t>(n, 1));
      cmx_check(cmx_canonical_default_swaps(m_dev->state, n_shell, long_range, n, swaps.data(), &n));
      cmx_check(cmx_canonical_set_swaps(m_dev->state, n, swaps.data()));
      m_swaps_set = true;
    }
    const Index steps_per_pass = occ_location.mol_size();
    Eigen::VectorXi &occupation = get_occupation(state);
    int64_t pass = 0;
    n_pairs_attempted = n_pairs_accepted = 0;

    run_manager.initialize(steps_per_pass);
    run_manager.sample_data_by_count_if_due(state);
    while (!run_manager.is_complete()) {
      run_manager.write_status_if_due();
      const Index n_passes = std::max<Index>(1, run_manager.passes_until_sample_due());
      int64_t n_attempt = 0, n_accept = 0;
      if (reference_order) {
        int64_t acc = 0;
        uint64_t hash = 0;
        cmx_check(cmx_metropolis_sequential(m_dev->state, 0, /*mode: canonical*/ 1, n_passes * steps_per_pass,
                                            seed + (uint64_t)pass, nullptr, 0, &acc, &hash));
        n_attempt = n_passes * steps_per_pass;
        n_accept = acc;
      } else {
        // A pass of this calculator is one sweep over the swap table.  It evaluates as many
        // unlike-species pairs as the configuration holds (not steps_per_pass of them), so the
        // RunManager gets the sweep's acceptance RATE scaled to the steps_per_pass steps of a
        // pass: its pass count, and with it the sampling schedule, stays exact.
        cmx_counters c;
        cmx_check(cmx_canonical_sweep(m_dev->state, n_passes, seed, pass, &c));  // the hot loop
        n_pairs_attempted += c.n_attempt;
        n_pairs_accepted += c.n_accept;
        n_attempt = n_passes * steps_per_pass;
        n_accept = c.n_attempt > 0 ? (int64_t)(((__int128)c.n_accept * n_attempt + c.n_attempt / 2) / c.n_attempt) : 0;
      }
      pass += n_passes;
      b200::count_steps(run_manager, n_passes, n_attempt, n_accept);
      cmx_check(cmx_state_download_occ(m_dev->state, 0, occupation.data()));
      run_manager.sample_data_by_count_if_due(state);
    }
    occ_location.initialize(occupation);  // BaseMonteCalculator.hh:222-228
    run_manager.finalize(state);
  }

  void run(int, std::vector<state_type> &, std::vector<monte::OccLocation> &,
           run_manager_type<engine_type> &) override {
    throw std::runtime_error(std::string("Error: ") + kName + " does not allow multi-state runs");
  }

 private:
  std::shared_ptr<DeviceState> m_dev = std::make_shared<DeviceState>();
  int m_max_occ = 0;
  bool m_swaps_set = false;

  void _reset() override {
    m_dev = std::make_shared<DeviceState>();
    m_swaps_set = false;
    b200::load_tables(*m_dev, params, kName);
    m_max_occ = b200::max_occupants(*this->system);
  }

  B200CanonicalCalculator *_clone() const override {
    auto *c = new B200CanonicalCalculator();
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
/// \brief Returns a clexmonte::BaseMonteCalculator* owning a B200CanonicalCalculator
CASM::clexmonte::BaseMonteCalculator *make_B200CanonicalCalculator() {
  return new CASM::clexmonte::B200CanonicalCalculator();
}
}
