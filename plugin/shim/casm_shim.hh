// Stand-in declarations for the part of libcasm the B200 plugin touches (TEST INFRASTRUCTURE).
//
// plugin/B200SemiGrandCanonicalCalculator.cc is written against the reference's plugin
// interface (include/casm/clexmonte/monte_calculator/BaseMonteCalculator.hh:26-264) and its
// neighbours.  Those headers only compile with libcasm-{global,crystallography,clexulator,
// configuration,composition,monte} installed, none of which exist in this image (SURVEY.md
// 0-1).  So that the plugin source is nevertheless COMPILED and DRIVEN in CI, this header
// declares, under the reference's own names and signatures, exactly the members the plugin
// uses -- nothing else -- each with the reference (or [EXT] library) location it mirrors.
// With libcasm installed the forwarding headers next to this file are not on the include
// path and the plugin compiles against the real ones.
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <random>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

// ---- [EXT] Eigen: only what the plugin calls ---------------------------------------------
namespace Eigen {
template <typename T>
struct VectorT {
  std::vector<T> v;
  VectorT() {}
  explicit VectorT(long n) : v((size_t)n, T()) {}
  long size() const { return (long)v.size(); }
  T *data() { return v.data(); }
  T const *data() const { return v.data(); }
  T &operator()(long i) { return v[(size_t)i]; }
  T const &operator()(long i) const { return v[(size_t)i]; }
  T &operator[](long i) { return v[(size_t)i]; }
  T const &operator[](long i) const { return v[(size_t)i]; }
  void setZero() { std::fill(v.begin(), v.end(), T()); }
  void resize(long n) { v.assign((size_t)n, T()); }
  T dot(VectorT const &o) const {
    T s = T();
    for (size_t i = 0; i < v.size(); ++i) s += v[i] * o.v[i];
    return s;
  }
};
typedef VectorT<int> VectorXi;
typedef VectorT<double> VectorXd;
struct MatrixXd {
  long r = 0, c = 0;
  std::vector<double> v;
  MatrixXd() {}
  MatrixXd(long rows, long cols) : r(rows), c(cols), v((size_t)(rows * cols), 0.0) {}
  long rows() const { return r; }
  long cols() const { return c; }
  double &operator()(long i, long j) { return v[(size_t)(i * c + j)]; }
  double operator()(long i, long j) const { return v[(size_t)(i * c + j)]; }
  VectorXd operator*(VectorXd const &x) const {
    VectorXd y(r);
    for (long i = 0; i < r; ++i)
      for (long j = 0; j < c; ++j) y(i) += (*this)(i, j) * x(j);
    return y;
  }
};
struct Matrix3l {
  long m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  long &operator()(int i, int j) { return m[i][j]; }
  long operator()(int i, int j) const { return m[i][j]; }
};
}  // namespace Eigen

namespace CASM {
typedef long Index;  // [EXT] casm/global/definitions.hh

// [EXT] casm/casm_io/json/jsonParser.hh: object lookup, get<T>
class jsonParser {
 public:
  std::map<std::string, std::string> strings;
  std::map<std::string, double> numbers;
  bool contains(std::string const &key) const { return strings.count(key) || numbers.count(key); }
  std::string get_string(std::string const &key) const { return strings.at(key); }
  double get_number(std::string const &key) const { return numbers.at(key); }
};

// casm/misc/Validator.hh [EXT libcasm-global]: error / warning sets
struct Validator {
  std::set<std::string> error, warning;
  bool valid() const { return error.empty(); }
};

namespace composition {
// [EXT] casm/composition/CompositionConverter.hh: x = R^T (n - n_0)
class CompositionConverter {
 public:
  std::vector<std::string> m_components;
  Eigen::VectorXd m_origin;     // n_0 [n_components]
  Eigen::MatrixXd m_Rt;         // dparam_dmol [k][n_components]
  std::vector<std::string> components() const { return m_components; }
  Index independent_compositions() const { return m_Rt.rows(); }
  Eigen::MatrixXd dparam_dmol() const { return m_Rt; }
  Eigen::VectorXd origin() const { return m_origin; }
  Eigen::VectorXd param_composition(Eigen::VectorXd const &n) const {
    Eigen::VectorXd d(n.size());
    for (long i = 0; i < n.size(); ++i) d(i) = n(i) - m_origin(i);
    return m_Rt * d;
  }
};
}  // namespace composition

namespace clexulator {
// [EXT] casm/clexulator/SparseCoefficients.hh
struct SparseCoefficients {
  std::vector<unsigned int> index;
  std::vector<double> value;
};
}  // namespace clexulator

namespace xtal {
// [EXT] casm/crystallography/UnitCellCoord.hh: xtal::UnitCell, integer prim coordinates of a unit cell
struct UnitCell {
  long v[3] = {0, 0, 0};
  long operator[](int a) const { return v[a]; }
  long operator()(int a) const { return v[a]; }
};
}  // namespace xtal

namespace monte {
// [EXT] casm/monte/Conversions.hh: the index conversions the potential uses
// (SemiGrandCanonicalCalculator.cc:202-209).  l = b * n_unitcells + unitl; the ORDER of the
// unit cells within a supercell is the reference's business (xtal::UnitCellIndexConverter,
// Smith normal form): this stand-in deliberately uses one of its own (see StateData), so a
// plugin that guessed the order instead of asking l_to_ijk fails the driver's checks.
class Conversions {
 public:
  Index m_n_unitcells = 0;
  std::vector<Index> m_b_to_asym;
  std::vector<std::vector<Index>> m_species_index;  // [asym][occ]
  std::vector<xtal::UnitCell> m_unitl_to_ijk;       // [unitl]
  xtal::UnitCell l_to_ijk(Index l) const { return m_unitl_to_ijk[(size_t)(l % m_n_unitcells)]; }
  Index l_to_unitl(Index l) const { return l % m_n_unitcells; }
  Index l_to_b(Index l) const { return l / m_n_unitcells; }
  Index l_to_asym(Index l) const { return m_b_to_asym[(size_t)(l / m_n_unitcells)]; }
  Index species_index(Index asym, Index occ) const { return m_species_index[(size_t)asym][(size_t)occ]; }
};

// [EXT] casm/monte/events/OccLocation.hh
class OccLocation {
 public:
  Index m_mol_size = 0;
  int n_initialize = 0;
  Index mol_size() const { return m_mol_size; }
  void initialize(Eigen::VectorXi const &occupation) {
    (void)occupation;
    ++n_initialize;
  }
};

// [EXT] casm/monte/ValueMap.hh
struct ValueMap {
  std::map<std::string, double> scalar_values;
  std::map<std::string, Eigen::VectorXd> vector_values;
};

// [EXT] casm/monte/State.hh
template <typename ConfigType>
struct State {
  ConfigType configuration;
  ValueMap conditions;
};

// [EXT] casm/monte/run_management/RunManager.hh -- the calls of
// methods/occupation_metropolis.hh:92-120, plus the bulk form INTEGRATION.md proposes
template <typename ConfigType, typename StatisticsType, typename EngineType>
class RunManager {
 public:
  std::shared_ptr<EngineType> engine;
  // stand-in sampling fixture: one sample every `sample_period` passes, complete after
  // `n_samples_max` samples; `sampler` is what the fixture's sampling functions would do
  Index sample_period = 1, n_samples_max = 1;
  std::function<void(State<ConfigType> const &)> sampler;
  Index steps_per_pass = 0, step = 0, pass = 0, n_accept = 0, n_reject = 0, n_samples = 0;
  bool initialized = false, finalized = false;
  Index last_sampled_pass = -1;

  void initialize(Index _steps_per_pass) {
    steps_per_pass = _steps_per_pass;
    step = pass = n_accept = n_reject = n_samples = 0;
    last_sampled_pass = -1;
    initialized = true;
  }
  bool is_complete() const { return n_samples >= n_samples_max; }
  void write_status_if_due() {}
  void increment_n_accept() { ++n_accept; }
  void increment_n_reject() { ++n_reject; }
  void increment_step() {
    if (++step == steps_per_pass) {
      step = 0;
      ++pass;
    }
  }
#ifdef CMX_HAVE_RUNMANAGER_BULK
  // proposed addition to libcasm-monte: whole passes at once
  void add_passes(Index n_passes, Index _n_accept, Index _n_reject) {
    pass += n_passes;
    n_accept += _n_accept;
    n_reject += _n_reject;
  }
#endif
  // passes until the sampling fixture wants the next sample (>= 1 unless one is due now)
  Index passes_until_sample_due() const {
    if (last_sampled_pass < 0) return 0;
    Index due = last_sampled_pass + sample_period;
    return due > pass ? due - pass : 0;
  }
  void sample_data_by_count_if_due(State<ConfigType> const &state) {
    if (step == 0 && passes_until_sample_due() == 0 && pass != last_sampled_pass) {
      if (sampler) sampler(state);
      last_sampled_pass = pass;
      ++n_samples;
    }
  }
  void finalize(State<ConfigType> const &state) {
    (void)state;
    finalized = true;
  }
};
}  // namespace monte

namespace clexmonte {
// include/casm/clexmonte/definitions.hh
typedef std::mt19937_64 default_engine_type;
struct Configuration {  // state/Configuration.hh: config::Configuration stand-in
  Eigen::Matrix3l transformation_matrix_to_super;
  struct {
    Eigen::VectorXi occupation;
  } dof_values;
};
typedef Configuration config_type;
typedef monte::State<config_type> state_type;
struct statistics_type {};
template <typename EngineType>
using run_manager_type = monte::RunManager<config_type, statistics_type, EngineType>;

inline Eigen::VectorXi &get_occupation(state_type &state) { return state.configuration.dof_values.occupation; }
inline Eigen::VectorXi const &get_occupation(state_type const &state) {
  return state.configuration.dof_values.occupation;
}
inline Eigen::Matrix3l const &get_transformation_matrix_to_super(state_type const &state) {
  return state.configuration.transformation_matrix_to_super;
}

// system/system_data.hh:27-46
struct ClexData {
  std::string basis_set_name;
  clexulator::SparseCoefficients coefficients;
};
// system/System.hh:24-240: only what the plugin reads
struct System {
  composition::CompositionConverter composition_converter;
  std::map<std::string, ClexData> clex_data;
  std::vector<Index> sublat_to_asym;                  // [EXT] via monte::Conversions
  std::vector<std::vector<Index>> occ_to_species;     // [sublattice][occupant] -> component index
};
typedef System system_type;
inline composition::CompositionConverter const &get_composition_converter(system_type const &s) {
  return s.composition_converter;
}
inline ClexData const &get_clex_data(system_type const &s, std::string const &key) {
  auto it = s.clex_data.find(key);
  if (it == s.clex_data.end()) throw std::runtime_error("System error: '" + key + "' is not a clex.");
  return it->second;
}

// monte_calculator/StateData.hh:14-58
struct StateData {
  StateData(std::shared_ptr<system_type> _system, state_type const *_state, monte::OccLocation const *_occ_location)
      : system(_system), state(_state), occ_location(_occ_location),
        transformation_matrix_to_super(get_transformation_matrix_to_super(*_state)) {
    auto const &T = transformation_matrix_to_super;
    n_unitcells = T(0, 0) * (T(1, 1) * T(2, 2) - T(1, 2) * T(2, 1)) - T(0, 1) * (T(1, 0) * T(2, 2) - T(1, 2) * T(2, 0)) +
                  T(0, 2) * (T(1, 0) * T(2, 1) - T(1, 1) * T(2, 0));
    if (n_unitcells < 0) n_unitcells = -n_unitcells;
    owned_convert.m_n_unitcells = n_unitcells;
    // the unit cells of the supercell: the lattice points whose coordinates in the supercell
    // lattice, T^-1 ijk = adj(T) ijk / det, lie in [0, 1)^3; numbered k-major in prim coordinates
    {
      long A[3][3];  // adjugate of T
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          const int r1 = (c + 1) % 3, r2 = (c + 2) % 3, c1 = (r + 1) % 3, c2 = (r + 2) % 3;
          A[r][c] = T(r1, c1) * T(r2, c2) - T(r1, c2) * T(r2, c1);
        }
      long det = 0;
      for (int c = 0; c < 3; ++c) det += T(0, c) * A[c][0];
      long lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};  // bounding box of the supercell's corners
      for (int m = 0; m < 8; ++m)
        for (int r = 0; r < 3; ++r) {
          long x = 0;
          for (int c = 0; c < 3; ++c) x += ((m >> c) & 1) ? T(r, c) : 0;
          lo[r] = std::min(lo[r], x);
          hi[r] = std::max(hi[r], x);
        }
      for (long k = lo[2]; k <= hi[2]; ++k)
        for (long j = lo[1]; j <= hi[1]; ++j)
          for (long i = lo[0]; i <= hi[0]; ++i) {
            bool inside = true;
            for (int r = 0; r < 3 && inside; ++r) {
              long f = A[r][0] * i + A[r][1] * j + A[r][2] * k;
              if (det < 0) f = -f;
              inside = f >= 0 && f < (det < 0 ? -det : det);
            }
            if (!inside) continue;
            xtal::UnitCell u;
            u.v[0] = i;
            u.v[1] = j;
            u.v[2] = k;
            owned_convert.m_unitl_to_ijk.push_back(u);
          }
      if ((Index)owned_convert.m_unitl_to_ijk.size() != n_unitcells)
        throw std::runtime_error("casm_shim: unit cell enumeration does not match det(T)");
    }
    owned_convert.m_b_to_asym = _system->sublat_to_asym;
    owned_convert.m_species_index.assign(_system->occ_to_species.size(), {});
    for (size_t b = 0; b < _system->occ_to_species.size(); ++b)
      owned_convert.m_species_index[(size_t)_system->sublat_to_asym[b]] = _system->occ_to_species[b];
    convert = &owned_convert;
  }
  std::shared_ptr<system_type> system;
  state_type const *state;
  monte::OccLocation const *occ_location;
  Eigen::Matrix3l transformation_matrix_to_super;
  Index n_unitcells;
  monte::Conversions const *convert;
  monte::Conversions owned_convert;
};

class MonteCalculator;

// monte_calculator/BaseMonteCalculator.hh:26-44
class BaseMontePotential {
 public:
  BaseMontePotential(std::shared_ptr<StateData> _state_data) : state_data(_state_data) {}
  virtual ~BaseMontePotential() {}
  std::shared_ptr<StateData> state_data;
  virtual double per_supercell() = 0;
  virtual double per_unitcell() = 0;
  virtual double occ_delta_per_supercell(std::vector<Index> const &linear_site_index,
                                         std::vector<int> const &new_occ) = 0;
};

// [EXT]-typed values of the sampling / analysis / fixture interface (libcasm-monte:
// StateSamplingFunction, jsonStateSamplingFunction, ResultsAnalysisFunction,
// StateModifyingFunction, SelectedEventFunctions, SamplingFixtureParams; typedefs in
// include/casm/clexmonte/definitions.hh).  Opaque here: a plugin only passes them through.
struct state_sampling_function_type {
  std::string name;
};
struct json_state_sampling_function_type {
  std::string name;
};
struct results_analysis_function_type {
  std::string name;
};
struct StateModifyingFunction {
  std::string name;
};
typedef std::map<std::string, StateModifyingFunction> StateModifyingFunctionMap;
struct sampling_fixture_params_type {
  std::string label;
  std::vector<std::string> sampler_names;
};
}  // namespace clexmonte
namespace monte {
struct SelectedEventFunctions {};
}  // namespace monte
namespace clexmonte {

// monte_calculator/BaseMonteCalculator.hh:47-264 (the selected-event and KMC data members are
// [EXT]-typed and not used by a Metropolis plugin: left out here)
class BaseMonteCalculator {
 public:
  typedef default_engine_type engine_type;
  explicit BaseMonteCalculator(std::string _calculator_name, std::set<std::string> _required_basis_set,
                               std::set<std::string> _required_local_basis_set, std::set<std::string> _required_clex,
                               std::set<std::string> _required_multiclex, std::set<std::string> _required_local_clex,
                               std::set<std::string> _required_local_multiclex,
                               std::set<std::string> _required_dof_spaces, std::set<std::string> _required_params,
                               std::set<std::string> _optional_params, bool _time_sampling_allowed, bool _update_atoms,
                               bool _save_atom_info, bool _is_multistate_method)
      : calculator_name(_calculator_name), required_basis_set(_required_basis_set),
        required_local_basis_set(_required_local_basis_set), required_clex(_required_clex),
        required_multiclex(_required_multiclex), required_local_clex(_required_local_clex),
        required_local_multiclex(_required_local_multiclex), required_dof_spaces(_required_dof_spaces),
        required_params(_required_params), optional_params(_optional_params),
        time_sampling_allowed(_time_sampling_allowed), update_atoms(_update_atoms), save_atom_info(_save_atom_info),
        is_multistate_method(_is_multistate_method) {}
  virtual ~BaseMonteCalculator() {}
  std::shared_ptr<engine_type> engine;
  std::string calculator_name;
  std::set<std::string> required_basis_set, required_local_basis_set, required_clex, required_multiclex,
      required_local_clex, required_local_multiclex, required_dof_spaces, required_params, optional_params;
  bool time_sampling_allowed, update_atoms, save_atom_info;
  jsonParser params;
  std::shared_ptr<system_type> system;
  void reset(jsonParser const &_params, std::shared_ptr<system_type> _system) {  // :107-113
    this->params = _params;
    this->system = _system;
    for (auto const &key : required_clex) get_clex_data(*system, key);           // _check_system
    for (auto const &key : required_params)                                     // _check_params
      if (!params.contains(key)) throw std::runtime_error("Error: missing required parameter '" + key + "'");
    this->_reset();
  }
  // :121-151 -- the sampling / analysis / fixture maps every calculator provides
  virtual std::map<std::string, state_sampling_function_type> standard_sampling_functions(
      std::shared_ptr<MonteCalculator> const &calculation) const = 0;
  virtual std::map<std::string, json_state_sampling_function_type> standard_json_sampling_functions(
      std::shared_ptr<MonteCalculator> const &calculation) const = 0;
  virtual std::map<std::string, results_analysis_function_type> standard_analysis_functions(
      std::shared_ptr<MonteCalculator> const &calculation) const = 0;
  virtual StateModifyingFunctionMap standard_modifying_functions(
      std::shared_ptr<MonteCalculator> const &calculation) const = 0;
  virtual std::optional<monte::SelectedEventFunctions> standard_selected_event_functions(
      std::shared_ptr<MonteCalculator> const &calculation) const = 0;
  virtual sampling_fixture_params_type make_default_sampling_fixture_params(
      std::shared_ptr<MonteCalculator> const &calculation, std::string label, bool write_results,
      bool write_trajectory, bool write_observations, bool write_status, std::optional<std::string> output_dir,
      std::optional<std::string> log_file, double log_frequency_in_s) const = 0;
  virtual Validator validate_configuration(state_type &state) const = 0;
  virtual Validator validate_conditions(state_type &state) const = 0;
  virtual Validator validate_state(state_type &state) const = 0;
  std::shared_ptr<StateData> state_data;
  std::shared_ptr<BaseMontePotential> potential;
  virtual void set_state_and_potential(state_type &state, monte::OccLocation *occ_location) = 0;
  virtual void set_event_data() = 0;
  virtual void run(state_type &state, monte::OccLocation &occ_location,
                   run_manager_type<engine_type> &run_manager) = 0;
  bool is_multistate_method;
  virtual void run(int current_state, std::vector<state_type> &states, std::vector<monte::OccLocation> &occ_locations,
                   run_manager_type<engine_type> &run_manager) = 0;
  std::unique_ptr<BaseMonteCalculator> clone() const { return std::unique_ptr<BaseMonteCalculator>(this->_clone()); }

 private:
  virtual void _reset() = 0;
  virtual BaseMonteCalculator *_clone() const = 0;
};

}  // namespace clexmonte
}  // namespace CASM
