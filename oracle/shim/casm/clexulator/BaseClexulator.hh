// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Stand-in for casm/clexulator/BaseClexulator.hh (libcasm-clexulator v3.0a1,
// not vendored under /root/reference).  It lets the CASM-*generated*
// Clexulator sources that ARE in the reference
// (tests/unit/clexmonte/data/**/*_Clexulator_*.cc) compile unmodified, so the
// reference's own arithmetic can be executed as the parity oracle.
//
// Only the surface the generated code touches is provided (signatures read off
// FCC_binary_vacancy_Clexulator_default.cc:182-267,306-439):
//   ctor (nlist_size, corr_size, n_point_corr), size_type, corr_size(),
//   _occ(n), the protected metadata members, and the private virtuals.
// The public calc_* entry points follow the [EXT] BaseClexulator contract:
// point the object at (occupation, neighbor-list row), then dispatch.
#ifndef ORACLE_SHIM_CASM_CLEXULATOR_BASECLEXULATOR_HH
#define ORACLE_SHIM_CASM_CLEXULATOR_BASECLEXULATOR_HH

#include <cstddef>
#include <set>
#include <vector>

#include "casm/clexulator/BasicClexParamPack.hh"
#include "casm/global/eigen.hh"

namespace CASM {
namespace xtal {

struct UnitCell {
  long c[3];
  UnitCell(long i, long j, long k) {
    c[0] = i;
    c[1] = j;
    c[2] = k;
  }
  bool operator<(UnitCell const &o) const {
    for (int a = 0; a < 3; ++a) {
      if (c[a] < o.c[a]) return true;
      if (c[a] > o.c[a]) return false;
    }
    return false;
  }
};

struct UnitCellCoord {
  long b;
  long c[3];
  UnitCellCoord(long _b, long i, long j, long k) : b(_b) {
    c[0] = i;
    c[1] = j;
    c[2] = k;
  }
  bool operator<(UnitCellCoord const &o) const {
    for (int a = 0; a < 3; ++a) {
      if (c[a] < o.c[a]) return true;
      if (c[a] > o.c[a]) return false;
    }
    return b < o.b;
  }
};

}  // namespace xtal

namespace clexulator {

class BaseClexulator {
 public:
  typedef unsigned int size_type;

  BaseClexulator(size_type nlist_size, size_type corr_size,
                 size_type n_point_corr)
      : m_nlist_size(nlist_size),
        m_corr_size(corr_size),
        m_n_point_corr(n_point_corr),
        m_n_sublattices(0),
        m_occ_ptr(nullptr),
        m_nlist_ptr(nullptr) {}
  virtual ~BaseClexulator() {}

  size_type nlist_size() const { return m_nlist_size; }
  size_type corr_size() const { return m_corr_size; }
  size_type n_point_corr() const { return m_n_point_corr; }
  size_type n_sublattices() const { return m_n_sublattices; }
  Eigen::Matrix3l const &weight_matrix() const { return m_weight_matrix; }
  std::set<int> const &sublat_indices() const { return m_sublat_indices; }
  std::set<xtal::UnitCell> const &neighborhood() const {
    return m_neighborhood;
  }
  std::vector<std::set<xtal::UnitCell>> const &orbit_neighborhood() const {
    return m_orbit_neighborhood;
  }
  std::vector<std::set<xtal::UnitCellCoord>> const &orbit_site_neighborhood()
      const {
    return m_orbit_site_neighborhood;
  }

  virtual ClexParamPack const &param_pack() const = 0;
  virtual ClexParamPack &param_pack() = 0;

  BaseClexulator *clone() const { return _clone(); }

  // -- public entry points: bind (occ, nlist row) then dispatch -------------
  void bind(int const *occ, long const *nlist_row) const {
    m_occ_ptr = occ;
    m_nlist_ptr = nlist_row;
  }
  void calc_global_corr_contribution(int const *occ, long const *nlist,
                                     double *corr) const {
    bind(occ, nlist);
    _calc_global_corr_contribution(corr);
  }
  void calc_restricted_global_corr_contribution(int const *occ,
                                                long const *nlist, double *corr,
                                                size_type const *ib,
                                                size_type const *ie) const {
    bind(occ, nlist);
    _calc_restricted_global_corr_contribution(corr, ib, ie);
  }
  void calc_point_corr(int const *occ, long const *nlist, int nlist_ind,
                       double *corr) const {
    bind(occ, nlist);
    _calc_point_corr(nlist_ind, corr);
  }
  void calc_restricted_point_corr(int const *occ, long const *nlist,
                                  int nlist_ind, double *corr,
                                  size_type const *ib,
                                  size_type const *ie) const {
    bind(occ, nlist);
    _calc_restricted_point_corr(nlist_ind, corr, ib, ie);
  }
  void calc_delta_point_corr(int const *occ, long const *nlist, int nlist_ind,
                             int occ_i, int occ_f, double *corr) const {
    bind(occ, nlist);
    _calc_delta_point_corr(nlist_ind, occ_i, occ_f, corr);
  }
  void calc_restricted_delta_point_corr(int const *occ, long const *nlist,
                                        int nlist_ind, int occ_i, int occ_f,
                                        double *corr, size_type const *ib,
                                        size_type const *ie) const {
    bind(occ, nlist);
    _calc_restricted_delta_point_corr(nlist_ind, occ_i, occ_f, corr, ib, ie);
  }

 protected:
  int const &_occ(int const &nlist_ind) const {
    return m_occ_ptr[m_nlist_ptr[nlist_ind]];
  }

  Eigen::Matrix3l m_weight_matrix;
  std::set<int> m_sublat_indices;
  size_type m_n_sublattices;
  std::set<xtal::UnitCell> m_neighborhood;
  std::vector<std::set<xtal::UnitCell>> m_orbit_neighborhood;
  std::vector<std::set<xtal::UnitCellCoord>> m_orbit_site_neighborhood;

 private:
  virtual BaseClexulator *_clone() const = 0;
  virtual void _calc_global_corr_contribution() const = 0;
  virtual void _calc_global_corr_contribution(double *corr_begin) const = 0;
  virtual void _calc_restricted_global_corr_contribution(
      size_type const *ind_list_begin,
      size_type const *ind_list_end) const = 0;
  virtual void _calc_restricted_global_corr_contribution(
      double *corr_begin, size_type const *ind_list_begin,
      size_type const *ind_list_end) const = 0;
  virtual void _calc_point_corr(int nlist_ind) const = 0;
  virtual void _calc_point_corr(int nlist_ind, double *corr_begin) const = 0;
  virtual void _calc_restricted_point_corr(
      int nlist_ind, size_type const *ind_list_begin,
      size_type const *ind_list_end) const = 0;
  virtual void _calc_restricted_point_corr(
      int nlist_ind, double *corr_begin, size_type const *ind_list_begin,
      size_type const *ind_list_end) const = 0;
  virtual void _calc_delta_point_corr(int nlist_ind, int occ_i,
                                      int occ_f) const = 0;
  virtual void _calc_delta_point_corr(int nlist_ind, int occ_i, int occ_f,
                                      double *corr_begin) const = 0;
  virtual void _calc_restricted_delta_point_corr(
      int nlist_ind, int occ_i, int occ_f, size_type const *ind_list_begin,
      size_type const *ind_list_end) const = 0;
  virtual void _calc_restricted_delta_point_corr(
      int nlist_ind, int occ_i, int occ_f, double *corr_begin,
      size_type const *ind_list_begin,
      size_type const *ind_list_end) const = 0;

  size_type m_nlist_size;
  size_type m_corr_size;
  size_type m_n_point_corr;
  mutable int const *m_occ_ptr;
  mutable long const *m_nlist_ptr;
};

}  // namespace clexulator
}  // namespace CASM

#endif
