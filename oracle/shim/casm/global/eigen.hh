// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Stand-in for casm/global/eigen.hh (libcasm-global, not vendored under
// /root/reference).  The CASM-generated Clexulator sources only need a 3x3
// integer matrix that supports `m.row(i) << a, b, c;`
// (e.g. FCC_binary_vacancy_Clexulator_default.cc:374-376).
#ifndef ORACLE_SHIM_CASM_GLOBAL_EIGEN_HH
#define ORACLE_SHIM_CASM_GLOBAL_EIGEN_HH

namespace Eigen {

class Matrix3l {
 public:
  long v[3][3];
  Matrix3l() {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) v[i][j] = 0;
  }

  class RowInit {
   public:
    RowInit(long *row, int pos) : m_row(row), m_pos(pos) {}
    RowInit operator,(long x) {
      m_row[m_pos] = x;
      return RowInit(m_row, m_pos + 1);
    }

   private:
    long *m_row;
    int m_pos;
  };

  class Row {
   public:
    explicit Row(long *row) : m_row(row) {}
    RowInit operator<<(long x) {
      m_row[0] = x;
      return RowInit(m_row, 1);
    }

   private:
    long *m_row;
  };

  Row row(int i) { return Row(v[i]); }
  long operator()(int i, int j) const { return v[i][j]; }
};

}  // namespace Eigen

#endif
