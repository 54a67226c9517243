// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Stand-in for casm/clexulator/BasicClexParamPack.hh (libcasm-clexulator
// v3.0a1, not vendored under /root/reference).  Implements exactly the calls
// the generated Clexulator sources make (census in SURVEY.md section 0-3):
//   allocate / eval_mode / pre_eval / post_eval / read(key,i[,j]) /
//   Val<T>::set(pack,key,i[,j],v) / Val<T>::get(pack,key,i) / READ / Key
#ifndef ORACLE_SHIM_CASM_CLEXULATOR_BASICCLEXPARAMPACK_HH
#define ORACLE_SHIM_CASM_CLEXULATOR_BASICCLEXPARAMPACK_HH

#include <cstddef>
#include <string>
#include <vector>

namespace CASM {
namespace clexulator {

class ClexParamPack {
 public:
  virtual ~ClexParamPack() {}
};

class BasicClexParamPack : public ClexParamPack {
 public:
  typedef int Key;
  enum EvalMode { DEFAULT = 0, READ = 1 };

  struct Block {
    std::string name;
    int rows;
    int cols;
    EvalMode mode;
    std::vector<double> data;  // row-major: data[i * cols + j]
  };

  Key allocate(std::string const &name, int rows, int cols, bool /*indep*/) {
    Block b;
    b.name = name;
    b.rows = rows;
    b.cols = cols;
    b.mode = DEFAULT;
    b.data.assign(static_cast<std::size_t>(rows) * cols, 0.0);
    m_blocks.push_back(b);
    return static_cast<Key>(m_blocks.size()) - 1;
  }

  EvalMode eval_mode(Key const &k) const { return m_blocks[k].mode; }
  void pre_eval() {}
  void post_eval() {}

  double const &read(Key const &k, int i) const { return m_blocks[k].data[i]; }
  double const &read(Key const &k, int i, int j) const {
    Block const &b = m_blocks[k];
    return b.data[static_cast<std::size_t>(i) * b.cols + j];
  }
  double &write(Key const &k, int i, int j) {
    Block &b = m_blocks[k];
    return b.data[static_cast<std::size_t>(i) * b.cols + j];
  }

  template <typename Scalar>
  struct Val {
    static void set(BasicClexParamPack &p, Key const &k, int i, Scalar v) {
      p.write(k, i, 0) = v;
    }
    static void set(BasicClexParamPack &p, Key const &k, int i, int j,
                    Scalar v) {
      p.write(k, i, j) = v;
    }
    static Scalar get(BasicClexParamPack const &p, Key const &k, int i) {
      return p.read(k, i, 0);
    }
    static Scalar get(BasicClexParamPack const &p, Key const &k, int i,
                      int j) {
      return p.read(k, i, j);
    }
  };

 private:
  std::vector<Block> m_blocks;
};

}  // namespace clexulator
}  // namespace CASM

#endif
