// IpTNLP.hpp — stand-in (TEST INFRASTRUCTURE ONLY, see oracle/refshim/README.md) for the Ipopt interface types the
// plugin derives from.  There is no solver behind it: OptimizeTNLP reports failure.  What is exercised is the
// reference's `problem` (the NLP callbacks Ipopt would call), not Ipopt.
#ifndef DPOSE_REFSHIM_IPOPT
#define DPOSE_REFSHIM_IPOPT
#include <string>
namespace Ipopt {
typedef int Index;
typedef double Number;

class ReferencedObject {
public:
  virtual ~ReferencedObject() = default;
  mutable int refs_ = 0;
};

template <typename T>
class SmartPtr {
public:
  SmartPtr() = default;
  SmartPtr(T* p) : p_(p) { acquire(); }
  SmartPtr(const SmartPtr& o) : p_(o.p_) { acquire(); }
  template <typename U>
  SmartPtr(const SmartPtr<U>& o) : p_(o.get()) {
    acquire();
  }
  ~SmartPtr() { release(); }
  SmartPtr& operator=(const SmartPtr& o) {
    if (o.p_ != p_) {
      release();
      p_ = o.p_;
      acquire();
    }
    return *this;
  }
  SmartPtr& operator=(T* p) { return *this = SmartPtr(p); }
  T* operator->() const { return p_; }
  T& operator*() const { return *p_; }
  T* get() const { return p_; }

private:
  void acquire() {
    if (p_) ++p_->refs_;
  }
  void release() {
    if (p_ && --p_->refs_ == 0) delete p_;
    p_ = nullptr;
  }
  T* p_ = nullptr;
};
template <typename T>
T*
GetRawPtr(const SmartPtr<T>& p) {
  return p.get();
}

enum SolverReturn { SUCCESS, MAXITER_EXCEEDED, INTERNAL_ERROR };
enum ApplicationReturnStatus { Solve_Succeeded = 0, Internal_Error = -199 };
class IpoptData;
class IpoptCalculatedQuantities;

class TNLP : public ReferencedObject {
public:
  enum IndexStyleEnum { C_STYLE = 0, FORTRAN_STYLE = 1 };
  virtual bool get_nlp_info(Index& n, Index& m, Index& nnz_jac_g, Index& nnz_h_lag, IndexStyleEnum& index_style) = 0;
  virtual bool get_bounds_info(Index n, Number* x_l, Number* x_u, Index m, Number* g_l, Number* g_u) = 0;
  virtual bool get_starting_point(Index n, bool init_x, Number* x, bool init_z, Number* z_L, Number* z_U, Index m,
                                  bool init_lambda, Number* lambda) = 0;
  virtual bool eval_f(Index n, const Number* x, bool new_x, Number& obj_value) = 0;
  virtual bool eval_grad_f(Index n, const Number* x, bool new_x, Number* grad_f) = 0;
  virtual bool eval_g(Index n, const Number* x, bool new_x, Index m, Number* g) = 0;
  virtual bool eval_jac_g(Index n, const Number* x, bool new_x, Index m, Index nele_jac, Index* iRow, Index* jCol,
                          Number* values) = 0;
  virtual void finalize_solution(SolverReturn status, Index n, const Number* x, const Number* z_L, const Number* z_U, Index m,
                                 const Number* g, const Number* lambda, Number obj_value, const IpoptData* ip_data,
                                 IpoptCalculatedQuantities* ip_cq) = 0;
};

class OptionsList : public ReferencedObject {
public:
  bool SetIntegerValue(const std::string&, Index) { return true; }
  bool SetNumericValue(const std::string&, Number) { return true; }
  bool SetStringValue(const std::string&, const std::string&) { return true; }
};

class IpoptApplication : public ReferencedObject {
public:
  SmartPtr<OptionsList> Options() { return options_; }
  ApplicationReturnStatus Initialize() { return Solve_Succeeded; }
  ApplicationReturnStatus OptimizeTNLP(const SmartPtr<TNLP>&) { return Internal_Error; }  // no solver here

private:
  SmartPtr<OptionsList> options_ = new OptionsList();
};
}  // namespace Ipopt

inline Ipopt::SmartPtr<Ipopt::IpoptApplication>
IpoptApplicationFactory() {
  return new Ipopt::IpoptApplication();
}
#endif
