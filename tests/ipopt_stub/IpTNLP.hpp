// TEST INFRASTRUCTURE — a minimal stand-in for Ipopt's public headers (Ipopt is not installed in this image).
// It reproduces the parts of the interface the reference's main() uses (KPR/armour_main.cu:238-317): the intrusive
// reference counting of Ipopt::ReferencedObject / SmartPtr (an object whose count drops to zero is DELETED — which is
// why the NLP must be heap-allocated), the TNLP callback signatures, and IpoptApplication with its option setters and
// return statuses.  Behind OptimizeTNLP sits the repo's stand-in Gauss-Newton solver, driven purely through the TNLP
// virtual interface, so the ARMOUR_HAVE_IPOPT branch of armour-dev_b200/host/armour_main.cpp is compiled AND executed
// in the tests exactly as it would be against the real library.  Not a solver to be trusted beyond that.
#pragma once
#include <cstddef>

namespace Ipopt {
typedef int Index;
typedef double Number;
enum SolverReturn { SUCCESS, MAXITER_EXCEEDED, CPUTIME_EXCEEDED, STOP_AT_TINY_STEP, STOP_AT_ACCEPTABLE_POINT, LOCAL_INFEASIBILITY, USER_REQUESTED_STOP,
                    FEASIBLE_POINT_FOUND, DIVERGING_ITERATES, RESTORATION_FAILURE, ERROR_IN_STEP_COMPUTATION, INVALID_NUMBER_DETECTED, TOO_FEW_DEGREES_OF_FREEDOM,
                    INVALID_OPTION, OUT_OF_MEMORY, INTERNAL_ERROR, UNASSIGNED, WALLTIME_EXCEEDED };
class IpoptData;
class IpoptCalculatedQuantities;

class ReferencedObject {
public:
    ReferencedObject() : count_(0) {}
    virtual ~ReferencedObject() {}
    void AddRef() const { ++count_; }
    int ReleaseRef() const { return --count_; }
    int ReferenceCount() const { return count_; }
private:
    mutable int count_;
};

template <class T>
class SmartPtr {
public:
    SmartPtr() : p_(nullptr) {}
    SmartPtr(T* p) : p_(p) { if (p_) p_->AddRef(); }
    SmartPtr(const SmartPtr& o) : p_(o.p_) { if (p_) p_->AddRef(); }
    template <class U> SmartPtr(const SmartPtr<U>& o) : p_(o.get()) { if (p_) p_->AddRef(); }
    ~SmartPtr() { release(); }
    SmartPtr& operator=(const SmartPtr& o) { if (o.p_) o.p_->AddRef(); release(); p_ = o.p_; return *this; }
    T* operator->() const { return p_; }
    T& operator*() const { return *p_; }
    T* get() const { return p_; }
private:
    void release() { if (p_ && p_->ReleaseRef() == 0) delete p_; p_ = nullptr; }   // like Ipopt: the last reference deletes the object
    T* p_;
};
template <class T> T* GetRawPtr(const SmartPtr<T>& p) { return p.get(); }
template <class T> bool IsValid(const SmartPtr<T>& p) { return p.get() != nullptr; }

class TNLP : public ReferencedObject {
public:
    enum IndexStyleEnum { C_STYLE = 0, FORTRAN_STYLE = 1 };
    virtual ~TNLP() {}
    virtual bool get_nlp_info(Index& n, Index& m, Index& nnz_jac_g, Index& nnz_h_lag, IndexStyleEnum& index_style) = 0;
    virtual bool get_bounds_info(Index n, Number* x_l, Number* x_u, Index m, Number* g_l, Number* g_u) = 0;
    virtual bool get_starting_point(Index n, bool init_x, Number* x, bool init_z, Number* z_L, Number* z_U, Index m, bool init_lambda, Number* lambda) = 0;
    virtual bool eval_f(Index n, const Number* x, bool new_x, Number& obj_value) = 0;
    virtual bool eval_grad_f(Index n, const Number* x, bool new_x, Number* grad_f) = 0;
    virtual bool eval_g(Index n, const Number* x, bool new_x, Index m, Number* g) = 0;
    virtual bool eval_jac_g(Index n, const Number* x, bool new_x, Index m, Index nele_jac, Index* iRow, Index* jCol, Number* values) = 0;
    virtual bool eval_h(Index, const Number*, bool, Number, Index, const Number*, bool, Index, Index*, Index*, Number*) { return false; }
    virtual void finalize_solution(SolverReturn status, Index n, const Number* x, const Number* z_L, const Number* z_U, Index m, const Number* g,
                                   const Number* lambda, Number obj_value, const IpoptData* ip_data, IpoptCalculatedQuantities* ip_cq) = 0;
};
}  // namespace Ipopt
