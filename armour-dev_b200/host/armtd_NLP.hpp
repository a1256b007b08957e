// armtd_NLP — host-side mirror of the reference's Ipopt TNLP (KPR/NLPclass.h:11-184, KPR/NLPclass.cu),
// backed by the device through the C ABI (include/armour_b200.h).
//
// With Ipopt on the include path the class derives from Ipopt::TNLP and can be handed to
// IpoptApplication::OptimizeTNLP exactly like the reference's (KPR/armour_main.cu:238-285).  Without Ipopt
// (this image) the same callbacks exist with the same signatures on a minimal stand-in base, and
// standin_solver.hpp drives them.
//
// Difference to the reference's set_parameters (KPR/NLPclass.cu:30-60): the BezierCurve, KinematicsDynamics,
// torque_radius and Obstacles objects it borrowed all live inside the armour_handle, so the call takes the
// handle instead of four pointers.  Ownership: the NLP borrows the handle; the caller keeps it alive through the
// solve (same rule as the reference, KPR/armour_main.cu:238-241).  Callbacks return false when the C ABI reports
// an error (the reference always returned true and never looked at CUDA errors).
#pragma once
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/armour_b200.h"

#if defined(__has_include)
#if __has_include("IpTNLP.hpp")
#include "IpTNLP.hpp"
#define ARMOUR_HAVE_IPOPT 1
#endif
#endif

#ifndef ARMOUR_HAVE_IPOPT
namespace Ipopt {   // just enough of Ipopt's vocabulary for the callbacks to keep their signatures
typedef int Index;
typedef double Number;
enum SolverReturn { SUCCESS, MAXITER_EXCEEDED, CPUTIME_EXCEEDED, STOP_AT_ACCEPTABLE_POINT, LOCAL_INFEASIBILITY, INTERNAL_ERROR };
class IpoptData;
class IpoptCalculatedQuantities;
class TNLP {
public:
    enum IndexStyleEnum { C_STYLE = 0, FORTRAN_STYLE = 1 };
    virtual ~TNLP() {}
};
}  // namespace Ipopt
#endif

class armtd_NLP : public Ipopt::TNLP {
public:
    typedef Ipopt::Index Index;
    typedef Ipopt::Number Number;

    armtd_NLP() {}
    virtual ~armtd_NLP() {}

    bool set_parameters(const double* q_des_input, double t_plan_input, armour_handle* handle_input) {
        for (int i = 0; i < ARMOUR_NUM_FACTORS; i++) q_des[i] = q_des_input[i];
        t_plan = t_plan_input;
        handle = handle_input;
        int n = 0, m = 0, nnz = 0, nh = 0;
        if (armour_get_nlp_info(handle, &n, &m, &nnz, &nh) != ARMOUR_OK) return false;
        constraint_number = m;
        g_copy.assign(m, 0.0);   // the reference leaks its g_copy when set_parameters is called twice; a vector does not
        return true;
    }

    virtual bool get_nlp_info(Index& n, Index& m, Index& nnz_jac_g, Index& nnz_h_lag, IndexStyleEnum& index_style) {
        int n_ = 0, m_ = 0, nnz_ = 0, nh_ = 0;
        if (armour_get_nlp_info(handle, &n_, &m_, &nnz_, &nh_) != ARMOUR_OK) return false;
        n = n_; m = m_; nnz_jac_g = nnz_; nnz_h_lag = nh_;
        index_style = TNLP::C_STYLE;
        return true;
    }
    virtual bool get_bounds_info(Index n, Number* x_l, Number* x_u, Index m, Number* g_l, Number* g_u) {
        if (n != ARMOUR_NUM_FACTORS) printf("*** Error wrong value of n in get_bounds_info!");
        if (m != constraint_number) printf("*** Error wrong value of m in get_bounds_info!");
        return armour_get_bounds_info(handle, x_l, x_u, g_l, g_u) == ARMOUR_OK;
    }
    virtual bool get_starting_point(Index n, bool init_x, Number* x, bool init_z, Number*, Number*, Index, bool init_lambda, Number*) {
        if (init_x == false || init_z == true || init_lambda == true) printf("*** Error wrong value of init in get_starting_point!");
        if (n != ARMOUR_NUM_FACTORS) printf("*** Error wrong value of n in get_starting_point!");
        return armour_get_starting_point(handle, x) == ARMOUR_OK;
    }
    virtual bool eval_f(Index n, const Number* x, bool, Number& obj_value) {
        if (n != ARMOUR_NUM_FACTORS) printf("*** Error wrong value of n in eval_f!");
        return armour_eval_f(handle, q_des, t_plan, x, &obj_value) == ARMOUR_OK;
    }
    virtual bool eval_grad_f(Index n, const Number* x, bool, Number* grad_f) {
        if (n != ARMOUR_NUM_FACTORS) printf("*** Error wrong value of n in eval_grad_f!");
        return armour_eval_grad_f(handle, q_des, t_plan, x, grad_f) == ARMOUR_OK;
    }
    virtual bool eval_g(Index n, const Number* x, bool, Index m, Number* g) {
        if (n != ARMOUR_NUM_FACTORS) printf("*** Error wrong value of n in eval_g!");
        if (m != constraint_number) printf("*** Error wrong value of m in eval_g!");
        return armour_eval_g(handle, x, g) == ARMOUR_OK;
    }
    virtual bool eval_jac_g(Index n, const Number* x, bool, Index m, Index, Index* iRow, Index* jCol, Number* values) {
        if (n != ARMOUR_NUM_FACTORS) printf("*** Error wrong value of n in eval_g!");
        if (m != constraint_number) printf("*** Error wrong value of m in eval_g!");
        if (values == NULL) {
            std::vector<int> r((size_t)m * n), c((size_t)m * n);
            if (armour_jac_structure(handle, r.data(), c.data()) != ARMOUR_OK) return false;
            for (size_t i = 0; i < r.size(); i++) { iRow[i] = r[i]; jCol[i] = c[i]; }
            return true;
        }
        return armour_eval_jac_g(handle, x, values) == ARMOUR_OK;
    }
    virtual bool eval_h(Index, const Number*, bool, Number, Index, const Number*, bool, Index, Index*, Index*, Number*) {
        return false;   // limited-memory Hessian approximation, like the reference (KPR/NLPclass.cu:402-417)
    }
    virtual void finalize_solution(Ipopt::SolverReturn, Index n, const Number* x, const Number*, const Number*, Index m, const Number* g,
                                   const Number*, Number obj_value, const Ipopt::IpoptData*, Ipopt::IpoptCalculatedQuantities*) {
        for (Index i = 0; i < n; i++) solution[i] = (double)x[i];
        if (verbose) printf("        CUDA & C++: Ipopt: final cost function value: %g\n", obj_value / 10.0);
        memcpy(g_copy.data(), g, (size_t)m * sizeof(Number));
        int f = 0;
        feasible = (armour_check_feasible(handle, g_copy.data(), &f) == ARMOUR_OK) && f != 0;
        // link_sliced_center at the final x, as written to armour_joint_position_center.out
        int nn = 0, mm = 0, a = 0, b = 0;
        armour_get_nlp_info(handle, &nn, &mm, &a, &b);
        std::vector<double> gtmp(mm);
        armour_eval_g(handle, x, gtmp.data());
        link_sliced_center.assign((size_t)num_time_steps() * ARMOUR_NUM_JOINTS * 3, 0.0);
        armour_get_link_sliced_center(handle, link_sliced_center.data());
        // Under cfg.pin_user_buffers the library page-locked the arrays the solver handed to the callbacks (and gtmp above).  The
        // solver frees them once the solve returns, so their registrations end here.
        armour_release_host_buffers(handle);
    }

    int num_time_steps() const { return (constraint_number - 4 * ARMOUR_NUM_FACTORS) > 0 ? time_steps : 0; }
    void set_time_steps(int T) { time_steps = T; }

    double t_plan = 1.0;
    bool verbose = false;   // the reference prints the final cost from finalize_solution (KPR/NLPclass.cu:443)
    double solution[ARMOUR_NUM_FACTORS] = {0};
    Index constraint_number = 0;
    std::vector<Number> g_copy;
    bool feasible = false;
    std::vector<double> link_sliced_center;   // [t*7 + link][3]

private:
    armtd_NLP(const armtd_NLP&);
    armtd_NLP& operator=(const armtd_NLP&);
    double q_des[ARMOUR_NUM_FACTORS] = {0};
    armour_handle* handle = nullptr;
    int time_steps = 128;
};
