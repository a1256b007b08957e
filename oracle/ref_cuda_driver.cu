// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// C entry points around the REFERENCE'S OWN complete planning path, CUDA kernels included: KPR/PZsparse.cu,
// Trajectory.cu, Dynamics.cu, CollisionChecking.cu and NLPclass.cu are compiled unmodified by nvcc from /root/reference
// (oracle/Makefile target `ref`) against the stand-in Eigen / Boost / Ipopt headers in oracle/shim.  The sequence
// below is the reference main's (KPR/armour_main.cu:87-241): reach sets, torque radius, Obstacles +
// initializeHyperPlane, armtd_NLP::set_parameters; the TNLP callbacks are then called directly (no solver).
// Needs a GPU (the reference's half-space tables and plane tests are CUDA kernels); used by the `-m gpu` tests in
// tests/test_reference_pin.py to pin constraints and Jacobians of the device path against the reference itself.
#include "NLPclass.h"

#include <chrono>
#include <cstdint>

namespace {
struct RefCudaPlan {
    BezierCurve* traj = nullptr;
    KinematicsDynamics* kd = nullptr;
    Obstacles* O = nullptr;
    armtd_NLP* nlp = nullptr;
    std::vector<double> obstacles;
    std::vector<Eigen::Matrix<double, 3, 3 + 3>> link_gens;
    Eigen::MatrixXd torque_radius;
    Eigen::VectorXd q_des;
    double build_ms = 0.0;   // the reference's own "Time taken by generating reachable sets" span (KPR/armour_main.cu:89-226)
    ~RefCudaPlan() { delete nlp; delete O; delete kd; delete traj; }
};
}  // namespace

extern "C" {
void* refcuda_build(const double* q0_in, const double* qd0_in, const double* qdd0_in, const double* q_des_in, double t_plan,
                    const double* obstacles, int num_obstacles, int num_threads) {
    Eigen::VectorXd q0(NUM_FACTORS), qd0(NUM_FACTORS), qdd0(NUM_FACTORS);
    RefCudaPlan* p = new RefCudaPlan();
    p->q_des.resize(NUM_FACTORS);
    for (int i = 0; i < NUM_FACTORS; i++) { q0[i] = q0_in[i]; qd0[i] = qd0_in[i]; qdd0[i] = qdd0_in[i]; p->q_des[i] = q_des_in[i]; }
    p->obstacles.assign(obstacles, obstacles + (size_t)num_obstacles * (MAX_OBSTACLE_GENERATOR_NUM + 1) * 3);
    omp_set_num_threads(num_threads > 0 ? num_threads : 1);
    try {
        p->O = new Obstacles(p->obstacles.data(), num_obstacles);   // KPR/armour_main.cu:87, before the reference starts its timer (:89)
        const auto t_start = std::chrono::high_resolution_clock::now();
        p->traj = new BezierCurve(q0, qd0, qdd0);
        int s = 0;
#pragma omp parallel for private(s) schedule(dynamic, 1)
        for (s = 0; s < NUM_TIME_STEPS; s++) p->traj->makePolyZono(s);
        p->kd = new KinematicsDynamics(p->traj);
        p->link_gens.resize((size_t)NUM_TIME_STEPS * NUM_JOINTS);
        KinematicsDynamics& kd = *p->kd;
#pragma omp parallel for private(s) schedule(dynamic)
        for (s = 0; s < NUM_TIME_STEPS; s++) {
            kd.fk(s);
            for (int i = 0; i < NUM_JOINTS; i++) p->link_gens[(size_t)s * NUM_JOINTS + i] = kd.links(i, s).reduce_link_PZ();
            kd.rnea_nominal(s);
            kd.rnea_interval(s);
            for (int i = 0; i < NUM_FACTORS; i++) kd.u_nom_int(i, s) = kd.u_nom_int(i, s) - kd.u_nom(i, s);
            for (int i = 0; i < NUM_FACTORS; i++) kd.u_nom(i, s).reduce();
        }
        p->torque_radius = Eigen::MatrixXd::Zero(NUM_FACTORS, NUM_TIME_STEPS);
        for (int t = 0; t < NUM_TIME_STEPS; t++) {
            Interval rho_max_temp = Interval(0.0);
            for (int i = 0; i < NUM_FACTORS; i++) {
                MatrixXInt temp = kd.u_nom_int(i, t).toInterval();
                rho_max_temp += temp(0) * temp(0);
                p->torque_radius(i, t) = alpha * (M_max - M_min) * eps + 0.5 * max(fabs(temp(0).lower()), fabs(temp(0).upper()));
            }
            rho_max_temp = sqrt(rho_max_temp);
            for (int i = 0; i < NUM_FACTORS; i++) p->torque_radius(i, t) += 0.5 * rho_max_temp.upper();
            for (int i = 0; i < NUM_FACTORS; i++) p->torque_radius(i, t) += kd.u_nom(i, t).independent(0);
            for (int i = 0; i < NUM_FACTORS; i++) p->torque_radius(i, t) += friction[i];
        }
        p->O->initializeHyperPlane(p->link_gens.data());
        p->build_ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t_start).count();   // :224-226
        p->nlp = new armtd_NLP();
        p->nlp->set_parameters(p->q_des, t_plan, p->traj, p->kd, &p->torque_radius, p->O);
    } catch (...) {
        delete p;
        return nullptr;
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { delete p; return nullptr; }
    return p;
}
void refcuda_destroy(void* h) { delete (RefCudaPlan*)h; }
int refcuda_num_time_steps() { return NUM_TIME_STEPS; }
double refcuda_k_range(int i) { return k_range[i]; }
double refcuda_mass_uncertainty() { return mass_uncertainty; }
double refcuda_last_build_ms(void* h) { return ((RefCudaPlan*)h)->build_ms; }

int refcuda_get_nlp_info(void* h, int* n, int* m, int* nnz_jac_g, int* nnz_h_lag) {
    Ipopt::TNLP::IndexStyleEnum style;
    return ((RefCudaPlan*)h)->nlp->get_nlp_info(*n, *m, *nnz_jac_g, *nnz_h_lag, style) ? 0 : -1;
}
int refcuda_get_bounds_info(void* h, int n, int m, double* x_l, double* x_u, double* g_l, double* g_u) {
    return ((RefCudaPlan*)h)->nlp->get_bounds_info(n, x_l, x_u, m, g_l, g_u) ? 0 : -1;
}
int refcuda_get_starting_point(void* h, int n, int m, double* x) {
    return ((RefCudaPlan*)h)->nlp->get_starting_point(n, true, x, false, nullptr, nullptr, m, false, nullptr) ? 0 : -1;
}
int refcuda_eval_f(void* h, const double* x, double* f, double* grad_f) {
    armtd_NLP* nlp = ((RefCudaPlan*)h)->nlp;
    bool ok = nlp->eval_f(NUM_FACTORS, x, true, *f);
    ok = ok && nlp->eval_grad_f(NUM_FACTORS, x, true, grad_f);
    return ok ? 0 : -1;
}
int refcuda_eval_g(void* h, const double* x, int m, double* g) { return ((RefCudaPlan*)h)->nlp->eval_g(NUM_FACTORS, x, true, m, g) ? 0 : -1; }
int refcuda_eval_jac_g(void* h, const double* x, int m, double* values) {
    return ((RefCudaPlan*)h)->nlp->eval_jac_g(NUM_FACTORS, x, true, m, m * NUM_FACTORS, nullptr, nullptr, values) ? 0 : -1;
}
// armtd_NLP::finalize_solution's feasibility verdict for (x, g) (KPR/NLPclass.cu:446-537); it prints the violated rows
int refcuda_check_feasible(void* h, const double* x, int m, const double* g) {
    armtd_NLP* nlp = ((RefCudaPlan*)h)->nlp;
    nlp->finalize_solution(Ipopt::SUCCESS, NUM_FACTORS, x, nullptr, nullptr, m, g, nullptr, 0.0, nullptr, nullptr);
    return nlp->feasible ? 1 : 0;
}
void refcuda_get_link_sliced_center(void* h, double* out) {
    armtd_NLP* nlp = ((RefCudaPlan*)h)->nlp;
    for (int e = 0; e < NUM_TIME_STEPS * NUM_JOINTS; e++) for (int a = 0; a < 3; a++) out[e * 3 + a] = nlp->link_sliced_center[e](a);
}
}
