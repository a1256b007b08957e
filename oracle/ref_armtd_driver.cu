// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// C entry points around the REFERENCE'S OWN ARMTD comparison planner (KPA =
// kinova_src/kinova_simulator_interfaces/kinova_planner_realtime_armtd_comparison): PZsparse.cu, Trajectory.cu,
// Dynamics.cu, CollisionChecking.cu, NLPclass.cu compiled unmodified by nvcc from /root/reference against the stand-in
// headers in oracle/shim (oracle/Makefile target `ref`).  The sequence is the reference main's (KPA/armtd_main.cu:110-178):
// ConstantAccelerationCurve from the offline JRS tables, makePolyZono, fk, reduce_link_PZ, initializeHyperPlane,
// armtd_NLP::set_parameters; the TNLP callbacks are then called directly.  Needs a GPU.  T = 100 at compile time.
#include "NLPclass.h"

#include <cstdint>

namespace {
struct RefArmtdPlan {
    std::vector<double> q0, qd0, jrs, k_range, q_des, obstacles;
    ConstantAccelerationCurve* traj = nullptr;
    KinematicsDynamics* kd = nullptr;
    Obstacles* O = nullptr;
    armtd_NLP* nlp = nullptr;
    std::vector<Eigen::Matrix<double, 3, 3 + 3>> link_gens;
    ~RefArmtdPlan() { delete nlp; delete O; delete kd; delete traj; }
};
}  // namespace

extern "C" {
int refarmtd_num_time_steps() { return NUM_TIME_STEPS; }

// jrs = [6][NUM_FACTORS][NUM_TIME_STEPS]: c_cos, g_cos, r_cos, c_sin, g_sin, r_sin (KPA/armtd_main.cu:41-90)
void* refarmtd_build(const double* q0, const double* qd0, const double* jrs, const double* k_range, const double* q_des,
                     const double* obstacles, int num_obstacles, int num_threads) {
    RefArmtdPlan* p = new RefArmtdPlan();
    const size_t NT = (size_t)NUM_FACTORS * NUM_TIME_STEPS;
    p->q0.assign(q0, q0 + NUM_FACTORS); p->qd0.assign(qd0, qd0 + NUM_FACTORS);
    p->k_range.assign(k_range, k_range + NUM_FACTORS); p->q_des.assign(q_des, q_des + NUM_FACTORS);
    p->jrs.assign(jrs, jrs + 6 * NT);
    p->obstacles.assign(obstacles, obstacles + (size_t)num_obstacles * (MAX_OBSTACLE_GENERATOR_NUM + 1) * 3);
    omp_set_num_threads(num_threads > 0 ? num_threads : 1);
    try {
        double* J = p->jrs.data();
        p->traj = new ConstantAccelerationCurve(p->q0.data(), p->qd0.data(), J, J + NT, J + 2 * NT, J + 3 * NT, J + 4 * NT, J + 5 * NT, p->k_range.data());
        p->O = new Obstacles(p->obstacles.data(), num_obstacles);
        p->kd = new KinematicsDynamics(p->traj);
        p->link_gens.resize((size_t)NUM_TIME_STEPS * NUM_JOINTS);
        int t = 0;
#pragma omp parallel for private(t) schedule(dynamic, 1)
        for (t = 0; t < NUM_TIME_STEPS; t++) p->traj->makePolyZono(t);
#pragma omp parallel for private(t) schedule(dynamic)
        for (t = 0; t < NUM_TIME_STEPS; t++) {
            p->kd->fk(t);
            for (int i = 0; i < NUM_JOINTS; i++) p->link_gens[(size_t)t * NUM_JOINTS + i] = p->kd->links(i, t).reduce_link_PZ();
        }
        p->O->initializeHyperPlane(p->link_gens.data());
        p->nlp = new armtd_NLP();
        p->nlp->set_parameters(p->q_des.data(), p->traj, p->kd, p->O);
    } catch (...) {
        delete p;
        return nullptr;
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { delete p; return nullptr; }
    return p;
}
void refarmtd_destroy(void* h) { delete (RefArmtdPlan*)h; }

// which: 2 = R, 3 = R_t, 7 = links (numbering of oracle_get_pz)
int refarmtd_get_pz(void* h, int which, int idx, int s, int* dims, uint64_t* keys, double* coeffs, double* center, double* independent) {
    RefArmtdPlan& p = *(RefArmtdPlan*)h;
    PZsparseArray* t = which == 2 ? &p.traj->R : which == 3 ? &p.traj->R_t : which == 7 ? &p.kd->links : nullptr;
    if (!t) return -1;
    const PZsparse& z = (*t)(idx, s);
    const int dim = z.NRows * z.NCols;
    if (dims) { dims[0] = z.NRows; dims[1] = z.NCols; }
    const int n = (int)z.polynomial.size();
    if (keys) for (int i = 0; i < n; i++) keys[i] = z.polynomial[i].degree;
    if (coeffs) for (int i = 0; i < n; i++) for (int a = 0; a < dim; a++) coeffs[(size_t)i * dim + a] = z.polynomial[i].coeff(a);
    if (center) for (int a = 0; a < dim; a++) center[a] = z.center(a);
    if (independent) for (int a = 0; a < dim; a++) independent[a] = z.independent(a);
    return n;
}
void refarmtd_get_link_generators(void* h, double* out) {
    RefArmtdPlan& p = *(RefArmtdPlan*)h;
    for (size_t e = 0; e < p.link_gens.size(); e++) for (int a = 0; a < 18; a++) out[e * 18 + a] = p.link_gens[e](a);
}
int refarmtd_get_nlp_info(void* h, int* n, int* m, int* nnz_jac_g, int* nnz_h_lag) {
    Ipopt::TNLP::IndexStyleEnum style;
    return ((RefArmtdPlan*)h)->nlp->get_nlp_info(*n, *m, *nnz_jac_g, *nnz_h_lag, style) ? 0 : -1;
}
int refarmtd_get_bounds_info(void* h, int n, int m, double* x_l, double* x_u, double* g_l, double* g_u) {
    return ((RefArmtdPlan*)h)->nlp->get_bounds_info(n, x_l, x_u, m, g_l, g_u) ? 0 : -1;
}
int refarmtd_eval_f(void* h, const double* x, double* f, double* grad_f) {
    armtd_NLP* nlp = ((RefArmtdPlan*)h)->nlp;
    bool ok = nlp->eval_f(NUM_FACTORS, x, true, *f);
    ok = ok && nlp->eval_grad_f(NUM_FACTORS, x, true, grad_f);
    return ok ? 0 : -1;
}
int refarmtd_eval_g(void* h, const double* x, int m, double* g) { return ((RefArmtdPlan*)h)->nlp->eval_g(NUM_FACTORS, x, true, m, g) ? 0 : -1; }
int refarmtd_eval_jac_g(void* h, const double* x, int m, double* values) {
    return ((RefArmtdPlan*)h)->nlp->eval_jac_g(NUM_FACTORS, x, true, m, m * NUM_FACTORS, nullptr, nullptr, values) ? 0 : -1;
}
int refarmtd_check_feasible(void* h, const double* x, int m, const double* g) {
    armtd_NLP* nlp = ((RefArmtdPlan*)h)->nlp;
    nlp->finalize_solution(Ipopt::SUCCESS, NUM_FACTORS, x, nullptr, nullptr, m, g, nullptr, 0.0, nullptr, nullptr);
    return nlp->feasible ? 1 : 0;
}
}
