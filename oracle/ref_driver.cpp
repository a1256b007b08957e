// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// Thin C entry points around the REFERENCE'S OWN reach-set code, compiled from the sources where they lie under
// /root/reference (KPR/PZsparse.cu, Trajectory.cu, Dynamics.cu; see oracle/Makefile target `ref`) against the stand-in
// Eigen / Boost headers in oracle/shim.  The build sequence below is the call sequence of the reference's main
// (KPR/armour_main.cu:94-205: makePolyZono per interval, fk, reduce_link_PZ, rnea_nominal, rnea_interval,
// u_nom_int - u_nom, reduce, torque radius); every PZ operation executed is the reference's.  Used by
// tests/test_reference_pin.py to pin oracle/oracle_armour.cpp (and, on the GPU box, the device path) against it.
// NUM_TIME_STEPS, k_range, thresholds are the reference's compile-time values (KPR/Parameters.h: T = 128, pi/48).
#include "Dynamics.h"

#include <cstdint>

namespace {
struct RefPlan {
    BezierCurve* traj = nullptr;
    KinematicsDynamics* kd = nullptr;
    std::vector<Eigen::MatrixXd> link_gens;   // [t*NUM_JOINTS + i], 3x6
    Eigen::MatrixXd torque_radius;            // [NUM_FACTORS x NUM_TIME_STEPS]
    ~RefPlan() { delete kd; delete traj; }
};
PZsparseArray* table(RefPlan& p, int which) {
    switch (which) {
        case 0: return &p.traj->cos_q_des; case 1: return &p.traj->sin_q_des; case 2: return &p.traj->R; case 3: return &p.traj->R_t;
        case 4: return &p.traj->qd_des; case 5: return &p.traj->qda_des; case 6: return &p.traj->qdda_des;
        case 7: return &p.kd->links; case 8: return &p.kd->u_nom; case 9: return &p.kd->u_nom_int;
    }
    return nullptr;
}
}  // namespace

extern "C" {
int ref_num_time_steps() { return NUM_TIME_STEPS; }
double ref_k_range(int i) { return k_range[i]; }

void* ref_build(const double* q0_in, const double* qd0_in, const double* qdd0_in, int num_threads) {
    Eigen::VectorXd q0(NUM_FACTORS), qd0(NUM_FACTORS), qdd0(NUM_FACTORS);
    for (int i = 0; i < NUM_FACTORS; i++) { q0[i] = q0_in[i]; qd0[i] = qd0_in[i]; qdd0[i] = qdd0_in[i]; }
    RefPlan* p = new RefPlan();
    omp_set_num_threads(num_threads > 0 ? num_threads : 1);
    try {
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
                p->torque_radius(i, t) = alpha * (M_max - M_min) * eps + 0.5 * max(std::fabs(temp(0).lower()), std::fabs(temp(0).upper()));   // the reference is built by nvcc, whose global abs(double) is the floating one; g++ alone would pick abs(int)
            }
            rho_max_temp = sqrt(rho_max_temp);
            for (int i = 0; i < NUM_FACTORS; i++) p->torque_radius(i, t) += 0.5 * rho_max_temp.upper();
            for (int i = 0; i < NUM_FACTORS; i++) p->torque_radius(i, t) += kd.u_nom(i, t).independent(0);
            for (int i = 0; i < NUM_FACTORS; i++) p->torque_radius(i, t) += friction[i];
        }
    } catch (...) {
        delete p;
        return nullptr;
    }
    return p;
}
void ref_destroy(void* h) { delete (RefPlan*)h; }

// same layout as oracle_get_pz (oracle/oracle_armour.cpp)
int ref_get_pz(void* h, int which, int idx, int s, int* dims, uint64_t* keys, double* coeffs, double* center, double* independent) {
    RefPlan& p = *(RefPlan*)h;
    PZsparseArray* t = table(p, which);
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
void ref_get_torque_radius(void* h, double* out) {   // out[t*NUM_FACTORS + i]
    RefPlan& p = *(RefPlan*)h;
    for (int t = 0; t < NUM_TIME_STEPS; t++) for (int i = 0; i < NUM_FACTORS; i++) out[t * NUM_FACTORS + i] = p.torque_radius(i, t);
}
void ref_get_link_generators(void* h, double* out) {   // out[(t*NUM_JOINTS + i)*18 + col*3 + row]
    RefPlan& p = *(RefPlan*)h;
    for (size_t e = 0; e < p.link_gens.size(); e++) for (int a = 0; a < 18; a++) out[e * 18 + a] = p.link_gens[e](a);
}
// PZsparse::slice at k (KPR/PZsparse.cu:404-435) of one table entry: lower/upper per component
int ref_slice(void* h, int which, int idx, int s, const double* factor, double* lo, double* hi) {
    RefPlan& p = *(RefPlan*)h;
    PZsparseArray* t = table(p, which);
    if (!t) return -1;
    PZsparse& z = (*t)(idx, s);
    MatrixXInt r = z.slice(factor);
    for (int a = 0; a < r.size(); a++) { lo[a] = r(a).lower(); hi[a] = r(a).upper(); }
    return r.size();
}
}
