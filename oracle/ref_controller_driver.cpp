// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// C entry points around the REFERENCE'S OWN robust controller (KRC =
// kinova_src/kinova_simulator_interfaces/kinova_robust_controllers_mex): robot_models.cpp, rnea.cpp, robust_controller.cpp
// (which pull in spatial.cpp / spatial_interval.cpp) compiled unmodified from /root/reference against the stand-in
// Eigen / Boost headers in oracle/shim (oracle/Makefile target `ref`).  The body of refctrl_update is what the two MEX
// gateways do per call (KRC/kinova_controller.cpp:15-84, kinova_controller_ALTHOFF.cpp:15-90), minus the mxArray
// plumbing; the Robot is built once per handle instead of once per tick.
#include "robust_controller.hpp"

static_assert(sizeof(Eigen::Vector3d) == 3 * sizeof(double), "fixed-size stand-in matrices must be plain arrays");

namespace {
struct RefController {
    Robot* robot = nullptr;
    ~RefController() { delete robot; }
};
}  // namespace

extern "C" {
void* refctrl_create(const char* model_file, double eps) {
    RefController* c = new RefController();
    try {
        c->robot = new Robot(std::string(model_file), eps);
    } catch (...) {
        delete c;
        return nullptr;
    }
    return c;
}
void refctrl_destroy(void* h) { delete (RefController*)h; }
int refctrl_num_joints(void* h) { return ((RefController*)h)->robot->numJoints; }

// method 0: kinova_controller (ARMOUR robust input), par = alpha, V_max, r_norm_threshold
// method 1: kinova_controller_ALTHOFF, par = Kp[0], Kp[1], Ki[0], Ki[1], maxError
int refctrl_update(void* h, int method, int count, const double* Kr_in, const double* par, const double* q_in, const double* q_d_in,
                   const double* qd_in, const double* qd_d_in, const double* qd_dd_in, double* u_out, double* u_nominal_out, double* v_out) {
    Robot* robot = ((RefController*)h)->robot;
    const int n = robot->numJoints;
    Eigen::MatrixXd Kr = Eigen::MatrixXd::Identity(n, n);
    for (int i = 0; i < n; i++) Kr(i, i) = Kr_in[i];
    for (int s = 0; s < count; s++) {
        Eigen::VectorXd q(n), q_d(n), qd(n), qd_d(n), qd_dd(n);
        for (int i = 0; i < n; i++) {
            q(i) = q_in[s * n + i]; q_d(i) = q_d_in[s * n + i]; qd(i) = qd_in[s * n + i]; qd_d(i) = qd_d_in[s * n + i]; qd_dd(i) = qd_dd_in[s * n + i];
        }
        Eigen::VectorXd u;
        if (method == 0) {
            RobustController c(Kr, par[0], par[1], par[2]);
            c.applyFriction = false;
            u = c.update(robot, q, q_d, qd, qd_d, qd_dd);
            for (int i = 0; i < n; i++) { u_out[s * n + i] = u(i); u_nominal_out[s * n + i] = c.u_nominal(i); v_out[s * n + i] = c.v(i); }
        } else {
            Eigen::Vector2d Kp, Ki;
            Kp[0] = par[0]; Kp[1] = par[1]; Ki[0] = par[2]; Ki[1] = par[3];
            RobustController c(Kr, Kp, Ki, par[4]);
            c.applyFriction = false;
            u = c.update(robot, q, q_d, qd, qd_d, qd_dd);
            for (int i = 0; i < n; i++) { u_out[s * n + i] = u(i); u_nominal_out[s * n + i] = c.u_nominal(i); v_out[s * n + i] = c.v(i); }
        }
    }
    return 0;
}
// passRNEA / passRNEA_Int (KRC/rnea.cpp) for one sample
void refctrl_rnea(void* h, const double* q_in, const double* qd_in, const double* qda_in, const double* qdd_in, int gravity, double* tau, double* tau_lo_hi) {
    Robot* robot = ((RefController*)h)->robot;
    const int n = robot->numJoints;
    Eigen::VectorXd q(n), qd(n), qda(n), qdd(n), t(n);
    for (int i = 0; i < n; i++) { q(i) = q_in[i]; qd(i) = qd_in[i]; qda(i) = qda_in[i]; qdd(i) = qdd_in[i]; }
    passRNEA(t, robot->RobotModelPtr, q, qd, qda, qdd, false, gravity != 0);
    for (int i = 0; i < n; i++) tau[i] = t(i);
    VectorXint ti(n);
    passRNEA_Int(ti, robot->IntRobotModelPtr, q, qd, qda, qdd, false, gravity != 0);
    for (int i = 0; i < n; i++) { tau_lo_hi[2 * i] = ti(i).lower(); tau_lo_hi[2 * i + 1] = ti(i).upper(); }
}
}
