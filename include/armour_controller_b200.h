/* armour_controller_b200.h — C ABI of the B200-native robust low-level controller.
 *
 * Drop-in boundary for kinova_src/kinova_simulator_interfaces/kinova_robust_controllers_mex/ ("KRC"), whose two
 * MEX functions the MATLAB controller uarmtd_robust_CBF_LLC.m:166-168 calls once per control tick:
 *   [u, tau, v] = kinova_controller        (Kr, alpha, V_max, r_norm_threshold, q, qd, q_des, qd_des, qdd_des, eps)
 *   [u, tau, v] = kinova_controller_ALTHOFF(Kr, Kp,    Ki,    maxError,         q, qd, q_des, qd_des, qdd_des, eps)
 * Each call is: nominal passivity RNEA, interval passivity RNEA over the +-eps inertial-parameter box, the error
 * bound between them and the robust input v (RobustController::update, KRC/robust_controller.cpp:62-171).  Here
 * `count` ticks / simulation instances are evaluated per launch, one thread per sample (closed-loop sweeps and
 * Monte-Carlo simulation batch them; count = 1 is the reference's call).  Same library as armour_b200.h
 * (libarmour_b200.so), same conventions: plain pointers and sizes, 0 or a negative ARMOUR_E_* code,
 * armour_last_error() has the text, no CPU fallback.
 */
#ifndef ARMOUR_CONTROLLER_B200_H
#define ARMOUR_CONTROLLER_B200_H

#include "armour_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ARMOUR_CONTROLLER_MAX_JOINTS 7

typedef struct armour_controller armour_controller;

/* Robot::Robot(filename, eps): Model(filename) + IntModel(model, eps) (KRC/robot_models.cpp:20-151, 168-249, 257-261).
 * `robot_model_file` is the reference's text description (e.g. KRC/kinova_without_gripper.txt); the conversion to
 * body-CoM frames, the interval model and the q-independent part of the RNEA (body-to-world transforms, screw axes in
 * body frames, KRC/rnea.cpp:38-48) are computed once, on the device.  device = -1: current device. */
int armour_controller_create(const char* robot_model_file, double model_uncertainty, int device, armour_controller** out);
void armour_controller_destroy(armour_controller* c);
int armour_controller_num_joints(armour_controller* c, int* num_joints);

/* kinova_controller (KRC/kinova_controller.cpp:15-84), ROBUST_INPUT_METHOD::ARMOUR.
 * Kr[n] is the diagonal of the gain matrix; state arrays are [count][n] in the MEX argument order
 * q, q_d (measured), qd, qd_d, qd_dd (desired).  Outputs [count][n]: u = u_nominal - v, u_nominal (the MEX's `tau`), v.
 * Optional (may be NULL) debug outputs: u_interval [count][n][2] = lower, upper of the interval torque,
 * V_sup [count] = upper bound of 0.5 r' M(q) r (0 where ||r|| <= r_norm_threshold).
 * *outside (may be NULL) = number of samples whose nominal torque left the interval torque — the reference prints
 * "Nominal model output falls outside interval output" and throws there; the call returns ARMOUR_E_NUMERIC. */
int armour_controller_update(armour_controller* c, int count, const double* Kr, double alpha, double V_max, double r_norm_threshold,
                             const double* q, const double* q_d, const double* qd, const double* qd_d, const double* qd_dd,
                             double* u, double* u_nominal, double* v, double* u_interval, double* V_sup, int* outside);

/* kinova_controller_ALTHOFF (KRC/kinova_controller_ALTHOFF.cpp:15-90), ROBUST_INPUT_METHOD::ALTHOFF with
 * deltaT = eAcc = 0 as that MEX calls update(): v = -(Kp[1] ||bound|| + Kp[0]) r.  Ki and max_error are accepted for
 * signature parity and, like there, do not influence a stateless call. */
int armour_controller_update_althoff(armour_controller* c, int count, const double* Kr, const double* Kp, const double* Ki, double max_error,
                                     const double* q, const double* q_d, const double* qd, const double* qd_d, const double* qd_dd,
                                     double* u, double* u_nominal, double* v, double* u_interval, int* outside);

/* passRNEA / passRNEA_Int on their own (KRC/rnea.cpp:6-93, 95-185): tau [count][n], tau_interval [count][n][2]
 * (either may be NULL).  apply_gravity = 0 with qd = qda = 0 gives M(q) qdd. */
int armour_controller_rnea(armour_controller* c, int count, const double* q, const double* qd, const double* qda, const double* qdd,
                           int apply_gravity, double* tau, double* tau_interval);

/* Device-resident variant for closed-loop sweeps and bench.py: the five state arrays are uploaded once
 * ([5][count][n] = q, q_d, qd, qd_d, qd_dd), armour_controller_update_resident runs the ARMOUR-method kernel on them and
 * leaves u / u_nominal / v on the device; armour_controller_download copies them out ([3][count][n]). */
int armour_controller_upload(armour_controller* c, int count, const double* states);
int armour_controller_update_resident(armour_controller* c, const double* Kr, double alpha, double V_max, double r_norm_threshold);
int armour_controller_download(armour_controller* c, double* u_unom_v, int* outside);
/* Calls with count >= 32768 run as a two-stream pipeline of 32768-tick chunks (copy in, kernel, copy out) and page-lock the
 * caller's arrays the first time they are seen (closed-loop sweeps pass the same arrays every tick).  The arrays stay
 * page-locked until this call or armour_controller_destroy: call it before freeing them. */
int armour_controller_release_host_buffers(armour_controller* c);
/* milliseconds of the last update: the kernel alone for count < 32768, the whole pipeline (copies included) above */
int armour_controller_last_ms(armour_controller* c, double* kernel_ms);

#ifdef __cplusplus
}
#endif
#endif
