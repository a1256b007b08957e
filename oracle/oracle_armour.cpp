// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (C++17 + OpenMP, no third-party dependencies) of ARMOUR's online
// reachability + constraint path, following (KPR = kinova_src/kinova_simulator_interfaces/
// kinova_planner_realtime in the reference tree):
//   KPR/KinovaWithoutGripperInfo.h:10-112   robot constants
//   KPR/Parameters.h:10-59                  sizes / thresholds (runtime here)
//   KPR/Trajectory.cu:15-61                 BezierCurve ctor (k-independent extrema)
//   KPR/Trajectory.cu:63-254                makePolyZono
//   KPR/Trajectory.cu:256-540               joint position/velocity extrema + gradients
//   KPR/Trajectory.cu:542-822               Bezier helpers + generated k-derivatives
//   KPR/Dynamics.cu:6-181                   KinematicsDynamics ctor, fk, rnea
//   KPR/armour_main.cu:94-211               build loop + torque radius
//   KPR/CollisionChecking.cu:26-39,136-299  pair table, bufferObstacles, polytope_PH, checkCollision
//   KPR/NLPclass.cu:30-538                  TNLP callbacks
//
// PARITY: pinned against the reference's own sources compiled here against stand-in Eigen / Boost / Ipopt headers
// (oracle/_ref, tests/test_reference_pin.py); the ARMTD comparison mode (build_armtd) is pinned by properties only.
// See oracle/README.md.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.
#include "oracle_pz.hpp"
#include <chrono>
#include <memory>
#include <omp.h>

using namespace orc;

namespace {

// ---- robot constants: KPR/KinovaWithoutGripperInfo.h ------------------------------------
const int axes[NJ] = {3, 3, 3, 3, 3, 3, 3};
const double trans[(NJ + 1) * 3] = {0, 0, 0.15643,  0, 0.005375, -0.12838,  0, -0.21038, -0.006375,  0, 0.006375, -0.21038,
                                    0, -0.20843, -0.006375,  0, 0.00017505, -0.10593,  0, -0.10593, -0.00017505,  0, 0, 0};
const double rots[NJ * 3] = {M_PI, 0, 0,  M_PI * 0.5, 0, 0,  -M_PI * 0.5, 0, 0,  M_PI * 0.5, 0, 0,  -M_PI * 0.5, 0, 0,  M_PI * 0.5, 0, 0,  -M_PI * 0.5, 0, 0};
const double mass[NJ] = {1.3773, 1.1636, 1.1636, 0.9302, 0.6781, 0.6781, 0.5};
const double com[NJ * 3] = {-0.000023, -0.010364, -0.07336,  -0.000044, -0.09958, -0.013278,  -0.000044, -0.006641, -0.117892,
                            -0.000018, -0.075478, -0.015006,  0.000001, -0.009432, -0.063883,  0.000001, -0.045483, -0.00965,
                            0.000281, 0.011402, -0.029798};
const double inertia[NJ * 9] = {
    0.00457, 0.000001, 0.000002, 0.000001, 0.004831, 0.000448, 0.000002, 0.000448, 0.001409,
    0.011088, 0.000005, 0, 0.000005, 0.001072, -0.000691, 0, -0.000691, 0.011255,
    0.010932, 0, -0.000007, 0, 0.011127, 0.000606, -0.000007, 0.000606, 0.001043,
    0.008147, -0.000001, 0, -0.000001, 0.000631, -0.0005, 0, -0.0005, 0.008316,
    0.001596, 0, 0, 0, 0.001607, 0.000256, 0, 0.000256, 0.000399,
    0.001641, 0, 0, 0, 0.00041, -0.000278, 0, -0.000278, 0.001641,
    0.000587, 0.000003, 0.000003, 0.000003, 0.000369, -0.000118, 0.000003, -0.000118, 0.000609};
const double friction[NJ] = {0.0};
const double damping[NJ] = {0.0};
const double armature[NJ] = {8.03, 11.9962024615303644, 9.0025427861751517, 11.5806439316706360, 8.4665040917914123, 8.8537069373742430, 8.8587303664685315};
const double state_limits_lb[NF] = {-1000.0, -2.41, -1000.0, -2.66, -1000.0, -2.23, -1000.0};
const double state_limits_ub[NF] = {1000.0, 2.41, 1000.0, 2.66, 1000.0, 2.23, 1000.0};
const double speed_limits[NF] = {1.3963, 1.3963, 1.3963, 1.3963, 1.2218, 1.2218, 1.2218};
const double torque_limits[NF] = {56.7, 56.7, 56.7, 56.7, 29.4, 29.4, 29.4};
const double gravity = 9.81;
const double link_zonotope_center[NJ][3] = {{0.000000, -0.001297, -0.088375}, {0.000000, -0.089400, -0.007877}, {0.000000, -0.001502, -0.129375},
                                            {0.000000, -0.087450, -0.013648}, {0.000001, -0.009023, -0.071752}, {0.000000, -0.041661, -0.009251},
                                            {0.000000, -0.018585, -0.033462}};
const double link_zonotope_generators[NJ][3] = {{0.046358, 0.047354, 0.086000}, {0.046000, 0.135400, 0.047501}, {0.046000, 0.047501, 0.127000},
                                                {0.046000, 0.133450, 0.042293}, {0.034999, 0.044023, 0.069252}, {0.035000, 0.076739, 0.044076},
                                                {0.045500, 0.056085, 0.030963}};
const double alpha_ub = 10.0, V_m = 1e-2, M_max = 15.79635774, M_min = 5.095620491878957;
const double eps_ub = std::sqrt(2 * V_m / M_min);
const double K_ub = 5.0;
const double qe = eps_ub / K_ub, qde = 2 * eps_ub, qdae = eps_ub, qddae = 2 * K_ub * eps_ub;

const double DURATION = 1.0;                                  // KPR/Parameters.h:14
const double QDD_DES_K_DEP_MAXIMA = (0.5 - std::sqrt(3) / 6);  // KPR/Trajectory.h:7-8
const double QDD_DES_K_DEP_MINIMA = (0.5 + std::sqrt(3) / 6);
const int OBS_GEN = 3, BUF_GEN = OBS_GEN + 6, COMB = BUF_GEN * (BUF_GEN - 1) / 2;   // KPR/CollisionChecking.h:6-7
const double COLLISION_THRESHOLD = 1e-4, TORQUE_THRESHOLD = 1e-2, COST_SCALE = 10.0;   // KPR/Parameters.h:38-44

using std::pow; using std::sqrt; using std::min; using std::max; using std::swap;

// ---- Bezier helpers: KPR/Trajectory.cu:542-599, 812-822 -----------------------------------
double q_des_func(double q0, double Tqd0, double TTqdd0, double k, double t) {
    double B0 = -pow(t - 1, 5), B1 = 5 * t * pow(t - 1, 4), B2 = -10 * pow(t, 2) * pow(t - 1, 3);
    double B3 = 10 * pow(t, 3) * pow(t - 1, 2), B4 = -5 * pow(t, 4) * (t - 1), B5 = pow(t, 5);
    double b0 = q0, b1 = q0 + Tqd0 / 5, b2 = q0 + (2 * Tqd0) / 5 + TTqdd0 / 20, b3 = q0 + k, b4 = q0 + k, b5 = q0 + k;
    return B0 * b0 + B1 * b1 + B2 * b2 + B3 * b3 + B4 * b4 + B5 * b5;
}
double qd_des_func(double q0, double Tqd0, double TTqdd0, double k, double t) {
    double dB0 = pow(t - 1.0, 4.0) * -5.0;
    double dB1 = t * pow(t - 1.0, 3.0) * 2.0E+1 + pow(t - 1.0, 4.0) * 5.0;
    double dB2 = t * pow(t - 1.0, 3.0) * -2.0E+1 - (t * t) * pow(t - 1.0, 2.0) * 3.0E+1;
    double dB3 = pow(t, 3.0) * (t * 2.0 - 2.0) * 1.0E+1 + (t * t) * pow(t - 1.0, 2.0) * 3.0E+1;
    double dB4 = pow(t, 3.0) * (t - 1.0) * -2.0E+1 - pow(t, 4.0) * 5.0;
    double dB5 = pow(t, 4.0) * 5.0;
    double b0 = q0, b1 = q0 + Tqd0 / 5, b2 = q0 + (2 * Tqd0) / 5 + TTqdd0 / 20, b3 = q0 + k, b4 = q0 + k, b5 = q0 + k;
    return dB0 * b0 + dB1 * b1 + dB2 * b2 + dB3 * b3 + dB4 * b4 + dB5 * b5;
}
double qdd_des_func(double q0, double Tqd0, double TTqdd0, double k, double t) {
    double t2 = t * 2.0, t3 = t * t, t4 = t * t * t, t5 = t - 1.0, t6 = t2 - 2.0, t7 = t4 * 2.0E+1, t8 = t5 * t5, t9 = t5 * t5 * t5;
    double t10 = t9 * 2.0E+1, t11 = t * t8 * 6.0E+1, t12 = -t10;
    double ddB0 = t12, ddB1 = t9 * 4.0E+1 + t11, ddB2 = t12 - t * t8 * 1.2E+2 - t3 * t6 * 3.0E+1;
    double ddB3 = t7 + t11 + t3 * t6 * 6.0E+1, ddB4 = t4 * -4.0E+1 - t3 * t5 * 6.0E+1, ddB5 = t7;
    double b0 = q0, b1 = q0 + Tqd0 / 5, b2 = q0 + (2 * Tqd0) / 5 + TTqdd0 / 20, b3 = q0 + k, b4 = q0 + k, b5 = q0 + k;
    return ddB0 * b0 + ddB1 * b1 + ddB2 * b2 + ddB3 * b3 + ddB4 * b4 + ddB5 * b5;
}
double q_des_k_indep(double q0, double Tqd0, double TTqdd0, double s) {
    return q0 + Tqd0 * s - 6 * Tqd0 * pow(s, 3) + 8 * Tqd0 * pow(s, 4) - 3 * Tqd0 * pow(s, 5) + (TTqdd0 * pow(s, 2)) * 0.5 - (3 * TTqdd0 * pow(s, 3)) * 0.5 +
           (3 * TTqdd0 * pow(s, 4)) * 0.5 - (TTqdd0 * pow(s, 5)) * 0.5;
}
double qd_des_k_indep(double, double Tqd0, double TTqdd0, double s) {
    return (pow(s - 1, 2) * (2 * Tqd0 + 4 * Tqd0 * s + 2 * TTqdd0 * s - 30 * Tqd0 * pow(s, 2) - 5 * TTqdd0 * pow(s, 2))) * 0.5 / DURATION;
}
double qdd_des_k_indep(double, double Tqd0, double TTqdd0, double s) {
    return -(s - 1.0) * (TTqdd0 - (36 * Tqd0 + 8 * TTqdd0) * s + (60 * Tqd0 + 10 * TTqdd0) * pow(s, 2)) / (DURATION * DURATION);
}

// ---- generated k-derivatives of the interior extrema: KPR/Trajectory.cu:601-810 -----------
// The reference's expressions are the symbolic total derivative d/dk q_des(t*(k), k) (resp. qd_des)
// expanded over the Bernstein form above; they are restated here term for term.  `sg` = +1 for the
// "+sqrt" stationary point (extrema2), -1 for the "-sqrt" one (extrema3).
double q_des_extrema_k_derivative(double q0, double Tqd0, double TTqdd0, double k, int sg) {
    const double qk = k + q0;                       // beta3..5
    const double b1 = q0 + Tqd0 / 5.0;              // beta1
    const double b2 = q0 + Tqd0 * (2.0 / 5.0) + TTqdd0 / 2.0E+1;   // beta2
    const double den = TTqdd0 + Tqd0 * 6.0 + (-(k * 1.2E+1));
    const double disc = TTqdd0 * TTqdd0 + Tqd0 * TTqdd0 * 1.4E+1 + (Tqd0 * Tqd0) * 6.4E+1 + (-(k * Tqd0 * 1.2E+2));
    const double i1 = 1.0 / den, rt = sqrt(disc), ir = 1.0 / rt;
    const double i2 = i1 * i1, i3 = i1 * i1 * i1, i5 = i1 * i1 * i1 * i1 * i1, i4 = i2 * i2;
    const double num = (sg > 0) ? (TTqdd0 + Tqd0 * 2.0 + rt) : (TTqdd0 + Tqd0 * 2.0 + (-rt));
    const double n2 = num * num, n3 = num * num * num, n5 = num * num * num * num * num, n4 = n2 * n2;
    const double a = Tqd0 * i1 * ir * 1.2E+1;          // |d sqrt-part / dk| / (5 den)
    const double b = i2 * num * (1.2E+1 / 5.0);
    const double tm1 = (i1 * num) / 5.0 - 1.0;          // t* - 1
    const double m2 = tm1 * tm1, m3 = tm1 * tm1 * tm1, m4 = m2 * m2;
    if (sg > 0) {
        const double d = a - b;                         // = -dt*/dk
        return (i5 * n5) / 3.125E+3 + qk * (i2 * i2 * i2) * n5 * (1.2E+1 / 6.25E+2) + i3 * n3 * m2 * (2.0 / 2.5E+1) - (i4 * n4 * tm1) / 1.25E+2 +
               q0 * m4 * d * 5.0 + qk * i4 * n3 * m2 * (7.2E+1 / 2.5E+1) - qk * i5 * n4 * tm1 * (4.8E+1 / 1.25E+2) + b1 * i2 * num * m4 * 1.2E+1 -
               b2 * i3 * n2 * m3 * (4.8E+1 / 5.0) + (qk * i4 * n4 * d) / 1.25E+2 - Tqd0 * qk * i5 * ir * n4 * (1.2E+1 / 1.25E+2) -
               Tqd0 * b1 * i1 * ir * m4 * 6.0E+1 - qk * i3 * n3 * tm1 * d * (4.0 / 2.5E+1) - b1 * i1 * num * m3 * d * 4.0 +
               b2 * i2 * n2 * m2 * d * (6.0 / 5.0) - Tqd0 * qk * i3 * ir * n2 * m2 * (7.2E+1 / 5.0) + Tqd0 * qk * i4 * ir * n3 * tm1 * (4.8E+1 / 2.5E+1) +
               Tqd0 * b2 * i2 * ir * num * m3 * 4.8E+1;
    }
    const double d = a + b;                             // = +dt*/dk
    return (i5 * n5) / 3.125E+3 + qk * (i2 * i2 * i2) * n5 * (1.2E+1 / 6.25E+2) - q0 * m4 * d * 5.0 + i3 * n3 * m2 * (2.0 / 2.5E+1) -
           (i4 * n4 * tm1) / 1.25E+2 + qk * i4 * n3 * m2 * (7.2E+1 / 2.5E+1) - qk * i5 * n4 * tm1 * (4.8E+1 / 1.25E+2) - (qk * i4 * n4 * d) / 1.25E+2 +
           b1 * i2 * num * m4 * 1.2E+1 - b2 * i3 * n2 * m3 * (4.8E+1 / 5.0) + Tqd0 * qk * i5 * ir * n4 * (1.2E+1 / 1.25E+2) +
           Tqd0 * b1 * i1 * ir * m4 * 6.0E+1 + qk * i3 * n3 * tm1 * d * (4.0 / 2.5E+1) + b1 * i1 * num * m3 * d * 4.0 -
           b2 * i2 * n2 * m2 * d * (6.0 / 5.0) + Tqd0 * qk * i3 * ir * n2 * m2 * (7.2E+1 / 5.0) - Tqd0 * qk * i4 * ir * n3 * tm1 * (4.8E+1 / 2.5E+1) -
           Tqd0 * b2 * i2 * ir * num * m3 * 4.8E+1;
}
double qd_des_extrema_k_derivative(double q0, double Tqd0, double TTqdd0, double k, int sg) {
    const double qk = k + q0;
    const double b1 = q0 + Tqd0 / 5.0;
    const double b2 = q0 + Tqd0 * (2.0 / 5.0) + TTqdd0 / 2.0E+1;
    const double s6 = sqrt(6.0);
    const double den = TTqdd0 + Tqd0 * 6.0 + (-(k * 1.2E+1));
    const double lin = TTqdd0 * 2.0E+1 + Tqd0 * 1.8E+2 + (-(k * 3.0E+2));
    const double disc = TTqdd0 * TTqdd0 + Tqd0 * TTqdd0 * 1.4E+1 + (Tqd0 * Tqd0) * 5.4E+1 + k * TTqdd0 * -2.0E+1 + (k * k) * 1.5E+2 + k * Tqd0 * -1.8E+2;
    const double i1 = 1.0 / den, rt = sqrt(disc), ir = 1.0 / rt;
    const double i2 = i1 * i1, i3 = i1 * i1 * i1, i5 = i1 * i1 * i1 * i1 * i1, i4 = i2 * i2;
    const double sr = s6 * rt;
    if (sg > 0) {
        const double num = TTqdd0 * 4.0 + Tqd0 * 1.8E+1 + (-(k * 3.0E+1)) + sr;
        const double h = (s6 * lin * ir) / 2.0;
        const double n2 = num * num, n3 = num * num * num, n4 = n2 * n2;
        const double e = h + 3.0E+1;
        const double u5 = (i1 * num) / 5.0, g = i2 * num * (6.0 / 5.0), u10 = (i1 * num) / 1.0E+1;
        const double w = u5 - 2.0, tm1 = u10 - 1.0;
        const double f = (i1 * e) / 1.0E+1;
        const double m2 = tm1 * tm1, m3 = tm1 * tm1 * tm1;
        const double A = i2 * num * m3 * 2.4E+1;
        const double B = i3 * n2 * m2 * (3.6E+1 / 5.0);
        const double dd = g + (-f);
        const double C = -(i1 * e * m3 * 2.0);
        const double D = -(i2 * num * e * m2 * (3.0 / 5.0));
        const double E = i1 * num * m2 * dd * 6.0;
        const double F = i2 * n2 * tm1 * dd * (3.0 / 5.0);
        return b1 * (A + C + E + m3 * dd * 2.0E+1) +
               qk * (B + D + F + (i3 * n3 * (i2 * num * (1.2E+1 / 5.0) - (i1 * e) / 5.0)) / 1.0E+2 + i4 * n3 * w * (9.0 / 2.5E+1) - i3 * n2 * e * w * (3.0 / 1.0E+2)) -
               qk * (i5 * n4 * (3.0 / 1.25E+2) - (i4 * n3 * e) / 5.0E+2 + i4 * n3 * tm1 * (1.8E+1 / 2.5E+1) + (i3 * n3 * dd) / 5.0E+1 - i3 * n2 * e * tm1 * (3.0 / 5.0E+1)) -
               b2 * (A + B + C + D + E + F) - q0 * m3 * dd * 2.0E+1 + qk * i5 * n4 * (3.0 / 1.25E+2) + i2 * n2 * m2 * (3.0 / 1.0E+1) + (i3 * n3 * w) / 1.0E+2 -
               (i3 * n3 * tm1) / 5.0E+1 - (qk * i4 * n3 * e) / 5.0E+2;
    }
    // "-sqrt" branch: the generated code works with P = 4*TTqdd0 - 30k + 18*Tqd0 - sqrt(6)*sqrt(disc)
    const double P = TTqdd0 * 4.0 - k * 3.0E+1 + Tqd0 * 1.8E+1 - sr;
    const double P2 = pow(P, 2.0), P3 = pow(P, 3.0), P4 = P2 * P2;
    const double h = (s6 * lin * ir) / 2.0;
    const double e = h - 3.0E+1;
    const double tm1 = (i1 * P) / 1.0E+1 - 1.0;
    const double w = (i1 * P) / 5.0 - 2.0;
    const double m2 = pow(tm1, 2.0), m3 = pow(tm1, 3.0);
    const double f = (i1 * e) / 1.0E+1;
    const double dd = f + i2 * P * (6.0 / 5.0);
    const double A = i2 * m3 * P * 2.4E+1;
    const double B = i3 * P2 * m2 * (3.6E+1 / 5.0);
    const double C = i1 * e * m3 * 2.0;
    const double D = i2 * e * m2 * P * (3.0 / 5.0);
    const double E = i1 * m2 * dd * P * 6.0;
    const double F = i2 * P2 * tm1 * dd * (3.0 / 5.0);
    return b1 * (A + C + E + m3 * dd * 2.0E+1) +
           qk * (B + D + F + i4 * w * P3 * (9.0 / 2.5E+1) + (i3 * ((i1 * e) / 5.0 + i2 * P * (1.2E+1 / 5.0)) * P3) / 1.0E+2 + i3 * P2 * e * w * (3.0 / 1.0E+2)) -
           qk * (i5 * P4 * (3.0 / 1.25E+2) + i4 * tm1 * P3 * (1.8E+1 / 2.5E+1) + (i4 * e * P3) / 5.0E+2 + (i3 * dd * P3) / 5.0E+1 + i3 * P2 * e * tm1 * (3.0 / 5.0E+1)) -
           b2 * (A + B + C + D + E + F) + (i3 * w * P3) / 1.0E+2 - (i3 * tm1 * P3) / 5.0E+1 + qk * i5 * P4 * (3.0 / 1.25E+2) + i2 * P2 * m2 * (3.0 / 1.0E+1) -
           q0 * m3 * dd * 2.0E+1 + (qk * i4 * e * P3) / 5.0E+2;
}

// ---- plan state ------------------------------------------------------------------------------
struct Plan {
    // configuration (runtime; compile-time macros in the reference)
    int T = 128;
    double k_range[NF];
    double mass_uncertainty = 0.03, inertia_uncertainty = 0.03;
    double threshold = 5e-4;
    int num_threads = 0;

    // BezierCurve state: KPR/Trajectory.h:32-55
    double q0[NF], qd0[NF], qdd0[NF], Tqd0[NF], TTqdd0[NF];
    double q_ext[2][NF], q_extv[2][NF], qd_ext[2][NF], qd_extv[2][NF], qdd_ext[2][NF], qdd_extv[2][NF];
    double ds = 0;
    std::vector<PZ> cos_q, sin_q, R, R_t, qd_des, qda_des, qdda_des;   // [joint * T + s]
    // raw Boost intervals of the Taylor remainders, exported for containment tests
    std::vector<double> cos_rem, sin_rem;   // [(i*T+s)*2 + {lo,hi}]

    // KinematicsDynamics state: KPR/Dynamics.h:6-31
    Mat trans_matrix[NJ + 1], com_matrix[NJ];
    PZ mass_nominal[NJ], mass_uncertain[NJ], I_nominal[NJ], I_uncertain[NJ];
    PZ link0[NJ];
    std::vector<PZ> links, u_nom, u_nom_int;
    std::vector<Mat> link_gens;   // [s * NJ + i], 3x6
    std::vector<double> torque_radius;   // (j, s) -> [j + s * NF]

    // Obstacles: KPR/CollisionChecking.h:14-46
    int n_obs = 0;
    std::vector<double> obstacles;   // n_obs * 12
    std::vector<double> A, d, delta;   // [((t*NJ+link)*n_obs+obs)*COMB + p] (A has 3 per entry)

    // NLP scratch
    std::vector<Mat> link_sliced_center;      // [t*NJ + l]
    std::vector<Mat> dk_link_sliced_center;   // [(t*NJ + l)*NF + k]

    // second trajectory family: ARMTD comparison planner (KPA = kinova_planner_realtime_armtd_comparison)
    int mode = 0;                 // 0 ARMOUR (Bezier + RNEA), 1 ARMTD (constant acceleration, offline JRS tables, FK only)
    std::vector<double> jrs;      // [6][NF][T]: c_cos, g_cos, r_cos, c_sin, g_sin, r_sin (KPA/armtd_main.cu:41-47)

    OpStats stats;
    double build_ms = 0;

    PZ& at(std::vector<PZ>& a, int i, int s) { return a[(size_t)i * T + s]; }
};

void bezier_init(Plan& p, const double* q0, const double* qd0, const double* qdd0) {   // KPR/Trajectory.cu:15-61
    const int T = p.T;
    for (int i = 0; i < NF; i++) {
        p.q0[i] = q0[i]; p.qd0[i] = qd0[i]; p.qdd0[i] = qdd0[i];
        p.Tqd0[i] = qd0[i] * DURATION;
        p.TTqdd0[i] = qdd0[i] * DURATION * DURATION;
    }
    p.cos_q.assign((size_t)NF * T, PZ()); p.sin_q.assign((size_t)NF * T, PZ());
    p.R.assign((size_t)(NJ + 1) * T, PZ()); p.R_t.assign((size_t)NJ * T, PZ());
    p.qd_des.assign((size_t)NF * T, PZ()); p.qda_des.assign((size_t)NF * T, PZ()); p.qdda_des.assign((size_t)NF * T, PZ());
    p.cos_rem.assign((size_t)NF * T * 2, 0.0); p.sin_rem.assign((size_t)NF * T * 2, 0.0);
    const double *a = p.Tqd0, *b = p.TTqdd0;
    for (int i = 0; i < NF; i++) {
        p.q_ext[0][i] = (2 * a[i] + b[i] + sqrt(64 * pow(a[i], 2) + 14 * a[i] * b[i] + pow(b[i], 2))) / (5 * (6 * a[i] + b[i]));
        p.q_ext[1][i] = (2 * a[i] + b[i] - sqrt(64 * pow(a[i], 2) + 14 * a[i] * b[i] + pow(b[i], 2))) / (5 * (6 * a[i] + b[i]));
        p.q_extv[0][i] = q_des_k_indep(q0[i], a[i], b[i], p.q_ext[0][i]);
        p.q_extv[1][i] = q_des_k_indep(q0[i], a[i], b[i], p.q_ext[1][i]);
    }
    for (int i = 0; i < NF; i++) {
        p.qd_ext[0][i] = (18 * a[i] + 4 * b[i] + sqrt(6 * (54 * pow(a[i], 2) + 14 * a[i] * b[i] + pow(b[i], 2)))) / (10 * (6 * a[i] + b[i]));
        p.qd_ext[1][i] = (18 * a[i] + 4 * b[i] - sqrt(6 * (54 * pow(a[i], 2) + 14 * a[i] * b[i] + pow(b[i], 2)))) / (10 * (6 * a[i] + b[i]));
        p.qd_extv[0][i] = qd_des_k_indep(q0[i], a[i], b[i], p.qd_ext[0][i]);
        p.qd_extv[1][i] = qd_des_k_indep(q0[i], a[i], b[i], p.qd_ext[1][i]);
    }
    for (int i = 0; i < NF; i++) {
        p.qdd_ext[0][i] = (32 * a[i] + 6 * b[i] + sqrt(2 * (152 * pow(a[i], 2) + 42 * a[i] * b[i] + 3 * pow(b[i], 2)))) / (10 * (6 * a[i] + b[i]));
        p.qdd_ext[1][i] = (32 * a[i] + 6 * b[i] - sqrt(2 * (152 * pow(a[i], 2) + 42 * a[i] * b[i] + 3 * pow(b[i], 2)))) / (10 * (6 * a[i] + b[i]));
        p.qdd_extv[0][i] = qdd_des_k_indep(q0[i], a[i], b[i], p.qdd_ext[0][i]);
        p.qdd_extv[1][i] = qdd_des_k_indep(q0[i], a[i], b[i], p.qdd_ext[1][i]);
    }
    p.ds = 1.0 / T;
}

// bound the k-independent part over [s_lb, s_ub]: endpoints + interior stationary points
void bound_k_indep(double lb_val, double ub_val, double s_lb, double s_ub, double e1, double v1, double e2, double v2, double& lo, double& hi) {
    lo = lb_val; hi = ub_val;
    if (lo > hi) swap(lo, hi);
    if (s_lb < e1 && e1 < s_ub) { lo = min(lo, v1); hi = max(hi, v1); }
    if (s_lb < e2 && e2 < s_ub) { lo = min(lo, v2); hi = max(hi, v2); }
}

void makePolyZono(Plan& p, int s_ind) {   // KPR/Trajectory.cu:63-254
    assert(s_ind < p.T);
    const int T = p.T;
    const double s_lb = s_ind * p.ds, s_ub = (s_ind + 1) * p.ds;
    for (int i = 0; i < NF; i++) {
        const double kr = p.k_range[i];
        const double q0 = p.q0[i], a = p.Tqd0[i], b = p.TTqdd0[i];
        // Part 1: q_des
        double kd_lb = pow(s_lb, 3) * (6 * pow(s_lb, 2) - 15 * s_lb + 10);
        double kd_ub = pow(s_ub, 3) * (6 * pow(s_ub, 2) - 15 * s_ub + 10);
        double kd_center = (kd_ub + kd_lb) * 0.5;
        double kd_radius = (kd_ub - kd_lb) * 0.5 * kr;
        double ki_lb, ki_ub;
        bound_k_indep(q_des_k_indep(q0, a, b, s_lb), q_des_k_indep(q0, a, b, s_ub), s_lb, s_ub, p.q_ext[0][i], p.q_extv[0][i], p.q_ext[1][i], p.q_extv[1][i], ki_lb, ki_ub);
        double ki_radius = (ki_ub - ki_lb) * 0.5;
        double q_des_center = (ki_lb + ki_ub) * 0.5;
        Interval q_rad(-kd_radius - ki_radius - qe, kd_radius + ki_radius + qe);

        // Part 1.a: cos(q_des), first-order Taylor + Lagrange remainder
        double cos_center = std::cos(q_des_center);
        Interval cos_rad = -q_rad * std::sin(q_des_center)
                           - 0.5 * cos(q_des_center + kd_center * Interval(-kr, kr) + q_rad) * pow2(q_rad + kd_center * Interval(-kr, kr));
        p.cos_rem[((size_t)i * T + s_ind) * 2] = cos_rad.lo; p.cos_rem[((size_t)i * T + s_ind) * 2 + 1] = cos_rad.hi;
        cos_center += getCenter(cos_rad);
        cos_rad = cos_rad - getCenter(cos_rad);
        double cos_coeff[] = {-kd_center * kr * std::sin(q_des_center), getRadius(cos_rad)};
        uint64_t cos_deg[2][NVAR] = {{0}};
        cos_deg[0][i] = 1; cos_deg[1][i + NF * 4] = 1;
        p.at(p.cos_q, i, s_ind) = PZ(cos_center, cos_coeff, cos_deg, 2);

        // Part 1.b: sin(q_des)
        double sin_center = std::sin(q_des_center);
        Interval sin_rad = q_rad * std::cos(q_des_center)
                           - 0.5 * sin(q_des_center + kd_center * Interval(-kr, kr) + q_rad) * pow2(q_rad + kd_center * Interval(-kr, kr));
        p.sin_rem[((size_t)i * T + s_ind) * 2] = sin_rad.lo; p.sin_rem[((size_t)i * T + s_ind) * 2 + 1] = sin_rad.hi;
        sin_center += getCenter(sin_rad);
        sin_rad = sin_rad - getCenter(sin_rad);
        double sin_coeff[] = {kd_center * kr * std::cos(q_des_center), getRadius(sin_rad)};
        uint64_t sin_deg[2][NVAR] = {{0}};
        sin_deg[0][i] = 1; sin_deg[1][i + NF * 5] = 1;
        p.at(p.sin_q, i, s_ind) = PZ(sin_center, sin_coeff, sin_deg, 2);

        p.at(p.R, i, s_ind) = PZ(rots[i * 3], rots[i * 3 + 1], rots[i * 3 + 2]);
        if (axes[i] != 0)
            p.at(p.R, i, s_ind) = p.at(p.R, i, s_ind) * PZ(cos_center, cos_coeff, cos_deg, 2, sin_center, sin_coeff, sin_deg, 2, axes[i]);
        p.at(p.R_t, i, s_ind) = p.at(p.R, i, s_ind).transpose();

        // Part 2: qd_des
        kd_lb = (30 * pow(s_lb, 2) * pow(s_lb - 1, 2)) / DURATION;
        kd_ub = (30 * pow(s_ub, 2) * pow(s_ub - 1, 2)) / DURATION;
        if (kd_ub < kd_lb) swap(kd_lb, kd_ub);
        kd_center = (kd_ub + kd_lb) * 0.5 * kr;
        kd_radius = (kd_ub - kd_lb) * 0.5 * kr;
        bound_k_indep(qd_des_k_indep(q0, a, b, s_lb), qd_des_k_indep(q0, a, b, s_ub), s_lb, s_ub, p.qd_ext[0][i], p.qd_extv[0][i], p.qd_ext[1][i], p.qd_extv[1][i], ki_lb, ki_ub);
        ki_radius = (ki_ub - ki_lb) * 0.5;
        double qd_des_center = (ki_lb + ki_ub) * 0.5;
        double qd_coeff[] = {kd_center, kd_radius + ki_radius + qde};
        uint64_t qd_deg[2][NVAR] = {{0}};
        qd_deg[0][i] = 1; qd_deg[1][i + NF * 1] = 1;
        p.at(p.qd_des, i, s_ind) = PZ(qd_des_center, qd_coeff, qd_deg, 2);
        double qda_coeff[] = {kd_center, kd_radius + ki_radius + qdae};
        uint64_t qda_deg[2][NVAR] = {{0}};
        qda_deg[0][i] = 1; qda_deg[1][i + NF * 2] = 1;
        p.at(p.qda_des, i, s_ind) = PZ(qd_des_center, qda_coeff, qda_deg, 2);

        // Part 3: qdd_des
        double t_lb = (60 * s_lb * (2 * pow(s_lb, 2) - 3 * s_lb + 1)) / DURATION / DURATION;
        double t_ub = (60 * s_ub * (2 * pow(s_ub, 2) - 3 * s_ub + 1)) / DURATION / DURATION;
        if (s_ub <= QDD_DES_K_DEP_MAXIMA) { kd_lb = t_lb; kd_ub = t_ub; }
        else if (s_lb <= QDD_DES_K_DEP_MAXIMA) {
            kd_lb = min(t_lb, t_ub);
            kd_ub = (60 * QDD_DES_K_DEP_MAXIMA * (2 * pow(QDD_DES_K_DEP_MAXIMA, 2) - 3 * QDD_DES_K_DEP_MAXIMA + 1)) / DURATION / DURATION;
        }
        else if (s_ub <= QDD_DES_K_DEP_MINIMA) { kd_lb = t_ub; kd_ub = t_lb; }
        else if (s_lb <= QDD_DES_K_DEP_MINIMA) {
            kd_lb = (60 * QDD_DES_K_DEP_MINIMA * (2 * pow(QDD_DES_K_DEP_MINIMA, 2) - 3 * QDD_DES_K_DEP_MINIMA + 1)) / DURATION / DURATION;
            kd_ub = max(t_lb, t_ub);
        }
        else { kd_lb = t_lb; kd_ub = t_ub; }
        kd_center = (kd_ub + kd_lb) * 0.5 * kr;
        kd_radius = (kd_ub - kd_lb) * 0.5 * kr;
        bound_k_indep(qdd_des_k_indep(q0, a, b, s_lb), qdd_des_k_indep(q0, a, b, s_ub), s_lb, s_ub, p.qdd_ext[0][i], p.qdd_extv[0][i], p.qdd_ext[1][i], p.qdd_extv[1][i], ki_lb, ki_ub);
        ki_radius = (ki_ub - ki_lb) * 0.5;
        double qdd_des_center = (ki_lb + ki_ub) * 0.5;
        double qdd_coeff[] = {kd_center, kd_radius + ki_radius + qddae};
        uint64_t qdd_deg[2][NVAR] = {{0}};
        qdd_deg[0][i] = 1; qdd_deg[1][i + NF * 3] = 1;
        p.at(p.qdda_des, i, s_ind) = PZ(qdd_des_center, qdd_coeff, qdd_deg, 2);
    }
    for (int i = NF; i < NJ; i++) {
        p.at(p.R, i, s_ind) = PZ(rots[i * 3], rots[i * 3 + 1], rots[i * 3 + 2]);
        p.at(p.R_t, i, s_ind) = p.at(p.R, i, s_ind).transpose();
    }
    p.at(p.R, NJ, s_ind) = PZ(0.0, 0.0, 0.0);
}

void kd_init(Plan& p) {   // KPR/Dynamics.cu:6-67
    const int T = p.T;
    p.links.assign((size_t)NJ * T, PZ());
    p.u_nom.assign((size_t)NF * T, PZ());
    p.u_nom_int.assign((size_t)NF * T, PZ());
    for (int i = 0; i < NJ; i++) {
        p.trans_matrix[i] = Mat::Zero(3, 1);
        for (int a = 0; a < 3; a++) p.trans_matrix[i](a) = trans[3 * i + a];
        p.com_matrix[i] = Mat::Zero(3, 1);
        for (int a = 0; a < 3; a++) p.com_matrix[i](a) = com[3 * i + a];
        Mat m(1, 1); m(0) = mass[i];
        p.mass_nominal[i] = PZ(m);
        p.mass_uncertain[i] = PZ(m, p.mass_uncertainty);
        Mat I(3, 3);
        for (int j = 0; j < 9; j++) I(j) = inertia[i * 9 + j];
        p.I_nominal[i] = PZ(I);
        p.I_uncertain[i] = PZ(I, p.inertia_uncertainty);
    }
    p.trans_matrix[NJ] = Mat::Zero(3, 1);
    for (int a = 0; a < 3; a++) p.trans_matrix[NJ](a) = trans[3 * NJ + a];
    for (int i = 0; i < NJ; i++) {
        PZ link[3];
        for (int j = 0; j < 3; j++) {
            uint64_t deg[1][NVAR] = {{0}};
            deg[0][NF * (j + 1)] = 1;   // qde_0, qdae_0, qddae_0 stand for the x, y, z box generators
            double g = link_zonotope_generators[i][j];
            link[j] = PZ(link_zonotope_center[i][j], &g, deg, 1);
        }
        p.link0[i] = stack3(link);
        for (int s = 0; s < T; s++) p.at(p.links, i, s) = p.link0[i];
    }
}

void fk(Plan& p, int s) {   // KPR/Dynamics.cu:69-81
    PZ FK_R = PZ(0.0, 0.0, 0.0);
    PZ FK_T(3, 1);
    for (int i = 0; i < NJ; i++) {
        PZ P(p.trans_matrix[i]);
        FK_T = FK_T + FK_R * P;
        FK_R = FK_R * p.at(p.R, i, s);
        p.at(p.links, i, s) = FK_R * p.at(p.links, i, s) + FK_T;
    }
}

void rnea(Plan& p, int s, PZ* mass_arr, PZ* I_arr, std::vector<PZ>& u, bool setGravity = true) {   // KPR/Dynamics.cu:83-181
    PZ w(3, 1), wdot(3, 1), w_aux(3, 1), linear_acc(3, 1);
    PZ F[NJ], N[NJ];
    if (setGravity) linear_acc.center(2) = gravity;
    for (int i = 0; i < NJ; i++) {
        PZ& Rt = p.at(p.R_t, i, s);
        if (axes[i] != 0) {
            linear_acc = Rt * (linear_acc + cross(wdot, p.trans_matrix[i]) + cross(w, cross(w_aux, p.trans_matrix[i])));
            w = Rt * w;
            w.addOneDimPZ(p.at(p.qd_des, i, s), std::abs(axes[i]) - 1, 0);
            w_aux = Rt * w_aux;
            wdot = Rt * wdot;
            PZ temp(3, 1);
            temp.addOneDimPZ(p.at(p.qd_des, i, s), std::abs(axes[i]) - 1, 0);
            wdot = wdot + cross(w_aux, temp);
            wdot.addOneDimPZ(p.at(p.qdda_des, i, s), std::abs(axes[i]) - 1, 0);
            w_aux.addOneDimPZ(p.at(p.qda_des, i, s), std::abs(axes[i]) - 1, 0);
        }
        else {
            linear_acc = Rt * (linear_acc + cross(wdot, p.trans_matrix[i]) + cross(w, cross(w_aux, p.trans_matrix[i])));
            w = Rt * w;
            w_aux = Rt * w_aux;
            wdot = Rt * wdot;
        }
        F[i] = mass_arr[i] * (linear_acc + cross(wdot, p.com_matrix[i]) + cross(w, cross(w_aux, p.com_matrix[i])));
        N[i] = I_arr[i] * wdot + cross(w_aux, (I_arr[i] * w));
    }
    PZ f(3, 1), n(3, 1);
    for (int i = NJ - 1; i >= 0; i--) {
        PZ& Rn = p.at(p.R, i + 1, s);
        n = N[i] + Rn * n + cross(p.com_matrix[i], F[i]) + cross(p.trans_matrix[i + 1], Rn * f);
        f = Rn * f + F[i];
        if (axes[i] != 0) {
            u[(size_t)i * p.T + s] = n(std::abs(axes[i]) - 1, 0);
            u[(size_t)i * p.T + s] = u[(size_t)i * p.T + s] + armature[i] * p.at(p.qdda_des, i, s);
            u[(size_t)i * p.T + s] = u[(size_t)i * p.T + s] + damping[i] * p.at(p.qd_des, i, s);
        }
    }
}

void torque_radius(Plan& p) {   // KPR/armour_main.cu:173-211
    const int T = p.T;
    p.torque_radius.assign((size_t)NF * T, 0.0);
    for (int t = 0; t < T; t++) {
        Interval rho(0.0);
        for (int i = 0; i < NF; i++) {
            Mat c, r; p.at(p.u_nom_int, i, t).toInterval(c, r);
            Interval temp(c(0) - r(0), c(0) + r(0));
            rho += temp * temp;
            p.torque_radius[i + (size_t)t * NF] = alpha_ub * (M_max - M_min) * eps_ub + 0.5 * max(std::fabs(temp.lower()), std::fabs(temp.upper()));
        }
        rho = sqrt(rho);
        for (int i = 0; i < NF; i++) p.torque_radius[i + (size_t)t * NF] += 0.5 * rho.upper();
        for (int i = 0; i < NF; i++) p.torque_radius[i + (size_t)t * NF] += p.at(p.u_nom, i, t).independent(0);
        for (int i = 0; i < NF; i++) p.torque_radius[i + (size_t)t * NF] += friction[i];
    }
}

// pair table of the 9 buffered generators: KPR/CollisionChecking.cu:26-39
void comb_table(unsigned* ca, unsigned* cb) {
    unsigned a = 0, b = 1;
    for (int i = 0; i < COMB; i++) {
        ca[i] = a; cb[i] = b;
        if (b < (unsigned)BUF_GEN - 1) b++;
        else { a++; b = a + 1; }
    }
}

void initializeHyperPlane(Plan& p) {   // KPR/CollisionChecking.cu:74-88 + kernels :136-228
    const int T = p.T, n_obs = p.n_obs;
    if (n_obs == 0) return;
    p.A.assign((size_t)T * NJ * n_obs * COMB * 3, 0.0);
    p.d.assign((size_t)T * NJ * n_obs * COMB, 0.0);
    p.delta.assign((size_t)T * NJ * n_obs * COMB, 0.0);
    unsigned ca[COMB], cb[COMB];
    comb_table(ca, cb);
    for (int link = 0; link < NJ; link++)
        for (int t = 0; t < T; t++)
            for (int o = 0; o < n_obs; o++) {
                double G[BUF_GEN][3], c[3];
                const double* ob = &p.obstacles[(size_t)o * (OBS_GEN + 1) * 3];
                for (int a = 0; a < 3; a++) c[a] = ob[a];
                for (int g = 0; g < OBS_GEN; g++) for (int a = 0; a < 3; a++) G[g][a] = ob[(g + 1) * 3 + a];
                const Mat& lg = p.link_gens[(size_t)t * NJ + link];
                for (int g = 0; g < 6; g++) for (int a = 0; a < 3; a++) G[g + OBS_GEN][a] = lg(a, g);
                for (int pid = 0; pid < COMB; pid++) {
                    const unsigned ia = ca[pid], ib = cb[pid];
                    double cr[3];
                    cr[0] = G[ia][1] * G[ib][2] - G[ia][2] * G[ib][1];
                    cr[1] = G[ia][2] * G[ib][0] - G[ia][0] * G[ib][2];
                    cr[2] = G[ia][0] * G[ib][1] - G[ia][1] * G[ib][0];
                    const double nrm = sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
                    double C[3] = {0, 0, 0};
                    if (nrm > 0) for (int a = 0; a < 3; a++) C[a] = cr[a] / nrm;
                    const size_t idx = (((size_t)t * NJ + link) * n_obs + o) * COMB + pid;
                    for (int a = 0; a < 3; a++) p.A[idx * 3 + a] = C[a];
                    p.d[idx] = C[0] * c[0] + C[1] * c[1] + C[2] * c[2];
                    double dl = 0.0;
                    for (int j = 0; j < BUF_GEN; j++) dl += std::fabs(C[0] * G[j][0] + C[1] * G[j][1] + C[2] * G[j][2]);
                    p.delta[idx] = dl;
                }
            }
}

// KPR/CollisionChecking.cu:90-134 + checkCollisionKernel :230-299
void linkFRSConstraints(Plan& p, double* link_c, double* grad_link_c) {
    const int T = p.T, n_obs = p.n_obs;
    if (n_obs == 0) return;
    const bool grad = grad_link_c != nullptr;
    for (int link = 0; link < NJ; link++)
        for (int t = 0; t < T; t++)
            for (int o = 0; o < n_obs; o++) {
                const size_t base = (((size_t)t * NJ + link) * n_obs + o) * COMB;
                const Mat& ce = p.link_sliced_center[(size_t)t * NJ + link];
                double max_elt = -100000000;
                unsigned max_id = 0;
                bool neg = false;
                for (int i = 0; i < COMB; i++) {
                    const double* Ae = &p.A[(base + i) * 3];
                    double pos_res, neg_res;
                    if (sqrt(Ae[0] * Ae[0] + Ae[1] * Ae[1] + Ae[2] * Ae[2]) > 0) {
                        const double dot = Ae[0] * ce(0) + Ae[1] * ce(1) + Ae[2] * ce(2);
                        pos_res = dot - (p.d[base + i] + p.delta[base + i]);
                        neg_res = -dot - (-p.d[base + i] + p.delta[base + i]);
                    }
                    else { pos_res = -100000000; neg_res = -100000000; }
                    if (pos_res > max_elt) { max_elt = pos_res; max_id = i; neg = false; }
                    if (neg_res > max_elt) { max_elt = neg_res; max_id = i; neg = true; }
                }
                const size_t out = ((size_t)link * T + t) * n_obs + o;
                if (link_c) link_c[out] = -max_elt;
                if (grad) {
                    const double* Am = &p.A[(base + max_id) * 3];
                    for (int k = 0; k < NF; k++) {
                        const Mat& dk = p.dk_link_sliced_center[((size_t)t * NJ + link) * NF + k];
                        const double dot = Am[0] * dk(0) + Am[1] * dk(1) + Am[2] * dk(2);
                        grad_link_c[out * NF + k] = neg ? dot : -dot;
                    }
                }
            }
}

// joint position / velocity extrema over the whole trajectory: KPR/Trajectory.cu:256-540
struct Extremum { double mn, mx; int mnId, mxId; double e2, e3; };
Extremum joint_extremum(const Plan& p, int i, double k_actual, bool velocity) {
    const double a = p.Tqd0[i], b = p.TTqdd0[i], q0 = p.q0[i];
    double e2, e3;
    if (!velocity) {
        e2 = (2 * a + b + sqrt(64 * pow(a, 2) + 14 * a * b - 120 * k_actual * a + pow(b, 2))) / (5 * (6 * a - 12 * k_actual + b));
        e3 = (2 * a + b - sqrt(64 * pow(a, 2) + 14 * a * b - 120 * k_actual * a + pow(b, 2))) / (5 * (6 * a - 12 * k_actual + b));
    }
    else {
        e2 = (18 * a - 30 * k_actual + 4 * b + sqrt(6 * (150 * pow(k_actual, 2) - 180 * k_actual * a - 20 * k_actual * b + 54 * pow(a, 2) + 14 * a * b + pow(b, 2)))) / (10 * (6 * a - 12 * k_actual + b));
        e3 = (18 * a - 30 * k_actual + 4 * b - sqrt(6 * (150 * pow(k_actual, 2) - 180 * k_actual * a - 20 * k_actual * b + 54 * pow(a, 2) + 14 * a * b + pow(b, 2)))) / (10 * (6 * a - 12 * k_actual + b));
    }
    auto f = velocity ? qd_des_func : q_des_func;
    const double v1 = f(q0, a, b, k_actual, 0), v2 = f(q0, a, b, k_actual, e2), v3 = f(q0, a, b, k_actual, e3), v4 = f(q0, a, b, k_actual, 1);
    Extremum r;
    r.e2 = e2; r.e3 = e3;
    if (v1 < v4) { r.mn = v1; r.mnId = 1; r.mx = v4; r.mxId = 4; }
    else { r.mn = v4; r.mnId = 4; r.mx = v1; r.mxId = 1; }
    if (0 <= e2 && e2 <= 1) {
        if (v2 < r.mn) { r.mn = v2; r.mnId = 2; }
        if (r.mx < v2) { r.mx = v2; r.mxId = 2; }
    }
    if (0 <= e3 && e3 <= 1) {
        if (v3 < r.mn) { r.mn = v3; r.mnId = 3; }
        if (r.mx < v3) { r.mx = v3; r.mxId = 3; }
    }
    return r;
}
// The value variant (:256-288, :399-431) uses min/max chains; with NaN-free inputs it selects the same numbers.
void returnJointExtremum(const Plan& p, double* extremum, const double* k, bool velocity) {
    for (int i = 0; i < NF; i++) {
        const double k_actual = p.k_range[i] * k[i];
        const double a = p.Tqd0[i], b = p.TTqdd0[i], q0 = p.q0[i];
        Extremum e = joint_extremum(p, i, k_actual, velocity);
        auto f = velocity ? qd_des_func : q_des_func;
        const double v1 = f(q0, a, b, k_actual, 0), v2 = f(q0, a, b, k_actual, e.e2), v3 = f(q0, a, b, k_actual, e.e3), v4 = f(q0, a, b, k_actual, 1);
        double mn = min(v1, v4), mx = max(v1, v4);
        if (0 <= e.e2 && e.e2 <= 1) { mn = min(mn, v2); mx = max(mx, v2); }
        if (0 <= e.e3 && e.e3 <= 1) { mn = min(mn, v3); mx = max(mx, v3); }
        extremum[i] = velocity ? mn / DURATION : mn;
        extremum[i + NF] = velocity ? mx / DURATION : mx;
    }
}
void returnJointExtremumGradient(const Plan& p, double* grad, const double* k, bool velocity) {
    for (int i = 0; i < NF; i++) {
        const double k_actual = p.k_range[i] * k[i];
        const double a = p.Tqd0[i], b = p.TTqdd0[i], q0 = p.q0[i];
        Extremum e = joint_extremum(p, i, k_actual, velocity);
        auto dk = [&](int id) -> double {
            switch (id) {
                case 1: return 0.0;
                case 2: return velocity ? qd_des_extrema_k_derivative(q0, a, b, k_actual, +1) : q_des_extrema_k_derivative(q0, a, b, k_actual, +1);
                case 3: return velocity ? qd_des_extrema_k_derivative(q0, a, b, k_actual, -1) : q_des_extrema_k_derivative(q0, a, b, k_actual, -1);
                default: return 1.0;
            }
        };
        const double gmn = dk(e.mnId), gmx = dk(e.mxId);
        for (int j = 0; j < NF; j++) {
            if (i == j) {
                grad[(i) * NF + j] = velocity ? gmn * p.k_range[i] / DURATION : gmn * p.k_range[i];
                grad[(i + NF) * NF + j] = velocity ? gmx * p.k_range[i] / DURATION : gmx * p.k_range[i];
            }
            else { grad[(i) * NF + j] = 0.0; grad[(i + NF) * NF + j] = 0.0; }
        }
    }
}


// ---- ARMTD comparison planner: KPA/Trajectory.cu ------------------------------------------------------------
// joint rotation PZs from the offline JRS zonotopes of cos / sin of the displacement (KPA/Trajectory.cu:29-81)
void makePolyZono_armtd(Plan& p, int s) {
    const int T = p.T;
    auto tab = [&](int which, int i) { return p.jrs[((size_t)which * NF + i) * T + s]; };
    for (int i = 0; i < NF; i++) {
        const double cos_q0 = std::cos(p.q0[i]), sin_q0 = std::sin(p.q0[i]);
        double cos_center = cos_q0 * tab(0, i) - sin_q0 * tab(3, i);
        double cos_coeff[2];
        cos_coeff[0] = cos_q0 * tab(1, i) - sin_q0 * tab(4, i);
        cos_coeff[1] = std::fabs(cos_q0) * tab(2, i) + std::fabs(sin_q0) * tab(5, i);
        cos_coeff[1] *= 5.0;
        uint64_t cos_deg[2][NVAR] = {{0}};
        cos_deg[0][i] = 1; cos_deg[1][i + NF * 4] = 1;
        double sin_center = cos_q0 * tab(3, i) + sin_q0 * tab(0, i);
        double sin_coeff[2];
        sin_coeff[0] = cos_q0 * tab(4, i) + sin_q0 * tab(1, i);
        sin_coeff[1] = std::fabs(cos_q0) * tab(5, i) + std::fabs(sin_q0) * tab(2, i);
        sin_coeff[1] *= 5.0;
        uint64_t sin_deg[2][NVAR] = {{0}};
        sin_deg[0][i] = 1; sin_deg[1][i + NF * 5] = 1;
        p.at(p.cos_q, i, s) = PZ(cos_center, cos_coeff, cos_deg, 2);   // not kept by the reference; exported for parity tests
        p.at(p.sin_q, i, s) = PZ(sin_center, sin_coeff, sin_deg, 2);
        p.at(p.R, i, s) = PZ(rots[i * 3], rots[i * 3 + 1], rots[i * 3 + 2]);
        if (axes[i] != 0)
            p.at(p.R, i, s) = p.at(p.R, i, s) * PZ(cos_center, cos_coeff, cos_deg, 2, sin_center, sin_coeff, sin_deg, 2, axes[i]);
        p.at(p.R_t, i, s) = p.at(p.R, i, s).transpose();
    }
    p.at(p.R, NJ, s) = PZ(0.0, 0.0, 0.0);
}
// min / max joint position and velocity over the move-then-brake trajectory and the derivative of the selected
// branch (KPA/Trajectory.cu:83-411).  Note: the reference's derivatives are with respect to k_actual = k_range * k
// (no k_range factor) and its gradient routine zeroes only 4*NF*NF BYTES of the output; both are kept as they are
// (off-diagonal entries are written as 0 here).
void armtd_state_extremum(const Plan& p, const double* k, double* ext, double* grad) {
    const double t_move = 0.5, t_total = 1.0, t_to_stop = t_total - t_move;
    for (int i = 0; i < NF; i++) {
        const double q0 = p.q0[i], qd0 = p.qd0[i];
        const double k_actual = p.k_range[i] * k[i];
        const double q_peak = q0 + qd0 * t_move + k_actual * t_move * t_move * 0.5;
        const double q_dot_peak = qd0 + k_actual * t_move;
        const double q_ddot_to_stop = -q_dot_peak / t_to_stop;
        const double q_stop = q_peak + q_dot_peak * t_to_stop + 0.5 * q_ddot_to_stop * t_to_stop * t_to_stop;
        const double t_mm = -qd0 / k_actual;
        double qe0, qe1, ge0, ge1;
        if (q_peak >= q0) { qe0 = q0; qe1 = q_peak; ge0 = 0; ge1 = 0.5 * t_move * t_move; }
        else { qe0 = q_peak; qe1 = q0; ge0 = 0.5 * t_move * t_move; ge1 = 0; }
        double q_min_pk, q_max_pk, g_min_pk, g_max_pk;
        if (t_mm > 0 && t_mm < t_move) {
            if (k_actual >= 0) {
                q_min_pk = q0 + qd0 * t_mm + 0.5 * k_actual * t_mm * t_mm; q_max_pk = qe1;
                g_min_pk = (0.5 * qd0 * qd0) / (k_actual * k_actual); g_max_pk = ge1;
            }
            else {
                q_min_pk = qe0; q_max_pk = q0 + qd0 * t_mm + 0.5 * k_actual * t_mm * t_mm;
                g_min_pk = ge0; g_max_pk = (0.5 * qd0 * qd0) / (k_actual * k_actual);
            }
        }
        else { q_min_pk = qe0; q_max_pk = qe1; g_min_pk = ge0; g_max_pk = ge1; }
        double v_min_pk, v_max_pk, gv_min_pk, gv_max_pk;
        if (q_dot_peak >= qd0) { v_min_pk = qd0; v_max_pk = q_dot_peak; gv_min_pk = 0; gv_max_pk = t_move; }
        else { v_min_pk = q_dot_peak; v_max_pk = qd0; gv_min_pk = t_move; gv_max_pk = 0; }
        double q_min_st, q_max_st, g_min_st, g_max_st;
        if (q_stop >= q_peak) { q_min_st = q_peak; q_max_st = q_stop; g_min_st = 0.5 * t_move * t_move; g_max_st = 0.5 * t_move * t_move + 0.5 * t_move * t_to_stop; }
        else { q_min_st = q_stop; q_max_st = q_peak; g_min_st = 0.5 * t_move * t_move + 0.5 * t_move * t_to_stop; g_max_st = 0.5 * t_move * t_move; }
        double v_min_st, v_max_st, gv_min_st, gv_max_st;
        if (q_dot_peak >= 0) { v_min_st = 0; v_max_st = q_dot_peak; gv_min_st = 0; gv_max_st = t_move; }
        else { v_min_st = q_dot_peak; v_max_st = 0; gv_min_st = t_move; gv_max_st = 0; }
        const bool a = q_min_pk <= q_min_st, b = q_max_pk >= q_max_st, c = v_min_pk <= v_min_st, d = v_max_pk >= v_max_st;
        if (ext) {
            ext[i] = a ? q_min_pk : q_min_st;
            ext[i + NF] = b ? q_max_pk : q_max_st;
            ext[i + 2 * NF] = c ? v_min_pk : v_min_st;
            ext[i + 3 * NF] = d ? v_max_pk : v_max_st;
        }
        if (grad) {
            for (int r = 0; r < 4; r++) for (int j = 0; j < NF; j++) grad[(i + r * NF) * NF + j] = 0.0;
            grad[i * NF + i] = a ? g_min_pk : g_min_st;
            grad[(i + NF) * NF + i] = b ? g_max_pk : g_max_st;
            grad[(i + 2 * NF) * NF + i] = c ? gv_min_pk : gv_min_st;
            grad[(i + 3 * NF) * NF + i] = d ? gv_max_pk : gv_max_st;
        }
    }
}

double wrap_to_pi(double angle) {   // KPR/NLPclass.cu:6-15
    double w = angle;
    while (w < -M_PI) w += 2 * M_PI;
    while (w > M_PI) w -= 2 * M_PI;
    return w;
}

int constraint_number(const Plan& p) {   // KPR/NLPclass.cu:47-49; KPA/NLPclass.cu:42-43
    return (p.mode == 0 ? NF * p.T : 0) + NJ * p.T * p.n_obs + NF * 4;
}

}  // namespace

// =================================================================================================
// C ABI (mirrors include/armour_b200.h so tests can drive either backend)
// =================================================================================================
extern "C" {

struct oracle_config {
    int num_time_steps;
    double k_range[7];
    double mass_uncertainty;
    double inertia_uncertainty;
    double simplify_threshold;
    int num_threads;   // 0 = omp default
};

void* oracle_create(const oracle_config* cfg) {
    Plan* p = new Plan();
    p->T = cfg->num_time_steps;
    for (int i = 0; i < NF; i++) p->k_range[i] = cfg->k_range[i];
    p->mass_uncertainty = cfg->mass_uncertainty;
    p->inertia_uncertainty = cfg->inertia_uncertainty;
    p->threshold = cfg->simplify_threshold;
    p->num_threads = cfg->num_threads;
    return p;
}
void oracle_destroy(void* h) { delete (Plan*)h; }

// stages A-D: the reference's `duration1` span (KPR/armour_main.cu:89-226)
int oracle_build(void* h, const double* q0, const double* qd0, const double* qdd0, const double* obstacles, int n_obs) {
    Plan& p = *(Plan*)h;
    if (n_obs < 0) return -1;
    PZ::threshold() = p.threshold;
    p.mode = 0;
    p.n_obs = n_obs;
    p.obstacles.assign(obstacles, obstacles + (size_t)n_obs * 12);
    const int T = p.T;
    auto t0 = std::chrono::high_resolution_clock::now();
    if (p.num_threads > 0) omp_set_num_threads(p.num_threads);
    int nthreads = omp_get_max_threads();
    std::vector<OpStats> tstats(nthreads);
    bezier_init(p, q0, qd0, qdd0);
    int err = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int s = 0; s < T; s++) {
        try { makePolyZono(p, s); } catch (int) { err = -1; }
    }
    if (err) return err;
    kd_init(p);
    p.link_gens.assign((size_t)T * NJ, Mat());
#pragma omp parallel for schedule(dynamic)
    for (int s = 0; s < T; s++) {
        tls_stats() = &tstats[omp_get_thread_num()];
        try {
            fk(p, s);
            for (int i = 0; i < NJ; i++) p.link_gens[(size_t)s * NJ + i] = p.at(p.links, i, s).reduce_link_PZ();
            rnea(p, s, p.mass_nominal, p.I_nominal, p.u_nom);
            rnea(p, s, p.mass_uncertain, p.I_uncertain, p.u_nom_int);
            for (int i = 0; i < NF; i++) p.at(p.u_nom_int, i, s) = p.at(p.u_nom_int, i, s) - p.at(p.u_nom, i, s);
            for (int i = 0; i < NF; i++) p.at(p.u_nom, i, s).reduce();
        } catch (int) { err = -1; }
        tls_stats() = nullptr;
    }
    if (err) return err;
    torque_radius(p);
    initializeHyperPlane(p);
    auto t1 = std::chrono::high_resolution_clock::now();
    p.build_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    p.stats = OpStats();
    for (auto& s : tstats) p.stats.add(s);
    p.link_sliced_center.assign((size_t)T * NJ, Mat(3, 1));
    p.dk_link_sliced_center.assign((size_t)T * NJ * NF, Mat(3, 1));
    return 0;
}

// ARMTD comparison planner build (KPA/armtd_main.cu:110-160): JRS tables -> rotation PZs -> FK -> half-space tables.
// jrs: [6][7][T] doubles in the order c_cos, g_cos, r_cos, c_sin, g_sin, r_sin; k_range is a per-problem input there.
int oracle_build_armtd(void* h, const double* q0, const double* qd0, const double* jrs, const double* k_range, const double* obstacles, int n_obs) {
    Plan& p = *(Plan*)h;
    if (n_obs < 0) return -1;
    PZ::threshold() = p.threshold;
    p.mode = 1;
    p.n_obs = n_obs;
    p.obstacles.assign(obstacles, obstacles + (size_t)n_obs * 12);
    const int T = p.T;
    p.jrs.assign(jrs, jrs + (size_t)6 * NF * T);
    for (int i = 0; i < NF; i++) { p.q0[i] = q0[i]; p.qd0[i] = qd0[i]; p.qdd0[i] = 0; p.k_range[i] = k_range[i]; }
    auto t0 = std::chrono::high_resolution_clock::now();
    if (p.num_threads > 0) omp_set_num_threads(p.num_threads);
    p.cos_q.assign((size_t)NF * T, PZ()); p.sin_q.assign((size_t)NF * T, PZ());
    p.R.assign((size_t)(NJ + 1) * T, PZ()); p.R_t.assign((size_t)NJ * T, PZ());
    p.qd_des.assign((size_t)NF * T, PZ(1, 1)); p.qda_des.assign((size_t)NF * T, PZ(1, 1)); p.qdda_des.assign((size_t)NF * T, PZ(1, 1));
    int err = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int s = 0; s < T; s++) {
        try { makePolyZono_armtd(p, s); } catch (int) { err = -1; }
    }
    if (err) return err;
    kd_init(p);
    p.link_gens.assign((size_t)T * NJ, Mat());
#pragma omp parallel for schedule(dynamic)
    for (int s = 0; s < T; s++) {
        try {
            fk(p, s);
            for (int i = 0; i < NJ; i++) p.link_gens[(size_t)s * NJ + i] = p.at(p.links, i, s).reduce_link_PZ();
        } catch (int) { err = -1; }
    }
    if (err) return err;
    for (auto& z : p.u_nom) z = PZ(1, 1);
    for (auto& z : p.u_nom_int) z = PZ(1, 1);
    p.torque_radius.assign((size_t)NF * T, 0.0);
    initializeHyperPlane(p);
    auto t1 = std::chrono::high_resolution_clock::now();
    p.build_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    p.link_sliced_center.assign((size_t)T * NJ, Mat(3, 1));
    p.dk_link_sliced_center.assign((size_t)T * NJ * NF, Mat(3, 1));
    return 0;
}
double oracle_build_ms(void* h) { return ((Plan*)h)->build_ms; }
int oracle_num_threads() { return omp_get_max_threads(); }

// out[8]: n_mul, pair_products, flops, n_simplify, simplify_in, simplify_out, max_simplify_in, max_simplify_out
void oracle_op_stats(void* h, uint64_t* out) {
    const OpStats& s = ((Plan*)h)->stats;
    out[0] = s.n_mul; out[1] = s.pair_products; out[2] = s.flops; out[3] = s.n_simplify;
    out[4] = s.simplify_in; out[5] = s.simplify_out; out[6] = s.max_simplify_in; out[7] = s.max_simplify_out;
}

// ---- TNLP callbacks (KPR/NLPclass.cu) -------------------------------------------------------------
int oracle_get_nlp_info(void* h, int* n, int* m, int* nnz_jac_g, int* nnz_h_lag) {   // :62-82
    Plan& p = *(Plan*)h;
    *n = NF; *m = constraint_number(p); *nnz_jac_g = *m * *n; *nnz_h_lag = 0;
    return 0;
}
int oracle_get_bounds_info(void* h, double* x_l, double* x_u, double* g_l, double* g_u) {   // :87-165
    Plan& p = *(Plan*)h;
    const int T = p.T;
    for (int i = 0; i < NF; i++) { x_l[i] = -1.0; x_u[i] = 1.0; }
    int offset = 0;
    if (p.mode == 0) {
        for (int i = 0; i < T; i++)
            for (int j = 0; j < NF; j++) {
                g_l[i * NF + j] = -torque_limits[j] + p.torque_radius[j + (size_t)i * NF];
                g_u[i * NF + j] = torque_limits[j] - p.torque_radius[j + (size_t)i * NF];
            }
        offset += NF * T;
    }
    for (int i = offset; i < offset + T * NJ * p.n_obs; i++) { g_l[i] = -1e19; g_u[i] = 0; }
    offset += T * NJ * p.n_obs;
    for (int rep = 0; rep < 2; rep++) {
        for (int i = offset; i < offset + NF; i++) { g_l[i] = state_limits_lb[i - offset] + qe; g_u[i] = state_limits_ub[i - offset] - qe; }
        offset += NF;
    }
    for (int rep = 0; rep < 2; rep++) {
        for (int i = offset; i < offset + NF; i++) { g_l[i] = -speed_limits[i - offset] + qde; g_u[i] = speed_limits[i - offset] - qde; }
        offset += NF;
    }
    return 0;
}
int oracle_get_starting_point(void*, double* x) { for (int i = 0; i < NF; i++) x[i] = 0.0; return 0; }   // :170-202

int oracle_eval_f(void* h, const double* q_des, double t_plan, const double* x, double* obj) {   // :207-236
    Plan& p = *(Plan*)h;
    double qp[NF];
    for (int i = 0; i < NF; i++)
        qp[i] = p.mode == 0 ? q_des_func(p.q0[i], p.Tqd0[i], p.TTqdd0[i], p.k_range[i] * x[i], t_plan)
                            : p.q0[i] + p.qd0[i] * 0.5 + p.k_range[i] * x[i] * 0.125;   // KPA/NLPclass.cu:197
    double v = pow(wrap_to_pi(q_des[0] - qp[0]), 2) + pow(wrap_to_pi(q_des[2] - qp[2]), 2) + pow(wrap_to_pi(q_des[4] - qp[4]), 2) +
               pow(wrap_to_pi(q_des[6] - qp[6]), 2) + pow(q_des[1] - qp[1], 2) + pow(q_des[3] - qp[3], 2) + pow(q_des[5] - qp[5], 2);
    *obj = v * COST_SCALE;
    return 0;
}
int oracle_eval_grad_f(void* h, const double* q_des, double t_plan, const double* x, double* grad) {   // :241-267
    Plan& p = *(Plan*)h;
    for (int i = 0; i < NF; i++) {
        double qp = p.mode == 0 ? q_des_func(p.q0[i], p.Tqd0[i], p.TTqdd0[i], p.k_range[i] * x[i], t_plan)
                                : p.q0[i] + p.qd0[i] * 0.5 + p.k_range[i] * x[i] * 0.125;
        double dk = p.mode == 0 ? pow(t_plan, 3) * (6 * pow(t_plan, 2) - 15 * t_plan + 10) * p.k_range[i] : p.k_range[i] * 0.125;   // KPA/NLPclass.cu:229-230
        grad[i] = (i % 2 == 0) ? (2 * wrap_to_pi(qp - q_des[i]) * dk) : (2 * (qp - q_des[i]) * dk);
        grad[i] *= COST_SCALE;
    }
    return 0;
}
int oracle_eval_g(void* h, const double* x, double* g) {   // :272-324
    Plan& p = *(Plan*)h;
    const int T = p.T;
    if (p.mode == 1) {   // KPA/NLPclass.cu:263-277
#pragma omp parallel for schedule(dynamic)
        for (int i = 0; i < T; i++)
            for (int l = 0; l < NJ; l++) p.link_sliced_center[(size_t)i * NJ + l] = p.at(p.links, l, i).sliceCenter(x);
        linkFRSConstraints(p, g, nullptr);
        armtd_state_extremum(p, x, g + NJ * T * p.n_obs, nullptr);
        return 0;
    }
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < T; i++) {
        for (int k = 0; k < NF; k++) g[i * NF + k] = p.at(p.u_nom, k, i).sliceCenter(x)(0);
        for (int l = 0; l < NJ; l++) p.link_sliced_center[(size_t)i * NJ + l] = p.at(p.links, l, i).sliceCenter(x);
    }
    linkFRSConstraints(p, g + T * NF, nullptr);
    returnJointExtremum(p, g + T * NF + T * NJ * p.n_obs, x, false);
    returnJointExtremum(p, g + T * NF + T * NJ * p.n_obs + NF * 2, x, true);
    return 0;
}
int oracle_eval_jac_g(void* h, const double* x, double* values) {   // :330-396 (values != NULL branch)
    Plan& p = *(Plan*)h;
    const int T = p.T;
    if (p.mode == 1) {   // KPA/NLPclass.cu eval_jac_g
#pragma omp parallel for schedule(dynamic)
        for (int i = 0; i < T; i++)
            for (int l = 0; l < NJ; l++) {
                p.link_sliced_center[(size_t)i * NJ + l] = p.at(p.links, l, i).sliceCenter(x);
                p.at(p.links, l, i).sliceGradient(&p.dk_link_sliced_center[((size_t)i * NJ + l) * NF], x);
            }
        linkFRSConstraints(p, nullptr, values);
        armtd_state_extremum(p, x, nullptr, values + (size_t)NJ * T * p.n_obs * NF);
        return 0;
    }
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < T; i++) {
        for (int k = 0; k < NF; k++) {
            Mat gr[NF];
            p.at(p.u_nom, k, i).sliceGradient(gr, x);
            for (int j = 0; j < NF; j++) values[(i * NF + k) * NF + j] = gr[j](0);
        }
        for (int l = 0; l < NJ; l++) {
            p.link_sliced_center[(size_t)i * NJ + l] = p.at(p.links, l, i).sliceCenter(x);
            p.at(p.links, l, i).sliceGradient(&p.dk_link_sliced_center[((size_t)i * NJ + l) * NF], x);
        }
    }
    linkFRSConstraints(p, nullptr, values + (size_t)T * NF * NF);
    returnJointExtremumGradient(p, values + ((size_t)T * NF + (size_t)T * NJ * p.n_obs) * NF, x, false);
    returnJointExtremumGradient(p, values + ((size_t)T * NF + (size_t)T * NJ * p.n_obs + NF * 2) * NF, x, true);
    return 0;
}
int oracle_jac_structure(void* h, int* iRow, int* jCol) {   // :348-357
    Plan& p = *(Plan*)h;
    const int m = constraint_number(p);
    for (int i = 0; i < m; i++) for (int j = 0; j < NF; j++) { iRow[i * NF + j] = i; jCol[i * NF + j] = j; }
    return 0;
}
// finalize_solution's feasibility re-check (:446-537); returns 1 feasible / 0 infeasible
int oracle_check_feasible(void* h, const double* g) {
    Plan& p = *(Plan*)h;
    const int T = p.T;
    int offset = 0;
    if (p.mode == 0) {
        for (int i = 0; i < T; i++)
            for (int j = 0; j < NF; j++) {
                const double r = p.torque_radius[j + (size_t)i * NF];
                if (g[i * NF + j] < -torque_limits[j] + r - TORQUE_THRESHOLD || g[i * NF + j] > torque_limits[j] - r + TORQUE_THRESHOLD) return 0;
            }
        offset += NF * T;
    }
    const int links_checked = p.mode == 0 ? NJ : NF - 1;   // the ARMTD planner re-checks links 0..NUM_FACTORS-2 only (KPA/NLPclass.cu:388)
    for (int i = 0; i < links_checked; i++)
        for (int j = 0; j < T; j++)
            for (int o = 0; o < p.n_obs; o++)
                if (g[(i * T + j) * p.n_obs + o + offset] > COLLISION_THRESHOLD) return 0;
    offset += NJ * T * p.n_obs;
    for (int rep = 0; rep < 2; rep++) {
        for (int i = offset; i < offset + NF; i++)
            if (g[i] < state_limits_lb[i - offset] + qe || g[i] > state_limits_ub[i - offset] - qe) return 0;
        offset += NF;
    }
    for (int rep = 0; rep < 2; rep++) {
        for (int i = offset; i < offset + NF; i++)
            if (g[i] < -speed_limits[i - offset] + qde || g[i] > speed_limits[i - offset] - qde) return 0;
        offset += NF;
    }
    return 1;
}

// ---- table getters for parity tests -----------------------------------------------------------------
// which: 0 cos_q, 1 sin_q, 2 R, 3 R_t, 4 qd_des, 5 qda_des, 6 qdda_des, 7 links (after reduce_link_PZ),
//        8 u_nom (after reduce), 9 u_nom_int (disturbance, after subtraction)
static std::vector<PZ>* table(Plan& p, int which) {
    switch (which) {
        case 0: return &p.cos_q; case 1: return &p.sin_q; case 2: return &p.R; case 3: return &p.R_t;
        case 4: return &p.qd_des; case 5: return &p.qda_des; case 6: return &p.qdda_des;
        case 7: return &p.links; case 8: return &p.u_nom; case 9: return &p.u_nom_int;
    }
    return nullptr;
}
// returns number of monomials; dims[0..1] = rows, cols.  Pass null outputs to query sizes only.
int oracle_get_pz(void* h, int which, int idx, int s, int* dims, uint64_t* keys, double* coeffs, double* center, double* independent) {
    Plan& p = *(Plan*)h;
    std::vector<PZ>* t = table(p, which);
    if (!t) return -1;
    const PZ& z = (*t)[(size_t)idx * p.T + s];
    const int dim = z.NRows * z.NCols;
    if (dims) { dims[0] = z.NRows; dims[1] = z.NCols; }
    const int n = (int)z.polynomial.size();
    if (keys) for (int i = 0; i < n; i++) keys[i] = z.polynomial[i].degree;
    if (coeffs) for (int i = 0; i < n; i++) for (int a = 0; a < dim; a++) coeffs[(size_t)i * dim + a] = z.polynomial[i].coeff.v[a];
    if (center) for (int a = 0; a < dim; a++) center[a] = z.center.v[a];
    if (independent) for (int a = 0; a < dim; a++) independent[a] = z.independent.v[a];
    return n;
}
void oracle_get_torque_radius(void* h, double* out) { Plan& p = *(Plan*)h; std::copy(p.torque_radius.begin(), p.torque_radius.end(), out); }
// out[(s*NJ + i)*18 + col*3 + row]  (column-major 3x6 like Eigen)
void oracle_get_link_generators(void* h, double* out) {
    Plan& p = *(Plan*)h;
    for (size_t e = 0; e < p.link_gens.size(); e++) std::copy(p.link_gens[e].v.begin(), p.link_gens[e].v.end(), out + e * 18);
}
void oracle_get_taylor_remainders(void* h, double* cos_rem, double* sin_rem) {
    Plan& p = *(Plan*)h;
    std::copy(p.cos_rem.begin(), p.cos_rem.end(), cos_rem);
    std::copy(p.sin_rem.begin(), p.sin_rem.end(), sin_rem);
}
void oracle_get_hyperplanes(void* h, double* A, double* d, double* delta) {
    Plan& p = *(Plan*)h;
    std::copy(p.A.begin(), p.A.end(), A); std::copy(p.d.begin(), p.d.end(), d); std::copy(p.delta.begin(), p.delta.end(), delta);
}
void oracle_get_link_sliced_center(void* h, double* out) {
    Plan& p = *(Plan*)h;
    for (size_t e = 0; e < p.link_sliced_center.size(); e++) for (int a = 0; a < 3; a++) out[e * 3 + a] = p.link_sliced_center[e](a);
}

// ---- stand-alone PZ primitives on flat arrays (for primitive-level parity tests) -----------------------
static PZ pz_from_flat(int rows, int cols, int n, const uint64_t* keys, const double* coeffs, const double* center, const double* indep) {
    PZ z(rows, cols);
    const int dim = rows * cols;
    for (int a = 0; a < dim; a++) { z.center.v[a] = center[a]; z.independent.v[a] = indep[a]; }
    for (int i = 0; i < n; i++) {
        Mat c(rows, cols);
        for (int a = 0; a < dim; a++) c.v[a] = coeffs[(size_t)i * dim + a];
        z.polynomial.emplace_back(c, keys[i]);
    }
    return z;
}
static int pz_to_flat(const PZ& z, int cap, int* dims, uint64_t* keys, double* coeffs, double* center, double* indep) {
    const int dim = z.NRows * z.NCols, n = (int)z.polynomial.size();
    dims[0] = z.NRows; dims[1] = z.NCols;
    if (n > cap) return -n;
    for (int i = 0; i < n; i++) { keys[i] = z.polynomial[i].degree; for (int a = 0; a < dim; a++) coeffs[(size_t)i * dim + a] = z.polynomial[i].coeff.v[a]; }
    for (int a = 0; a < dim; a++) { center[a] = z.center.v[a]; indep[a] = z.independent.v[a]; }
    return n;
}
// op: 0 mul, 1 add, 2 sub, 3 cross(a,b) (3x1 each), 4 simplify(a), 5 reduce(a), 6 transpose(a)
int oracle_pz_binary(int op, double threshold,
                     int ar, int ac, int an, const uint64_t* akeys, const double* acoef, const double* acen, const double* aind,
                     int br, int bc, int bn, const uint64_t* bkeys, const double* bcoef, const double* bcen, const double* bind,
                     int cap, int* dims, uint64_t* keys, double* coeffs, double* center, double* indep) {
    PZ::threshold() = threshold;
    PZ a = pz_from_flat(ar, ac, an, akeys, acoef, acen, aind);
    PZ r;
    if (op == 4) { r = a; r.simplify(); }
    else if (op == 5) { r = a; r.reduce(); }
    else if (op == 6) { r = a.transpose(); }
    else {
        PZ b = pz_from_flat(br, bc, bn, bkeys, bcoef, bcen, bind);
        if (op == 0) r = a * b;
        else if (op == 1) r = a + b;
        else if (op == 2) r = a - b;
        else if (op == 3) r = cross(a, b);
        else if (op >= 7 && op <= 9) { r = a; r.addOneDimPZ(b, op - 7, 0); }          // KPR/PZsparse.cu:1068-1085
        else if (op == 10) r = cross(b.center, a);                                       // :1118-1132, the constant is b's centre
        else if (op == 11) r = cross(a, b.center);                                       // :1153-1167
        else return -1;
    }
    return pz_to_flat(r, cap, dims, keys, coeffs, center, indep);
}

// per-op trace of one interval (FK + nominal RNEA), for sizing the device engine. out rows: kind,dim,na,nb,nin,nout,nza,nzb
int oracle_trace_interval(void* h, int s, int cap_rows, int* out) {
    Plan& p = *(Plan*)h;
    PZ::threshold() = p.threshold;
    std::vector<OpTrace> tr;
    tls_trace() = &tr;
    for (int i = 0; i < NJ; i++) p.at(p.links, i, s) = p.link0[i];
    fk(p, s);
    tr.push_back({9, 0, 0, 0, 0, 0});   // marker: end of FK
    std::vector<PZ> u((size_t)NF * p.T);
    rnea(p, s, p.mass_nominal, p.I_nominal, u);
    tls_trace() = nullptr;
    for (int i = 0; i < NJ; i++) p.at(p.links, i, s).reduce_link_PZ();
    const int n = (int)std::min<size_t>(tr.size(), cap_rows);
    for (int i = 0; i < n; i++) { int* o = out + i * 8; o[0] = tr[i].kind; o[1] = tr[i].dim; o[2] = tr[i].na; o[3] = tr[i].nb; o[4] = tr[i].nin; o[5] = tr[i].nout; o[6] = tr[i].nza; o[7] = tr[i].nzb; }
    return (int)tr.size();
}

}  // extern "C"
