// Robust low-level controller on the device: C ABI of include/armour_controller_b200.h.
//
// Reference (KRC = kinova_src/kinova_simulator_interfaces/kinova_robust_controllers_mex):
//   KRC/robot_models.cpp:20-151, 168-249   model file -> body-CoM model -> interval model
//   KRC/spatial.cpp, spatial_interval.cpp   spatial algebra over double / boost interval
//   KRC/rnea.cpp:6-93, 95-185               passivity RNEA (nominal, interval)
//   KRC/robust_controller.cpp:62-171        RobustController::update
//
// Layout of the computation.  Everything that does not depend on the state — the conversion of the model file to
// body-CoM frames, the +-eps interval model, the body-to-world transforms Xbw and the body-frame screw axes Sb
// (KRC/rnea.cpp:38-48 recomputes them every call) — is evaluated once by model_setup_kernel, in the same operation
// order and the same arithmetic (round-to-nearest doubles, directed-rounding intervals), and travels to the update
// kernel as a __grid_constant__ argument, so the per-sample code reads it from the constant bank.
// controller_update_kernel then runs one sample per thread: nominal RNEA, interval RNEA and (ARMOUR method) the
// M(q) r interval pass sharing the joint transforms, the error bound and the robust input.  The per-joint transforms
// and link wrenches a backward pass needs live in thread-local memory (L1-resident); the joint loops stay rolled.
// The work is fp64-pipe bound (about 5e4 directed-rounding multiplies / adds / min-max per sample against 448 bytes
// of HBM traffic), so the roofline it is reported against is the fp64 one.
//
// Restriction: serial chains only (parent[i] == i - 1), which covers every robot file the reference ships.
#include <cuda_runtime.h>

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "../../include/armour_controller_b200.h"
#include "interval.cuh"

namespace armour {
int capi_fail(int code, const std::string& msg);   // armour_capi.cu: sets armour_last_error()

namespace ctrl {
constexpr int MAXJ = ARMOUR_CONTROLLER_MAX_JOINTS;
#ifndef CTRL_MINB
#define CTRL_MINB 4
#endif

// ---- scalar adaptors: S is double or Itv ---------------------------------------------------------------------
template <class S> __device__ __forceinline__ S cst(double v);
template <> __device__ __forceinline__ double cst<double>(double v) { return v; }
template <> __device__ __forceinline__ Itv cst<Itv>(double v) { return itv(v, v); }

template <class S> struct V3 { S x[3]; };
template <class S> struct M3 { S a[9]; };   // row-major
template <class S> struct Twist { V3<S> w, v; };
template <class S> struct Wrench { V3<S> tau, f; };
template <class S> struct Xf { M3<S> R; V3<S> p; };
template <class S> struct Inertia { S m; M3<S> I_bar, mch; };

#define DEV __device__ __forceinline__
template <class S> DEV V3<S> vzero() { V3<S> r; for (int i = 0; i < 3; i++) r.x[i] = cst<S>(0.0); return r; }
template <class S> DEV M3<S> mzero() { M3<S> r; for (int i = 0; i < 9; i++) r.a[i] = cst<S>(0.0); return r; }
template <class S> DEV M3<S> mident() { M3<S> r = mzero<S>(); r.a[0] = r.a[4] = r.a[8] = cst<S>(1.0); return r; }
template <class S> DEV V3<S> operator+(const V3<S>& a, const V3<S>& b) { V3<S> r; for (int i = 0; i < 3; i++) r.x[i] = a.x[i] + b.x[i]; return r; }
template <class S> DEV V3<S> operator-(const V3<S>& a, const V3<S>& b) { V3<S> r; for (int i = 0; i < 3; i++) r.x[i] = a.x[i] - b.x[i]; return r; }
template <class S> DEV V3<S> operator-(const V3<S>& a) { V3<S> r; for (int i = 0; i < 3; i++) r.x[i] = -a.x[i]; return r; }
template <class S> DEV M3<S> operator+(const M3<S>& a, const M3<S>& b) { M3<S> r; for (int i = 0; i < 9; i++) r.a[i] = a.a[i] + b.a[i]; return r; }
template <class S> DEV M3<S> operator-(const M3<S>& a, const M3<S>& b) { M3<S> r; for (int i = 0; i < 9; i++) r.a[i] = a.a[i] - b.a[i]; return r; }
template <class S> DEV M3<S> operator-(const M3<S>& a) { M3<S> r; for (int i = 0; i < 9; i++) r.a[i] = -a.a[i]; return r; }
// scalar factors: a double promoted to the scalar type (Eigen's promotion of the MEX's doubles), or a model scalar
template <class S> DEV V3<S> scaled(const V3<S>& a, double s) { V3<S> r; for (int i = 0; i < 3; i++) r.x[i] = a.x[i] * s; return r; }
template <class S> DEV M3<S> scaled(const M3<S>& a, double s) { M3<S> r; for (int i = 0; i < 9; i++) r.a[i] = a.a[i] * s; return r; }
template <class S> DEV V3<S> scaled_by(const S& s, const V3<S>& a) { V3<S> r; for (int i = 0; i < 3; i++) r.x[i] = s * a.x[i]; return r; }
template <class S> DEV M3<S> scaled_by(const S& s, const M3<S>& a) { M3<S> r; for (int i = 0; i < 9; i++) r.a[i] = s * a.a[i]; return r; }
template <class S> DEV M3<S> tr(const M3<S>& a) { M3<S> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.a[i * 3 + j] = a.a[j * 3 + i]; return r; }
// The products below keep their loops rolled (#pragma unroll 1): fully inlined, one interval RNEA is ~27 k instructions
// (430 KB), far beyond the 32 KB instruction cache, and ncu showed 65 % of the stall samples waiting for instruction
// fetch.  Rolled, the operands are indexed dynamically and live in thread-local memory (L1), the code is ~10x smaller.
template <class S> DEV V3<S> operator*(const M3<S>& m, const V3<S>& v) {
    V3<S> r;
#pragma unroll 1
    for (int i = 0; i < 3; i++) r.x[i] = (m.a[i * 3] * v.x[0] + m.a[i * 3 + 1] * v.x[1]) + m.a[i * 3 + 2] * v.x[2];
    return r;
}
template <class S> DEV M3<S> operator*(const M3<S>& a, const M3<S>& b) {
    M3<S> r;
#pragma unroll 1
    for (int e = 0; e < 9; e++) {
        const int i = e / 3, j = e - 3 * i;
        r.a[e] = (a.a[i * 3] * b.a[j] + a.a[i * 3 + 1] * b.a[3 + j]) + a.a[i * 3 + 2] * b.a[6 + j];
    }
    return r;
}
template <class S> DEV V3<S> cross(const V3<S>& a, const V3<S>& b) {
    V3<S> r;
#pragma unroll 1
    for (int i = 0; i < 3; i++) {
        const int j = i == 2 ? 0 : i + 1, k = i == 0 ? 2 : i - 1;
        r.x[i] = a.x[j] * b.x[k] - a.x[k] * b.x[j];
    }
    return r;
}
template <class S> DEV S dot(const V3<S>& a, const V3<S>& b) { return (a.x[0] * b.x[0] + a.x[1] * b.x[1]) + a.x[2] * b.x[2]; }
template <class S> DEV M3<S> hat(const V3<S>& w) {
    M3<S> m;
    m.a[0] = cst<S>(0.0); m.a[1] = -w.x[2]; m.a[2] = w.x[1];
    m.a[3] = w.x[2]; m.a[4] = cst<S>(0.0); m.a[5] = -w.x[0];
    m.a[6] = -w.x[1]; m.a[7] = w.x[0]; m.a[8] = cst<S>(0.0);
    return m;
}
template <class S> DEV Twist<S> operator+(const Twist<S>& a, const Twist<S>& b) { Twist<S> r; r.w = a.w + b.w; r.v = a.v + b.v; return r; }
template <class S> DEV Twist<S> scaled(const Twist<S>& a, double s) { Twist<S> r; r.w = scaled(a.w, s); r.v = scaled(a.v, s); return r; }
template <class S> DEV Wrench<S> operator+(const Wrench<S>& a, const Wrench<S>& b) { Wrench<S> r; r.tau = a.tau + b.tau; r.f = a.f + b.f; return r; }
template <class S> DEV S dot(const Twist<S>& t, const Wrench<S>& f) { return dot(t.w, f.tau) + dot(t.v, f.f); }   // KRC/spatial.cpp:78-81
// KRC/spatial.cpp:83-87.  w_hat * x is formed as the cross product w x x: the hat matrix only contributes exact zeros
// (0 * x, and "+ 0"), (-w_k) * x_j == -(w_k * x_j) and a + (-b) == a - b hold exactly in both arithmetics.
template <class S> DEV Twist<S> cross(const Twist<S>& a, const Twist<S>& b) {
    Twist<S> r; r.w = cross(a.w, b.w); r.v = cross(a.w, b.v) + cross(a.v, b.w); return r;
}
// KRC/spatial.cpp:143-147.  first_moment_zero: m c^ is exactly zero (the model is expressed in body-CoM frames, so it
// cancels exactly for the robot files checked); its two products are then exact zeros and x + 0 == x, x - 0 == x.
template <class S> DEV Wrench<S> apply(const Inertia<S>& I, const Twist<S>& z, bool first_moment_zero = false) {
    Wrench<S> w;
    if (first_moment_zero) { w.tau = I.I_bar * z.w; w.f = scaled_by(I.m, z.v); return w; }
    w.tau = I.I_bar * z.w + I.mch * z.v;
    w.f = scaled_by(I.m, z.v) - I.mch * z.w;
    return w;
}
// y = M^T x without materialising the transpose (same terms, same association as tr(M) * x)
template <class S> DEV V3<S> mul_t(const M3<S>& m, const V3<S>& v) {
    V3<S> r;
#pragma unroll 1
    for (int i = 0; i < 3; i++) r.x[i] = (m.a[i] * v.x[0] + m.a[3 + i] * v.x[1]) + m.a[6 + i] * v.x[2];
    return r;
}
// Rodrigues rotation about a body-frame screw axis (KRC/spatial.cpp:156-172, spatial_interval.cpp:145-156):
//   R = (I + w^ sin) + ((1 - cos) w^) w^ ;  p = -R' (((I - R) w^) v)
// evaluated entry by entry in that association; the intermediate matrices ((1 - cos) w^, I - R, (I - R) w^, -R') are
// formed a row at a time in registers instead of as thread-local 3x3 temporaries.
template <class S> DEV Xf<S> joint_xf(const Twist<S>& z, double sin_theta, double one_minus_cos) {
    const M3<S> wh = hat(z.w);
    Xf<S> x;
#pragma unroll 1
    for (int e = 0; e < 9; e++) {
        const int i = e / 3, j = e - 3 * i;
        const S t = ((wh.a[i * 3] * one_minus_cos) * wh.a[j] + (wh.a[i * 3 + 1] * one_minus_cos) * wh.a[3 + j]) + (wh.a[i * 3 + 2] * one_minus_cos) * wh.a[6 + j];
        x.R.a[e] = (cst<S>(i == j ? 1.0 : 0.0) + wh.a[e] * sin_theta) + t;
    }
    V3<S> p0;
#pragma unroll 1
    for (int i = 0; i < 3; i++) {
        S m1[3], m2[3];
#pragma unroll
        for (int k = 0; k < 3; k++) m1[k] = cst<S>(i == k ? 1.0 : 0.0) - x.R.a[i * 3 + k];
#pragma unroll
        for (int j = 0; j < 3; j++) m2[j] = (m1[0] * wh.a[j] + m1[1] * wh.a[3 + j]) + m1[2] * wh.a[6 + j];
        p0.x[i] = (m2[0] * z.v.x[0] + m2[1] * z.v.x[1]) + m2[2] * z.v.x[2];
    }
#pragma unroll 1
    for (int i = 0; i < 3; i++) x.p.x[i] = ((-x.R.a[i]) * p0.x[0] + (-x.R.a[3 + i]) * p0.x[1]) + (-x.R.a[6 + i]) * p0.x[2];
    return x;
}
template <class S> DEV Twist<S> apply(const Xf<S>& X, const Twist<S>& z) { Twist<S> r; r.w = X.R * z.w; r.v = X.R * (z.v - cross(X.p, z.w)); return r; }
template <class S> DEV Twist<S> invapply(const Xf<S>& X, const Twist<S>& z) {
    const M3<S> Rt = tr(X.R);
    Twist<S> r; r.w = Rt * z.w; r.v = Rt * z.v + cross(X.p, r.w); return r;
}
template <class S> DEV Wrench<S> invapply(const Xf<S>& X, const Wrench<S>& f) {   // KRC/spatial.cpp:210-214
    Wrench<S> r; r.f = mul_t(X.R, f.f); r.tau = mul_t(X.R, f.tau) + cross(X.p, r.f); return r;
}
template <class S> DEV Xf<S> compose(const Xf<S>& X, const Xf<S>& x2) { Xf<S> r; r.R = X.R * x2.R; r.p = x2.p + mul_t(x2.R, X.p); return r; }   // :236-243
template <class S> DEV Xf<S> inverse(const Xf<S>& X) { Xf<S> r; r.R = tr(X.R); r.p = (-X.R) * X.p; return r; }
// rigid inertia seen from a shifted/rotated frame (KRC/spatial.cpp:220-234); model set-up only, doubles only
DEV Inertia<double> apply(const Xf<double>& X, const Inertia<double>& I) {
    const M3<double> ph = hat(X.p), Rt = tr(X.R), mR = scaled_by(I.m, X.R);
    Inertia<double> n;
    n.m = I.m;
    n.mch = ((X.R * I.mch) * Rt) - ((mR * ph) * Rt);
    n.I_bar = ((X.R * (I.I_bar + scaled_by(2.0, I.mch) * ph)) - ((mR * ph) * ph)) * Rt;
    return n;
}

// ---- model ---------------------------------------------------------------------------------------------------
struct RawModel {   // the model file as parsed on the host (KRC/robot_models.cpp:20-122)
    int n;
    int parent[MAXJ];
    double twist[MAXJ][6], gravity[3];
    double m[MAXJ], I_bar[MAXJ][9], mch[MAXJ][9], X_R[MAXJ][9], X_p[MAXJ][3], com[MAXJ][3];
    double transI[MAXJ], friction[MAXJ], damping[MAXJ];
};
template <class S> struct JointConst {
    Twist<S> Sb;        // screw axis in the body frame           (KRC/rnea.cpp:41)
    Xf<S> XTinv;        // XTree[i].inverse()                     (:47)
    Inertia<S> I;
    S transI;
};
struct ControllerModel {
    int n;
    int mch_zero[MAXJ];   // 1: every entry of m c^ of joint i is exactly zero in both models
    double damping[MAXJ], friction[MAXJ];
    Twist<double> neg_gravity;
    Twist<Itv> neg_gravity_itv;
    JointConst<double> nom[MAXJ];
    JointConst<Itv> unc[MAXJ];
};
static_assert(sizeof(ControllerModel) < 32000, "ControllerModel must fit a kernel parameter");

template <class A> DEV V3<Itv> to_itv(const V3<A>& a) { V3<Itv> r; for (int i = 0; i < 3; i++) r.x[i] = cst<Itv>(a.x[i]); return r; }
template <class A> DEV M3<Itv> to_itv(const M3<A>& a) { M3<Itv> r; for (int i = 0; i < 9; i++) r.a[i] = cst<Itv>(a.a[i]); return r; }

__global__ void model_setup_kernel(RawModel raw, double eps, ControllerModel* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ControllerModel& M = *out;
    const int n = raw.n;
    M.n = n;
    Twist<double> S0[MAXJ];
    Xf<double> XT0[MAXJ], com[MAXJ];
    for (int i = 0; i < n; i++) {
        for (int e = 0; e < 3; e++) { S0[i].w.x[e] = raw.twist[i][e]; S0[i].v.x[e] = raw.twist[i][3 + e]; XT0[i].p.x[e] = raw.X_p[i][e]; com[i].p.x[e] = raw.com[i][e]; }
        for (int e = 0; e < 9; e++) XT0[i].R.a[e] = raw.X_R[i][e];
        com[i].R = mident<double>();
    }
    // joint-frame description -> Featherstone-style body-CoM description (KRC/robot_models.cpp:124-151)
    Twist<double> S[MAXJ];
    Xf<double> XT[MAXJ];
    Inertia<double> I[MAXJ];
    for (int i = 0; i < n; i++) {
        Xf<double> Xwj = XT0[i];
        for (int p = raw.parent[i]; p > -1; p = raw.parent[p]) Xwj = compose(Xwj, XT0[p]);
        S[i] = invapply(Xwj, S0[i]);
        Inertia<double> Ij;
        Ij.m = raw.m[i];
        for (int e = 0; e < 9; e++) { Ij.I_bar.a[e] = raw.I_bar[i][e]; Ij.mch.a[e] = raw.mch[i][e]; }
        I[i] = apply(com[i], Ij);
        Xf<double> prev; prev.R = mident<double>(); prev.p = vzero<double>();
        if (raw.parent[i] != -1) prev = com[raw.parent[i]];
        XT[i] = compose(prev, compose(inverse(XT0[i]), inverse(com[i])));
    }
    Twist<double> g; g.w = vzero<double>(); for (int e = 0; e < 3; e++) g.v.x[e] = raw.gravity[e];
    M.neg_gravity.w = -g.w; M.neg_gravity.v = -g.v;
    M.neg_gravity_itv.w = -to_itv(g.w); M.neg_gravity_itv.v = -to_itv(g.v);
    // nominal constants + interval model (KRC/robot_models.cpp:168-249) + q-independent RNEA terms (KRC/rnea.cpp:38-48)
    const double lowP = 1 - eps, highP = 1 + eps;
    Xf<double> Xbw;
    Xf<Itv> Xbw_i;
    for (int i = 0; i < n; i++) {
        M.damping[i] = raw.damping[i]; M.friction[i] = raw.friction[i];
        Xf<Itv> XTi; XTi.R = to_itv(XT[i].R); XTi.p = to_itv(XT[i].p);
        Twist<Itv> Si; Si.w = to_itv(S[i].w); Si.v = to_itv(S[i].v);
        if (i == 0) { Xbw = XT[i]; Xbw_i = XTi; }
        else { Xbw = compose(Xbw, XT[i]); Xbw_i = compose(Xbw_i, XTi); }
        M.nom[i].Sb = invapply(Xbw, S[i]);
        M.nom[i].XTinv = inverse(XT[i]);
        M.nom[i].I = I[i];
        M.nom[i].transI = raw.transI[i];
        M.unc[i].Sb = invapply(Xbw_i, Si);
        M.unc[i].XTinv = inverse(XTi);
        M.unc[i].I.m = itv(I[i].m * lowP, I[i].m * highP);
        M.unc[i].I.mch = to_itv(I[i].mch);
        for (int e = 0; e < 9; e++) {
            const double val = I[i].I_bar.a[e];
            M.unc[i].I.I_bar.a[e] = val >= 0 ? itv(val * lowP, val * highP) : itv(val * highP, val * lowP);
        }
        M.unc[i].transI = cst<Itv>(raw.transI[i]);
        int z = 1;
        for (int e = 0; e < 9; e++) if (I[i].mch.a[e] != 0.0) z = 0;
        M.mch_zero[i] = z;
    }
}

// ---- passivity RNEA, one sample, serial chain (KRC/rnea.cpp:6-93 / 95-185) ---------------------------------------
// MR: also carry the second pass RobustController::update makes for M(q) r (qd = qda = 0, qdd = r, no gravity,
// KRC/robust_controller.cpp:138-139) through the same joint transforms.  Its velocity-dependent terms are exact
// zeros in the reference (products with [0,0]) and are not formed here.
template <class S, bool MR>
__device__ void rnea_chain(const ControllerModel& M, const JointConst<S>* __restrict__ J, const Twist<S>& neg_gravity, bool gravity, bool friction,
                           const double* q, const double* qd, const double* qda, const double* qdd, const double* r, S* tau, S* Mr) {
    const int n = M.n;
    Xf<S> Xl[MAXJ];
    Wrench<S> f[MAXJ];
    Wrench<S> f2[MR ? MAXJ : 1];
    // parent-frame quantities carried down the chain: v, va, a and (MR) the acceleration of the M r pass.  Starting
    // them at 0, 0, -gravity, 0 makes the root joint the general case: Xl * 0 is an exact zero and 0 + x == x, and
    // at the root va == Sb * qda, so v x va is the v x (Sb * qda) of the other joints (KRC/rnea.cpp:50-73).
    constexpr int NCH = MR ? 4 : 3;
    Twist<S> ch[NCH];
    for (int c = 0; c < NCH; c++) { ch[c].w = vzero<S>(); ch[c].v = vzero<S>(); }
    if (gravity) ch[2] = neg_gravity;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        const JointConst<S>& Ji = J[i];
        double s, c;
        sincos(-q[i], &s, &c);
        Xl[i] = compose(joint_xf(Ji.Sb, s, 1 - c), Ji.XTinv);
#pragma unroll 1
        for (int k = 0; k < NCH; k++) ch[k] = apply(Xl[i], ch[k]);
        const Twist<S> sa = scaled(Ji.Sb, qda[i]);
        ch[0] = ch[0] + scaled(Ji.Sb, qd[i]);
        ch[1] = ch[1] + sa;
        ch[2] = (ch[2] + scaled(Ji.Sb, qdd[i])) + cross(ch[0], sa);
        if (MR) ch[NCH - 1] = ch[NCH - 1] + scaled(Ji.Sb, r[i]);
        const Twist<S>&v = ch[0], &va = ch[1];
        Wrench<S> vIv;
        vIv.tau = cross(va.w, Ji.I.I_bar * v.w);
        vIv.tau = vIv.tau + Ji.I.I_bar * cross(va.w, v.w);
        vIv.f = scaled_by(Ji.I.m, cross(va.w, v.v));
        const bool fmz = M.mch_zero[i] != 0;
        f[i] = apply(Ji.I, ch[2], fmz) + vIv;
        if (MR) f2[i] = apply(Ji.I, ch[NCH - 1], fmz);
    }
    Wrench<S> acc, acc2;
#pragma unroll 1
    for (int i = n - 1; i >= 0; i--) {
        const JointConst<S>& Ji = J[i];
        acc = (i == n - 1) ? f[i] : f[i] + acc;
        S t = dot(Ji.Sb, acc) + Ji.transI * cst<S>(qdd[i]);
        t = t + cst<S>(M.damping[i] * qd[i]);
        if (friction) t = t + cst<S>(M.friction[i] * (double)((qd[i] > 0) - (qd[i] < 0)));
        tau[i] = t;
        if (MR) {
            acc2 = (i == n - 1) ? f2[i] : f2[i] + acc2;
            Mr[i] = dot(Ji.Sb, acc2) + Ji.transI * cst<S>(r[i]);
            if (i > 0) acc2 = invapply(Xl[i], acc2);
        }
        if (i > 0) acc = invapply(Xl[i], acc);
    }
}

DEV double wrap_pi(double x) {   // clamp(), KRC/robust_controller.hpp:11-16
    const double two_pi = 6.283185307179586476925286766559;
    while (x >= 3.14159265358979323846) x -= two_pi;
    while (x < -3.14159265358979323846) x += two_pi;
    return x;
}

struct UpdateArgs {
    int count, method;   // 0 ARMOUR, 1 ALTHOFF
    double Kr[MAXJ];
    double par[3];       // ARMOUR: alpha, V_max, r_norm_threshold; ALTHOFF: Kp[0], Kp[1]
    const double *q, *q_d, *qd, *qd_d, *qd_dd;
    double *u, *u_nominal, *v, *u_interval, *V_sup;
    int* outside;
};

// RobustController::update (KRC/robust_controller.cpp:62-171), one sample per thread
template <bool ARMOUR_METHOD>
__global__ void __launch_bounds__(128, CTRL_MINB) controller_update_kernel(const __grid_constant__ ControllerModel M, const __grid_constant__ UpdateArgs A) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.count) return;
    const int n = M.n;
    const size_t o = (size_t)s * n;
    const double *q = A.q + o, *q_d = A.q_d + o;
    double qa_d[MAXJ], qa_dd[MAXJ], r[MAXJ], un[MAXJ];
    Itv ui[MAXJ], Mr[MAXJ];
    double r_sq = 0;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        const double q_diff = wrap_pi(A.qd[o + i] - q[i]);
        const double e_d = A.qd_d[o + i] - q_d[i];
        qa_d[i] = A.qd_d[o + i] + A.Kr[i] * q_diff;
        qa_dd[i] = A.qd_dd[o + i] + A.Kr[i] * e_d;
        r[i] = e_d + A.Kr[i] * q_diff;
        r_sq += r[i] * r[i];
    }
    // applyFriction = false in both MEX entry points (KRC/kinova_controller.cpp:49)
    rnea_chain<double, false>(M, M.nom, M.neg_gravity, true, false, q, q_d, qa_d, qa_dd, nullptr, un, nullptr);
    rnea_chain<Itv, ARMOUR_METHOD>(M, M.unc, M.neg_gravity_itv, true, false, q, q_d, qa_d, qa_dd, r, ui, Mr);
    bool outside = false;
    double bound_sq = 0;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        outside |= (un[i] > ui[i].hi) || (un[i] < ui[i].lo);
        const Itv phi = ui[i] - cst<Itv>(un[i]);
        const double b = fmax(fabs(phi.lo), fabs(phi.hi));
        bound_sq += b * b;
        if (A.u_interval) { A.u_interval[2 * (o + i)] = ui[i].lo; A.u_interval[2 * (o + i) + 1] = ui[i].hi; }
    }
    if (outside) atomicAdd(A.outside, 1);
    const double bound_norm = sqrt(bound_sq);
    double V_sup = 0, lambda = 0, r_norm = 1;
    bool active = true;
    if (ARMOUR_METHOD) {
        r_norm = sqrt(r_sq);
        active = r_norm > A.par[2];
        Itv V = cst<Itv>(0.0);
#pragma unroll 1
        for (int i = 0; i < n; i++) V = V + (0.5 * r[i]) * Mr[i];
        V_sup = active ? V.hi : 0.0;
        const double h = -V_sup + A.par[1];
        lambda = fmax(0.0, -A.par[0] * h / r_norm + bound_norm);
        if (A.V_sup) A.V_sup[s] = V_sup;
    }
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        double vi;
        if (ARMOUR_METHOD) vi = active ? -lambda * r[i] / r_norm : 0.0;
        else vi = -(A.par[1] * bound_norm + A.par[0]) * r[i];
        A.v[o + i] = vi;
        A.u_nominal[o + i] = un[i];
        A.u[o + i] = un[i] - vi;
    }
}

// passRNEA / passRNEA_Int on their own
__global__ void __launch_bounds__(128) controller_rnea_kernel(const __grid_constant__ ControllerModel M, int count, const double* q, const double* qd, const double* qda,
                                                              const double* qdd, int gravity, double* tau, double* tau_interval) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= count) return;
    const size_t o = (size_t)s * M.n;
    if (tau) {
        double t[MAXJ];
        rnea_chain<double, false>(M, M.nom, M.neg_gravity, gravity != 0, false, q + o, qd + o, qda + o, qdd + o, nullptr, t, nullptr);
        for (int i = 0; i < M.n; i++) tau[o + i] = t[i];
    }
    if (tau_interval) {
        Itv t[MAXJ];
        rnea_chain<Itv, false>(M, M.unc, M.neg_gravity_itv, gravity != 0, false, q + o, qd + o, qda + o, qdd + o, nullptr, t, nullptr);
        for (int i = 0; i < M.n; i++) { tau_interval[2 * (o + i)] = t[i].lo; tau_interval[2 * (o + i) + 1] = t[i].hi; }
    }
}

// ---- host side -----------------------------------------------------------------------------------------------
// Line format of the model file: "<field> [index] <v0 v1 ...>" (KRC/robot_models.cpp:28-122)
static bool parse_model_file(const char* path, RawModel& R, std::string& why) {
    std::ifstream in(path);
    if (!in.is_open()) { why = std::string("could not open the robot model file ") + path; return false; }
    memset(&R, 0, sizeof(R));
    for (int i = 0; i < MAXJ; i++) { R.parent[i] = -1; R.X_R[i][0] = R.X_R[i][4] = R.X_R[i][8] = 1.0; }
    std::string line;
    bool have_n = false;
    while (std::getline(in, line)) {
        std::string field, index, inside;
        bool in_array = false;
        for (char ch : line) {
            if (ch == '>') break;
            if (in_array) inside += ch;
            else if (ch == '<') in_array = true;
            else if (std::isalpha((unsigned char)ch) || ch == '_') field += ch;
            else if (std::isdigit((unsigned char)ch)) index += ch;
        }
        std::vector<double> v;
        { std::stringstream ss(inside); std::string tok; while (ss >> tok) v.push_back(strtod(tok.c_str(), nullptr)); }
        const int k = index.empty() ? 0 : atoi(index.c_str());
        auto need = [&](size_t cnt) { return v.size() >= cnt && k >= 0 && k < R.n; };
        if (field == "numJoints") {
            if (v.empty() || v[0] < 1 || v[0] > MAXJ) { why = "numJoints must be between 1 and 7"; return false; }
            R.n = (int)v[0]; have_n = true;
            continue;
        }
        if (!have_n) continue;
        bool ok = true;
        if (field == "twist") { ok = need(6); if (ok) for (int e = 0; e < 6; e++) R.twist[k][e] = v[e]; }
        else if (field == "gravity") { ok = v.size() >= 3; if (ok) for (int e = 0; e < 3; e++) R.gravity[e] = v[e]; }
        else if (field == "inertia") { ok = need(19); if (ok) { R.m[k] = v[0]; for (int e = 0; e < 9; e++) { R.I_bar[k][e] = v[1 + e]; R.mch[k][e] = v[10 + e]; } } }
        else if (field == "Xtree") { ok = need(12); if (ok) { for (int e = 0; e < 9; e++) R.X_R[k][e] = v[e]; for (int e = 0; e < 3; e++) R.X_p[k][e] = v[9 + e]; } }
        else if (field == "parent") { ok = (int)v.size() >= R.n; if (ok) for (int j = 0; j < R.n; j++) R.parent[j] = (int)v[j]; }
        else if (field == "CoM") { ok = need(3); if (ok) for (int e = 0; e < 3; e++) R.com[k][e] = v[e]; }
        else if (field == "transI") { ok = (int)v.size() >= R.n; if (ok) for (int j = 0; j < R.n; j++) R.transI[j] = v[j]; }
        else if (field == "friction") { ok = (int)v.size() >= R.n; if (ok) for (int j = 0; j < R.n; j++) R.friction[j] = v[j]; }
        else if (field == "damping") { ok = (int)v.size() >= R.n; if (ok) for (int j = 0; j < R.n; j++) R.damping[j] = v[j]; }
        if (!ok) { why = "malformed '" + field + "' line in the robot model file"; return false; }
    }
    if (!have_n) { why = "the robot model file has no numJoints line"; return false; }
    for (int i = 0; i < R.n; i++)
        if (R.parent[i] != i - 1) { why = "only serial chains are supported (parent[i] must be i - 1)"; return false; }
    return true;
}
}  // namespace ctrl
}  // namespace armour

using namespace armour;
using namespace armour::ctrl;

struct armour_controller {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ControllerModel model;
    int* d_outside = nullptr;
    double* d_in = nullptr;    // staging / resident inputs  [5][cap][n]
    double* d_out = nullptr;   // outputs [3][cap][n], then u_interval [cap][n][2], V_sup [cap]
    int cap = 0, resident = 0;
    float last_ms = 0;
    // large host-buffer calls: chunks ping-pong over two streams so that copies overlap the kernel; caller buffers are
    // page-locked once and remembered (closed-loop sweeps reuse them every tick)
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_join = nullptr;
    std::vector<std::pair<const char*, size_t>> pinned;
};

#define CK(call)                                                                                            \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess) return capi_fail(ARMOUR_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

static int ensure_capacity(armour_controller* c, int count) {
    if (count <= c->cap) return ARMOUR_OK;
    const size_t n = c->model.n;
    if (c->d_in) cudaFree(c->d_in);
    if (c->d_out) cudaFree(c->d_out);
    c->d_in = c->d_out = nullptr; c->cap = 0; c->resident = 0;
    CK(cudaMalloc(&c->d_in, sizeof(double) * 5 * count * n));
    CK(cudaMalloc(&c->d_out, sizeof(double) * ((size_t)3 * count * n + (size_t)2 * count * n + count)));
    c->cap = count;
    return ARMOUR_OK;
}

static int launch_update(armour_controller* c, int method, int count, const double* Kr, const double par[3], bool want_interval, bool want_V,
                         cudaStream_t stream = nullptr, size_t first = 0, bool whole_call = true) {
    if (!stream) stream = c->stream;
    const size_t n = c->model.n, N = (size_t)c->cap * n, o = first * n;
    UpdateArgs A;
    A.count = count; A.method = method;
    for (int i = 0; i < MAXJ; i++) A.Kr[i] = i < (int)n ? Kr[i] : 0.0;
    for (int i = 0; i < 3; i++) A.par[i] = par[i];
    A.q = c->d_in + o; A.q_d = c->d_in + N + o; A.qd = c->d_in + 2 * N + o; A.qd_d = c->d_in + 3 * N + o; A.qd_dd = c->d_in + 4 * N + o;
    A.u = c->d_out + o; A.u_nominal = c->d_out + N + o; A.v = c->d_out + 2 * N + o;
    A.u_interval = want_interval ? c->d_out + 3 * N + 2 * o : nullptr;
    A.V_sup = want_V ? c->d_out + 5 * N + first : nullptr;
    A.outside = c->d_outside;
    if (whole_call) CK(cudaMemsetAsync(c->d_outside, 0, sizeof(int), stream));
    // ARMOUR_TUNE_CTRL_THREADS / ARMOUR_TUNE_CTRL_SMEM: tuning knobs (block size; dummy dynamic shared memory that
    // caps the resident blocks per SM so the thread-local working set stays in L1/L2)
    static const int threads_env = getenv("ARMOUR_TUNE_CTRL_THREADS") ? atoi(getenv("ARMOUR_TUNE_CTRL_THREADS")) : 128;
    const int threads = (threads_env == 32 || threads_env == 64) ? threads_env : 128;   // the kernel is bounded for 128 threads
    static const int smem = getenv("ARMOUR_TUNE_CTRL_SMEM") ? atoi(getenv("ARMOUR_TUNE_CTRL_SMEM")) : 0;
    static bool attr_set = false;
    if (!attr_set && smem > 48 * 1024) {
        CK(cudaFuncSetAttribute(controller_update_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(controller_update_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    attr_set = true;
    const int blocks = (count + threads - 1) / threads;
    if (whole_call) CK(cudaEventRecord(c->ev0, stream));
    if (method == 0) controller_update_kernel<true><<<blocks, threads, smem, stream>>>(c->model, A);
    else controller_update_kernel<false><<<blocks, threads, smem, stream>>>(c->model, A);
    CK(cudaGetLastError());
    if (whole_call) CK(cudaEventRecord(c->ev1, stream));
    return ARMOUR_OK;
}

static int finish_update(armour_controller* c, int* outside) {
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, c->d_outside, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    if (outside) *outside = bad;
    if (bad) return capi_fail(ARMOUR_E_NUMERIC, "nominal model output falls outside interval output for " + std::to_string(bad) + " sample(s)");
    return ARMOUR_OK;
}

static int upload_states(armour_controller* c, int count, const double* const src[5]) {
    const size_t n = c->model.n, N = (size_t)c->cap * n;
    for (int a = 0; a < 5; a++) CK(cudaMemcpyAsync(c->d_in + a * N, src[a], sizeof(double) * count * n, cudaMemcpyHostToDevice, c->stream));
    return ARMOUR_OK;
}

// page-lock a caller buffer once (best effort: a buffer that cannot be registered is simply copied pageable)
static void pin_buffer(armour_controller* c, const void* p, size_t bytes) {
    const char* b = (const char*)p;
    for (auto& r : c->pinned) if (b >= r.first && b + bytes <= r.first + r.second) return;
    if (cudaHostRegister((void*)p, bytes, cudaHostRegisterPortable) == cudaSuccess) c->pinned.push_back({b, bytes});
    else cudaGetLastError();
}
static const int PIPE_MIN = 1 << 15, PIPE_CHUNK = 1 << 15;   // ticks

// count >= PIPE_MIN: chunks of PIPE_CHUNK ticks alternate between two streams (H2D, kernel, D2H per chunk)
static int host_update_pipelined(armour_controller* c, int method, int count, const double* Kr, const double par[3], const double* const src[5],
                                 double* u, double* u_nominal, double* v, double* u_interval, double* V_sup, int* outside) {
    const size_t n = c->model.n, N = (size_t)c->cap * n;
    for (int a = 0; a < 5; a++) pin_buffer(c, src[a], sizeof(double) * count * n);
    double* dst[3] = {u, u_nominal, v};
    for (int a = 0; a < 3; a++) pin_buffer(c, dst[a], sizeof(double) * count * n);
    if (u_interval) pin_buffer(c, u_interval, 2 * sizeof(double) * count * n);
    if (V_sup) pin_buffer(c, V_sup, sizeof(double) * count);
    CK(cudaMemsetAsync(c->d_outside, 0, sizeof(int), c->stream));
    CK(cudaEventRecord(c->ev0, c->stream));
    CK(cudaStreamWaitEvent(c->stream2, c->ev0, 0));
    int k = 0;
    for (size_t first = 0; first < (size_t)count; first += PIPE_CHUNK, k++) {
        const int m = (int)std::min<size_t>(PIPE_CHUNK, count - first);
        cudaStream_t st = (k & 1) ? c->stream2 : c->stream;
        const size_t o = first * n, bytes = sizeof(double) * m * n;
        for (int a = 0; a < 5; a++) CK(cudaMemcpyAsync(c->d_in + a * N + o, src[a] + o, bytes, cudaMemcpyHostToDevice, st));
        int rc = launch_update(c, method, m, Kr, par, u_interval != nullptr, V_sup != nullptr, st, first, false);
        if (rc) return rc;
        for (int a = 0; a < 3; a++) CK(cudaMemcpyAsync(dst[a] + o, c->d_out + a * N + o, bytes, cudaMemcpyDeviceToHost, st));
        if (u_interval) CK(cudaMemcpyAsync(u_interval + 2 * o, c->d_out + 3 * N + 2 * o, 2 * bytes, cudaMemcpyDeviceToHost, st));
        if (V_sup) CK(cudaMemcpyAsync(V_sup + first, c->d_out + 5 * N + first, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaEventRecord(c->ev_join, c->stream2));
    CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    CK(cudaEventRecord(c->ev1, c->stream));
    return finish_update(c, outside);
}

static int host_update(armour_controller* c, int method, int count, const double* Kr, const double par[3], const double* q, const double* q_d, const double* qd,
                       const double* qd_d, const double* qd_dd, double* u, double* u_nominal, double* v, double* u_interval, double* V_sup, int* outside) {
    if (!c || !Kr || !q || !q_d || !qd || !qd_d || !qd_dd || !u || !u_nominal || !v) return capi_fail(ARMOUR_E_INVALID, "null argument");
    if (count < 1) return capi_fail(ARMOUR_E_INVALID, "count must be positive");
    CK(cudaSetDevice(c->device));
    int rc = ensure_capacity(c, count);
    if (rc) return rc;
    c->resident = 0;
    const double* src[5] = {q, q_d, qd, qd_d, qd_dd};
    if (count >= PIPE_MIN) return host_update_pipelined(c, method, count, Kr, par, src, u, u_nominal, v, u_interval, V_sup, outside);
    if ((rc = upload_states(c, count, src))) return rc;
    if ((rc = launch_update(c, method, count, Kr, par, u_interval != nullptr, V_sup != nullptr))) return rc;
    const size_t n = c->model.n, N = (size_t)c->cap * n, bytes = sizeof(double) * count * n;
    CK(cudaMemcpyAsync(u, c->d_out, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(u_nominal, c->d_out + N, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(v, c->d_out + 2 * N, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (u_interval) CK(cudaMemcpyAsync(u_interval, c->d_out + 3 * N, 2 * bytes, cudaMemcpyDeviceToHost, c->stream));
    if (V_sup) CK(cudaMemcpyAsync(V_sup, c->d_out + 5 * N, sizeof(double) * count, cudaMemcpyDeviceToHost, c->stream));
    return finish_update(c, outside);
}

extern "C" {

int armour_controller_create(const char* robot_model_file, double model_uncertainty, int device, armour_controller** out) {
    if (!robot_model_file || !out) return capi_fail(ARMOUR_E_INVALID, "null argument");
    if (!(model_uncertainty >= 0) || model_uncertainty >= 1) return capi_fail(ARMOUR_E_INVALID, "model_uncertainty must be in [0, 1)");
    RawModel raw;
    std::string why;
    if (!parse_model_file(robot_model_file, raw, why)) return capi_fail(ARMOUR_E_INVALID, why);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return capi_fail(ARMOUR_E_CUDA, "no usable CUDA device (there is no CPU fallback)");
    if (device < 0) CK(cudaGetDevice(&device));
    if (device >= ndev) return capi_fail(ARMOUR_E_INVALID, "device ordinal out of range");
    CK(cudaSetDevice(device));
    armour_controller* c = new armour_controller();
    c->device = device;
    ControllerModel* d_model = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev_join);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_outside, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&d_model, sizeof(ControllerModel));
    if (e == cudaSuccess) {
        model_setup_kernel<<<1, 32, 0, c->stream>>>(raw, model_uncertainty, d_model);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&c->model, d_model, sizeof(ControllerModel), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (d_model) cudaFree(d_model);
    if (e != cudaSuccess) {
        armour_controller_destroy(c);
        return capi_fail(ARMOUR_E_CUDA, std::string("controller set-up: ") + cudaGetErrorString(e));
    }
    *out = c;
    return ARMOUR_OK;
}

void armour_controller_destroy(armour_controller* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->d_in) cudaFree(c->d_in);
    if (c->d_out) cudaFree(c->d_out);
    if (c->d_outside) cudaFree(c->d_outside);
    for (auto& r : c->pinned) cudaHostUnregister((void*)r.first);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int armour_controller_num_joints(armour_controller* c, int* num_joints) {
    if (!c || !num_joints) return capi_fail(ARMOUR_E_INVALID, "null argument");
    *num_joints = c->model.n;
    return ARMOUR_OK;
}

int armour_controller_update(armour_controller* c, int count, const double* Kr, double alpha, double V_max, double r_norm_threshold,
                             const double* q, const double* q_d, const double* qd, const double* qd_d, const double* qd_dd,
                             double* u, double* u_nominal, double* v, double* u_interval, double* V_sup, int* outside) {
    const double par[3] = {alpha, V_max, r_norm_threshold};
    return host_update(c, 0, count, Kr, par, q, q_d, qd, qd_d, qd_dd, u, u_nominal, v, u_interval, V_sup, outside);
}

int armour_controller_update_althoff(armour_controller* c, int count, const double* Kr, const double* Kp, const double* Ki, double max_error,
                                     const double* q, const double* q_d, const double* qd, const double* qd_d, const double* qd_dd,
                                     double* u, double* u_nominal, double* v, double* u_interval, int* outside) {
    if (!Kp || !Ki) return capi_fail(ARMOUR_E_INVALID, "null argument");
    (void)max_error;
    const double par[3] = {Kp[0], Kp[1], 0.0};   // phi_t = Kp[0] + Ki[0]*eAcc, kappa_t = Kp[1] + Ki[1]*eAcc with eAcc = 0
    return host_update(c, 1, count, Kr, par, q, q_d, qd, qd_d, qd_dd, u, u_nominal, v, u_interval, nullptr, outside);
}

int armour_controller_rnea(armour_controller* c, int count, const double* q, const double* qd, const double* qda, const double* qdd,
                           int apply_gravity, double* tau, double* tau_interval) {
    if (!c || !q || !qd || !qda || !qdd || (!tau && !tau_interval)) return capi_fail(ARMOUR_E_INVALID, "null argument");
    if (count < 1) return capi_fail(ARMOUR_E_INVALID, "count must be positive");
    CK(cudaSetDevice(c->device));
    int rc = ensure_capacity(c, count);
    if (rc) return rc;
    c->resident = 0;
    const size_t n = c->model.n, N = (size_t)c->cap * n, bytes = sizeof(double) * count * n;
    const double* src[5] = {q, qd, qda, qdd, qdd};
    if ((rc = upload_states(c, count, src))) return rc;
    const int threads = 128, blocks = (count + threads - 1) / threads;
    controller_rnea_kernel<<<blocks, threads, 0, c->stream>>>(c->model, count, c->d_in, c->d_in + N, c->d_in + 2 * N, c->d_in + 3 * N, apply_gravity,
                                                              tau ? c->d_out : nullptr, tau_interval ? c->d_out + 3 * N : nullptr);
    CK(cudaGetLastError());
    if (tau) CK(cudaMemcpyAsync(tau, c->d_out, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (tau_interval) CK(cudaMemcpyAsync(tau_interval, c->d_out + 3 * N, 2 * bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return ARMOUR_OK;
}

int armour_controller_upload(armour_controller* c, int count, const double* states) {
    if (!c || !states) return capi_fail(ARMOUR_E_INVALID, "null argument");
    if (count < 1) return capi_fail(ARMOUR_E_INVALID, "count must be positive");
    CK(cudaSetDevice(c->device));
    int rc = ensure_capacity(c, count);
    if (rc) return rc;
    const size_t n = c->model.n;
    const double* src[5];
    for (int a = 0; a < 5; a++) src[a] = states + (size_t)a * count * n;
    if ((rc = upload_states(c, count, src))) return rc;
    CK(cudaStreamSynchronize(c->stream));
    c->resident = count;
    return ARMOUR_OK;
}

int armour_controller_update_resident(armour_controller* c, const double* Kr, double alpha, double V_max, double r_norm_threshold) {
    if (!c || !Kr) return capi_fail(ARMOUR_E_INVALID, "null argument");
    if (c->resident < 1) return capi_fail(ARMOUR_E_STATE, "armour_controller_upload has not been called");
    CK(cudaSetDevice(c->device));
    const double par[3] = {alpha, V_max, r_norm_threshold};
    int rc = launch_update(c, 0, c->resident, Kr, par, false, false);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    return ARMOUR_OK;
}

int armour_controller_download(armour_controller* c, double* u_unom_v, int* outside) {
    if (!c || !u_unom_v) return capi_fail(ARMOUR_E_INVALID, "null argument");
    if (c->resident < 1) return capi_fail(ARMOUR_E_STATE, "armour_controller_upload has not been called");
    CK(cudaSetDevice(c->device));
    const size_t n = c->model.n, N = (size_t)c->cap * n, bytes = sizeof(double) * c->resident * n;
    for (int a = 0; a < 3; a++) CK(cudaMemcpyAsync(u_unom_v + (size_t)a * c->resident * n, c->d_out + a * N, bytes, cudaMemcpyDeviceToHost, c->stream));
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, c->d_outside, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (outside) *outside = bad;
    return ARMOUR_OK;
}

int armour_controller_release_host_buffers(armour_controller* c) {
    if (!c) return capi_fail(ARMOUR_E_INVALID, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    for (auto& r : c->pinned) cudaHostUnregister((void*)r.first);
    cudaGetLastError();
    c->pinned.clear();
    return ARMOUR_OK;
}

int armour_controller_last_ms(armour_controller* c, double* kernel_ms) {
    if (!c || !kernel_ms) return capi_fail(ARMOUR_E_INVALID, "null argument");
    *kernel_ms = c->last_ms;
    return ARMOUR_OK;
}

}  // extern "C"
