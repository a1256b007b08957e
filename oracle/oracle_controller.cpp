// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the reference's robust low-level controller,
// KRC = kinova_src/kinova_simulator_interfaces/kinova_robust_controllers_mex:
//   KRC/spatial.cpp:6-250, KRC/spatial_interval.cpp:6-231   twists, wrenches, rigid inertias, transforms
//   KRC/robot_models.cpp:20-151                            model text file + conversion to body-CoM frames
//   KRC/robot_models.cpp:168-249                           interval model (mass / inertia uncertainty)
//   KRC/rnea.cpp:6-93, 95-185                              passivity RNEA, nominal and interval
//   KRC/robust_controller.cpp:62-171                       RobustController::update, both robust-input methods
//   KRC/kinova_controller.cpp:15-84, kinova_controller_ALTHOFF.cpp:15-90   the two MEX entry points
//
// PARITY: pinned against the reference's own KRC sources compiled against the stand-in Eigen / Boost headers
// (oracle/_ref/libref_controller.so, tests/test_reference_pin.py: u_nominal and the interval RNEA bit-identical, u and v
// within 7e-15), and by properties (tests/test_controller.py): the nominal torque equals an independent numeric RNEA of
// the planner's robot constants, the interval torque encloses the torque of sampled models inside the uncertainty box.
//
// The spatial algebra is written once over the scalar type (double or orc::Interval); Eigen's evaluation order
// for 3-vectors / 3x3 matrices is followed (sums in ascending index order, products before sums).
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "oracle_pz.hpp"

namespace {
using orc::Interval;

// ---- scalar helpers so that one template serves both arithmetic types ----
inline double neg(double a) { return -a; }
inline Interval neg(const Interval& a) { return -a; }

template <class S> struct V3 { S x[3]; S& operator[](int i) { return x[i]; } const S& operator[](int i) const { return x[i]; } };
template <class S> struct M3 {
    S a[3][3];
    S& operator()(int r, int c) { return a[r][c]; }
    const S& operator()(int r, int c) const { return a[r][c]; }
};

template <class S> V3<S> zero3() { V3<S> v; for (int i = 0; i < 3; i++) v[i] = S(0.0); return v; }
template <class S> M3<S> zero33() { M3<S> m; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m(i, j) = S(0.0); return m; }
template <class S> M3<S> ident33() { M3<S> m = zero33<S>(); for (int i = 0; i < 3; i++) m(i, i) = S(1.0); return m; }
template <class S> V3<S> add(const V3<S>& a, const V3<S>& b) { V3<S> r; for (int i = 0; i < 3; i++) r[i] = a[i] + b[i]; return r; }
template <class S> V3<S> sub(const V3<S>& a, const V3<S>& b) { V3<S> r; for (int i = 0; i < 3; i++) r[i] = a[i] - b[i]; return r; }
template <class S> V3<S> neg(const V3<S>& a) { V3<S> r; for (int i = 0; i < 3; i++) r[i] = neg(a[i]); return r; }
template <class S> V3<S> scale(const V3<S>& a, const S& s) { V3<S> r; for (int i = 0; i < 3; i++) r[i] = a[i] * s; return r; }
template <class S> V3<S> scale(const S& s, const V3<S>& a) { V3<S> r; for (int i = 0; i < 3; i++) r[i] = s * a[i]; return r; }
template <class S> M3<S> add(const M3<S>& a, const M3<S>& b) { M3<S> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r(i, j) = a(i, j) + b(i, j); return r; }
template <class S> M3<S> sub(const M3<S>& a, const M3<S>& b) { M3<S> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r(i, j) = a(i, j) - b(i, j); return r; }
template <class S> M3<S> neg(const M3<S>& a) { M3<S> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r(i, j) = neg(a(i, j)); return r; }
template <class S> M3<S> scale(const M3<S>& a, const S& s) { M3<S> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r(i, j) = a(i, j) * s; return r; }
template <class S> M3<S> scale(const S& s, const M3<S>& a) { M3<S> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r(i, j) = s * a(i, j); return r; }
template <class S> M3<S> tr(const M3<S>& a) { M3<S> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r(i, j) = a(j, i); return r; }
template <class S> V3<S> mul(const M3<S>& m, const V3<S>& v) {
    V3<S> r;
    for (int i = 0; i < 3; i++) r[i] = (m(i, 0) * v[0] + m(i, 1) * v[1]) + m(i, 2) * v[2];
    return r;
}
template <class S> M3<S> mul(const M3<S>& a, const M3<S>& b) {
    M3<S> r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r(i, j) = (a(i, 0) * b(0, j) + a(i, 1) * b(1, j)) + a(i, 2) * b(2, j);
    return r;
}
template <class S> V3<S> cross(const V3<S>& a, const V3<S>& b) {
    V3<S> r;
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
    return r;
}
template <class S> S dot(const V3<S>& a, const V3<S>& b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
template <class S> M3<S> hat(const V3<S>& w) {   // the w_hat member every (Int)Twist carries (KRC/spatial.cpp:44-50)
    M3<S> m;
    m(0, 0) = S(0.0); m(0, 1) = neg(w[2]); m(0, 2) = w[1];
    m(1, 0) = w[2];   m(1, 1) = S(0.0);    m(1, 2) = neg(w[0]);
    m(2, 0) = neg(w[1]); m(2, 1) = w[0];   m(2, 2) = S(0.0);
    return m;
}

template <class S> struct Wrench { V3<S> tau, f; };
template <class S> Wrench<S> zero_wrench() { Wrench<S> w; w.tau = zero3<S>(); w.f = zero3<S>(); return w; }
template <class S> Wrench<S> add(const Wrench<S>& a, const Wrench<S>& b) { Wrench<S> w; w.tau = add(a.tau, b.tau); w.f = add(a.f, b.f); return w; }

template <class S> struct Twist { V3<S> w, v; M3<S> w_hat; };
template <class S> Twist<S> make_twist(const V3<S>& w, const V3<S>& v) { Twist<S> t; t.w = w; t.v = v; t.w_hat = hat(w); return t; }
template <class S> Twist<S> zero_twist() { Twist<S> t; t.w = zero3<S>(); t.v = zero3<S>(); t.w_hat = zero33<S>(); return t; }
template <class S> Twist<S> add(const Twist<S>& a, const Twist<S>& b) { return make_twist(add(a.w, b.w), add(a.v, b.v)); }
template <class S> Twist<S> scale(const Twist<S>& a, const S& s) { return make_twist(scale(a.w, s), scale(a.v, s)); }
template <class S> Twist<S> neg(const Twist<S>& a) { return make_twist(neg(a.w), neg(a.v)); }
template <class S> S dot(const Twist<S>& t, const Wrench<S>& f) { return dot(t.w, f.tau) + dot(t.v, f.f); }   // KRC/spatial.cpp:78-81
template <class S> Twist<S> cross(const Twist<S>& a, const Twist<S>& b) {   // KRC/spatial.cpp:83-87
    return make_twist(mul(a.w_hat, b.w), add(mul(a.w_hat, b.v), cross(a.v, b.w)));
}

template <class S> struct Inertia { S m; M3<S> I_bar, m_c_hat; };
template <class S> Wrench<S> apply(const Inertia<S>& I, const Twist<S>& z) {   // KRC/spatial.cpp:143-147
    Wrench<S> w;
    w.tau = add(mul(I.I_bar, z.w), mul(I.m_c_hat, z.v));
    w.f = sub(scale(I.m, z.v), mul(I.m_c_hat, z.w));
    return w;
}

template <class S> struct Xf { M3<S> R; V3<S> p; };
template <class S> Xf<S> ident_xf() { Xf<S> x; x.R = ident33<S>(); x.p = zero3<S>(); return x; }
// Rodrigues rotation about a screw axis (KRC/spatial.cpp:156-172, spatial_interval.cpp:145-156)
template <class S> Xf<S> joint_xf(const Twist<S>& z, double theta) {
    Xf<S> x;
    const S s(std::sin(theta)), c1(1 - std::cos(theta));
    x.R = add(add(ident33<S>(), scale(z.w_hat, s)), mul(scale(c1, z.w_hat), z.w_hat));
    const V3<S> p = mul(mul(sub(ident33<S>(), x.R), z.w_hat), z.v);
    x.p = mul(neg(tr(x.R)), p);
    return x;
}
template <class S> Twist<S> apply(const Xf<S>& X, const Twist<S>& z) { return make_twist(mul(X.R, z.w), mul(X.R, sub(z.v, cross(X.p, z.w)))); }
template <class S> Twist<S> invapply(const Xf<S>& X, const Twist<S>& z) {
    const M3<S> Rt = tr(X.R);
    const V3<S> w = mul(Rt, z.w);
    return make_twist(w, add(mul(Rt, z.v), cross(X.p, w)));
}
template <class S> Wrench<S> invapply(const Xf<S>& X, const Wrench<S>& f) {   // KRC/spatial.cpp:210-214
    const M3<S> Rt = tr(X.R);
    Wrench<S> r;
    r.tau = add(mul(Rt, f.tau), cross(X.p, mul(Rt, f.f)));
    r.f = mul(Rt, f.f);
    return r;
}
template <class S> Xf<S> compose(const Xf<S>& X, const Xf<S>& x2) {   // Transform::apply(Transform), KRC/spatial.cpp:236-243
    Xf<S> r;
    r.R = mul(X.R, x2.R);
    r.p = add(x2.p, mul(tr(x2.R), X.p));
    return r;
}
template <class S> Xf<S> inverse(const Xf<S>& X) { Xf<S> r; r.R = tr(X.R); r.p = mul(neg(X.R), X.p); return r; }
// rigid inertia seen from a shifted/rotated frame (KRC/spatial.cpp:220-234); only used while loading the model
Inertia<double> apply(const Xf<double>& X, const Inertia<double>& I) {
    const M3<double> p_hat = hat(X.p), Rt = tr(X.R);
    const M3<double> mRp_hat = mul(scale(I.m, X.R), p_hat);
    Inertia<double> n;
    n.m = I.m;
    n.m_c_hat = sub(mul(mul(X.R, I.m_c_hat), Rt), mul(mul(scale(I.m, X.R), p_hat), Rt));
    n.I_bar = mul(sub(mul(X.R, add(I.I_bar, mul(scale(2.0, I.m_c_hat), p_hat))), mul(mRp_hat, p_hat)), Rt);
    return n;
}

template <class S> struct Model {
    int n = 0;
    std::vector<Twist<S>> S_;
    std::vector<Inertia<S>> I;
    std::vector<Xf<S>> XTree;
    std::vector<int> lam;
    std::vector<S> transI;
    std::vector<double> friction, damping;
    Twist<S> gravity;
};

// KRC/robot_models.cpp:20-151.  Line format: "<field> [index] <v0 v1 ...>".
bool load_model(const char* path, Model<double>& M) {
    std::ifstream in(path);
    if (!in.is_open()) return false;
    std::vector<Xf<double>> com;
    std::string line;
    M.gravity = zero_twist<double>();
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
        { std::stringstream ss(inside); std::string tok; while (ss >> tok) v.push_back(std::stod(tok)); }
        const int k = index.empty() ? 0 : std::stoi(index);
        if (field == "numJoints") {
            M.n = (int)v.at(0);
            M.S_.assign(M.n, zero_twist<double>());
            Inertia<double> zi; zi.m = 0; zi.I_bar = zero33<double>(); zi.m_c_hat = zero33<double>();
            M.I.assign(M.n, zi);
            M.XTree.assign(M.n, ident_xf<double>());
            com.assign(M.n, ident_xf<double>());
            M.lam.assign(M.n, -1);
            M.transI.assign(M.n, 0.0);
            M.friction.assign(M.n, 0.0);
            M.damping.assign(M.n, 0.0);
            continue;
        }
        if (M.n == 0) continue;
        if (field == "twist") { V3<double> w{{v.at(0), v.at(1), v.at(2)}}, q{{v.at(3), v.at(4), v.at(5)}}; M.S_.at(k) = make_twist(w, q); }
        else if (field == "gravity") { V3<double> g{{v.at(0), v.at(1), v.at(2)}}; M.gravity = make_twist(zero3<double>(), g); }
        else if (field == "inertia") {
            M.I.at(k).m = v.at(0);
            for (int e = 0; e < 9; e++) { M.I[k].I_bar(e / 3, e % 3) = v.at(1 + e); M.I[k].m_c_hat(e / 3, e % 3) = v.at(10 + e); }
        }
        else if (field == "Xtree") { for (int e = 0; e < 9; e++) M.XTree.at(k).R(e / 3, e % 3) = v.at(e); for (int e = 0; e < 3; e++) M.XTree[k].p[e] = v.at(9 + e); }
        else if (field == "parent") { for (int j = 0; j < M.n; j++) M.lam[j] = (int)v.at(j); }
        else if (field == "CoM") { for (int e = 0; e < 3; e++) com.at(k).p[e] = v.at(e); }
        else if (field == "transI") { for (int j = 0; j < M.n; j++) M.transI[j] = v.at(j); }
        else if (field == "friction") { for (int j = 0; j < M.n; j++) M.friction[j] = v.at(j); }
        else if (field == "damping") { for (int j = 0; j < M.n; j++) M.damping[j] = v.at(j); }
    }
    if (M.n == 0) return false;
    // joint-frame description -> Featherstone-style body-CoM description (KRC/robot_models.cpp:124-151)
    std::vector<Twist<double>> S2(M.n);
    std::vector<Inertia<double>> I2(M.n);
    std::vector<Xf<double>> X2(M.n);
    for (int i = 0; i < M.n; i++) {
        Xf<double> Xwj = M.XTree[i];
        for (int p = M.lam[i]; p > -1; p = M.lam[p]) Xwj = compose(Xwj, M.XTree[p]);
        S2[i] = invapply(Xwj, M.S_[i]);
        I2[i] = apply(com[i], M.I[i]);
        const Xf<double> prev = M.lam[i] != -1 ? com[M.lam[i]] : ident_xf<double>();
        X2[i] = compose(prev, compose(inverse(M.XTree[i]), inverse(com[i])));
    }
    M.S_ = S2; M.I = I2; M.XTree = X2;
    return true;
}

template <class A, class B> V3<B> cast(const V3<A>& a) { V3<B> r; for (int i = 0; i < 3; i++) r[i] = B(a[i]); return r; }
template <class A, class B> M3<B> cast(const M3<A>& a) { M3<B> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r(i, j) = B(a(i, j)); return r; }

// KRC/robot_models.cpp:168-249
void make_interval_model(const Model<double>& M, double eps, Model<Interval>& Q) {
    Q.n = M.n; Q.lam = M.lam; Q.friction = M.friction; Q.damping = M.damping;
    Q.S_.resize(M.n); Q.I.resize(M.n); Q.XTree.resize(M.n); Q.transI.resize(M.n);
    const double lowP = 1 - eps, highP = 1 + eps;
    for (int i = 0; i < M.n; i++) {
        Q.S_[i].w = cast<double, Interval>(M.S_[i].w);
        Q.S_[i].v = cast<double, Interval>(M.S_[i].v);
        Q.S_[i].w_hat = cast<double, Interval>(M.S_[i].w_hat);
        Q.I[i].m = Interval(M.I[i].m * lowP, M.I[i].m * highP);
        Q.I[i].m_c_hat = cast<double, Interval>(M.I[i].m_c_hat);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) {
                const double val = M.I[i].I_bar(r, c);
                Q.I[i].I_bar(r, c) = val >= 0 ? Interval(val * lowP, val * highP) : Interval(val * highP, val * lowP);
            }
        Q.XTree[i].R = cast<double, Interval>(M.XTree[i].R);
        Q.XTree[i].p = cast<double, Interval>(M.XTree[i].p);
        Q.transI[i] = Interval(M.transI[i]);
    }
    Q.gravity.w = cast<double, Interval>(M.gravity.w);
    Q.gravity.v = cast<double, Interval>(M.gravity.v);
    Q.gravity.w_hat = cast<double, Interval>(M.gravity.w_hat);
}

// passRNEA / passRNEA_Int (KRC/rnea.cpp:6-93, 95-185)
template <class S>
void pass_rnea(const Model<S>& M, const double* q, const double* qd, const double* qda, const double* qdd, bool friction, bool gravity, S* tau) {
    const int n = M.n;
    std::vector<Twist<S>> v(n), va(n), a(n), Sb(n);
    std::vector<Wrench<S>> f(n);
    std::vector<Xf<S>> Xbw(n), Xl(n);
    const Twist<S> neg_gravity = gravity ? neg(M.gravity) : zero_twist<S>();
    for (int i = 0; i < n; i++) {
        const int li = M.lam[i];
        Xbw[i] = li != -1 ? compose(Xbw[li], M.XTree[i]) : M.XTree[i];
        Sb[i] = invapply(Xbw[i], M.S_[i]);
        Xl[i] = compose(joint_xf(Sb[i], -q[i]), inverse(M.XTree[i]));
        const Twist<S> sa = scale(Sb[i], S(qda[i]));
        if (li == -1) {
            v[i] = scale(Sb[i], S(qd[i]));
            va[i] = sa;
            a[i] = add(add(apply(Xl[i], neg_gravity), scale(Sb[i], S(qdd[i]))), cross(v[i], va[i]));
        } else {
            v[i] = add(apply(Xl[i], v[li]), scale(Sb[i], S(qd[i])));
            va[i] = add(apply(Xl[i], va[li]), sa);
            a[i] = add(add(apply(Xl[i], a[li]), scale(Sb[i], S(qdd[i]))), cross(v[i], sa));
        }
        Wrench<S> vIv;
        vIv.tau = cross(va[i].w, mul(M.I[i].I_bar, v[i].w));
        vIv.tau = add(vIv.tau, mul(M.I[i].I_bar, cross(va[i].w, v[i].w)));
        vIv.f = scale(M.I[i].m, cross(va[i].w, v[i].v));
        f[i] = add(apply(M.I[i], a[i]), vIv);
    }
    for (int i = n - 1; i >= 0; i--) {
        tau[i] = dot(Sb[i], f[i]) + M.transI[i] * S(qdd[i]);
        tau[i] = tau[i] + S(M.damping[i] * qd[i]);
        if (friction) tau[i] = tau[i] + S(M.friction[i] * (double)((qd[i] > 0) - (qd[i] < 0)));
        if (M.lam[i] != -1) f[M.lam[i]] = add(f[M.lam[i]], invapply(Xl[i], f[i]));
    }
}

struct Controller {
    Model<double> nominal;
    Model<Interval> interval;
};

inline double wrap_pi(double x) {   // clamp(), KRC/robust_controller.hpp:11-16
    const double two_pi = 6.283185307179586476925286766559;
    while (x >= M_PI) x -= two_pi;
    while (x < -M_PI) x += two_pi;
    return x;
}

// RobustController::update for one sample (KRC/robust_controller.cpp:62-171); method 0 = ARMOUR, 1 = ALTHOFF
// (deltaT = eAcc = 0 as both MEX files call it).  Returns 1 when the nominal torque leaves the interval torque
// (the reference prints and throws there).
int update_one(const Controller& C, int method, const double* Kr, const double* par, const double* q, const double* q_d, const double* qd,
               const double* qd_d, const double* qd_dd, double* u, double* u_nominal, double* v, double* u_itv, double* V_sup_out) {
    const int n = C.nominal.n;
    std::vector<double> qa_d(n), qa_dd(n), r(n), zero(n, 0.0), bound(n);
    for (int i = 0; i < n; i++) {
        const double q_diff = wrap_pi(qd[i] - q[i]);
        qa_d[i] = qd_d[i] + Kr[i] * q_diff;
        qa_dd[i] = qd_dd[i] + Kr[i] * (qd_d[i] - q_d[i]);
        r[i] = (qd_d[i] - q_d[i]) + Kr[i] * q_diff;
    }
    std::vector<Interval> ui(n);
    pass_rnea<double>(C.nominal, q, q_d, qa_d.data(), qa_dd.data(), false, true, u_nominal);   // applyFriction = false in both MEX files
    pass_rnea<Interval>(C.interval, q, q_d, qa_d.data(), qa_dd.data(), false, true, ui.data());
    int outside = 0;
    double bound_sq = 0;
    for (int i = 0; i < n; i++) {
        if (u_nominal[i] > ui[i].hi || u_nominal[i] < ui[i].lo) outside = 1;
        const Interval phi = ui[i] - Interval(u_nominal[i]);
        bound[i] = std::max(std::fabs(phi.lo), std::fabs(phi.hi));
        bound_sq += bound[i] * bound[i];
        if (u_itv) { u_itv[2 * i] = ui[i].lo; u_itv[2 * i + 1] = ui[i].hi; }
    }
    const double bound_norm = std::sqrt(bound_sq);
    for (int i = 0; i < n; i++) v[i] = 0;
    if (V_sup_out) *V_sup_out = 0;
    if (method == 1) {
        const double phi_t = par[0], kappa_t = par[1];   // Kp[0] + Ki[0]*0, Kp[1] + Ki[1]*0
        for (int i = 0; i < n; i++) v[i] = -(kappa_t * bound_norm + phi_t) * r[i];
    } else {
        const double alpha = par[0], V_max = par[1], r_norm_threshold = par[2];
        double rs = 0;
        for (int i = 0; i < n; i++) rs += r[i] * r[i];
        const double r_norm = std::sqrt(rs);
        if (r_norm > r_norm_threshold) {
            std::vector<Interval> Mr(n);
            pass_rnea<Interval>(C.interval, q, zero.data(), zero.data(), r.data(), false, false, Mr.data());
            Interval V_int(0.0);
            for (int i = 0; i < n; i++) V_int += (0.5 * r[i]) * Mr[i];
            const double V_sup = V_int.hi;
            if (V_sup_out) *V_sup_out = V_sup;
            const double h = -V_sup + V_max;
            const double lambda = std::max(0.0, -alpha * h / r_norm + bound_norm);
            for (int i = 0; i < n; i++) v[i] = -lambda * r[i] / r_norm;
        }
    }
    for (int i = 0; i < n; i++) u[i] = u_nominal[i] - v[i];
    return outside;
}
}  // namespace

extern "C" {
void* oracle_controller_create(const char* model_file, double eps) {
    Controller* C = new Controller();
    if (!load_model(model_file, C->nominal)) { delete C; return nullptr; }
    make_interval_model(C->nominal, eps, C->interval);
    return C;
}
void oracle_controller_destroy(void* h) { delete (Controller*)h; }
int oracle_controller_num_joints(void* h) { return ((Controller*)h)->nominal.n; }
// converted model as doubles: per joint S.w[3] S.v[3] m I_bar[9] m_c_hat[9] XTree.R[9] XTree.p[3] (row-major) = 40 values
void oracle_controller_get_model(void* h, double* out) {
    const Model<double>& M = ((Controller*)h)->nominal;
    for (int i = 0; i < M.n; i++) {
        double* o = out + 40 * i;
        for (int e = 0; e < 3; e++) { o[e] = M.S_[i].w[e]; o[3 + e] = M.S_[i].v[e]; o[37 + e] = M.XTree[i].p[e]; }
        o[6] = M.I[i].m;
        for (int e = 0; e < 9; e++) { o[7 + e] = M.I[i].I_bar(e / 3, e % 3); o[16 + e] = M.I[i].m_c_hat(e / 3, e % 3); o[28 + e] = M.XTree[i].R(e / 3, e % 3); }
        o[25] = M.transI[i]; o[26] = M.friction[i]; o[27] = M.damping[i];
    }
}
// arrays are [count][n]; u_itv ([count][n][2]) and V_sup ([count]) may be null.  Returns the number of samples whose
// nominal torque fell outside the interval torque.
int oracle_controller_update(void* h, int method, int count, const double* Kr, const double* par, const double* q, const double* q_d,
                             const double* qd, const double* qd_d, const double* qd_dd, double* u, double* u_nominal, double* v,
                             double* u_itv, double* V_sup, int num_threads) {
    const Controller& C = *(Controller*)h;
    const int n = C.nominal.n;
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad) num_threads(num_threads > 0 ? num_threads : 1)
    for (int s = 0; s < count; s++) {
        const size_t o = (size_t)s * n;
        bad += update_one(C, method, Kr, par, q + o, q_d + o, qd + o, qd_d + o, qd_dd + o, u + o, u_nominal + o, v + o,
                          u_itv ? u_itv + 2 * o : nullptr, V_sup ? V_sup + s : nullptr);
    }
    return bad;
}
// plain passivity RNEA of the nominal model (for the property tests)
void oracle_controller_rnea(void* h, const double* q, const double* qd, const double* qda, const double* qdd, int gravity, double* tau) {
    const Controller& C = *(Controller*)h;
    pass_rnea<double>(C.nominal, q, qd, qda, qdd, false, gravity != 0, tau);
}
void oracle_controller_rnea_interval(void* h, const double* q, const double* qd, const double* qda, const double* qdd, int gravity, double* tau_lo_hi) {
    const Controller& C = *(Controller*)h;
    std::vector<Interval> t(C.nominal.n);
    pass_rnea<Interval>(C.interval, q, qd, qda, qdd, false, gravity != 0, t.data());
    for (int i = 0; i < C.nominal.n; i++) { tau_lo_hi[2 * i] = t[i].lo; tau_lo_hi[2 * i + 1] = t[i].hi; }
}
}
