// Reach-set build kernel: one CTA per (problem, time interval).
//
// Stages, all on the device, for one interval s of one planning problem:
//   A  joint reach sets as PZs      BezierCurve::BezierCurve + makePolyZono   KPR/Trajectory.cu:15-254
//   B  PZ forward kinematics        KinematicsDynamics::fk                     KPR/Dynamics.cu:69-81
//      + reduce_link_PZ                                                        KPR/PZsparse.cu:370-402
//   C  PZ-RNEA                      KinematicsDynamics::rnea                   KPR/Dynamics.cu:83-181
//      The reference runs it twice (nominal / +-3 % inertial parameters, KPR/armour_main.cu:129-132).
//      Centres and monomials of the two runs are identical — the uncertainty only enters the interval
//      radius (KPR/PZsparse.cu:93-98, 944-989) — so one pass carries two radii.
//   M  disturbance radius, reduce(), torque radius                             KPR/armour_main.cu:135-205
// The half-space tables (stage D) are built by hyperplane_kernel in constraint_kernels.cu.
#include <cstdio>
#include "armour_launch.h"
#include "pz_engine.cuh"
#include "interval.cuh"
#include "hyperplane.cuh"

namespace armour {

__constant__ RobotModel c_robot;



__device__ Itv icos(Itv x) {   // boost::numeric::cos(interval), restated in oracle/oracle_pz.hpp
    bool negate = false;   // the library recurses once through -cos(x - pi); unrolled here (no device recursion)
    for (int it = 0; it < 3; it++) {
        const double yb = (x.lo < 0) ? PI2_LO : PI2_HI;
        const double n = floor(__ddiv_rd(x.lo, yb));
        const Itv tmp = isub(x, imul(n, itv(PI2_LO, PI2_HI)));
        if (__dadd_ru(tmp.hi, -tmp.lo) >= PI2_LO) return itv(-1.0, 1.0);
        if (tmp.lo >= PI_HI) { x = isub(tmp, itv(PI_LO, PI_HI)); negate = !negate; continue; }
        const double l = tmp.lo, u = tmp.hi;
        Itv r;
        if (u <= PI_LO) r = itv(widen_dn(cos(u)), widen_up(cos(l)));
        else if (u <= PI2_LO) r = itv(-1.0, widen_up(cos(fmin(__dadd_rd(PI2_LO, -u), l))));
        else r = itv(-1.0, 1.0);
        return negate ? ineg(r) : r;
    }
    return itv(-1.0, 1.0);
}
__device__ __forceinline__ Itv isin(Itv x) { return icos(isub(x, itv(PI_HALF_LO, PI_HALF_HI))); }

// ---- Bezier pieces (KPR/Trajectory.cu:812-822), expression order preserved ------------------
__device__ __forceinline__ double q_des_k_indep(double q0, double a, double b, double s) {
    const double s2 = s * s, s3 = s2 * s, s4 = s2 * s2, s5 = s4 * s;
    return q0 + a * s - 6 * a * s3 + 8 * a * s4 - 3 * a * s5 + (b * s2) * 0.5 - (3 * b * s3) * 0.5 + (3 * b * s4) * 0.5 - (b * s5) * 0.5;
}
__device__ __forceinline__ double qd_des_k_indep(double a, double b, double s) {
    const double sm1 = s - 1, s2 = s * s;
    return ((sm1 * sm1) * (2 * a + 4 * a * s + 2 * b * s - 30 * a * s2 - 5 * b * s2)) * 0.5;
}
__device__ __forceinline__ double qdd_des_k_indep(double a, double b, double s) {
    return -(s - 1.0) * (b - (36 * a + 8 * b) * s + (60 * a + 10 * b) * (s * s));
}
__device__ __forceinline__ void bound_k_indep(double lbv, double ubv, double s_lb, double s_ub, double e1, double v1, double e2, double v2, double& lo, double& hi) {
    lo = lbv; hi = ubv;
    if (lo > hi) { const double t = lo; lo = hi; hi = t; }
    if (s_lb < e1 && e1 < s_ub) { lo = fmin(lo, v1); hi = fmax(hi, v1); }
    if (s_lb < e2 && e2 < s_ub) { lo = fmin(lo, v2); hi = fmax(hi, v2); }
}

// scalar PZ from (centre, {k_i: c0, key1: c1}) with simplify()  (KPR/PZsparse.cu:120-136)
__device__ void small_scalar(PZ<1>& z, double center, u64 k0, double c0, u64 k1, double c1, double thr_sq) {
    int n = 0;
    double ind = 0.0, abss = 0.0;
    if (norm1(&c0) > thr_sq) { z.keys[n] = k0; z.coef[n] = c0; abss = __dadd_ru(abss, fabs(c0)); n++; } else ind = __dadd_ru(ind, fabs(c0));
    if (norm1(&c1) > thr_sq) { z.keys[n] = k1; z.coef[n] = c1; abss = __dadd_ru(abss, fabs(c1)); n++; } else ind = __dadd_ru(ind, fabs(c1));
    ind = __dmul_ru(ind, 1.0 + 0x1p-40);   // a dropped coefficient built from device libm values may be an ulp below the host's
    z.n = n; z.divM = FastDiv::magic(n); z.center[0] = center; z.ind[0][0] = ind; z.ind[1][0] = ind; z.abss[0] = abss; z.ormask = k0 | k1;
}
template <int D>
__device__ void export_small(SmallRec& r, const PZ<D>& z) {
    r.n = z.n; r.dim = D;
    for (int i = 0; i < z.n && i < SMALL_CAP; i++) { r.keys[i] = z.keys[i]; for (int c = 0; c < D; c++) r.coef[i][c] = z.coef[c * z.cap + i]; }
    for (int c = 0; c < D; c++) { r.center[c] = z.center[c]; r.ind[c] = z.ind[0][c]; }
}

// ARMTD comparison planner (KPA = kinova_planner_realtime_armtd_comparison): rotation PZs of joint i from the offline
// JRS zonotopes of cos / sin of the displacement (KPA/Trajectory.cu:29-81)
__device__ void rotation_from_cos_sin(int i, double cos_center, double cos_c0, double cos_c1, double sin_center, double sin_c0, double sin_c1, PZ<9>& R, PZ<9>& Rt, double thr);
__device__ void make_poly_zono_armtd(const Tables& tb, int prob, int s, int i, PZ<9>& R, PZ<9>& Rt, PZ<1>& cosq, PZ<1>& sinq, double thr) {
    const double* st = tb.state + (size_t)prob * 21;
    const int T = tb.T;
    const double* J = tb.jrs + (size_t)prob * 6 * NF * T;
    auto tab = [&](int which) { return J[((size_t)which * NF + i) * T + s]; };
    const double cos_q0 = cos(st[i]), sin_q0 = sin(st[i]);
    const double cos_center = cos_q0 * tab(0) - sin_q0 * tab(3);
    const double cos_c0 = cos_q0 * tab(1) - sin_q0 * tab(4);
    double cos_c1 = fabs(cos_q0) * tab(2) + fabs(sin_q0) * tab(5);
    cos_c1 *= 5.0;
    const double sin_center = cos_q0 * tab(3) + sin_q0 * tab(0);
    const double sin_c0 = cos_q0 * tab(4) + sin_q0 * tab(1);
    double sin_c1 = fabs(cos_q0) * tab(5) + fabs(sin_q0) * tab(2);
    sin_c1 *= 5.0;
    small_scalar(cosq, cos_center, key_k(i), cos_c0, key_cosqe(i), cos_c1, thr);
    small_scalar(sinq, sin_center, key_k(i), sin_c0, key_sinqe(i), sin_c1, thr);
    rotation_from_cos_sin(i, cos_center, cos_c0, cos_c1, sin_center, sin_c0, sin_c1, R, Rt, thr);
}

// R = R0(rpy) * R_axis(cos, sin) and its transpose   (Trajectory.cu:136-144; 3x3 ctor PZsparse.cu:179-205; operator* :864-994)
__device__ void rotation_from_cos_sin(int i, double cos_center, double cos_c0, double cos_c1, double sin_center, double sin_c0, double sin_c1, PZ<9>& R, PZ<9>& Rt, double thr) {
    const RobotModel& rm = c_robot;

        const int axis = rm.axes[i];
        const double* R0 = rm.R0[i];
        double Rzc[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        double Mk[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, Mc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, Ms[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        // index pairs of the rotation block for axis x / y / z (makeRotationMatrix, PZsparse.cu:211-250)
        int pcc0 = 0, pcc1 = 4, pns = 3, pps = 1;   // z: (0,0),(1,1) cos; (0,1) -sin; (1,0) +sin
        if (axis == 1) { pcc0 = 4; pcc1 = 8; pns = 7; pps = 5; }        // x: (1,1),(2,2); (1,2) -sin; (2,1) +sin
        else if (axis == 2) { pcc0 = 0; pcc1 = 8; pns = 2; pps = 6; }   // y: (0,0),(2,2); (2,0) -sin; (0,2) +sin
        int n = 0;
        double zind[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        u64 zk[3]; double zc[3][9];
        if (axis != 0) {
            Rzc[pcc0] = cos_center; Rzc[pcc1] = cos_center; Rzc[pns] = -sin_center; Rzc[pps] = sin_center;
            Mk[pcc0] = cos_c0; Mk[pcc1] = cos_c0; Mk[pns] = -sin_c0; Mk[pps] = sin_c0;   // k_i terms of cos and sin merge
            Mc[pcc0] = cos_c1; Mc[pcc1] = cos_c1;
            Ms[pns] = -sin_c1; Ms[pps] = sin_c1;
            const double* cand[3] = {Mk, Mc, Ms};
            const u64 ck[3] = {key_k(i), key_cosqe(i), key_sinqe(i)};
            for (int m = 0; m < 3; m++) {
                if (norm9(cand[m]) <= thr) { for (int c = 0; c < 9; c++) zind[c] = __dadd_ru(zind[c], __dmul_ru(fabs(cand[m][c]), 1.0 + 0x1p-40)); }
                else { zk[n] = ck[m]; for (int c = 0; c < 9; c++) zc[n][c] = cand[m][c]; n++; }
            }
        }
        double cen[9], ind[9], absR0[9], abss[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        matmat_rn(R0, Rzc, cen);
        for (int c = 0; c < 9; c++) absR0[c] = fabs(R0[c]);
        matmat_ru(absR0, zind, ind);
        int nr = 0;
        for (int m = 0; m < n; m++) {
            double pm[9];
            matmat_rn(R0, zc[m], pm);
            if (norm9(pm) <= thr) { for (int c = 0; c < 9; c++) ind[c] = __dadd_ru(ind[c], __dmul_ru(fabs(pm[c]), 1.0 + 0x1p-40)); }
            else {
                R.keys[nr] = zk[m];
                for (int c = 0; c < 9; c++) { R.coef[c * R.cap + nr] = pm[c]; abss[c] = __dadd_ru(abss[c], fabs(pm[c])); }
                nr++;
            }
        }
        R.n = nr; R.divM = FastDiv::magic(nr); R.ormask = key_k(i) | key_cosqe(i) | key_sinqe(i);
        for (int c = 0; c < 9; c++) { R.center[c] = cen[c]; R.ind[0][c] = ind[c]; R.ind[1][c] = ind[c]; R.abss[c] = abss[c]; }
        // R_t = R.transpose()
        Rt.n = nr; Rt.divM = FastDiv::magic(nr); Rt.ormask = R.ormask;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) {
                const int src = r + 3 * c, dstp = c + 3 * r;
                Rt.center[dstp] = cen[src]; Rt.ind[0][dstp] = ind[src]; Rt.ind[1][dstp] = ind[src]; Rt.abss[dstp] = abss[src];
                for (int m = 0; m < nr; m++) Rt.coef[dstp * Rt.cap + m] = R.coef[src * R.cap + m];
            }
        for (int m = 0; m < nr; m++) Rt.keys[m] = R.keys[m];
    }

// joint i, interval s: everything makePolyZono produces for that joint (KPR/Trajectory.cu:63-254)
__device__ void make_poly_zono_joint(const Tables& tb, int prob, int s, int i, PZ<9>& R, PZ<9>& Rt, PZ<1>& qd, PZ<1>& qda, PZ<1>& qdda,
                                     PZ<1>& cosq, PZ<1>& sinq, double thr) {   // thr: squared-domain threshold (Scratch::thr_sq)
    const RobotModel& rm = c_robot;
    const double* st = tb.state + (size_t)prob * 21;
    const double q0 = st[i], a = st[7 + i] * 1.0, b = st[14 + i] * 1.0 * 1.0;   // Tqd0, TTqdd0 with DURATION = 1
    const double ds = 1.0 / tb.T;
    const double s_lb = s * ds, s_ub = (s + 1) * ds;
    const double kr = tb.k_range[i];
    // k-independent stationary points (BezierCurve ctor, Trajectory.cu:36-58)
    double ext[6], extv[6];
    {
        const double den5 = 5 * (6 * a + b), den10 = 10 * (6 * a + b);
        const double r1 = sqrt(64 * (a * a) + 14 * a * b + (b * b));
        ext[0] = (2 * a + b + r1) / den5; ext[1] = (2 * a + b - r1) / den5;
        const double r2 = sqrt(6 * (54 * (a * a) + 14 * a * b + (b * b)));
        ext[2] = (18 * a + 4 * b + r2) / den10; ext[3] = (18 * a + 4 * b - r2) / den10;
        const double r3 = sqrt(2 * (152 * (a * a) + 42 * a * b + 3 * (b * b)));
        ext[4] = (32 * a + 6 * b + r3) / den10; ext[5] = (32 * a + 6 * b - r3) / den10;
        extv[0] = q_des_k_indep(q0, a, b, ext[0]); extv[1] = q_des_k_indep(q0, a, b, ext[1]);
        extv[2] = qd_des_k_indep(a, b, ext[2]); extv[3] = qd_des_k_indep(a, b, ext[3]);
        extv[4] = qdd_des_k_indep(a, b, ext[4]); extv[5] = qdd_des_k_indep(a, b, ext[5]);
    }
    // Part 1: q_des
    const double sl2 = s_lb * s_lb, sl3 = sl2 * s_lb, su2 = s_ub * s_ub, su3 = su2 * s_ub;
    double kd_lb = sl3 * (6 * sl2 - 15 * s_lb + 10);
    double kd_ub = su3 * (6 * su2 - 15 * s_ub + 10);
    double kd_center = (kd_ub + kd_lb) * 0.5;
    double kd_radius = (kd_ub - kd_lb) * 0.5 * kr;
    double ki_lb, ki_ub;
    bound_k_indep(q_des_k_indep(q0, a, b, s_lb), q_des_k_indep(q0, a, b, s_ub), s_lb, s_ub, ext[0], extv[0], ext[1], extv[1], ki_lb, ki_ub);
    double ki_radius = (ki_ub - ki_lb) * 0.5;
    const double qc = (ki_lb + ki_ub) * 0.5;
    // The reference's bound of the k-independent part may use an interior stationary value computed with
    // libm pow (Trajectory.cu:812-814); the device evaluates the same polynomial with products, so the radius
    // can differ by a few ulp of |q|.  It only feeds the interval remainders below: widen it by 2^-50 max(|q|, 1) (sound).
    const double rq = __dadd_ru(kd_radius + ki_radius + rm.qe, fmax(fmax(fabs(ki_lb), fabs(ki_ub)), 1.0) * 0x1p-50);
    const Itv q_rad = itv(-rq, rq);
    const Itv kspan = imul(kd_center, itv(-kr, kr));
    const double cqc = cos(qc), sqc = sin(qc);
    // Part 1.a / 1.b: first-order Taylor expansion with Lagrange remainder in interval arithmetic
    Itv cos_rad = isub(imul(ineg(q_rad), point_trig(sqc)), imul(imul(0.5, icos(iadd(iadd(qc, kspan), q_rad))), ipow2(iadd(q_rad, kspan))));
    Itv sin_rad = isub(imul(q_rad, point_trig(cqc)), imul(imul(0.5, isin(iadd(iadd(qc, kspan), q_rad))), ipow2(iadd(q_rad, kspan))));
    if (tb.cos_rem) {
        double* cr = tb.cos_rem + (((size_t)prob * NJ + i) * tb.T + s) * 2;
        double* sr = tb.sin_rem + (((size_t)prob * NJ + i) * tb.T + s) * 2;
        cr[0] = cos_rad.lo; cr[1] = cos_rad.hi; sr[0] = sin_rad.lo; sr[1] = sin_rad.hi;
    }
    const double cmid = (cos_rad.lo + cos_rad.hi) * 0.5;
    const double cos_center = cqc + cmid;
    cos_rad = isub(cos_rad, cmid);
    const double cos_c0 = -kd_center * kr * sqc, cos_c1 = (cos_rad.hi - cos_rad.lo) * 0.5;
    const double smid = (sin_rad.lo + sin_rad.hi) * 0.5;
    const double sin_center = sqc + smid;
    sin_rad = isub(sin_rad, smid);
    const double sin_c0 = kd_center * kr * cqc, sin_c1 = (sin_rad.hi - sin_rad.lo) * 0.5;
    small_scalar(cosq, cos_center, key_k(i), cos_c0, key_cosqe(i), cos_c1, thr);
    small_scalar(sinq, sin_center, key_k(i), sin_c0, key_sinqe(i), sin_c1, thr);

    rotation_from_cos_sin(i, cos_center, cos_c0, cos_c1, sin_center, sin_c0, sin_c1, R, Rt, thr);

    // Part 2: qd_des
    {
        const double slm1 = s_lb - 1, sum1 = s_ub - 1;
        kd_lb = 30 * sl2 * (slm1 * slm1);
        kd_ub = 30 * su2 * (sum1 * sum1);
        if (kd_ub < kd_lb) { const double t = kd_lb; kd_lb = kd_ub; kd_ub = t; }
        kd_center = (kd_ub + kd_lb) * 0.5 * kr;
        kd_radius = (kd_ub - kd_lb) * 0.5 * kr;
        bound_k_indep(qd_des_k_indep(a, b, s_lb), qd_des_k_indep(a, b, s_ub), s_lb, s_ub, ext[2], extv[2], ext[3], extv[3], ki_lb, ki_ub);
        ki_radius = (ki_ub - ki_lb) * 0.5;
        const double c = (ki_lb + ki_ub) * 0.5;
        small_scalar(qd, c, key_k(i), kd_center, key_qde(i), kd_radius + ki_radius + rm.qde, thr);
        small_scalar(qda, c, key_k(i), kd_center, key_qdae(i), kd_radius + ki_radius + rm.qdae, thr);
    }
    // Part 3: qdd_des
    {
        const double t_lb = 60 * s_lb * (2 * sl2 - 3 * s_lb + 1);
        const double t_ub = 60 * s_ub * (2 * su2 - 3 * s_ub + 1);
        if (s_ub <= rm.qdd_k_maxima) { kd_lb = t_lb; kd_ub = t_ub; }
        else if (s_lb <= rm.qdd_k_maxima) { kd_lb = fmin(t_lb, t_ub); kd_ub = rm.qdd_k_maxima_val; }
        else if (s_ub <= rm.qdd_k_minima) { kd_lb = t_ub; kd_ub = t_lb; }
        else if (s_lb <= rm.qdd_k_minima) { kd_lb = rm.qdd_k_minima_val; kd_ub = fmax(t_lb, t_ub); }
        else { kd_lb = t_lb; kd_ub = t_ub; }
        kd_center = (kd_ub + kd_lb) * 0.5 * kr;
        kd_radius = (kd_ub - kd_lb) * 0.5 * kr;
        bound_k_indep(qdd_des_k_indep(a, b, s_lb), qdd_des_k_indep(a, b, s_ub), s_lb, s_ub, ext[4], extv[4], ext[5], extv[5], ki_lb, ki_ub);
        ki_radius = (ki_ub - ki_lb) * 0.5;
        const double c = (ki_lb + ki_ub) * 0.5;
        small_scalar(qdda, c, key_k(i), kd_center, key_qddae(i), kd_radius + ki_radius + rm.qddae, thr);
    }
}

// reduce_link_PZ + export (KPR/PZsparse.cu:370-402, armour_main.cu:123-126)
template <int NT>
__device__ __noinline__ void export_link(Scratch& S, const Tables& tb, size_t rec, PZ<3>& L) {
    const int n = L.n;
    u16* flag = n <= S.scap ? S.sidx(0) : S.gidx[0];
    double red[3] = {0, 0, 0};
    double* gens = tb.gens + rec * 18;
    for (int i = gtid<NT>(); i < 18; i += NT) gens[i] = 0.0;
    gsync<NT>();
    for (int i = gtid<NT>(); i < n; i += NT) {
        const u64 k = L.keys[i];
        u16 f = 0;
        if (k < KEY_K_ONLY) f = 1;
        else if (k < KEY_K_LINKS_ONLY && (k & KEY_K_MASK) == 0) {
            // pure link-box generator: keys qde_0, qdae_0, qddae_0 in ascending order -> columns 0, 1, 2
            const int col = (k == key_qde(0)) ? 0 : (k == key_qdae(0)) ? 1 : (k == key_qddae(0)) ? 2 : -1;
            if (col < 0) set_err(S, ERR_LINK_GEN);
            else for (int c = 0; c < 3; c++) gens[col * 3 + c] = L.coef[c * L.cap + i];
        }
        else for (int c = 0; c < 3; c++) red[c] = __dadd_ru(red[c], fabs(L.coef[c * L.cap + i]));
        flag[i] = f;
    }
    gsync<NT>();
    // the link-generator columns must be packed in order of appearance (j++ in the reference): with the
    // three fixed keys all present (checked on the host model) appearance order equals column order.
    const int ipt = (n + NT - 1) / NT;
    const int g0 = min(gtid<NT>() * ipt, n), g1 = min(g0 + ipt, n);
    int cnt = 0;
    for (int g = g0; g < g1; g++) cnt += flag[g];
    int total;
    int off = block_scan_sum<NT, 3>(S, cnt, red, total);
    const int LCAP = tb.lcap;
    if (total > LCAP) { if (gtid<NT>() == 0) set_err(S, ERR_LTABLE_CAP); total = 0; }
    else {
        u64* ok = tb.l_keys + rec * LCAP;
        double* oc = tb.l_coef + rec * 3 * LCAP;
        for (int g = g0; g < g1; g++)
            if (flag[g]) { ok[off] = L.keys[g]; for (int c = 0; c < 3; c++) oc[c * LCAP + off] = L.coef[c * L.cap + g]; off++; }
    }
    if (gtid<NT>() < 3) {
        const int c = gtid<NT>();
        // final radius: inflate by 2^-40 so that last-bit libm differences upstream cannot make it
        // smaller than the host restatement's (see DESIGN.md "soundness of radii")
        const double r = __dmul_ru(__dadd_ru(L.ind[0][c], inflate(block_total<NT, 3>(S, c), n)), 1.0 + 0x1p-40);
        tb.l_center[rec * 3 + c] = L.center[c];
        tb.l_ind[rec * 3 + c] = r;
        gens[(3 + c) * 3 + c] = r;
        if (c == 0) tb.l_n[rec] = total;
    }
    gsync<NT>();
    phase_mark(PH_EXPORT);
}

// disturbance radius, reduce() and export of one torque PZ (KPR/armour_main.cu:135-142, PZsparse.cu:352-368)
template <int NT>
__device__ __noinline__ void export_torque(Scratch& S, const Tables& tb, size_t rec, PZ<1>& U) {
    const int n = U.n;
    u16* flag = n <= S.scap ? S.sidx(0) : S.gidx[0];
    double red[1] = {0};
    for (int i = gtid<NT>(); i < n; i += NT) {
        const u64 k = U.keys[i];
        const u16 f = (k < KEY_K_ONLY) ? 1 : 0;
        if (!f) red[0] = __dadd_ru(red[0], fabs(U.coef[i]));
        flag[i] = f;
    }
    gsync<NT>();
    const int ipt = (n + NT - 1) / NT;
    const int g0 = min(gtid<NT>() * ipt, n), g1 = min(g0 + ipt, n);
    int cnt = 0;
    for (int g = g0; g < g1; g++) cnt += flag[g];
    int total;
    int off = block_scan_sum<NT, 1>(S, cnt, red, total);
    const int UCAP = tb.ucap;
    if (total > UCAP) { if (gtid<NT>() == 0) set_err(S, ERR_TABLE_CAP); total = 0; }
    else {
        u64* ok = tb.u_keys + rec * UCAP;
        double* oc = tb.u_coef + rec * UCAP;
        for (int g = g0; g < g1; g++)
            if (flag[g]) { ok[off] = U.keys[g]; oc[off] = U.coef[g]; off++; }
    }
    if (gtid<NT>() == 0) {
        tb.u_n[rec] = total;
        tb.u_center[rec] = U.center[0];
        tb.dist_rad[rec] = __dmul_ru(__dadd_ru(U.ind[1][0], U.ind[0][0]), 1.0 + 0x1p-40);
        tb.u_ind[rec] = __dmul_ru(__dadd_ru(U.ind[0][0], inflate(block_total<NT, 1>(S, 0), n)), 1.0 + 0x1p-40);
    }
    gsync<NT>();
    phase_mark(PH_EXPORT);
}

// ---------------------------------------------------------------------------------------------
// Working set of one interval.  The RNEA joint state (w, wdot, w_aux, linear_acc) is double-buffered by joint
// parity so that the force/moment computation of joint i can run beside the chain update of joint i + 1.
// Descriptors are split by how often they are touched.  The hot ones (the recursion state, read and rewritten by nearly every
// operation) always live in shared memory, next to the per-group temporaries; the per-joint ones (each used by a handful of
// operations of one joint) live in shared memory in the one-plan shapes and in the CTA's arena slice (global memory, L1-cached)
// in the narrow sweep shapes, where 12-24 CTAs share an SM's shared memory.
struct HotSlots {
    PZ<3> W[2], WD[2], WA[2], LA[2];
    PZ<3> Fv, Nv, FKT;
    PZ<9> FKR;
};
struct ColdSlots {
    PZ<3> C1[NJ], C2[NJ];                // two-group hand-off: cross(com, F_i), cross(trans_{i+1}, R_{i+1} f_{i+1})
    PZ<3> F[NJ], N[NJ], LINK[NJ], link0[NJ];
    PZ<9> R[NJ + 1], Rt[NJ];
    PZ<1> qd[NJ], qda[NJ], qdda[NJ], u[NJ], cosq[NJ], sinq[NJ];
};
struct Slots {
    HotSlots& h;
    ColdSlots& c;
};
constexpr int N_BIG3 = 8 + 10 + 3 + 5 * NJ;   // W/WD/WA/LA x 2, two temporary sets, Fv/Nv/FKT, C1/C2/F/N/LINK per joint

// Debug builds (-DARMOUR_ARENA_CANARY, `make canary`): every PZ slot of the arena is followed by a guard word that the kernel
// checks after its last interval; a write past a slot's capacity — which the capacity checks of the engine are there to
// prevent — sets ERR_CANARY.  (compute-sanitizer is closed on this GPU pool: profiles/r2_sanitizer_refusal.txt.)
#ifdef ARMOUR_ARENA_CANARY
constexpr u64 CANARY_WORD = 0xA5C3F00DDEADBEEFull;
constexpr int CANARY_MAX = 192;
__shared__ u64* g_canary[CANARY_MAX];
__shared__ int g_ncanary;
#define CANARY_BYTES 16
#else
#define CANARY_BYTES 0
#endif
__host__ __device__ constexpr size_t cold_arena_bytes() { return (sizeof(ColdSlots) + 255) & ~(size_t)255; }
size_t arena_bytes(int mcap, int ncap, int groups) {
    size_t b = cold_arena_bytes();                                // per-joint descriptors (narrow sweep shapes)
    b += (size_t)CANARY_BYTES * 160;                              // guard words of the debug build (one per slot)
    b += (size_t)N_BIG3 * mcap * (8 + 3 * 8);                    // big 3-vectors
    b += (size_t)NJ * SMALL_CAP * (8 + 3 * 8);                    // link0
    b += (size_t)mcap * (8 + 9 * 8);                              // FK_R
    b += (size_t)(2 * NJ + 1) * SMALL_CAP * (8 + 9 * 8);          // R, R_t
    b += (size_t)5 * NJ * SMALL_CAP * 16;                         // qd, qda, qdda, cos, sin
    b += (size_t)NJ * mcap * 16;                                  // u
    b += (size_t)groups * Scratch::gmem_bytes(ncap);              // per-group sort / staging buffers of large operations
    return (b + 255) & ~(size_t)255;
}

template <int D>
__device__ __noinline__ char* carve(PZ<D>& z, char* p, int cap) {
    z.cap = cap; z.n = 0; z.divM = FastDiv::magic(0); z.ormask = 0;
    z.keys = (u64*)p; p += (size_t)cap * 8;
    z.coef = (double*)p; p += (size_t)cap * 8 * D;
#ifdef ARMOUR_ARENA_CANARY
    if (g_ncanary < CANARY_MAX) { g_canary[g_ncanary++] = (u64*)p; *(u64*)p = CANARY_WORD; }
    p += CANARY_BYTES;
#endif
    for (int c = 0; c < D; c++) { z.center[c] = 0; z.ind[0][c] = 0; z.ind[1][c] = 0; z.abss[c] = 0; }
    return p;
}

// ---- hand-off between the two thread groups of a CTA --------------------------------------------------------
// Monotonic counters in shared memory.  The producing group signals after an operation (whose closing group
// barrier has made every thread's writes visible to its thread 0); the consuming group's thread 0 polls, then the
// group barrier releases the rest.  The poll is bounded: a lost signal becomes an error word, never a hang.
template <int NT>
__device__ __forceinline__ void group_signal(volatile int* flag, int value) {
    if (gtid<NT>() == 0) { __threadfence_block(); *flag = value; }
}
template <int NT>
__device__ __forceinline__ void group_wait(Scratch& S, volatile int* flag, int needed) {
    if (gtid<NT>() == 0) {
        const long long t0 = clock64();
        while (*flag < needed) {
            if (clock64() - t0 > (1ll << 32)) { set_err(S, ERR_SYNC); break; }
        }
        __threadfence_block();
    }
    gsync<NT>();
}

// group-uniform test of a hand-off counter (thread 0 reads, the group barrier broadcasts)
template <int NT>
__device__ __forceinline__ bool group_ready(Scratch& S, volatile int* flag, int needed) {
    if (gtid<NT>() == 0) { const int ok = *flag >= needed; if (ok) __threadfence_block(); S.iscan[17] = ok; }
    gsync<NT>();
    const bool r = S.iscan[17] != 0;
    gsync<NT>();
    return r;
}

// Opt-in experiment (-DARMOUR_SHARED_FK=1), measured SLOWER and therefore off: forward kinematics is a chain of its own (FK_R,
// FK_T -> link PZs) that nothing else in the interval reads, so either group could advance it whenever it would otherwise wait
// for the other one (by default group 0 runs it in its waits and at its tail; group 0 is busy 99 % of the kernel, group 1 85 %).
// One joint at a time: a group claims the chain (fk_lock), runs the joint with its own temporaries, publishes fk_next and
// releases.  Parity identical (51 GPU tests), but 0.69-0.73 ms against 0.655-0.695 ms per plan: a joint taken by group 1 in a
// wait delays the force computations group 0 is about to need, and the chain itself stays sequential whoever runs it.
#ifndef ARMOUR_SHARED_FK
#define ARMOUR_SHARED_FK 0
#endif
template <int NT>
__device__ __noinline__ bool try_fk(Scratch& S, Slots& Z, PZ<3>* T, const Tables& tb, size_t rec0, volatile int* fk_next, int* fk_lock);
template <int NT>
__device__ __forceinline__ void wait_or_fk(Scratch& S, Slots& Z, PZ<3>* T, const Tables& tb, size_t rec0, volatile int* flag, int needed, volatile int* fk_next, int* fk_lock);
template <int NT>
__device__ __forceinline__ void drain_fk(Scratch& S, Slots& Z, PZ<3>* T, const Tables& tb, size_t rec0, volatile int* fk_next, int* fk_lock);

// ---- the per-interval program, in four pieces --------------------------------------------------------------
// RNEA forward recursion, joint chain (KPR/Dynamics.cu:102-137): state `p` (after joint i-1) -> state `c`
template <int NT>
__device__ void chain_joint(Scratch& S, Slots& Z, PZ<3>* T, int i, int p, int c) {
    const RobotModel& rm = c_robot;
    const double zero3[3] = {0, 0, 0};
    const int axis = rm.axes[i];
    const int row = (axis < 0 ? -axis : axis) - 1;
    // linear_acc = R_t * (linear_acc + cross(wdot, trans) + cross(w, cross(w_aux, trans)))   (line 16)
    pz_cross_const<NT>(S, T[0], Z.h.WD[p], rm.trans[i], false);
    pz_cross_const<NT>(S, T[1], Z.h.WA[p], rm.trans[i], false);
    pz_cross_pp<NT>(S, T[2], Z.h.W[p], T[1]);
    pz_add3<NT>(S, T[3], Z.h.LA[p], T[0]);
    pz_add3<NT>(S, T[3], T[3], T[2]);
    pz_mul<NT, 9, 3, 3>(S, Z.h.LA[c], Z.c.Rt[i], T[3]);
    // w = R_t * w (+ qd_des on the joint axis)                                               (line 13)
    pz_mul<NT, 9, 3, 3>(S, Z.h.W[c], Z.c.Rt[i], Z.h.W[p]);
    if (axis != 0) pz_add_one_dim<NT>(S, Z.h.W[c], Z.h.W[c], Z.c.qd[i], row);
    // w_aux = R_t * w_aux                                                                     (line 14)
    pz_mul<NT, 9, 3, 3>(S, Z.h.WA[c], Z.c.Rt[i], Z.h.WA[p]);
    // wdot = R_t * wdot (+ cross(w_aux, qd_des * z) + qdda_des on the axis)                   (line 15)
    pz_mul<NT, 9, 3, 3>(S, Z.h.WD[c], Z.c.Rt[i], Z.h.WD[p]);
    if (axis != 0) {
        pz_set_const<NT, 3>(T[0], zero3);
        pz_add_one_dim<NT>(S, T[0], T[0], Z.c.qd[i], row);
        pz_cross_pp<NT>(S, T[1], Z.h.WA[c], T[0]);
        pz_add3<NT>(S, Z.h.WD[c], Z.h.WD[c], T[1]);
        pz_add_one_dim<NT>(S, Z.h.WD[c], Z.h.WD[c], Z.c.qdda[i], row);
        pz_add_one_dim<NT>(S, Z.h.WA[c], Z.h.WA[c], Z.c.qda[i], row);
    }
}
// F = m * (linear_acc + cross(wdot, com) + cross(w, cross(w_aux, com))), N = I * wdot + cross(w_aux, I * w)
// from the state after joint i (KPR/Dynamics.cu:146-154)
template <int NT>
__device__ void force_joint(Scratch& S, Slots& Z, PZ<3>* T, const Tables& tb, int i, int c, const PZ<3>& LA) {
    const RobotModel& rm = c_robot;
    pz_cross_const<NT>(S, T[0], Z.h.WD[c], rm.com[i], false);
    pz_cross_const<NT>(S, T[1], Z.h.WA[c], rm.com[i], false);
    pz_cross_pp<NT>(S, T[2], Z.h.W[c], T[1]);
    pz_add3<NT>(S, T[3], LA, T[0]);
    pz_add3<NT>(S, T[3], T[3], T[2]);
    {
        const double m0 = 0.0, m1 = __dmul_ru(tb.mass_unc, fabs(rm.mass[i]));
        pz_const_left<NT>(S, Z.c.F[i], &rm.mass[i], &m0, &m1, true, T[3]);
    }
    {
        double I0[9], I1[9];
        for (int k = 0; k < 9; k++) { I0[k] = 0.0; I1[k] = __dmul_ru(tb.inertia_unc, fabs(rm.inertia[i][k])); }
        pz_const_left<NT>(S, T[0], rm.inertia[i], I0, I1, false, Z.h.WD[c]);
        pz_const_left<NT>(S, T[1], rm.inertia[i], I0, I1, false, Z.h.W[c]);
    }
    pz_cross_pp<NT>(S, T[2], Z.h.WA[c], T[1]);
    pz_add3<NT>(S, Z.c.N[i], T[0], T[2]);
}
// forward kinematics + reduce_link_PZ, one joint (KPR/Dynamics.cu:69-81, armour_main.cu:123-126)
template <int NT>
__device__ void fk_joint(Scratch& S, Slots& Z, PZ<3>* T, const Tables& tb, size_t rec0, int i) {
    const RobotModel& rm = c_robot;
    const double zero3[3] = {0, 0, 0};
    if (i == 0) { pz_set_const<NT, 9>(Z.h.FKR, rm.R0[NJ]); pz_set_const<NT, 3>(Z.h.FKT, zero3); }
    pz_const_right<NT>(S, T[4], Z.h.FKR, rm.trans[i]);          // FK_R * P
    pz_add3<NT>(S, Z.h.FKT, Z.h.FKT, T[4]);                       // FK_T = FK_T + FK_R * P
    pz_mul<NT, 9, 9, 9>(S, Z.h.FKR, Z.h.FKR, Z.c.R[i]);             // FK_R = FK_R * R_i
    pz_mul<NT, 9, 3, 3>(S, T[4], Z.h.FKR, Z.c.link0[i]);          // FK_R * link_i
    pz_add3<NT>(S, Z.c.LINK[i], T[4], Z.h.FKT);                   //          + FK_T
    export_link<NT>(S, tb, rec0 + i, Z.c.LINK[i]);
}
template <int NT>
__device__ void forward_kinematics(Scratch& S, Slots& Z, PZ<3>* T, const Tables& tb, size_t rec0) {
    #pragma unroll 1
    for (int i = 0; i < NJ; i++) fk_joint<NT>(S, Z, T, tb, rec0, i);
}
template <int NT>
__device__ __noinline__ bool try_fk(Scratch& S, Slots& Z, PZ<3>* T, const Tables& tb, size_t rec0, volatile int* fk_next, int* fk_lock) {
    if (gtid<NT>() == 0) {
        int i = -1;
        if (*fk_next < NJ && atomicCAS(fk_lock, 0, 1) == 0) {
            __threadfence_block();
            i = *fk_next;
            if (i >= NJ) { i = -1; atomicExch(fk_lock, 0); }
        }
        S.iscan[17] = i;
    }
    gsync<NT>();
    const int i = S.iscan[17];
    gsync<NT>();
    if (i < 0) return false;
    fk_joint<NT>(S, Z, T, tb, rec0, i);
    if (gtid<NT>() == 0) { __threadfence_block(); *fk_next = i + 1; __threadfence_block(); atomicExch(fk_lock, 0); }
    return true;
}
// wait for a hand-off counter, advancing the forward kinematics while it is not there yet
template <int NT>
__device__ __forceinline__ void wait_or_fk(Scratch& S, Slots& Z, PZ<3>* T, const Tables& tb, size_t rec0, volatile int* flag, int needed, volatile int* fk_next, int* fk_lock) {
    while (!group_ready<NT>(S, flag, needed)) {
        if (!try_fk<NT>(S, Z, T, tb, rec0, fk_next, fk_lock)) { group_wait<NT>(S, flag, needed); return; }
    }
}
// end of a group's program: the chain must be complete before the interval's results are used
template <int NT>
__device__ __forceinline__ void drain_fk(Scratch& S, Slots& Z, PZ<3>* T, const Tables& tb, size_t rec0, volatile int* fk_next, int* fk_lock) {
    for (;;) {
        if (try_fk<NT>(S, Z, T, tb, rec0, fk_next, fk_lock)) continue;
        if (gtid<NT>() == 0) {   // nothing claimed: finished, or the other group is inside a joint — wait for it to release
            const long long t0 = clock64();
            bool lost = false;
            while (*fk_next < NJ && *(volatile int*)fk_lock != 0) {
                if (clock64() - t0 > (1ll << 32)) { set_err(S, ERR_SYNC); lost = true; break; }
            }
            __threadfence_block();
            S.iscan[17] = (*fk_next >= NJ || lost) ? 1 : 0;
        }
        gsync<NT>();
        const bool done = S.iscan[17] != 0;
        gsync<NT>();
        if (done) return;
    }
}
// RNEA reverse recursion for joint i (KPR/Dynamics.cu:161-180)
template <int NT>
__device__ void backward_joint(Scratch& S, Slots& Z, PZ<3>* T, int i) {
    const RobotModel& rm = c_robot;
    const int axis = rm.axes[i];
    const int row = (axis < 0 ? -axis : axis) - 1;
    // n = N + R * n + cross(com, F) + cross(trans_{i+1}, R * f)                              (line 29)
    pz_mul<NT, 9, 3, 3>(S, T[0], Z.c.R[i + 1], Z.h.Nv);
    pz_add3<NT>(S, T[1], Z.c.N[i], T[0]);
    pz_cross_const<NT>(S, T[2], Z.c.F[i], rm.com[i], true);
    pz_add3<NT>(S, T[1], T[1], T[2]);
    pz_mul<NT, 9, 3, 3>(S, T[3], Z.c.R[i + 1], Z.h.Fv);             // R * f (the reference evaluates it twice)
    pz_cross_const<NT>(S, T[2], T[3], rm.trans[i + 1], true);
    pz_add3<NT>(S, Z.h.Nv, T[1], T[2]);
    // f = R * f + F                                                                           (line 28)
    pz_add3<NT>(S, Z.h.Fv, T[3], Z.c.F[i]);
    if (axis != 0) {
        // u = n(axis) + armature * qdda_des + damping * qd_des
        pz_merge<NT, 3, 1, 1>(S, Z.c.u[i], view_extract(Z.h.Nv, row), view_scaled(Z.c.qdda[i], rm.armature[i]), false);
        // + damping * qd_des: with a damping of exactly zero (the Kinova model) the addend is all zeros — centre, coefficients and
        // |scale| * radius — so u is unchanged bit for bit (its monomials are re-thresholded against the same values) and the pass is skipped
        if (rm.damping[i] != 0.0) pz_merge<NT, 1, 1, 1>(S, Z.c.u[i], view(Z.c.u[i]), view_scaled(Z.c.qd[i], rm.damping[i]), false);
    }
}

// ---- the same program cut for two thread groups ------------------------------------------------------------
// Group 0 owns the angular recurrence (w, w_aux, wdot), the moment recursion n and — in the time it would spend
// waiting for group 1 — the forward kinematics; group 1 owns linear_acc, the link forces F / N and the force recursion f.  Operands and operation order of every PZ
// operation are exactly those of chain_joint / force_joint / backward_joint.
template <int NT>
__device__ void angular_joint(Scratch& S, Slots& Z, PZ<3>* T, int i, int p, int c) {
    const RobotModel& rm = c_robot;
    const double zero3[3] = {0, 0, 0};
    const int axis = rm.axes[i];
    const int row = (axis < 0 ? -axis : axis) - 1;
    pz_mul<NT, 9, 3, 3>(S, Z.h.W[c], Z.c.Rt[i], Z.h.W[p]);
    if (axis != 0) pz_add_one_dim<NT>(S, Z.h.W[c], Z.h.W[c], Z.c.qd[i], row);
    pz_mul<NT, 9, 3, 3>(S, Z.h.WA[c], Z.c.Rt[i], Z.h.WA[p]);
    pz_mul<NT, 9, 3, 3>(S, Z.h.WD[c], Z.c.Rt[i], Z.h.WD[p]);
    if (axis != 0) {
        pz_set_const<NT, 3>(T[0], zero3);
        pz_add_one_dim<NT>(S, T[0], T[0], Z.c.qd[i], row);
        pz_cross_pp<NT>(S, T[1], Z.h.WA[c], T[0]);
        pz_add3<NT>(S, Z.h.WD[c], Z.h.WD[c], T[1]);
        pz_add_one_dim<NT>(S, Z.h.WD[c], Z.h.WD[c], Z.c.qdda[i], row);
        pz_add_one_dim<NT>(S, Z.h.WA[c], Z.h.WA[c], Z.c.qda[i], row);
    }
}
// group 1: linear_acc = R_t * (linear_acc + cross(wdot, trans) + cross(w, cross(w_aux, trans))) from the PREVIOUS
// angular state (set p)   (KPR/Dynamics.cu:110-112, line 16)
template <int NT>
__device__ void linacc_joint(Scratch& S, Slots& Z, PZ<3>* T, int i, int p, PZ<3>& LA) {
    const RobotModel& rm = c_robot;
    pz_cross_const<NT>(S, T[0], Z.h.WD[p], rm.trans[i], false);
    pz_cross_const<NT>(S, T[1], Z.h.WA[p], rm.trans[i], false);
    pz_cross_pp<NT>(S, T[2], Z.h.W[p], T[1]);
    pz_add3<NT>(S, T[3], LA, T[0]);
    pz_add3<NT>(S, T[3], T[3], T[2]);
    pz_mul<NT, 9, 3, 3>(S, LA, Z.c.Rt[i], T[3]);
}
// group 1, reverse recursion of f and the two cross terms of n (KPR/Dynamics.cu:163-169)
template <int NT>
__device__ void side_joint(Scratch& S, Slots& Z, PZ<3>* T, int i) {
    const RobotModel& rm = c_robot;
    pz_mul<NT, 9, 3, 3>(S, T[3], Z.c.R[i + 1], Z.h.Fv);                       // R * f
    pz_cross_const<NT>(S, Z.c.C2[i], T[3], rm.trans[i + 1], true);         // cross(trans, R * f)
    pz_cross_const<NT>(S, Z.c.C1[i], Z.c.F[i], rm.com[i], true);             // cross(com, F)
    pz_add3<NT>(S, Z.h.Fv, T[3], Z.c.F[i]);                                   // f = R * f + F
}
// group 0, reverse recursion of n and the torque (KPR/Dynamics.cu:163-179)
template <int NT>
__device__ void moment_joint(Scratch& S, Slots& Z, PZ<3>* T, int i) {
    const RobotModel& rm = c_robot;
    const int axis = rm.axes[i];
    const int row = (axis < 0 ? -axis : axis) - 1;
    pz_mul<NT, 9, 3, 3>(S, T[0], Z.c.R[i + 1], Z.h.Nv);
    pz_add3<NT>(S, T[1], Z.c.N[i], T[0]);
    pz_add3<NT>(S, T[1], T[1], Z.c.C1[i]);
    pz_add3<NT>(S, Z.h.Nv, T[1], Z.c.C2[i]);
    if (axis != 0) {
        pz_merge<NT, 3, 1, 1>(S, Z.c.u[i], view_extract(Z.h.Nv, row), view_scaled(Z.c.qdda[i], rm.armature[i]), false);
        // + damping * qd_des: with a damping of exactly zero (the Kinova model) the addend is all zeros — centre, coefficients and
        // |scale| * radius — so u is unchanged bit for bit (its monomials are re-thresholded against the same values) and the pass is skipped
        if (rm.damping[i] != 0.0) pz_merge<NT, 1, 1, 1>(S, Z.c.u[i], view(Z.c.u[i]), view_scaled(Z.c.qd[i], rm.damping[i]), false);
    }
}

// GROUPS == 1: one group of NT threads runs the whole program in the reference's order.
// GROUPS == 2: see angular_joint / linacc_joint / side_joint / moment_joint above: the two groups run side by side and
// hand PZs over through shared-memory counters; the critical path is roughly half of the 280 operations.
// COLD_GLOBAL: the per-joint descriptors live in the CTA's arena slice instead of shared memory (narrow sweep shapes).
template <int NT, int MINB, int GROUPS, bool COLD_GLOBAL>
__global__ void __launch_bounds__(NT * GROUPS, MINB) reach_build_kernel(Tables tb, char* arena, size_t arena_stride, int mcap, int ncap, int scap, int tcap, int n_work) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Scratch SS[GROUPS];
    __shared__ HotSlots ZH;
    __shared__ PZ<3> TT[GROUPS][5];                        // temporaries, one set per thread group
    __shared__ volatile int sig_la, sig_state, sig_force, sig_side, sig_u, fk_next_s;
    __shared__ int fk_lock_s;
    const RobotModel& rm = c_robot;
    const int group = threadIdx.x / NT;
    Scratch& S = SS[group];
    PZ<3>* T = TT[group];
    constexpr size_t COLD_SMEM = COLD_GLOBAL ? 0 : (sizeof(ColdSlots) + 15) & ~(size_t)15;
    ColdSlots& ZC = COLD_GLOBAL ? *reinterpret_cast<ColdSlots*>(arena + (size_t)blockIdx.x * arena_stride)
                                : *reinterpret_cast<ColdSlots*>(smem_raw);
    Slots Z{ZH, ZC};
    if (threadIdx.x == 0) {
#ifdef ARMOUR_ARENA_CANARY
        g_ncanary = 0;
#endif
        char* g = arena + (size_t)blockIdx.x * arena_stride + cold_arena_bytes();
        for (int k = 0; k < 2; k++) { g = carve<3>(Z.h.W[k], g, mcap); g = carve<3>(Z.h.WD[k], g, mcap); g = carve<3>(Z.h.WA[k], g, mcap); g = carve<3>(Z.h.LA[k], g, mcap); }
        for (int k = 0; k < GROUPS; k++) for (int t = 0; t < 5; t++) g = carve<3>(TT[k][t], g, mcap);
        for (int i = 0; i < NJ; i++) { g = carve<3>(Z.c.C1[i], g, mcap); g = carve<3>(Z.c.C2[i], g, mcap); }
        g = carve<3>(Z.h.Fv, g, mcap); g = carve<3>(Z.h.Nv, g, mcap); g = carve<3>(Z.h.FKT, g, mcap);
        for (int i = 0; i < NJ; i++) { g = carve<3>(Z.c.F[i], g, mcap); g = carve<3>(Z.c.N[i], g, mcap); g = carve<3>(Z.c.LINK[i], g, mcap); g = carve<3>(Z.c.link0[i], g, SMALL_CAP); }
        g = carve<9>(Z.h.FKR, g, mcap);
        for (int i = 0; i <= NJ; i++) g = carve<9>(Z.c.R[i], g, SMALL_CAP);
        for (int i = 0; i < NJ; i++) g = carve<9>(Z.c.Rt[i], g, SMALL_CAP);
        #pragma unroll 1
        for (int i = 0; i < NJ; i++) {
            g = carve<1>(Z.c.qd[i], g, SMALL_CAP); g = carve<1>(Z.c.qda[i], g, SMALL_CAP); g = carve<1>(Z.c.qdda[i], g, SMALL_CAP);
            g = carve<1>(Z.c.cosq[i], g, SMALL_CAP); g = carve<1>(Z.c.sinq[i], g, SMALL_CAP);
            g = carve<1>(Z.c.u[i], g, mcap);
        }
        const double thr_sq = squared_threshold(tb.thr);
        const size_t group_smem = Scratch::smem_bytes(scap, tcap, NT / 32);
        for (int k = 0; k < GROUPS; k++) {
            SS[k].bind(smem_raw + COLD_SMEM + (size_t)k * group_smem, scap, tcap, g + (size_t)k * Scratch::gmem_bytes(ncap), ncap);
            SS[k].thr = tb.thr; SS[k].thr_sq = thr_sq; SS[k].gerr = tb.err; SS[k].no_structured = tb.no_structured;
        }
    }
    __syncthreads();
    const double zero3[3] = {0, 0, 0};
#ifdef ARMOUR_PHASE_TIMING
    if (threadIdx.x == 0) g_phase_clock.last = clock64();
#endif
    // persistent CTAs: the first gridDim.x work items are taken by block index, the rest from a global counter (interval
    // cost varies about 2x along the trajectory and between problems, so a static stride leaves CTAs idle at the end)
    __shared__ int next_work;
    for (int work = blockIdx.x; work < n_work; work = next_work) {
        const int prob = work / tb.T, s = work - prob * tb.T;
        const size_t rec0 = ((size_t)prob * tb.T + s) * NJ;
        if (threadIdx.x == 0) { sig_la = 0; sig_state = 0; sig_force = 0; sig_side = 0; sig_u = 0; fk_next_s = 0; fk_lock_s = 0; }
#ifdef ARMOUR_KEYHASH
        if (threadIdx.x == 0) for (int k = 0; k < GROUPS; k++) { SS[k].dbg_work = work; SS[k].dbg_seq = 0; }
#endif
        __syncthreads();
        // ---- stage A: joint reach sets (one thread per joint; tiny scalar work) -------------------
        if (threadIdx.x < NJ) {
            const int i = threadIdx.x;
            if (tb.mode == 1) {
                make_poly_zono_armtd(tb, prob, s, i, Z.c.R[i], Z.c.Rt[i], Z.c.cosq[i], Z.c.sinq[i], S.thr_sq);
                PZ<1>* unused[3] = {&Z.c.qd[i], &Z.c.qda[i], &Z.c.qdda[i]};
                for (PZ<1>* z : unused) { z->n = 0; z->divM = FastDiv::magic(0); z->ormask = 0; z->center[0] = 0; z->ind[0][0] = 0; z->ind[1][0] = 0; z->abss[0] = 0; }
            }
            else make_poly_zono_joint(tb, prob, s, i, Z.c.R[i], Z.c.Rt[i], Z.c.qd[i], Z.c.qda[i], Z.c.qdda[i], Z.c.cosq[i], Z.c.sinq[i], S.thr_sq);
            if (tb.traj) {
                SmallRec* rec = tb.traj + (((size_t)prob * tb.T + s) * TRAJ_TABLES) * NJ;
                export_small<1>(rec[TRAJ_COS * NJ + i], Z.c.cosq[i]); export_small<1>(rec[TRAJ_SIN * NJ + i], Z.c.sinq[i]);
                export_small<9>(rec[TRAJ_R * NJ + i], Z.c.R[i]); export_small<9>(rec[TRAJ_RT * NJ + i], Z.c.Rt[i]);
                export_small<1>(rec[TRAJ_QD * NJ + i], Z.c.qd[i]); export_small<1>(rec[TRAJ_QDA * NJ + i], Z.c.qda[i]);
                export_small<1>(rec[TRAJ_QDDA * NJ + i], Z.c.qdda[i]);
            }
            // original link boxes: stack of three scalar PZs, one generator each (KPR/Dynamics.cu:51-66)
            PZ<3>& L0 = Z.c.link0[i];
            int n = 0;
            double ind[3] = {0, 0, 0}, abss[3] = {0, 0, 0};
            const u64 gk[3] = {key_qde(0), key_qdae(0), key_qddae(0)};
            for (int j = 0; j < 3; j++) {
                const double g = rm.link_g[i][j];
                if (norm1(&g) > S.thr_sq) {   // both the scalar ctor's simplify and stack()'s see |g|
                    L0.keys[n] = gk[j];
                    for (int c = 0; c < 3; c++) L0.coef[c * L0.cap + n] = (c == j) ? g : 0.0;
                    abss[j] = fabs(g);
                    n++;
                }
                else ind[j] = fabs(g);
            }
            L0.n = n; L0.divM = FastDiv::magic(n); L0.ormask = gk[0] | gk[1] | gk[2];
            for (int c = 0; c < 3; c++) { L0.center[c] = rm.link_c[i][c]; L0.ind[0][c] = ind[c]; L0.ind[1][c] = ind[c]; L0.abss[c] = abss[c]; }
        }
        if (threadIdx.x == NJ) {   // R(NUM_JOINTS) = identity; initial RNEA state (KPR/Dynamics.cu:87-99) in set 1 (= "before joint 0")
            PZ<9>& R = Z.c.R[NJ];
            R.n = 0; R.divM = FastDiv::magic(0); R.ormask = 0;
            for (int c = 0; c < 9; c++) { R.center[c] = rm.R0[NJ][c]; R.ind[0][c] = 0; R.ind[1][c] = 0; R.abss[c] = 0; }
            PZ<3>* init[6] = {&Z.h.W[1], &Z.h.WD[1], &Z.h.WA[1], &Z.h.LA[1], &Z.h.Fv, &Z.h.Nv};
            for (PZ<3>* z : init) { z->n = 0; z->divM = FastDiv::magic(0); z->ormask = 0; for (int c = 0; c < 3; c++) { z->center[c] = 0; z->ind[0][c] = 0; z->ind[1][c] = 0; z->abss[c] = 0; } }
            Z.h.LA[1].center[2] = rm.gravity;
        }
        __syncthreads();
        phase_mark(PH_STAGE_A);

#ifdef ARMOUR_PHASE_TIMING
#define PIECE_T0() long long pt0__ = clock64()
#define PIECE(acc) do { const long long t__ = clock64(); acc += t__ - pt0__; pt0__ = t__; } while (0)
        long long pc_chain = 0, pc_force = 0, pc_fk = 0, pc_back = 0, pc_wait = 0;
#else
#define PIECE_T0()
#define PIECE(acc)
#endif
        PIECE_T0();
        if (tb.mode == 1) {   // ARMTD comparison planner: forward kinematics only (KPA/armtd_main.cu:127-145)
            if (group == 0) forward_kinematics<NT>(S, Z, T, tb, rec0);
        }
        else if (GROUPS == 1) {
            forward_kinematics<NT>(S, Z, T, tb, rec0);                                    // stage B
            #pragma unroll 1
            for (int i = 0; i < NJ; i++) {                                                 // stage C forward
                chain_joint<NT>(S, Z, T, i, (i + 1) & 1, i & 1);
                force_joint<NT>(S, Z, T, tb, i, i & 1, Z.h.LA[i & 1]);
            }
            #pragma unroll 1
            for (int i = NJ - 1; i >= 0; i--) backward_joint<NT>(S, Z, T, i);             // stage C backward
        }
        else if (group == 0) {
            // angular recurrence, then the moment recursion; forward-kinematics joints fill the time spent waiting
#if ARMOUR_SHARED_FK
            #pragma unroll 1
            for (int i = 0; i < NJ; i++) {
                // set i&1 still holds the state of joint i-2, read by group 1 until its linear_acc step of joint i-1 is done
                if (i >= 1) wait_or_fk<NT>(S, Z, T, tb, rec0, &sig_la, i, &fk_next_s, &fk_lock_s);
                PIECE(pc_wait);
                angular_joint<NT>(S, Z, T, i, (i + 1) & 1, i & 1);
                group_signal<NT>(&sig_state, i + 1);
                PIECE(pc_chain);
            }
            #pragma unroll 1
            for (int i = NJ - 1; i >= 0; i--) {
                wait_or_fk<NT>(S, Z, T, tb, rec0, &sig_side, NJ - i, &fk_next_s, &fk_lock_s);
                PIECE(pc_wait);
                moment_joint<NT>(S, Z, T, i);
                group_signal<NT>(&sig_u, NJ - i);          // u_i is final: group 1 exports it (stage M) beside the next joint
                PIECE(pc_back);
            }
            drain_fk<NT>(S, Z, T, tb, rec0, &fk_next_s, &fk_lock_s);
            PIECE(pc_fk);
#else
            int fk_next = 0;
            #pragma unroll 1
            for (int i = 0; i < NJ; i++) {
                // set i&1 still holds the state of joint i-2, read by group 1 until its linear_acc step of joint i-1 is done
                if (i >= 1) {
                    while (fk_next < NJ && !group_ready<NT>(S, &sig_la, i)) { fk_joint<NT>(S, Z, T, tb, rec0, fk_next++); PIECE(pc_fk); }
                    group_wait<NT>(S, &sig_la, i);
                }
                PIECE(pc_wait);
                angular_joint<NT>(S, Z, T, i, (i + 1) & 1, i & 1);
                group_signal<NT>(&sig_state, i + 1);
                PIECE(pc_chain);
            }
            #pragma unroll 1
            for (int i = NJ - 1; i >= 0; i--) {
                while (fk_next < NJ && !group_ready<NT>(S, &sig_side, NJ - i)) { fk_joint<NT>(S, Z, T, tb, rec0, fk_next++); PIECE(pc_fk); }
                group_wait<NT>(S, &sig_side, NJ - i);
                PIECE(pc_wait);
                moment_joint<NT>(S, Z, T, i);
                group_signal<NT>(&sig_u, NJ - i);          // u_i is final: group 1 exports it (stage M) beside the next joint
                PIECE(pc_back);
            }
            while (fk_next < NJ) { fk_joint<NT>(S, Z, T, tb, rec0, fk_next++); PIECE(pc_fk); }
#endif
        }
        else {
            PZ<3>& LA = Z.h.LA[1];   // group 1's private linear_acc, initialised with gravity in stage A
            #pragma unroll 1
            for (int i = 0; i < NJ; i++) {
#if ARMOUR_SHARED_FK
                wait_or_fk<NT>(S, Z, T, tb, rec0, &sig_state, i, &fk_next_s, &fk_lock_s);          // state of joint i-1 (the initial state for i = 0)
#else
                group_wait<NT>(S, &sig_state, i);          // state of joint i-1 (the initial state for i = 0)
#endif
                PIECE(pc_wait);
                linacc_joint<NT>(S, Z, T, i, (i + 1) & 1, LA);
                group_signal<NT>(&sig_la, i + 1);
                PIECE(pc_chain);
#if ARMOUR_SHARED_FK
                wait_or_fk<NT>(S, Z, T, tb, rec0, &sig_state, i + 1, &fk_next_s, &fk_lock_s);
#else
                group_wait<NT>(S, &sig_state, i + 1);
#endif
                PIECE(pc_wait);
                force_joint<NT>(S, Z, T, tb, i, i & 1, LA);
                group_signal<NT>(&sig_force, i + 1);
                PIECE(pc_force);
            }
            #pragma unroll 1
            for (int i = NJ - 1; i >= 0; i--) {
                side_joint<NT>(S, Z, T, i);
                group_signal<NT>(&sig_side, NJ - i);
                PIECE(pc_back);
            }
            // stage M exports (disturbance radius, reduce(), torque table) in the time this group would idle
            #pragma unroll 1
            for (int i = NJ - 1; i >= 0; i--) {
#if ARMOUR_SHARED_FK
                wait_or_fk<NT>(S, Z, T, tb, rec0, &sig_u, NJ - i, &fk_next_s, &fk_lock_s);
#else
                group_wait<NT>(S, &sig_u, NJ - i);
#endif
                export_torque<NT>(S, tb, rec0 + i, Z.c.u[i]);
            }
#if ARMOUR_SHARED_FK
            drain_fk<NT>(S, Z, T, tb, rec0, &fk_next_s, &fk_lock_s);
#endif
        }
#ifdef ARMOUR_PHASE_TIMING
        if (blockIdx.x == 64 && gtid<NT>() == 0)
            printf("PIECES group %d: chain %lld force %lld fk %lld backward %lld wait %lld (cycles)\n", group, pc_chain, pc_force, pc_fk, pc_back, pc_wait);
#endif

        // ---- stage M: disturbance, reduce(), torque radius (KPR/armour_main.cu:135-205) -------------
        if (GROUPS == 2 && tb.mode == 0) __syncthreads();   // group 1's exports are visible to thread 0
        if (group == 0 && tb.mode == 0) {
            if (GROUPS == 1) for (int i = 0; i < NF; i++) export_torque<NT>(S, tb, rec0 + i, Z.c.u[i]);
            if (threadIdx.x == 0) {
                const size_t base = rec0;
                // rho = sqrt(sum_i [-r_i, r_i]^2): only the upper end is used; everything rounded up
                double rho = 0.0;
                for (int i = 0; i < NF; i++) { const double r = tb.dist_rad[base + i]; rho = __dadd_ru(rho, __dmul_ru(r, r)); }
                rho = __dsqrt_ru(rho);
                const double c0 = __dmul_ru(__dmul_ru(rm.alpha, __dadd_ru(rm.M_max, -rm.M_min)), rm.eps);
                for (int i = 0; i < NF; i++) {
                    double tr = __dadd_ru(c0, __dmul_ru(0.5, tb.dist_rad[base + i]));
                    tr = __dadd_ru(tr, __dmul_ru(0.5, rho));
                    tr = __dadd_ru(tr, tb.u_ind[base + i]);
                    tr = __dadd_ru(tr, rm.friction[i]);
                    tb.torque_radius[base + i] = tr;
                }
            }
        }
        // ---- stage D fused: this interval's half-space tables, from the generator blocks stage B just exported -----------
        if (tb.fuse_planes && tb.n_obs > 0) interval_half_spaces(tb, prob, rec0, reinterpret_cast<double*>(smem_raw + COLD_SMEM));
        if (threadIdx.x == 0) next_work = tb.static_stride ? work + (int)gridDim.x : (int)gridDim.x + atomicAdd(tb.err + 1, 1);
        __syncthreads();
    }
#ifdef ARMOUR_ARENA_CANARY
    if (threadIdx.x == 0) {
        bool ok = g_ncanary > 100 && g_ncanary < CANARY_MAX;
        for (int i = 0; i < g_ncanary && i < CANARY_MAX; i++) ok = ok && (*g_canary[i] == CANARY_WORD);
        if (!ok) atomicOr(tb.err, ERR_CANARY);
        else atomicAdd(tb.err + 2, g_ncanary);   // guard words verified (read back by the canary test)
    }
#endif
    (void)zero3;
}

// ---- stand-alone PZsparse arithmetic (PZsparse facade / primitive parity tests) -----------------------
// Flat operand record in global memory: [n, dim] ints, then keys[cap], coef[dim][cap], center[9], ind[9]
template <int D>
__device__ void load_flat(PZ<D>& z, const FlatPZ& f) {
    if (threadIdx.x == 0) {
        z.n = f.n; z.divM = FastDiv::magic(f.n); z.cap = f.cap; z.keys = f.keys; z.coef = f.coef;
        z.ormask = 0;
        for (int i = 0; i < f.n; i++) z.ormask |= f.keys[i];
        for (int c = 0; c < D; c++) {
            z.center[c] = f.center[c]; z.ind[0][c] = f.ind[c]; z.ind[1][c] = f.ind[c];
            double s = 0.0;
            for (int i = 0; i < f.n; i++) s = __dadd_ru(s, fabs(f.coef[c * f.cap + i]));
            z.abss[c] = s;
        }
    }
}
template <int D>
__device__ void store_flat(FlatPZ& f, const PZ<D>& z) {
    if (threadIdx.x == 0) { f.n = z.n; f.dim = D; for (int c = 0; c < D; c++) { f.center[c] = z.center[c]; f.ind[c] = z.ind[0][c]; } }
}
template <int NT>
__global__ void __launch_bounds__(NT) pz_binary_kernel(int op, FlatPZ a, FlatPZ b, FlatPZ r, FlatOut* out, char* gmem, int ncap, int scap, int tcap, double thr, int* err) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Scratch S;
    __shared__ PZ<1> a1, b1, r1;
    __shared__ PZ<3> a3, b3, r3;
    __shared__ PZ<9> a9, b9, r9;
    if (threadIdx.x == 0) {
        S.bind(smem_raw, scap, tcap, gmem, ncap);
        S.thr = thr; S.thr_sq = squared_threshold(thr); S.gerr = err;
        r1.cap = r3.cap = r9.cap = r.cap; r1.keys = r3.keys = r9.keys = r.keys; r1.coef = r3.coef = r9.coef = r.coef;
    }
    __syncthreads();
    FlatOut o; o.n = -1; o.dim = 0;
    if (op == 0) {
        if (a.dim == 9 && b.dim == 3) { load_flat<9>(a9, a); load_flat<3>(b3, b); __syncthreads(); pz_mul<NT, 9, 3, 3>(S, r3, a9, b3); o.n = r3.n; o.dim = 3; store_flat<3>(r, r3); }
        else if (a.dim == 9 && b.dim == 9) { load_flat<9>(a9, a); load_flat<9>(b9, b); __syncthreads(); pz_mul<NT, 9, 9, 9>(S, r9, a9, b9); o.n = r9.n; o.dim = 9; store_flat<9>(r, r9); }
        else if (a.dim == 1 && b.dim == 1) { load_flat<1>(a1, a); load_flat<1>(b1, b); __syncthreads(); pz_mul<NT, 1, 1, 1>(S, r1, a1, b1); o.n = r1.n; o.dim = 1; store_flat<1>(r, r1); }
    }
    else if (op == 1 || op == 2) {
        if (a.dim == 3 && b.dim == 3) { load_flat<3>(a3, a); load_flat<3>(b3, b); __syncthreads(); pz_merge<NT, 3, 3, 3>(S, r3, view(a3), view(b3), op == 2); o.n = r3.n; o.dim = 3; store_flat<3>(r, r3); }
        else if (a.dim == 1 && b.dim == 1) { load_flat<1>(a1, a); load_flat<1>(b1, b); __syncthreads(); pz_merge<NT, 1, 1, 1>(S, r1, view(a1), view(b1), op == 2); o.n = r1.n; o.dim = 1; store_flat<1>(r, r1); }
    }
    else if (op == 3 && a.dim == 3 && b.dim == 3) { load_flat<3>(a3, a); load_flat<3>(b3, b); __syncthreads(); pz_cross_pp<NT>(S, r3, a3, b3); o.n = r3.n; o.dim = 3; store_flat<3>(r, r3); }
    else if (op == 4) {   // simplify
        if (a.dim == 1) { load_flat<1>(a1, a); __syncthreads(); pz_simplify<NT, 1>(S, r1, a1); o.n = r1.n; o.dim = 1; store_flat<1>(r, r1); }
        else if (a.dim == 3) { load_flat<3>(a3, a); __syncthreads(); pz_simplify<NT, 3>(S, r3, a3); o.n = r3.n; o.dim = 3; store_flat<3>(r, r3); }
        else if (a.dim == 9) { load_flat<9>(a9, a); __syncthreads(); pz_simplify<NT, 9>(S, r9, a9); o.n = r9.n; o.dim = 9; store_flat<9>(r, r9); }
    }
    else if (op >= 7 && op <= 9 && a.dim == 3 && b.dim == 1) {   // addOneDimPZ(b, row, 0)
        load_flat<3>(a3, a); load_flat<1>(b1, b); __syncthreads(); pz_add_one_dim<NT>(S, r3, a3, b1, op - 7); o.n = r3.n; o.dim = 3; store_flat<3>(r, r3);
    }
    else if ((op == 10 || op == 11) && a.dim == 3 && b.dim == 3) {   // cross(constant, a) / cross(a, constant): the constant is b's centre
        load_flat<3>(a3, a); load_flat<3>(b3, b); __syncthreads();
        const double kv[3] = {b3.center[0], b3.center[1], b3.center[2]};
        pz_cross_const<NT>(S, r3, a3, kv, op == 10); o.n = r3.n; o.dim = 3; store_flat<3>(r, r3);
    }
    __syncthreads();
    if (threadIdx.x == 0) { if (o.dim == 1) o.n = r1.n; else if (o.dim == 3) o.n = r3.n; else if (o.dim == 9) o.n = r9.n; *out = o; }
}

// ---- host-side launch helpers -----------------------------------------------------------------
cudaError_t upload_robot_model(const RobotModel& rm) { return cudaMemcpyToSymbol(c_robot, &rm, sizeof(RobotModel)); }

size_t reach_gmem_bytes(int ncap) { return Scratch::gmem_bytes(ncap); }

// kernel variants: (threads per group, CTAs per SM the register allocation is bounded for, thread groups per CTA,
// per-joint descriptors in global memory).  One plan: 2 groups x 256 threads, 1 CTA per SM.  Sweeps: one group of 32 / 64 /
// 128 threads per interval and as many CTAs per SM as shared memory and registers allow.
typedef void (*ReachKernel)(Tables, char*, size_t, int, int, int, int, int);
struct ReachVariant { ReachKernel k; int threads; int nwarps_per_group; int groups; bool cold_global; };
static ReachVariant pick_reach_kernel(int nt, int minb, int groups) {
    if (groups == 2) return {reach_build_kernel<256, 1, 2, false>, 512, 8, 2, false};
    if (nt == 32) {
        if (minb >= 24) return {reach_build_kernel<32, 24, 1, true>, 32, 1, 1, true};
        if (minb >= 16) return {reach_build_kernel<32, 16, 1, true>, 32, 1, 1, true};
        return {reach_build_kernel<32, 12, 1, true>, 32, 1, 1, true};
    }
    if (nt == 64) {
        if (minb >= 12) return {reach_build_kernel<64, 12, 1, true>, 64, 2, 1, true};
        return {reach_build_kernel<64, 8, 1, true>, 64, 2, 1, true};
    }
    if (nt == 128) {
        if (minb >= 6) return {reach_build_kernel<128, 6, 1, true>, 128, 4, 1, true};
        if (minb == 5) return {reach_build_kernel<128, 5, 1, false>, 128, 4, 1, false};
        if (minb == 4) return {reach_build_kernel<128, 4, 1, false>, 128, 4, 1, false};
        if (minb == 3) return {reach_build_kernel<128, 3, 1, false>, 128, 4, 1, false};
        return {reach_build_kernel<128, 2, 1, false>, 128, 4, 1, false};
    }
    if (nt == 512) return {reach_build_kernel<512, 1, 1, false>, 512, 16, 1, false};
    if (minb >= 2) return {reach_build_kernel<256, 2, 1, false>, 256, 8, 1, false};
    return {reach_build_kernel<256, 1, 1, false>, 256, 8, 1, false};
}
static size_t variant_smem(const ReachVariant& v, int scap, int tcap) {
    return (v.cold_global ? 0 : ((sizeof(ColdSlots) + 15) & ~(size_t)15)) + (size_t)v.groups * Scratch::smem_bytes(scap, tcap, v.nwarps_per_group);
}
cudaError_t launch_reach_build(const Tables& tb, char* arena, size_t arena_stride, int mcap, int ncap, int scap, int tcap, int n_work, int grid, int nt, int minb, int groups, cudaStream_t stream) {
    const ReachVariant v = pick_reach_kernel(nt, minb, groups);
    const size_t smem = variant_smem(v, scap, tcap);
    cudaError_t e = cudaFuncSetAttribute(v.k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    v.k<<<grid, v.threads, smem, stream>>>(tb, arena, arena_stride, mcap, ncap, scap, tcap, n_work);
    return cudaGetLastError();
}

cudaError_t launch_pz_binary(int op, const FlatPZ& a, const FlatPZ& b, const FlatPZ& r, FlatOut* out, char* gmem, int ncap, int scap, int tcap, double thr, int* err, cudaStream_t stream) {
    const size_t smem = Scratch::smem_bytes(scap, tcap, 8);
    cudaError_t e = cudaFuncSetAttribute(pz_binary_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pz_binary_kernel<256><<<1, 256, smem, stream>>>(op, a, b, r, out, gmem, ncap, scap, tcap, thr, err);
    return cudaGetLastError();
}

// profiling builds only: cycles and call counts per phase, summed over CTAs (zeros otherwise)
#ifdef ARMOUR_KEYHASH
extern "C" int armour_debug_keyhash(unsigned long long* out, int work_items) {   // out[work][KEYHASH_OPS][2]
    if (work_items > KEYHASH_WORK) work_items = KEYHASH_WORK;
    return (int)cudaMemcpyFromSymbol(out, g_keyhash, sizeof(u64) * (size_t)work_items * KEYHASH_OPS * 2);
}
#endif
#ifdef ARMOUR_SEGSTAT
extern "C" int armour_debug_segstat(unsigned long long* out4, int reset) {
    int rc = (int)cudaMemcpyFromSymbol(out4, g_segstat, sizeof(unsigned long long) * 4);
    if (reset) { unsigned long long z[4] = {0, 0, 0, 0}; cudaMemcpyToSymbol(g_segstat, z, sizeof(z)); }
    return rc;
}
#endif
void read_phase_cycles(unsigned long long* cycles, unsigned long long* calls, bool reset) {
#ifdef ARMOUR_PHASE_TIMING
    cudaMemcpyFromSymbol(cycles, armour_phase_cycles, sizeof(unsigned long long) * PH_COUNT);
    cudaMemcpyFromSymbol(calls, armour_phase_calls, sizeof(unsigned long long) * PH_COUNT);
    if (reset) {
        unsigned long long z[PH_COUNT] = {0};
        cudaMemcpyToSymbol(armour_phase_cycles, z, sizeof(z));
        cudaMemcpyToSymbol(armour_phase_calls, z, sizeof(z));
    }
#else
    for (int i = 0; i < PH_COUNT; i++) { cycles[i] = 0; calls[i] = 0; }
    (void)reset;
#endif
}

// resident CTAs per SM for a variant; 0 when it does not fit (e.g. two groups with large sort buffers)
int reach_max_ctas_per_sm(int nt, int minb, int groups, int scap, int tcap) {
    int n = 0;
    const ReachVariant v = pick_reach_kernel(nt, minb, groups);
    const size_t smem = variant_smem(v, scap, tcap);
    if (cudaFuncSetAttribute(v.k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, v.k, v.threads, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

}  // namespace armour

