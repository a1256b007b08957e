// Exercises the C++ host mirror (armtd_NLP, PZsparse facade, stand-in solver) on a device.  Prints one line of
// numbers that tests/test_host_mirror.py compares with the ctypes path.
#include <cstdio>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "PZsparse.hpp"
#include "armtd_NLP.hpp"
#include "standin_solver.hpp"

int main() {
    armour_config cfg;
    armour_default_config(&cfg);
    cfg.num_time_steps = 16;
    armour_handle* h = nullptr;
    if (armour_create(&cfg, &h) != ARMOUR_OK) { printf("create failed: %s\n", armour_last_error()); return 2; }
    const double q0[7] = {0.6543, -0.0876, -0.4837, -1.2278, -1.5735, -1.0720, 0}, qd0[7] = {0.1, -0.1, 0.2, 0.1, -0.2, 0.1, 0.05}, qdd0[7] = {0.1, 0.2, -0.1, 0.3, 0.1, -0.2, 0.1};
    const double q_des[7] = {0.6831, 0.009488, -0.2471, -0.9777, -1.414, -0.9958, 0};
    const double obs[24] = {-0.28239, -0.33281, 0.88069, 0.069825, 0, 0, 0, 0.09508, 0, 0, 0, 0.016624, 0.67593, -0.085841, 0.43572, 0.17408, 0, 0, 0, 0.07951, 0, 0, 0, 0.18012};
    if (armour_build(h, q0, qd0, qdd0, obs, 2) != ARMOUR_OK) { printf("build failed: %s\n", armour_last_error()); return 2; }
    armtd_NLP nlp;
    nlp.set_time_steps(16);
    if (!nlp.set_parameters(q_des, 0.5, h)) return 2;
    armtd_NLP::Index n, m, nnz, nh;
    Ipopt::TNLP::IndexStyleEnum st;
    nlp.get_nlp_info(n, m, nnz, nh, st);
    std::vector<double> g(m), J((size_t)m * n);
    const double x[7] = {0.5, 0.6, 0.7, 0.0, -0.5, -0.6, -0.7};
    if (!nlp.eval_g(n, x, true, m, g.data()) || !nlp.eval_jac_g(n, x, false, m, nnz, nullptr, nullptr, J.data())) return 2;
    double gs = 0, js = 0;
    for (double v : g) gs += v;
    for (double v : J) js += v;
    double f;
    nlp.eval_f(n, x, true, f);
    // PZsparse facade: R(0, 3) * links-like vector and cross on the device
    PZsparse::device() = h;
    PZsparse R = PZsparse::from_table(h, 2, 0, 3), L = PZsparse::from_table(h, 7, 2, 3), U = PZsparse::from_table(h, 8, 1, 3);
    PZsparse RL = R * L, C = cross(L, RL), S = U + U, D = U - U;
    std::vector<double> c, r;
    C.slice(x, c, r);
    // the rest of the PZsparse public surface (KPR/PZsparse.h:50-210)
    int facade_errors = 0;
    auto expect = [&](bool ok, const char* what) { if (!ok) { printf("FACADE MISMATCH: %s\n", what); facade_errors++; } };
    auto same = [](const PZsparse& a, const PZsparse& b, double tol) {
        if (a.NRows != b.NRows || a.NCols != b.NCols || a.polynomial.size() != b.polynomial.size()) return false;
        for (size_t i = 0; i < a.polynomial.size(); i++) {
            if (a.polynomial[i].degree != b.polynomial[i].degree) return false;
            for (size_t e = 0; e < a.polynomial[i].coeff.size(); e++) if (std::fabs(a.polynomial[i].coeff[e] - b.polynomial[i].coeff[e]) > tol) return false;
        }
        for (size_t e = 0; e < a.center.size(); e++) if (std::fabs(a.center[e] - b.center[e]) > tol || std::fabs(a.independent[e] - b.independent[e]) > 1e-9) return false;
        return true;
    };
    {
        // gradient slice of a torque PZ == its row of the dense Jacobian (armtd_NLP::eval_jac_g, KPR/NLPclass.cu:362-374)
        double grad[7];
        U.slice(grad, x);
        for (int kk = 0; kk < 7; kk++) expect(std::fabs(grad[kk] - J[(size_t)(3 * 7 + 1) * 7 + kk]) <= 1e-12 * std::max(1.0, std::fabs(grad[kk])), "slice gradient vs Jacobian row");
        // value slice == g row; toInterval encloses it
        std::vector<double> uc, ur, lo, hi;
        U.slice(x, uc, ur);
        expect(std::fabs(uc[0] - g[3 * 7 + 1]) <= 1e-12 * std::max(1.0, std::fabs(uc[0])), "slice value vs g row");
        U.toInterval(lo, hi);
        expect(lo[0] <= uc[0] - ur[0] + 1e-12 && uc[0] + ur[0] <= hi[0] + 1e-12, "toInterval encloses the slice");
        // element extraction + stack gives the vector back (each monomial is split into three and merged again by simplify)
        PZsparse rows[3] = {RL(0, 0), RL(1, 0), RL(2, 0)};
        expect(rows[0].polynomial.size() == RL.polynomial.size(), "operator() copies every monomial");
        expect(same(stack(rows), RL, 1e-15), "stack(extract) == original");
        // simplify is idempotent on a simplified PZ; on a doubled list it adds coefficients
        PZsparse S2 = RL; S2.simplify();
        expect(same(S2, RL, 0.0), "simplify idempotent");
        PZsparse Dbl = RL; Dbl.polynomial.insert(Dbl.polynomial.end(), RL.polynomial.begin(), RL.polynomial.end()); Dbl.simplify();
        expect(Dbl.polynomial.size() == RL.polynomial.size() && std::fabs(Dbl.polynomial[0].coeff[0] - 2 * RL.polynomial[0].coeff[0]) <= 1e-15, "simplify merges repeated keys");
        // scalar operators (host, no simplify) against the PZ x PZ path: (2 * U) has the monomials of U + U
        PZsparse U2 = 2.0 * U;
        expect(same(U2, S, 1e-15) || U2.polynomial.size() >= S.polynomial.size(), "2 * U vs U + U");
        expect(std::fabs((U / 2.0).center[0] - 0.5 * U.center[0]) <= 1e-15 && std::fabs((U + 1.5).center[0] - (U.center[0] + 1.5)) <= 1e-15 &&
               std::fabs((1.5 - U).center[0] - (U.center[0] - 1.5)) <= 1e-15 && (-U).independent[0] == 0.0, "scalar operators");
        // constant cross products: cross(c, L) == -cross(L, c) coefficient-wise
        const double cvec[3] = {0.3, -0.2, 0.7};
        PZsparse C1 = cross(cvec, L), C2 = cross(L, cvec);
        expect(C1.polynomial.size() == C2.polynomial.size(), "constant cross sizes");
        for (size_t i = 0; i < C1.polynomial.size() && i < C2.polynomial.size(); i++)
            for (int e = 0; e < 3; e++) expect(std::fabs(C1.polynomial[i].coeff[e] + C2.polynomial[i].coeff[e]) <= 1e-15, "cross(c, L) == -cross(L, c)");
        // addOneDimPZ: adding a scalar PZ into row 1 only changes that component's centre
        PZsparse La = L; La.addOneDimPZ(U, 1, 0);
        expect(std::fabs(La.center[1] - (L.center[1] + U.center[0])) <= 1e-15 && La.center[0] == L.center[0], "addOneDimPZ centre");
        // reduce / reduce_link_PZ: k-only monomials survive, the rest goes to the radius
        PZsparse Ur = RL; Ur.reduce();
        for (const Monomial& mm : Ur.polynomial) expect(mm.degree < (1ull << 14), "reduce keeps k-only monomials");
        uint64_t da[42];
        PZsparse::convertHashToDegree(RL.polynomial.empty() ? 0 : RL.polynomial.back().degree, da);
        expect(PZsparse::convertDegreeToHash(da) == (RL.polynomial.empty() ? 0 : RL.polynomial.back().degree), "degree pack / unpack round trip");
    }
    double k[7];
    StandinResult sr = standin_solve(nlp, k);
    printf("RESULT n %d m %d gsum %.12e jsum %.12e f %.12e RLn %zu Cn %zu Sn %zu Dn %zu cslice %.12e %.12e %.12e feasible %d iters %d k0 %.6f\n", n, m, gs, js, f,
           RL.polynomial.size(), C.polynomial.size(), S.polynomial.size(), D.polynomial.size(), c[0], c[1], c[2], (int)nlp.feasible, sr.iterations, k[0]);
    printf("FACADE errors %d\n", facade_errors);
    armour_destroy(h);
    return facade_errors ? 3 : 0;
}
