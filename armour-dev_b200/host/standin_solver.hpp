// Stand-in NLP solver used ONLY when Ipopt is not available (this image): quadratic-penalty Gauss-Newton with
// box projection, 7 variables.  It drives the same TNLP callbacks Ipopt would (eval_f, eval_grad_f, eval_g,
// eval_jac_g) so the whole device path is exercised end to end, but it is NOT Ipopt: its k differs from the
// reference's solver and it is labelled as a stand-in wherever its result is reported.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "armtd_NLP.hpp"

struct StandinResult { int iterations = 0; int evaluations = 0; bool converged = false; double objective = 0; double max_violation = 0; };

// NLP: anything with the Ipopt::TNLP callbacks (armtd_NLP directly, or an Ipopt::TNLP& through its virtual interface)
// work: optional caller-owned scratch of m * (n + 2) doubles for the constraint values (current and trial point) and the dense
// Jacobian — the arrays the callbacks write.  armour_standin_solve passes page-locked memory of the handle here, so that every
// callback is a direct device write without registering solver temporaries.
template <class NLP>
inline StandinResult standin_solve(NLP& nlp, double* x_out, int max_iter = 60, double* work = nullptr) {
    typedef Ipopt::Index Index;
    Index n, m, nnz, nh;
    Ipopt::TNLP::IndexStyleEnum st;
    nlp.get_nlp_info(n, m, nnz, nh, st);
    std::vector<double> xl(n), xu(n), gl(m), gu(m), x(n), grad(n), own;
    if (!work) { own.resize((size_t)m * (n + 2)); work = own.data(); }
    double *g = work, *gt = work + m, *J = work + 2 * (size_t)m;
    nlp.get_bounds_info(n, xl.data(), xu.data(), m, gl.data(), gu.data());
    nlp.get_starting_point(n, true, x.data(), false, nullptr, nullptr, m, false, nullptr);
    StandinResult res;
    double mu = 1e2;
    auto merit = [&](const std::vector<double>& xx, double& f, double& viol, double* gg) {
        nlp.eval_f(n, xx.data(), true, f);
        nlp.eval_g(n, xx.data(), true, m, gg);
        res.evaluations++;
        double p = 0; viol = 0;
        for (Index i = 0; i < m; i++) {
            const double v = gg[i] < gl[i] ? gl[i] - gg[i] : gg[i] > gu[i] ? gg[i] - gu[i] : 0.0;
            p += v * v; viol = std::max(viol, v);
        }
        return f + 0.5 * mu * p;
    };
    double f, viol;
    double phi = merit(x, f, viol, g);
    for (int it = 0; it < max_iter; it++) {
        res.iterations = it + 1;
        nlp.eval_grad_f(n, x.data(), false, grad.data());
        nlp.eval_jac_g(n, x.data(), false, m, nnz, nullptr, nullptr, J);
        // Gauss-Newton system (H + mu J_a^T J_a) dx = -(grad + mu J_a^T v)
        double H[7][7] = {{0}}, rhs[7] = {0};
        for (Index a = 0; a < n; a++) { H[a][a] = 1e-6 + 20.0; rhs[a] = -grad[a]; }   // objective curvature scale (10 * 2 * dq/dk^2 <= 20)
        for (Index i = 0; i < m; i++) {
            const double v = g[i] < gl[i] ? g[i] - gl[i] : g[i] > gu[i] ? g[i] - gu[i] : 0.0;
            if (v == 0.0) continue;
            const double* Ji = &J[(size_t)i * n];
            for (Index a = 0; a < n; a++) { rhs[a] -= mu * Ji[a] * v; for (Index b = 0; b < n; b++) H[a][b] += mu * Ji[a] * Ji[b]; }
        }
        // Cholesky solve (7x7)
        double L[7][7] = {{0}}, y[7], dx[7];
        bool ok = true;
        for (int a = 0; a < n && ok; a++)
            for (int b = 0; b <= a; b++) {
                double s = H[a][b];
                for (int k = 0; k < b; k++) s -= L[a][k] * L[b][k];
                if (a == b) { if (s <= 0) { ok = false; break; } L[a][a] = std::sqrt(s); } else L[a][b] = s / L[b][b];
            }
        if (!ok) break;
        for (int a = 0; a < n; a++) { double s = rhs[a]; for (int k = 0; k < a; k++) s -= L[a][k] * y[k]; y[a] = s / L[a][a]; }
        for (int a = n - 1; a >= 0; a--) { double s = y[a]; for (int k = a + 1; k < n; k++) s -= L[k][a] * dx[k]; dx[a] = s / L[a][a]; }
        double step = 1.0, best = phi;
        std::vector<double> xt(n);
        bool moved = false;
        for (int ls = 0; ls < 12; ls++, step *= 0.5) {
            for (Index a = 0; a < n; a++) xt[a] = std::min(xu[a], std::max(xl[a], x[a] + step * dx[a]));
            double ft, vt;
            const double pt = merit(xt, ft, vt, gt);
            if (pt < best - 1e-12) { x = xt; std::swap(g, gt); f = ft; viol = vt; phi = pt; moved = true; break; }
        }
        double nrm = 0;
        for (Index a = 0; a < n; a++) nrm = std::max(nrm, std::fabs(step * dx[a]));
        if (!moved || nrm < 1e-7) {
            if (viol > 1e-6 && mu < 1e9) { mu *= 10; phi = merit(x, f, viol, g); continue; }
            res.converged = true;
            break;
        }
    }
    res.objective = f; res.max_violation = viol;
    for (Index a = 0; a < n; a++) x_out[a] = x[a];
    std::vector<double> lam;
    nlp.finalize_solution(res.converged ? Ipopt::SUCCESS : Ipopt::MAXITER_EXCEEDED, n, x.data(), nullptr, nullptr, m, g, nullptr, f, nullptr, nullptr);
    return res;
}
