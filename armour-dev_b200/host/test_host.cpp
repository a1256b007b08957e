// Exercises the C++ host mirror (armtd_NLP, PZsparse facade, stand-in solver) on a device.  Prints one line of
// numbers that tests/test_host_mirror.py compares with the ctypes path.
#include <cstdio>
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
    double k[7];
    StandinResult sr = standin_solve(nlp, k);
    printf("RESULT n %d m %d gsum %.12e jsum %.12e f %.12e RLn %zu Cn %zu Sn %zu Dn %zu cslice %.12e %.12e %.12e feasible %d iters %d k0 %.6f\n", n, m, gs, js, f,
           RL.polynomial.size(), C.polynomial.size(), S.polynomial.size(), D.polynomial.size(), c[0], c[1], c[2], (int)nlp.feasible, sr.iterations, k[0]);
    armour_destroy(h);
    return 0;
}
