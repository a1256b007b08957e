// armour_main — drop-in for the reference executable the MATLAB driver shells out to
// (uarmtd_planner.m:200; KPR/armour_main.cu).  Same text protocol:
//   in : <buffer>/armour.in  = 7 q0, 7 qd0, 7 qdd0, 7 q_des, n_obs, n_obs*12 doubles      (KPR/armour_main.cu:47-79)
//   out: armour.out (7 k or -1, then total ms), armour_joint_position_center.out, armour_joint_position_radius.out,
//        armour_control_input_radius.out, armour_constraints.out                              (KPR/armour_main.cu:324-397)
// Exit code 0 also for "infeasible" (-1 in armour.out); non-zero on I/O or device errors, like the reference.
// The buffer directory is the reference's generated BufferPath.h constant; here: argv[1], else $ARMOUR_BUFFER_PATH,
// else ./buffer/ .  Solver: Ipopt when available at build time (-DARMOUR_HAVE_IPOPT via armtd_NLP.hpp), otherwise the
// stand-in of standin_solver.hpp (clearly labelled on stdout).
#include <chrono>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <algorithm>
#include <string>
#include <vector>

#include "armtd_NLP.hpp"
#include "standin_solver.hpp"
#ifdef ARMOUR_HAVE_IPOPT
#include "IpIpoptApplication.hpp"
#endif

int main(int argc, char** argv) {
    std::string pathname = argc > 1 ? argv[1] : (getenv("ARMOUR_BUFFER_PATH") ? getenv("ARMOUR_BUFFER_PATH") : "./buffer/");
    if (!pathname.empty() && pathname.back() != '/') pathname += '/';
    const int NUM_TIME_STEPS = getenv("ARMOUR_NUM_TIME_STEPS") ? atoi(getenv("ARMOUR_NUM_TIME_STEPS")) : 128;
    const int MAX_OBSTACLE_NUM = 40;
    // declared first so that a new output always exists (KPR/armour_main.cu:36-37)
    std::ofstream out1(pathname + "armour.out");
    double q0[7] = {0}, qd0[7] = {0}, qdd0[7] = {0}, q_des[7] = {0};
    int num_obstacles = 0;
    std::vector<double> obstacles(MAX_OBSTACLE_NUM * 12, 0.0);
    std::ifstream in(pathname + "armour.in");
    if (!in.is_open()) { printf("        CUDA & C++: Error reading input files !\n"); out1 << -1; out1.close(); return 1; }
    for (double& v : q0) in >> v;
    for (double& v : qd0) in >> v;
    for (double& v : qdd0) in >> v;
    for (double& v : q_des) in >> v;
    in >> num_obstacles;
    if (num_obstacles > MAX_OBSTACLE_NUM || num_obstacles < 0) { printf("Number of obstacles larger than MAX_OBSTACLE_NUM !\n"); out1 << -1; out1.close(); return 1; }
    for (int i = 0; i < num_obstacles * 12; i++) in >> obstacles[i];
    in.close();
    const double t_plan = 0.5;   // KPR/armour_main.cu:81

    armour_config cfg;
    armour_default_config(&cfg);
    cfg.num_time_steps = NUM_TIME_STEPS;
#ifdef ARMOUR_HAVE_IPOPT
    // Ipopt hands the same g / Jacobian arrays to every callback of a solve: let the kernels write them directly (page-locked on
    // first sight, released by armtd_NLP::finalize_solution) instead of staging each result through a 1.2 MB host copy
    cfg.pin_user_buffers = 1;
#endif
    armour_handle* h = nullptr;
    if (armour_create(&cfg, &h) != ARMOUR_OK) { printf("        CUDA & C++: %s\n", armour_last_error()); out1 << -1 << '\n'; out1.close(); armour_destroy(h); return 1; }

    auto start1 = std::chrono::high_resolution_clock::now();
    if (armour_build(h, q0, qd0, qdd0, obstacles.data(), num_obstacles) != ARMOUR_OK) {
        printf("        CUDA & C++: Error computing link PZs and nominal torque PZs! %s\n", armour_last_error());
        out1 << -1 << '\n'; out1.close(); armour_destroy(h); return 1;
    }
    auto stop1 = std::chrono::high_resolution_clock::now();
    const long duration1 = std::chrono::duration_cast<std::chrono::milliseconds>(stop1 - start1).count();
    std::cout << "        CUDA & C++: Time taken by generating reachable sets: " << duration1 << " milliseconds" << std::endl;

    // KPR/armour_main.cu:228-230: the deadline formula is commented out in the reference; 10 s is what it passes to Ipopt
    double time_for_optimization = 10;
    time_for_optimization = std::max(time_for_optimization, 0.0);
    std::cout << "        CUDA & C++: Time allocated for Ipopt: " << time_for_optimization * 1000.0 << " milliseconds" << std::endl;

    auto start2 = std::chrono::high_resolution_clock::now();
    // The NLP lives on the heap, like the reference's (`SmartPtr<armtd_NLP> mynlp = new armtd_NLP()`, KPR/armour_main.cu:238):
    // Ipopt's SmartPtr deletes the object when the last reference goes away, so it must never point at a stack object.
#ifdef ARMOUR_HAVE_IPOPT
    Ipopt::SmartPtr<armtd_NLP> mynlp = new armtd_NLP();
#else
    std::unique_ptr<armtd_NLP> mynlp(new armtd_NLP());
#endif
    mynlp->verbose = true;
    mynlp->set_time_steps(NUM_TIME_STEPS);
    if (!mynlp->set_parameters(q_des, t_plan, h)) {
        printf("        CUDA & C++: Error initializing Ipopt! Check previous error message!\n");
        out1 << -1 << '\n'; out1.close(); armour_destroy(h); return 1;
    }
    bool report_solve_time = true;
#ifdef ARMOUR_HAVE_IPOPT
    {
        Ipopt::SmartPtr<Ipopt::IpoptApplication> app = Ipopt::IpoptApplicationFactory();   // options: KPR/armour_main.cu:256-261
        app->Options()->SetNumericValue("tol", 1e-4);
        app->Options()->SetNumericValue("max_wall_time", time_for_optimization);
        app->Options()->SetIntegerValue("print_level", 0);
        app->Options()->SetStringValue("mu_strategy", "adaptive");
        app->Options()->SetStringValue("linear_solver", "ma97");
        app->Options()->SetStringValue("hessian_approximation", "limited-memory");
        Ipopt::ApplicationReturnStatus status = app->Initialize();
        if (status != Ipopt::Solve_Succeeded) {   // :276-281 (the reference then rethrows and aborts; here: a clean non-zero exit)
            printf("Error during initialization!\n");
            out1 << -1 << '\n'; out1.close(); armour_destroy(h); return 1;
        }
        status = app->OptimizeTNLP(mynlp);
        // :295-317
        if (status == Ipopt::Invalid_Option) {
            std::cout << "        CUDA & C++: Cannot find HSL library! Need to put libcoinhsl.so in proper path!\n";
            report_solve_time = false;
        }
        else if (status == Ipopt::Maximum_CpuTime_Exceeded) {
            std::cout << "        CUDA & C++: Ipopt maximum CPU time exceeded!\n";
            std::cout << (mynlp->feasible ? "        CUDA & C++: Found a feasible solution!\n" : "        CUDA & C++: Did not find a feasible solution!\n");
        }
        else std::cout << (mynlp->feasible ? "        CUDA & C++: Found an optimal solution!\n" : "        CUDA & C++: Problem infeasible!\n");
    }
#else
    {
        double k_opt[7] = {0};
        StandinResult r = standin_solve(*mynlp, k_opt);
        std::cout << "        CUDA & C++: stand-in solver (Ipopt not available at build time): " << r.iterations << " iterations, " << r.evaluations
                  << " constraint evaluations, max violation " << r.max_violation << std::endl;
        std::cout << (mynlp->feasible ? "        CUDA & C++: Found an optimal solution!\n" : "        CUDA & C++: Problem infeasible!\n");
    }
#endif
    auto stop2 = std::chrono::high_resolution_clock::now();
    const long duration2 = std::chrono::duration_cast<std::chrono::milliseconds>(stop2 - start2).count();
    if (report_solve_time) std::cout << "        CUDA & C++: Time taken by Ipopt: " << duration2 << " milliseconds" << std::endl;
    armtd_NLP& nlp = *mynlp;
    if (nlp.link_sliced_center.empty()) nlp.link_sliced_center.assign((size_t)NUM_TIME_STEPS * 7 * 3, 0.0);   // no solve happened (Invalid_Option)

    out1 << std::setprecision(10);
    if (nlp.feasible) for (int i = 0; i < 7; i++) out1 << nlp.solution[i] << '\n';
    else out1 << -1 << '\n';
    out1 << duration1 + duration2;
    out1.close();

    const int T = NUM_TIME_STEPS;
    std::ofstream out2(pathname + "armour_joint_position_center.out");
    out2 << std::setprecision(10);
    for (int i = 0; i < T; i++) for (int j = 0; j < 7; j++) { for (int l = 0; l < 3; l++) out2 << nlp.link_sliced_center[((size_t)i * 7 + j) * 3 + l] << ' '; out2 << '\n'; }
    out2.close();
    std::vector<double> gens((size_t)T * 7 * 18), tr((size_t)T * 7);
    armour_get_link_generators(h, gens.data());
    armour_get_torque_radius(h, tr.data());
    std::ofstream out3(pathname + "armour_joint_position_radius.out");
    out3 << std::setprecision(10);
    for (int i = 0; i < T; i++) for (int j = 0; j < 7; j++) for (int k = 0; k < 3; k++) { for (int l = 0; l < 6; l++) out3 << gens[((size_t)i * 7 + j) * 18 + l * 3 + k] << ' '; out3 << '\n'; }
    out3.close();
    std::ofstream out4(pathname + "armour_control_input_radius.out");
    out4 << std::setprecision(10);
    for (int i = 0; i < T; i++) { for (int j = 0; j < 7; j++) out4 << tr[(size_t)i * 7 + j] << ' '; out4 << '\n'; }
    out4.close();
    std::ofstream out5(pathname + "armour_constraints.out");
    out5 << std::setprecision(6);
    for (int i = 0; i < nlp.constraint_number; i++) out5 << nlp.g_copy[i] << '\n';
    {   // the 14 position and 14 velocity bounds the MATLAB side reads back (KPR/armour_main.cu:381-394)
        std::vector<double> xl(7), xu(7), gl(nlp.constraint_number), gu(nlp.constraint_number);
        armour_get_bounds_info(h, xl.data(), xu.data(), gl.data(), gu.data());
        const int off = nlp.constraint_number - 28;
        for (int i = 0; i < 7; i++) { out5 << gl[off + i] << '\n' << gu[off + i] << '\n'; }
        for (int i = 0; i < 7; i++) { out5 << gl[off + 14 + i] << '\n' << gu[off + 14 + i] << '\n'; }
    }
    out5.close();
    armour_destroy(h);
    return 0;
}
