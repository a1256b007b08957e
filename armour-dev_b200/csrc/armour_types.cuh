// Device/host shared plain structs of the reachability + constraint path.
#pragma once
#include <stdint.h>

namespace armour {

constexpr int NJ = 7;       // NUM_JOINTS   (KPR/KinovaWithoutGripperInfo.h:10)
constexpr int NF = 7;       // NUM_FACTORS  (:14)
constexpr int COMB = 36;    // pairs of the 9 buffered generators (KPR/CollisionChecking.h:6-7)
constexpr int DEFAULT_UCAP = 128;   // k-only monomials kept per torque PZ after reduce(); runtime (Tables::ucap), grown on overflow
constexpr int DEFAULT_LCAP = 64;    // k-only monomials kept per link PZ after reduce_link_PZ(); runtime (Tables::lcap)
constexpr int SMALL_CAP = 4;

// Robot constants (KPR/KinovaWithoutGripperInfo.h:10-112) after host-side preparation.
struct RobotModel {
    int axes[NJ];
    double trans[NJ + 1][3];
    double R0[NJ + 1][9];        // fixed rpy rotation of each joint frame, column-major; [NJ] = identity (PZsparse.cu:160-176)
    double mass[NJ];
    double com[NJ][3];
    double inertia[NJ][9];       // column-major, filled by linear index like Dynamics.cu:36-38
    double armature[NJ], damping[NJ], friction[NJ];
    double link_c[NJ][3], link_g[NJ][3];
    double state_lb[NF], state_ub[NF], speed[NF], torque[NF];
    double gravity;
    double alpha, M_max, M_min, eps, qe, qde, qdae, qddae;
    double qdd_k_maxima, qdd_k_minima;        // KPR/Trajectory.h:7-8
    double qdd_k_maxima_val, qdd_k_minima_val;   // 60 t (2 t^2 - 3 t + 1) at those points (Trajectory.cu:202,210)
};

// fixed-size record of a tiny PZ (trajectory tables: at most 4 monomials)
struct SmallRec {
    int n, dim;
    unsigned long long keys[SMALL_CAP];
    double coef[SMALL_CAP][9];
    double center[9];
    double ind[9];
};
enum { TRAJ_COS = 0, TRAJ_SIN = 1, TRAJ_R = 2, TRAJ_RT = 3, TRAJ_QD = 4, TRAJ_QDA = 5, TRAJ_QDDA = 6, TRAJ_TABLES = 7 };

struct Tables {
    int T, P, n_obs;
    int ucap, lcap;            // capacities of the k-only monomial tables below
    int static_stride;         // 1: persistent CTAs take work items blockIdx + k * gridDim instead of a shared counter (experiments)
    int no_structured;         // 1: every product takes the generic sort path (A/B measurements; ARMOUR_TUNE_NO_STRUCTURED)
    int fuse_planes;           // 1: reach_build_kernel also writes the interval's half-space tables (stage D) instead of a separate hyperplane_kernel launch
    double k_range[NF];
    double mass_unc, inertia_unc, thr;
    // inputs
    int mode;                  // 0 ARMOUR (Bezier trajectory + RNEA); 1 ARMTD comparison planner (offline JRS tables, FK only)
    const double* jrs;         // mode 1: [P][6][NF][T] c_cos, g_cos, r_cos, c_sin, g_sin, r_sin (KPA/armtd_main.cu:41-47)
    const double* k_range_in;  // mode 1: [P][NF] (k_range is an input of that planner)
    const double* state;       // [P][21] q0, qd0, qdd0
    const double* obstacles;   // [P][n_obs][12]
    // trajectory tables (export)
    SmallRec* traj;            // [P][T][TRAJ_TABLES][NJ]
    double* cos_rem;           // [P][NJ][T][2]
    double* sin_rem;
    // torque PZs after reduce(): k-only monomials
    int* u_n;                  // [P][T][NF]
    unsigned long long* u_keys;   // [P][T][NF][ucap]
    double* u_coef;            // [P][T][NF][ucap]
    double* u_center;          // [P][T][NF]
    double* u_ind;             // [P][T][NF] radius after reduce()
    double* dist_rad;          // [P][T][NF] radius of u_nom_int - u_nom
    double* torque_radius;     // [P][T][NF]
    // link PZs after reduce_link_PZ()
    int* l_n;                  // [P][T][NJ]
    unsigned long long* l_keys;   // [P][T][NJ][lcap]
    double* l_coef;            // [P][T][NJ][3][lcap]
    double* l_center;          // [P][T][NJ][3]
    double* l_ind;             // [P][T][NJ][3]
    double* gens;              // [P][T][NJ][18] 3x6 column-major
    // half-space tables
    double* A;                 // [P][T][NJ][n_obs][COMB][3]
    double* d;                 // [P][T][NJ][n_obs][COMB]
    double* delta;
    // per-problem Bezier constants needed by the limit rows of the constraint kernel are recomputed there
    int* err;                  // [3] error word, work counter of the persistent reach kernel, guard words verified (debug build)
};

}  // namespace armour
