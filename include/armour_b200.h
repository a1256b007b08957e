/* armour_b200.h — C ABI of the B200-native ARMOUR reachability + constraint path.
 *
 * Drop-in boundary for the path the reference implements in
 * kinova_src/kinova_simulator_interfaces/kinova_planner_realtime/ ("KPR"):
 *   BezierCurve::makePolyZono            KPR/Trajectory.cu:63-254
 *   KinematicsDynamics::fk / rnea        KPR/Dynamics.cu:69-181
 *   reach-set build loop + torque radius KPR/armour_main.cu:94-211
 *   Obstacles (half-space tables)        KPR/CollisionChecking.cu:74-134
 *   armtd_NLP : Ipopt::TNLP callbacks    KPR/NLPclass.cu:62-538
 * A maintainer binds these entry points from the reference's host code:
 * plain pointers and sizes only, no C++ or torch types.  All functions return 0 on success and a
 * negative ARMOUR_E_* code on failure; nothing throws across this boundary.  There is no CPU
 * fallback: every entry point that computes fails with ARMOUR_E_CUDA when no device is usable.
 *
 * Conventions shared with the reference:
 *   - 7 joints / 7 trajectory parameters k in [-1,1]; T time intervals (NUM_TIME_STEPS, runtime here).
 *   - obstacles: n_obs x 12 doubles = centre(3), generator1(3), generator2(3), generator3(3)
 *     (KPR/armour_main.cu:47-79, column-major 3x4 zonotope matrix written by uarmtd_planner.m:189).
 *   - constraint vector g (m = 7T + 7T*n_obs + 28 rows, KPR/NLPclass.cu:47-49):
 *       [0, 7T)                 torque PZ centres sliced at k, index t*7 + joint        (:306-309)
 *       [7T, 7T + 7T*n_obs)     obstacle rows, index (link*T + t)*n_obs + obs          (:317)
 *       then 7 min-position, 7 max-position, 7 min-velocity, 7 max-velocity rows       (:319-320)
 *   - Jacobian: dense, row-major, values[row*7 + col]                                  (:348-357)
 * The handle is not re-entrant: one host thread at a time, one CUDA stream per handle.
 */
#ifndef ARMOUR_B200_H
#define ARMOUR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARMOUR_NUM_JOINTS 7
#define ARMOUR_NUM_FACTORS 7
#define ARMOUR_COMB_NUM 36 /* pairs of the 9 buffered generators, KPR/CollisionChecking.h:6-7 */

#define ARMOUR_OK 0
#define ARMOUR_E_INVALID (-1)  /* bad argument (null pointer, n_obs out of range, odd T ...) */
#define ARMOUR_E_CUDA (-2)     /* CUDA runtime error or no usable device; see armour_last_error() */
#define ARMOUR_E_CAPACITY (-3) /* a monomial list outgrew the configured capacities even after the automatic retries (every
                                * capacity — monomials per PZ, candidates per operation, k-only table rows — doubles and the build re-runs) */
#define ARMOUR_E_STATE (-4)    /* call order violated (e.g. eval before build) */
#define ARMOUR_E_NUMERIC (-5)  /* monomial degree overflow or more than 3 pure link generators */

typedef struct armour_handle armour_handle;

/* Runtime equivalents of the reference's compile-time macros (KPR/Parameters.h, KinovaWithoutGripperInfo.h). */
typedef struct armour_config {
    int num_time_steps;         /* NUM_TIME_STEPS, even; default 128                    (Parameters.h:17) */
    double k_range[7];          /* default pi/48 each                                   (Parameters.h:21) */
    double mass_uncertainty;    /* default 0.03                     (KinovaWithoutGripperInfo.h:41) */
    double inertia_uncertainty; /* default 0.03                     (KinovaWithoutGripperInfo.h:61) */
    double simplify_threshold;  /* SIMPLIFY_THRESHOLD, default 5e-4                     (Parameters.h:10) */
    int max_obstacles;          /* MAX_OBSTACLE_NUM, default 40, at most 64             (Parameters.h:26) */
    int max_monomials;          /* capacity of one PZ's monomial list; default 1024 (0 = default)  */
    int max_entries;            /* capacity of one sort (candidate monomials of one op); default 8192, at most 65535 (0 = default).
                                 * Operations up to 2048 candidates sort in shared memory, larger ones in global memory. */
    int threads_per_cta;        /* threads working on one time interval: 32, 64, 128, 256 or 512; 0 = default (one plan: two groups of
                                 * 256; a batch: the measured sweep shape) */
    int device;                 /* CUDA device ordinal; -1 = current device              */
    int batch;                  /* problems one handle builds per armour_build_batch call; default 1 */
    int pin_user_buffers;       /* 1: armour_eval_g_jac page-locks the caller's g / values arrays (cudaHostRegister) and the
                                 * kernel writes into them directly, no staging copy.  CONTRACT: the arrays must stay allocated
                                 * until armour_release_host_buffers / armour_destroy — a registered array that is freed and whose
                                 * address the allocator hands out again cannot be detected.  Pass the same arrays every call (Ipopt
                                 * does); at most 8 distinct arrays are kept registered.  Default 0 (staging through pinned buffers). */
    int export_trajectory_tables; /* the joint trajectory PZs (cos q, sin q, R, R_t, qd_des, qda_des, qdda_des; armour_get_pz tables
                                 * 0..6) are only read by tests and the PZsparse facade: 1 = write them during the build, -1 = never,
                                 * 0 = default (only for single-problem handles, cfg.batch == 1) */
} armour_config;

void armour_default_config(armour_config* cfg);

int armour_create(const armour_config* cfg, armour_handle** out);
void armour_destroy(armour_handle* h);
const char* armour_last_error(void);

/* Stages A-D of the reference's main() (KPR/armour_main.cu:89-226): joint reach sets, PZ forward
 * kinematics, PZ-RNEA (nominal + uncertain parameters in one pass), torque radius, half-space tables.
 * Inputs are host pointers; everything is computed on the device.  Replaces BezierCurve::BezierCurve +
 * makePolyZono, KinematicsDynamics::{fk,rnea_nominal,rnea_interval}, PZsparse::reduce[_link_PZ] and
 * Obstacles::initializeHyperPlane. */
int armour_build(armour_handle* h, const double* q0, const double* qd0, const double* qdd0, const double* obstacles, int n_obs);

/* Same for `count` independent problems (count <= cfg.batch); arrays are [count][7] and
 * [count][n_obs][12], every problem with the same n_obs.  Problem `p` is then selected for the
 * callbacks below with armour_select_problem. */
int armour_build_batch(armour_handle* h, int count, const double* q0, const double* qd0, const double* qdd0, const double* obstacles, int n_obs);
int armour_select_problem(armour_handle* h, int p);

/* Second trajectory family: the ARMTD comparison planner (kinova_planner_realtime_armtd_comparison, "KPA"):
 * constant-acceleration trajectory whose joint reach sets come from OFFLINE JRS tables, forward occupancy only
 * (no RNEA, no torque rows).  jrs = [6][7][T] doubles in the order c_cos, g_cos, r_cos, c_sin, g_sin, r_sin and
 * k_range[7] are inputs of that planner (KPA/armtd_main.cu:41-90); use cfg.num_time_steps = 100 for its defaults.
 * Replaces ConstantAccelerationCurve::makePolyZono (KPA/Trajectory.cu:29-81), KinematicsDynamics::fk,
 * reduce_link_PZ and Obstacles::initializeHyperPlane.  The TNLP entry points then follow KPA/NLPclass.cu:
 * m = 7*T*n_obs + 28, obstacle rows first, then the joint state extrema of returnJointStateExtremum[Gradient]
 * (KPA/Trajectory.cu:83-411; their k-derivatives carry no k_range factor in the reference and none here). */
int armour_build_armtd(armour_handle* h, const double* q0, const double* qd0, const double* jrs, const double* k_range, const double* obstacles, int n_obs);

/* armtd_NLP::get_nlp_info (KPR/NLPclass.cu:62-82) */
int armour_get_nlp_info(armour_handle* h, int* n, int* m, int* nnz_jac_g, int* nnz_h_lag);
/* armtd_NLP::get_bounds_info (:87-165) */
int armour_get_bounds_info(armour_handle* h, double* x_l, double* x_u, double* g_l, double* g_u);
/* armtd_NLP::get_starting_point (:170-202) */
int armour_get_starting_point(armour_handle* h, double* x);
/* armtd_NLP::eval_f / eval_grad_f (:207-267); q_des and t_plan are set_parameters' arguments (:30-60) */
int armour_eval_f(armour_handle* h, const double* q_des, double t_plan, const double* x, double* obj_value);
int armour_eval_grad_f(armour_handle* h, const double* q_des, double t_plan, const double* x, double* grad_f);
/* armtd_NLP::eval_g (:272-324) and eval_jac_g (:330-396).  One kernel slices the torque and link PZs at k and tests every link
 * against every obstacle; each call computes and transfers only what it is asked for (Ipopt calls eval_g at every trial point
 * and eval_jac_g once per accepted iterate).  Without cfg.pin_user_buffers a repeated request at the same x is served from the
 * handle's pinned result buffers.  armour_eval_g_jac asks for both in one launch. */
int armour_eval_g(armour_handle* h, const double* x, double* g);
int armour_eval_jac_g(armour_handle* h, const double* x, double* values);
int armour_eval_g_jac(armour_handle* h, const double* x, double* g, double* values);
int armour_jac_structure(armour_handle* h, int* iRow, int* jCol);
/* undo the page-locking done under cfg.pin_user_buffers (call before freeing those arrays) */
int armour_release_host_buffers(armour_handle* h);
/* caller arrays currently page-locked by this handle (cfg.pin_user_buffers); arrays that the library itself pinned for the
 * duration of a call (armour_standin_solve) are released before that call returns */
int armour_pinned_buffer_count(armour_handle* h, int* count);
/* armtd_NLP::finalize_solution's feasibility re-check (:446-537): *feasible = 1 or 0 */
int armour_check_feasible(armour_handle* h, const double* g, int* feasible);

/* ---- results the reference exposes as public members / output files --------------------------------- */
/* torque_radius(j, t) (KPR/armour_main.cu:173-211) as out[t*7 + j] */
int armour_get_torque_radius(armour_handle* h, double* out);
/* link_independent_generators[t*7 + link] (3x6, column-major) as out[(t*7+link)*18 + col*3 + row] */
int armour_get_link_generators(armour_handle* h, double* out);
/* armtd_NLP::link_sliced_center[t*7 + link] at the x of the last eval (KPR/NLPclass.h:150) */
int armour_get_link_sliced_center(armour_handle* h, double* out);
/* half-space tables A, d, delta indexed ((t*7+link)*n_obs+obs)*36 + pair (KPR/CollisionChecking.cu:215-227) */
int armour_get_hyperplanes(armour_handle* h, double* A, double* d, double* delta);
/* Boost-interval Taylor remainders of cos/sin(q_des) before re-centring (KPR/Trajectory.cu:104-106,121-123),
 * out[(joint*T + t)*2 + {lo,hi}] — exported for enclosure tests */
int armour_get_taylor_remainders(armour_handle* h, double* cos_rem, double* sin_rem);

/* PZsparse tables (the reference's PZsparseArray members), for the PZsparse facade and parity tests.
 * which: 0 cos_q_des, 1 sin_q_des, 2 R, 3 R_t, 4 qd_des, 5 qda_des, 6 qdda_des   (KPR/Trajectory.h:57-70)
 *        7 links after reduce_link_PZ, 8 u_nom after reduce, 9 u_nom_int after the subtraction (KPR/Dynamics.h:21-26)
 * Returns the monomial count (>= 0) or a negative error.  Null output pointers are skipped.
 * coeffs[i*rows*cols + (row + col*rows)], keys ascending. */
int armour_get_pz(armour_handle* h, int which, int idx, int t, int* dims, uint64_t* keys, double* coeffs, double* center, double* independent);

/* ---- stand-alone PZsparse arithmetic on the device (PZsparse facade; primitive parity tests) --------- */
/* op: 0 a*b, 1 a+b, 2 a-b, 3 cross(a,b) for 3x1 operands, 4 simplify(a) (any monomial list: unsorted, repeated keys; b is ignored,
 * pass a 1x1 with b_n = 0), 7/8/9 a.addOneDimPZ(b, row 0/1/2, 0) for a 3x1 and b 1x1, 10 cross(c, a) and 11 cross(a, c) with the
 * constant 3-vector c = b_center (b 3x1, b_n = 0).  (5 and 6 — reduce, transpose — are pure data movement and live in the host
 * facade.)  Shapes supported: (3x3)*(3x1), (3x3)*(3x3), (1x1)*(1x1); + and - for 1x1 and 3x1; simplify for 1x1, 3x1, 3x3;
 * any other shape is rejected with ARMOUR_E_INVALID before anything is read.
 * Returns the monomial count (>= 0) or a negative ARMOUR_E_* code: ARMOUR_E_CAPACITY when the result has more than `cap`
 * monomials or the candidate list exceeds cfg.max_entries, ARMOUR_E_NUMERIC when a monomial degree outgrows its key field. */
int armour_pz_binary(armour_handle* h, int op,
                     int a_rows, int a_cols, int a_n, const uint64_t* a_keys, const double* a_coeffs, const double* a_center, const double* a_independent,
                     int b_rows, int b_cols, int b_n, const uint64_t* b_keys, const double* b_coeffs, const double* b_center, const double* b_independent,
                     int cap, int* dims, uint64_t* keys, double* coeffs, double* center, double* independent);

/* Stand-in NLP solve for hosts without Ipopt (this image): quadratic-penalty Gauss-Newton over the same callbacks
 * Ipopt would call (armour-dev_b200/host/standin_solver.hpp).  NOT the reference's solver — results are labelled
 * "stand-in" wherever reported.  k_opt[7]; *feasible as armour_check_feasible at the returned point. */
int armour_standin_solve(armour_handle* h, const double* q_des, double t_plan, double* k_opt, int* feasible, int* iterations, int* evaluations);

/* ---- timing / accounting ------------------------------------------------------------------------------ */
/* device time of the last build / eval in milliseconds (CUDA events on the handle's stream) */
int armour_last_build_ms(armour_handle* h, float* total_ms, float* reach_kernel_ms, float* hyperplane_kernel_ms);
int armour_last_eval_ms(armour_handle* h, float* kernel_ms);
/* Host-buffer evaluations (armour_eval_g / _jac_g / _g_jac) record CUDA events for armour_last_eval_ms only when asked to:
 * two event records cost about 4 us per call on the per-iteration path.  enabled = 1 turns them on, 0 (default) off
 * (armour_last_eval_ms then reports -1 after a host-buffer evaluation; armour_eval_resident is always timed). */
int armour_set_kernel_timing(armour_handle* h, int enabled);
/* wall-clock microseconds spent inside the last evaluation call (launch, PCIe transfer, completion wait) */
int armour_last_eval_host_us(armour_handle* h, double* microseconds);
/* kernels launched by this handle since creation */
int armour_kernel_launches(armour_handle* h, uint64_t* launches);
/* build with inputs already resident on the device (armour_build without the host<->device copies);
 * used by bench.py's device-resident timing */
int armour_upload_problems(armour_handle* h, int count, const double* q0, const double* qd0, const double* qdd0, const double* obstacles, int n_obs);
int armour_build_resident(armour_handle* h);
int armour_eval_resident(armour_handle* h, const double* x); /* x == NULL: reuse the x uploaded by armour_upload_x */
int armour_upload_x(armour_handle* h, const double* x);
/* `launches` device-resident evaluations back to back between two CUDA events: average kernel time without the ~6 us that
 * a single launch bracketed by two events reads high (roofline measurement of bench.py) */
int armour_eval_resident_burst(armour_handle* h, const double* x, int launches, float* ms_per_launch);
/* Batched evaluation after armour_build_batch / armour_build_resident: ONE launch evaluates problems first .. first+count-1
 * of the last batch, problem y at x[7 y .. 7 y + 6] (the per-iteration call of a batched solver; the reference has no
 * counterpart — it runs one armtd_NLP::eval_g / eval_jac_g (KPR/NLPclass.cu:272-396) per problem and process).  Rows of
 * problem first + y land at g[y * m ..] and values[y * 7 m ..]; either may be NULL.  Bit-identical to `count` single calls. */
int armour_eval_batch(armour_handle* h, int first, int count, const double* x, double* g, double* values);
int armour_last_eval_batch_ms(armour_handle* h, float* ms); /* device time of the last armour_eval_batch kernel */
/* The same launch with the Jacobians left on the device for a device-side consumer (a batched QP / line search): only the
 * constraint values cross PCIe (8 m instead of 64 m bytes per problem).  g may be NULL.  armour_batch_jacobian_device hands
 * out the device array ([problems][7 m] doubles, valid until the next batched evaluation or build);
 * armour_get_batch_jacobian copies one problem's rows to the host (tests, debugging). */
int armour_eval_batch_resident(armour_handle* h, int first, int count, const double* x, double* g);
int armour_batch_jacobian_device(armour_handle* h, const double** values_dev, int* problems);
int armour_get_batch_jacobian(armour_handle* h, int y, double* values);
/* profiling builds (-DARMOUR_PHASE_TIMING) only: cycles / calls per engine phase summed over CTAs; zeros otherwise.
 * phases: 0 fill, 1 sort level, 2 segment walk, 3 scan+compact, 4 element-wise, 5 stage A, 6 export, 7 other */
int armour_debug_phase_cycles(uint64_t* cycles8, uint64_t* calls8, int reset);
/* debug builds (-DARMOUR_ARENA_CANARY, `make canary`) only: guard words behind the arena's PZ slots that the last build
 * found intact, summed over CTAs (0 in the product build); an overwritten guard word fails the build with ARMOUR_E_CUDA */
int armour_debug_canaries_verified(armour_handle* h, int* count);
/* fp64 FMA micro-benchmark (TFLOP/s) used as the fp64 roofline denominator */
int armour_measure_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* ARMOUR_B200_H */
