// C ABI of the B200-native ARMOUR reachability + constraint path (see include/armour_b200.h).
// Host-side glue only: device memory, one stream per handle, pinned staging buffers, the cheap
// host-side TNLP callbacks (bounds, objective) and table getters.  All heavy work is in
// reach_kernels.cu / constraint_kernels.cu.  There is no CPU fallback anywhere in this file.
#include "../../include/armour_b200.h"
#include "armour_launch.h"
#include "../host/standin_solver.hpp"

#include <chrono>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost nanoseconds unless a profiler is attached
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <algorithm>
#include <vector>

using namespace armour;

namespace {

thread_local std::string g_last_error;
int fail(int code, const std::string& msg) { g_last_error = msg; return code; }
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) return fail(ARMOUR_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

}  // namespace
namespace armour {
int capi_fail(int code, const std::string& msg) { return fail(code, msg); }   // used by controller_kernels.cu
}
namespace {

// Kinova Gen3 without gripper (KPR/KinovaWithoutGripperInfo.h:10-112)
void kinova_model(RobotModel& m) {
    memset(&m, 0, sizeof(m));
    const int axes[NJ] = {3, 3, 3, 3, 3, 3, 3};
    const double trans[(NJ + 1) * 3] = {0, 0, 0.15643, 0, 0.005375, -0.12838, 0, -0.21038, -0.006375, 0, 0.006375, -0.21038,
                                        0, -0.20843, -0.006375, 0, 0.00017505, -0.10593, 0, -0.10593, -0.00017505, 0, 0, 0};
    const double rots[NJ * 3] = {M_PI, 0, 0, M_PI * 0.5, 0, 0, -M_PI * 0.5, 0, 0, M_PI * 0.5, 0, 0, -M_PI * 0.5, 0, 0, M_PI * 0.5, 0, 0, -M_PI * 0.5, 0, 0};
    const double mass[NJ] = {1.3773, 1.1636, 1.1636, 0.9302, 0.6781, 0.6781, 0.5};
    const double com[NJ * 3] = {-0.000023, -0.010364, -0.07336, -0.000044, -0.09958, -0.013278, -0.000044, -0.006641, -0.117892, -0.000018, -0.075478, -0.015006,
                                0.000001, -0.009432, -0.063883, 0.000001, -0.045483, -0.00965, 0.000281, 0.011402, -0.029798};
    const double inertia[NJ * 9] = {0.00457, 0.000001, 0.000002, 0.000001, 0.004831, 0.000448, 0.000002, 0.000448, 0.001409,
                                    0.011088, 0.000005, 0, 0.000005, 0.001072, -0.000691, 0, -0.000691, 0.011255,
                                    0.010932, 0, -0.000007, 0, 0.011127, 0.000606, -0.000007, 0.000606, 0.001043,
                                    0.008147, -0.000001, 0, -0.000001, 0.000631, -0.0005, 0, -0.0005, 0.008316,
                                    0.001596, 0, 0, 0, 0.001607, 0.000256, 0, 0.000256, 0.000399,
                                    0.001641, 0, 0, 0, 0.00041, -0.000278, 0, -0.000278, 0.001641,
                                    0.000587, 0.000003, 0.000003, 0.000003, 0.000369, -0.000118, 0.000003, -0.000118, 0.000609};
    const double armature[NJ] = {8.03, 11.9962024615303644, 9.0025427861751517, 11.5806439316706360, 8.4665040917914123, 8.8537069373742430, 8.8587303664685315};
    const double lb[NF] = {-1000.0, -2.41, -1000.0, -2.66, -1000.0, -2.23, -1000.0};
    const double ub[NF] = {1000.0, 2.41, 1000.0, 2.66, 1000.0, 2.23, 1000.0};
    const double speed[NF] = {1.3963, 1.3963, 1.3963, 1.3963, 1.2218, 1.2218, 1.2218};
    const double torque[NF] = {56.7, 56.7, 56.7, 56.7, 29.4, 29.4, 29.4};
    const double lzc[NJ][3] = {{0.000000, -0.001297, -0.088375}, {0.000000, -0.089400, -0.007877}, {0.000000, -0.001502, -0.129375}, {0.000000, -0.087450, -0.013648},
                               {0.000001, -0.009023, -0.071752}, {0.000000, -0.041661, -0.009251}, {0.000000, -0.018585, -0.033462}};
    const double lzg[NJ][3] = {{0.046358, 0.047354, 0.086000}, {0.046000, 0.135400, 0.047501}, {0.046000, 0.047501, 0.127000}, {0.046000, 0.133450, 0.042293},
                               {0.034999, 0.044023, 0.069252}, {0.035000, 0.076739, 0.044076}, {0.045500, 0.056085, 0.030963}};
    for (int i = 0; i < NJ; i++) {
        m.axes[i] = axes[i];
        m.mass[i] = mass[i]; m.armature[i] = armature[i]; m.damping[i] = 0.0; m.friction[i] = 0.0;
        for (int a = 0; a < 3; a++) { m.com[i][a] = com[3 * i + a]; m.link_c[i][a] = lzc[i][a]; m.link_g[i][a] = lzg[i][a]; }
        for (int a = 0; a < 9; a++) m.inertia[i][a] = inertia[9 * i + a];
        m.state_lb[i] = lb[i]; m.state_ub[i] = ub[i]; m.speed[i] = speed[i]; m.torque[i] = torque[i];
    }
    for (int i = 0; i <= NJ; i++) for (int a = 0; a < 3; a++) m.trans[i][a] = trans[3 * i + a];
    // fixed frame rotations from roll/pitch/yaw, host libm like the reference (KPR/PZsparse.cu:160-176); column-major
    for (int i = 0; i <= NJ; i++) {
        const double roll = i < NJ ? rots[3 * i] : 0.0, pitch = i < NJ ? rots[3 * i + 1] : 0.0, yaw = i < NJ ? rots[3 * i + 2] : 0.0;
        double* R = m.R0[i];
        R[0 + 3 * 0] = cos(pitch) * cos(yaw);
        R[0 + 3 * 1] = -cos(pitch) * sin(yaw);
        R[0 + 3 * 2] = sin(pitch);
        R[1 + 3 * 0] = cos(roll) * sin(yaw) + cos(yaw) * sin(pitch) * sin(roll);
        R[1 + 3 * 1] = cos(roll) * cos(yaw) - sin(pitch) * sin(roll) * sin(yaw);
        R[1 + 3 * 2] = -cos(pitch) * sin(roll);
        R[2 + 3 * 0] = sin(roll) * sin(yaw) - cos(roll) * cos(yaw) * sin(pitch);
        R[2 + 3 * 1] = cos(yaw) * sin(roll) + cos(roll) * sin(pitch) * sin(yaw);
        R[2 + 3 * 2] = cos(pitch) * cos(roll);
    }
    m.gravity = 9.81;
    m.alpha = 10.0; m.M_max = 15.79635774; m.M_min = 5.095620491878957;
    const double V_m = 1e-2, K = 5.0;
    m.eps = sqrt(2 * V_m / m.M_min);
    m.qe = m.eps / K; m.qde = 2 * m.eps; m.qdae = m.eps; m.qddae = 2 * K * m.eps;
    m.qdd_k_maxima = (0.5 - sqrt(3) / 6); m.qdd_k_minima = (0.5 + sqrt(3) / 6);   // KPR/Trajectory.h:7-8
    m.qdd_k_maxima_val = (60 * m.qdd_k_maxima * (2 * pow(m.qdd_k_maxima, 2) - 3 * m.qdd_k_maxima + 1)) / 1.0 / 1.0;   // Trajectory.cu:202
    m.qdd_k_minima_val = (60 * m.qdd_k_minima * (2 * pow(m.qdd_k_minima, 2) - 3 * m.qdd_k_minima + 1)) / 1.0 / 1.0;   // Trajectory.cu:210
}

const double TORQUE_THRESHOLD = 1e-2, COLLISION_THRESHOLD = 1e-4, COST_SCALE = 10.0;   // KPR/Parameters.h:38-44

double wrap_to_pi(double angle) {   // KPR/NLPclass.cu:6-15
    double w = angle;
    while (w < -M_PI) w += 2 * M_PI;
    while (w > M_PI) w -= 2 * M_PI;
    return w;
}
double q_des_host(double q0, double a, double b, double k, double t) {   // KPR/Trajectory.cu:542-556
    const double B0 = -pow(t - 1, 5), B1 = 5 * t * pow(t - 1, 4), B2 = -10 * pow(t, 2) * pow(t - 1, 3);
    const double B3 = 10 * pow(t, 3) * pow(t - 1, 2), B4 = -5 * pow(t, 4) * (t - 1), B5 = pow(t, 5);
    const double b0 = q0, b1 = q0 + a / 5, b2 = q0 + (2 * a) / 5 + b / 20, b3 = q0 + k;
    return B0 * b0 + B1 * b1 + B2 * b2 + B3 * b3 + B4 * b3 + B5 * b3;
}

template <class T>
cudaError_t dalloc(T** p, size_t count) { return cudaMalloc((void**)p, count * sizeof(T) > 0 ? count * sizeof(T) : 16); }

}  // namespace

struct armour_handle {
    armour_config cfg;
    int device = 0, sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    RobotModel model;
    Tables tb;
    int P = 1, T = 128, max_obs = 40;
    int mcap = 1024, ncap = 8192, nt = 256, minb = 1, groups = 2, groups_cfg = 2;
    int scap = 2048, tcap = 512;   // shared-memory sort / staging capacity per thread group (larger operations use global buffers)
    char* arena = nullptr;
    size_t arena_stride = 0;
    int grid = 0;
    // device buffers
    double *d_jrs = nullptr, *d_krange = nullptr, *h_jrs = nullptr, *h_krange = nullptr;
    int mode = 0;   // 0 ARMOUR, 1 ARMTD comparison planner
    double *d_state = nullptr, *d_obs = nullptr, *d_x = nullptr, *d_g = nullptr, *d_jac = nullptr, *d_link_center = nullptr;
    int* d_err = nullptr;
    unsigned* d_done = nullptr;                    // blocks-finished counter of constraint_eval_kernel
    volatile unsigned long long* h_done = nullptr; // completion word the kernel's last block writes (pinned, mapped)
    unsigned long long* d_done_flag = nullptr;     // device alias of h_done
    unsigned long long eval_seq = 0;
    // cuStreamWriteValue64, resolved through the runtime (no link-time dependency on libcuda): writes the completion word
    // after the constraint kernel in stream order, so the kernel needs no system-wide fences of its own
    int (*write_value64)(cudaStream_t, unsigned long long, unsigned long long, unsigned) = nullptr;
    int host_write = 0;                            // how results reach host memory, see launch_eval
    double eval_host_us = 0.0;                     // wall-clock time spent inside the last evaluation call
    double *a_g = nullptr, *a_jac = nullptr;       // device aliases of the pinned staging buffers h_g / h_jac
    double *d_bx = nullptr, *d_bg = nullptr, *d_bjac = nullptr, *h_bx = nullptr;   // armour_eval_batch: decision vectors and result rows of a whole batch
    size_t bx_cap = 0, bg_cap = 0, bjac_cap = 0, bjac_rows = 0;   // bjac_rows: problems whose Jacobian the last batched evaluation left in d_bjac
    double *h_ws = nullptr, *a_ws = nullptr;   // page-locked solver workspace (armour_standin_solve) and its device alias
    size_t ws_doubles = 0;
    bool fuse_planes = false;  // ARMOUR_TUNE_FUSE_PLANES=1: stage D inside reach_build_kernel instead of the separate hyperplane_kernel launch
    float batch_eval_ms = 0;
    int eval_bps_host = 0;                         // resident blocks per SM of the constraint kernel when it writes to host memory (waves overlap compute and PCIe)
    bool eval_timed = false;                       // events of the last evaluation are pending in ev[3], ev[4]
    bool time_kernels = false;                     // record CUDA events around the per-iteration kernel (armour_set_kernel_timing)
    int ucap = DEFAULT_UCAP, lcap = DEFAULT_LCAP;  // k-only table capacities (grown on overflow)
    // pinned host buffers
    double *h_state = nullptr, *h_obs = nullptr, *h_x = nullptr, *h_g = nullptr, *h_jac = nullptr, *h_torque_radius = nullptr;
    int* h_err = nullptr;
    // state
    int count = 0, n_obs = 0, sel = 0;
    bool built = false, have_eval = false;
    int staged = 0;   // which results of the evaluation at last_x sit in the pinned staging buffers: bit 0 g, bit 1 Jacobian
    double last_x[NF];
    float build_ms = 0, reach_ms = 0, hyper_ms = 0, eval_ms = 0;
    uint64_t launches = 0;
    // lazy host mirrors for getters
    bool mirror_valid = false;
    std::vector<SmallRec> m_traj;
    std::vector<int> m_un, m_ln;
    std::vector<unsigned long long> m_ukeys, m_lkeys;
    std::vector<double> m_ucoef, m_ucenter, m_uind, m_dist, m_lcoef, m_lcenter, m_lind;
    // caller arrays page-locked under cfg.pin_user_buffers: host pointer, bytes, device alias
    struct Pinned { const void* host; size_t bytes; void* dev; };
    std::vector<Pinned> pinned;
    // pz_binary scratch
    char* bin_buf = nullptr;
    size_t bin_bytes = 0;
};

namespace {

int m_of(const armour_handle* h) { return (h->mode == 0 ? NF * h->T : 0) + NJ * h->T * h->n_obs + NF * 4; }

void free_arena(armour_handle* h) { if (h->arena) cudaFree(h->arena); h->arena = nullptr; }
int alloc_arena(armour_handle* h) {
    free_arena(h);
    // two thread groups per CTA (joint chain || forces + FK) when their sort buffers fit in shared memory
    h->groups = h->groups_cfg;
    int per_sm = h->groups == 2 ? reach_max_ctas_per_sm(h->nt, h->minb, 2, h->scap, h->tcap) : 0;
    if (per_sm < 1) { h->groups = 1; per_sm = reach_max_ctas_per_sm(h->nt, h->minb, 1, h->scap, h->tcap); }
    if (per_sm < 1) return fail(ARMOUR_E_CUDA, "reach_build_kernel does not fit on an SM with these capacities");
    h->arena_stride = arena_bytes(h->mcap, h->ncap, h->groups);
    const int n_work = h->P * h->T;
    h->grid = std::min(n_work, per_sm * h->sm_count);
    CU(cudaMalloc((void**)&h->arena, h->arena_stride * (size_t)h->grid));
    return ARMOUR_OK;
}

// k-only monomial tables of the torque / link PZs (read by every constraint evaluation); capacities are runtime and grow
int alloc_konly_tables(armour_handle* h) {
    Tables& tb = h->tb;
    for (void* p : {(void*)tb.u_keys, (void*)tb.u_coef, (void*)tb.l_keys, (void*)tb.l_coef}) if (p) cudaFree(p);
    tb.u_keys = nullptr; tb.u_coef = nullptr; tb.l_keys = nullptr; tb.l_coef = nullptr;
    const size_t P = h->P, T = h->T;
    tb.ucap = h->ucap; tb.lcap = h->lcap;
    CU(dalloc(&tb.u_keys, P * T * NF * h->ucap)); CU(dalloc(&tb.u_coef, P * T * NF * h->ucap));
    CU(dalloc(&tb.l_keys, P * T * NJ * h->lcap)); CU(dalloc(&tb.l_coef, P * T * NJ * 3 * h->lcap));
    return ARMOUR_OK;
}

struct NvtxRange { explicit NvtxRange(const char* name) { nvtxRangePushA(name); } ~NvtxRange() { nvtxRangePop(); } };

int run_build(armour_handle* h) {   // kernels only; inputs already on the device
    NvtxRange range("armour_build: reach sets + half-space tables");
    const int n_work = h->count * h->T;
    Tables tb = h->tb;
    tb.P = h->count; tb.n_obs = h->n_obs; tb.mode = h->mode; tb.jrs = h->d_jrs; tb.k_range_in = h->d_krange;
    for (int attempt = 0; attempt < 6; attempt++) {
        tb.u_keys = h->tb.u_keys; tb.u_coef = h->tb.u_coef; tb.l_keys = h->tb.l_keys; tb.l_coef = h->tb.l_coef; tb.ucap = h->tb.ucap; tb.lcap = h->tb.lcap;
        CU(cudaMemsetAsync(h->d_err, 0, 3 * sizeof(int), h->stream));   // error word, work counter, guard words verified
        CU(cudaEventRecord(h->ev[0], h->stream));
        // opt-in experiment: stage D in the tail of each interval's CTA (no second launch) when its staging fits the CTA's sort buffers
        tb.fuse_planes = (h->fuse_planes && (size_t)(NJ * 18 + 12 * h->n_obs) * sizeof(double) <= (size_t)h->scap * 20) ? 1 : 0;
        CU(launch_reach_build(tb, h->arena, h->arena_stride, h->mcap, h->ncap, h->scap, h->tcap, n_work, std::min(h->grid, n_work), h->nt, h->minb, h->groups, h->stream));
        CU(cudaEventRecord(h->ev[1], h->stream));
        if (!tb.fuse_planes) CU(launch_hyperplanes(tb, h->stream));
        CU(cudaEventRecord(h->ev[2], h->stream));
        h->launches += (h->n_obs > 0 && !tb.fuse_planes) ? 2 : 1;
        CU(cudaMemcpyAsync(h->h_err, h->d_err, 3 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(h->h_torque_radius, h->tb.torque_radius, sizeof(double) * (size_t)h->count * h->T * NF, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        cudaEventElapsedTime(&h->reach_ms, h->ev[0], h->ev[1]);
        cudaEventElapsedTime(&h->hyper_ms, h->ev[1], h->ev[2]);
        cudaEventElapsedTime(&h->build_ms, h->ev[0], h->ev[2]);
        const int err = *h->h_err;
        if (err == 0) { h->built = true; h->have_eval = false; h->mirror_valid = false; return ARMOUR_OK; }
        if (err & 16) return fail(ARMOUR_E_NUMERIC, "reach-set build: a monomial degree outgrew its key field (more than 3 for k / cos / sin error symbols, more than 1 for the others; KPR/PZsparse.h:23-40)");
        if (err & 8) return fail(ARMOUR_E_NUMERIC, "reach-set build: a link PZ has a pure link generator other than the three box generators");
        if (err & 128) return fail(ARMOUR_E_CUDA, "reach-set build: an arena guard word was overwritten (debug build, ARMOUR_ARENA_CANARY)");
        if (err & 32) return fail(ARMOUR_E_CUDA, "reach-set build: a hand-off between the two thread groups of a CTA timed out (internal error)");
        // a capacity was exceeded: grow what overflowed and retry (documented in armour_b200.h)
        if ((err & 1) && h->ncap >= 65534) return fail(ARMOUR_E_CAPACITY, "an operation has more than 65535 candidate monomials");
        if (err & (4 | 64)) {
            if (err & 4) h->ucap *= 2;
            if (err & 64) h->lcap *= 2;
            h->mirror_valid = false;
            int rc = alloc_konly_tables(h);
            if (rc != ARMOUR_OK) return rc;
        }
        if (err & 1) h->ncap = std::min(h->ncap * 2, 65534);
        if (err & 2) h->mcap *= 2;
        if (h->ncap < 2 * h->mcap) h->ncap = std::min(2 * h->mcap, 65534);
        if (err & (1 | 2)) {
            int rc = alloc_arena(h);
            if (rc != ARMOUR_OK) return rc;
        }
    }
    return fail(ARMOUR_E_CAPACITY, "monomial capacities exceeded after retries");
}

// One fused eval_g + eval_jac_g launch.
//   hg / hj == nullptr: device-resident evaluation into d_g / d_jac (bench.py's `value` timing).
//   otherwise hg / hj are PAGE-LOCKED host destinations (the handle's staging buffers or registered caller arrays) and
//   ag / aj their device aliases.  How the 1.2 MB of results cross PCIe is h->host_write:
//     0  the kernel stores g and the Jacobian straight into host memory (zero-copy)                        [default]
//     1  the kernel writes device buffers, two copy-engine transfers follow in stream order
//     2  g (8 m bytes) zero-copy from the kernel, the Jacobian (56 m bytes) by one copy-engine transfer
//   Measured on the B200 box (T = 128, 20 obstacles, in-library wall clock per call, profiles/r2_eval_latency.md):
//   0: 42.8 us, 1: 54.1 us, 2: 48.7 us — a stand-alone copy-engine transfer of 1.2 MB is ~20 % faster than SM stores
//   (22 vs 26 us), but each copy adds ~3 us of start-up in stream order and nothing overlaps the kernel.
//   Completion: a 64-bit sequence number written to a mapped word in stream order (cuStreamWriteValue64; when the driver
//   entry point is missing, by the kernel's last block after a system-wide fence); the host spins on that word instead of
//   calling cudaStreamSynchronize, which saves the driver's wake-up latency on the per-iteration path.
int launch_eval(armour_handle* h, const double* x, double* hg, double* hj, double* ag, double* aj, int what = 3) {
    if (!h->built) return fail(ARMOUR_E_STATE, "eval before build");
    NvtxRange range("armour_eval: constraints + Jacobian");
    const auto t_begin = std::chrono::steady_clock::now();
    if (x) memcpy(h->h_x, x, sizeof(double) * NF);
    Tables tb = h->tb;
    tb.P = h->count; tb.n_obs = h->n_obs; tb.mode = h->mode; tb.jrs = h->d_jrs; tb.k_range_in = h->d_krange;
    const bool host_visible = hg != nullptr || hj != nullptr;
    const int mode = host_visible ? h->host_write : -1;
    const bool timed = h->time_kernels || !host_visible;   // events are recorded here, read lazily by armour_last_eval_ms
    if (timed) CU(cudaEventRecord(h->ev[3], h->stream));
    const unsigned long long seq = ++h->eval_seq;
    const bool flag_in_kernel = host_visible && !h->write_value64 && mode == 0;
    double* kg = !host_visible ? h->d_g : (mode == 1 ? h->d_g : ag);
    double* kj = !host_visible ? h->d_jac : (mode == 0 ? aj : h->d_jac);
    CU(launch_constraint_eval(tb, h->sel, h->h_x, kg, kj, h->d_link_center, what, h->d_done, flag_in_kernel ? h->d_done_flag : nullptr, seq, host_visible ? h->eval_bps_host : 0, h->stream));
    h->launches += 1;
    if (host_visible) {
        const size_t m = (size_t)m_of(h);
        if (mode == 1 && (what & 1)) CU(cudaMemcpyAsync(hg, h->d_g, sizeof(double) * m, cudaMemcpyDeviceToHost, h->stream));
        if (mode >= 1 && (what & 2)) CU(cudaMemcpyAsync(hj, h->d_jac, sizeof(double) * m * NF, cudaMemcpyDeviceToHost, h->stream));
    }
    if (timed) CU(cudaEventRecord(h->ev[4], h->stream));
    if (host_visible) {
        bool polled = false;
        if (h->write_value64) {
            if (h->write_value64(h->stream, (unsigned long long)(uintptr_t)h->d_done_flag, seq, 0) != 0) return fail(ARMOUR_E_CUDA, "cuStreamWriteValue64 failed");
            polled = true;
        }
        else if (flag_in_kernel) polled = true;
        bool done = false;
        if (polled) {   // bounded spin (about 50 ms), then fall back to the stream so that a faulted kernel surfaces as an error, not a hang
            for (long spin = 0; spin < 20000000L; spin++) {
                if (*h->h_done == seq) { done = true; break; }
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
        }
        if (!done) {
            CU(cudaStreamSynchronize(h->stream));
            if (polled && *h->h_done != seq) return fail(ARMOUR_E_CUDA, "constraint evaluation did not signal completion");
        }
    }
    else CU(cudaStreamSynchronize(h->stream));
    h->eval_timed = timed;
    h->eval_host_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count();
    return ARMOUR_OK;
}
int run_eval(armour_handle* h, const double* x, bool to_host) {
    int rc = to_host ? launch_eval(h, x, h->h_g, h->h_jac, h->a_g, h->a_jac) : launch_eval(h, x, nullptr, nullptr, nullptr, nullptr);
    if (rc != ARMOUR_OK) return rc;
    if (x && to_host) { memcpy(h->last_x, x, sizeof(double) * NF); h->have_eval = true; h->staged = 3; }
    return ARMOUR_OK;
}

// Device alias of a caller array under cfg.pin_user_buffers, registering it on first sight.  `keep` is an alias resolved
// earlier in the same call: its entry is never the one evicted.  A cached entry is re-validated with the driver before use
// (cudaHostGetDevicePointer fails once the registration is gone); what the driver cannot see is a caller that frees a
// registered array WITHOUT armour_release_host_buffers and gets the same address back from the allocator — the header
// states that contract.
void* pinned_alias(armour_handle* h, void* host, size_t bytes, const void* keep = nullptr) {
    for (size_t i = 0; i < h->pinned.size(); i++) {
        auto& p = h->pinned[i];
        if (p.host != host) continue;
        void* dev = nullptr;
        if (p.bytes >= bytes && cudaHostGetDevicePointer(&dev, host, 0) == cudaSuccess && dev == p.dev) {
            if (i + 1 != h->pinned.size()) { auto e = p; h->pinned.erase(h->pinned.begin() + i); h->pinned.push_back(e); }   // most recently used last
            return dev;
        }
        cudaGetLastError();
        cudaHostUnregister(host); cudaGetLastError();
        h->pinned.erase(h->pinned.begin() + i);
        break;
    }
    if (h->pinned.size() >= 8) {   // evict the least recently used entry that this call does not depend on
        size_t victim = 0;
        while (victim < h->pinned.size() && h->pinned[victim].dev == keep) victim++;
        if (victim == h->pinned.size()) return nullptr;
        cudaHostUnregister((void*)h->pinned[victim].host); cudaGetLastError();
        h->pinned.erase(h->pinned.begin() + victim);
    }
    static const bool debug = getenv("ARMOUR_DEBUG_PINNING") != nullptr;
    cudaError_t e = cudaHostRegister(host, bytes, cudaHostRegisterMapped);
    if (e != cudaSuccess) { if (debug) fprintf(stderr, "armour: cudaHostRegister(%p, %zu): %s\n", host, bytes, cudaGetErrorString(e)); cudaGetLastError(); return nullptr; }
    void* dev = nullptr;
    e = cudaHostGetDevicePointer(&dev, host, 0);
    if (e != cudaSuccess) { if (debug) fprintf(stderr, "armour: cudaHostGetDevicePointer(%p): %s\n", host, cudaGetErrorString(e)); cudaGetLastError(); cudaHostUnregister(host); return nullptr; }
    h->pinned.push_back({host, bytes, dev});
    return dev;
}

int ensure_mirror(armour_handle* h) {
    if (h->mirror_valid) return ARMOUR_OK;
    if (!h->built) return fail(ARMOUR_E_STATE, "tables requested before build");
    const size_t T = h->T, p = h->sel;
    h->m_traj.resize(T * TRAJ_TABLES * NJ);
    h->m_un.resize(T * NF); h->m_ln.resize(T * NJ);
    const size_t UCAP = h->tb.ucap, LCAP = h->tb.lcap;
    h->m_ukeys.resize(T * NF * UCAP); h->m_ucoef.resize(T * NF * UCAP); h->m_ucenter.resize(T * NF); h->m_uind.resize(T * NF); h->m_dist.resize(T * NF);
    h->m_lkeys.resize(T * NJ * LCAP); h->m_lcoef.resize(T * NJ * 3 * LCAP); h->m_lcenter.resize(T * NJ * 3); h->m_lind.resize(T * NJ * 3);
    const Tables& tb = h->tb;
    if (tb.traj) CU(cudaMemcpy(h->m_traj.data(), tb.traj + p * T * TRAJ_TABLES * NJ, sizeof(SmallRec) * h->m_traj.size(), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h->m_un.data(), tb.u_n + p * T * NF, sizeof(int) * T * NF, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h->m_ln.data(), tb.l_n + p * T * NJ, sizeof(int) * T * NJ, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h->m_ukeys.data(), tb.u_keys + p * T * NF * UCAP, 8 * h->m_ukeys.size(), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h->m_ucoef.data(), tb.u_coef + p * T * NF * UCAP, 8 * h->m_ucoef.size(), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h->m_ucenter.data(), tb.u_center + p * T * NF, 8 * T * NF, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h->m_uind.data(), tb.u_ind + p * T * NF, 8 * T * NF, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h->m_dist.data(), tb.dist_rad + p * T * NF, 8 * T * NF, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h->m_lkeys.data(), tb.l_keys + p * T * NJ * LCAP, 8 * h->m_lkeys.size(), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h->m_lcoef.data(), tb.l_coef + p * T * NJ * 3 * LCAP, 8 * h->m_lcoef.size(), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h->m_lcenter.data(), tb.l_center + p * T * NJ * 3, 8 * T * NJ * 3, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h->m_lind.data(), tb.l_ind + p * T * NJ * 3, 8 * T * NJ * 3, cudaMemcpyDeviceToHost));
    h->mirror_valid = true;
    return ARMOUR_OK;
}

}  // namespace

extern "C" {

const char* armour_last_error(void) { return g_last_error.c_str(); }

void armour_default_config(armour_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->num_time_steps = 128;
    for (int i = 0; i < 7; i++) cfg->k_range[i] = M_PI / 48;
    cfg->mass_uncertainty = 0.03;
    cfg->inertia_uncertainty = 0.03;
    cfg->simplify_threshold = 5e-4;
    cfg->max_obstacles = 40;
    cfg->max_monomials = 1024;
    cfg->max_entries = 8192;
    cfg->threads_per_cta = 256;
    cfg->device = -1;
    cfg->batch = 1;
}

int armour_create(const armour_config* cfg_in, armour_handle** out) {
    if (!cfg_in || !out) return fail(ARMOUR_E_INVALID, "null argument");
    armour_config cfg = *cfg_in;
    if (cfg.num_time_steps <= 0 || (cfg.num_time_steps & 1)) return fail(ARMOUR_E_INVALID, "num_time_steps must be a positive even number");
    if (cfg.max_obstacles < 0) return fail(ARMOUR_E_INVALID, "max_obstacles < 0");
    if (cfg.max_obstacles > eval_max_obstacles()) return fail(ARMOUR_E_INVALID, "max_obstacles above 64 (the constraint kernel stages one link's half-space slab in shared memory; the reference's MAX_OBSTACLE_NUM is 40)");
    if (cfg.max_monomials <= 0) cfg.max_monomials = 1024;
    if (cfg.max_entries <= 0) cfg.max_entries = 8192;
    if (cfg.batch <= 0) cfg.batch = 1;
    if (const char* e = getenv("ARMOUR_TUNE_NT")) cfg.threads_per_cta = atoi(e);
    const int nt_in = cfg.threads_per_cta;
    const bool nt_default = nt_in != 32 && nt_in != 64 && nt_in != 128 && nt_in != 256 && nt_in != 512;
    if (nt_default) cfg.threads_per_cta = cfg.batch > 1 ? 128 : 256;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(ARMOUR_E_CUDA, "no CUDA device: this library has no CPU fallback");
    armour_handle* h = new armour_handle();
    h->cfg = cfg;
    bool no_structured_env = false;
    bool fuse_planes_env = false;   // measured: no gain for one plan (the ~20 us move into the reach kernel's tail), -4.5 % on the sweep
    if (cfg.device >= 0) { if (cudaSetDevice(cfg.device) != cudaSuccess) { delete h; return fail(ARMOUR_E_CUDA, "cudaSetDevice failed"); } }
    cudaGetDevice(&h->device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, h->device) != cudaSuccess) { delete h; return fail(ARMOUR_E_CUDA, "cudaGetDeviceProperties failed"); }
    h->sm_count = prop.multiProcessorCount;
    h->P = cfg.batch; h->T = cfg.num_time_steps; h->max_obs = cfg.max_obstacles;
    h->mcap = cfg.max_monomials; h->ncap = std::min(cfg.max_entries, 65534) & ~1; h->nt = cfg.threads_per_cta;
    // register budget: one plan is latency-bound (1 CTA/SM, all registers); a batch wants more resident CTAs
    // (measured, scripts/tune_sweep.py: 128 threads per interval with 1408-entry shared sort buffers is the fastest sweep shape;
    // the narrow shapes — 64 / 32 threads per interval, 8-24 CTAs per SM — run 12 % / 48 % slower)
    const int nt = cfg.threads_per_cta;
    // sweep shape: 128 threads x 4 CTAs per SM (128 registers, 16 resident warps).  With the round-1 operation code 3 CTAs at 168
    // registers were 2 % ahead (the sweep is bound by instruction fetch, not by resident warps); with the unified 3-vector
    // operation 4 CTAs lead by 4.5 % (profiles/README.md, round-2 experiments)
    h->minb = nt == 32 ? 16 : nt == 64 ? 8 : cfg.batch > 1 ? (nt == 128 ? 4 : 2) : 1;
    if (cfg.batch > 1 && nt == 128) { h->scap = 1536; h->tcap = 300; }   // 1536: the largest operations of a heavy interval (~1400-1500 candidates) still sort in shared memory
    if (nt == 64) { h->scap = 768; h->tcap = 192; }
    if (nt == 32) { h->scap = 512; h->tcap = 128; }
    if (const char* e = getenv("ARMOUR_TUNE_MINB")) h->minb = atoi(e);
    if (const char* e = getenv("ARMOUR_TUNE_MCAP")) h->mcap = std::max(64, atoi(e));
    if (const char* e = getenv("ARMOUR_TUNE_EVAL_BPS")) h->eval_bps_host = atoi(e);
    if (const char* e = getenv("ARMOUR_TUNE_NO_STRUCTURED")) no_structured_env = atoi(e) != 0;
    if (const char* e = getenv("ARMOUR_TUNE_FUSE_PLANES")) fuse_planes_env = atoi(e) != 0;
    if (const char* e = getenv("ARMOUR_TUNE_HOST_WRITE")) h->host_write = std::min(2, std::max(0, atoi(e)));
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        const char* off = getenv("ARMOUR_TUNE_FLAG_IN_KERNEL");
        if (!(off && atoi(off)) && cudaGetDriverEntryPoint("cuStreamWriteValue64", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess && fn)
            h->write_value64 = (int (*)(cudaStream_t, unsigned long long, unsigned long long, unsigned))fn;
        cudaGetLastError();
    }
    if (const char* e = getenv("ARMOUR_TUNE_UCAP")) h->ucap = std::max(1, atoi(e));   // tests force the grow-and-retry path with tiny tables
    if (const char* e = getenv("ARMOUR_TUNE_LCAP")) h->lcap = std::max(1, atoi(e));
    if (const char* e = getenv("ARMOUR_TUNE_NCAP")) h->ncap = std::min(std::max(256, atoi(e)), 65534) & ~1;
    // one plan (latency): two thread groups per CTA; a batch (throughput): one group per CTA and several resident CTAs per SM
    h->groups_cfg = (cfg.batch > 1 || nt != 256) ? 1 : 2;
    if (const char* e = getenv("ARMOUR_TUNE_GROUPS")) h->groups_cfg = atoi(e) == 2 ? 2 : 1;
    if (const char* e = getenv("ARMOUR_TUNE_SCAP")) h->scap = std::max(128, atoi(e)) & ~1;
    if (const char* e = getenv("ARMOUR_TUNE_TCAP")) h->tcap = std::max(32, atoi(e)) & ~1;
    kinova_model(h->model);
    *out = h;   // so that armour_destroy can clean up after a partial failure
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto& e : h->ev) CU(cudaEventCreate(&e));
    CU(upload_robot_model(h->model));
    const size_t P = h->P, T = h->T, O = std::max(h->max_obs, 1);
    Tables& tb = h->tb;
    memset(&tb, 0, sizeof(tb));
    tb.T = h->T; tb.P = h->P; tb.n_obs = 0;
    for (int i = 0; i < NF; i++) tb.k_range[i] = cfg.k_range[i];
    tb.mass_unc = cfg.mass_uncertainty; tb.inertia_unc = cfg.inertia_uncertainty; tb.thr = cfg.simplify_threshold;
    tb.no_structured = no_structured_env ? 1 : 0;
    h->fuse_planes = fuse_planes_env;
    if (const char* e = getenv("ARMOUR_TUNE_STATIC_STRIDE")) tb.static_stride = atoi(e) != 0;
    CU(dalloc(&h->d_state, P * 21)); CU(dalloc(&h->d_obs, P * O * 12));
    CU(dalloc(&h->d_jrs, (size_t)6 * NF * T)); CU(dalloc(&h->d_krange, (size_t)NF));
    CU(cudaMallocHost((void**)&h->h_jrs, sizeof(double) * 6 * NF * T)); CU(cudaMallocHost((void**)&h->h_krange, sizeof(double) * NF));
    tb.state = h->d_state; tb.obstacles = h->d_obs;
    const bool want_traj = cfg.export_trajectory_tables > 0 || (cfg.export_trajectory_tables == 0 && cfg.batch == 1);
    if (want_traj) CU(dalloc(&tb.traj, P * T * TRAJ_TABLES * NJ));
    CU(dalloc(&tb.cos_rem, P * NJ * T * 2)); CU(dalloc(&tb.sin_rem, P * NJ * T * 2));
    CU(dalloc(&tb.u_n, P * T * NF));
    CU(dalloc(&tb.u_center, P * T * NF)); CU(dalloc(&tb.u_ind, P * T * NF)); CU(dalloc(&tb.dist_rad, P * T * NF)); CU(dalloc(&tb.torque_radius, P * T * NF));
    CU(dalloc(&tb.l_n, P * T * NJ));
    { int rc = alloc_konly_tables(h); if (rc != ARMOUR_OK) return rc; }
    CU(dalloc(&tb.l_center, P * T * NJ * 3)); CU(dalloc(&tb.l_ind, P * T * NJ * 3)); CU(dalloc(&tb.gens, P * T * NJ * 18));
    CU(dalloc(&tb.A, P * T * NJ * O * COMB * 3)); CU(dalloc(&tb.d, P * T * NJ * O * COMB)); CU(dalloc(&tb.delta, P * T * NJ * O * COMB));
    CU(dalloc(&h->d_err, 3)); tb.err = h->d_err;
    const size_t mmax = NF * T + NJ * T * O + NF * 4;
    CU(dalloc(&h->d_x, NF)); CU(dalloc(&h->d_g, mmax)); CU(dalloc(&h->d_jac, mmax * NF)); CU(dalloc(&h->d_link_center, T * NJ * 3));
    CU(cudaMallocHost((void**)&h->h_state, sizeof(double) * P * 21)); CU(cudaMallocHost((void**)&h->h_obs, sizeof(double) * P * O * 12));
    CU(cudaMallocHost((void**)&h->h_x, sizeof(double) * NF)); CU(cudaMallocHost((void**)&h->h_g, sizeof(double) * mmax));
    CU(cudaMallocHost((void**)&h->h_jac, sizeof(double) * mmax * NF)); CU(cudaMallocHost((void**)&h->h_torque_radius, sizeof(double) * P * T * NF));
    CU(cudaMallocHost((void**)&h->h_err, 3 * sizeof(int)));
    CU(cudaHostGetDevicePointer((void**)&h->a_g, h->h_g, 0)); CU(cudaHostGetDevicePointer((void**)&h->a_jac, h->h_jac, 0));
    CU(dalloc(&h->d_done, 1)); CU(cudaMemset(h->d_done, 0, sizeof(unsigned)));
    CU(cudaHostAlloc((void**)&h->h_done, sizeof(unsigned long long), cudaHostAllocMapped));
    *h->h_done = 0;
    CU(cudaHostGetDevicePointer((void**)&h->d_done_flag, (void*)h->h_done, 0));
    int rc = alloc_arena(h);
    if (rc != ARMOUR_OK) return rc;
    return ARMOUR_OK;
}

void armour_destroy(armour_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (auto& p : h->pinned) cudaHostUnregister((void*)p.host);
    cudaGetLastError();
    Tables& tb = h->tb;
    void* dev[] = {h->d_jrs, h->d_krange, h->d_state, h->d_obs, tb.traj, tb.cos_rem, tb.sin_rem, tb.u_n, tb.u_keys, tb.u_coef, tb.u_center, tb.u_ind, tb.dist_rad, tb.torque_radius,
                   tb.l_n, tb.l_keys, tb.l_coef, tb.l_center, tb.l_ind, tb.gens, tb.A, tb.d, tb.delta, h->d_err, h->d_x, h->d_g, h->d_jac, h->d_link_center, h->arena, h->bin_buf, h->d_done, h->d_bx, h->d_bg, h->d_bjac};
    for (void* p : dev) if (p) cudaFree(p);
    void* pinned[] = {h->h_jrs, h->h_krange, h->h_state, h->h_obs, h->h_x, h->h_g, h->h_jac, h->h_torque_radius, h->h_err, (void*)h->h_done, h->h_bx, h->h_ws};
    for (void* p : pinned) if (p) cudaFreeHost(p);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int armour_upload_problems(armour_handle* h, int count, const double* q0, const double* qd0, const double* qdd0, const double* obstacles, int n_obs) {
    if (!h || !q0 || !qd0 || !qdd0) return fail(ARMOUR_E_INVALID, "null argument");
    if (count < 1 || count > h->P) return fail(ARMOUR_E_INVALID, "count outside [1, cfg.batch]");
    if (n_obs < 0 || n_obs > h->max_obs) return fail(ARMOUR_E_INVALID, "number of obstacles outside [0, max_obstacles]");   // KPR/armour_main.cu:66-72
    if (n_obs > 0 && !obstacles) return fail(ARMOUR_E_INVALID, "null obstacles");
    CU(cudaSetDevice(h->device));
    for (int p = 0; p < count; p++)
        for (int i = 0; i < NF; i++) { h->h_state[p * 21 + i] = q0[p * NF + i]; h->h_state[p * 21 + 7 + i] = qd0[p * NF + i]; h->h_state[p * 21 + 14 + i] = qdd0[p * NF + i]; }
    if (n_obs > 0) memcpy(h->h_obs, obstacles, sizeof(double) * (size_t)count * n_obs * 12);
    h->count = count; h->n_obs = n_obs; h->sel = 0; h->built = false; h->have_eval = false; h->mirror_valid = false; h->mode = 0;
    CU(cudaMemcpyAsync(h->d_state, h->h_state, sizeof(double) * (size_t)count * 21, cudaMemcpyHostToDevice, h->stream));
    if (n_obs > 0) CU(cudaMemcpyAsync(h->d_obs, h->h_obs, sizeof(double) * (size_t)count * n_obs * 12, cudaMemcpyHostToDevice, h->stream));
    return ARMOUR_OK;
}
int armour_build_resident(armour_handle* h) {
    if (!h || h->count < 1) return fail(ARMOUR_E_STATE, "no problems uploaded");
    CU(cudaSetDevice(h->device));
    return run_build(h);
}
int armour_build_batch(armour_handle* h, int count, const double* q0, const double* qd0, const double* qdd0, const double* obstacles, int n_obs) {
    int rc = armour_upload_problems(h, count, q0, qd0, qdd0, obstacles, n_obs);
    if (rc != ARMOUR_OK) return rc;
    return run_build(h);
}
int armour_build(armour_handle* h, const double* q0, const double* qd0, const double* qdd0, const double* obstacles, int n_obs) {
    return armour_build_batch(h, 1, q0, qd0, qdd0, obstacles, n_obs);
}
int armour_build_armtd(armour_handle* h, const double* q0, const double* qd0, const double* jrs, const double* k_range, const double* obstacles, int n_obs) {
    if (!h || !q0 || !qd0 || !jrs || !k_range) return fail(ARMOUR_E_INVALID, "null argument");
    const double zero[NF] = {0, 0, 0, 0, 0, 0, 0};
    int rc = armour_upload_problems(h, 1, q0, qd0, zero, obstacles, n_obs);
    if (rc != ARMOUR_OK) return rc;
    h->mode = 1;
    memcpy(h->h_jrs, jrs, sizeof(double) * 6 * NF * h->T);
    memcpy(h->h_krange, k_range, sizeof(double) * NF);
    CU(cudaMemcpyAsync(h->d_jrs, h->h_jrs, sizeof(double) * 6 * NF * h->T, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_krange, h->h_krange, sizeof(double) * NF, cudaMemcpyHostToDevice, h->stream));
    return run_build(h);
}
int armour_select_problem(armour_handle* h, int p) {
    if (!h || !h->built) return fail(ARMOUR_E_STATE, "select before build");
    if (p < 0 || p >= h->count) return fail(ARMOUR_E_INVALID, "problem index out of range");
    h->sel = p; h->have_eval = false; h->mirror_valid = false;
    return ARMOUR_OK;
}

int armour_get_nlp_info(armour_handle* h, int* n, int* m, int* nnz_jac_g, int* nnz_h_lag) {
    if (!h || !n || !m || !nnz_jac_g || !nnz_h_lag) return fail(ARMOUR_E_INVALID, "null argument");
    *n = NF; *m = m_of(h); *nnz_jac_g = *m * NF; *nnz_h_lag = 0;
    return ARMOUR_OK;
}
int armour_get_bounds_info(armour_handle* h, double* x_l, double* x_u, double* g_l, double* g_u) {
    if (!h || !x_l || !x_u || !g_l || !g_u) return fail(ARMOUR_E_INVALID, "null argument");
    if (!h->built) return fail(ARMOUR_E_STATE, "bounds before build");
    const RobotModel& rm = h->model;
    const int T = h->T;
    const double* tr = h->h_torque_radius + (size_t)h->sel * T * NF;
    for (int i = 0; i < NF; i++) { x_l[i] = -1.0; x_u[i] = 1.0; }
    int offset = 0;
    if (h->mode == 0) {
        for (int i = 0; i < T; i++)
            for (int j = 0; j < NF; j++) { g_l[i * NF + j] = -rm.torque[j] + tr[i * NF + j]; g_u[i * NF + j] = rm.torque[j] - tr[i * NF + j]; }
        offset += NF * T;
    }
    for (int i = offset; i < offset + T * NJ * h->n_obs; i++) { g_l[i] = -1e19; g_u[i] = 0; }
    offset += T * NJ * h->n_obs;
    for (int rep = 0; rep < 2; rep++) { for (int i = 0; i < NF; i++) { g_l[offset + i] = rm.state_lb[i] + rm.qe; g_u[offset + i] = rm.state_ub[i] - rm.qe; } offset += NF; }
    for (int rep = 0; rep < 2; rep++) { for (int i = 0; i < NF; i++) { g_l[offset + i] = -rm.speed[i] + rm.qde; g_u[offset + i] = rm.speed[i] - rm.qde; } offset += NF; }
    return ARMOUR_OK;
}
int armour_get_starting_point(armour_handle* h, double* x) {
    if (!h || !x) return fail(ARMOUR_E_INVALID, "null argument");
    for (int i = 0; i < NF; i++) x[i] = 0.0;
    return ARMOUR_OK;
}
int armour_eval_f(armour_handle* h, const double* q_des, double t_plan, const double* x, double* obj_value) {
    if (!h || !q_des || !x || !obj_value) return fail(ARMOUR_E_INVALID, "null argument");
    if (h->count < 1) return fail(ARMOUR_E_STATE, "no problem");
    const double* st = h->h_state + (size_t)h->sel * 21;
    double qp[NF];
    for (int i = 0; i < NF; i++)
        qp[i] = h->mode == 0 ? q_des_host(st[i], st[7 + i], st[14 + i], h->cfg.k_range[i] * x[i], t_plan)
                             : st[i] + st[7 + i] * 0.5 + h->h_krange[i] * x[i] * 0.125;   // KPA/NLPclass.cu:197
    double v = pow(wrap_to_pi(q_des[0] - qp[0]), 2) + pow(wrap_to_pi(q_des[2] - qp[2]), 2) + pow(wrap_to_pi(q_des[4] - qp[4]), 2) + pow(wrap_to_pi(q_des[6] - qp[6]), 2) +
               pow(q_des[1] - qp[1], 2) + pow(q_des[3] - qp[3], 2) + pow(q_des[5] - qp[5], 2);
    *obj_value = v * COST_SCALE;
    return ARMOUR_OK;
}
int armour_eval_grad_f(armour_handle* h, const double* q_des, double t_plan, const double* x, double* grad_f) {
    if (!h || !q_des || !x || !grad_f) return fail(ARMOUR_E_INVALID, "null argument");
    if (h->count < 1) return fail(ARMOUR_E_STATE, "no problem");
    const double* st = h->h_state + (size_t)h->sel * 21;
    for (int i = 0; i < NF; i++) {
        const double qp = h->mode == 0 ? q_des_host(st[i], st[7 + i], st[14 + i], h->cfg.k_range[i] * x[i], t_plan)
                                       : st[i] + st[7 + i] * 0.5 + h->h_krange[i] * x[i] * 0.125;
        const double dk = h->mode == 0 ? pow(t_plan, 3) * (6 * pow(t_plan, 2) - 15 * t_plan + 10) * h->cfg.k_range[i] : h->h_krange[i] * 0.125;   // KPA/NLPclass.cu:229-230
        grad_f[i] = (i % 2 == 0) ? (2 * wrap_to_pi(qp - q_des[i]) * dk) : (2 * (qp - q_des[i]) * dk);
        grad_f[i] *= COST_SCALE;
    }
    return ARMOUR_OK;
}
int armour_pinned_buffer_count(armour_handle* h, int* count) {
    if (!h || !count) return fail(ARMOUR_E_INVALID, "null argument");
    *count = (int)h->pinned.size();
    return ARMOUR_OK;
}
int armour_release_host_buffers(armour_handle* h) {
    if (!h) return fail(ARMOUR_E_INVALID, "null argument");
    cudaSetDevice(h->device);
    for (auto& p : h->pinned) cudaHostUnregister((void*)p.host);
    h->pinned.clear();
    cudaGetLastError();
    return ARMOUR_OK;
}
// Ipopt asks for g and for the Jacobian in separate callbacks (eval_g at every trial point, eval_jac_g once per accepted
// iterate), so each entry point computes what it is asked for: `what` bit 0 = g, bit 1 = Jacobian.  With cfg.pin_user_buffers
// the kernel writes the caller's page-locked arrays directly; otherwise results pass through the handle's pinned staging
// buffers, and a repeated request at the same x is served from there.
// device alias of an address inside the handle's own page-locked solver workspace (nullptr otherwise)
static double* ws_alias(armour_handle* h, double* host, size_t bytes) {
    if (!h->h_ws || host < h->h_ws || (char*)host + bytes > (char*)(h->h_ws + h->ws_doubles)) return nullptr;
    return h->a_ws + (host - h->h_ws);
}
static int eval_into(armour_handle* h, const double* x, double* g, double* values) {
    const int what = (g ? 1 : 0) | (values ? 2 : 0);
    if (!h->built) return fail(ARMOUR_E_STATE, "eval before build");
    const int m = m_of(h);
    const bool in_ws = (!g || ws_alias(h, g, sizeof(double) * m)) && (!values || ws_alias(h, values, sizeof(double) * (size_t)m * NF));
    if (h->cfg.pin_user_buffers || in_ws) {   // zero staging: the kernel writes the caller's arrays
        double* dg = !g ? nullptr : in_ws ? ws_alias(h, g, sizeof(double) * m) : (double*)pinned_alias(h, g, sizeof(double) * m);
        double* dj = !values ? nullptr : in_ws ? ws_alias(h, values, sizeof(double) * (size_t)m * NF) : (double*)pinned_alias(h, values, sizeof(double) * (size_t)m * NF, dg);
        if ((!g || dg) && (!values || dj)) {
            int rc = launch_eval(h, x, g, values, dg, dj, what);
            if (rc != ARMOUR_OK) return rc;
            h->have_eval = true; h->staged = 0;      // link_sliced_center is current, nothing is staged
            memcpy(h->last_x, x, sizeof(double) * NF);
            return ARMOUR_OK;
        }
    }
    const bool same_x = h->have_eval && memcmp(h->last_x, x, sizeof(double) * NF) == 0;
    const int missing = same_x ? (what & ~h->staged) : what;
    if (missing) {
        int rc = launch_eval(h, x, (missing & 1) ? h->h_g : nullptr, (missing & 2) ? h->h_jac : nullptr, h->a_g, h->a_jac, missing);
        if (rc != ARMOUR_OK) return rc;
        h->staged = same_x ? (h->staged | missing) : missing;
        memcpy(h->last_x, x, sizeof(double) * NF);
        h->have_eval = true;
    }
    if (g) memcpy(g, h->h_g, sizeof(double) * m);
    if (values) memcpy(values, h->h_jac, sizeof(double) * (size_t)m * NF);
    return ARMOUR_OK;
}
int armour_eval_g_jac(armour_handle* h, const double* x, double* g, double* values) {
    if (!h || !x || (!g && !values)) return fail(ARMOUR_E_INVALID, "null argument");
    CU(cudaSetDevice(h->device));
    return eval_into(h, x, g, values);
}
int armour_eval_g(armour_handle* h, const double* x, double* g) { if (!g) return fail(ARMOUR_E_INVALID, "null argument"); return armour_eval_g_jac(h, x, g, nullptr); }
int armour_eval_jac_g(armour_handle* h, const double* x, double* values) { if (!values) return fail(ARMOUR_E_INVALID, "null argument"); return armour_eval_g_jac(h, x, nullptr, values); }
int armour_eval_resident(armour_handle* h, const double* x) {
    if (!h) return fail(ARMOUR_E_INVALID, "null argument");
    CU(cudaSetDevice(h->device));
    return run_eval(h, x, false);
}
int armour_eval_resident_burst(armour_handle* h, const double* x, int launches, float* ms_per_launch) {
    if (!h || !ms_per_launch || launches < 1) return fail(ARMOUR_E_INVALID, "bad argument");
    if (!h->built) return fail(ARMOUR_E_STATE, "eval before build");
    CU(cudaSetDevice(h->device));
    if (x) memcpy(h->h_x, x, sizeof(double) * NF);
    Tables tb = h->tb;
    tb.P = h->count; tb.n_obs = h->n_obs; tb.mode = h->mode; tb.jrs = h->d_jrs; tb.k_range_in = h->d_krange;
    CU(launch_constraint_eval(tb, h->sel, h->h_x, h->d_g, h->d_jac, h->d_link_center, 3, h->d_done, nullptr, 0, 0, h->stream));   // warm
    CU(cudaEventRecord(h->ev[3], h->stream));
    for (int i = 0; i < launches; i++)
        CU(launch_constraint_eval(tb, h->sel, h->h_x, h->d_g, h->d_jac, h->d_link_center, 3, h->d_done, nullptr, 0, 0, h->stream));
    CU(cudaEventRecord(h->ev[4], h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->launches += launches + 1;
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]);
    *ms_per_launch = ms / launches;
    h->eval_timed = false;
    return ARMOUR_OK;
}
// Device -> caller array.  A direct copy fails ("invalid argument") when part of the destination is still page-locked by a
// stale registration — memory that some handle registered under cfg.pin_user_buffers and the caller freed without
// armour_release_host_buffers; in that case the rows go through the handle's own pinned staging buffer in pieces.
static int copy_out(armour_handle* h, double* dst, const double* src_dev, size_t bytes) {
    if (cudaMemcpy(dst, src_dev, bytes, cudaMemcpyDeviceToHost) == cudaSuccess) return ARMOUR_OK;
    cudaGetLastError();
    const size_t piece = sizeof(double) * (size_t)m_of(h) * NF;   // capacity of h_jac for the current problem shape
    for (size_t off = 0; off < bytes; off += piece) {
        const size_t nb = std::min(piece, bytes - off);
        CU(cudaMemcpy(h->h_jac, (const char*)src_dev + off, nb, cudaMemcpyDeviceToHost));
        memcpy((char*)dst + off, h->h_jac, nb);
    }
    return ARMOUR_OK;
}
// One launch for `count` problems of the last batch build, problem y evaluated at x[y][0..6]: the call a batched solver (or a
// sweep that steps all its solvers in lockstep) makes once per iteration.  Rows of problem first + y go to g + y * m and
// values + y * 7 m (either may be NULL).  With cfg.pin_user_buffers the kernel writes the caller's arrays; otherwise the rows
// pass through device buffers and one copy each.
static int eval_batch_impl(armour_handle* h, int first, int count, const double* x, double* g, double* values, bool jac_resident) {
    if (!h || !x || (!g && !values && !jac_resident)) return fail(ARMOUR_E_INVALID, "null argument");
    if (!h->built) return fail(ARMOUR_E_STATE, "eval before build");
    if (first < 0 || count < 1 || first + count > h->count) return fail(ARMOUR_E_INVALID, "problem range outside the last batch build");
    CU(cudaSetDevice(h->device));
    NvtxRange range("armour_eval_batch");
    const size_t m = (size_t)m_of(h), n = (size_t)count;
    const int what = (g ? 1 : 0) | ((values || jac_resident) ? 2 : 0);
    if (h->bx_cap < n) {
        if (h->d_bx) cudaFree(h->d_bx);
        if (h->h_bx) cudaFreeHost(h->h_bx);
        h->d_bx = nullptr; h->h_bx = nullptr; h->bx_cap = 0;
        CU(dalloc(&h->d_bx, n * NF));
        CU(cudaMallocHost((void**)&h->h_bx, sizeof(double) * n * NF));
        h->bx_cap = n;
    }
    double *kg = nullptr, *kj = nullptr;
    if (h->cfg.pin_user_buffers) {
        kg = g ? (double*)pinned_alias(h, g, sizeof(double) * m * n) : nullptr;
        kj = values ? (double*)pinned_alias(h, values, sizeof(double) * m * NF * n, kg) : nullptr;
    }
    const bool copy_g = g && !kg, copy_j = values && !kj;
    if (copy_g) {
        if (h->bg_cap < n * m) { if (h->d_bg) cudaFree(h->d_bg); h->d_bg = nullptr; h->bg_cap = 0; CU(dalloc(&h->d_bg, n * m)); h->bg_cap = n * m; }
        kg = h->d_bg;
    }
    if (copy_j || jac_resident) {
        if (h->bjac_cap < n * m * NF) { if (h->d_bjac) cudaFree(h->d_bjac); h->d_bjac = nullptr; h->bjac_cap = 0; CU(dalloc(&h->d_bjac, n * m * NF)); h->bjac_cap = n * m * NF; }
        kj = h->d_bjac;
    }
    memcpy(h->h_bx, x, sizeof(double) * n * NF);
    Tables tb = h->tb;
    tb.P = h->count; tb.n_obs = h->n_obs; tb.mode = h->mode; tb.jrs = h->d_jrs; tb.k_range_in = h->d_krange;
    CU(cudaMemcpyAsync(h->d_bx, h->h_bx, sizeof(double) * n * NF, cudaMemcpyHostToDevice, h->stream));
    CU(cudaEventRecord(h->ev[3], h->stream));
    CU(launch_constraint_eval_batch(tb, first, count, h->d_bx, kg, kj, nullptr, what, h->stream));
    CU(cudaEventRecord(h->ev[4], h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (copy_g) { int rc = copy_out(h, g, h->d_bg, sizeof(double) * n * m); if (rc != ARMOUR_OK) return rc; }
    if (copy_j) { int rc = copy_out(h, values, h->d_bjac, sizeof(double) * n * m * NF); if (rc != ARMOUR_OK) return rc; }
    h->launches += 1;
    cudaEventElapsedTime(&h->batch_eval_ms, h->ev[3], h->ev[4]);
    h->have_eval = false;          // the single-problem staging buffers and link_sliced_center were not touched
    h->bjac_rows = (jac_resident || copy_j) ? n : 0;
    return ARMOUR_OK;
}
int armour_eval_batch(armour_handle* h, int first, int count, const double* x, double* g, double* values) {
    if (!g && !values) return fail(ARMOUR_E_INVALID, "null argument");
    return eval_batch_impl(h, first, count, x, g, values, false);
}
// Same launch with the Jacobians left on the device (count x 7 m doubles, problem-major) for a device-side consumer; only the
// constraint values (8 m bytes per problem instead of 64 m) cross PCIe.  g may be NULL.
int armour_eval_batch_resident(armour_handle* h, int first, int count, const double* x, double* g) {
    return eval_batch_impl(h, first, count, x, g, nullptr, true);
}
int armour_batch_jacobian_device(armour_handle* h, const double** values_dev, int* problems) {
    if (!h || !values_dev || !problems) return fail(ARMOUR_E_INVALID, "null argument");
    *values_dev = h->bjac_rows ? h->d_bjac : nullptr;
    *problems = (int)h->bjac_rows;
    return ARMOUR_OK;
}
int armour_get_batch_jacobian(armour_handle* h, int y, double* values) {
    if (!h || !values) return fail(ARMOUR_E_INVALID, "null argument");
    if (y < 0 || (size_t)y >= h->bjac_rows) return fail(ARMOUR_E_STATE, "no device-resident Jacobian for that problem (armour_eval_batch_resident first)");
    CU(cudaSetDevice(h->device));
    const size_t mj = (size_t)m_of(h) * NF;
    return copy_out(h, values, h->d_bjac + (size_t)y * mj, sizeof(double) * mj);
}
int armour_last_eval_batch_ms(armour_handle* h, float* ms) {
    if (!h || !ms) return fail(ARMOUR_E_INVALID, "null argument");
    *ms = h->batch_eval_ms;
    return ARMOUR_OK;
}
int armour_upload_x(armour_handle* h, const double* x) {
    if (!h || !x) return fail(ARMOUR_E_INVALID, "null argument");
    CU(cudaSetDevice(h->device));
    memcpy(h->h_x, x, sizeof(double) * NF);   // k travels as a kernel argument
    h->have_eval = false;
    return ARMOUR_OK;
}
int armour_jac_structure(armour_handle* h, int* iRow, int* jCol) {
    if (!h || !iRow || !jCol) return fail(ARMOUR_E_INVALID, "null argument");
    const int m = m_of(h);
    for (int i = 0; i < m; i++) for (int j = 0; j < NF; j++) { iRow[i * NF + j] = i; jCol[i * NF + j] = j; }
    return ARMOUR_OK;
}
int armour_check_feasible(armour_handle* h, const double* g, int* feasible) {
    if (!h || !g || !feasible) return fail(ARMOUR_E_INVALID, "null argument");
    if (!h->built) return fail(ARMOUR_E_STATE, "feasibility check before build");
    const RobotModel& rm = h->model;
    const int T = h->T;
    const double* tr = h->h_torque_radius + (size_t)h->sel * T * NF;
    *feasible = 0;
    int offset = 0;
    if (h->mode == 0) {
        for (int i = 0; i < T; i++)
            for (int j = 0; j < NF; j++) {
                const double r = tr[i * NF + j];
                if (g[i * NF + j] < -rm.torque[j] + r - TORQUE_THRESHOLD || g[i * NF + j] > rm.torque[j] - r + TORQUE_THRESHOLD) return ARMOUR_OK;
            }
        offset += NF * T;
    }
    const int links_checked = h->mode == 0 ? NJ : NF - 1;   // the ARMTD planner re-checks links 0..NUM_FACTORS-2 only (KPA/NLPclass.cu:388)
    for (int i = 0; i < links_checked; i++) for (int j = 0; j < T; j++) for (int o = 0; o < h->n_obs; o++)
        if (g[(i * T + j) * h->n_obs + o + offset] > COLLISION_THRESHOLD) return ARMOUR_OK;
    offset += NJ * T * h->n_obs;
    for (int rep = 0; rep < 2; rep++) { for (int i = 0; i < NF; i++) if (g[offset + i] < rm.state_lb[i] + rm.qe || g[offset + i] > rm.state_ub[i] - rm.qe) return ARMOUR_OK; offset += NF; }
    for (int rep = 0; rep < 2; rep++) { for (int i = 0; i < NF; i++) if (g[offset + i] < -rm.speed[i] + rm.qde || g[offset + i] > rm.speed[i] - rm.qde) return ARMOUR_OK; offset += NF; }
    *feasible = 1;
    return ARMOUR_OK;
}

int armour_get_torque_radius(armour_handle* h, double* out) {
    if (!h || !out) return fail(ARMOUR_E_INVALID, "null argument");
    if (!h->built) return fail(ARMOUR_E_STATE, "not built");
    memcpy(out, h->h_torque_radius + (size_t)h->sel * h->T * NF, sizeof(double) * h->T * NF);
    return ARMOUR_OK;
}
int armour_get_link_generators(armour_handle* h, double* out) {
    if (!h || !out) return fail(ARMOUR_E_INVALID, "null argument");
    if (!h->built) return fail(ARMOUR_E_STATE, "not built");
    CU(cudaMemcpy(out, h->tb.gens + (size_t)h->sel * h->T * NJ * 18, sizeof(double) * h->T * NJ * 18, cudaMemcpyDeviceToHost));
    return ARMOUR_OK;
}
int armour_get_link_sliced_center(armour_handle* h, double* out) {
    if (!h || !out) return fail(ARMOUR_E_INVALID, "null argument");
    if (!h->have_eval) return fail(ARMOUR_E_STATE, "no evaluation yet");
    CU(cudaMemcpy(out, h->d_link_center, sizeof(double) * h->T * NJ * 3, cudaMemcpyDeviceToHost));
    return ARMOUR_OK;
}
int armour_get_hyperplanes(armour_handle* h, double* A, double* d, double* delta) {
    if (!h || !A || !d || !delta) return fail(ARMOUR_E_INVALID, "null argument");
    if (!h->built) return fail(ARMOUR_E_STATE, "not built");
    const size_t n = (size_t)h->T * NJ * h->n_obs * COMB, off = (size_t)h->sel * n;
    if (n == 0) return ARMOUR_OK;
    CU(cudaMemcpy(A, h->tb.A + off * 3, 8 * n * 3, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(d, h->tb.d + off, 8 * n, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(delta, h->tb.delta + off, 8 * n, cudaMemcpyDeviceToHost));
    return ARMOUR_OK;
}
int armour_get_taylor_remainders(armour_handle* h, double* cos_rem, double* sin_rem) {
    if (!h || !cos_rem || !sin_rem) return fail(ARMOUR_E_INVALID, "null argument");
    if (!h->built) return fail(ARMOUR_E_STATE, "not built");
    const size_t n = (size_t)NJ * h->T * 2, off = (size_t)h->sel * n;
    CU(cudaMemcpy(cos_rem, h->tb.cos_rem + off, 8 * n, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(sin_rem, h->tb.sin_rem + off, 8 * n, cudaMemcpyDeviceToHost));
    return ARMOUR_OK;
}

int armour_get_pz(armour_handle* h, int which, int idx, int t, int* dims, uint64_t* keys, double* coeffs, double* center, double* independent) {
    if (!h) return fail(ARMOUR_E_INVALID, "null argument");
    if (t < 0 || t >= h->T || idx < 0) return fail(ARMOUR_E_INVALID, "index out of range");
    int rc = ensure_mirror(h);
    if (rc != ARMOUR_OK) return rc;
    const int T = h->T;
    if (h->mode == 1 && (which == 8 || which == 9)) {   // no torque PZs in the ARMTD comparison planner
        if (dims) { dims[0] = 1; dims[1] = 1; }
        if (center) center[0] = 0.0;
        if (independent) independent[0] = 0.0;
        return 0;
    }
    if (which >= 0 && which <= 6) {
        if (which == 2 && idx == NJ) {   // R(NUM_JOINTS) = identity (KPR/Trajectory.cu:253)
            if (dims) { dims[0] = 3; dims[1] = 3; }
            for (int c = 0; c < 9; c++) { if (center) center[c] = h->model.R0[NJ][c]; if (independent) independent[c] = 0.0; }
            return 0;
        }
        if (idx >= NJ) return fail(ARMOUR_E_INVALID, "joint index out of range");
        if (!h->tb.traj) return fail(ARMOUR_E_STATE, "trajectory tables were not exported (cfg.export_trajectory_tables)");
        const SmallRec& r = h->m_traj[((size_t)t * TRAJ_TABLES + which) * NJ + idx];
        const int dim = r.dim;
        if (dims) { dims[0] = dim == 9 ? 3 : 1; dims[1] = dim == 9 ? 3 : 1; }
        for (int i = 0; i < r.n; i++) { if (keys) keys[i] = r.keys[i]; if (coeffs) for (int c = 0; c < dim; c++) coeffs[(size_t)i * dim + c] = r.coef[i][c]; }
        for (int c = 0; c < dim; c++) { if (center) center[c] = r.center[c]; if (independent) independent[c] = r.ind[c]; }
        return r.n;
    }
    if (idx >= NJ) return fail(ARMOUR_E_INVALID, "joint index out of range");
    const size_t rec = (size_t)t * NJ + idx;
    const size_t UCAP = h->tb.ucap, LCAP = h->tb.lcap;
    if (which == 7) {
        const int n = h->m_ln[rec];
        if (dims) { dims[0] = 3; dims[1] = 1; }
        for (int i = 0; i < n; i++) { if (keys) keys[i] = h->m_lkeys[rec * LCAP + i]; if (coeffs) for (int c = 0; c < 3; c++) coeffs[(size_t)i * 3 + c] = h->m_lcoef[(rec * 3 + c) * LCAP + i]; }
        for (int c = 0; c < 3; c++) { if (center) center[c] = h->m_lcenter[rec * 3 + c]; if (independent) independent[c] = h->m_lind[rec * 3 + c]; }
        return n;
    }
    if (which == 8) {
        const int n = h->m_un[rec];
        if (dims) { dims[0] = 1; dims[1] = 1; }
        for (int i = 0; i < n; i++) { if (keys) keys[i] = h->m_ukeys[rec * UCAP + i]; if (coeffs) coeffs[i] = h->m_ucoef[rec * UCAP + i]; }
        if (center) center[0] = h->m_ucenter[rec];
        if (independent) independent[0] = h->m_uind[rec];
        return n;
    }
    if (which == 9) {   // u_nom_int - u_nom: the polynomial parts cancel exactly, only the radius survives
        if (dims) { dims[0] = 1; dims[1] = 1; }
        if (center) center[0] = 0.0;
        if (independent) independent[0] = h->m_dist[rec];
        return 0;
    }
    (void)T;
    return fail(ARMOUR_E_INVALID, "unknown table");
}

int armour_pz_binary(armour_handle* h, int op,
                     int a_rows, int a_cols, int a_n, const uint64_t* a_keys, const double* a_coeffs, const double* a_center, const double* a_independent,
                     int b_rows, int b_cols, int b_n, const uint64_t* b_keys, const double* b_coeffs, const double* b_center, const double* b_independent,
                     int cap, int* dims, uint64_t* keys, double* coeffs, double* center, double* independent) {
    if (!h || !dims || !keys || !coeffs || !center || !independent || !a_center || !b_center || !a_independent || !b_independent) return fail(ARMOUR_E_INVALID, "null argument");
    if (op < 0 || op > 11 || op == 5 || op == 6 || a_n < 0 || b_n < 0) return fail(ARMOUR_E_INVALID, "bad op");
    {   // shapes are validated before any operand array is read (each array holds rows * cols values per monomial)
        const bool a11 = a_rows == 1 && a_cols == 1, a31 = a_rows == 3 && a_cols == 1, a33 = a_rows == 3 && a_cols == 3;
        const bool b11 = b_rows == 1 && b_cols == 1, b31 = b_rows == 3 && b_cols == 1, b33 = b_rows == 3 && b_cols == 3;
        const bool ok = op == 0 ? ((a33 && (b31 || b33)) || (a11 && b11)) : op == 3 ? (a31 && b31) : op == 4 ? (a11 || a31 || a33)
                      : (op >= 7 && op <= 9) ? (a31 && b11) : (op == 10 || op == 11) ? (a31 && b31 && b_n == 0) : ((a31 && b31) || (a11 && b11));
        if (!ok) return fail(ARMOUR_E_INVALID, "unsupported operand shapes for this operation");
        if ((a_n > 0 && (!a_keys || !a_coeffs)) || (b_n > 0 && (!b_keys || !b_coeffs)) || cap < 0) return fail(ARMOUR_E_INVALID, "null operand arrays");
    }
    CU(cudaSetDevice(h->device));
    const int da = a_rows * a_cols, db = b_rows * b_cols;
    const int rcap = std::max(cap, 1), ncap = h->ncap;
    // layout of one staging buffer: A keys/coef/center/ind | B ... | R ... | out | tmp | err
    auto pz_bytes = [](int n, int d) { return (size_t)std::max(n, 1) * 8 * (1 + d) + 18 * 8; };
    const size_t need = pz_bytes(a_n, da) + pz_bytes(b_n, db) + pz_bytes(rcap, 9) + 64 + reach_gmem_bytes(ncap) + 64;
    if (need > h->bin_bytes) { if (h->bin_buf) cudaFree(h->bin_buf); h->bin_buf = nullptr; CU(cudaMalloc((void**)&h->bin_buf, need)); h->bin_bytes = need; }
    std::vector<char> host(need, 0);
    char* base = h->bin_buf;
    size_t off = 0;
    auto place = [&](int n, int d, const uint64_t* k, const double* c, const double* cen, const double* ind, FlatPZ& f) {
        const int capn = std::max(n, 1);
        f.n = n; f.dim = d; f.cap = capn;
        f.keys = (u64*)(base + off);
        if (k) memcpy(host.data() + off, k, (size_t)n * 8);
        off += (size_t)capn * 8;
        f.coef = (double*)(base + off);
        if (c) for (int i = 0; i < n; i++) for (int e = 0; e < d; e++) ((double*)(host.data() + off))[(size_t)e * capn + i] = c[(size_t)i * d + e];   // AoS -> SoA planes
        off += (size_t)capn * 8 * d;
        f.center = (double*)(base + off);
        if (cen) memcpy(host.data() + off, cen, 8 * d);
        off += 9 * 8;
        f.ind = (double*)(base + off);
        if (ind) memcpy(host.data() + off, ind, 8 * d);
        off += 9 * 8;
    };
    FlatPZ fa, fb, fr;
    place(a_n, da, a_keys, a_coeffs, a_center, a_independent, fa);
    place(b_n, db, b_keys, b_coeffs, b_center, b_independent, fb);
    const size_t r_off = off;
    place(rcap, 9, nullptr, nullptr, nullptr, nullptr, fr);
    fr.n = 0; fr.cap = rcap;
    FlatOut* d_out = (FlatOut*)(base + off); const size_t out_off = off; off += 64;
    char* d_gmem = base + off; off += reach_gmem_bytes(ncap);
    int* d_err = (int*)(base + off); const size_t err_off = off; off += 64;
    CU(cudaMemcpyAsync(base, host.data(), need, cudaMemcpyHostToDevice, h->stream));
    CU(launch_pz_binary(op, fa, fb, fr, d_out, d_gmem, ncap, std::min(h->scap, 2048), std::min(h->tcap, 512), h->cfg.simplify_threshold, d_err, h->stream));
    h->launches += 1;
    CU(cudaMemcpyAsync(host.data(), base, need, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    const int err = *(int*)(host.data() + err_off);
    const FlatOut o = *(FlatOut*)(host.data() + out_off);
    if (err & 16) return fail(ARMOUR_E_NUMERIC, "a monomial degree outgrew its key field (KPR/PZsparse.h:23-40)");
    if (err & 2) return fail(ARMOUR_E_CAPACITY, "result does not fit in cap");
    if (err) return fail(ARMOUR_E_CAPACITY, "candidate list exceeds max_entries");
    if (o.n < 0) return fail(ARMOUR_E_INVALID, "unsupported operand shapes");
    const int d = o.dim;
    dims[0] = d == 9 ? 3 : d; dims[1] = d == 9 ? 3 : 1;
    const char* r = host.data() + r_off;
    memcpy(keys, r, (size_t)o.n * 8);
    const double* rc = (const double*)(r + (size_t)rcap * 8);
    for (int i = 0; i < o.n; i++) for (int e = 0; e < d; e++) coeffs[(size_t)i * d + e] = rc[(size_t)e * rcap + i];
    const double* rcen = (const double*)(r + (size_t)rcap * 8 * 10);
    memcpy(center, rcen, 8 * d);
    memcpy(independent, rcen + 9, 8 * d);
    return o.n;
}

int armour_standin_solve(armour_handle* h, const double* q_des, double t_plan, double* k_opt, int* feasible, int* iterations, int* evaluations) {
    if (!h || !q_des || !k_opt) return fail(ARMOUR_E_INVALID, "null argument");
    if (!h->built) return fail(ARMOUR_E_STATE, "solve before build");
    armtd_NLP nlp;
    nlp.set_time_steps(h->T);
    if (!nlp.set_parameters(q_des, t_plan, h)) return fail(ARMOUR_E_STATE, "set_parameters failed");
    // the solver's g / trial-g / Jacobian arrays are the handle's page-locked workspace: every callback is a direct device write
    // (eval_into resolves workspace addresses without registering anything), whatever cfg.pin_user_buffers says
    const size_t need = (size_t)m_of(h) * (NF + 2);
    if (h->ws_doubles < need) {
        if (h->h_ws) cudaFreeHost(h->h_ws);
        h->h_ws = nullptr; h->a_ws = nullptr; h->ws_doubles = 0;
        CU(cudaHostAlloc((void**)&h->h_ws, sizeof(double) * need, cudaHostAllocMapped));
        CU(cudaHostGetDevicePointer((void**)&h->a_ws, h->h_ws, 0));
        h->ws_doubles = need;
    }
    // arrays of the adapter object itself (g_copy in finalize_solution) are registered under cfg.pin_user_buffers like any caller
    // array; the adapter dies with this call, so those registrations go with it (a registration left on freed memory makes a
    // later registration of whatever the allocator places there fail: "already mapped")
    std::vector<const void*> before;
    for (auto& e : h->pinned) before.push_back(e.host);
    StandinResult r = standin_solve(nlp, k_opt, 60, h->h_ws);
    for (size_t i = h->pinned.size(); i-- > 0;) {
        if (std::find(before.begin(), before.end(), (const void*)h->pinned[i].host) != before.end()) continue;
        cudaHostUnregister((void*)h->pinned[i].host);
        h->pinned.erase(h->pinned.begin() + i);
    }
    cudaGetLastError();
    if (feasible) *feasible = nlp.feasible ? 1 : 0;
    if (iterations) *iterations = r.iterations;
    if (evaluations) *evaluations = r.evaluations;
    return ARMOUR_OK;
}

int armour_last_build_ms(armour_handle* h, float* total_ms, float* reach_kernel_ms, float* hyperplane_kernel_ms) {
    if (!h) return fail(ARMOUR_E_INVALID, "null argument");
    if (total_ms) *total_ms = h->build_ms;
    if (reach_kernel_ms) *reach_kernel_ms = h->reach_ms;
    if (hyperplane_kernel_ms) *hyperplane_kernel_ms = h->hyper_ms;
    return ARMOUR_OK;
}
int armour_last_eval_ms(armour_handle* h, float* kernel_ms) {
    if (!h || !kernel_ms) return fail(ARMOUR_E_INVALID, "null argument");
    if (h->eval_timed) {
        CU(cudaEventSynchronize(h->ev[4]));
        cudaEventElapsedTime(&h->eval_ms, h->ev[3], h->ev[4]);
        h->eval_timed = false;
    }
    *kernel_ms = h->eval_ms;
    return ARMOUR_OK;
}
int armour_debug_canaries_verified(armour_handle* h, int* count) {
    if (!h || !count) return fail(ARMOUR_E_INVALID, "null argument");
    *count = h->h_err ? h->h_err[2] : 0;
    return ARMOUR_OK;
}
int armour_last_eval_host_us(armour_handle* h, double* microseconds) {
    if (!h || !microseconds) return fail(ARMOUR_E_INVALID, "null argument");
    *microseconds = h->eval_host_us;
    return ARMOUR_OK;
}
int armour_set_kernel_timing(armour_handle* h, int enabled) {
    if (!h) return fail(ARMOUR_E_INVALID, "null argument");
    h->time_kernels = enabled != 0;
    if (!h->time_kernels) { h->eval_ms = -1.0f; h->eval_timed = false; }
    return ARMOUR_OK;
}
int armour_kernel_launches(armour_handle* h, uint64_t* launches) {
    if (!h || !launches) return fail(ARMOUR_E_INVALID, "null argument");
    *launches = h->launches;
    return ARMOUR_OK;
}
int armour_debug_phase_cycles(uint64_t* cycles8, uint64_t* calls8, int reset) {
    if (!cycles8 || !calls8) return fail(ARMOUR_E_INVALID, "null argument");
    read_phase_cycles((unsigned long long*)cycles8, (unsigned long long*)calls8, reset != 0);
    return ARMOUR_OK;
}
int armour_measure_fp64_peak(int device, double* tflops) {
    if (!tflops) return fail(ARMOUR_E_INVALID, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(ARMOUR_E_CUDA, "no CUDA device");
    if (device >= 0) CU(cudaSetDevice(device));
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, dev));
    *tflops = measure_fp64_tflops(prop.multiProcessorCount);
    return *tflops > 0 ? ARMOUR_OK : fail(ARMOUR_E_CUDA, "fp64 micro-benchmark failed");
}

}  // extern "C"
