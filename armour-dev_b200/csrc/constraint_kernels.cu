// Constraint path kernels.
//   hyperplane_kernel       Obstacles::initializeHyperPlane = bufferObstaclesKernel + polytope_PH
//                           (KPR/CollisionChecking.cu:74-88, 136-228), one launch for all links
//   constraint_eval_kernel  armtd_NLP::eval_g + eval_jac_g fused (KPR/NLPclass.cu:272-396):
//                           PZsparse::slice value + gradient (KPR/PZsparse.cu:404-555) of the torque and
//                           link PZs at k, checkCollisionKernel (KPR/CollisionChecking.cu:230-299) for all
//                           seven links, and the joint position / velocity limit rows
//                           (KPR/Trajectory.cu:256-540), one launch per Ipopt iteration.
#include <algorithm>
#include "armour_types.cuh"
#include "hyperplane.cuh"

namespace armour {

typedef unsigned long long u64;

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }

// One block per (problem, t, link) record, HYPER_OBS obstacles x 36 generator pairs per pass, looping over the obstacles: the link's
// 3x6 generator block and all obstacles of the problem are staged in shared memory once (one barrier), then every thread computes
// one plane per pass with no further synchronisation — 896 blocks for one plan, i.e. a single wave at 10 resident blocks per SM.
// (Round 1: a flat 64-bit index decomposed with five 64-bit divisions per thread; first round-2 version: one block per record
// AND group of 8 obstacles, 3.6 waves of one-plane threads, 22 us for 20 obstacles.)
constexpr int EVAL_MAX_OBS = 64;      // shared-memory slab of the half-space table: 64 x 1440 B = 90 KB
constexpr int HYPER_OBS = 4;
constexpr int HYPER_NT = HYPER_OBS * COMB;   // 144
__global__ void __launch_bounds__(HYPER_NT, 7) hyperplane_kernel(Tables tb) {
    __shared__ double Lg[18];                       // the link's six generators (bufferObstaclesKernel appends them to the obstacle's three)
    __shared__ double Ob[EVAL_MAX_OBS * 12];        // centre + three generators of every obstacle of this problem
    const int n_obs = tb.n_obs;
    const size_t rec = blockIdx.x;                      // (prob * T + t) * NJ + link
    const int prob = (int)(rec / ((size_t)tb.T * NJ));
    for (int e = threadIdx.x; e < 18; e += HYPER_NT) Lg[e] = tb.gens[rec * 18 + e];
    for (int e = threadIdx.x; e < n_obs * 12; e += HYPER_NT) Ob[e] = tb.obstacles[(size_t)prob * n_obs * 12 + e];
    __syncthreads();
    // (transposing the normals through shared memory for fully coalesced stores was measured slower: one barrier per pass)
    const int p = threadIdx.x % COMB, ol = threadIdx.x / COMB;
    #pragma unroll 2
    for (int o = ol; o < n_obs; o += HYPER_OBS)
        write_half_space(tb, (rec * n_obs + o) * COMB + p, Ob + o * 12 + 3, Lg, Ob + o * 12, p);
}

// ---- Bezier curve pieces used by the limit rows (KPR/Trajectory.cu:542-599) -----------------------
__device__ double q_des_func(double q0, double a, double b, double k, double t) {
    const double tm = t - 1, tm2 = tm * tm, tm3 = tm2 * tm, tm4 = tm2 * tm2, tm5 = tm4 * tm;
    const double t2 = t * t, t3 = t2 * t, t4 = t2 * t2, t5 = t4 * t;
    const double B0 = -tm5, B1 = 5 * t * tm4, B2 = -10 * t2 * tm3, B3 = 10 * t3 * tm2, B4 = -5 * t4 * tm, B5 = t5;
    const double b0 = q0, b1 = q0 + a / 5, b2 = q0 + (2 * a) / 5 + b / 20, b3 = q0 + k;
    return B0 * b0 + B1 * b1 + B2 * b2 + B3 * b3 + B4 * b3 + B5 * b3;
}
__device__ double qd_des_func(double q0, double a, double b, double k, double t) {
    const double tm = t - 1.0, tm2 = tm * tm, tm3 = tm2 * tm, tm4 = tm2 * tm2;
    const double t2 = t * t, t3 = t2 * t, t4 = t2 * t2;
    const double dB0 = tm4 * -5.0;
    const double dB1 = t * tm3 * 2.0E+1 + tm4 * 5.0;
    const double dB2 = t * tm3 * -2.0E+1 - t2 * tm2 * 3.0E+1;
    const double dB3 = t3 * (t * 2.0 - 2.0) * 1.0E+1 + t2 * tm2 * 3.0E+1;
    const double dB4 = t3 * tm * -2.0E+1 - t4 * 5.0;
    const double dB5 = t4 * 5.0;
    const double b0 = q0, b1 = q0 + a / 5, b2 = q0 + (2 * a) / 5 + b / 20, b3 = q0 + k;
    return dB0 * b0 + dB1 * b1 + dB2 * b2 + dB3 * b3 + dB4 * b3 + dB5 * b3;
}
__device__ double qdd_des_func(double q0, double a, double b, double k, double t) {
    const double t3 = t * t, t4 = t3 * t, t5 = t - 1.0, t6 = t * 2.0 - 2.0, t7 = t4 * 2.0E+1, t8 = t5 * t5, t9 = t8 * t5;
    const double t11 = t * t8 * 6.0E+1, t12 = -(t9 * 2.0E+1);
    const double ddB0 = t12, ddB1 = t9 * 4.0E+1 + t11, ddB2 = t12 - t * t8 * 1.2E+2 - t3 * t6 * 3.0E+1;
    const double ddB3 = t7 + t11 + t3 * t6 * 6.0E+1, ddB4 = t4 * -4.0E+1 - t3 * t5 * 6.0E+1, ddB5 = t7;
    const double b0 = q0, b1 = q0 + a / 5, b2 = q0 + (2 * a) / 5 + b / 20, b3 = q0 + k;
    return ddB0 * b0 + ddB1 * b1 + ddB2 * b2 + ddB3 * b3 + ddB4 * b3 + ddB5 * b3;
}

// min / max of q_des (velocity = false) or qd_des (true) over t in [0,1] and the k-derivative of the
// selected branch (returnJoint{Position,Velocity}Extremum[Gradient], KPR/Trajectory.cu:256-540).
// The reference's generated derivative expressions (:601-810) are the total derivative of the value at
// the interior stationary point t*(k); here that is evaluated as  d/dk = dF/dk|_t + dF/dt * dt*/dk.
__device__ void joint_extremum(double q0, double a, double b, double k, bool velocity, double* mn, double* mx, double* gmn, double* gmx) {
    double e2, e3, de2, de3;
    const double den = 6 * a - 12 * k + b;
    if (!velocity) {
        const double D = 64 * (a * a) + 14 * a * b - 120 * k * a + (b * b);
        const double rt = sqrt(D);
        const double n2 = 2 * a + b + rt, n3 = 2 * a + b - rt;
        e2 = n2 / (5 * den); e3 = n3 / (5 * den);
        const double drt = -60 * a / rt;   // d sqrt(D) / dk
        de2 = (drt * den + 12 * n2) / (5 * den * den);
        de3 = (-drt * den + 12 * n3) / (5 * den * den);
    }
    else {
        const double E = 150 * (k * k) - 180 * k * a - 20 * k * b + 54 * (a * a) + 14 * a * b + (b * b);
        const double rt = sqrt(6 * E);
        const double n2 = 18 * a - 30 * k + 4 * b + rt, n3 = 18 * a - 30 * k + 4 * b - rt;
        e2 = n2 / (10 * den); e3 = n3 / (10 * den);
        const double drt = 3 * (300 * k - 180 * a - 20 * b) / rt;   // d sqrt(6E) / dk
        de2 = ((-30 + drt) * den + 12 * n2) / (10 * den * den);
        de3 = ((-30 - drt) * den + 12 * n3) / (10 * den * den);
    }
    const double v1 = velocity ? qd_des_func(q0, a, b, k, 0.0) : q_des_func(q0, a, b, k, 0.0);
    const double v2 = velocity ? qd_des_func(q0, a, b, k, e2) : q_des_func(q0, a, b, k, e2);
    const double v3 = velocity ? qd_des_func(q0, a, b, k, e3) : q_des_func(q0, a, b, k, e3);
    const double v4 = velocity ? qd_des_func(q0, a, b, k, 1.0) : q_des_func(q0, a, b, k, 1.0);
    double vmn, vmx; int imn, imx;
    if (v1 < v4) { vmn = v1; imn = 1; vmx = v4; imx = 4; } else { vmn = v4; imn = 4; vmx = v1; imx = 1; }
    if (0 <= e2 && e2 <= 1) { if (v2 < vmn) { vmn = v2; imn = 2; } if (vmx < v2) { vmx = v2; imx = 2; } }
    if (0 <= e3 && e3 <= 1) { if (v3 < vmn) { vmn = v3; imn = 3; } if (vmx < v3) { vmx = v3; imx = 3; } }
    auto dk = [&](int id) -> double {
        if (id == 1) return 0.0;
        if (id == 4) return 1.0;
        const double t = (id == 2) ? e2 : e3, dt = (id == 2) ? de2 : de3;
        const double tm = t - 1.0, t2 = t * t;
        if (!velocity) {
            const double dFdk = 10 * t2 * t * tm * tm - 5 * t2 * t2 * tm + t2 * t2 * t;   // B3 + B4 + B5
            return dFdk + qd_des_func(q0, a, b, k, t) * dt;
        }
        const double dFdk = (t2 * t * (t * 2.0 - 2.0) * 10.0 + t2 * tm * tm * 30.0) + (t2 * t * tm * -20.0 - t2 * t2 * 5.0) + t2 * t2 * 5.0;
        return dFdk + qdd_des_func(q0, a, b, k, t) * dt;
    };
    *mn = vmn; *mx = vmx; *gmn = dk(imn); *gmx = dk(imx);
}

// ARMTD comparison planner: min / max joint position and velocity over the move-then-brake trajectory and the
// derivative of the selected branch with respect to k_actual (KPA/Trajectory.cu:83-411; no k_range factor there)
__device__ void armtd_state_extremum(double q0, double qd0, double k_actual, double* ext, double* grad) {
    const double t_move = 0.5, t_to_stop = 0.5;
    const double q_peak = q0 + qd0 * t_move + k_actual * t_move * t_move * 0.5;
    const double q_dot_peak = qd0 + k_actual * t_move;
    const double q_ddot_to_stop = -q_dot_peak / t_to_stop;
    const double q_stop = q_peak + q_dot_peak * t_to_stop + 0.5 * q_ddot_to_stop * t_to_stop * t_to_stop;
    const double t_mm = -qd0 / k_actual;
    double qe0, qe1, ge0, ge1;
    if (q_peak >= q0) { qe0 = q0; qe1 = q_peak; ge0 = 0; ge1 = 0.5 * t_move * t_move; }
    else { qe0 = q_peak; qe1 = q0; ge0 = 0.5 * t_move * t_move; ge1 = 0; }
    double q_min_pk = qe0, q_max_pk = qe1, g_min_pk = ge0, g_max_pk = ge1;
    if (t_mm > 0 && t_mm < t_move) {
        const double qm = q0 + qd0 * t_mm + 0.5 * k_actual * t_mm * t_mm, gm = (0.5 * qd0 * qd0) / (k_actual * k_actual);
        if (k_actual >= 0) { q_min_pk = qm; g_min_pk = gm; q_max_pk = qe1; g_max_pk = ge1; }
        else { q_max_pk = qm; g_max_pk = gm; q_min_pk = qe0; g_min_pk = ge0; }
    }
    double v_min_pk, v_max_pk, gv_min_pk, gv_max_pk;
    if (q_dot_peak >= qd0) { v_min_pk = qd0; v_max_pk = q_dot_peak; gv_min_pk = 0; gv_max_pk = t_move; }
    else { v_min_pk = q_dot_peak; v_max_pk = qd0; gv_min_pk = t_move; gv_max_pk = 0; }
    double q_min_st, q_max_st, g_min_st, g_max_st;
    if (q_stop >= q_peak) { q_min_st = q_peak; q_max_st = q_stop; g_min_st = 0.5 * t_move * t_move; g_max_st = 0.5 * t_move * t_move + 0.5 * t_move * t_to_stop; }
    else { q_min_st = q_stop; q_max_st = q_peak; g_min_st = 0.5 * t_move * t_move + 0.5 * t_move * t_to_stop; g_max_st = 0.5 * t_move * t_move; }
    double v_min_st, v_max_st, gv_min_st, gv_max_st;
    if (q_dot_peak >= 0) { v_min_st = 0; v_max_st = q_dot_peak; gv_min_st = 0; gv_max_st = t_move; }
    else { v_min_st = q_dot_peak; v_max_st = 0; gv_min_st = t_move; gv_max_st = 0; }
    const bool a = q_min_pk <= q_min_st, b = q_max_pk >= q_max_st, cc = v_min_pk <= v_min_st, d = v_max_pk >= v_max_st;
    ext[0] = a ? q_min_pk : q_min_st; grad[0] = a ? g_min_pk : g_min_st;
    ext[1] = b ? q_max_pk : q_max_st; grad[1] = b ? g_max_pk : g_max_st;
    ext[2] = cc ? v_min_pk : v_min_st; grad[2] = cc ? gv_min_pk : gv_min_st;
    ext[3] = d ? v_max_pk : v_max_st; grad[3] = d ? gv_max_pk : gv_max_st;
}

// x^d for d in 0..3 the way std::pow returns it for these exponents (exact products of x)
__device__ __forceinline__ double powi(double x, int d) { return d == 0 ? 1.0 : d == 1 ? x : d == 2 ? x * x : x * x * x; }

// term and k-gradient of one k-only monomial at x (PZsparse::slice, KPR/PZsparse.cu:404-555):
//   value:   coeff * prod_j pow(x_j, d_j)            (multiplied in j order)
//   grad[k]: coeff * prod_j (j == k ? d_j pow(x_j, d_j - 1) : pow(x_j, d_j)),  0 when d_k == 0
__device__ __forceinline__ double slice_term(double coef, u64 key, const double* x, int which) {
    double r = coef;
#pragma unroll
    for (int j = 0; j < NF; j++) {
        const int d = (int)((key >> (2 * j)) & 3);
        if (j == which) {
            if (d == 0) r = 0.0;
            else r = r * ((double)d * powi(x[j], d - 1));
        }
        else r = r * powi(x[j], d);
    }
    return r;
}

// ---- fused eval_g + eval_jac_g --------------------------------------------------------------------------------------
// One block per (t, link) plus one for the 28 limit rows; every block of the benchmark configurations is resident at once
// (7 blocks per SM x 148 SMs >= 7 T + 1 at T = 128, up to 20 obstacles).  The block's slab of the half-space table
// (n_obs x 36 planes x 40 B, contiguous in A, d and delta) is fetched by three bulk asynchronous copies (TMA, completion on an
// mbarrier) issued before anything else, so the table streams in while the block slices its two polynomial zonotopes at k.
constexpr int EVAL_NT = 128;
constexpr int EVAL_LCHUNK = 5;        // link monomials staged per pass (x 24 outputs)
// torque monomials staged per pass (x 8 outputs): as many as fit beside the table slab while seven blocks stay resident per SM
// (7 x (static + dynamic + 1 KB) <= 228 KB), between 16 and the table capacity
__host__ __device__ inline int eval_uchunk(int n_obs, int ucap) {
    const long budget = 32073 - 256 - (long)n_obs * COMB * 40 - 24 * EVAL_LCHUNK * 8;
    long u = budget / 64;
    u = u < 16 ? 16 : u;
    u = u > ucap ? ucap : u;
    u = u < 1 ? 1 : u;
    return (int)(u > 16 ? (u & ~15L) : u);
}
struct XArg { double x[NF]; };        // k passed by value: no host-to-device copy on the per-iteration path

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(mbar), "r"(parity) : "memory");
}

// dynamic shared memory of constraint_eval_kernel: the table slab, then a staging area used first for slice terms and then
// for the block's output rows
__host__ __device__ inline size_t eval_smem_bytes(int n_obs, int uchunk) {
    const size_t slab = (size_t)n_obs * COMB * 40;
    const size_t terms = (size_t)(8 * uchunk + 24 * EVAL_LCHUNK) * 8;
    const size_t rows = (size_t)n_obs * 8 * 8;
    return slab + (terms > rows ? terms : rows);
}

// grid: T * NJ + 1 blocks for the selected problem: block (t, j) handles torque row (t, j) and link j at interval t against
// every obstacle; the extra block handles the 28 limit rows.  g / jac may point to device memory or to pinned host memory
// (UVA): the rows of one block are staged in shared memory and written out as contiguous runs, so over PCIe they leave as
// full-width posted writes while the rest of the grid is still computing (no device-to-host copy after the kernel).
// done_counter / done_flag: when done_flag != nullptr the last block to finish stores `seq` there (mapped pinned host memory)
// after a system-wide fence — the host polls that word instead of synchronising the stream.
// Batched form (armour_eval_batch): gridDim.y problems starting at `prob`, each with its own decision vector x_batch[y][7]
// and its own block of rows in g / jac / link_center_out (problem-major) — one launch steps every solver of a sweep.
__global__ void __launch_bounds__(EVAL_NT, 7) constraint_eval_kernel(Tables tb, int prob, XArg xarg, const double* __restrict__ x_batch, double* __restrict__ g, double* __restrict__ jac,
                                                                     double* __restrict__ link_center_out, int what, int uchunk, int lchunk,
                                                                     unsigned* done_counter, volatile unsigned long long* done_flag, unsigned long long seq) {
    extern __shared__ __align__(128) unsigned char eval_smem[];
    __shared__ __align__(8) unsigned long long mbar_storage;
    __shared__ double lc[3], ldk[NF][3];
    const int T = tb.T, n_obs = tb.n_obs;
    const int tid = threadIdx.x;
    const int blk = blockIdx.x;
    double x[NF];
#pragma unroll
    for (int i = 0; i < NF; i++) x[i] = xarg.x[i];
    const size_t off_obs = tb.mode == 0 ? (size_t)NF * T : 0, off_lim = off_obs + (size_t)NJ * T * n_obs;
    if (x_batch != nullptr) {
        const size_t y = blockIdx.y, m_rows = off_lim + 4 * NF;
        prob += (int)y;
#pragma unroll
        for (int i = 0; i < NF; i++) x[i] = x_batch[y * NF + i];
        if (g) g += y * m_rows;
        if (jac) jac += y * m_rows * NF;
        if (link_center_out) link_center_out += y * (size_t)T * NJ * 3;
    }
    const int n_planes = n_obs * COMB;
    double* sA = reinterpret_cast<double*>(eval_smem);
    double* sd = sA + (size_t)n_planes * 3;
    double* sdl = sd + n_planes;
    double* stage = sdl + n_planes;              // slice terms, later the block's output rows
    const bool want_g = what & 1, want_j = what & 2;   // Ipopt asks for g and for the Jacobian in separate callbacks: each launch computes what is asked for
    if (blk == T * NJ) {   // limit rows (KPR/NLPclass.cu:319-320, 393-394)
        if (tb.mode == 1) {   // KPA/NLPclass.cu:275-276
            if (tid < NF) {
                const double* st = tb.state + (size_t)prob * 21;
                double ext[4], gr[4];
                armtd_state_extremum(st[tid], st[7 + tid], tb.k_range_in[(size_t)prob * NF + tid] * x[tid], ext, gr);
                for (int r = 0; r < 4; r++) {
                    const size_t row = off_lim + r * NF + tid;
                    if (want_g) g[row] = ext[r];
                    if (want_j) for (int j = 0; j < NF; j++) jac[row * NF + j] = (j == tid) ? gr[r] : 0.0;
                }
            }
        }
        else if (tid < 2 * NF) {
            const int i = tid % NF;
            const bool velocity = tid >= NF;
            const double* st = tb.state + (size_t)prob * 21;
            const double kr = tb.k_range[i];
            double mn, mx, gmn, gmx;
            joint_extremum(st[i], st[7 + i], st[14 + i], kr * x[i], velocity, &mn, &mx, &gmn, &gmx);
            const size_t r0 = off_lim + (velocity ? 2 * NF : 0) + i, r1 = r0 + NF;
            if (want_g) { g[r0] = mn; g[r1] = mx; }   // DURATION == 1
            if (want_j) for (int j = 0; j < NF; j++) { jac[r0 * NF + j] = (i == j) ? gmn * kr : 0.0; jac[r1 * NF + j] = (i == j) ? gmx * kr : 0.0; }
        }
    }
    else {
        const int t = blk / NJ, j = blk - t * NJ;
        const size_t rec = ((size_t)prob * T + t) * NJ + j;
        const unsigned mbar = smem_u32(&mbar_storage);
        // ---- table slab: three bulk copies in flight from the first instruction on -----------------------------
        if (tid == 0 && n_obs > 0) {
            mbar_init(mbar, 1);
            const size_t base = rec * n_obs * COMB;
            mbar_expect_tx(mbar, (unsigned)n_planes * 40u);
            bulk_g2s(smem_u32(sA), tb.A + base * 3, (unsigned)n_planes * 24u, mbar);
            bulk_g2s(smem_u32(sd), tb.d + base, (unsigned)n_planes * 8u, mbar);
            bulk_g2s(smem_u32(sdl), tb.delta + base, (unsigned)n_planes * 8u, mbar);
        }
        // ---- slices (PZsparse::slice, KPR/PZsparse.cu:404-555): terms in parallel, sums sequential in key order ----
        const int UCAP = tb.ucap, LCAP = tb.lcap;
        // operands of this thread's first torque / link term, requested BEFORE the monomial counts arrive: the tables are
        // allocated to capacity, so any row below it is readable, and the dependent round trip (count, then operands) becomes one
        const int um0 = tid >> 3, lm0 = tid / 24, lw0 = tid - lm0 * 24;
        double uc0 = 0.0, lc0 = 0.0;
        u64 uk0 = 0, lk0 = 0;
        if (tb.mode == 0 && um0 < UCAP) { uc0 = tb.u_coef[rec * UCAP + um0]; uk0 = tb.u_keys[rec * UCAP + um0]; }
        if (lm0 < LCAP) { lc0 = tb.l_coef[(rec * 3 + lw0 % 3) * LCAP + lm0]; lk0 = tb.l_keys[rec * LCAP + lm0]; }
        const int un = tb.mode == 0 ? tb.u_n[rec] : 0;
        const int ln = tb.l_n[rec];
        double* uterm = stage;                      // [8][uchunk]
        double* lterm = stage + 8 * uchunk;         // [24][lchunk]
        // threads 0..7: torque value + 7 derivatives; threads 32..55: link value (3) + derivatives (21)
        double acc = 0.0;
        const bool sum_u = tid < 8, sum_l = tid >= 32 && tid < 56;
        const int lw = tid - 32;
        if (sum_u && tid == 0 && tb.mode == 0) acc = tb.u_center[rec];
        if (sum_l && lw < 3) acc = tb.l_center[rec * 3 + lw];
        const int u_passes = (un + uchunk - 1) / uchunk, l_passes = (ln + lchunk - 1) / lchunk;
        const int passes = u_passes > l_passes ? u_passes : l_passes;
        for (int ps = 0; ps < passes; ps++) {
            const int ub = ps * uchunk, uc = min(uchunk, un - ub);
            const int lb = ps * lchunk, lcn = min(lchunk, ln - lb);
            for (int e = tid; e < uc * 8; e += EVAL_NT) {
                const int m = e >> 3, w = e & 7;
                if (!((want_j || w == 0) && (want_g || w != 0))) continue;
                const bool pre = ps == 0 && e == tid;
                uterm[w * uchunk + m] = slice_term(pre ? uc0 : tb.u_coef[rec * UCAP + ub + m], pre ? uk0 : tb.u_keys[rec * UCAP + ub + m], x, w - 1);
            }
            for (int e = tid; e < lcn * 24; e += EVAL_NT) {
                const int m = e / 24, w = e - m * 24;
                const int c = w % 3, which = w / 3;   // which 0: value, 1..7: d/dk_{which-1}
                if (!(want_j || w < 3)) continue;     // the link centre (value) is needed for both g and the arg-max of the Jacobian rows
                const bool pre = ps == 0 && e == tid;
                lterm[w * lchunk + m] = slice_term(pre ? lc0 : tb.l_coef[(rec * 3 + c) * LCAP + lb + m], pre ? lk0 : tb.l_keys[rec * LCAP + lb + m], x, which - 1);
            }
            __syncthreads();
            if (sum_u) for (int m = 0; m < uc; m++) acc = acc + uterm[tid * uchunk + m];
            if (sum_l) for (int m = 0; m < lcn; m++) acc = acc + lterm[lw * lchunk + m];
            __syncthreads();
        }
        if (sum_u && tb.mode == 0) {
            const size_t row = (size_t)t * NF + j;
            if (tid == 0) {
                const double r = tb.u_ind[rec];
                if (want_g) g[row] = ((acc - r) + (acc + r)) * 0.5;   // getCenter(Interval(c - r, c + r))
            }
            else if (want_j) jac[row * NF + (tid - 1)] = acc;
        }
        if (sum_l) {
            const int c = lw % 3, which = lw / 3;
            if (which == 0) {
                const double r = tb.l_ind[rec * 3 + c];
                const double v = ((acc - r) + (acc + r)) * 0.5;
                lc[c] = v;
                if (link_center_out) link_center_out[((size_t)t * NJ + j) * 3 + c] = v;
            }
            else ldk[which - 1][c] = acc;
        }
        __syncthreads();
        // ---- obstacle rows (checkCollisionKernel, KPR/CollisionChecking.cu:230-299): four lanes per obstacle, nine planes
        // each; the reference scans pos_0, neg_0, pos_1, ... and keeps the FIRST maximum, i.e. the maximum value with the
        // smallest order index 2 * plane + (neg ? 1 : 0) — a max-then-min reduction, independent of the reduction order.
        if (n_obs > 0) {
            mbar_wait(mbar, 0);
            double* sg = stage;                  // [n_obs]
            double* sjac = stage + n_obs;        // [n_obs][7]
            const double c0 = lc[0], c1 = lc[1], c2 = lc[2];
            const int q = tid & 3;
            for (int o = tid >> 2; o < ((n_obs + 31) & ~31); o += EVAL_NT / 4) {   // whole warps stay in the loop for the shuffles
                double best = -100000000.0;
                int best_e = 0x7fffffff;
                if (o < n_obs) {
                    const int p0 = o * COMB + q * 9;
#pragma unroll 3
                    for (int i = 0; i < 9; i++) {
                        const double a0 = sA[(p0 + i) * 3], a1 = sA[(p0 + i) * 3 + 1], a2 = sA[(p0 + i) * 3 + 2];
                        double pos = -100000000.0, neg = -100000000.0;
                        // A.norm() > 0 (KPR/CollisionChecking.cu:252): the rows of this table are unit vectors or exactly zero
                        // (hyperplane_kernel), for which "norm > 0" is "some component is not +-0" — three bit tests, no arithmetic
                        if (((__double_as_longlong(a0) | __double_as_longlong(a1) | __double_as_longlong(a2)) << 1) != 0) {
                            const double dot = dadd(dadd(dmul(a0, c0), dmul(a1, c1)), dmul(a2, c2));
                            const double dd = sd[p0 + i], dl = sdl[p0 + i];
                            pos = dadd(dot, -dadd(dd, dl));
                            neg = dadd(-dot, -dadd(-dd, dl));
                        }
                        if (pos > best) { best = pos; best_e = 2 * (q * 9 + i); }
                        if (neg > best) { best = neg; best_e = 2 * (q * 9 + i) + 1; }
                    }
                }
#pragma unroll
                for (int sft = 1; sft <= 2; sft <<= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, sft);
                    const int oe = __shfl_xor_sync(0xffffffffu, best_e, sft);
                    if (ob > best || (ob == best && oe < best_e)) { best = ob; best_e = oe; }
                }
                // no candidate beat the sentinel: the reference's thread 0 reports max_id 0 / pos, i.e. order index 0
                if (best_e == 0x7fffffff) best_e = 0;
                if (o < n_obs) {
                    if (q == 0) sg[o] = -best;
                    const int p = o * COMB + (best_e >> 1);
                    const bool neg = best_e & 1;
                    const double a0 = sA[p * 3], a1 = sA[p * 3 + 1], a2 = sA[p * 3 + 2];
                    if (want_j) for (int k = q; k < NF; k += 4) {
                        const double dot = dadd(dadd(dmul(a0, ldk[k][0]), dmul(a1, ldk[k][1])), dmul(a2, ldk[k][2]));
                        sjac[o * NF + k] = neg ? dot : -dot;
                    }
                }
            }
            __syncthreads();
            // rows (j*T + t)*n_obs + [0, n_obs) are contiguous in g and in jac: coalesced write-out
            const size_t row0 = off_obs + ((size_t)j * T + t) * n_obs;
            if (want_g) for (int e = tid; e < n_obs; e += EVAL_NT) g[row0 + e] = sg[e];
            if (want_j) for (int e = tid; e < n_obs * NF; e += EVAL_NT) jac[row0 * NF + e] = sjac[e];
        }
    }
    // ---- completion word for the polling host (see above) ----------------------------------------------------------
    if (done_flag != nullptr) {
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
            const unsigned prev = atomicAdd(done_counter, 1u);
            if (prev == gridDim.x * gridDim.y - 1) {
                *done_counter = 0;               // ready for the next launch (stream order)
                __threadfence_system();
                *done_flag = seq;
            }
        }
    }
}

cudaError_t launch_hyperplanes(const Tables& tb, cudaStream_t stream) {
    const size_t recs = (size_t)tb.P * tb.T * NJ;
    if (recs == 0 || tb.n_obs == 0) return cudaSuccess;
    if (tb.n_obs > EVAL_MAX_OBS) return cudaErrorInvalidValue;
    hyperplane_kernel<<<(unsigned)recs, HYPER_NT, 0, stream>>>(tb);
    return cudaGetLastError();
}
int eval_max_obstacles() { return EVAL_MAX_OBS; }
// blocks_per_sm > 0 caps the resident blocks per SM by padding the dynamic shared-memory request.  With every block resident
// at once (the default, best for device-resident results) all blocks finish together; when the results go to host memory
// over PCIe it pays to run the grid in a few waves, so that the first rows are on the wire while later blocks still compute.
static cudaError_t eval_opt_in(size_t smem, size_t smem_max) {
    if (smem > 48 * 1024) {   // opt in to a large dynamic shared-memory request, once per device
        static bool opted_in[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !opted_in[dev]) {
            cudaError_t e = cudaFuncSetAttribute(constraint_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) opted_in[dev] = true;
        }
    }
    return cudaSuccess;
}
cudaError_t launch_constraint_eval(const Tables& tb, int prob, const double* x_host, double* g, double* jac, double* link_center, int what, unsigned* done_counter,
                                   unsigned long long* done_flag, unsigned long long seq, int blocks_per_sm, cudaStream_t stream) {
    if (tb.n_obs > EVAL_MAX_OBS) return cudaErrorInvalidValue;
    XArg xa;
    for (int i = 0; i < NF; i++) xa.x[i] = x_host[i];
    const int uchunk = eval_uchunk(tb.n_obs, tb.ucap);
    const size_t smem_max = eval_smem_bytes(EVAL_MAX_OBS, 128);
    size_t smem = eval_smem_bytes(tb.n_obs, uchunk);
    if (blocks_per_sm > 0) smem = std::max(smem, std::min((size_t)(227 * 1024) / blocks_per_sm - 1280, smem_max));
    cudaError_t e = eval_opt_in(smem, smem_max);
    if (e != cudaSuccess) return e;
    constraint_eval_kernel<<<tb.T * NJ + 1, EVAL_NT, smem, stream>>>(tb, prob, xa, nullptr, g, jac, link_center, what, uchunk, EVAL_LCHUNK, done_counter, done_flag, seq);
    return cudaGetLastError();
}
// `count` problems starting at `first` in one launch: x_dev[count][7] in device-visible memory, rows of problem y at
// g + y * m, jac + y * 7 m, link_center + y * 21 T (device memory or mapped pinned host memory).
cudaError_t launch_constraint_eval_batch(const Tables& tb, int first, int count, const double* x_dev, double* g, double* jac, double* link_center, int what, cudaStream_t stream) {
    if (tb.n_obs > EVAL_MAX_OBS || count < 1 || count > 65535 || x_dev == nullptr) return cudaErrorInvalidValue;
    XArg xa;
    for (int i = 0; i < NF; i++) xa.x[i] = 0.0;
    const int uchunk = eval_uchunk(tb.n_obs, tb.ucap);
    const size_t smem_max = eval_smem_bytes(EVAL_MAX_OBS, 128), smem = eval_smem_bytes(tb.n_obs, uchunk);
    cudaError_t e = eval_opt_in(smem, smem_max);
    if (e != cudaSuccess) return e;
    constraint_eval_kernel<<<dim3(tb.T * NJ + 1, count), EVAL_NT, smem, stream>>>(tb, first, xa, x_dev, g, jac, link_center, what, uchunk, EVAL_LCHUNK, nullptr, nullptr, 0ull);
    return cudaGetLastError();
}

// ---- fp64 FMA micro-benchmark: denominator of the fp64 roofline -----------------------------------------
__global__ void fp64_fma_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
double measure_fp64_tflops(int sm_count) {
    const int nt = 512, grid = sm_count * 4, iters = 1 << 14;
    double* out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * nt * grid) != cudaSuccess) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fp64_fma_kernel<<<grid, nt>>>(out, iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        fp64_fma_kernel<<<grid, nt>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    const double flops = 2.0 * 8 * (double)iters * nt * grid;
    return flops / (best * 1e-3) / 1e12;
}

}  // namespace armour
