// Task-scheduled variant of the reach-set build (single-plan latency path).  Included at the end of reach_kernels.cu.
//
// One CTA still owns one (problem, time interval), but the ~270 PZ operations of the interval are no longer cut by hand
// for two thread groups: they are grouped into 99 small tasks (2-4 operations each, see the table below) over a
// dependency graph, and G thread groups of NT threads pick ready tasks at run time (list scheduling, highest priority
// first).  A single interval is a dependency chain of small sorts — latency-bound — and the chain through the graph
// (about 58 operations: the wdot / linear_acc recurrences forward, the moment recursion backward) is less than half
// of what either hand-cut group executed; the rest (forces, moments, cross terms, forward kinematics) fills the other
// groups.  Every PZ operation has exactly the operands and order of chain_joint / force_joint / backward_joint /
// fk_joint (and therefore of KPR/Dynamics.cu:69-181); only who executes it, and when, changes.
//
// Hazards: every task writes slots nobody else writes (the joint state is indexed by joint instead of double-buffered,
// each group has private temporaries), so dependencies are pure read-after-write and the `done` flags are the only
// synchronisation.  A task is published by its group's thread 0 after the closing group barrier of its last operation
// (__threadfence_block, then state = 2); the picking warp of another group reads the flags, fences, and the group
// barrier hands the visibility to the rest of that group.  Program order is a topological order and every group runs
// its tasks to completion, so the earliest unfinished task is always runnable: no deadlock.
#pragma once

namespace armour {

enum TaskType : int {   // priority order: critical recurrences first, forward kinematics as filler, stage M last
    TK_A2 = 0,   // WAm[i]   = R_t * WA[i]                                              (w_aux before qda is added)
    TK_A5,       // WA[i+1]  = WAm[i] + qda_i
    TK_A3,       // X[i]     = cross(WAm[i], qd_i * z)
    TK_A4,       // WD[i+1]  = R_t * WD[i] + X[i] + qdda_i
    TK_A1,       // W[i+1]   = R_t * W[i] + qd_i
    TK_L1,       // LT0[i]   = cross(WD[i], trans_i);  LT2[i] = cross(W[i], cross(WA[i], trans_i))
    TK_L2,       // LA[i+1]  = R_t * ((LA[i] + LT0[i]) + LT2[i])
    TK_B1,       // Rf[i]    = R_{i+1} * f_{i+1};  C2[i] = cross(trans_{i+1}, Rf[i]);  f_i = Rf[i] + F[i]
    TK_B3,       // n_i      = ((N[i] + R_{i+1} * n_{i+1}) + C1[i]) + C2[i]
    TK_F1,       // FT0[i]   = cross(WD[i+1], com_i);  FT2[i] = cross(W[i+1], cross(WA[i+1], com_i))
    TK_F2,       // F[i]     = m_i * ((LA[i+1] + FT0[i]) + FT2[i])
    TK_N,        // N[i]     = I_i * WD[i+1] + cross(WA[i+1], I_i * W[i+1])
    TK_B2,       // C1[i]    = cross(com_i, F[i])
    TK_B4,       // u_i      = n_i(axis) + armature * qdda + damping * qd;  export (disturbance radius, reduce())
    TK_K,        // forward kinematics of joint i + reduce_link_PZ export
    TK_TYPES,
};
constexpr int N_TASKS = TK_TYPES * NJ + 1;   // + stage M (torque radius) at the very end
constexpr int TASK_M = TK_TYPES * NJ;
__host__ __device__ constexpr int task_id(int type, int joint) { return type * NJ + joint; }

struct TaskDeps { short d[4]; };
// dependencies of task (type, i); -1 = none
__device__ __forceinline__ TaskDeps task_deps(int type, int i) {
    TaskDeps r = {{-1, -1, -1, -1}};
    const bool first = i == 0, last = i == NJ - 1;
    switch (type) {
        case TK_A2: if (!first) r.d[0] = task_id(TK_A5, i - 1); break;
        case TK_A5: r.d[0] = task_id(TK_A2, i); break;
        case TK_A3: r.d[0] = task_id(TK_A2, i); break;
        case TK_A4: r.d[0] = task_id(TK_A3, i); if (!first) r.d[1] = task_id(TK_A4, i - 1); break;
        case TK_A1: if (!first) r.d[0] = task_id(TK_A1, i - 1); break;
        case TK_L1: if (!first) { r.d[0] = task_id(TK_A4, i - 1); r.d[1] = task_id(TK_A1, i - 1); r.d[2] = task_id(TK_A5, i - 1); } break;
        case TK_L2: r.d[0] = task_id(TK_L1, i); if (!first) r.d[1] = task_id(TK_L2, i - 1); break;
        case TK_B1: r.d[0] = task_id(TK_F2, i); if (!last) r.d[1] = task_id(TK_B1, i + 1); break;
        case TK_B3: r.d[0] = task_id(TK_N, i); r.d[1] = task_id(TK_B2, i); r.d[2] = task_id(TK_B1, i); if (!last) r.d[3] = task_id(TK_B3, i + 1); break;
        case TK_F1: r.d[0] = task_id(TK_A4, i); r.d[1] = task_id(TK_A1, i); r.d[2] = task_id(TK_A5, i); break;
        case TK_F2: r.d[0] = task_id(TK_L2, i); r.d[1] = task_id(TK_F1, i); break;
        case TK_N:  r.d[0] = task_id(TK_A4, i); r.d[1] = task_id(TK_A1, i); r.d[2] = task_id(TK_A5, i); break;
        case TK_B2: r.d[0] = task_id(TK_F2, i); break;
        case TK_B4: r.d[0] = task_id(TK_B3, i); if (!last) r.d[1] = task_id(TK_B4, i + 1); break;   // chained so that B4(0) done = all exported
        case TK_K:  if (!first) r.d[0] = task_id(TK_K, i - 1); break;
    }
    return r;
}

template <int G>
struct TaskSlots {
    PZ<3> W[NJ + 1], WA[NJ + 1], WD[NJ + 1], LA[NJ + 1];   // state before joint i at index i (index 0 = base, KPR/Dynamics.cu:87-99)
    PZ<3> WAm[NJ], X[NJ], LT0[NJ], LT2[NJ], FT0[NJ], FT2[NJ];
    PZ<3> F[NJ], N[NJ], Rf[NJ], C1[NJ], C2[NJ];
    PZ<3> Fv[NJ + 1], Nv[NJ + 1];                          // f, n of joint i at index i; index NJ = 0
    PZ<3> FKT[NJ + 1], LINK[NJ], link0[NJ];
    PZ<3> T[G][4];                                         // private temporaries of each thread group
    PZ<3> Zero;
    PZ<9> FKR[NJ + 1], R[NJ + 1], Rt[NJ];
    PZ<1> qd[NJ], qda[NJ], qdda[NJ], u[NJ], u0[NJ], cosq[NJ], sinq[NJ];
};
constexpr int task_big3(int G) { return 4 * (NJ + 1) + 6 * NJ + 5 * NJ + 2 * (NJ + 1) + (NJ + 1) + NJ + 4 * G; }

size_t task_arena_bytes(int mcap, int ncap, int G) {
    size_t b = 0;
    b += (size_t)task_big3(G) * mcap * (8 + 3 * 8);
    b += (size_t)(NJ + 1) * SMALL_CAP * (8 + 3 * 8);               // link0, Zero
    b += (size_t)(NJ + 1) * mcap * (8 + 9 * 8);                    // FK_R
    b += (size_t)(2 * NJ + 1) * SMALL_CAP * (8 + 9 * 8);           // R, R_t
    b += (size_t)5 * NJ * SMALL_CAP * 16;                          // qd, qda, qdda, cos, sin
    b += (size_t)2 * NJ * mcap * 16;                               // u, u0
    b += (size_t)G * Scratch::gmem_bytes(ncap);
    return (b + 255) & ~(size_t)255;
}

template <int D> __device__ __forceinline__ void slot_zero(PZ<D>& z) {
    z.n = 0; z.divM = FastDiv::magic(0);
    for (int c = 0; c < D; c++) { z.center[c] = 0; z.ind[0][c] = 0; z.ind[1][c] = 0; z.abss[c] = 0; }
}

// one task = the operations of the reference program between two hand-over points, run by one thread group
template <int NT, int G>
__device__ void run_task(Scratch& S, TaskSlots<G>& Z, PZ<3>* T, const Tables& tb, size_t rec0, int type, int i) {
    const RobotModel& rm = c_robot;
    const int axis = rm.axes[i];
    const int row = (axis < 0 ? -axis : axis) - 1;
    switch (type) {
        case TK_A2: pz_mul<NT, 9, 3, 3>(S, Z.WAm[i], Z.Rt[i], Z.WA[i]); break;
        case TK_A5:
            if (axis != 0) pz_add_one_dim<NT>(S, Z.WA[i + 1], Z.WAm[i], Z.qda[i], row);
            else pz_add3<NT>(S, Z.WA[i + 1], Z.WAm[i], Z.Zero);
            break;
        case TK_A3:
            if (axis != 0) {
                pz_set_const<NT, 3>(T[0], nullptr);
                pz_add_one_dim<NT>(S, T[0], T[0], Z.qd[i], row);
                pz_cross_pp<NT>(S, Z.X[i], Z.WAm[i], T[0]);
            }
            break;
        case TK_A4:
            if (axis != 0) {
                pz_mul<NT, 9, 3, 3>(S, T[0], Z.Rt[i], Z.WD[i]);
                pz_add3<NT>(S, T[1], T[0], Z.X[i]);
                pz_add_one_dim<NT>(S, Z.WD[i + 1], T[1], Z.qdda[i], row);
            }
            else pz_mul<NT, 9, 3, 3>(S, Z.WD[i + 1], Z.Rt[i], Z.WD[i]);
            break;
        case TK_A1:
            if (axis != 0) {
                pz_mul<NT, 9, 3, 3>(S, T[0], Z.Rt[i], Z.W[i]);
                pz_add_one_dim<NT>(S, Z.W[i + 1], T[0], Z.qd[i], row);
            }
            else pz_mul<NT, 9, 3, 3>(S, Z.W[i + 1], Z.Rt[i], Z.W[i]);
            break;
        case TK_L1:
            pz_cross_const<NT>(S, Z.LT0[i], Z.WD[i], rm.trans[i], false);
            pz_cross_const<NT>(S, T[1], Z.WA[i], rm.trans[i], false);
            pz_cross_pp<NT>(S, Z.LT2[i], Z.W[i], T[1]);
            break;
        case TK_L2:
            pz_add3<NT>(S, T[0], Z.LA[i], Z.LT0[i]);
            pz_add3<NT>(S, T[1], T[0], Z.LT2[i]);
            pz_mul<NT, 9, 3, 3>(S, Z.LA[i + 1], Z.Rt[i], T[1]);
            break;
        case TK_F1:
            pz_cross_const<NT>(S, Z.FT0[i], Z.WD[i + 1], rm.com[i], false);
            pz_cross_const<NT>(S, T[1], Z.WA[i + 1], rm.com[i], false);
            pz_cross_pp<NT>(S, Z.FT2[i], Z.W[i + 1], T[1]);
            break;
        case TK_F2: {
            pz_add3<NT>(S, T[0], Z.LA[i + 1], Z.FT0[i]);
            pz_add3<NT>(S, T[1], T[0], Z.FT2[i]);
            const double m0 = 0.0, m1 = __dmul_ru(tb.mass_unc, fabs(rm.mass[i]));
            pz_const_left<NT>(S, Z.F[i], &rm.mass[i], &m0, &m1, true, T[1]);
            break;
        }
        case TK_N: {
            double I0[9], I1[9];
            for (int k = 0; k < 9; k++) { I0[k] = 0.0; I1[k] = __dmul_ru(tb.inertia_unc, fabs(rm.inertia[i][k])); }
            pz_const_left<NT>(S, T[0], rm.inertia[i], I0, I1, false, Z.WD[i + 1]);
            pz_const_left<NT>(S, T[1], rm.inertia[i], I0, I1, false, Z.W[i + 1]);
            pz_cross_pp<NT>(S, T[2], Z.WA[i + 1], T[1]);
            pz_add3<NT>(S, Z.N[i], T[0], T[2]);
            break;
        }
        case TK_B1:
            pz_mul<NT, 9, 3, 3>(S, Z.Rf[i], Z.R[i + 1], Z.Fv[i + 1]);
            pz_cross_const<NT>(S, Z.C2[i], Z.Rf[i], rm.trans[i + 1], true);
            pz_add3<NT>(S, Z.Fv[i], Z.Rf[i], Z.F[i]);
            break;
        case TK_B2: pz_cross_const<NT>(S, Z.C1[i], Z.F[i], rm.com[i], true); break;
        case TK_B3:
            pz_mul<NT, 9, 3, 3>(S, T[0], Z.R[i + 1], Z.Nv[i + 1]);
            pz_add3<NT>(S, T[1], Z.N[i], T[0]);
            pz_add3<NT>(S, T[2], T[1], Z.C1[i]);
            pz_add3<NT>(S, Z.Nv[i], T[2], Z.C2[i]);
            break;
        case TK_B4:
            if (axis != 0) {
                pz_merge<NT, 3, 1, 1>(S, Z.u0[i], view_extract(Z.Nv[i], row), view_scaled(Z.qdda[i], rm.armature[i]), false);
                pz_merge<NT, 1, 1, 1>(S, Z.u[i], view(Z.u0[i]), view_scaled(Z.qd[i], rm.damping[i]), false);
            }
            export_torque<NT>(S, tb, rec0 + i, Z.u[i]);
            break;
        case TK_K:
            pz_const_right<NT>(S, T[0], Z.FKR[i], rm.trans[i]);            // FK_R * P
            pz_add3<NT>(S, Z.FKT[i + 1], Z.FKT[i], T[0]);                  // FK_T = FK_T + FK_R * P
            pz_mul<NT, 9, 9, 9>(S, Z.FKR[i + 1], Z.FKR[i], Z.R[i]);        // FK_R = FK_R * R_i
            pz_mul<NT, 9, 3, 3>(S, T[1], Z.FKR[i + 1], Z.link0[i]);        // FK_R * link_i
            pz_add3<NT>(S, Z.LINK[i], T[1], Z.FKT[i + 1]);                 //          + FK_T
            export_link<NT>(S, tb, rec0 + i, Z.LINK[i]);
            break;
    }
}

// Warp 0 of the calling group picks the lowest-numbered free task whose dependencies are done and claims it.
// Returns the task id, -1 when no free task is left, -2 when free tasks exist but none is ready yet.
template <int NT>
__device__ int pick_task(Scratch& S, volatile int* tstate, const TaskDeps* deps) {
    if (gtid<NT>() < 32) {
        const int lane = gtid<NT>();
        int result = -1;
        for (int base = 0; base < N_TASKS; base += 32) {
            const int k = base + lane;
            bool free_ = false, ready = false;
            if (k < N_TASKS) {
                free_ = tstate[k] == 0;
                if (free_) {
                    const TaskDeps d = deps[k];
                    ready = true;
#pragma unroll
                    for (int j = 0; j < 4; j++) if (d.d[j] >= 0 && tstate[d.d[j]] != 2) ready = false;
                }
            }
            unsigned m = __ballot_sync(0xffffffffu, ready);
            const bool any_free = __ballot_sync(0xffffffffu, free_) != 0;
            while (m) {
                const int first = __ffs(m) - 1;
                int won = 0;
                if (lane == first) won = atomicCAS((int*)&tstate[base + first], 0, 1) == 0;
                won = __shfl_sync(0xffffffffu, won, first);
                if (won) { result = base + first; break; }
                m &= m - 1;
            }
            if (result >= 0) break;
            if (any_free) result = -2;
        }
        if (lane == 0) { if (result >= 0) __threadfence_block(); S.iscan[33] = result; }
    }
    gsync<NT>();
    const int r = S.iscan[33];
    gsync<NT>();
    return r;
}

template <int NT, int G>
__global__ void __launch_bounds__(NT * G, 1) reach_task_kernel(Tables tb, char* arena, size_t arena_stride, int mcap, int ncap, int scap, int tcap, int n_work) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Scratch SS[G];
    __shared__ volatile int tstate[N_TASKS];
    __shared__ TaskDeps deps[N_TASKS];
    const RobotModel& rm = c_robot;
    const int group = threadIdx.x / NT;
    Scratch& S = SS[group];
    // the PZ descriptors (~32 KB) live in dynamic shared memory behind the groups' sort buffers
    TaskSlots<G>& Z = *reinterpret_cast<TaskSlots<G>*>(smem_raw + (size_t)G * Scratch::smem_bytes(scap, tcap));
    PZ<3>* T = Z.T[group];
    if (threadIdx.x == 0) {
        char* g = arena + (size_t)blockIdx.x * arena_stride;
        for (int i = 0; i <= NJ; i++) { g = carve<3>(Z.W[i], g, mcap); g = carve<3>(Z.WA[i], g, mcap); g = carve<3>(Z.WD[i], g, mcap); g = carve<3>(Z.LA[i], g, mcap); }
        for (int i = 0; i < NJ; i++) {
            g = carve<3>(Z.WAm[i], g, mcap); g = carve<3>(Z.X[i], g, mcap); g = carve<3>(Z.LT0[i], g, mcap); g = carve<3>(Z.LT2[i], g, mcap);
            g = carve<3>(Z.FT0[i], g, mcap); g = carve<3>(Z.FT2[i], g, mcap);
            g = carve<3>(Z.F[i], g, mcap); g = carve<3>(Z.N[i], g, mcap); g = carve<3>(Z.Rf[i], g, mcap); g = carve<3>(Z.C1[i], g, mcap); g = carve<3>(Z.C2[i], g, mcap);
            g = carve<3>(Z.LINK[i], g, mcap); g = carve<3>(Z.link0[i], g, SMALL_CAP);
        }
        for (int i = 0; i <= NJ; i++) { g = carve<3>(Z.Fv[i], g, mcap); g = carve<3>(Z.Nv[i], g, mcap); g = carve<3>(Z.FKT[i], g, mcap); g = carve<9>(Z.FKR[i], g, mcap); }
        for (int k = 0; k < G; k++) for (int t = 0; t < 4; t++) g = carve<3>(Z.T[k][t], g, mcap);
        g = carve<3>(Z.Zero, g, SMALL_CAP);
        for (int i = 0; i <= NJ; i++) g = carve<9>(Z.R[i], g, SMALL_CAP);
        for (int i = 0; i < NJ; i++) g = carve<9>(Z.Rt[i], g, SMALL_CAP);
        for (int i = 0; i < NJ; i++) {
            g = carve<1>(Z.qd[i], g, SMALL_CAP); g = carve<1>(Z.qda[i], g, SMALL_CAP); g = carve<1>(Z.qdda[i], g, SMALL_CAP);
            g = carve<1>(Z.cosq[i], g, SMALL_CAP); g = carve<1>(Z.sinq[i], g, SMALL_CAP);
            g = carve<1>(Z.u[i], g, mcap); g = carve<1>(Z.u0[i], g, mcap);
        }
        const double thr_sq = squared_threshold(tb.thr);
        for (int k = 0; k < G; k++) {
            SS[k].bind(smem_raw + (size_t)k * Scratch::smem_bytes(scap, tcap), scap, tcap, g + (size_t)k * Scratch::gmem_bytes(ncap), ncap);
            SS[k].thr = tb.thr; SS[k].thr_sq = thr_sq; SS[k].gerr = tb.err;
        }
    }
    for (int k = threadIdx.x; k < N_TASKS; k += NT * G) {
        TaskDeps d = {{-1, -1, -1, -1}};
        if (k == TASK_M) d.d[0] = task_id(TK_B4, 0);
        else d = task_deps(k / NJ, k % NJ);
        deps[k] = d;
    }
    __syncthreads();
    for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
        const int prob = work / tb.T, s = work - prob * tb.T;
        const size_t rec0 = ((size_t)prob * tb.T + s) * NJ;
        for (int k = threadIdx.x; k < N_TASKS; k += NT * G) tstate[k] = 0;
        // ---- stage A: joint reach sets (one thread per joint; tiny scalar work) -------------------
        if (threadIdx.x < NJ) {
            const int i = threadIdx.x;
            make_poly_zono_joint(tb, prob, s, i, Z.R[i], Z.Rt[i], Z.qd[i], Z.qda[i], Z.qdda[i], Z.cosq[i], Z.sinq[i], S.thr_sq);
            if (tb.traj) {
                SmallRec* rec = tb.traj + (((size_t)prob * tb.T + s) * TRAJ_TABLES) * NJ;
                export_small<1>(rec[TRAJ_COS * NJ + i], Z.cosq[i]); export_small<1>(rec[TRAJ_SIN * NJ + i], Z.sinq[i]);
                export_small<9>(rec[TRAJ_R * NJ + i], Z.R[i]); export_small<9>(rec[TRAJ_RT * NJ + i], Z.Rt[i]);
                export_small<1>(rec[TRAJ_QD * NJ + i], Z.qd[i]); export_small<1>(rec[TRAJ_QDA * NJ + i], Z.qda[i]);
                export_small<1>(rec[TRAJ_QDDA * NJ + i], Z.qdda[i]);
            }
            PZ<3>& L0 = Z.link0[i];   // original link boxes (KPR/Dynamics.cu:51-66), as in reach_build_kernel
            int n = 0;
            double ind[3] = {0, 0, 0}, abss[3] = {0, 0, 0};
            const u64 gk[3] = {key_qde(0), key_qdae(0), key_qddae(0)};
            for (int j = 0; j < 3; j++) {
                const double gj = rm.link_g[i][j];
                if (norm1(&gj) > S.thr_sq) {
                    L0.keys[n] = gk[j];
                    for (int c = 0; c < 3; c++) L0.coef[c * L0.cap + n] = (c == j) ? gj : 0.0;
                    abss[j] = fabs(gj);
                    n++;
                }
                else ind[j] = fabs(gj);
            }
            L0.n = n; L0.divM = FastDiv::magic(n);
            for (int c = 0; c < 3; c++) { L0.center[c] = rm.link_c[i][c]; L0.ind[0][c] = ind[c]; L0.ind[1][c] = ind[c]; L0.abss[c] = abss[c]; }
            slot_zero(Z.u[i]); slot_zero(Z.u0[i]);
        }
        if (threadIdx.x == NJ) {   // R(NUM_JOINTS) = identity, initial RNEA state (KPR/Dynamics.cu:87-99), FK start (:72-73)
            PZ<9>& R = Z.R[NJ];
            slot_zero(R);
            for (int c = 0; c < 9; c++) R.center[c] = rm.R0[NJ][c];
            slot_zero(Z.W[0]); slot_zero(Z.WA[0]); slot_zero(Z.WD[0]); slot_zero(Z.LA[0]); slot_zero(Z.Fv[NJ]); slot_zero(Z.Nv[NJ]); slot_zero(Z.Zero);
            Z.LA[0].center[2] = rm.gravity;
            slot_zero(Z.FKR[0]);
            for (int c = 0; c < 9; c++) Z.FKR[0].center[c] = rm.R0[NJ][c];
            slot_zero(Z.FKT[0]);
        }
        __syncthreads();

        int idle = 0;
#ifdef ARMOUR_TASK_TIMING
        long long t_busy = 0, t_start = clock64(), t_type[TK_TYPES + 1] = {0};
        int n_run = 0;
#endif
        for (;;) {
            const int k = pick_task<NT>(S, tstate, deps);
            if (k == -1) break;
            if (k == -2) {   // nothing ready yet; bounded, so that a scheduling bug becomes an error word and never a hang
                if (++idle > (1 << 21)) { if (gtid<NT>() == 0) set_err(S, ERR_SYNC); break; }
                __nanosleep(100);
                continue;
            }
            if (k == TASK_M) {
                if (gtid<NT>() == 0) {   // stage M: torque radius (KPR/armour_main.cu:176-205); rho: only the upper end is used
                    double rho = 0.0;
                    for (int i = 0; i < NF; i++) { const double r = tb.dist_rad[rec0 + i]; rho = __dadd_ru(rho, __dmul_ru(r, r)); }
                    rho = __dsqrt_ru(rho);
                    const double c0 = __dmul_ru(__dmul_ru(rm.alpha, __dadd_ru(rm.M_max, -rm.M_min)), rm.eps);
                    for (int i = 0; i < NF; i++) {
                        double tr = __dadd_ru(c0, __dmul_ru(0.5, tb.dist_rad[rec0 + i]));
                        tr = __dadd_ru(tr, __dmul_ru(0.5, rho));
                        tr = __dadd_ru(tr, tb.u_ind[rec0 + i]);
                        tr = __dadd_ru(tr, rm.friction[i]);
                        tb.torque_radius[rec0 + i] = tr;
                    }
                }
            }
            else {
#ifdef ARMOUR_TASK_TIMING
                const long long t0 = clock64();
#endif
                run_task<NT, G>(S, Z, T, tb, rec0, k / NJ, k % NJ);
#ifdef ARMOUR_TASK_TIMING
                const long long dt = clock64() - t0;
                t_busy += dt; t_type[k / NJ] += dt; n_run++;
#endif
            }
            gsync<NT>();
            if (gtid<NT>() == 0) { __threadfence_block(); tstate[k] = 2; }
        }
#ifdef ARMOUR_TASK_TIMING
        if (blockIdx.x == 64 && gtid<NT>() == 0) {
            printf("TASKS group %d: total %lld busy %lld tasks %d |", group, clock64() - t_start, t_busy, n_run);
            for (int t = 0; t < TK_TYPES; t++) printf(" %lld", t_type[t] / 1000);
            printf(" (kcycles per type)\n");
        }
#endif
        __syncthreads();
    }
}

template <int G> size_t task_smem_bytes(int scap, int tcap) { return (size_t)G * Scratch::smem_bytes(scap, tcap) + sizeof(TaskSlots<G>); }

template <int NT, int G>
static cudaError_t launch_tasks_variant(const Tables& tb, char* arena, size_t arena_stride, int mcap, int ncap, int scap, int tcap, int n_work, int grid, cudaStream_t stream) {
    const size_t smem = task_smem_bytes<G>(scap, tcap);
    cudaError_t e = cudaFuncSetAttribute(reach_task_kernel<NT, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    reach_task_kernel<NT, G><<<grid, NT * G, smem, stream>>>(tb, arena, arena_stride, mcap, ncap, scap, tcap, n_work);
    return cudaGetLastError();
}
// four groups of 128 threads, one CTA per SM (2 x 256 and 8 x 64 were measured slower and are not instantiated)
cudaError_t launch_reach_tasks(const Tables& tb, char* arena, size_t arena_stride, int mcap, int ncap, int scap, int tcap, int n_work, int grid, int groups, cudaStream_t stream) {
    (void)groups;
    return launch_tasks_variant<128, 4>(tb, arena, arena_stride, mcap, ncap, scap, tcap, n_work, grid, stream);
}
bool reach_tasks_fit(int groups, int scap, int tcap) {
    (void)groups;
    return task_smem_bytes<4>(scap, tcap) + (size_t)4 * 4700 + 2048 <= 227 * 1024;   // static shared memory (Scratch per group, flags) on top
}

}  // namespace armour
