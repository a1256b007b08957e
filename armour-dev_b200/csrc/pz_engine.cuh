// CTA-cooperative sparse polynomial-zonotope algebra for sm_100a.
//
// One CTA owns one time interval's whole FK + RNEA computation.  A polynomial zonotope (PZ) is a
// structure of arrays in the CTA's private slice of HBM (L1/L2 resident): packed 63-bit monomial
// keys (uint64, ascending, unique) beside fp64 coefficient planes, plus centre, two interval
// radii ("independent", one for the nominal and one for the uncertain inertial parameters) and
// the running sum of |coefficients|.
//
// What the reference does per operation (KPR/PZsparse.cu:864-994 operator*, :743-834 +/-,
// :1068-1167 addOneDimPZ / stack / cross, :284-350 simplify) is: build a list of candidate
// monomials, std::sort it by key, add up equal keys left to right, and move every monomial whose
// coefficient norm is <= SIMPLIFY_THRESHOLD into the interval radius.  Here the candidate list
// is never materialised with its coefficients.  Only (key, origin-index) pairs are generated, in
// runs that are already sorted because both operands are sorted, and merged with a rank-based
// merge sort in shared memory; ties break on the origin index, which makes the order identical
// to a stable sort of the reference's list.  One thread per distinct key then walks its segment
// in that order and recomputes each coefficient product from the operands on the fly, so the
// fp64 sums are rounded exactly like the CPU restatement's (compile with -fmad=false).
//
// cross(PZ, PZ) — six scalar multiplies, three subtractions and a stack in the reference, ten
// sort+merge passes — is one sort here: the six products share the same key multiset, so one
// sorted order serves six accumulators per key and the three threshold stages are applied per
// key in registers (bit-exact keep/drop decisions, see cross_pp()).
//
// Interval radii are accumulated with round-up intrinsics (__dadd_ru/__dmul_ru) so that a
// different summation order can only enlarge them (SURVEY.md §7 "hard parts").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace armour {

typedef unsigned long long u64;
typedef unsigned short u16;

// monomial key layout: KPR/PZsparse.h:23-40
__device__ __forceinline__ constexpr u64 key_k(int j) { return 1ull << (2 * j); }
__device__ __forceinline__ constexpr u64 key_qde(int j) { return 1ull << (14 + j); }
__device__ __forceinline__ constexpr u64 key_qdae(int j) { return 1ull << (21 + j); }
__device__ __forceinline__ constexpr u64 key_qddae(int j) { return 1ull << (28 + j); }
__device__ __forceinline__ constexpr u64 key_cosqe(int j) { return 1ull << (35 + 2 * j); }
__device__ __forceinline__ constexpr u64 key_sinqe(int j) { return 1ull << (49 + 2 * j); }
static constexpr u64 KEY_K_ONLY = 1ull << 14;         // key < this  <=> depends on k only   (PZsparse.h:38)
static constexpr u64 KEY_K_LINKS_ONLY = 1ull << 35;   // (PZsparse.h:40)
static constexpr u64 KEY_K_MASK = KEY_K_ONLY - 1;

enum { ERR_NONE = 0, ERR_ENTRY_CAP = 1, ERR_MONO_CAP = 2, ERR_TABLE_CAP = 4, ERR_LINK_GEN = 8, ERR_DEGREE = 16, ERR_SYNC = 32, ERR_LTABLE_CAP = 64, ERR_CANARY = 128 };

// Degree-overflow guard for key addition.  The reference adds keys without checking ("do not have to check carry",
// KPR/PZsparse.cu:938-940): a degree that outgrows its field (3 for the 2-bit fields, 1 for the 1-bit ones) silently corrupts
// the neighbouring variable.  The carries INTO each bit of a + b are (a + b) ^ a ^ b; a carry into the first bit of a field (or
// out of bit 62) is exactly an overflow of the field below it, the carry inside a 2-bit field (1 + 1 = 2) is legitimate.
__host__ __device__ constexpr u64 field_start_mask() {
    u64 m = 0;
    for (int j = 1; j < 7; j++) m |= 1ull << (2 * j);            // k_1 .. k_6
    for (int b = 14; b <= 35; b++) m |= 1ull << b;               // qde, qdae, qddae (1 bit each) and cosqe_0
    for (int j = 1; j < 7; j++) m |= 1ull << (35 + 2 * j);       // cosqe_1 .. cosqe_6
    for (int j = 0; j < 7; j++) m |= 1ull << (49 + 2 * j);       // sinqe_0 .. sinqe_6
    m |= 1ull << 63;                                             // out of sinqe_6
    return m;
}
static constexpr u64 KEY_FIELD_STARTS = field_start_mask();
__device__ __forceinline__ u64 key_add_checked(u64 a, u64 b, u64& bad) {
    const u64 s = a + b;
#ifndef ARMOUR_NO_DEGREE_CHECK
    bad |= (s ^ a ^ b) & KEY_FIELD_STARTS;
#endif
    return s;
}

template <int D>
struct PZ {
    int n;          // monomials
    int cap;        // capacity (plane stride)
    u64* keys;      // [cap]
    double* coef;   // [D][cap], column-major element order inside a 3x3 (row + 3*col), like Eigen
    double center[D];
    double ind[2][D];   // interval radius: [0] nominal inertial parameters, [1] uncertain ones
    double abss[D];     // sum_i |coef_i| rounded up
    u64 divM;           // FastDiv magic for division by n (kept with the descriptor: one division per op, off the critical path)
    u64 ormask;         // superset of the OR of all keys (union of the operands' masks: thresholding only removes monomials); lets a
                        // product prove that its operands share no variable (pz_mul_structured)
};

// per-CTA scratch (shared memory) ------------------------------------------------------------
struct Scratch {
    // Sort ping-pong buffers live in dynamic shared memory.  They are kept as shared-window addresses and
    // turned back into pointers with __cvta_shared_to_generic at each use, so that the compiler emits
    // LDS/STS instead of generic loads (a pointer loaded from a struct has no known address space).
    // Operations with more than `scap` candidate monomials (a few per interval) use the same-shaped buffers in
    // global memory instead: keeping the shared footprint small leaves the SM's L1 to the PZ working set.
    unsigned skey_a[2];   // shared u64[scap]
    unsigned sidx_a[2];   // shared u16[scap]
    unsigned stmp_a;      // shared double[3][tcap]: per-key results before compaction (small operations)
    unsigned red_a;       // shared double[nwarps][32]: per-warp partial sums of block_scan_sum
    u64* gkey[2];         // global u64[ncap]
    u16* gidx[2];         // global u16[ncap]
    double* tmp;          // global double[9][ncap]: per-key results before compaction (large operations)
    int scap, tcap, ncap;
    int no_structured;   // 1: products always take the generic sort path (A/B measurements, the parity test of the structured path)
    double thr;
    double thr_sq;    // largest x with RN(sqrt(x)) <= thr: "norm <= thr" is tested as "squared norm <= thr_sq", exactly
    int* gerr;        // global error word
    int iscan[18];         // per-warp scan totals (<= 16 warps); [17] is a group-uniform flag (group_ready)
#ifdef ARMOUR_KEYHASH
    int dbg_work, dbg_seq;   // measurement build: work item and running operation number (see g_keyhash)
#endif
    __device__ __forceinline__ u64* skey(int b) const { return (u64*)__cvta_shared_to_generic((size_t)skey_a[b]); }
    __device__ __forceinline__ u16* sidx(int b) const { return (u16*)__cvta_shared_to_generic((size_t)sidx_a[b]); }
    __device__ __forceinline__ double* stmp() const { return (double*)__cvta_shared_to_generic((size_t)stmp_a); }
    __device__ __forceinline__ double* red() const { return (double*)__cvta_shared_to_generic((size_t)red_a); }
    // scap and tcap must be even (8-byte alignment of the regions that follow the 2-byte index buffers)
    static __host__ __device__ size_t smem_bytes(int scap_, int tcap_, int nwarps) { return (size_t)scap_ * 20 + (size_t)3 * tcap_ * 8 + (size_t)nwarps * 32 * 8; }
    static __host__ __device__ size_t gmem_bytes(int ncap_) { return (size_t)ncap_ * 20 + (size_t)9 * ncap_ * 8; }
    __device__ void bind(unsigned char* smem, int scap_, int tcap_, char* gmem, int ncap_) {
        const unsigned base = (unsigned)__cvta_generic_to_shared(smem);
        skey_a[0] = base; skey_a[1] = base + (unsigned)scap_ * 8;
        sidx_a[0] = base + (unsigned)scap_ * 16; sidx_a[1] = base + (unsigned)scap_ * 18;
        stmp_a = base + (unsigned)scap_ * 20;
        red_a = stmp_a + (unsigned)tcap_ * 24;
        scap = scap_; tcap = tcap_; ncap = ncap_; no_structured = 0;
        gkey[0] = (u64*)gmem; gkey[1] = gkey[0] + ncap_;
        gidx[0] = (u16*)(gkey[1] + ncap_); gidx[1] = gidx[0] + ncap_;
        tmp = (double*)(gmem + (size_t)ncap_ * 20);
    }
    // per-key staging buffer and its plane stride for an operation with n candidates and D output components
    __device__ __forceinline__ double* staging(int n, int D, int& stride) const {
        if (D <= 3 && n <= tcap) { stride = tcap; return stmp(); }
        stride = ncap; return tmp;
    }
};
// sort buffers of an operation: shared (BIG = false) or global (BIG = true)
template <bool BIG> struct Buf {
    static __device__ __forceinline__ u64* key(const Scratch& S, int b) { return BIG ? S.gkey[b] : S.skey(b); }
    static __device__ __forceinline__ u16* idx(const Scratch& S, int b) { return BIG ? S.gidx[b] : S.sidx(b); }
};

// unified 3-vector operation (defined near the end of this file); the named operations below forward to it
#ifndef ARMOUR_UNIFIED_OP3
#define ARMOUR_UNIFIED_OP3 1
#endif
template <int NT> __device__ __forceinline__ void op3_mul93(Scratch& S, PZ<3>& dst, const PZ<9>& A, const PZ<3>& B);
template <int NT> __device__ __forceinline__ void op3_add(Scratch& S, PZ<3>& dst, const PZ<3>& A, const PZ<3>& B);
template <int NT> __device__ __forceinline__ void op3_add_one_dim(Scratch& S, PZ<3>& dst, const PZ<3>& A, const PZ<1>& s, int row);
template <int NT> __device__ __forceinline__ void op3_cross(Scratch& S, PZ<3>& dst, const PZ<3>& A, const PZ<3>& B);
template <int NT> __device__ __forceinline__ void op3_cross_const(Scratch& S, PZ<3>& dst, const PZ<3>& Z, const double* kvec, bool const_first);
template <int NT> __device__ __forceinline__ void op3_const_left(Scratch& S, PZ<3>& dst, const double* Mc, const double* Mi0, const double* Mi1, bool scalar, const PZ<3>& V);
template <int NT> __device__ __forceinline__ void op3_const_right(Scratch& S, PZ<3>& dst, const PZ<9>& R, const double* pvec);

// exact n / d for n, d < 2^16 with one multiply: q = (n * M) >> 32, M = floor((2^32 - 1) / d) + 1
struct FastDiv {
    u64 M;
    // (a shared non-inlined copy of this 32-bit division was measured slower: the call's register traffic sits in every operation's epilogue)
    __device__ __forceinline__ static u64 magic(int d) { return (u64)(0xFFFFFFFFu / (unsigned)(d > 0 ? d : 1)) + 1; }
    __device__ __forceinline__ explicit FastDiv(u64 M_) : M(M_) {}
    __device__ __forceinline__ int div(int n) const { return (int)(((u64)(unsigned)n * M) >> 32); }
};

__device__ __forceinline__ void set_err(Scratch& S, int e) { atomicOr(S.gerr, e); }

// A CTA may hold several independent groups of NT threads (reach_kernels.cu runs the RNEA joint chain and the
// force/FK computations of one interval side by side).  Every cooperative routine below works within one group:
// gtid<NT>() is the thread's index in its group, gsync<NT>() a named barrier (id 1 + group) over the group's NT
// threads.  With a single group this is __syncthreads() under another name.
template <int NT> __device__ __forceinline__ int gtid() { return (int)(threadIdx.x & (NT - 1)); }
template <int NT> __device__ __forceinline__ void gsync() {
    if (NT == 32) __syncwarp();   // a one-warp group: warp-level barrier with memory ordering among its lanes
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x / NT)), "n"(NT) : "memory");
}

// sqrt is monotone, so {x : RN(sqrt(x)) <= thr} is a down-set {x <= X}.  Find X once per kernel; the hot loops then
// compare squared norms and never take a square root, with bit-identical keep/drop decisions.
__device__ inline double squared_threshold(double thr) {
    double x = __dmul_rn(thr, thr);
    for (int i = 0; i < 8 && __dsqrt_rn(x) > thr; i++) x = __longlong_as_double(__double_as_longlong(x) - 1);
    for (int i = 0; i < 8; i++) {
        const double up = __longlong_as_double(__double_as_longlong(x) + 1);
        if (__dsqrt_rn(up) <= thr) x = up; else break;
    }
    return x;
}

// Optional per-phase cycle accounting (profiling builds only: -DARMOUR_PHASE_TIMING).  Thread 0 of every CTA
// charges the cycles since the previous mark to a phase; armour_phase_cycles[] is summed over CTAs.
enum { PH_FILL = 0, PH_SORT = 1, PH_SEGMENT = 2, PH_COMPACT = 3, PH_ELEMENTWISE = 4, PH_STAGE_A = 5, PH_EXPORT = 6, PH_OTHER = 7, PH_COUNT = 8 };
#ifdef ARMOUR_PHASE_TIMING
__device__ unsigned long long armour_phase_cycles[PH_COUNT];
__device__ unsigned long long armour_phase_calls[PH_COUNT];
struct PhaseClock { long long last; };
__shared__ PhaseClock g_phase_clock;
__device__ __forceinline__ void phase_mark(int ph) {
    if ((threadIdx.x & 255) == 0 && threadIdx.x < 256) {
        const long long t = clock64();
        atomicAdd(&armour_phase_cycles[ph], (unsigned long long)(t - g_phase_clock.last));
        atomicAdd(&armour_phase_calls[ph], 1ull);
        g_phase_clock.last = t;
    }
}
#else
__device__ __forceinline__ void phase_mark(int) {}
#endif

// ---- small fp helpers (round-to-nearest, no contraction; -fmad=false is also set) -----------
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
// squared Frobenius norms in Eigen 3.3's reduction order (see oracle/oracle_pz.hpp Mat::squaredNorm)
// squared norms; compare against Scratch::thr_sq
__device__ __forceinline__ double norm1(const double* v) { return mul_rn(v[0], v[0]); }
__device__ __forceinline__ double norm3(const double* v) {
    return add_rn(add_rn(mul_rn(v[0], v[0]), mul_rn(v[1], v[1])), mul_rn(v[2], v[2]));
}
__device__ __forceinline__ double norm9(const double* v) {
    double p0a = add_rn(mul_rn(v[0], v[0]), mul_rn(v[4], v[4]));
    double p0b = add_rn(mul_rn(v[1], v[1]), mul_rn(v[5], v[5]));
    double p1a = add_rn(mul_rn(v[2], v[2]), mul_rn(v[6], v[6]));
    double p1b = add_rn(mul_rn(v[3], v[3]), mul_rn(v[7], v[7]));
    double s = add_rn(add_rn(p0a, p1a), add_rn(p0b, p1b));
    return add_rn(s, mul_rn(v[8], v[8]));
}
template <int D>
__device__ __forceinline__ double normD(const double* v) {
    if (D == 1) return norm1(v);
    if (D == 3) return norm3(v);
    return norm9(v);
}
// 3x3 (column-major) times 3-vector, k ascending like Eigen's coefficient-based product
__device__ __forceinline__ void matvec_rn(const double* M, const double* x, double* y) {
#pragma unroll
    for (int r = 0; r < 3; r++) y[r] = add_rn(add_rn(mul_rn(M[r], x[0]), mul_rn(M[r + 3], x[1])), mul_rn(M[r + 6], x[2]));
}
__device__ __forceinline__ void matmat_rn(const double* A, const double* B, double* C) {
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++)
            C[r + 3 * c] = add_rn(add_rn(mul_rn(A[r], B[3 * c]), mul_rn(A[r + 3], B[3 * c + 1])), mul_rn(A[r + 6], B[3 * c + 2]));
}
__device__ __forceinline__ void matvec_ru(const double* M, const double* x, double* y) {   // non-negative operands
#pragma unroll
    for (int r = 0; r < 3; r++) y[r] = __dadd_ru(__dadd_ru(__dmul_ru(M[r], x[0]), __dmul_ru(M[r + 3], x[1])), __dmul_ru(M[r + 6], x[2]));
}
__device__ __forceinline__ void matmat_ru(const double* A, const double* B, double* C) {
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++)
            C[r + 3 * c] = __dadd_ru(__dadd_ru(__dmul_ru(A[r], B[3 * c]), __dmul_ru(A[r + 3], B[3 * c + 1])), __dmul_ru(A[r + 6], B[3 * c + 2]));
}
// A parallel round-up sum of n non-negative terms is >= the exact sum; the reference's sequential
// round-to-nearest sum is <= exact * (1 + n * 2^-53).  Inflating by (1 + (n + 2) * 2^-52) therefore
// dominates the reference's value whatever the order.
__device__ __forceinline__ double inflate(double s, int n) { return __dmul_ru(s, __dadd_ru(1.0, (double)(n + 2) * 0x1p-52)); }

// ---- block-wide aggregation ---------------------------------------------------------------------
// Warp-level sum of V values per lane, rounded up, with a transposing butterfly: at each of the first
// log2(VP) steps a lane keeps one half of its values and trades the other half with its partner, so the
// whole reduction costs VP - 1 + (5 - log2 VP) exchanges instead of 5 * V.  Afterwards lane L holds in
// v[0] the warp total of value index L >> (5 - log2 VP).  The summation order is fixed (deterministic).
template <int VP>
__device__ __forceinline__ void warp_multi_sum_ru(double (&v)[VP]) {
    const int lane = threadIdx.x & 31;
    int off = 16;
#pragma unroll
    for (int half = VP / 2; half >= 1; half >>= 1, off >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; i++) {
            const double send = hi ? v[i] : v[i + half];
            const double keep = hi ? v[i + half] : v[i];
            v[i] = __dadd_ru(keep, __shfl_xor_sync(0xffffffffu, send, off));
        }
    }
#pragma unroll
    for (; off >= 1; off >>= 1) v[0] = __dadd_ru(v[0], __shfl_xor_sync(0xffffffffu, v[0], off));
}
template <int V> struct Pow2Ceil { static constexpr int value = V <= 1 ? 1 : V <= 2 ? 2 : V <= 4 ? 4 : V <= 8 ? 8 : V <= 16 ? 16 : 32; };
template <int VP> struct Log2 { static constexpr int value = VP == 1 ? 0 : VP == 2 ? 1 : VP == 4 ? 2 : VP == 8 ? 3 : VP == 16 ? 4 : 5; };

// One barrier: exclusive scan of `cnt` over the block and block-wide round-up sums of red[0..V).
// Returns this thread's scan offset; `total` is the block total.  After the call, value k's per-warp partial
// sums are in S.red[w * VP + k] for w < NT/32 (block_total() adds them up).
template <int NT, int V>
__device__ __forceinline__ int block_scan_sum(Scratch& S, int cnt, const double (&red)[V], int& total) {
    constexpr int VP = Pow2Ceil<V>::value;
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = gtid<NT>() >> 5;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    // (a plain xor-butterfly per value is a quarter of the code but three times the shuffles: measured 5 % slower in the sweep, 8 % on one plan)
    double v[VP];
#pragma unroll
    for (int k = 0; k < VP; k++) v[k] = k < V ? red[k] : 0.0;
    warp_multi_sum_ru<VP>(v);
    if (lane == 31) S.iscan[warp] = incl;
    if ((lane & ((32 >> Log2<VP>::value) - 1)) == 0) S.red()[warp * VP + (lane >> (5 - Log2<VP>::value))] = v[0];
    gsync<NT>();
    int before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) { const int t = S.iscan[w]; all += t; if (w < warp) before += t; }
    total = all;
    return before + incl - cnt;
}
template <int NT, int V>
__device__ __forceinline__ double block_total(const Scratch& S, int k) {
    constexpr int VP = Pow2Ceil<V>::value;
    double s = 0.0;
    const double* red = S.red();
#pragma unroll
    for (int w = 0; w < NT / 32; w++) s = __dadd_ru(s, red[w * VP + k]);
    return s;
}

// Merge levels with at least this many candidates per thread use the merge-path merge (one diagonal search per thread,
// then a sequential merge of the thread's output chunk) instead of one binary search per candidate.
#ifndef ARMOUR_MERGE_PATH_MIN
#define ARMOUR_MERGE_PATH_MIN 2
#endif
template <int NT> struct MergePathPolicy {
    static __device__ __forceinline__ bool use(int N) {
        if (NT >= 256) return N > NT;
        if (NT == 128) return false;   // measured in round 1: the per-candidate search is 2 % faster at 128 threads x 4 CTAs per SM
        return N > ARMOUR_MERGE_PATH_MIN * NT;
    }
};

// ---- run-structured merge sort on (key, idx) ------------------------------------------------
// Buffer 0 holds N entries laid out as sorted runs of width W (the last run may be shorter).
// Returns the buffer that holds the fully sorted sequence.  All (key, idx) pairs are distinct, so
// ordering by (key, idx) equals a stable sort by key of the list in origin order.
template <int NT, bool BIG>
__device__ __forceinline__ int merge_sort_runs(Scratch& S, int N, int W, u64 magicW) {
    int cur = 0;
    const FastDiv fd(magicW);
    int level = 0;
    for (int w = W; w < N; w <<= 1, level++) {
        const u64* ki = Buf<BIG>::key(S, cur);
        const u16* ii = Buf<BIG>::idx(S, cur);
        u64* ko = Buf<BIG>::key(S, cur ^ 1);
        u16* io = Buf<BIG>::idx(S, cur ^ 1);
        if (MergePathPolicy<NT>::use(N)) {   // more than MIN candidates per thread (one plan at 256 threads: -4 %; the narrow sweep shapes run several candidates per lane)
        // merge path: every thread owns a chunk of consecutive OUTPUT positions; one binary search along the chunk's
        // diagonal finds how many elements of each run precede it, then the chunk is merged sequentially.  (One search
        // per thread and level instead of one per candidate; order by (key, origin index), all pairs distinct.)
        {
            const int ipt = (N + NT - 1) / NT;
            int o = min(gtid<NT>() * ipt, N);
            const int oend = min(o + ipt, N);
            const int w2 = w << 1;
            #pragma unroll 1
            while (o < oend) {
                const int base = (fd.div(o) >> (level + 1)) * w2;
                const int la = min(w, N - base), lb = max(0, min(w, N - base - w));
                const int cend = min(oend, base + la + lb);
                const int bb = base + w;
                const int d = o - base;
                int lo = max(0, d - lb), hi = min(d, la);
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const u64 ka = ki[base + mid], kb = ki[bb + d - 1 - mid];
                    bool less = ka < kb;
                    if (ka == kb) less = ii[base + mid] < ii[bb + d - 1 - mid];
                    if (less) lo = mid + 1; else hi = mid;
                }
                int i = lo, j = d - lo;
                u64 ka = i < la ? ki[base + i] : ~0ull, kb = j < lb ? ki[bb + j] : ~0ull;
                #pragma unroll 1
                for (; o < cend; o++) {
                    bool ta;
                    if (j >= lb) ta = true;
                    else if (i >= la) ta = false;
                    else { ta = ka < kb; if (ka == kb) ta = ii[base + i] < ii[bb + j]; }
                    if (ta) { ko[o] = ka; io[o] = ii[base + i]; i++; ka = i < la ? ki[base + i] : ~0ull; }
                    else { ko[o] = kb; io[o] = ii[bb + j]; j++; kb = j < lb ? ki[bb + j] : ~0ull; }
                }
            }
        }
        }
        else {
        // rank of each element among its sibling run (binary search on the key; the origin index breaks exact ties)
        #pragma unroll 1
        for (int g = gtid<NT>(); g < N; g += NT) {
            const int r = fd.div(g) >> level;      // g / (W << level)
            const int base = r * w;
            const int sb = (r ^ 1) * w;
            const u64 k = ki[g];
            const unsigned id = ii[g];
            int pos = g;
            if (sb < N) {
                const int end = min(sb + w, N);
                int lo = sb, hi = end;
                while (lo < hi) {   // lower bound on the key alone
                    const int mid = (lo + hi) >> 1;
                    if (ki[mid] < k) lo = mid + 1; else hi = mid;
                }
                // exact ties: equal keys of a run are adjacent and ordered by origin index, so the entries that precede
                // (k, id) are a prefix of the equal range (one extra load per candidate instead of tie logic in every search step)
                while (lo < end && ki[lo] == k && ii[lo] < id) lo++;
                pos = min(base, sb) + (g - base) + (lo - sb);
            }
            ko[pos] = k;
            io[pos] = (u16)id;
        }
        }
        gsync<NT>();
        phase_mark(PH_SORT);
        cur ^= 1;
    }
    return cur;
}

#ifdef ARMOUR_OP_TRACE
__device__ long long g_tr[8];
#define TR(i) do { if (blockIdx.x == 64 && threadIdx.x == 0) g_tr[i] = clock64(); } while (0)
#else
#define TR(i)
#endif

// ---- generic "segment-reduce, threshold, compact" -------------------------------------------
// Op interface:
//   static constexpr int NACC;                          accumulators per key
//   void first(unsigned idx, double* acc);              acc  = contribution of the segment's first entry
//   void next (unsigned idx, double* acc);              acc += contribution (round-to-nearest, in order)
//   bool finish(const double* acc, double* out, double* drop);   threshold logic; out[DOUT]; drop[DOUT] = |dropped|
// Epi interface (scalar epilogue, run by thread c < DOUT only, off the other threads' critical path):
//   void operator()(int c, double& cen, double& r0, double& r1)   centre and the two radii of component c before the
//                                                                  dropped monomials are added, read from the operands
// dst may alias an operand: the epilogue lanes read the operands at entry (before any barrier) and write the
// descriptor only after the scan barrier, when every thread is done with the operands' descriptors.
// The epilogue is owned by lanes 0..DOUT-1 of the LAST warp (the warp most likely to be idle in the segment pass):
// begin() reads the operands before any barrier, finish() adds the block totals and writes the descriptor.
template <int NT, int DOUT>
struct ScalarEpilogue {
    double cen, r0, r1;
    u64 om;
    int c;
    template <class Epi>
    __device__ __forceinline__ void begin(const Epi& epi) {
        c = gtid<NT>() - (NT - 32);
        cen = 0; r0 = 0; r1 = 0; om = 0;
        if (c >= 0 && c < DOUT) { epi(c, cen, r0, r1); if (c == 0) om = epi.mask(); }
    }
    __device__ __forceinline__ void finish(const Scratch& S, PZ<DOUT>& dst, int n_in, int total) const {
        if (c >= 0 && c < DOUT) {
            const double drop = inflate(block_total<NT, 2 * DOUT>(S, c), n_in);
            dst.abss[c] = inflate(block_total<NT, 2 * DOUT>(S, DOUT + c), total);
            dst.center[c] = cen;
            dst.ind[0][c] = __dadd_ru(r0, drop);
            dst.ind[1][c] = __dadd_ru(r1, drop);
            if (c == 0) { dst.n = total; dst.divM = FastDiv::magic(total); dst.ormask = om; }
        }
    }
};

// Shared tail of every operation: compaction of the kept keys (blocked ranges keep the order), block-wide radius sums,
// descriptor update.  key / flag / tmp are indexed by sorted candidate position; one barrier inside, one at the end.
#ifdef ARMOUR_KEYHASH
// Measurement build (scripts/keyhash_experiment.py): per (work item, operation number) a hash of the RESULT's key list and the
// candidate count, to measure how often consecutive intervals of one problem produce identical key lists (the condition under
// which a cached sort order of the previous interval could be reused).
constexpr int KEYHASH_OPS = 512, KEYHASH_WORK = 512;
__device__ u64 g_keyhash[KEYHASH_WORK][KEYHASH_OPS][2];
template <int NT, int D>
__device__ __forceinline__ void keyhash_record(Scratch& S, const PZ<D>& dst, int N, int total) {
    if (gtid<NT>() == 0) {
        u64 hsh = 1469598103934665603ull ^ (u64)total;
        for (int i = 0; i < total; i++) hsh = (hsh ^ dst.keys[i]) * 1099511628211ull;
        if (S.dbg_work < KEYHASH_WORK && S.dbg_seq < KEYHASH_OPS) { g_keyhash[S.dbg_work][S.dbg_seq][0] = hsh; g_keyhash[S.dbg_work][S.dbg_seq][1] = (u64)N; }
        S.dbg_seq++;
    }
}
#endif
#ifdef ARMOUR_SEGSTAT
// Measurement build (scripts/segstat_experiment.py): how unevenly the per-key segment walk loads the lanes of a warp.  Per warp
// iteration: candidates, segment heads, longest segment (= the iteration's duration in term computations today) — summed in
// g_segstat[0..3] = {warp iterations, candidates, heads, sum of longest segment}.
__device__ unsigned long long g_segstat[4];
__device__ __forceinline__ void segstat_record(int N, int g, bool head, const u64* key) {
    const unsigned act = __activemask();
    int len = 0;
    if (head) { const u64 k = key[g]; len = 1; for (int e = g + 1; e < N && key[e] == k; e++) len++; }
    int mx = len;
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(act, mx, o));   // partial masks only in an operation's last iteration: approximate there
    const unsigned heads = __ballot_sync(act, head);
    if ((threadIdx.x & 31) == (__ffs(act) - 1)) {
        atomicAdd(&g_segstat[0], 1ull); atomicAdd(&g_segstat[1], (unsigned long long)__popc(act));
        atomicAdd(&g_segstat[2], (unsigned long long)__popc(heads)); atomicAdd(&g_segstat[3], (unsigned long long)mx);
    }
}
#endif
#ifndef ARMOUR_PREFETCH_DST
#define ARMOUR_PREFETCH_DST 0
#endif
template <int NT, int DOUT>
__device__ __forceinline__ void compact_emit(Scratch& S, int N, const u64* key, const u16* flag, const double* tmp, int ncap, const double (&red)[2 * DOUT],
                                             const ScalarEpilogue<NT, DOUT>& se, PZ<DOUT>& dst) {
    const int ipt = (N + NT - 1) / NT;
    const int g0 = min(gtid<NT>() * ipt, N), g1 = min(g0 + ipt, N);
    int cnt = 0;
    #pragma unroll 1
    for (int g = g0; g < g1; g++) cnt += flag[g];
    int total;
    int off = block_scan_sum<NT, 2 * DOUT>(S, cnt, red, total);
    const int dcap = dst.cap;
    if (total > dcap) { if (gtid<NT>() == 0) set_err(S, ERR_MONO_CAP); total = 0; }
    else {
        u64* dk = dst.keys;
        double* dc = dst.coef;
        #pragma unroll 1
        for (int g = g0; g < g1; g++) {
            if (flag[g]) {
                dk[off] = key[g];
#pragma unroll
                for (int c = 0; c < DOUT; c++) dc[c * dcap + off] = tmp[c * ncap + g];
                off++;
            }
        }
#if ARMOUR_PREFETCH_DST
        // stores go through to L2; pull the lines just written into L1, where the next operation (usually the consumer) reads them
        if (cnt > 0) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(dk + off - 1));
#pragma unroll
            for (int c = 0; c < DOUT; c++) asm volatile("prefetch.global.L1 [%0];" ::"l"(dc + c * dcap + off - 1));
        }
#endif
    }
    se.finish(S, dst, N, total);
    TR(6);
    gsync<NT>();
    TR(7);
#ifdef ARMOUR_KEYHASH
    keyhash_record<NT, DOUT>(S, dst, N, total);
#endif
    phase_mark(PH_COMPACT);
}

// Barriers: one after the segment pass, one inside block_scan_sum, one at the end.
template <int NT, int DOUT, bool BIG, class Op, class Epi>
__device__ __forceinline__ void reduce_emit(Scratch& S, int buf, int N, Op& op, PZ<DOUT>& dst, const Epi& epi) {
    const u64* key = Buf<BIG>::key(S, buf);
    const u16* idx = Buf<BIG>::idx(S, buf);
    u16* flag = Buf<BIG>::idx(S, buf ^ 1);
    int ncap;
    double* tmp = S.staging(N, DOUT, ncap);
    double red[2 * DOUT];
#pragma unroll
    for (int c = 0; c < 2 * DOUT; c++) red[c] = 0.0;
    ScalarEpilogue<NT, DOUT> se;
    se.begin(epi);
    TR(3);
    // pass 1: one thread per segment head
    #pragma unroll 1
    for (int g = gtid<NT>(); g < N; g += NT) {
        const u64 k = key[g];
        u16 f = 0;
#ifdef ARMOUR_SEGSTAT
        segstat_record(N, g, g == 0 || key[g - 1] != k, key);
#endif
        if (g == 0 || key[g - 1] != k) {
            double acc[Op::NACC];
            op.first(idx[g], acc);
            #pragma unroll 1
            for (int e = g + 1; e < N && key[e] == k; e++) op.next(idx[e], acc);
            double out[DOUT], dr[DOUT];
#pragma unroll
            for (int c = 0; c < DOUT; c++) dr[c] = 0.0;
            const bool keep = op.finish(acc, out, dr);
#pragma unroll
            for (int c = 0; c < DOUT; c++) red[c] = __dadd_ru(red[c], dr[c]);
            if (keep) {
                f = 1;
#pragma unroll
                for (int c = 0; c < DOUT; c++) { tmp[c * ncap + g] = out[c]; red[DOUT + c] = __dadd_ru(red[DOUT + c], fabs(out[c])); }
            }
        }
        flag[g] = f;
    }
    TR(4);
    gsync<NT>();
    TR(5);
    phase_mark(PH_SEGMENT);
    compact_emit<NT, DOUT>(S, N, key, flag, tmp, ncap, red, se, dst);
}

// elementwise variant: no sort, keys are those of `src` in order; op computes out from index i.
//   bool Op::finish(int i, double* out, double* drop)
template <int NT, int DOUT, bool BIG, class Op, class Epi>
__device__ __forceinline__ void elementwise_emit(Scratch& S, int n, const u64* src_keys, Op& op, PZ<DOUT>& dst, const Epi& epi) {
    u16* flag = Buf<BIG>::idx(S, 0);
    u64* kcopy = Buf<BIG>::key(S, 0);
    double red[2 * DOUT];
#pragma unroll
    for (int c = 0; c < 2 * DOUT; c++) red[c] = 0.0;
    if (n > S.ncap) { if (gtid<NT>() == 0) set_err(S, ERR_ENTRY_CAP); n = 0; }
    int ncap;
    double* tmp = S.staging(n, DOUT, ncap);
    ScalarEpilogue<NT, DOUT> se;
    se.begin(epi);
    #pragma unroll 1
    for (int i = gtid<NT>(); i < n; i += NT) {
        double out[DOUT], dr[DOUT];
#pragma unroll
        for (int c = 0; c < DOUT; c++) dr[c] = 0.0;
        const bool keep = op.finish(i, out, dr);
#pragma unroll
        for (int c = 0; c < DOUT; c++) red[c] = __dadd_ru(red[c], dr[c]);
        if (keep) {
#pragma unroll
            for (int c = 0; c < DOUT; c++) { tmp[c * ncap + i] = out[c]; red[DOUT + c] = __dadd_ru(red[DOUT + c], fabs(out[c])); }
        }
        flag[i] = keep ? 1 : 0;
        kcopy[i] = src_keys[i];   // dst may alias src
    }
    gsync<NT>();
    phase_mark(PH_ELEMENTWISE);
    compact_emit<NT, DOUT>(S, n, kcopy, flag, tmp, ncap, red, se, dst);
}

// load coefficient vector i through a cached (pointer, stride) pair: the descriptor lives in shared memory, and every
// shared-memory store in between would otherwise force the compiler to re-read coef / cap before each access
template <int D>
__device__ __forceinline__ void ldc(const double* __restrict__ coef, int cap, int i, double* v) {
    __builtin_assume(__isGlobal(coef));   // PZ storage is always in the global arena: LDG instead of generic LD
#pragma unroll
    for (int c = 0; c < D; c++) v[c] = coef[c * cap + i];
}
// load coefficient vector i of a PZ
template <int D>
__device__ __forceinline__ void ldc(const PZ<D>& z, int i, double* v) {
#pragma unroll
    for (int c = 0; c < D; c++) v[c] = z.coef[c * z.cap + i];
}

// =============================================================================================
// General product  dst = A * B   (KPR/PZsparse.cu:864-994)
//   shapes: <9,3,3> 3x3 * 3x1,  <9,9,9> 3x3 * 3x3,  <1,1,1> scalar * scalar
// =============================================================================================
template <int DA, int DB, int DO>
__device__ __forceinline__ void coef_mul(const double* a, const double* b, double* o) {
    if (DA == 9 && DB == 3) matvec_rn(a, b, o);
    else if (DA == 9 && DB == 9) matmat_rn(a, b, o);
    else o[0] = mul_rn(a[0], b[0]);
}
template <int DA, int DB, int DO>
struct MulOp {
    static constexpr int NACC = DO;
    const PZ<DA>& A;
    const PZ<DB>& B;
    int na, nb;
    FastDiv fdb;
    double thr;   // squared-domain threshold (Scratch::thr_sq)
    double ca[DA], cb[DB];
    const double* pa; const double* pb;   // coefficient planes and strides, read once
    int cpa, cpb;
    __device__ MulOp(const PZ<DA>& a, const PZ<DB>& b, double t) : A(a), B(b), na(a.n), nb(b.n), fdb(b.divM), thr(t), pa(a.coef), pb(b.coef), cpa(a.cap), cpb(b.cap) {
#pragma unroll
        for (int c = 0; c < DA; c++) ca[c] = a.center[c];
#pragma unroll
        for (int c = 0; c < DB; c++) cb[c] = b.center[c];
    }
    __device__ __forceinline__ void term(unsigned idx, double* o) const {
        double a[DA], b[DB];
        if ((int)idx < na) { ldc<DA>(pa, cpa, idx, a); coef_mul<DA, DB, DO>(a, cb, o); }
        else if ((int)idx < na + nb) { ldc<DB>(pb, cpb, idx - na, b); coef_mul<DA, DB, DO>(ca, b, o); }
        else {
            const int p = idx - na - nb;
            const int i = fdb.div(p), j = p - i * nb;
            ldc<DA>(pa, cpa, i, a); ldc<DB>(pb, cpb, j, b);
            coef_mul<DA, DB, DO>(a, b, o);
        }
    }
    __device__ __forceinline__ void first(unsigned idx, double* acc) const { term(idx, acc); }
    __device__ __forceinline__ void next(unsigned idx, double* acc) const {
        double t[DO];
        term(idx, t);
#pragma unroll
        for (int c = 0; c < DO; c++) acc[c] = add_rn(acc[c], t[c]);
    }
    __device__ __forceinline__ bool finish(const double* acc, double* out, double* drop) const {
        if (normD<DO>(acc) <= thr) {
#pragma unroll
            for (int c = 0; c < DO; c++) drop[c] = fabs(acc[c]);
            return false;
        }
#pragma unroll
        for (int c = 0; c < DO; c++) out[c] = acc[c];
        return true;
    }
};

// fill the sort buffer for a product: rows over the smaller operand, width W = max(na, nb)
template <int NT, bool BIG>
__device__ __forceinline__ int fill_product_keys(Scratch& S, const u64* ka, int na, u64 magic_a, const u64* kb, int nb, u64 magic_b, int& W, u64& magicW) {
    const int N = na + nb + na * nb;
    u64* key = Buf<BIG>::key(S, 0);
    u16* idx = Buf<BIG>::idx(S, 0);
    if (na == 0 || nb == 0) {   // single sorted run
        W = N > 0 ? N : 1;
        magicW = na ? magic_a : magic_b;
        #pragma unroll 1
        for (int g = gtid<NT>(); g < N; g += NT) { key[g] = na ? ka[g] : kb[g]; idx[g] = (u16)g; }
        return N;
    }
    u64 bad = 0;   // degree-overflow guard (key_add_checked)
    if (nb >= na) {   // runs: i = 0..na-1 (width nb), then B's own list (nb), then A's own list (na <= nb, last)
        W = nb; magicW = magic_b;
        const FastDiv fd(magic_b);
        #pragma unroll 1
        for (int g = gtid<NT>(); g < na * nb; g += NT) {
            const int i = fd.div(g), j = g - i * nb;
            key[g] = key_add_checked(ka[i], kb[j], bad);   // degrees add (KPR/PZsparse.cu:938-940); an overflowing field sets ERR_DEGREE
            idx[g] = (u16)(na + nb + g);
        }
        const int o1 = na * nb;
        #pragma unroll 1
        for (int j = gtid<NT>(); j < nb; j += NT) { key[o1 + j] = kb[j]; idx[o1 + j] = (u16)(na + j); }
        const int o0 = o1 + nb;
        #pragma unroll 1
        for (int i = gtid<NT>(); i < na; i += NT) { key[o0 + i] = ka[i]; idx[o0 + i] = (u16)i; }
    }
    else {            // runs: j = 0..nb-1 (width na), then A's own list (na), then B's own list (nb < na, last)
        W = na; magicW = magic_a;
        const FastDiv fd(magic_a);
        #pragma unroll 1
        for (int g = gtid<NT>(); g < na * nb; g += NT) {
            const int j = fd.div(g), i = g - j * na;
            key[g] = key_add_checked(ka[i], kb[j], bad);
            idx[g] = (u16)(na + nb + i * nb + j);
        }
        const int o0 = na * nb;
        #pragma unroll 1
        for (int i = gtid<NT>(); i < na; i += NT) { key[o0 + i] = ka[i]; idx[o0 + i] = (u16)i; }
        const int o1 = o0 + na;
        #pragma unroll 1
        for (int j = gtid<NT>(); j < nb; j += NT) { key[o1 + j] = kb[j]; idx[o1 + j] = (u16)(na + j); }
    }
    if (bad) set_err(S, ERR_DEGREE);
    return N;
}

// radius of component c of a product, rounded up (KPR/PZsparse.cu:944-989):
//   ind_a * ind_b + ((|c_a| + sum|a_i|) * ind_b + ind_a * (|c_b| + sum|b_j|))
// with matrix products for matrix shapes.  Element (r, col) of a 3x3 result is index r + 3*col.
template <int DA, int DB, int DO>
__device__ __forceinline__ double product_radius_c(int c, const double* ca, const double* abssa, const double* inda, const double* cb, const double* abssb, const double* indb) {
    if (DA == 1) {   // scalar times scalar / vector
        const double ma = __dadd_ru(fabs(ca[0]), abssa[0]), mb = __dadd_ru(fabs(cb[c]), abssb[c]);
        return __dadd_ru(__dmul_ru(inda[0], indb[c]), __dadd_ru(__dmul_ru(ma, indb[c]), __dmul_ru(inda[0], mb)));
    }
    const int r = (DB == 9) ? c % 3 : c, col = (DB == 9) ? c / 3 : 0;
    double r1 = 0, r2 = 0, r3 = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int ia = r + 3 * k, ib = k + 3 * col;
        const double ma = __dadd_ru(fabs(ca[ia]), abssa[ia]), mb = __dadd_ru(fabs(cb[ib]), abssb[ib]);
        const double t1 = __dmul_ru(inda[ia], indb[ib]), t2 = __dmul_ru(ma, indb[ib]), t3 = __dmul_ru(inda[ia], mb);
        r1 = k ? __dadd_ru(r1, t1) : t1; r2 = k ? __dadd_ru(r2, t2) : t2; r3 = k ? __dadd_ru(r3, t3) : t3;
    }
    return __dadd_ru(r1, __dadd_ru(r2, r3));
}
template <int DA, int DB, int DO>
__device__ __forceinline__ double product_center_c(int c, const double* ca, const double* cb) {
    if (DA == 1) return mul_rn(ca[0], cb[c]);
    const int r = (DB == 9) ? c % 3 : c, col = (DB == 9) ? c / 3 : 0;
    return add_rn(add_rn(mul_rn(ca[r], cb[3 * col]), mul_rn(ca[r + 3], cb[3 * col + 1])), mul_rn(ca[r + 6], cb[3 * col + 2]));
}
template <int DA, int DB, int DO>
struct MulEpi {
    const PZ<DA>& A;
    const PZ<DB>& B;
    __device__ __forceinline__ void operator()(int c, double& cen, double& r0, double& r1) const {
        cen = product_center_c<DA, DB, DO>(c, A.center, B.center);
        r0 = product_radius_c<DA, DB, DO>(c, A.center, A.abss, A.ind[0], B.center, B.abss, B.ind[0]);
        r1 = product_radius_c<DA, DB, DO>(c, A.center, A.abss, A.ind[1], B.center, B.abss, B.ind[1]);
    }
    __device__ __forceinline__ u64 mask() const { return A.ormask | B.ormask; }
};

template <int NT, int DA, int DB, int DO, bool BIG>
__device__ __forceinline__ void pz_mul_impl(Scratch& S, PZ<DO>& dst, const PZ<DA>& A, const PZ<DB>& B, int N) {
    const int na = A.n, nb = B.n;
    int W = 1;
    u64 magicW = 0;
    TR(0);
    if (N > 0) fill_product_keys<NT, BIG>(S, A.keys, na, A.divM, B.keys, nb, B.divM, W, magicW);
    MulOp<DA, DB, DO> op(A, B, S.thr_sq);
    TR(1);
    gsync<NT>();
    TR(2);
    phase_mark(PH_FILL);
    const int buf = merge_sort_runs<NT, BIG>(S, N, W, magicW);
    reduce_emit<NT, DO, BIG, MulOp<DA, DB, DO>, MulEpi<DA, DB, DO>>(S, buf, N, op, dst, MulEpi<DA, DB, DO>{A, B});
#ifdef ARMOUR_OP_TRACE
    if (blockIdx.x == 64 && threadIdx.x == 0)
        printf("MUL d%d%d na %d nb %d N %d out %d big %d | fill %lld bar %lld sort %lld walk %lld bar %lld scan+emit %lld bar %lld\n", DA, DB, na, nb, N, dst.n, (int)BIG,
               g_tr[1] - g_tr[0], g_tr[2] - g_tr[1], g_tr[3] - g_tr[2], g_tr[4] - g_tr[3], g_tr[5] - g_tr[4], g_tr[6] - g_tr[5], g_tr[7] - g_tr[6]);
#endif
}
// =============================================================================================
// Structured product: the operands share no variable and the small one has at most three monomials, each ONE variable to
// the first power — R_i, R_i^T (k_i, cosqe_i, sinqe_i) and the link boxes (three generator symbols) against anything built
// from other joints: the forward RNEA recursion (R_i^T * state of joints < i) and the forward kinematics (FK_R * R_i,
// FK_R * link_i).  Then every candidate monomial of (c_S + S_1 + .. + S_ns)(c_L + L_1 + .. + L_nl) has its own key — nothing
// merges — and its position in the sorted result follows in closed form, so there is no sort and no segment walk:
//   key(r, j) = L'[j] | bit_r over the extended list L' = [0, keys of L] (0 = the centre), r = 0 for the centre of S.
//   Against a run with a lower (or no) bit, (r, j) sorts after exactly the L' entries whose bits ABOVE bit_r are <= its own;
//   against a run with a higher bit, after those whose bits above that higher bit are < its own.  With gstart_p(j) / gend_p(j)
//   the bounds of j's group of equal "bits above p" in L' (sorted, so groups are contiguous):
//       rank(r, j) = j + r * gend_{p_r}(j) + sum_{s > r} gstart_{p_s}(j)          (runs ordered by bit position, r = 0 first)
//   (checked against a sort on random inputs; tests/test_gpu_parity.py::test_structured_product_equals_generic).
// Group bounds come from ONE packed prefix sum of three head flags (10 bits each) and a start-of-group table.
// Results are identical to the generic path: same keys, same single-term coefficients, same thresholding.
// =============================================================================================
template <int NT, int DA, int DB, int DO, bool S_IS_A>
__device__ __forceinline__ void pz_mul_structured(Scratch& S, PZ<DO>& dst, const PZ<DA>& A, const PZ<DB>& B, int ns, int nl) {
    constexpr int DS = S_IS_A ? DA : DB, DL = S_IS_A ? DB : DA;
    const int M = nl + 1;                       // extended list of the large operand
    const int N = (ns + 1) * M - 1;             // candidates (the centre * centre term is not a monomial)
    u64* lkey = S.skey(0);                      // [M]
    u16* start = (u16*)(lkey + M);              // [3][M + 1] start index of each group; fits: N <= scap and ns >= 1 give M <= scap / 2
    unsigned* gid = (unsigned*)S.sidx(0);       // [M] packed 1-based group numbers (3 x 10 bits); sidx(0) and sidx(1) are contiguous
    u64* okey = S.skey(1);                      // [N] keys by sorted position
    u16* flag = S.sidx(1);                      // [N] keep flags by sorted position (gid needs 4 M <= 2 scap bytes: structured_ok)
    const u64* skeys = S_IS_A ? A.keys : B.keys;
    const u64* lkeys = S_IS_A ? B.keys : A.keys;
    u64 sbit[3] = {0, 0, 0};
    int sh[3] = {63, 63, 63};
#pragma unroll
    for (int r = 0; r < 3; r++) if (r < ns) { sbit[r] = skeys[r]; sh[r] = __ffsll((long long)sbit[r]); }   // bits above p: x >> (p + 1)
    int ncap;
    double* tmp = S.staging(N, DO, ncap);
    double red[2 * DO];
#pragma unroll
    for (int c = 0; c < 2 * DO; c++) red[c] = 0.0;
    ScalarEpilogue<NT, DO> se;
    se.begin(MulEpi<DA, DB, DO>{A, B});
    #pragma unroll 1
    for (int j = gtid<NT>(); j < M; j += NT) lkey[j] = j ? lkeys[j - 1] : 0ull;
    gsync<NT>();
    phase_mark(PH_FILL);
    // group numbers: packed inclusive prefix sum of the head flags, chunk by chunk with a running carry
    {
        unsigned carry = 0;
        const int lane = threadIdx.x & 31, warp = gtid<NT>() >> 5;
        #pragma unroll 1
        for (int base = 0; base < M; base += NT) {
            const int j = base + gtid<NT>();
            unsigned f = 0;
            if (j < M) {
                const u64 k = lkey[j], kp = j ? lkey[j - 1] : 0ull;
#pragma unroll
                for (int r = 0; r < 3; r++) if (r < ns && (j == 0 || (k >> sh[r]) != (kp >> sh[r]))) f |= 1u << (10 * r);
            }
            unsigned incl = f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            if (lane == 31) S.iscan[warp] = (int)incl;
            gsync<NT>();
            unsigned before = carry, all = carry;
#pragma unroll
            for (int w = 0; w < NT / 32; w++) { const unsigned t = (unsigned)S.iscan[w]; all += t; if (w < warp) before += t; }
            const unsigned g = before + incl;
            if (j < M) {
                gid[j] = g;
#pragma unroll
                for (int r = 0; r < 3; r++) if (f & (1u << (10 * r))) start[r * (M + 1) + ((g >> (10 * r)) & 1023u) - 1] = (u16)j;
                if (j == M - 1) {
#pragma unroll
                    for (int r = 0; r < 3; r++) if (r < ns) start[r * (M + 1) + ((g >> (10 * r)) & 1023u)] = (u16)M;   // sentinel: end of the last group
                }
            }
            carry = all;
            gsync<NT>();
        }
    }
    phase_mark(PH_SORT);
    // every candidate: position, key, coefficient, threshold
    const double thr = S.thr_sq;
    const double* ps = S_IS_A ? (const double*)A.coef : (const double*)B.coef;
    const double* pl = S_IS_A ? (const double*)B.coef : (const double*)A.coef;
    const int cps = S_IS_A ? A.cap : B.cap, cpl = S_IS_A ? B.cap : A.cap;
    #pragma unroll 1
    for (int j = gtid<NT>(); j < M; j += NT) {
        const u64 kj = lkey[j];
        const unsigned g = gid[j];
        int gs[3], ge[3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            gs[r] = 0; ge[r] = 0;
            if (r < ns) { const int q = (int)((g >> (10 * r)) & 1023u); gs[r] = start[r * (M + 1) + q - 1]; ge[r] = start[r * (M + 1) + q]; }
        }
        double lv[DL];
        if (j == 0) {
#pragma unroll
            for (int c = 0; c < DL; c++) lv[c] = S_IS_A ? B.center[c < DB ? c : 0] : A.center[c < DA ? c : 0];
        }
        else ldc<DL>(pl, cpl, j - 1, lv);
#pragma unroll
        for (int r = 0; r <= 3; r++) {   // runs are ordered by bit position: run r >= 1 carries sbit[r - 1]; fully unrolled (static register indices)
            if (r > ns || (r == 0 && j == 0)) continue;
            int rank = j;
            if (r >= 1) rank += r * ge[r >= 1 ? r - 1 : 0];
#pragma unroll
            for (int s2 = 0; s2 < 3; s2++) if (s2 + 1 > r && s2 < ns) rank += gs[s2];
            const int pos = rank - 1;
            double sv[DS];
            if (r == 0) {
#pragma unroll
                for (int c = 0; c < DS; c++) sv[c] = S_IS_A ? A.center[c < DA ? c : 0] : B.center[c < DB ? c : 0];
            }
            else ldc<DS>(ps, cps, r - 1, sv);
            double o[DO];
            if (S_IS_A) coef_mul<DA, DB, DO>(sv, lv, o); else coef_mul<DA, DB, DO>(lv, sv, o);
            okey[pos] = r ? (kj | sbit[r >= 1 ? r - 1 : 0]) : kj;
            u16 f = 0;
            if (normD<DO>(o) <= thr) {
#pragma unroll
                for (int c = 0; c < DO; c++) red[c] = __dadd_ru(red[c], fabs(o[c]));
            }
            else {
                f = 1;
#pragma unroll
                for (int c = 0; c < DO; c++) { tmp[c * ncap + pos] = o[c]; red[DO + c] = __dadd_ru(red[DO + c], fabs(o[c])); }
            }
            flag[pos] = f;
        }
    }
    gsync<NT>();
    phase_mark(PH_SEGMENT);
    compact_emit<NT, DO>(S, N, okey, flag, tmp, ncap, red, se, dst);
}
// can `Sm` play the small operand against `L`?  (group-uniform: every thread evaluates the same descriptors and keys)
template <int DSm, int DLg>
__device__ __forceinline__ bool structured_ok(const Scratch& S, const PZ<DSm>& Sm, const PZ<DLg>& L) {
    const int ns = Sm.n, nl = L.n;
    if (ns < 1 || ns > 3 || nl < 1) return false;
    if ((ns + 1) * (nl + 1) - 1 > S.scap || 2 * (nl + 1) > S.scap || nl + 1 > 1023) return false;
    u64 fields = 0;
    for (int r = 0; r < ns; r++) {
        const u64 k = Sm.keys[r];
        if (__popcll(k) != 1) return false;
        const int p = __ffsll((long long)k) - 1;
        const bool one_bit = p >= 14 && p < 35;
        if (!one_bit && ((p < 14 ? p : p - 35) & 1)) return false;   // degree 2: the high bit of a 2-bit field
        fields |= one_bit ? k : (k | (k << 1));
    }
    return (L.ormask & fields) == 0;
}

#ifndef ARMOUR_STRUCTURED_PRODUCTS
#define ARMOUR_STRUCTURED_PRODUCTS 1
#endif
// R * v: the sort-free structured path when it applies (one-plan shapes), else the unified 3-vector operation; one copy per kernel
template <int NT>
__device__ __noinline__ void pz_mul93(Scratch& S, PZ<3>& dst, const PZ<9>& A, const PZ<3>& B) {
    if (ARMOUR_STRUCTURED_PRODUCTS && NT >= 256 && !S.no_structured) {
        if (structured_ok<9, 3>(S, A, B)) { pz_mul_structured<NT, 9, 3, 3, true>(S, dst, A, B, A.n, B.n); return; }
        if (structured_ok<3, 9>(S, B, A)) { pz_mul_structured<NT, 9, 3, 3, false>(S, dst, A, B, B.n, A.n); return; }
    }
    op3_mul93<NT>(S, dst, A, B);
}
template <int NT, int DA, int DB, int DO, int CP>
__device__ __noinline__ void pz_mul_cp(Scratch& S, PZ<DO>& dst, const PZ<DA>& A, const PZ<DB>& B);
template <int NT, int DA, int DB, int DO>
__device__ __forceinline__ void pz_mul(Scratch& S, PZ<DO>& dst, const PZ<DA>& A, const PZ<DB>& B) {
#if ARMOUR_UNIFIED_OP3
    if constexpr (DA == 9 && DB == 3 && NT <= 128) { pz_mul93<NT>(S, dst, A, B); return; }
#endif
#ifdef ARMOUR_DUP_CODE
    if (((size_t)&dst >> 7) & 1) { pz_mul_cp<NT, DA, DB, DO, 1>(S, dst, A, B); return; }
#endif
    pz_mul_cp<NT, DA, DB, DO, 0>(S, dst, A, B);
}
template <int NT, int DA, int DB, int DO, int CP>
__device__ __noinline__ void pz_mul_cp(Scratch& S, PZ<DO>& dst, const PZ<DA>& A, const PZ<DB>& B) {
    // one-plan shapes only (256 threads per group): measured -3.3 % on one plan; in the 128-thread sweep shape the sort-free path
    // executes 7 % fewer instructions but runs 7 % slower (its per-candidate chain is longer and the extra code costs 1.5 % even
    // when unused), so it is compiled out there
    if (ARMOUR_STRUCTURED_PRODUCTS && NT >= 256 && !S.no_structured) {
        if (structured_ok<DA, DB>(S, A, B)) { pz_mul_structured<NT, DA, DB, DO, true>(S, dst, A, B, A.n, B.n); return; }
        if (structured_ok<DB, DA>(S, B, A)) { pz_mul_structured<NT, DA, DB, DO, false>(S, dst, A, B, B.n, A.n); return; }
    }
    int N = A.n + B.n + A.n * B.n;
    if (N > S.ncap || N > 65535) { if (gtid<NT>() == 0) set_err(S, ERR_ENTRY_CAP); N = 0; }
    if (N <= S.scap) pz_mul_impl<NT, DA, DB, DO, false>(S, dst, A, B, N);
    else pz_mul_impl<NT, DA, DB, DO, true>(S, dst, A, B, N);
}

// =============================================================================================
// Two-run merges:  dst = viewA + viewB  (operator+, operator-, addOneDimPZ, element extraction
// followed by +; KPR/PZsparse.cu:678-697, 743-834, 1068-1085)
// =============================================================================================
enum ViewMode { VIEW_SAME = 0, VIEW_EXTRACT = 1, VIEW_PLACE = 2 };
template <int D>
struct View {
    const PZ<D>* p;
    int mode;       // VIEW_SAME: D == DOUT; VIEW_EXTRACT: D == 3 -> DOUT == 1 (row); VIEW_PLACE: D == 1 -> DOUT == 3 (row)
    int row;
    double scale;   // coefficients, centre scaled by `scale`, radius by |scale| (operator*(double), :996-1030); 1.0 = none
    bool scaled;
};
template <int D> __device__ __forceinline__ View<D> view(const PZ<D>& p) { return View<D>{&p, VIEW_SAME, 0, 1.0, false}; }
template <int D> __device__ __forceinline__ View<D> view_scaled(const PZ<D>& p, double s) { return View<D>{&p, VIEW_SAME, 0, s, true}; }
__device__ __forceinline__ View<3> view_extract(const PZ<3>& p, int row) { return View<3>{&p, VIEW_EXTRACT, row, 1.0, false}; }
__device__ __forceinline__ View<1> view_place(const PZ<1>& p, int row) { return View<1>{&p, VIEW_PLACE, row, 1.0, false}; }

template <int D, int DO>
__device__ __forceinline__ void view_vec(const View<D>& v, const double* src, double* o) {   // map a D-vector of the source to DO
    if (v.mode == VIEW_SAME) {
#pragma unroll
        for (int c = 0; c < DO; c++) o[c] = v.scaled ? mul_rn(v.scale, src[c < D ? c : 0]) : src[c < D ? c : 0];
    }
    else if (v.mode == VIEW_EXTRACT) { o[0] = v.scaled ? mul_rn(v.scale, src[v.row < D ? v.row : 0]) : src[v.row < D ? v.row : 0]; }
    else {
#pragma unroll
        for (int c = 0; c < DO; c++) o[c] = 0.0;
        o[v.row < DO ? v.row : 0] = v.scaled ? mul_rn(v.scale, src[0]) : src[0];
    }
}
// component c of a view's centre (round-to-nearest scaling) / radius (|scale|, rounded up)
template <int D>
__device__ __forceinline__ double view_comp(const View<D>& v, const double* src, int c, bool radius) {
    int i;
    if (v.mode == VIEW_SAME) i = c;
    else if (v.mode == VIEW_EXTRACT) i = v.row;
    else { if (c != v.row) return 0.0; i = 0; }
    i = i < D ? i : 0;
    if (!v.scaled) return src[i];
    return radius ? __dmul_ru(fabs(v.scale), src[i]) : mul_rn(v.scale, src[i]);
}
template <int DA, int DB, int DO>
struct MergeOp {
    static constexpr int NACC = DO;
    View<DA> A;
    View<DB> B;
    int na;
    bool negb;
    double thr;   // squared-domain threshold (Scratch::thr_sq)
    const double* pa; const double* pb;   // coefficient planes and strides, read once
    int cpa, cpb;
    __device__ __forceinline__ void term(unsigned idx, double* o) const {
        if ((int)idx < na) { double a[DA]; ldc<DA>(pa, cpa, idx, a); view_vec<DA, DO>(A, a, o); }
        else {
            double b[DB]; ldc<DB>(pb, cpb, idx - na, b); view_vec<DB, DO>(B, b, o);
            if (negb) {
#pragma unroll
                for (int c = 0; c < DO; c++) o[c] = -o[c];
            }
        }
    }
    __device__ __forceinline__ void first(unsigned idx, double* acc) const { term(idx, acc); }
    __device__ __forceinline__ void next(unsigned idx, double* acc) const {
        double t[DO];
        term(idx, t);
#pragma unroll
        for (int c = 0; c < DO; c++) acc[c] = add_rn(acc[c], t[c]);
    }
    __device__ __forceinline__ bool finish(const double* acc, double* out, double* drop) const {
        if (normD<DO>(acc) <= thr) {
#pragma unroll
            for (int c = 0; c < DO; c++) drop[c] = fabs(acc[c]);
            return false;
        }
#pragma unroll
        for (int c = 0; c < DO; c++) out[c] = acc[c];
        return true;
    }
};
template <int DA, int DB>
struct MergeEpi {
    View<DA> A;
    View<DB> B;
    bool negb;
    __device__ __forceinline__ void operator()(int c, double& cen, double& r0, double& r1) const {
        const double ca = view_comp<DA>(A, A.p->center, c, false), cb = view_comp<DB>(B, B.p->center, c, false);
        cen = negb ? add_rn(ca, -cb) : add_rn(ca, cb);
        r0 = __dadd_ru(view_comp<DA>(A, A.p->ind[0], c, true), view_comp<DB>(B, B.p->ind[0], c, true));
        r1 = __dadd_ru(view_comp<DA>(A, A.p->ind[1], c, true), view_comp<DB>(B, B.p->ind[1], c, true));
    }
    __device__ __forceinline__ u64 mask() const { return A.p->ormask | B.p->ormask; }
};
// dst = A (+/-) B through views.  Centre and radii of a VIEW_PLACE / VIEW_EXTRACT source are mapped the same way.
template <int NT, int DA, int DB, int DO, bool BIG>
__device__ __forceinline__ void pz_merge_impl(Scratch& S, PZ<DO>& dst, const View<DA>& A, const View<DB>& B, bool negb, int N) {
    const int na = A.p->n, nb = B.p->n;
    u64* key = Buf<BIG>::key(S, 0);
    u16* idx = Buf<BIG>::idx(S, 0);
    int W = 1;
    u64 magicW = 0;
    if (N > 0) {
        const u64* ka = A.p->keys; const u64* kb = B.p->keys;
        if (na >= nb) {   // the longer run first: runs must have uniform width except the last
            W = na > 0 ? na : 1; magicW = A.p->divM;
            #pragma unroll 1
            for (int i = gtid<NT>(); i < na; i += NT) { key[i] = ka[i]; idx[i] = (u16)i; }
            #pragma unroll 1
            for (int j = gtid<NT>(); j < nb; j += NT) { key[na + j] = kb[j]; idx[na + j] = (u16)(na + j); }
        }
        else {
            W = nb; magicW = B.p->divM;
            #pragma unroll 1
            for (int j = gtid<NT>(); j < nb; j += NT) { key[j] = kb[j]; idx[j] = (u16)(na + j); }
            #pragma unroll 1
            for (int i = gtid<NT>(); i < na; i += NT) { key[nb + i] = ka[i]; idx[nb + i] = (u16)i; }
        }
    }
    MergeOp<DA, DB, DO> op{A, B, na, negb, S.thr_sq, A.p->coef, B.p->coef, A.p->cap, B.p->cap};
    gsync<NT>();
    phase_mark(PH_FILL);
    const int buf = merge_sort_runs<NT, BIG>(S, N, W, magicW);
    reduce_emit<NT, DO, BIG, MergeOp<DA, DB, DO>, MergeEpi<DA, DB>>(S, buf, N, op, dst, MergeEpi<DA, DB>{A, B, negb});
}
template <int NT, int DA, int DB, int DO, int CP>
__device__ __noinline__ void pz_merge_cp(Scratch& S, PZ<DO>& dst, const View<DA> A, const View<DB> B, bool negb);
template <int NT, int DA, int DB, int DO>
__device__ __forceinline__ void pz_merge(Scratch& S, PZ<DO>& dst, const View<DA> A, const View<DB> B, bool negb) {
#ifdef ARMOUR_DUP_CODE
    if (((size_t)&dst >> 7) & 1) { pz_merge_cp<NT, DA, DB, DO, 1>(S, dst, A, B, negb); return; }
#endif
    pz_merge_cp<NT, DA, DB, DO, 0>(S, dst, A, B, negb);
}
template <int NT, int DA, int DB, int DO, int CP>
__device__ __noinline__ void pz_merge_cp(Scratch& S, PZ<DO>& dst, const View<DA> A, const View<DB> B, bool negb) {
    int N = A.p->n + B.p->n;
    if (N > S.ncap || N > 65535) { if (gtid<NT>() == 0) set_err(S, ERR_ENTRY_CAP); N = 0; }
    if (N <= S.scap) pz_merge_impl<NT, DA, DB, DO, false>(S, dst, A, B, negb, N);
    else pz_merge_impl<NT, DA, DB, DO, true>(S, dst, A, B, negb, N);
}
// simplify() of an arbitrary monomial list (unsorted, repeated keys allowed): KPR/PZsparse.cu:284-350.  The engine's own
// operations never need it (their operands are sorted and unique by construction); it backs PZsparse::simplify / stack of
// the host facade.  A full merge sort from runs of width 1, ties on the list position (= stable), then the shared walk.
template <int D>
struct SimplifyEpi {
    const PZ<D>& A;
    __device__ __forceinline__ void operator()(int c, double& cen, double& r0, double& r1) const { cen = A.center[c]; r0 = A.ind[0][c]; r1 = A.ind[1][c]; }
    __device__ __forceinline__ u64 mask() const { return A.ormask; }
};
template <int NT, int D, bool BIG>
__device__ __forceinline__ void pz_simplify_impl(Scratch& S, PZ<D>& dst, const PZ<D>& A, int N) {
    u64* key = Buf<BIG>::key(S, 0);
    u16* idx = Buf<BIG>::idx(S, 0);
    const u64* ka = A.keys;
    #pragma unroll 1
    for (int i = gtid<NT>(); i < N; i += NT) { key[i] = ka[i]; idx[i] = (u16)i; }
    const View<D> va = view(A);
    MergeOp<D, D, D> op{va, va, N, false, S.thr_sq, A.coef, A.coef, A.cap, A.cap};   // every origin index is below na = N: the second view is never read
    gsync<NT>();
    const int buf = merge_sort_runs<NT, BIG>(S, N, 1, FastDiv::magic(1));
    reduce_emit<NT, D, BIG, MergeOp<D, D, D>, SimplifyEpi<D>>(S, buf, N, op, dst, SimplifyEpi<D>{A});
}
template <int NT, int D>
__device__ __noinline__ void pz_simplify(Scratch& S, PZ<D>& dst, const PZ<D>& A) {
    int N = A.n;
    if (N > S.ncap || N > 65535) { if (gtid<NT>() == 0) set_err(S, ERR_ENTRY_CAP); N = 0; }
    if (N <= S.scap) pz_simplify_impl<NT, D, false>(S, dst, A, N);
    else pz_simplify_impl<NT, D, true>(S, dst, A, N);
}
template <int NT> __device__ __forceinline__ void pz_add3(Scratch& S, PZ<3>& dst, const PZ<3>& a, const PZ<3>& b) {
#if ARMOUR_UNIFIED_OP3
    if constexpr (NT <= 128) { op3_add<NT>(S, dst, a, b); return; }
#endif
    pz_merge<NT, 3, 3, 3>(S, dst, view(a), view(b), false);
}
// dst = a with the scalar PZ s added into row `row`   (addOneDimPZ)
template <int NT> __device__ __forceinline__ void pz_add_one_dim(Scratch& S, PZ<3>& dst, const PZ<3>& a, const PZ<1>& s, int row) {
#if ARMOUR_UNIFIED_OP3
    if constexpr (NT <= 128) { op3_add_one_dim<NT>(S, dst, a, s, row); return; }
#endif
    pz_merge<NT, 3, 1, 3>(S, dst, view(a), view_place(s, row), false);
}

// =============================================================================================
// cross(PZ a, PZ b) for 3x1 operands, fused  (KPR/PZsparse.cu:1134-1151: six scalar products,
// three differences, one stack — each followed by simplify in the reference)
// =============================================================================================
struct CrossPPOp {
    // accumulators: s[0]=a1*b2, s[1]=a2*b1, s[2]=a2*b0, s[3]=a0*b2, s[4]=a0*b1, s[5]=a1*b0
    static constexpr int NACC = 6;
    const PZ<3>& A;
    const PZ<3>& B;
    int na, nb;
    FastDiv fdb;
    double thr;   // squared-domain threshold (Scratch::thr_sq)
    double ca[3], cb[3];
    const double* pa; const double* pb;
    int cpa, cpb;
    __device__ CrossPPOp(const PZ<3>& a, const PZ<3>& b, double t) : A(a), B(b), na(a.n), nb(b.n), fdb(b.divM), thr(t), pa(a.coef), pb(b.coef), cpa(a.cap), cpb(b.cap) {
        for (int c = 0; c < 3; c++) { ca[c] = a.center[c]; cb[c] = b.center[c]; }
    }
    __device__ __forceinline__ void term(unsigned idx, double* o) const {
        double a[3], b[3];
        if ((int)idx < na) { ldc<3>(pa, cpa, idx, a); b[0] = cb[0]; b[1] = cb[1]; b[2] = cb[2]; }
        else if ((int)idx < na + nb) { ldc<3>(pb, cpb, idx - na, b); a[0] = ca[0]; a[1] = ca[1]; a[2] = ca[2]; }
        else {
            const int p = idx - na - nb;
            const int i = fdb.div(p), j = p - i * nb;
            ldc<3>(pa, cpa, i, a); ldc<3>(pb, cpb, j, b);
        }
        o[0] = mul_rn(a[1], b[2]); o[1] = mul_rn(a[2], b[1]);
        o[2] = mul_rn(a[2], b[0]); o[3] = mul_rn(a[0], b[2]);
        o[4] = mul_rn(a[0], b[1]); o[5] = mul_rn(a[1], b[0]);
    }
    __device__ __forceinline__ void first(unsigned idx, double* acc) const { term(idx, acc); }
    __device__ __forceinline__ void next(unsigned idx, double* acc) const {
        double t[6];
        term(idx, t);
#pragma unroll
        for (int c = 0; c < 6; c++) acc[c] = add_rn(acc[c], t[c]);
    }
    // stage 1: each scalar product's simplify; stage 2: the difference's simplify; stage 3: stack's simplify
    __device__ __forceinline__ bool finish(const double* acc, double* out, double* drop) const {
        bool any = false;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double p = acc[2 * c], q = acc[2 * c + 1];
            const bool kp = norm1(&p) > thr, kq = norm1(&q) > thr;
            double d = 0.0;
            if (!kp) d = __dadd_ru(d, fabs(p));
            if (!kq) d = __dadd_ru(d, fabs(q));
            double v = 0.0;
            bool present = false;
            if (kp || kq) {
                v = kp ? (kq ? add_rn(p, -q) : p) : -q;
                if (norm1(&v) > thr) present = true;
                else { d = __dadd_ru(d, fabs(v)); v = 0.0; }
            }
            out[c] = v;
            drop[c] = d;
            any |= present;
        }
        if (!any) return false;
        if (norm3(out) <= thr) {
#pragma unroll
            for (int c = 0; c < 3; c++) drop[c] = __dadd_ru(drop[c], fabs(out[c]));
            return false;
        }
        return true;
    }
};
struct CrossPPEpi {
    const PZ<3>& A;
    const PZ<3>& B;
    // component c of the result is P - Q with P = a[c+1]*b[c+2], Q = a[c+2]*b[c+1]
    __device__ __forceinline__ void operator()(int c, double& cen, double& r0, double& r1) const {
        const int i1 = (c + 1) % 3, i2 = (c + 2) % 3;
        cen = add_rn(mul_rn(A.center[i1], B.center[i2]), -mul_rn(A.center[i2], B.center[i1]));
        double r[2];
#pragma unroll
        for (int v = 0; v < 2; v++) {
            const double rp = product_radius_c<1, 1, 1>(0, &A.center[i1], &A.abss[i1], &A.ind[v][i1], &B.center[i2], &B.abss[i2], &B.ind[v][i2]);
            const double rq = product_radius_c<1, 1, 1>(0, &A.center[i2], &A.abss[i2], &A.ind[v][i2], &B.center[i1], &B.abss[i1], &B.ind[v][i1]);
            r[v] = __dadd_ru(rp, rq);
        }
        r0 = r[0]; r1 = r[1];
    }
    __device__ __forceinline__ u64 mask() const { return A.ormask | B.ormask; }
};
template <int NT, bool BIG>
__device__ __forceinline__ void pz_cross_pp_impl(Scratch& S, PZ<3>& dst, const PZ<3>& A, const PZ<3>& B, int N) {
    const int na = A.n, nb = B.n;
    int W = 1;
    u64 magicW = 0;
    if (N > 0) fill_product_keys<NT, BIG>(S, A.keys, na, A.divM, B.keys, nb, B.divM, W, magicW);
    CrossPPOp op(A, B, S.thr_sq);
    gsync<NT>();
    phase_mark(PH_FILL);
    const int buf = merge_sort_runs<NT, BIG>(S, N, W, magicW);
    reduce_emit<NT, 3, BIG, CrossPPOp, CrossPPEpi>(S, buf, N, op, dst, CrossPPEpi{A, B});
}
template <int NT, int CP>
__device__ __noinline__ void pz_cross_pp_cp(Scratch& S, PZ<3>& dst, const PZ<3>& A, const PZ<3>& B);
template <int NT>
__device__ __forceinline__ void pz_cross_pp(Scratch& S, PZ<3>& dst, const PZ<3>& A, const PZ<3>& B) {
#if ARMOUR_UNIFIED_OP3
    if constexpr (NT <= 128) { op3_cross<NT>(S, dst, A, B); return; }
#endif
#ifdef ARMOUR_DUP_CODE
    if (((size_t)&dst >> 7) & 1) { pz_cross_pp_cp<NT, 1>(S, dst, A, B); return; }
#endif
    pz_cross_pp_cp<NT, 0>(S, dst, A, B);
}
template <int NT, int CP>
__device__ __noinline__ void pz_cross_pp_cp(Scratch& S, PZ<3>& dst, const PZ<3>& A, const PZ<3>& B) {
    int N = A.n + B.n + A.n * B.n;
    if (N > S.ncap || N > 65535) { if (gtid<NT>() == 0) set_err(S, ERR_ENTRY_CAP); N = 0; }
    if (N <= S.scap) pz_cross_pp_impl<NT, false>(S, dst, A, B, N);
    else pz_cross_pp_impl<NT, true>(S, dst, A, B, N);
}

// =============================================================================================
// Elementwise operations (operand lists with identical keys: no sort needed)
// =============================================================================================
// cross(PZ a, const b) and cross(const a, PZ b)   (KPR/PZsparse.cu:1118-1132, 1153-1167)
__device__ __forceinline__ void cross_const_comp(const double* k, bool const_first, const double* z, double* r) {
    if (const_first) {   // r_c = k[c+1]*z[c+2] - k[c+2]*z[c+1]
        r[0] = add_rn(mul_rn(k[1], z[2]), -mul_rn(k[2], z[1]));
        r[1] = add_rn(mul_rn(k[2], z[0]), -mul_rn(k[0], z[2]));
        r[2] = add_rn(mul_rn(k[0], z[1]), -mul_rn(k[1], z[0]));
    }
    else {               // r_c = z[c+1]*k[c+2] - z[c+2]*k[c+1]
        r[0] = add_rn(mul_rn(k[2], z[1]), -mul_rn(k[1], z[2]));
        r[1] = add_rn(mul_rn(k[0], z[2]), -mul_rn(k[2], z[0]));
        r[2] = add_rn(mul_rn(k[1], z[0]), -mul_rn(k[0], z[1]));
    }
}
struct CrossConstOp {
    const PZ<3>& Z;
    double k[3];
    bool const_first;   // true: cross(k, Z); false: cross(Z, k)
    double thr;   // squared-domain threshold (Scratch::thr_sq)
    __device__ __forceinline__ bool finish(int i, double* out, double* drop) const {
        double z[3], r[3];
        ldc<3>(Z, i, z);
        cross_const_comp(k, const_first, z, r);
        bool any = false;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (norm1(&r[c]) > thr) { out[c] = r[c]; any = true; drop[c] = 0.0; }
            else { out[c] = 0.0; drop[c] = fabs(r[c]); }
        }
        if (!any) return false;
        if (norm3(out) <= thr) {
#pragma unroll
            for (int c = 0; c < 3; c++) drop[c] = __dadd_ru(drop[c], fabs(out[c]));
            return false;
        }
        return true;
    }
};
struct CrossConstEpi {
    const PZ<3>& Z;
    double k[3];
    bool const_first;
    __device__ __forceinline__ void operator()(int c, double& cen, double& r0, double& r1) const {
        double r[3];
        cross_const_comp(k, const_first, Z.center, r);
        cen = r[c];
        const int i1 = (c + 1) % 3, i2 = (c + 2) % 3;
        // operator*(double) scales the radius by |k|, operator- adds the two radii
        r0 = __dadd_ru(__dmul_ru(Z.ind[0][i1], fabs(k[i2])), __dmul_ru(Z.ind[0][i2], fabs(k[i1])));
        r1 = __dadd_ru(__dmul_ru(Z.ind[1][i1], fabs(k[i2])), __dmul_ru(Z.ind[1][i2], fabs(k[i1])));
    }
    __device__ __forceinline__ u64 mask() const { return Z.ormask; }
};
template <int NT, int CP>
__device__ __noinline__ void pz_cross_const_cp(Scratch& S, PZ<3>& dst, const PZ<3>& Z, const double* kvec, bool const_first);
template <int NT>
__device__ __forceinline__ void pz_cross_const(Scratch& S, PZ<3>& dst, const PZ<3>& Z, const double* kvec, bool const_first) {
#if ARMOUR_UNIFIED_OP3
    if constexpr (NT <= 128) { op3_cross_const<NT>(S, dst, Z, kvec, const_first); return; }
#endif
#ifdef ARMOUR_DUP_CODE
    if (((size_t)&dst >> 7) & 1) { pz_cross_const_cp<NT, 1>(S, dst, Z, kvec, const_first); return; }
#endif
    pz_cross_const_cp<NT, 0>(S, dst, Z, kvec, const_first);
}
template <int NT, int CP>
__device__ __noinline__ void pz_cross_const_cp(Scratch& S, PZ<3>& dst, const PZ<3>& Z, const double* kvec, bool const_first) {
    CrossConstOp op{Z, {kvec[0], kvec[1], kvec[2]}, const_first, S.thr_sq};
    if (Z.n <= S.scap) elementwise_emit<NT, 3, false, CrossConstOp, CrossConstEpi>(S, Z.n, Z.keys, op, dst, CrossConstEpi{Z, {kvec[0], kvec[1], kvec[2]}, const_first});
    else elementwise_emit<NT, 3, true, CrossConstOp, CrossConstEpi>(S, Z.n, Z.keys, op, dst, CrossConstEpi{Z, {kvec[0], kvec[1], kvec[2]}, const_first});
}

// dst(3x1) = M * v with M a monomial-free 3x3 PZ (centre Mc, radii Mi[2]) — I_arr(i) * w
// dst(3x1) = m * v with m a monomial-free scalar PZ                         — mass_arr(i) * (...)
struct ConstLeftOp {
    const PZ<3>& V;
    double M[9];
    bool scalar;
    double thr;   // squared-domain threshold (Scratch::thr_sq)
    __device__ __forceinline__ bool finish(int i, double* out, double* drop) const {
        double z[3], r[3];
        ldc<3>(V, i, z);
        if (scalar) { r[0] = mul_rn(M[0], z[0]); r[1] = mul_rn(M[0], z[1]); r[2] = mul_rn(M[0], z[2]); }
        else matvec_rn(M, z, r);
        if (norm3(r) <= thr) { for (int c = 0; c < 3; c++) drop[c] = fabs(r[c]); return false; }
        for (int c = 0; c < 3; c++) out[c] = r[c];
        return true;
    }
};
struct ConstLeftEpi {
    const PZ<3>& V;
    const double* Mc;
    const double* Mi0;
    const double* Mi1;
    bool scalar;
    __device__ __forceinline__ void operator()(int c, double& cen, double& r0, double& r1) const {
        const double zero9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (scalar) {
            cen = mul_rn(Mc[0], V.center[c]);
            r0 = product_radius_c<1, 3, 3>(c, Mc, zero9, Mi0, V.center, V.abss, V.ind[0]);
            r1 = product_radius_c<1, 3, 3>(c, Mc, zero9, Mi1, V.center, V.abss, V.ind[1]);
        }
        else {
            cen = product_center_c<9, 3, 3>(c, Mc, V.center);
            r0 = product_radius_c<9, 3, 3>(c, Mc, zero9, Mi0, V.center, V.abss, V.ind[0]);
            r1 = product_radius_c<9, 3, 3>(c, Mc, zero9, Mi1, V.center, V.abss, V.ind[1]);
        }
    }
    __device__ __forceinline__ u64 mask() const { return V.ormask; }
};
template <int NT, int CP>
__device__ __noinline__ void pz_const_left_cp(Scratch& S, PZ<3>& dst, const double* Mc, const double* Mi0, const double* Mi1, bool scalar, const PZ<3>& V);
template <int NT>
__device__ __forceinline__ void pz_const_left(Scratch& S, PZ<3>& dst, const double* Mc, const double* Mi0, const double* Mi1, bool scalar, const PZ<3>& V) {
#if ARMOUR_UNIFIED_OP3
    if constexpr (NT <= 128) { op3_const_left<NT>(S, dst, Mc, Mi0, Mi1, scalar, V); return; }
#endif
#ifdef ARMOUR_DUP_CODE
    if (((size_t)&dst >> 7) & 1) { pz_const_left_cp<NT, 1>(S, dst, Mc, Mi0, Mi1, scalar, V); return; }
#endif
    pz_const_left_cp<NT, 0>(S, dst, Mc, Mi0, Mi1, scalar, V);
}
template <int NT, int CP>
__device__ __noinline__ void pz_const_left_cp(Scratch& S, PZ<3>& dst, const double* Mc, const double* Mi0, const double* Mi1, bool scalar, const PZ<3>& V) {
    ConstLeftOp op{V, {0}, scalar, S.thr_sq};
    const int DM = scalar ? 1 : 9;
    for (int c = 0; c < DM; c++) op.M[c] = Mc[c];
    if (V.n <= S.scap) elementwise_emit<NT, 3, false, ConstLeftOp, ConstLeftEpi>(S, V.n, V.keys, op, dst, ConstLeftEpi{V, Mc, Mi0, Mi1, scalar});
    else elementwise_emit<NT, 3, true, ConstLeftOp, ConstLeftEpi>(S, V.n, V.keys, op, dst, ConstLeftEpi{V, Mc, Mi0, Mi1, scalar});
}

// dst(3x1) = R * p with R a 3x3 PZ and p a constant vector — FK_R * P   (KPR/Dynamics.cu:76)
struct ConstRightOp {
    const PZ<9>& R;
    double p[3];
    double thr;   // squared-domain threshold (Scratch::thr_sq)
    __device__ __forceinline__ bool finish(int i, double* out, double* drop) const {
        double m[9], r[3];
        ldc<9>(R, i, m);
        matvec_rn(m, p, r);
        if (norm3(r) <= thr) { for (int c = 0; c < 3; c++) drop[c] = fabs(r[c]); return false; }
        for (int c = 0; c < 3; c++) out[c] = r[c];
        return true;
    }
};
struct ConstRightEpi {
    const PZ<9>& R;
    double p[3];
    __device__ __forceinline__ void operator()(int c, double& cen, double& r0, double& r1) const {
        const double zero3[3] = {0, 0, 0};
        cen = product_center_c<9, 3, 3>(c, R.center, p);
        r0 = product_radius_c<9, 3, 3>(c, R.center, R.abss, R.ind[0], p, zero3, zero3);
        r1 = product_radius_c<9, 3, 3>(c, R.center, R.abss, R.ind[1], p, zero3, zero3);
    }
    __device__ __forceinline__ u64 mask() const { return R.ormask; }
};
template <int NT>
__device__ __forceinline__ void pz_const_right(Scratch& S, PZ<3>& dst, const PZ<9>& R, const double* pvec) {
#if ARMOUR_UNIFIED_OP3
    if constexpr (NT <= 128) { op3_const_right<NT>(S, dst, R, pvec); return; }
#endif
    ConstRightOp op{R, {pvec[0], pvec[1], pvec[2]}, S.thr_sq};
    if (R.n <= S.scap) elementwise_emit<NT, 3, false, ConstRightOp, ConstRightEpi>(S, R.n, R.keys, op, dst, ConstRightEpi{R, {pvec[0], pvec[1], pvec[2]}});
    else elementwise_emit<NT, 3, true, ConstRightOp, ConstRightEpi>(S, R.n, R.keys, op, dst, ConstRightEpi{R, {pvec[0], pvec[1], pvec[2]}});
}

// =============================================================================================
// Unified 3-vector operation.  The sweep shape is bound by instruction fetch, not by issue slots or memory: with the hot code
// of the kernel duplicated (same executed instructions, twice the footprint) it runs 32 % slower (profiles/README.md, round 2).
// Every operation above inlines its own copy of the key fill, the merge sort, the segment walk and the compaction; here all
// seven operations that produce a 3-vector — R * v, v + v, v + scalar-in-a-row, cross(v, v), cross with a constant, constant
// matrix / scalar times v, R * constant — share ONE copy, and only the few instructions that differ (the term of a candidate,
// the threshold stages, the scalar epilogue) are selected by `kind`.  Arithmetic, candidate order and results are those of the
// dedicated operations (the element-wise ones become a single sorted run whose every key is its own segment).
// =============================================================================================
enum Op3Kind { K3_MUL93 = 0, K3_MERGE33 = 1, K3_MERGE31 = 2, K3_CROSS = 3, K3_CROSS_CONST = 4, K3_CONST_LEFT = 5, K3_CONST_RIGHT = 6 };
struct Op3 {
    int kind;
    const void* A; const void* B;       // operand descriptors (PZ<9> / PZ<3> / PZ<1> by kind), for the scalar epilogue
    const u64* ka; const u64* kb; int na, nb; u64 ma, mb;     // keys, counts, division magics
    const double* pa; const double* pb; int cpa, cpb;          // coefficient planes and strides
    double m[9];   // MUL93: centre of A; CROSS: centre of A; CROSS_CONST: the constant; CONST_LEFT: M (or the scalar in m[0]); CONST_RIGHT: p
    double v[3];   // MUL93, CROSS: centre of B
    int row;       // MERGE31: row that receives the scalar
    bool flag;     // CROSS_CONST: constant first; CONST_LEFT: scalar
    const double* Mc; const double* Mi0; const double* Mi1;   // CONST_LEFT: centre and the two radii of the constant operand
};
__device__ __forceinline__ void op3_term(const Op3& d, unsigned idx, double* o) {   // o[6]; entries 3..5 only for K3_CROSS
    switch (d.kind) {
    case K3_MUL93: {
        double a[9], b[3];
        if ((int)idx < d.na) { ldc<9>(d.pa, d.cpa, idx, a); matvec_rn(a, d.v, o); }
        else if ((int)idx < d.na + d.nb) { ldc<3>(d.pb, d.cpb, idx - d.na, b); matvec_rn(d.m, b, o); }
        else {
            const int p = idx - d.na - d.nb;
            const int i = FastDiv(d.mb).div(p), j = p - i * d.nb;
            ldc<9>(d.pa, d.cpa, i, a); ldc<3>(d.pb, d.cpb, j, b);
            matvec_rn(a, b, o);
        }
        break;
    }
    case K3_MERGE33: {
        if ((int)idx < d.na) ldc<3>(d.pa, d.cpa, idx, o); else ldc<3>(d.pb, d.cpb, idx - d.na, o);
        break;
    }
    case K3_MERGE31: {
        if ((int)idx < d.na) ldc<3>(d.pa, d.cpa, idx, o);
        else { const double s = d.pb[idx - d.na]; o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; if (d.row == 0) o[0] = s; else if (d.row == 1) o[1] = s; else o[2] = s; }
        break;
    }
    case K3_CROSS: {
        double a[3], b[3];
        if ((int)idx < d.na) { ldc<3>(d.pa, d.cpa, idx, a); b[0] = d.v[0]; b[1] = d.v[1]; b[2] = d.v[2]; }
        else if ((int)idx < d.na + d.nb) { ldc<3>(d.pb, d.cpb, idx - d.na, b); a[0] = d.m[0]; a[1] = d.m[1]; a[2] = d.m[2]; }
        else {
            const int p = idx - d.na - d.nb;
            const int i = FastDiv(d.mb).div(p), j = p - i * d.nb;
            ldc<3>(d.pa, d.cpa, i, a); ldc<3>(d.pb, d.cpb, j, b);
        }
        o[0] = mul_rn(a[1], b[2]); o[1] = mul_rn(a[2], b[1]);
        o[2] = mul_rn(a[2], b[0]); o[3] = mul_rn(a[0], b[2]);
        o[4] = mul_rn(a[0], b[1]); o[5] = mul_rn(a[1], b[0]);
        break;
    }
    case K3_CROSS_CONST: { double z[3]; ldc<3>(d.pa, d.cpa, idx, z); cross_const_comp(d.m, d.flag, z, o); break; }
    case K3_CONST_LEFT: {
        double z[3];
        ldc<3>(d.pa, d.cpa, idx, z);
        if (d.flag) { o[0] = mul_rn(d.m[0], z[0]); o[1] = mul_rn(d.m[0], z[1]); o[2] = mul_rn(d.m[0], z[2]); }
        else matvec_rn(d.m, z, o);
        break;
    }
    default: { double r9[9]; ldc<9>(d.pa, d.cpa, idx, r9); matvec_rn(r9, d.m, o); break; }   // K3_CONST_RIGHT
    }
}
// threshold stages of the operation's simplify() calls; out[3], drop[3] (drop = |dropped| per component)
__device__ __forceinline__ bool op3_finish(const Op3& d, const double* acc, double* out, double* drop, double thr) {
    if (d.kind == K3_CROSS) {   // each scalar product, then the difference, then the stacked vector (see CrossPPOp)
        bool any = false;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double p = acc[2 * c], q = acc[2 * c + 1];
            const bool kp = norm1(&p) > thr, kq = norm1(&q) > thr;
            double dd = 0.0;
            if (!kp) dd = __dadd_ru(dd, fabs(p));
            if (!kq) dd = __dadd_ru(dd, fabs(q));
            double vv = 0.0;
            bool present = false;
            if (kp || kq) {
                vv = kp ? (kq ? add_rn(p, -q) : p) : -q;
                if (norm1(&vv) > thr) present = true;
                else { dd = __dadd_ru(dd, fabs(vv)); vv = 0.0; }
            }
            out[c] = vv; drop[c] = dd; any |= present;
        }
        if (!any) return false;
        if (norm3(out) <= thr) {
#pragma unroll
            for (int c = 0; c < 3; c++) drop[c] = __dadd_ru(drop[c], fabs(out[c]));
            return false;
        }
        return true;
    }
    if (d.kind == K3_CROSS_CONST) {   // component-wise simplify of the scaled differences, then the stacked vector (see CrossConstOp)
        bool any = false;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (norm1(&acc[c]) > thr) { out[c] = acc[c]; any = true; drop[c] = 0.0; }
            else { out[c] = 0.0; drop[c] = fabs(acc[c]); }
        }
        if (!any) return false;
        if (norm3(out) <= thr) {
#pragma unroll
            for (int c = 0; c < 3; c++) drop[c] = __dadd_ru(drop[c], fabs(out[c]));
            return false;
        }
        return true;
    }
    if (norm3(acc) <= thr) {
#pragma unroll
        for (int c = 0; c < 3; c++) drop[c] = fabs(acc[c]);
        return false;
    }
#pragma unroll
    for (int c = 0; c < 3; c++) out[c] = acc[c];
    return true;
}
template <int NT>
__device__ __forceinline__ void op3_epilogue_begin(ScalarEpilogue<NT, 3>& se, const Op3& d) {
    switch (d.kind) {
    case K3_MUL93: se.begin(MulEpi<9, 3, 3>{*(const PZ<9>*)d.A, *(const PZ<3>*)d.B}); break;
    case K3_MERGE33: se.begin(MergeEpi<3, 3>{view(*(const PZ<3>*)d.A), view(*(const PZ<3>*)d.B), false}); break;
    case K3_MERGE31: se.begin(MergeEpi<3, 1>{view(*(const PZ<3>*)d.A), view_place(*(const PZ<1>*)d.B, d.row), false}); break;
    case K3_CROSS: se.begin(CrossPPEpi{*(const PZ<3>*)d.A, *(const PZ<3>*)d.B}); break;
    case K3_CROSS_CONST: se.begin(CrossConstEpi{*(const PZ<3>*)d.A, {d.m[0], d.m[1], d.m[2]}, d.flag}); break;
    case K3_CONST_LEFT: se.begin(ConstLeftEpi{*(const PZ<3>*)d.A, d.Mc, d.Mi0, d.Mi1, d.flag}); break;
    default: se.begin(ConstRightEpi{*(const PZ<9>*)d.A, {d.m[0], d.m[1], d.m[2]}}); break;
    }
}
template <int NT, bool BIG>
__device__ __forceinline__ void pz_op3_impl(Scratch& S, PZ<3>& dst, const Op3& d, int N) {
    u64* key0 = Buf<BIG>::key(S, 0);
    u16* idx0 = Buf<BIG>::idx(S, 0);
    int W = 1;
    u64 magicW = 0;
    const int na = d.na, nb = d.nb;
    if (N > 0) {
        if (d.kind == K3_MUL93 || d.kind == K3_CROSS) fill_product_keys<NT, BIG>(S, d.ka, na, d.ma, d.kb, nb, d.mb, W, magicW);
        else if (d.kind == K3_MERGE33 || d.kind == K3_MERGE31) {
            if (na >= nb) {   // the longer run first: runs must have uniform width except the last
                W = na > 0 ? na : 1; magicW = d.ma;
                #pragma unroll 1
                for (int i = gtid<NT>(); i < na; i += NT) { key0[i] = d.ka[i]; idx0[i] = (u16)i; }
                #pragma unroll 1
                for (int j = gtid<NT>(); j < nb; j += NT) { key0[na + j] = d.kb[j]; idx0[na + j] = (u16)(na + j); }
            }
            else {
                W = nb; magicW = d.mb;
                #pragma unroll 1
                for (int j = gtid<NT>(); j < nb; j += NT) { key0[j] = d.kb[j]; idx0[j] = (u16)(na + j); }
                #pragma unroll 1
                for (int i = gtid<NT>(); i < na; i += NT) { key0[nb + i] = d.ka[i]; idx0[nb + i] = (u16)i; }
            }
        }
        else {   // element-wise: one sorted run, every key its own segment
            W = N; magicW = d.ma;
            #pragma unroll 1
            for (int i = gtid<NT>(); i < N; i += NT) { key0[i] = d.ka[i]; idx0[i] = (u16)i; }
        }
    }
    gsync<NT>();
    phase_mark(PH_FILL);
    const int buf = merge_sort_runs<NT, BIG>(S, N, W, magicW);
    const u64* key = Buf<BIG>::key(S, buf);
    const u16* idx = Buf<BIG>::idx(S, buf);
    u16* flag = Buf<BIG>::idx(S, buf ^ 1);
    int ncap;
    double* tmp = S.staging(N, 3, ncap);
    double red[6];
#pragma unroll
    for (int c = 0; c < 6; c++) red[c] = 0.0;
    ScalarEpilogue<NT, 3> se;
    op3_epilogue_begin<NT>(se, d);
    const double thr = S.thr_sq;
    const bool six = d.kind == K3_CROSS;
    #pragma unroll 1
    for (int g = gtid<NT>(); g < N; g += NT) {
        const u64 k = key[g];
        u16 f = 0;
#ifdef ARMOUR_SEGSTAT
        segstat_record(N, g, g == 0 || key[g - 1] != k, key);
#endif
        if (g == 0 || key[g - 1] != k) {
            double acc[6];
            op3_term(d, idx[g], acc);
            #pragma unroll 1
            for (int e = g + 1; e < N && key[e] == k; e++) {
                double t[6];
                op3_term(d, idx[e], t);
#pragma unroll
                for (int c = 0; c < 3; c++) acc[c] = add_rn(acc[c], t[c]);
                if (six) {
#pragma unroll
                    for (int c = 3; c < 6; c++) acc[c] = add_rn(acc[c], t[c]);
                }
            }
            double out[3], dr[3];
#pragma unroll
            for (int c = 0; c < 3; c++) dr[c] = 0.0;
            const bool keep = op3_finish(d, acc, out, dr, thr);
#pragma unroll
            for (int c = 0; c < 3; c++) red[c] = __dadd_ru(red[c], dr[c]);
            if (keep) {
                f = 1;
#pragma unroll
                for (int c = 0; c < 3; c++) { tmp[c * ncap + g] = out[c]; red[3 + c] = __dadd_ru(red[3 + c], fabs(out[c])); }
            }
        }
        flag[g] = f;
    }
    gsync<NT>();
    phase_mark(PH_SEGMENT);
    compact_emit<NT, 3>(S, N, key, flag, tmp, ncap, red, se, dst);
}
// The descriptor is assembled INSIDE the one non-inlined function, from a handful of scalar arguments: a struct handed across
// the call boundary by reference would live in local memory, and every field access in the inner loops would be a load.
template <int NT>
__device__ __noinline__ void pz_op3(Scratch& S, PZ<3>& dst, int kind, const void* A, const void* B, const double* konst, int row, bool flag,
                                    const double* Mi0, const double* Mi1) {
    Op3 d;
    d.kind = kind; d.A = A; d.B = B; d.row = row; d.flag = flag; d.Mc = konst; d.Mi0 = Mi0; d.Mi1 = Mi1;
    d.kb = nullptr; d.nb = 0; d.mb = 0; d.pb = nullptr; d.cpb = 0;
#pragma unroll
    for (int c = 0; c < 9; c++) d.m[c] = 0.0;
#pragma unroll
    for (int c = 0; c < 3; c++) d.v[c] = 0.0;
    if (kind == K3_MUL93 || kind == K3_CONST_RIGHT) {
        const PZ<9>& a = *(const PZ<9>*)A;
        d.ka = a.keys; d.na = a.n; d.ma = a.divM; d.pa = a.coef; d.cpa = a.cap;
        if (kind == K3_MUL93) {
#pragma unroll
            for (int c = 0; c < 9; c++) d.m[c] = a.center[c];
        }
    }
    else {
        const PZ<3>& a = *(const PZ<3>*)A;
        d.ka = a.keys; d.na = a.n; d.ma = a.divM; d.pa = a.coef; d.cpa = a.cap;
        if (kind == K3_CROSS) { d.m[0] = a.center[0]; d.m[1] = a.center[1]; d.m[2] = a.center[2]; }
    }
    if (kind == K3_MUL93 || kind == K3_MERGE33 || kind == K3_CROSS) {
        const PZ<3>& b = *(const PZ<3>*)B;
        d.kb = b.keys; d.nb = b.n; d.mb = b.divM; d.pb = b.coef; d.cpb = b.cap;
        if (kind != K3_MERGE33) { d.v[0] = b.center[0]; d.v[1] = b.center[1]; d.v[2] = b.center[2]; }
    }
    else if (kind == K3_MERGE31) {
        const PZ<1>& b = *(const PZ<1>*)B;
        d.kb = b.keys; d.nb = b.n; d.mb = b.divM; d.pb = b.coef; d.cpb = b.cap;
    }
    if (kind == K3_CROSS_CONST || kind == K3_CONST_RIGHT) { d.m[0] = konst[0]; d.m[1] = konst[1]; d.m[2] = konst[2]; }
    else if (kind == K3_CONST_LEFT) {
        d.m[0] = konst[0];
        if (!flag) {
#pragma unroll
            for (int c = 1; c < 9; c++) d.m[c] = konst[c];
        }
    }
    int N = (kind == K3_MUL93 || kind == K3_CROSS) ? d.na + d.nb + d.na * d.nb : (kind == K3_MERGE33 || kind == K3_MERGE31) ? d.na + d.nb : d.na;
    if (N > S.ncap || N > 65535) { if (gtid<NT>() == 0) set_err(S, ERR_ENTRY_CAP); N = 0; }
    if (N <= S.scap) pz_op3_impl<NT, false>(S, dst, d, N);
    else pz_op3_impl<NT, true>(S, dst, d, N);
}
// the seven operations
template <int NT> __device__ __forceinline__ void op3_mul93(Scratch& S, PZ<3>& dst, const PZ<9>& A, const PZ<3>& B) { pz_op3<NT>(S, dst, K3_MUL93, &A, &B, nullptr, 0, false, nullptr, nullptr); }
template <int NT> __device__ __forceinline__ void op3_add(Scratch& S, PZ<3>& dst, const PZ<3>& A, const PZ<3>& B) { pz_op3<NT>(S, dst, K3_MERGE33, &A, &B, nullptr, 0, false, nullptr, nullptr); }
template <int NT> __device__ __forceinline__ void op3_add_one_dim(Scratch& S, PZ<3>& dst, const PZ<3>& A, const PZ<1>& s, int row) { pz_op3<NT>(S, dst, K3_MERGE31, &A, &s, nullptr, row, false, nullptr, nullptr); }
template <int NT> __device__ __forceinline__ void op3_cross(Scratch& S, PZ<3>& dst, const PZ<3>& A, const PZ<3>& B) { pz_op3<NT>(S, dst, K3_CROSS, &A, &B, nullptr, 0, false, nullptr, nullptr); }
template <int NT> __device__ __forceinline__ void op3_cross_const(Scratch& S, PZ<3>& dst, const PZ<3>& Z, const double* kvec, bool const_first) {
    pz_op3<NT>(S, dst, K3_CROSS_CONST, &Z, nullptr, kvec, 0, const_first, nullptr, nullptr);
}
template <int NT> __device__ __forceinline__ void op3_const_left(Scratch& S, PZ<3>& dst, const double* Mc, const double* Mi0, const double* Mi1, bool scalar, const PZ<3>& V) {
    pz_op3<NT>(S, dst, K3_CONST_LEFT, &V, nullptr, Mc, 0, scalar, Mi0, Mi1);
}
template <int NT> __device__ __forceinline__ void op3_const_right(Scratch& S, PZ<3>& dst, const PZ<9>& R, const double* pvec) {
    pz_op3<NT>(S, dst, K3_CONST_RIGHT, &R, nullptr, pvec, 0, false, nullptr, nullptr);
}

// reset to a monomial-free PZ with centre c (all threads call; thread 0 writes)
template <int NT, int D>
__device__ void pz_set_const(PZ<D>& z, const double* c) {
    if (gtid<NT>() == 0) {
        z.n = 0; z.divM = FastDiv::magic(0); z.ormask = 0;
        for (int i = 0; i < D; i++) { z.center[i] = c ? c[i] : 0.0; z.ind[0][i] = 0.0; z.ind[1][i] = 0.0; z.abss[i] = 0.0; }
    }
    gsync<NT>();
}

}  // namespace armour
