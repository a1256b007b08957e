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

enum { ERR_NONE = 0, ERR_ENTRY_CAP = 1, ERR_MONO_CAP = 2, ERR_TABLE_CAP = 4, ERR_LINK_GEN = 8, ERR_DEGREE = 16 };

template <int D>
struct PZ {
    int n;          // monomials
    int cap;        // capacity (plane stride)
    u64* keys;      // [cap]
    double* coef;   // [D][cap], column-major element order inside a 3x3 (row + 3*col), like Eigen
    double center[D];
    double ind[2][D];   // interval radius: [0] nominal inertial parameters, [1] uncertain ones
    double abss[D];     // sum_i |coef_i| rounded up
};

// per-CTA scratch (shared memory) ------------------------------------------------------------
struct Scratch {
    // Sort ping-pong buffers live in dynamic shared memory.  They are kept as shared-window addresses and
    // turned back into pointers with __cvta_shared_to_generic at each use, so that the compiler emits
    // LDS/STS instead of generic loads (a pointer loaded from a struct has no known address space).
    unsigned skey_a[2];   // u64[ncap]
    unsigned sidx_a[2];   // u16[ncap]
    double* tmp;      // global [9][ncap]: per-key results before compaction
    int ncap;
    double thr;
    int* gerr;        // global error word
    double red[32 * 32];
    int iscan[34];
    __device__ __forceinline__ u64* skey(int b) const { return (u64*)__cvta_shared_to_generic((size_t)skey_a[b]); }
    __device__ __forceinline__ u16* sidx(int b) const { return (u16*)__cvta_shared_to_generic((size_t)sidx_a[b]); }
    __device__ void bind(unsigned char* smem, int ncap_) {
        const unsigned base = (unsigned)__cvta_generic_to_shared(smem);
        skey_a[0] = base; skey_a[1] = base + (unsigned)ncap_ * 8;
        sidx_a[0] = base + (unsigned)ncap_ * 16; sidx_a[1] = base + (unsigned)ncap_ * 18;
        ncap = ncap_;
    }
};

// exact n / d for n, d < 2^16 with one multiply: q = (n * M) >> 32, M = floor((2^32 - 1) / d) + 1
struct FastDiv {
    u64 M; int d;
    __device__ __forceinline__ explicit FastDiv(int d_) : M((u64)(0xFFFFFFFFu / (unsigned)(d_ > 0 ? d_ : 1)) + 1), d(d_ > 0 ? d_ : 1) {}
    __device__ __forceinline__ int div(int n) const { return (int)(((u64)(unsigned)n * M) >> 32); }
};

__device__ __forceinline__ void set_err(Scratch& S, int e) { atomicOr(S.gerr, e); }

// ---- small fp helpers (round-to-nearest, no contraction; -fmad=false is also set) -----------
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
// Frobenius norms in Eigen 3.3's reduction order (see oracle/oracle_pz.hpp Mat::squaredNorm)
__device__ __forceinline__ double norm1(const double* v) { return __dsqrt_rn(mul_rn(v[0], v[0])); }
__device__ __forceinline__ double norm3(const double* v) {
    return __dsqrt_rn(add_rn(add_rn(mul_rn(v[0], v[0]), mul_rn(v[1], v[1])), mul_rn(v[2], v[2])));
}
__device__ __forceinline__ double norm9(const double* v) {
    double p0a = add_rn(mul_rn(v[0], v[0]), mul_rn(v[4], v[4]));
    double p0b = add_rn(mul_rn(v[1], v[1]), mul_rn(v[5], v[5]));
    double p1a = add_rn(mul_rn(v[2], v[2]), mul_rn(v[6], v[6]));
    double p1b = add_rn(mul_rn(v[3], v[3]), mul_rn(v[7], v[7]));
    double s = add_rn(add_rn(p0a, p1a), add_rn(p0b, p1b));
    return __dsqrt_rn(add_rn(s, mul_rn(v[8], v[8])));
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
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    double v[VP];
#pragma unroll
    for (int k = 0; k < VP; k++) v[k] = k < V ? red[k] : 0.0;
    warp_multi_sum_ru<VP>(v);
    if (lane == 31) S.iscan[warp] = incl;
    if ((lane & ((32 >> Log2<VP>::value) - 1)) == 0) S.red[warp * VP + (lane >> (5 - Log2<VP>::value))] = v[0];
    __syncthreads();
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
#pragma unroll
    for (int w = 0; w < NT / 32; w++) s = __dadd_ru(s, S.red[w * VP + k]);
    return s;
}

// ---- run-structured merge sort on (key, idx) ------------------------------------------------
// Buffer 0 holds N entries laid out as sorted runs of width W (the last run may be shorter).
// Returns the buffer that holds the fully sorted sequence.  All (key, idx) pairs are distinct, so
// ordering by (key, idx) equals a stable sort by key of the list in origin order.
template <int NT>
__device__ __forceinline__ int merge_sort_runs(Scratch& S, int N, int W) {
    int cur = 0;
    const FastDiv fd(W);
    int level = 0;
    for (int w = W; w < N; w <<= 1, level++) {
        const u64* ki = S.skey(cur);
        const u16* ii = S.sidx(cur);
        u64* ko = S.skey(cur ^ 1);
        u16* io = S.sidx(cur ^ 1);
        for (int g = threadIdx.x; g < N; g += NT) {
            const int r = fd.div(g) >> level;      // g / (W << level)
            const int base = r * w;
            const int sb = (r ^ 1) * w;
            const u64 k = ki[g];
            const unsigned id = ii[g];
            int pos = g;
            if (sb < N) {
                int lo = sb, hi = min(sb + w, N);
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const u64 km = ki[mid];
                    const bool less = (km < k) || (km == k && ii[mid] < id);
                    if (less) lo = mid + 1; else hi = mid;
                }
                pos = min(base, sb) + (g - base) + (lo - sb);
            }
            ko[pos] = k;
            io[pos] = (u16)id;
        }
        __syncthreads();
        cur ^= 1;
    }
    return cur;
}

// ---- generic "segment-reduce, threshold, compact" -------------------------------------------
// Op interface:
//   static constexpr int NACC;                          accumulators per key
//   void first(unsigned idx, double* acc);              acc  = contribution of the segment's first entry
//   void next (unsigned idx, double* acc);              acc += contribution (round-to-nearest, in order)
//   bool finish(const double* acc, double* out, double* drop);   threshold logic; out[DOUT]; drop[DOUT] = |dropped|
// cen / rad: centre and the two radii of the result before the dropped monomials are added (computed by the
// caller from the operands before anything is written, because dst may alias an operand).
// Barriers: one after the segment pass, one inside block_scan_sum, one at the end.
template <int NT, int DOUT, class Op>
__device__ __forceinline__ void reduce_emit(Scratch& S, int buf, int N, Op& op, PZ<DOUT>& dst, const double* cen, const double (*rad)[DOUT]) {
    const u64* key = S.skey(buf);
    const u16* idx = S.sidx(buf);
    u16* flag = S.sidx(buf ^ 1);
    double* tmp = S.tmp;
    const int ncap = S.ncap;
    double red[2 * DOUT];
#pragma unroll
    for (int c = 0; c < 2 * DOUT; c++) red[c] = 0.0;
    // pass 1: one thread per segment head
    for (int g = threadIdx.x; g < N; g += NT) {
        const u64 k = key[g];
        u16 f = 0;
        if (g == 0 || key[g - 1] != k) {
            double acc[Op::NACC];
            op.first(idx[g], acc);
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
    __syncthreads();
    // compaction of the kept keys (blocked ranges keep the order)
    const int ipt = (N + NT - 1) / NT;
    const int g0 = min(threadIdx.x * ipt, N), g1 = min(g0 + ipt, N);
    int cnt = 0;
    for (int g = g0; g < g1; g++) cnt += flag[g];
    int total;
    int off = block_scan_sum<NT, 2 * DOUT>(S, cnt, red, total);
    const int dcap = dst.cap;
    if (total > dcap) { if (threadIdx.x == 0) set_err(S, ERR_MONO_CAP); total = 0; }
    else {
        u64* dk = dst.keys;
        double* dc = dst.coef;
        for (int g = g0; g < g1; g++) {
            if (flag[g]) {
                dk[off] = key[g];
#pragma unroll
                for (int c = 0; c < DOUT; c++) dc[c * dcap + off] = tmp[c * ncap + g];
                off++;
            }
        }
    }
    if (threadIdx.x < DOUT) {   // scalar epilogue, one component per thread
        const int c = threadIdx.x;
        const double drop = inflate(block_total<NT, 2 * DOUT>(S, c), N);
        dst.abss[c] = inflate(block_total<NT, 2 * DOUT>(S, DOUT + c), total);
        dst.center[c] = cen[c];
        dst.ind[0][c] = __dadd_ru(rad[0][c], drop);
        dst.ind[1][c] = __dadd_ru(rad[1][c], drop);
        if (c == 0) dst.n = total;
    }
    __syncthreads();
}

// elementwise variant: no sort, keys are those of `src` in order; op computes out from index i.
//   bool Op::finish(int i, double* out, double* drop)
template <int NT, int DOUT, class Op>
__device__ __forceinline__ void elementwise_emit(Scratch& S, int n, const u64* src_keys, Op& op, PZ<DOUT>& dst, const double* cen, const double (*rad)[DOUT]) {
    u16* flag = S.sidx(0);
    u64* kcopy = S.skey(0);
    double* tmp = S.tmp;
    const int ncap = S.ncap;
    double red[2 * DOUT];
#pragma unroll
    for (int c = 0; c < 2 * DOUT; c++) red[c] = 0.0;
    if (n > ncap) { if (threadIdx.x == 0) set_err(S, ERR_ENTRY_CAP); n = 0; }
    for (int i = threadIdx.x; i < n; i += NT) {
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
    __syncthreads();
    const int ipt = (n + NT - 1) / NT;
    const int g0 = min(threadIdx.x * ipt, n), g1 = min(g0 + ipt, n);
    int cnt = 0;
    for (int g = g0; g < g1; g++) cnt += flag[g];
    int total;
    int off = block_scan_sum<NT, 2 * DOUT>(S, cnt, red, total);
    const int dcap = dst.cap;
    if (total > dcap) { if (threadIdx.x == 0) set_err(S, ERR_MONO_CAP); total = 0; }
    else {
        u64* dk = dst.keys;
        double* dc = dst.coef;
        for (int g = g0; g < g1; g++) {
            if (flag[g]) {
                dk[off] = kcopy[g];
#pragma unroll
                for (int c = 0; c < DOUT; c++) dc[c * dcap + off] = tmp[c * ncap + g];
                off++;
            }
        }
    }
    if (threadIdx.x < DOUT) {
        const int c = threadIdx.x;
        const double drop = inflate(block_total<NT, 2 * DOUT>(S, c), n);
        dst.abss[c] = inflate(block_total<NT, 2 * DOUT>(S, DOUT + c), total);
        dst.center[c] = cen[c];
        dst.ind[0][c] = __dadd_ru(rad[0][c], drop);
        dst.ind[1][c] = __dadd_ru(rad[1][c], drop);
        if (c == 0) dst.n = total;
    }
    __syncthreads();
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
    double thr;
    double ca[DA], cb[DB];
    __device__ MulOp(const PZ<DA>& a, const PZ<DB>& b, double t) : A(a), B(b), na(a.n), nb(b.n), fdb(b.n), thr(t) {
#pragma unroll
        for (int c = 0; c < DA; c++) ca[c] = a.center[c];
#pragma unroll
        for (int c = 0; c < DB; c++) cb[c] = b.center[c];
    }
    __device__ __forceinline__ void term(unsigned idx, double* o) const {
        double a[DA], b[DB];
        if ((int)idx < na) { ldc<DA>(A, idx, a); coef_mul<DA, DB, DO>(a, cb, o); }
        else if ((int)idx < na + nb) { ldc<DB>(B, idx - na, b); coef_mul<DA, DB, DO>(ca, b, o); }
        else {
            const int p = idx - na - nb;
            const int i = fdb.div(p), j = p - i * nb;
            ldc<DA>(A, i, a); ldc<DB>(B, j, b);
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
template <int NT>
__device__ int fill_product_keys(Scratch& S, const u64* ka, int na, const u64* kb, int nb, int& W) {
    const int N = na + nb + na * nb;
    u64* key = S.skey(0);
    u16* idx = S.sidx(0);
    if (na == 0 || nb == 0) {   // single sorted run
        W = N > 0 ? N : 1;
        for (int g = threadIdx.x; g < N; g += NT) { key[g] = na ? ka[g] : kb[g]; idx[g] = (u16)g; }
        return N;
    }
    if (nb >= na) {   // runs: i = 0..na-1 (width nb), then B's own list (nb), then A's own list (na <= nb, last)
        W = nb;
        const FastDiv fd(nb);
        for (int g = threadIdx.x; g < na * nb; g += NT) {
            const int i = fd.div(g), j = g - i * nb;
            key[g] = ka[i] + kb[j];   // degrees add; no carry by construction (KPR/PZsparse.cu:938-940)
            idx[g] = (u16)(na + nb + g);
        }
        const int o1 = na * nb;
        for (int j = threadIdx.x; j < nb; j += NT) { key[o1 + j] = kb[j]; idx[o1 + j] = (u16)(na + j); }
        const int o0 = o1 + nb;
        for (int i = threadIdx.x; i < na; i += NT) { key[o0 + i] = ka[i]; idx[o0 + i] = (u16)i; }
    }
    else {            // runs: j = 0..nb-1 (width na), then A's own list (na), then B's own list (nb < na, last)
        W = na;
        const FastDiv fd(na);
        for (int g = threadIdx.x; g < na * nb; g += NT) {
            const int j = fd.div(g), i = g - j * na;
            key[g] = ka[i] + kb[j];
            idx[g] = (u16)(na + nb + i * nb + j);
        }
        const int o0 = na * nb;
        for (int i = threadIdx.x; i < na; i += NT) { key[o0 + i] = ka[i]; idx[o0 + i] = (u16)i; }
        const int o1 = o0 + na;
        for (int j = threadIdx.x; j < nb; j += NT) { key[o1 + j] = kb[j]; idx[o1 + j] = (u16)(na + j); }
    }
    return N;
}

// radius of a product, rounded up (KPR/PZsparse.cu:944-989), for radius variant v
template <int DA, int DB, int DO>
__device__ void product_radius(const double* ca, const double* abssa, const double* inda, const double* cb, const double* abssb, const double* indb,
                               const double* drop, double* out) {
    double ma[DA], mb[DB];
#pragma unroll
    for (int c = 0; c < DA; c++) ma[c] = __dadd_ru(fabs(ca[c]), abssa[c]);
#pragma unroll
    for (int c = 0; c < DB; c++) mb[c] = __dadd_ru(fabs(cb[c]), abssb[c]);
    double r2[DO], r3[DO], r1[DO];
    if (DA == 9 && DB == 3) { matvec_ru(ma, indb, r2); matvec_ru(inda, mb, r3); matvec_ru(inda, indb, r1); }
    else if (DA == 9 && DB == 9) { matmat_ru(ma, indb, r2); matmat_ru(inda, mb, r3); matmat_ru(inda, indb, r1); }
    else if (DA == 1 && DB == 1) { r2[0] = __dmul_ru(ma[0], indb[0]); r3[0] = __dmul_ru(inda[0], mb[0]); r1[0] = __dmul_ru(inda[0], indb[0]); }
    else {   // DA == 1: scalar times vector
#pragma unroll
        for (int c = 0; c < DO; c++) { r2[c] = __dmul_ru(ma[0], indb[c]); r3[c] = __dmul_ru(inda[0], mb[c]); r1[c] = __dmul_ru(inda[0], indb[c]); }
    }
#pragma unroll
    for (int c = 0; c < DO; c++) out[c] = __dadd_ru(__dadd_ru(r1[c], __dadd_ru(r2[c], r3[c])), drop[c]);
}

template <int NT, int DA, int DB, int DO>
__device__ __noinline__ void pz_mul(Scratch& S, PZ<DO>& dst, const PZ<DA>& A, const PZ<DB>& B) {
    const int na = A.n, nb = B.n;
    int N = na + nb + na * nb;
    if (N > S.ncap || N > 65535) { if (threadIdx.x == 0) set_err(S, ERR_ENTRY_CAP); N = 0; }
    int W = 1;
    if (N > 0) fill_product_keys<NT>(S, A.keys, na, B.keys, nb, W);
    // everything the scalar epilogue needs from A and B is read before dst (which may alias) is written
    MulOp<DA, DB, DO> op(A, B, S.thr);
    double cen[DO], rad[2][DO];
    double drop0[DO];
#pragma unroll
    for (int c = 0; c < DO; c++) drop0[c] = 0.0;
    coef_mul<DA, DB, DO>(op.ca, op.cb, cen);
    for (int v = 0; v < 2; v++) product_radius<DA, DB, DO>(A.center, A.abss, A.ind[v], B.center, B.abss, B.ind[v], drop0, rad[v]);
    __syncthreads();
    const int buf = merge_sort_runs<NT>(S, N, W);
    reduce_emit<NT, DO, MulOp<DA, DB, DO>>(S, buf, N, op, dst, cen, rad);
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
template <int DA, int DB, int DO>
struct MergeOp {
    static constexpr int NACC = DO;
    View<DA> A;
    View<DB> B;
    int na;
    bool negb;
    double thr;
    __device__ __forceinline__ void term(unsigned idx, double* o) const {
        if ((int)idx < na) { double a[DA]; ldc<DA>(*A.p, idx, a); view_vec<DA, DO>(A, a, o); }
        else {
            double b[DB]; ldc<DB>(*B.p, idx - na, b); view_vec<DB, DO>(B, b, o);
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
// dst = A (+/-) B through views.  The centre of the VIEW_PLACE / VIEW_EXTRACT source is mapped the same way.
template <int D, int DO>
__device__ __forceinline__ void view_radius(const View<D>& v, const double* ind, double* o) {   // radii are non-negative: scale by |s| rounded up
    const double sc = fabs(v.scale);
    if (v.mode == VIEW_SAME) {
#pragma unroll
        for (int c = 0; c < DO; c++) o[c] = v.scaled ? __dmul_ru(sc, ind[c < D ? c : 0]) : ind[c < D ? c : 0];
    }
    else if (v.mode == VIEW_EXTRACT) o[0] = v.scaled ? __dmul_ru(sc, ind[v.row < D ? v.row : 0]) : ind[v.row < D ? v.row : 0];
    else {
#pragma unroll
        for (int c = 0; c < DO; c++) o[c] = 0.0;
        o[v.row < DO ? v.row : 0] = v.scaled ? __dmul_ru(sc, ind[0]) : ind[0];
    }
}
// dst = A (+/-) B through views.  Centre and radii of a VIEW_PLACE / VIEW_EXTRACT source are mapped the same way.
template <int NT, int DA, int DB, int DO>
__device__ __noinline__ void pz_merge(Scratch& S, PZ<DO>& dst, const View<DA> A, const View<DB> B, bool negb) {
    const int na = A.p->n, nb = B.p->n;
    int N = na + nb;
    if (N > S.ncap) { if (threadIdx.x == 0) set_err(S, ERR_ENTRY_CAP); N = 0; }
    u64* key = S.skey(0);
    u16* idx = S.sidx(0);
    int W = 1;
    if (N > 0) {
        const u64* ka = A.p->keys; const u64* kb = B.p->keys;
        if (na >= nb) {   // the longer run first: runs must have uniform width except the last
            W = na > 0 ? na : 1;
            for (int i = threadIdx.x; i < na; i += NT) { key[i] = ka[i]; idx[i] = (u16)i; }
            for (int j = threadIdx.x; j < nb; j += NT) { key[na + j] = kb[j]; idx[na + j] = (u16)(na + j); }
        }
        else {
            W = nb;
            for (int j = threadIdx.x; j < nb; j += NT) { key[j] = kb[j]; idx[j] = (u16)(na + j); }
            for (int i = threadIdx.x; i < na; i += NT) { key[nb + i] = ka[i]; idx[nb + i] = (u16)i; }
        }
    }
    MergeOp<DA, DB, DO> op{A, B, na, negb, S.thr};
    double cen[DO], rad[2][DO];
    {
        double ca[DO], cb[DO];
        view_vec<DA, DO>(A, A.p->center, ca);
        view_vec<DB, DO>(B, B.p->center, cb);
#pragma unroll
        for (int c = 0; c < DO; c++) cen[c] = negb ? add_rn(ca[c], -cb[c]) : add_rn(ca[c], cb[c]);
        for (int v = 0; v < 2; v++) {
            double ia[DO], ib[DO];
            view_radius<DA, DO>(A, A.p->ind[v], ia);
            view_radius<DB, DO>(B, B.p->ind[v], ib);
#pragma unroll
            for (int c = 0; c < DO; c++) rad[v][c] = __dadd_ru(ia[c], ib[c]);
        }
    }
    __syncthreads();
    const int buf = merge_sort_runs<NT>(S, N, W);
    reduce_emit<NT, DO, MergeOp<DA, DB, DO>>(S, buf, N, op, dst, cen, rad);
}
template <int NT> __device__ __forceinline__ void pz_add3(Scratch& S, PZ<3>& dst, const PZ<3>& a, const PZ<3>& b) { pz_merge<NT, 3, 3, 3>(S, dst, view(a), view(b), false); }
// dst = a with the scalar PZ s added into row `row`   (addOneDimPZ)
template <int NT> __device__ __forceinline__ void pz_add_one_dim(Scratch& S, PZ<3>& dst, const PZ<3>& a, const PZ<1>& s, int row) {
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
    double thr;
    double ca[3], cb[3];
    __device__ CrossPPOp(const PZ<3>& a, const PZ<3>& b, double t) : A(a), B(b), na(a.n), nb(b.n), fdb(b.n), thr(t) {
        for (int c = 0; c < 3; c++) { ca[c] = a.center[c]; cb[c] = b.center[c]; }
    }
    __device__ __forceinline__ void term(unsigned idx, double* o) const {
        double a[3], b[3];
        if ((int)idx < na) { ldc<3>(A, idx, a); b[0] = cb[0]; b[1] = cb[1]; b[2] = cb[2]; }
        else if ((int)idx < na + nb) { ldc<3>(B, idx - na, b); a[0] = ca[0]; a[1] = ca[1]; a[2] = ca[2]; }
        else {
            const int p = idx - na - nb;
            const int i = fdb.div(p), j = p - i * nb;
            ldc<3>(A, i, a); ldc<3>(B, j, b);
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
template <int NT>
__device__ __noinline__ void pz_cross_pp(Scratch& S, PZ<3>& dst, const PZ<3>& A, const PZ<3>& B) {
    const int na = A.n, nb = B.n;
    int N = na + nb + na * nb;
    if (N > S.ncap || N > 65535) { if (threadIdx.x == 0) set_err(S, ERR_ENTRY_CAP); N = 0; }
    int W = 1;
    if (N > 0) fill_product_keys<NT>(S, A.keys, na, B.keys, nb, W);
    CrossPPOp op(A, B, S.thr);
    double cen[3], rad[2][3];
    // component c of the result is P - Q with P = a[c+1]*b[c+2], Q = a[c+2]*b[c+1]
    for (int c = 0; c < 3; c++) {
        const int i1 = (c + 1) % 3, i2 = (c + 2) % 3;
        cen[c] = add_rn(mul_rn(op.ca[i1], op.cb[i2]), -mul_rn(op.ca[i2], op.cb[i1]));
        for (int v = 0; v < 2; v++) {
            double z = 0.0, rp, rq;
            product_radius<1, 1, 1>(&A.center[i1], &A.abss[i1], &A.ind[v][i1], &B.center[i2], &B.abss[i2], &B.ind[v][i2], &z, &rp);
            product_radius<1, 1, 1>(&A.center[i2], &A.abss[i2], &A.ind[v][i2], &B.center[i1], &B.abss[i1], &B.ind[v][i1], &z, &rq);
            rad[v][c] = __dadd_ru(rp, rq);
        }
    }
    __syncthreads();
    const int buf = merge_sort_runs<NT>(S, N, W);
    reduce_emit<NT, 3, CrossPPOp>(S, buf, N, op, dst, cen, rad);
}

// =============================================================================================
// Elementwise operations (operand lists with identical keys: no sort needed)
// =============================================================================================
// cross(PZ a, const b) and cross(const a, PZ b)   (KPR/PZsparse.cu:1118-1132, 1153-1167)
struct CrossConstOp {
    const PZ<3>& Z;
    double k[3];
    bool const_first;   // true: cross(k, Z); false: cross(Z, k)
    double thr;
    __device__ __forceinline__ void comp(const double* z, double* r) const {
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
    __device__ __forceinline__ bool finish(int i, double* out, double* drop) const {
        double z[3], r[3];
        ldc<3>(Z, i, z);
        comp(z, r);
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
template <int NT>
__device__ __noinline__ void pz_cross_const(Scratch& S, PZ<3>& dst, const PZ<3>& Z, const double* kvec, bool const_first) {
    CrossConstOp op{Z, {kvec[0], kvec[1], kvec[2]}, const_first, S.thr};
    double cen[3], rad[2][3];
    op.comp(Z.center, cen);
    for (int v = 0; v < 2; v++)
        for (int c = 0; c < 3; c++) {
            const int i1 = (c + 1) % 3, i2 = (c + 2) % 3;
            // operator*(double) scales the radius by |k|, operator- adds the two radii.
            // cross(Z, k): r_c = z[i1]*k[i2] - z[i2]*k[i1];  cross(k, Z): r_c = k[i1]*z[i2] - k[i2]*z[i1]
            rad[v][c] = const_first ? __dadd_ru(__dmul_ru(Z.ind[v][i2], fabs(kvec[i1])), __dmul_ru(Z.ind[v][i1], fabs(kvec[i2])))
                                    : __dadd_ru(__dmul_ru(Z.ind[v][i1], fabs(kvec[i2])), __dmul_ru(Z.ind[v][i2], fabs(kvec[i1])));
        }
    const int n = Z.n;
    const u64* keys = Z.keys;
    elementwise_emit<NT, 3, CrossConstOp>(S, n, keys, op, dst, cen, rad);
}

// dst(3x1) = M * v with M a monomial-free 3x3 PZ (centre Mc, radii Mi[2]) — I_arr(i) * w
// dst(3x1) = m * v with m a monomial-free scalar PZ                         — mass_arr(i) * (...)
struct ConstLeftOp {
    const PZ<3>& V;
    double M[9];
    bool scalar;
    double thr;
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
template <int NT>
__device__ __noinline__ void pz_const_left(Scratch& S, PZ<3>& dst, const double* Mc, const double* Mi0, const double* Mi1, bool scalar, const PZ<3>& V) {
    ConstLeftOp op{V, {0}, scalar, S.thr};
    const int DM = scalar ? 1 : 9;
    for (int c = 0; c < DM; c++) op.M[c] = Mc[c];
    double cen[3], rad[2][3];
    double zero9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, drop0[3] = {0, 0, 0};
    if (scalar) { for (int c = 0; c < 3; c++) cen[c] = mul_rn(Mc[0], V.center[c]); }
    else matvec_rn(Mc, V.center, cen);
    for (int v = 0; v < 2; v++) {
        const double* Mi = v ? Mi1 : Mi0;
        if (scalar) product_radius<1, 3, 3>(Mc, zero9, Mi, V.center, V.abss, V.ind[v], drop0, rad[v]);
        else product_radius<9, 3, 3>(Mc, zero9, Mi, V.center, V.abss, V.ind[v], drop0, rad[v]);
    }
    const int n = V.n;
    const u64* keys = V.keys;
    elementwise_emit<NT, 3, ConstLeftOp>(S, n, keys, op, dst, cen, rad);
}

// dst(3x1) = R * p with R a 3x3 PZ and p a constant vector — FK_R * P   (KPR/Dynamics.cu:76)
struct ConstRightOp {
    const PZ<9>& R;
    double p[3];
    double thr;
    __device__ __forceinline__ bool finish(int i, double* out, double* drop) const {
        double m[9], r[3];
        ldc<9>(R, i, m);
        matvec_rn(m, p, r);
        if (norm3(r) <= thr) { for (int c = 0; c < 3; c++) drop[c] = fabs(r[c]); return false; }
        for (int c = 0; c < 3; c++) out[c] = r[c];
        return true;
    }
};
template <int NT>
__device__ __noinline__ void pz_const_right(Scratch& S, PZ<3>& dst, const PZ<9>& R, const double* pvec) {
    ConstRightOp op{R, {pvec[0], pvec[1], pvec[2]}, S.thr};
    double cen[3], rad[2][3];
    double zero3[3] = {0, 0, 0}, drop0[3] = {0, 0, 0};
    matvec_rn(R.center, pvec, cen);
    for (int v = 0; v < 2; v++) product_radius<9, 3, 3>(R.center, R.abss, R.ind[v], pvec, zero3, zero3, drop0, rad[v]);
    const int n = R.n;
    const u64* keys = R.keys;
    elementwise_emit<NT, 3, ConstRightOp>(S, n, keys, op, dst, cen, rad);
}

// reset to a monomial-free PZ with centre c (all threads call; thread 0 writes)
template <int D>
__device__ void pz_set_const(PZ<D>& z, const double* c) {
    if (threadIdx.x == 0) {
        z.n = 0;
        for (int i = 0; i < D; i++) { z.center[i] = c ? c[i] : 0.0; z.ind[0][i] = 0.0; z.ind[1][i] = 0.0; z.abss[i] = 0.0; }
    }
    __syncthreads();
}
// dst = src (deep copy)
template <int NT, int D>
__device__ void pz_copy(PZ<D>& dst, const PZ<D>& src) {
    __syncthreads();
    const int n = src.n;
    for (int i = threadIdx.x; i < n; i += NT) {
        dst.keys[i] = src.keys[i];
        for (int c = 0; c < D; c++) dst.coef[c * dst.cap + i] = src.coef[c * src.cap + i];
    }
    if (threadIdx.x == 0) {
        dst.n = n;
        for (int c = 0; c < D; c++) { dst.center[c] = src.center[c]; dst.ind[0][c] = src.ind[0][c]; dst.ind[1][c] = src.ind[1][c]; dst.abss[c] = src.abss[c]; }
    }
    __syncthreads();
}

}  // namespace armour
