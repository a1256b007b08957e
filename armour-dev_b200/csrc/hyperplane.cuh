// One half-space of the buffered-obstacle polytope (Obstacles::initializeHyperPlane = bufferObstaclesKernel + polytope_PH,
// KPR/CollisionChecking.cu:74-88, 136-228): shared by hyperplane_kernel (stand-alone launch) and by the copy fused into the
// tail of reach_build_kernel, so that both produce bit-identical tables.
#pragma once
#include "armour_types.cuh"

namespace armour {

// pair (a, b), a < b, of the 9 buffered generators in the reference's enumeration order
// (KPR/CollisionChecking.cu:26-39): (0,1) (0,2) ... (0,8) (1,2) ... (7,8)
static __constant__ unsigned char c_pair_a[COMB] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 6, 6, 7};
static __constant__ unsigned char c_pair_b[COMB] = {1, 2, 3, 4, 5, 6, 7, 8, 2, 3, 4, 5, 6, 7, 8, 3, 4, 5, 6, 7, 8, 4, 5, 6, 7, 8, 5, 6, 7, 8, 6, 7, 8, 7, 8, 8};

// The nine generators of one (link, obstacle) pair: the obstacle's three (Go[3][3]) and the link's six (Gl[6][3]); cen: the
// obstacle centre.  Writes A (unit normal or zero), d = A . c and delta = sum_j |A . g_j| of plane `p` at table index idx.
__device__ __forceinline__ const double* buffered_generator(const double* Go, const double* Gl, int k) { return k < 3 ? Go + 3 * k : Gl + 3 * (k - 3); }
__device__ __forceinline__ void compute_half_space(const double* Go, const double* Gl, const double* cen, int p, double& C0, double& C1, double& C2, double& d, double& delta) {
    const double* ga = buffered_generator(Go, Gl, c_pair_a[p]);
    const double* gb = buffered_generator(Go, Gl, c_pair_b[p]);
    const double ga0 = ga[0], ga1 = ga[1], ga2 = ga[2];
    const double gb0 = gb[0], gb1 = gb[1], gb2 = gb[2];
    const double cr0 = __dadd_rn(__dmul_rn(ga1, gb2), -__dmul_rn(ga2, gb1));
    const double cr1 = __dadd_rn(__dmul_rn(ga2, gb0), -__dmul_rn(ga0, gb2));
    const double cr2 = __dadd_rn(__dmul_rn(ga0, gb1), -__dmul_rn(ga1, gb0));
    const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(cr0, cr0), __dmul_rn(cr1, cr1)), __dmul_rn(cr2, cr2)));
    C0 = 0; C1 = 0; C2 = 0;
    if (nrm > 0) { C0 = __ddiv_rn(cr0, nrm); C1 = __ddiv_rn(cr1, nrm); C2 = __ddiv_rn(cr2, nrm); }
    d = __dadd_rn(__dadd_rn(__dmul_rn(C0, cen[0]), __dmul_rn(C1, cen[1])), __dmul_rn(C2, cen[2]));
    double dl = 0.0;
#pragma unroll
    for (int j = 0; j < 9; j++) {
        const double* gj = buffered_generator(Go, Gl, j);
        dl = __dadd_rn(dl, fabs(__dadd_rn(__dadd_rn(__dmul_rn(C0, gj[0]), __dmul_rn(C1, gj[1])), __dmul_rn(C2, gj[2]))));
    }
    delta = dl;
}
__device__ __forceinline__ void write_half_space(const Tables& tb, size_t idx, const double* Go, const double* Gl, const double* cen, int p) {
    double C0, C1, C2, d, dl;
    compute_half_space(Go, Gl, cen, p, C0, C1, C2, d, dl);
    tb.A[idx * 3 + 0] = C0; tb.A[idx * 3 + 1] = C1; tb.A[idx * 3 + 2] = C2;
    tb.d[idx] = d;
    tb.delta[idx] = dl;
}

// All half-spaces of one (problem, interval): 7 links x n_obs obstacles x 36 pairs, by every thread of the CTA that built the
// interval's link PZs (the generator blocks tb.gens[rec0 .. rec0 + 6] were written by this CTA's export_link).  `stage` is
// CTA-shared scratch of at least 7 * 18 + 12 * n_obs doubles that nobody else uses between the two barriers.
__device__ __forceinline__ void interval_half_spaces(const Tables& tb, int prob, size_t rec0, double* stage) {
    const int n_obs = tb.n_obs, nthreads = (int)blockDim.x;
    __syncthreads();                                  // every group is done with the interval: gens are written, scratch is free
    double* Lg = stage;                               // [7][6][3]
    double* Ob = stage + NJ * 18;                     // [n_obs][12]: centre, three generators
    for (int e = threadIdx.x; e < NJ * 18; e += nthreads) Lg[e] = tb.gens[rec0 * 18 + e];
    for (int e = threadIdx.x; e < n_obs * 12; e += nthreads) Ob[e] = tb.obstacles[(size_t)prob * n_obs * 12 + e];
    __syncthreads();
    const int per_link = n_obs * COMB;
    #pragma unroll 1
    for (int q = threadIdx.x; q < NJ * per_link; q += nthreads) {
        const int j = q / per_link, r = q - j * per_link, o = r / COMB, p = r - o * COMB;
        write_half_space(tb, ((rec0 + j) * n_obs + o) * COMB + p, Ob + o * 12 + 3, Lg + j * 18, Ob + o * 12, p);
    }
}

}  // namespace armour
