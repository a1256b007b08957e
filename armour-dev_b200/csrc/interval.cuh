// Directed-rounding interval arithmetic shared by the reach-set kernels and the robust-controller kernels.
#pragma once

namespace armour {

// ---------------------------------------------------------------------------------------------
// Directed-rounding interval arithmetic (replaces boost::numeric::interval with
// rounded_transc_std<double>, KPR/Headers.h:30-36).  cos/sin endpoints are widened outward so the
// device enclosure contains the host libm one whatever the last-bit differences.
// ---------------------------------------------------------------------------------------------
struct Itv { double lo, hi; };
__device__ __forceinline__ Itv itv(double l, double h) { Itv r; r.lo = l; r.hi = h; return r; }
__device__ __forceinline__ Itv iadd(Itv a, Itv b) { return itv(__dadd_rd(a.lo, b.lo), __dadd_ru(a.hi, b.hi)); }
__device__ __forceinline__ Itv iadd(double a, Itv b) { return itv(__dadd_rd(a, b.lo), __dadd_ru(a, b.hi)); }
__device__ __forceinline__ Itv isub(Itv a, Itv b) { return itv(__dadd_rd(a.lo, -b.hi), __dadd_ru(a.hi, -b.lo)); }
__device__ __forceinline__ Itv isub(Itv a, double b) { return itv(__dadd_rd(a.lo, -b), __dadd_ru(a.hi, -b)); }
__device__ __forceinline__ Itv ineg(Itv a) { return itv(-a.hi, -a.lo); }
__device__ __forceinline__ Itv imul(double x, Itv y) {
    if (x < 0) return itv(__dmul_rd(x, y.hi), __dmul_ru(x, y.lo));
    if (x == 0) return itv(0.0, 0.0);
    return itv(__dmul_rd(x, y.lo), __dmul_ru(x, y.hi));
}
// min / max without fmin/fmax's NaN handling (one DSETP + selects); operands here are never NaN
__device__ __forceinline__ double dmin2(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double dmax2(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ Itv imul(Itv x, Itv y) {
    const double l = dmin2(dmin2(__dmul_rd(x.lo, y.lo), __dmul_rd(x.lo, y.hi)), dmin2(__dmul_rd(x.hi, y.lo), __dmul_rd(x.hi, y.hi)));
    const double u = dmax2(dmax2(__dmul_ru(x.lo, y.lo), __dmul_ru(x.lo, y.hi)), dmax2(__dmul_ru(x.hi, y.lo), __dmul_ru(x.hi, y.hi)));
    return itv(l, u);
}
__device__ __forceinline__ Itv ipow2(Itv x) {
    if (x.hi < 0) return itv(__dmul_rd(-x.hi, -x.hi), __dmul_ru(-x.lo, -x.lo));
    if (x.lo < 0) { const double m = fmax(-x.lo, x.hi); return itv(0.0, __dmul_ru(m, m)); }
    return itv(__dmul_rd(x.lo, x.lo), __dmul_ru(x.hi, x.hi));
}
// outward widening of a device libm value: CUDA cos/sin are within 2 ulp, glibc within 1 ulp of the true
// value, and the argument itself may differ from the host's by an ulp; 2^-50 relative + 2^-50 absolute
// covers all three with margin and is ~1e-15, far inside the 1e-9 tolerance on radii.
__device__ __forceinline__ double widen_dn(double v) { return fmax(-1.0, __dadd_rd(__dadd_rd(v, -fabs(v) * 0x1p-50), -0x1p-50)); }
__device__ __forceinline__ double widen_up(double v) { return fmin(1.0, __dadd_ru(__dadd_ru(v, fabs(v) * 0x1p-50), 0x1p-50)); }
__device__ __forceinline__ Itv point_trig(double v) { return itv(widen_dn(v), widen_up(v)); }

#define PI_LO 0x1.921fb54442d18p+1
#define PI_HI 0x1.921fb54442d19p+1
#define PI_HALF_LO 0x1.921fb54442d18p+0
#define PI_HALF_HI 0x1.921fb54442d19p+0
#define PI2_LO 0x1.921fb54442d18p+2
#define PI2_HI 0x1.921fb54442d19p+2

// operator forms (used by the scalar-generic spatial algebra of controller_kernels.cu)
__device__ __forceinline__ Itv operator+(Itv a, Itv b) { return iadd(a, b); }
__device__ __forceinline__ Itv operator-(Itv a, Itv b) { return isub(a, b); }
__device__ __forceinline__ Itv operator-(Itv a) { return ineg(a); }
__device__ __forceinline__ Itv operator*(Itv a, Itv b) { return imul(a, b); }
__device__ __forceinline__ Itv operator*(double a, Itv b) { return imul(a, b); }
__device__ __forceinline__ Itv operator*(Itv a, double b) { return imul(b, a); }

}  // namespace armour
