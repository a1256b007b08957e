// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// Restatement of boost::numeric::interval<double, policies<save_state<rounded_transc_std<double>>,
// checking_base<double>>> for the operations the reference uses (KPR/Headers.h:26-36).  Boost is not installed in
// this image; the formulas follow the published library.  Shared by the oracle (oracle_pz.hpp) and by the stand-in
// Boost header the reference's own sources are compiled against (oracle/shim/boost/numeric/interval.hpp).
#pragma once
#include <algorithm>
#include <cfenv>
#include <cmath>

namespace orc {

// ---------------------------------------------------------------------------------
// Interval with Boost's rounded_transc_std<double> + save_state semantics:
// every operation switches the FPU rounding mode, computes with ordinary arithmetic /
// libm, and restores round-to-nearest.  (KPR/Headers.h:30-36; Boost is not installed
// here, formulas restated from the published library.)
// ---------------------------------------------------------------------------------
#pragma STDC FENV_ACCESS ON
struct RoundGuard { int old; explicit RoundGuard() : old(fegetround()) {} ~RoundGuard() { fesetround(old); } };
// volatile forces the operation to happen under the current rounding mode (-frounding-math is also set)
inline double add_dn(double a, double b) { fesetround(FE_DOWNWARD); volatile double r = a + b; return r; }
inline double add_up(double a, double b) { fesetround(FE_UPWARD);   volatile double r = a + b; return r; }
inline double sub_dn(double a, double b) { fesetround(FE_DOWNWARD); volatile double r = a - b; return r; }
inline double sub_up(double a, double b) { fesetround(FE_UPWARD);   volatile double r = a - b; return r; }
inline double mul_dn(double a, double b) { fesetround(FE_DOWNWARD); volatile double r = a * b; return r; }
inline double mul_up(double a, double b) { fesetround(FE_UPWARD);   volatile double r = a * b; return r; }
inline double div_dn(double a, double b) { fesetround(FE_DOWNWARD); volatile double r = a / b; return r; }
inline double div_up(double a, double b) { fesetround(FE_UPWARD);   volatile double r = a / b; return r; }
inline double cos_dn(double a) { fesetround(FE_DOWNWARD); volatile double r = std::cos(a); return r; }
inline double cos_up(double a) { fesetround(FE_UPWARD);   volatile double r = std::cos(a); return r; }
inline double sqrt_dn(double a) { fesetround(FE_DOWNWARD); volatile double r = std::sqrt(a); return r; }
inline double sqrt_up(double a) { fesetround(FE_UPWARD);   volatile double r = std::sqrt(a); return r; }

struct Interval {
    double lo = 0, hi = 0;
    Interval() {}
    Interval(double v) : lo(v), hi(v) {}
    Interval(double l, double h) : lo(l), hi(h) {}
    double lower() const { return lo; }
    double upper() const { return hi; }
};
static const double PI_LO = 0x1.921fb54442d18p+1;   // Boost constants::pi_lower<double>
static const double PI_HI = 0x1.921fb54442d19p+1;   // pi_upper
static const double PI_HALF_LO = 0x1.921fb54442d18p+0, PI_HALF_HI = 0x1.921fb54442d19p+0;
static const double PI2_LO = 0x1.921fb54442d18p+2, PI2_HI = 0x1.921fb54442d19p+2;

inline Interval operator+(const Interval& a, const Interval& b) { RoundGuard g; return Interval(add_dn(a.lo, b.lo), add_up(a.hi, b.hi)); }
inline Interval operator+(double a, const Interval& b) { RoundGuard g; return Interval(add_dn(a, b.lo), add_up(a, b.hi)); }
inline Interval operator+(const Interval& a, double b) { return b + a; }
inline Interval operator-(const Interval& a, const Interval& b) { RoundGuard g; return Interval(sub_dn(a.lo, b.hi), sub_up(a.hi, b.lo)); }
inline Interval operator-(const Interval& a, double b) { RoundGuard g; return Interval(sub_dn(a.lo, b), sub_up(a.hi, b)); }
inline Interval operator-(const Interval& a) { return Interval(-a.hi, -a.lo); }
inline Interval operator*(double x, const Interval& y) {
    RoundGuard g;
    if (x < 0) return Interval(mul_dn(x, y.hi), mul_up(x, y.lo));
    if (x == 0) return Interval(0.0, 0.0);
    return Interval(mul_dn(x, y.lo), mul_up(x, y.hi));
}
inline Interval operator*(const Interval& y, double x) { return x * y; }
// Boost does a sign-case analysis; as a set this equals the min/max over the four
// directed-rounded endpoint products.
inline Interval operator*(const Interval& x, const Interval& y) {
    RoundGuard g;
    double l = std::min(std::min(mul_dn(x.lo, y.lo), mul_dn(x.lo, y.hi)), std::min(mul_dn(x.hi, y.lo), mul_dn(x.hi, y.hi)));
    double u = std::max(std::max(mul_up(x.lo, y.lo), mul_up(x.lo, y.hi)), std::max(mul_up(x.hi, y.lo), mul_up(x.hi, y.hi)));
    return Interval(l, u);
}
inline Interval& operator+=(Interval& a, const Interval& b) { a = a + b; return a; }
// boost::numeric::pow(interval, int) for the only exponent used (2): binary powering with mul_up / mul_dn
inline Interval pow2(const Interval& x) {
    RoundGuard g;
    if (x.hi < 0) return Interval(mul_dn(-x.hi, -x.hi), mul_up(-x.lo, -x.lo));
    if (x.lo < 0) { double m = std::max(-x.lo, x.hi); return Interval(0.0, mul_up(m, m)); }
    return Interval(mul_dn(x.lo, x.lo), mul_up(x.hi, x.hi));
}
inline Interval fmod_2pi(const Interval& x) {   // interval_lib fmod(x, pi_twice)
    double n;
    { RoundGuard g; const double yb = (x.lo < 0) ? PI2_LO : PI2_HI; n = std::floor(div_dn(x.lo, yb)); }
    return x - n * Interval(PI2_LO, PI2_HI);
}
inline Interval cos(const Interval& x) {
    Interval tmp = fmod_2pi(x);
    double width; { RoundGuard g; width = sub_up(tmp.hi, tmp.lo); }
    if (width >= PI2_LO) return Interval(-1.0, 1.0);
    if (tmp.lo >= PI_HI) return -cos(tmp - Interval(PI_LO, PI_HI));
    const double l = tmp.lo, u = tmp.hi;
    RoundGuard g;
    if (u <= PI_LO) return Interval(cos_dn(u), cos_up(l));
    if (u <= PI2_LO) return Interval(-1.0, cos_up(std::min(sub_dn(PI2_LO, u), l)));
    return Interval(-1.0, 1.0);
}
inline Interval sin(const Interval& x) { return cos(x - Interval(PI_HALF_LO, PI_HALF_HI)); }
inline Interval sqrt(const Interval& x) {
    RoundGuard g;
    double l = (x.lo <= 0) ? 0.0 : sqrt_dn(x.lo);
    return Interval(l, sqrt_up(x.hi));
}
}  // namespace orc
