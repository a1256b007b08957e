// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Minimal stand-in for <boost/numeric/interval.hpp> so that the reference's own sources compile unmodified into
// oracle/_ref.  Boost is not installed in this image; the arithmetic is the restatement in oracle/oracle_interval.hpp
// (rounded_transc_std<double> + save_state: every operation switches the FPU rounding mode and restores it).
#pragma once
#include "../../../oracle_interval.hpp"

namespace boost { namespace numeric {
namespace interval_lib {
template <class T> struct rounded_transc_std {};
template <class Rounding> struct save_state {};
template <class T> struct checking_base {};
template <class Rounding, class Checking> struct policies {};
}  // namespace interval_lib

template <class T, class Policies>
class interval {
    orc::Interval v;
public:
    interval() {}
    interval(const T& x) : v(x) {}
    interval(const T& l, const T& u) : v(l, u) {}
    explicit interval(const orc::Interval& o) : v(o) {}
    const T& lower() const { return v.lo; }
    const T& upper() const { return v.hi; }
    const orc::Interval& raw() const { return v; }
    void assign(const T& l, const T& u) { v = orc::Interval(l, u); }
    interval& operator+=(const interval& o) { v = v + o.v; return *this; }
    interval& operator-=(const interval& o) { v = v - o.v; return *this; }
    interval& operator*=(const interval& o) { v = v * o.v; return *this; }
};
#define ITV interval<T, P>
template <class T, class P> ITV operator+(const ITV& a, const ITV& b) { return ITV(a.raw() + b.raw()); }
template <class T, class P> ITV operator+(const ITV& a, const T& b) { return ITV(a.raw() + b); }
template <class T, class P> ITV operator+(const T& a, const ITV& b) { return ITV(a + b.raw()); }
template <class T, class P> ITV operator-(const ITV& a, const ITV& b) { return ITV(a.raw() - b.raw()); }
template <class T, class P> ITV operator-(const ITV& a, const T& b) { return ITV(a.raw() - b); }
template <class T, class P> ITV operator-(const T& a, const ITV& b) { return ITV(orc::Interval(a) - b.raw()); }
template <class T, class P> ITV operator-(const ITV& a) { return ITV(-a.raw()); }
template <class T, class P> ITV operator*(const ITV& a, const ITV& b) { return ITV(a.raw() * b.raw()); }
template <class T, class P> ITV operator*(const ITV& a, const T& b) { return ITV(b * a.raw()); }
template <class T, class P> ITV operator*(const T& a, const ITV& b) { return ITV(a * b.raw()); }
// division by an interval that does not contain zero (only reached through Eigen's normalize(), which the controller
// path compiles but never calls)
template <class T, class P> ITV operator/(const ITV& a, const ITV& b) {
    if (b.lower() <= 0 && b.upper() >= 0) throw "interval division by an interval containing zero";
    orc::RoundGuard g;
    const T l = std::min(std::min(orc::div_dn(a.lower(), b.lower()), orc::div_dn(a.lower(), b.upper())), std::min(orc::div_dn(a.upper(), b.lower()), orc::div_dn(a.upper(), b.upper())));
    const T u = std::max(std::max(orc::div_up(a.lower(), b.lower()), orc::div_up(a.lower(), b.upper())), std::max(orc::div_up(a.upper(), b.lower()), orc::div_up(a.upper(), b.upper())));
    return ITV(l, u);
}
template <class T, class P> ITV cos(const ITV& a) { return ITV(orc::cos(a.raw())); }
template <class T, class P> ITV sin(const ITV& a) { return ITV(orc::sin(a.raw())); }
template <class T, class P> ITV sqrt(const ITV& a) { return ITV(orc::sqrt(a.raw())); }
template <class T, class P> ITV pow(const ITV& a, int n) {
    if (n != 2) throw "interval pow: only the square is restated";
    return ITV(orc::pow2(a.raw()));
}
template <class T, class P> T lower(const ITV& a) { return a.lower(); }
template <class T, class P> T upper(const ITV& a) { return a.upper(); }
#undef ITV
}}  // namespace boost::numeric
