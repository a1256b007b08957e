// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the reference's sparse polynomial-zonotope algebra
// (KPR = kinova_src/kinova_simulator_interfaces/kinova_planner_realtime):
//   KPR/PZsparse.h:23-83      monomial key layout + storage
//   KPR/PZsparse.cu:10-40     getCenter / getRadius
//   KPR/PZsparse.cu:50-205    constructors
//   KPR/PZsparse.cu:284-350   simplify
//   KPR/PZsparse.cu:352-402   reduce / reduce_link_PZ
//   KPR/PZsparse.cu:404-555   slice (value + gradient)
//   KPR/PZsparse.cu:557-603   toInterval, key pack/unpack
//   KPR/PZsparse.cu:678-1167  arithmetic, transpose, addOneDimPZ, stack, cross
//   KPR/Headers.h:26-36       Boost.Interval policy (rounded_transc_std + save_state)
//
// PARITY: pinned against the reference's own sources compiled here (oracle/_ref, see oracle/README.md and
// tests/test_reference_pin.py: keys bit-exact, values <= 2e-15 at T = 128).  Boost.Interval and Eigen 3.3.7 are absent
// from this image; their arithmetic is restated (here and in oracle/shim), which is the one thing _ref cannot pin.
//
// Design notes (kept deliberately close to the reference so that it is an honest
// CPU baseline): one heap-allocated coefficient matrix per monomial, range-for by
// value, sort-then-merge simplify, asserts enabled.  Differences, all result-neutral:
//   * std::stable_sort instead of std::sort (the reference's equal-key summation
//     order is whatever libstdc++'s introsort leaves; stable order is canonical).
//   * sizes (time steps, uncertainty, threshold, k_range) are runtime values.
#pragma once
#include <algorithm>
#include <cassert>
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "oracle_interval.hpp"

namespace orc {

constexpr int NJ = 7;   // NUM_JOINTS  (KPR/KinovaWithoutGripperInfo.h:10)
constexpr int NF = 7;   // NUM_FACTORS (KPR/KinovaWithoutGripperInfo.h:14)
constexpr int NVAR = NF * 6;

// ---------------------------------------------------------------------------------
// op counters (algorithmic work is counted on the reference op sequence)
// ---------------------------------------------------------------------------------
struct OpStats {
    uint64_t n_mul = 0;            // PZ x PZ multiplies
    uint64_t pair_products = 0;    // (n_a+1)(n_b+1) summed over multiplies
    uint64_t flops = 0;            // sum (n_a+1)(n_b+1)*phi
    uint64_t n_simplify = 0;
    uint64_t simplify_in = 0;      // monomials entering simplify
    uint64_t simplify_out = 0;     // monomials surviving
    uint64_t max_simplify_in = 0;
    uint64_t max_simplify_out = 0;
    void add(const OpStats& o) {
        n_mul += o.n_mul; pair_products += o.pair_products; flops += o.flops;
        n_simplify += o.n_simplify; simplify_in += o.simplify_in; simplify_out += o.simplify_out;
        max_simplify_in = std::max(max_simplify_in, o.max_simplify_in);
        max_simplify_out = std::max(max_simplify_out, o.max_simplify_out);
    }
};
inline OpStats*& tls_stats() { static thread_local OpStats* p = nullptr; return p; }

// optional per-op trace (used once, offline, to size the GPU engine's buffers)
struct OpTrace { int kind; int dim; int na; int nb; int nin; int nout; int nza = 0; int nzb = 0; };
inline std::vector<OpTrace>*& tls_trace() { static thread_local std::vector<OpTrace>* p = nullptr; return p; }

// ---------------------------------------------------------------------------------
// Small dynamic column-major matrix standing in for Eigen::MatrixXd (heap per object)
// ---------------------------------------------------------------------------------
struct Mat {
    int r = 0, c = 0;
    std::vector<double> v;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), v((size_t)r_ * c_, 0.0) {}
    static Mat Zero(int r, int c) { return Mat(r, c); }
    static Mat Identity(int r, int c) { Mat m(r, c); for (int i = 0; i < std::min(r, c); i++) m(i, i) = 1.0; return m; }
    double& operator()(int i, int j) { return v[(size_t)i + (size_t)j * r]; }
    double operator()(int i, int j) const { return v[(size_t)i + (size_t)j * r]; }
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
    int size() const { return r * c; }
    Mat cwiseAbs() const { Mat m(r, c); for (int i = 0; i < size(); i++) m.v[i] = std::fabs(v[i]); return m; }
    Mat transpose() const { Mat m(c, r); for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) m(j, i) = (*this)(i, j); return m; }
    Mat block(int i, int j, int nr, int nc) const { Mat m(nr, nc); for (int a = 0; a < nr; a++) for (int b = 0; b < nc; b++) m(a, b) = (*this)(i + a, j + b); return m; }
    void setZero() { std::fill(v.begin(), v.end(), 0.0); }
    // Eigen 3.3 squaredNorm() on a dynamic matrix with SSE2 packets of two doubles:
    // two packet accumulators over aligned groups of four, packet sum, horizontal sum, then the
    // scalar tail.  For sizes 1 and 3 this equals the plain left-to-right sum.
    double squaredNorm() const {
        const int n = size();
        const int aligned = (n / 2) * 2, aligned2 = (n / 4) * 4;
        if (aligned == 0) { double s = v[0] * v[0]; for (int i = 1; i < n; i++) s = s + v[i] * v[i]; return s; }
        double p0a = v[0] * v[0], p0b = v[1] * v[1];
        if (aligned > 2) {
            double p1a = v[2] * v[2], p1b = v[3] * v[3];
            for (int i = 4; i < aligned2; i += 4) {
                p0a = p0a + v[i] * v[i];         p0b = p0b + v[i + 1] * v[i + 1];
                p1a = p1a + v[i + 2] * v[i + 2]; p1b = p1b + v[i + 3] * v[i + 3];
            }
            p0a = p0a + p1a; p0b = p0b + p1b;
            if (aligned > aligned2) { p0a = p0a + v[aligned2] * v[aligned2]; p0b = p0b + v[aligned2 + 1] * v[aligned2 + 1]; }
        }
        double s = p0a + p0b;
        for (int i = aligned; i < n; i++) s = s + v[i] * v[i];
        return s;
    }
    double norm() const { return std::sqrt(squaredNorm()); }
};
inline Mat operator+(const Mat& a, const Mat& b) { assert(a.r == b.r && a.c == b.c); Mat m(a.r, a.c); for (int i = 0; i < a.size(); i++) m.v[i] = a.v[i] + b.v[i]; return m; }
inline Mat operator-(const Mat& a, const Mat& b) { assert(a.r == b.r && a.c == b.c); Mat m(a.r, a.c); for (int i = 0; i < a.size(); i++) m.v[i] = a.v[i] - b.v[i]; return m; }
inline Mat operator-(const Mat& a) { Mat m(a.r, a.c); for (int i = 0; i < a.size(); i++) m.v[i] = -a.v[i]; return m; }
inline Mat operator*(const Mat& a, double s) { Mat m(a.r, a.c); for (int i = 0; i < a.size(); i++) m.v[i] = a.v[i] * s; return m; }
inline Mat operator*(double s, const Mat& a) { Mat m(a.r, a.c); for (int i = 0; i < a.size(); i++) m.v[i] = s * a.v[i]; return m; }
inline Mat operator/(const Mat& a, double s) { Mat m(a.r, a.c); for (int i = 0; i < a.size(); i++) m.v[i] = a.v[i] / s; return m; }
inline Mat& operator+=(Mat& a, const Mat& b) { assert(a.r == b.r && a.c == b.c); for (int i = 0; i < a.size(); i++) a.v[i] += b.v[i]; return a; }
// Eigen's coefficient-based small product: res(i,j) = sum_k a(i,k) b(k,j), k ascending, no FMA.
inline Mat matmul(const Mat& a, const Mat& b) {
    assert(a.c == b.r);
    Mat m(a.r, b.c);
    for (int j = 0; j < b.c; j++)
        for (int i = 0; i < a.r; i++) {
            double s = a(i, 0) * b(0, j);
            for (int k = 1; k < a.c; k++) s = s + a(i, k) * b(k, j);
            m(i, j) = s;
        }
    return m;
}

inline double getCenter(const Interval& a) { return (a.lo + a.hi) * 0.5; }   // KPR/PZsparse.cu:10-12
inline double getRadius(const Interval& a) { return (a.hi - a.lo) * 0.5; }   // KPR/PZsparse.cu:14-16

// ---------------------------------------------------------------------------------
// Monomial key (KPR/PZsparse.h:23-40): 2 bits per k_j, 1 bit per qde/qdae/qddae_j,
// 2 bits per cosqe_j / sinqe_j, little end first.
// ---------------------------------------------------------------------------------
static const uint64_t MOVE_BIT_INC[NVAR] = {2,2,2,2,2,2,2, 1,1,1,1,1,1,1, 1,1,1,1,1,1,1, 1,1,1,1,1,1,1, 2,2,2,2,2,2,2, 2,2,2,2,2,2,2};
static const uint64_t DEGREE_MASK[NVAR]  = {3,3,3,3,3,3,3, 1,1,1,1,1,1,1, 1,1,1,1,1,1,1, 1,1,1,1,1,1,1, 3,3,3,3,3,3,3, 3,3,3,3,3,3,3};
static const uint64_t KEY_K_ONLY = (uint64_t)1 << (2 * NF);          // max_hash_dependent_k_only
static const uint64_t KEY_K_LINKS_ONLY = (uint64_t)1 << (5 * NF);    // max_hash_dependent_k_links_only
static const uint64_t KEY_K_MASK = KEY_K_ONLY - 1;

inline uint64_t convertDegreeToHash(const uint64_t* deg) {   // KPR/PZsparse.cu:587-603
    uint64_t key = 0, shift = 0;
    for (int i = 0; i < NVAR; i++) {
        if (deg[i] > 1) { fprintf(stderr, "degree can not be larger than 1!\n"); throw -1; }
        key += deg[i] << shift;
        shift += MOVE_BIT_INC[i];
    }
    return key;
}
inline void convertHashToDegree(uint64_t key, uint64_t* deg) {   // KPR/PZsparse.cu:578-585
    for (int i = 0; i < NVAR; i++) { deg[i] = key & DEGREE_MASK[i]; key >>= MOVE_BIT_INC[i]; }
}

struct Monomial {
    Mat coeff;
    uint64_t degree = 0;
    Monomial(const Mat& c, uint64_t d) : coeff(c), degree(d) {}
    Monomial(double c, uint64_t d) : coeff(1, 1), degree(d) { coeff(0) = c; }
};

struct PZ {
    unsigned NRows = 0, NCols = 0;
    Mat center = Mat(1, 1);
    std::vector<Monomial> polynomial;
    Mat independent = Mat(1, 1);
    static double& threshold() { static double t = 5e-4; return t; }   // SIMPLIFY_THRESHOLD, KPR/Parameters.h:10

    PZ() {}
    PZ(unsigned r, unsigned c) : NRows(r), NCols(c), center(r, c), independent(r, c) {}
    explicit PZ(double c) : NRows(1), NCols(1), center(1, 1), independent(1, 1) { center(0) = c; }
    explicit PZ(const Mat& c) : NRows(c.r), NCols(c.c), center(c), independent(c.r, c.c) {}
    PZ(const Mat& c, double uncertainty) : NRows(c.r), NCols(c.c), center(c), independent(uncertainty * c.cwiseAbs()) {}   // :93-98
    // 1x1 with monomials (KPR/PZsparse.cu:120-136)
    PZ(double c, const double* coeff, const uint64_t (*deg)[NVAR], unsigned n) : NRows(1), NCols(1), center(1, 1), independent(1, 1) {
        center(0) = c;
        polynomial.reserve(n);
        for (unsigned i = 0; i < n; i++) polynomial.emplace_back(coeff[i], convertDegreeToHash(deg[i]));
        simplify();
    }
    // 3x3 from roll/pitch/yaw (KPR/PZsparse.cu:160-176)
    PZ(double roll, double pitch, double yaw) : NRows(3), NCols(3), center(3, 3), independent(3, 3) {
        center(0,0) = std::cos(pitch)*std::cos(yaw);
        center(0,1) = -std::cos(pitch)*std::sin(yaw);
        center(0,2) = std::sin(pitch);
        center(1,0) = std::cos(roll)*std::sin(yaw) + std::cos(yaw)*std::sin(pitch)*std::sin(roll);
        center(1,1) = std::cos(roll)*std::cos(yaw) - std::sin(pitch)*std::sin(roll)*std::sin(yaw);
        center(1,2) = -std::cos(pitch)*std::sin(roll);
        center(2,0) = std::sin(roll)*std::sin(yaw) - std::cos(roll)*std::cos(yaw)*std::sin(pitch);
        center(2,1) = std::cos(yaw)*std::sin(roll) + std::cos(roll)*std::sin(pitch)*std::sin(yaw);
        center(2,2) = std::cos(pitch)*std::cos(roll);
    }
    // 3x3 joint rotation from cos/sin PZs (KPR/PZsparse.cu:179-205)
    PZ(double cc, const double* ccoeff, const uint64_t (*cdeg)[NVAR], unsigned cn,
       double sc, const double* scoeff, const uint64_t (*sdeg)[NVAR], unsigned sn, unsigned axis)
        : NRows(3), NCols(3), independent(3, 3) {
        makeRotationMatrix(center, cc, sc, axis, false);
        polynomial.reserve(cn + sn);
        Mat tmp;
        for (unsigned i = 0; i < cn; i++) { makeRotationMatrix(tmp, ccoeff[i], 0, axis, true); polynomial.emplace_back(tmp, convertDegreeToHash(cdeg[i])); }
        for (unsigned i = 0; i < sn; i++) { makeRotationMatrix(tmp, 0, scoeff[i], axis, true); polynomial.emplace_back(tmp, convertDegreeToHash(sdeg[i])); }
        simplify();
    }

    static void makeRotationMatrix(Mat& R, double c, double s, unsigned axis, bool fromZero) {   // :211-250
        R = fromZero ? Mat::Zero(3, 3) : Mat::Identity(3, 3);
        const double ns = -1.0 * s;
        switch (axis) {
            case 0: return;
            case 1: R(1,1) = c; R(1,2) = ns; R(2,1) = s; R(2,2) = c; break;
            case 2: R(0,0) = c; R(0,2) = s; R(2,0) = ns; R(2,2) = c; break;
            case 3: R(0,0) = c; R(0,1) = ns; R(1,0) = s; R(1,1) = c; break;
            default: fprintf(stderr, "Undefined axis\n"); throw -1;
        }
    }

    bool internalCheck() const {   // :252-282
        if (center.r != (int)NRows || center.c != (int)NCols) return false;
        if (independent.r != (int)NRows || independent.c != (int)NCols) return false;
        for (int i = 0; i < independent.size(); i++) if (independent.v[i] < 0) return false;
        return true;
    }

    void simplify() {   // :284-350
        assert(internalCheck());
        std::stable_sort(polynomial.begin(), polynomial.end(), [](const Monomial& a, const Monomial& b) { return a.degree < b.degree; });
        Mat reduce_amount(NRows, NCols);
        std::vector<Monomial> polynomial_new;
        polynomial_new.reserve(polynomial.size());
        const size_t nin = polynomial.size();
        size_t i = 0;
        while (i < polynomial.size()) {
            size_t j;
            const uint64_t degree = polynomial[i].degree;
            for (j = i + 1; j < polynomial.size(); j++) {
                if (polynomial[j].degree != degree) break;
                polynomial[i].coeff += polynomial[j].coeff;
            }
            Mat temp = polynomial[i].coeff;
            if (temp.norm() <= threshold()) reduce_amount += temp.cwiseAbs();
            else polynomial_new.emplace_back(polynomial[i]);
            i = j;
        }
        polynomial = polynomial_new;
        if (reduce_amount.norm() != 0) independent = independent + reduce_amount;
        if (OpStats* s = tls_stats()) {
            s->n_simplify++; s->simplify_in += nin; s->simplify_out += polynomial.size();
            s->max_simplify_in = std::max<uint64_t>(s->max_simplify_in, nin);
            s->max_simplify_out = std::max<uint64_t>(s->max_simplify_out, polynomial.size());
        }
        if (auto* t = tls_trace()) t->push_back({0, (int)(NRows * NCols), 0, 0, (int)nin, (int)polynomial.size()});
    }

    void reduce() {   // :352-368
        assert(internalCheck());
        std::vector<Monomial> polynomial_new;
        polynomial_new.reserve(polynomial.size());
        for (auto it : polynomial) {
            if (it.degree < KEY_K_ONLY) polynomial_new.emplace_back(it.coeff, it.degree);
            else independent += it.coeff.cwiseAbs();
        }
        polynomial = polynomial_new;
    }

    Mat reduce_link_PZ() {   // :370-402
        assert(internalCheck());
        assert(NRows == 3 && NCols == 1);
        Mat gens(3, 6);
        std::vector<Monomial> polynomial_new;
        polynomial_new.reserve(polynomial.size());
        int j = 0;
        for (auto it : polynomial) {
            if (it.degree < KEY_K_ONLY) polynomial_new.emplace_back(it.coeff, it.degree);
            else if (it.degree < KEY_K_LINKS_ONLY && (it.degree & KEY_K_MASK) == 0) {
                assert(j < 3);
                for (int a = 0; a < 3; a++) gens(a, j) = it.coeff(a);
                j++;
            }
            else independent += it.coeff.cwiseAbs();
        }
        polynomial = polynomial_new;
        gens(0, 3) = independent(0); gens(1, 4) = independent(1); gens(2, 5) = independent(2);
        return gens;
    }

    // value slice (:404-435): centre and radius of the sliced interval matrix
    void slice(const double* factor, Mat& res_center, Mat& res_radius) const {
        assert(internalCheck());
        res_center = center; res_radius = independent;
        uint64_t deg[NVAR];
        for (auto it : polynomial) {
            Mat resTemp = it.coeff;
            if (it.degree < KEY_K_ONLY) {
                convertHashToDegree(it.degree, deg);
                for (int j = 0; j < NF; j++) resTemp = resTemp * std::pow(factor[j], (double)deg[j]);
                res_center += resTemp;
            }
            else res_radius += resTemp.cwiseAbs();
        }
    }
    // getCenter(slice(x)): Interval(c-r, c+r) then (lo+hi)/2  (KPR/NLPclass.cu:306-313)
    Mat sliceCenter(const double* factor) const {
        Mat c, r; slice(factor, c, r);
        Mat out(c.r, c.c);
        for (int i = 0; i < c.size(); i++) out.v[i] = getCenter(Interval(c.v[i] - r.v[i], c.v[i] + r.v[i]));
        return out;
    }
    // gradient slice (:437-555; the three overloads compute the same thing for their shapes)
    void sliceGradient(Mat* gradient, const double* factor) const {
        assert(internalCheck());
        for (int k = 0; k < NF; k++) gradient[k] = Mat::Zero(NRows, NCols);
        uint64_t deg[NVAR];
        Mat resTemp[NF];
        for (auto it : polynomial) {
            if (it.degree <= KEY_K_ONLY) {   // sic: "<=" in the reference
                for (int k = 0; k < NF; k++) resTemp[k] = it.coeff;
                convertHashToDegree(it.degree, deg);
                for (int j = 0; j < NF; j++)
                    for (int k = 0; k < NF; k++) {
                        if (j == k) {
                            if (deg[j] == 0) resTemp[k] = Mat::Zero(NRows, NCols);
                            else resTemp[k] = resTemp[k] * ((double)deg[j] * std::pow(factor[j], (double)(deg[j] - 1)));
                        }
                        else resTemp[k] = resTemp[k] * std::pow(factor[j], (double)deg[j]);
                    }
                for (int k = 0; k < NF; k++) gradient[k] += resTemp[k];
            }
        }
    }
    // toInterval (:557-576): centre and radius
    void toInterval(Mat& c, Mat& r) const {
        c = center; r = independent;
        for (auto it : polynomial) r += it.coeff.cwiseAbs();
    }

    PZ operator()(int row, int col) const {   // :678-697
        assert(internalCheck());
        PZ res(1, 1);
        res.center = center.block(row, col, 1, 1);
        res.polynomial.reserve(polynomial.size());
        for (auto it : polynomial) res.polynomial.emplace_back(it.coeff.block(row, col, 1, 1), it.degree);
        res.independent = independent.block(row, col, 1, 1);
        return res;
    }
    PZ operator+(const PZ& a) const {   // :743-764
        assert(internalCheck());
        PZ res(NRows, NCols);
        res.center = center + a.center;
        res.polynomial.reserve(polynomial.size() + a.polynomial.size());
        res.polynomial.insert(res.polynomial.end(), polynomial.begin(), polynomial.end());
        for (auto it : a.polynomial) res.polynomial.push_back(it);
        res.independent = independent + a.independent;
        res.simplify();
        return res;
    }
    PZ& operator+=(const PZ& a) {   // :794-811
        center += a.center;
        polynomial.reserve(polynomial.size() + a.polynomial.size());
        for (auto it : a.polynomial) polynomial.push_back(it);
        independent += a.independent;
        simplify();
        return *this;
    }
    PZ operator-(const PZ& a) const {   // :813-834
        assert(internalCheck());
        PZ res(NRows, NCols);
        res.center = center - a.center;
        res.polynomial.reserve(polynomial.size() + a.polynomial.size());
        res.polynomial.insert(res.polynomial.end(), polynomial.begin(), polynomial.end());
        for (auto it : a.polynomial) res.polynomial.emplace_back(-it.coeff, it.degree);
        res.independent = independent + a.independent;
        res.simplify();
        return res;
    }
    PZ operator*(const PZ& a) const {   // :864-994
        assert(internalCheck());
        const bool ls = (NRows == 1 && NCols == 1), rs = (a.NRows == 1 && a.NCols == 1);
        assert(NCols == a.NRows || ls || rs);
        PZ res;
        if (ls) { res.NRows = a.NRows; res.NCols = a.NCols; }
        else if (rs) { res.NRows = NRows; res.NCols = NCols; }
        else { res.NRows = NRows; res.NCols = a.NCols; }
        auto mulc = [&](const Mat& x, const Mat& y) -> Mat {
            if (ls) return x(0) * y;
            if (rs) return x * y(0);
            return matmul(x, y);
        };
        res.center = mulc(center, a.center);
        res.polynomial.reserve(polynomial.size() + a.polynomial.size() + polynomial.size() * a.polynomial.size());
        for (auto it : polynomial) res.polynomial.emplace_back(mulc(it.coeff, a.center), it.degree);
        for (auto it : a.polynomial) res.polynomial.emplace_back(mulc(center, it.coeff), it.degree);
        for (auto it1 : polynomial)
            for (auto it2 : a.polynomial)
                res.polynomial.emplace_back(mulc(it1.coeff, it2.coeff), it1.degree + it2.degree);   // key add, no carry check
        Mat ra2 = center.cwiseAbs();
        for (auto it : polynomial) ra2 += it.coeff.cwiseAbs();
        ra2 = mulc(ra2, a.independent);
        Mat ra3 = a.center.cwiseAbs();
        for (auto it : a.polynomial) ra3 += it.coeff.cwiseAbs();
        ra3 = mulc(independent, ra3);
        Mat ra = ra2 + ra3;
        res.independent = mulc(independent, a.independent) + ra;
        if (OpStats* s = tls_stats()) {
            const uint64_t pp = (uint64_t)(polynomial.size() + 1) * (a.polynomial.size() + 1);
            const uint64_t phi = (ls || rs) ? (uint64_t)res.NRows * res.NCols : 2ull * NRows * NCols * a.NCols;
            s->n_mul++; s->pair_products += pp; s->flops += pp * phi;
        }
        if (auto* t = tls_trace()) {
            int nza = 0, nzb = 0;
            for (auto& m : polynomial) { bool nz = false; for (double x : m.coeff.v) nz |= (x != 0.0); nza += nz; }
            for (auto& m : a.polynomial) { bool nz = false; for (double x : m.coeff.v) nz |= (x != 0.0); nzb += nz; }
            t->push_back({1, (int)(res.NRows * res.NCols), (int)polynomial.size(), (int)a.polynomial.size(), 0, 0, nza, nzb});
        }
        res.simplify();
        return res;
    }
    PZ operator*(double a) const {   // :996-1012
        PZ res(NRows, NCols);
        res.center = center * a;
        res.polynomial.reserve(polynomial.size());
        for (auto it : polynomial) res.polynomial.emplace_back(a * it.coeff, it.degree);
        res.independent = independent * std::fabs(a);
        return res;
    }
    PZ transpose() const {   // :1050-1066
        PZ res(NCols, NRows);
        res.center = center.transpose();
        res.polynomial.reserve(polynomial.size());
        for (auto it : polynomial) res.polynomial.emplace_back(it.coeff.transpose(), it.degree);
        res.independent = independent.transpose();
        return res;
    }
    void addOneDimPZ(const PZ& a, unsigned row, unsigned col) {   // :1068-1085
        assert(internalCheck());
        assert(a.NRows == 1 && a.NCols == 1 && row < NRows && col < NCols);
        center(row, col) += a.center(0);
        for (auto it : a.polynomial) {
            Mat t = Mat::Zero(NRows, NCols);
            t(row, col) = it.coeff(0);
            polynomial.emplace_back(t, it.degree);
        }
        independent(row, col) += a.independent(0);
        simplify();
    }
};
inline PZ operator*(double a, const PZ& b) {   // :1014-1030
    PZ res(b.NRows, b.NCols);
    res.center = b.center * a;
    res.polynomial.reserve(b.polynomial.size());
    for (auto it : b.polynomial) res.polynomial.emplace_back(a * it.coeff, it.degree);
    res.independent = b.independent * std::fabs(a);
    return res;
}
inline PZ stack3(const PZ* a) {   // stack(), :1087-1116, for the only size used (3)
    PZ res(3, 1);
    for (int i = 0; i < 3; i++) res.center(i, 0) = a[i].center(0);
    res.polynomial.reserve(3 * a[0].polynomial.size());
    for (int i = 0; i < 3; i++)
        for (auto it : a[i].polynomial) {
            Mat t = Mat::Zero(3, 1);
            t(i) = it.coeff(0);
            res.polynomial.emplace_back(t, it.degree);
        }
    for (int i = 0; i < 3; i++) res.independent(i, 0) = a[i].independent(0);
    res.simplify();
    return res;
}
inline PZ cross(const Mat& a, const PZ& b) {   // :1118-1132
    PZ r[3]; PZ b0 = b(0, 0), b1 = b(1, 0), b2 = b(2, 0);
    r[0] = a(1, 0) * b2 - a(2, 0) * b1;
    r[1] = a(2, 0) * b0 - a(0, 0) * b2;
    r[2] = a(0, 0) * b1 - a(1, 0) * b0;
    return stack3(r);
}
inline PZ cross(const PZ& a, const PZ& b) {   // :1134-1151
    PZ r[3]; PZ a0 = a(0, 0), a1 = a(1, 0), a2 = a(2, 0), b0 = b(0, 0), b1 = b(1, 0), b2 = b(2, 0);
    r[0] = a1 * b2 - a2 * b1;
    r[1] = a2 * b0 - a0 * b2;
    r[2] = a0 * b1 - a1 * b0;
    return stack3(r);
}
inline PZ cross(const PZ& a, const Mat& b) {   // :1153-1167
    PZ r[3]; PZ a0 = a(0, 0), a1 = a(1, 0), a2 = a(2, 0);
    r[0] = a1 * b(2, 0) - a2 * b(1, 0);
    r[1] = a2 * b(0, 0) - a0 * b(2, 0);
    r[2] = a0 * b(1, 0) - a1 * b(0, 0);
    return stack3(r);
}

}  // namespace orc
