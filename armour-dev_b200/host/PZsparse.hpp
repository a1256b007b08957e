// PZsparse — host facade with the reference's public surface (KPR/PZsparse.h:50-210) whose arithmetic runs on
// the device through armour_pz_binary (include/armour_b200.h).  Coefficients are stored flat, column-major
// inside a monomial like Eigen; keys are the packed 63-bit degree words of KPR/PZsparse.h:23-40.
// Supported shapes are the ones the hot path uses: 1x1, 3x1, 3x3 (products 3x3*3x1, 3x3*3x3, 1x1*1x1;
// + and - for 1x1 and 3x1; cross for 3x1).  Errors from the C ABI are thrown as int, the reference's own
// convention (KPR/PZsparse.cu:248, KPR/armour_main.cu:99-108).
//
// Where each public operation of the reference runs:
//   device (armour_pz_binary): operator* + - (PZ, PZ), the three cross() overloads, simplify(), addOneDimPZ(), stack()
//   host (pure data movement or one pass over the list, written out like the reference): transpose(), operator()(row, col),
//   reduce(), reduce_link_PZ(), slice() value and the three gradient overloads, toInterval(), the scalar operators
//   (* / + - with a double, unary minus, double op PZ) and the degree pack / unpack helpers.
// Reference quirks kept on purpose: unary minus drops `independent` (KPR/PZsparse.cu:725-741); `a - PZ` returns PZ - a
// (:849-862); the gradient slices test `degree <= 2^14` where the value slice tests `<` (:452 vs :413).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/armour_b200.h"

struct Monomial {
    std::vector<double> coeff;   // rows*cols, column-major
    uint64_t degree = 0;
};

class PZsparse {
public:
    unsigned NRows = 0, NCols = 0;
    std::vector<double> center, independent;
    std::vector<Monomial> polynomial;

    static armour_handle*& device() { static armour_handle* h = nullptr; return h; }   // bind once: PZsparse::device() = handle

    PZsparse() {}
    PZsparse(unsigned r, unsigned c) : NRows(r), NCols(c), center(r * c, 0.0), independent(r * c, 0.0) {}
    explicit PZsparse(double c) : NRows(1), NCols(1), center(1, c), independent(1, 0.0) {}

    // reach-set tables of a built handle: which as in armour_get_pz (0 cos_q_des ... 9 u_nom_int)
    static PZsparse from_table(armour_handle* h, int which, int idx, int t) {
        int dims[2] = {0, 0};
        std::vector<uint64_t> keys(256);
        std::vector<double> co(256 * 9), ce(9), in(9);
        const int n = armour_get_pz(h, which, idx, t, dims, keys.data(), co.data(), ce.data(), in.data());
        if (n < 0) throw n;
        return unflatten(dims, n, keys.data(), co.data(), ce.data(), in.data());
    }

    // ---- device-backed ---------------------------------------------------------------------------------------------
    void simplify() { *this = binary(4, *this, PZsparse(0.0)); }                                   // KPR/PZsparse.cu:284-350
    void addOneDimPZ(const PZsparse& a, unsigned row_id, unsigned col_id) {                        // :1068-1085
        if (NRows != 3 || NCols != 1 || col_id != 0 || row_id > 2 || a.NRows != 1 || a.NCols != 1) throw -1;
        *this = binary(7 + (int)row_id, *this, a);
    }
    friend PZsparse cross(const double (&a)[3], const PZsparse& b) { return binary(10, b, constant3(a)); }   // :1118-1132
    friend PZsparse cross(const PZsparse& a, const double (&b)[3]) { return binary(11, a, constant3(b)); }   // :1153-1167
    // stack N = 3 one-dimensional PZs (:1087-1116): concatenation on the host, simplify() on the device
    friend PZsparse stack(const PZsparse (&a)[3]) {
        PZsparse r(3, 1);
        for (int i = 0; i < 3; i++) {
            if (a[i].NRows != 1 || a[i].NCols != 1) throw -1;
            r.center[i] = a[i].center[0]; r.independent[i] = a[i].independent[0];
            for (const Monomial& m : a[i].polynomial) { Monomial t; t.degree = m.degree; t.coeff.assign(3, 0.0); t.coeff[i] = m.coeff[0]; r.polynomial.push_back(t); }
        }
        r.simplify();
        return r;
    }

    // ---- host: one pass over the monomial list, as the reference writes them ----------------------------------------
    PZsparse operator()(int row_id, int col_id) const {   // :678-697 — every monomial is copied, zero coefficients included, no simplify
        PZsparse r(1, 1);
        const size_t e = (size_t)row_id + (size_t)col_id * NRows;
        r.center[0] = center[e]; r.independent[0] = independent[e];
        for (const Monomial& m : polynomial) { Monomial t; t.degree = m.degree; t.coeff.assign(1, m.coeff[e]); r.polynomial.push_back(t); }
        return r;
    }
    void reduce() {   // :352-368
        std::vector<Monomial> keep;
        for (const Monomial& m : polynomial) {
            if (m.degree < (1ull << 14)) keep.push_back(m);
            else for (size_t e = 0; e < independent.size(); e++) independent[e] += std::fabs(m.coeff[e]);
        }
        polynomial.swap(keep);
    }
    std::vector<double> reduce_link_PZ() {   // :370-402 — returns the 3x6 block [g_x g_y g_z | diag(independent)], column-major
        if (NRows != 3 || NCols != 1) throw -1;
        std::vector<double> gens(18, 0.0);
        std::vector<Monomial> keep;
        int j = 0;
        for (const Monomial& m : polynomial) {
            if (m.degree < (1ull << 14)) keep.push_back(m);
            else if (m.degree < (1ull << 35) && (m.degree & ((1ull << 14) - 1)) == 0) { if (j >= 3) throw -5; for (int a = 0; a < 3; a++) gens[a + 3 * j] = m.coeff[a]; j++; }
            else for (int e = 0; e < 3; e++) independent[e] += std::fabs(m.coeff[e]);
        }
        polynomial.swap(keep);
        for (int a = 0; a < 3; a++) gens[a + 3 * (3 + a)] = independent[a];
        return gens;
    }
    // gradient of the slice with respect to k_0..k_6 (:437-555): gradient[k] has rows*cols entries; the three overloads of the
    // reference (MatrixXd*, Vector3d*, double*) differ only in the container
    void slice_gradient(std::vector<double> (&gradient)[7], const double* factor) const {
        for (int k = 0; k < 7; k++) gradient[k].assign(center.size(), 0.0);
        for (const Monomial& m : polynomial) {
            if (m.degree > (1ull << 14)) continue;   // sic: "<=" in the reference
            for (int k = 0; k < 7; k++) {
                double mon = 1.0;
                bool zero = false;
                for (int j = 0; j < 7; j++) {
                    const unsigned d = (unsigned)((m.degree >> (2 * j)) & 3);
                    if (j == k) { if (d == 0) zero = true; else mon *= (double)d * std::pow(factor[j], (double)(d - 1)); }
                    else mon *= std::pow(factor[j], (double)d);
                }
                if (!zero) for (size_t e = 0; e < center.size(); e++) gradient[k][e] += m.coeff[e] * mon;
            }
        }
    }
    void slice(double* gradient, const double* factor) const {   // 1-dim PZ (:515-555)
        if (NRows != 1 || NCols != 1) throw -1;
        std::vector<double> g[7];
        slice_gradient(g, factor);
        for (int k = 0; k < 7; k++) gradient[k] = g[k][0];
    }
    void toInterval(std::vector<double>& lower, std::vector<double>& upper) const {   // :557-576
        lower = center; upper = center;
        for (size_t e = 0; e < center.size(); e++) {
            double r = independent[e];
            for (const Monomial& m : polynomial) r += std::fabs(m.coeff[e]);
            lower[e] = center[e] - r; upper[e] = center[e] + r;
        }
    }
    static void convertHashToDegree(uint64_t degree, uint64_t (&degreeArray)[42]) {   // :578-585; variable order of KPR/PZsparse.h:23-40
        for (int j = 0; j < 7; j++) degreeArray[j] = (degree >> (2 * j)) & 3;
        for (int j = 0; j < 21; j++) degreeArray[7 + j] = (degree >> (14 + j)) & 1;
        for (int j = 0; j < 14; j++) degreeArray[28 + j] = (degree >> (35 + 2 * j)) & 3;
    }
    static uint64_t convertDegreeToHash(const uint64_t (&degreeArray)[42]) {          // :587-603
        uint64_t h = 0;
        for (int j = 0; j < 7; j++) h |= degreeArray[j] << (2 * j);
        for (int j = 0; j < 21; j++) h |= degreeArray[7 + j] << (14 + j);
        for (int j = 0; j < 14; j++) h |= degreeArray[28 + j] << (35 + 2 * j);
        return h;
    }
    PZsparse operator-() const {   // :725-741 — `independent` is NOT carried over (the reference comments that line out)
        PZsparse r(NRows, NCols);
        for (size_t e = 0; e < center.size(); e++) r.center[e] = -center[e];
        for (const Monomial& m : polynomial) { Monomial t = m; for (double& v : t.coeff) v = -v; r.polynomial.push_back(t); }
        return r;
    }
    PZsparse operator*(double a) const {   // :996-1012 — no simplify
        PZsparse r = *this;
        for (double& v : r.center) v *= a;
        for (Monomial& m : r.polynomial) for (double& v : m.coeff) v = a * v;
        for (double& v : r.independent) v *= std::fabs(a);
        return r;
    }
    friend PZsparse operator*(double a, const PZsparse& b) { return b * a; }   // :1014-1030
    PZsparse operator/(double a) const {   // :1032-1048
        PZsparse r = *this;
        for (double& v : r.center) v /= a;
        for (Monomial& m : r.polynomial) for (double& v : m.coeff) v /= a;
        for (double& v : r.independent) v /= std::fabs(a);
        return r;
    }
    PZsparse operator+(double a) const { PZsparse r = *this; for (double& v : r.center) v += a; return r; }   // :788-811
    friend PZsparse operator+(double a, const PZsparse& b) { return b + a; }
    PZsparse operator-(double a) const { PZsparse r = *this; for (double& v : r.center) v -= a; return r; }   // :836-847
    friend PZsparse operator-(double a, const PZsparse& b) { return b - a; }   // sic (:849-862): the reference returns b - a
    PZsparse& operator+=(const PZsparse& a) { *this = *this + a; return *this; }   // :813-834

    PZsparse operator+(const PZsparse& a) const { return binary(1, *this, a); }
    PZsparse operator-(const PZsparse& a) const { return binary(2, *this, a); }
    PZsparse operator*(const PZsparse& a) const { return binary(0, *this, a); }
    friend PZsparse cross(const PZsparse& a, const PZsparse& b) { return binary(3, a, b); }

    PZsparse transpose() const {   // KPR/PZsparse.cu:1050-1066 (pure data movement, host)
        PZsparse r(NCols, NRows);
        auto tr = [&](const std::vector<double>& v) { std::vector<double> o(v.size()); for (unsigned i = 0; i < NRows; i++) for (unsigned j = 0; j < NCols; j++) o[j + i * NCols] = v[i + j * NRows]; return o; };
        r.center = tr(center); r.independent = tr(independent);
        for (const Monomial& m : polynomial) { Monomial t; t.degree = m.degree; t.coeff = tr(m.coeff); r.polynomial.push_back(t); }
        return r;
    }
    // slice at k (KPR/PZsparse.cu:404-435): centre and radius of the resulting interval matrix
    void slice(const double* factor, std::vector<double>& c, std::vector<double>& r) const {
        c = center; r = independent;
        for (const Monomial& m : polynomial) {
            if (m.degree < (1ull << 14)) {
                double mon = 1.0;
                for (int j = 0; j < 7; j++) mon *= std::pow(factor[j], (double)((m.degree >> (2 * j)) & 3));
                for (size_t e = 0; e < c.size(); e++) c[e] += m.coeff[e] * mon;
            }
            else for (size_t e = 0; e < r.size(); e++) r[e] += std::fabs(m.coeff[e]);
        }
    }

private:
    static PZsparse constant3(const double (&v)[3]) { PZsparse c(3, 1); for (int i = 0; i < 3; i++) c.center[i] = v[i]; return c; }
    static PZsparse unflatten(const int* dims, int n, const uint64_t* keys, const double* co, const double* ce, const double* in) {
        PZsparse r(dims[0], dims[1]);
        const int d = dims[0] * dims[1];
        for (int e = 0; e < d; e++) { r.center[e] = ce[e]; r.independent[e] = in[e]; }
        r.polynomial.resize(n);
        for (int i = 0; i < n; i++) { r.polynomial[i].degree = keys[i]; r.polynomial[i].coeff.assign(co + (size_t)i * d, co + (size_t)(i + 1) * d); }
        return r;
    }
    static PZsparse binary(int op, const PZsparse& a, const PZsparse& b) {
        armour_handle* h = device();
        if (!h) throw -1;
        auto flat = [](const PZsparse& z, std::vector<uint64_t>& k, std::vector<double>& c) {
            const size_t d = (size_t)z.NRows * z.NCols;
            k.resize(z.polynomial.size()); c.resize(z.polynomial.size() * d);
            for (size_t i = 0; i < z.polynomial.size(); i++) { k[i] = z.polynomial[i].degree; for (size_t e = 0; e < d; e++) c[i * d + e] = z.polynomial[i].coeff[e]; }
        };
        std::vector<uint64_t> ka, kb;
        std::vector<double> ca, cb;
        flat(a, ka, ca); flat(b, kb, cb);
        const int cap = (int)(ka.size() + kb.size() + ka.size() * kb.size()) + 1;
        std::vector<uint64_t> kr(cap);
        std::vector<double> cr((size_t)cap * 9), ce(9), in(9);
        int dims[2];
        const int n = armour_pz_binary(h, op, a.NRows, a.NCols, (int)ka.size(), ka.data(), ca.data(), a.center.data(), a.independent.data(),
                                       b.NRows, b.NCols, (int)kb.size(), kb.data(), cb.data(), b.center.data(), b.independent.data(),
                                       cap, dims, kr.data(), cr.data(), ce.data(), in.data());
        if (n < 0) throw n;
        return unflatten(dims, n, kr.data(), cr.data(), ce.data(), in.data());
    }
};
