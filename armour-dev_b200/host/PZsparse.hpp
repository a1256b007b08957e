// PZsparse — host facade with the reference's public surface (KPR/PZsparse.h:50-210) whose arithmetic runs on
// the device through armour_pz_binary (include/armour_b200.h).  Coefficients are stored flat, column-major
// inside a monomial like Eigen; keys are the packed 63-bit degree words of KPR/PZsparse.h:23-40.
// Supported shapes are the ones the hot path uses: 1x1, 3x1, 3x3 (products 3x3*3x1, 3x3*3x3, 1x1*1x1;
// + and - for 1x1 and 3x1; cross for 3x1).  Errors from the C ABI are thrown as int, the reference's own
// convention (KPR/PZsparse.cu:248, KPR/armour_main.cu:99-108).
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
