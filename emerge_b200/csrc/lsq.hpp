// Small dense complex least squares on the host (no CUDA in this header: it is also compiled by the CPU tests,
// tests/hostcheck/lsq_host.cpp).  Used by the reduced-basis projection of recycle.cuh.
#pragma once
#include <cmath>
#include <complex>
#include <vector>

typedef std::complex<double> zc;

// min || g_v - G y_v ||, G: nq x m column-major (destroyed), g: nq x nv column-major (destroyed -> Q^H g).
// Householder QR; columns whose pivot is negligible get y = 0.  resid[v] = || g_v - G y_v ||.
static void ls_solve(int nq, int m, std::vector<zc>& G, int nv, std::vector<zc>& g, std::vector<zc>& y, std::vector<double>& resid) {
    const int kmax = m < nq ? m : nq;
    std::vector<double> piv((size_t)m, 0.0);
    double pmax = 0;
    for (int j = 0; j < kmax; ++j) {
        zc* x = &G[(size_t)j * nq];
        double nx2 = 0;
        for (int i = j; i < nq; ++i) nx2 += std::norm(x[i]);
        const double nx = std::sqrt(nx2);
        if (nx == 0) { piv[j] = 0; continue; }
        const double a0 = std::abs(x[j]);
        const zc ph = a0 > 0 ? x[j] / a0 : zc(1.0, 0.0);
        std::vector<zc> v((size_t)(nq - j));
        for (int i = j; i < nq; ++i) v[(size_t)(i - j)] = x[i];
        v[0] += ph * nx;
        double vn2 = 0;
        for (auto& t : v) vn2 += std::norm(t);
        if (vn2 == 0) { piv[j] = nx; continue; }
        auto reflect = [&](zc* colp) {
            zc w(0.0, 0.0);
            for (int i = j; i < nq; ++i) w += std::conj(v[(size_t)(i - j)]) * colp[i];
            w *= 2.0 / vn2;
            for (int i = j; i < nq; ++i) colp[i] -= w * v[(size_t)(i - j)];
        };
        for (int k = j + 1; k < m; ++k) reflect(&G[(size_t)k * nq]);
        for (int k = 0; k < nv; ++k) reflect(&g[(size_t)k * nq]);
        x[j] = -ph * nx;
        for (int i = j + 1; i < nq; ++i) x[i] = zc(0.0, 0.0);
        piv[j] = nx;
        if (nx > pmax) pmax = nx;
    }
    y.assign((size_t)m * nv, zc(0.0, 0.0));
    resid.assign((size_t)nv, 0.0);
    for (int k = 0; k < nv; ++k) {
        const zc* gk = &g[(size_t)k * nq];
        for (int j = kmax - 1; j >= 0; --j) {
            if (!(piv[j] > 1e-14 * pmax)) { y[(size_t)j * nv + k] = zc(0.0, 0.0); continue; }
            zc s = gk[j];
            for (int l = j + 1; l < kmax; ++l) s -= G[(size_t)l * nq + j] * y[(size_t)l * nv + k];
            y[(size_t)j * nv + k] = s / G[(size_t)j * nq + j];
        }
        double r2 = 0;
        for (int i = kmax; i < nq; ++i) r2 += std::norm(gk[i]);
        for (int j = 0; j < kmax; ++j)
            if (!(piv[j] > 1e-14 * pmax)) r2 += std::norm(gk[j]);
        resid[(size_t)k] = std::sqrt(r2);
    }
}


// Eigen-decomposition of a small complex Hermitian matrix H (n x n, row-major; destroyed) by cyclic Jacobi rotations:
// H = V diag(lam) V^H, V row-major with eigenvectors in its COLUMNS.  n <= a few dozen (Gram matrices of the reduced basis).
static void herm_eig_jacobi(int n, std::vector<zc>& H, std::vector<zc>& V, std::vector<double>& lam) {
    V.assign((size_t)n * n, zc(0.0, 0.0));
    for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = zc(1.0, 0.0);
    auto off2 = [&]() {
        double s = 0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j)
                if (i != j) s += std::norm(H[(size_t)i * n + j]);
        return s;
    };
    double diag2 = 0;
    for (int i = 0; i < n; ++i) diag2 += std::norm(H[(size_t)i * n + i]);
    const double tiny = 1e-30 * (diag2 + off2()) + 1e-300;
    for (int sweep = 0; sweep < 60 && off2() > tiny; ++sweep)
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const zc hpq = H[(size_t)p * n + q];
                const double a = std::abs(hpq);
                if (a == 0.0) continue;
                const double hpp = H[(size_t)p * n + p].real(), hqq = H[(size_t)q * n + q].real();
                const zc ph = hpq / a;                                   // H_pq = a * ph
                const double tau = (hqq - hpp) / (2.0 * a);
                const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
                const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = t * cs;
                // unitary rotation J: columns p, q of H and V;  J = [[cs, sn*ph], [-sn*conj(ph), cs]] acting on (p, q)
                for (int k = 0; k < n; ++k) {                           // H <- H J
                    const zc hkp = H[(size_t)k * n + p], hkq = H[(size_t)k * n + q];
                    H[(size_t)k * n + p] = cs * hkp - sn * std::conj(ph) * hkq;
                    H[(size_t)k * n + q] = sn * ph * hkp + cs * hkq;
                    const zc vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
                    V[(size_t)k * n + p] = cs * vkp - sn * std::conj(ph) * vkq;
                    V[(size_t)k * n + q] = sn * ph * vkp + cs * vkq;
                }
                for (int k = 0; k < n; ++k) {                           // H <- J^H H
                    const zc hpk = H[(size_t)p * n + k], hqk = H[(size_t)q * n + k];
                    H[(size_t)p * n + k] = cs * hpk - sn * ph * hqk;
                    H[(size_t)q * n + k] = sn * std::conj(ph) * hpk + cs * hqk;
                }
            }
    lam.resize((size_t)n);
    for (int i = 0; i < n; ++i) lam[(size_t)i] = H[(size_t)i * n + i].real();
}
