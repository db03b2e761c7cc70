// Host build of the small least-squares routine used by the reduced-basis projection (emerge_b200/csrc/lsq.hpp);
// checker only (tests/test_host_lsq.py compares it with numpy.linalg.lstsq).
#include "../../emerge_b200/csrc/lsq.hpp"

extern "C" void lsq_solve(int nq, int m, int nv, const double* G_ri, const double* g_ri, double* y_ri, double* resid) {
    std::vector<zc> G((size_t)nq * m), g((size_t)nq * nv), y;
    std::vector<double> res;
    for (size_t i = 0; i < G.size(); ++i) G[i] = zc(G_ri[2 * i], G_ri[2 * i + 1]);
    for (size_t i = 0; i < g.size(); ++i) g[i] = zc(g_ri[2 * i], g_ri[2 * i + 1]);
    ls_solve(nq, m, G, nv, g, y, res);
    for (size_t i = 0; i < y.size(); ++i) { y_ri[2 * i] = y[i].real(); y_ri[2 * i + 1] = y[i].imag(); }
    for (int k = 0; k < nv; ++k) resid[k] = res[(size_t)k];
}

extern "C" void herm_eig(int n, const double* H_ri, double* V_ri, double* lam) {
    std::vector<zc> H((size_t)n * n), V;
    std::vector<double> l;
    for (size_t i = 0; i < H.size(); ++i) H[i] = zc(H_ri[2 * i], H_ri[2 * i + 1]);
    herm_eig_jacobi(n, H, V, l);
    for (size_t i = 0; i < V.size(); ++i) { V_ri[2 * i] = V[i].real(); V_ri[2 * i + 1] = V[i].imag(); }
    for (int i = 0; i < n; ++i) lam[i] = l[(size_t)i];
}
