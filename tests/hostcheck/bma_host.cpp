// Host-side instantiation of the boundary-mode element matrices (emerge_b200/csrc/bma.cuh) so the restated element can be
// checked against the reference's generalized_matrix_GQ on a machine without a GPU.
// TEST INFRASTRUCTURE: built by tests/test_host_bma.py with g++; never part of the product library.
#include "../../emerge_b200/csrc/bma.cuh"

extern "C" void bma_host_element(const double* xy_3x2, const int* lmap_3x2, const cx* ur, const cx* er, double k0, cx* A, cx* B) {
    double p[3][2];
    int lm[3][2];
    cx mu[3][3], ep[3][3];
    for (int k = 0; k < 3; ++k) { p[k][0] = xy_3x2[2 * k]; p[k][1] = xy_3x2[2 * k + 1]; lm[k][0] = lmap_3x2[2 * k]; lm[k][1] = lmap_3x2[2 * k + 1]; }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { mu[i][j] = ur[i * 3 + j]; ep[i][j] = er[i * 3 + j]; }
    bma::TriData d;
    bma::tri_setup(p, lm, mu, ep, d);
    for (int r = 0; r < 14; ++r) bma::element_row(d, k0, r, A + r * 14, B + r * 14);
}
