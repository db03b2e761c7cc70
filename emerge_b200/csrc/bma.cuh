// Element matrices of the port boundary-mode analysis: mixed second-order Nedelec (tangential, 8 functions)
// + quadratic Lagrange (axial, 6 functions) triangle, generalised eigenproblem  A e = -beta^2 B e.
//
// Replaces generalized_matrix_GQ + _matrix_builder (reference fem/physics/edm/nedeleclegrange2.py:223-417), a numba
// prange loop over the port triangles.  Restated from the element definition, not transliterated:
//   barycentric lam_k = a_k + b_k x + c_k y (orientation-corrected, :186-209), edge (i, j) = local vertices of the
//   triangle's global edge in the edge's own direction (local_tri_to_edgeid, :26-29):
//     W_ij   = lam_j grad lam_i - lam_i grad lam_j
//     t-dofs 0-2 / 4-6 : lam_i W_ij / lam_j W_ij  (x edge length),  curl = 3 lam_{i|j} (b_j c_i - b_i c_j)
//     t-dof 3          : -lam_1 W_02 (x |v0 v2|),  t-dof 7: lam_2 W_01 (x |v0 v1|)          (:88-113, :155-187)
//     z-dofs 0-2       : lam_v (2 lam_v - 1),      z-dofs 3-5: 4 lam_i lam_j                  (:117-150)
//   A_tt = (curl N_a, ur^-1_zz curl N_b) - k0^2 (N_a, er_t N_b)
//   B_tt = (N_a, ur^-1_t N_b),  B_zt = (grad L_z, ur^-1_t N_b),  B_tz = B_zt^T (:371),
//   B_zz = (grad L_z, ur^-1_t grad L_z') - k0^2 (L_z, er_zz L_z')
// integrated with the reference's 6-point rule INCLUDING its 8-digit constants (:213-216; the rule is not exact for these
// degree-4 integrands, so the constants are part of the result).  Kept on purpose: the face-row / edge-column entries of
// the three tangential blocks are copies of the transposed entries (:320-346) and B_tz is the transpose of B_zt, which
// differs from the direct formula for non-symmetric tensors; matinv's adj*det for full tensors (ned2::matinv_ref); the
// face-face entries of the tangential block of B are zero (the reference never assigns them).
#pragma once
#include "ned2_tet.cuh"

namespace bma {

struct TriData {
    double lam[3][6];        // barycentric coordinates at the quadrature points
    double gx[3], gy[3];     // their gradients
    double w[6];             // weights
    double scale[14];        // row / column scaling (edge lengths for the tangential functions, 1 for the axial ones)
    double area;
    int ei[3], ej[3];        // local vertices of the three edges, in the direction of the global edge
    cx Ms[2][2], Mm[2][2], Msz, Mmz;
};

// xy: triangle vertices in port-local coordinates; lmap[e] = {local index of the edge's first global vertex, second}
EMB_HD void tri_setup(const double xy[3][2], const int lmap[3][2], const cx ur[3][3], const cx er[3][3], TriData& d) {
    const double W0 = 0.22338159, W1 = 0.10995174;
    const double P[3][6] = {{0.10810302, 0.44594849, 0.44594849, 0.81684757, 0.09157621, 0.09157621},
                            {0.44594849, 0.44594849, 0.10810302, 0.09157621, 0.09157621, 0.81684757},
                            {0.44594849, 0.10810302, 0.44594849, 0.09157621, 0.81684757, 0.09157621}};
    const double x1 = xy[0][0], x2 = xy[1][0], x3 = xy[2][0], y1 = xy[0][1], y2 = xy[1][1], y3 = xy[2][1];
    const double sA = 0.5 * ((x1 - x3) * (y2 - y1) - (x1 - x2) * (y3 - y1));
    const double sg = sA > 0 ? 1.0 : (sA < 0 ? -1.0 : 0.0);
    const double A = fabs(sA);
    const double i2A = sg / (2.0 * A);
    const double a[3] = {(x2 * y3 - y2 * x3) * i2A, (x3 * y1 - y3 * x1) * i2A, (x1 * y2 - y1 * x2) * i2A};
    d.gx[0] = (y2 - y3) * i2A; d.gx[1] = (y3 - y1) * i2A; d.gx[2] = (y1 - y2) * i2A;
    d.gy[0] = (x3 - x2) * i2A; d.gy[1] = (x1 - x3) * i2A; d.gy[2] = (x2 - x1) * i2A;
    d.area = A;
    for (int q = 0; q < 6; ++q) {
        const double xq = x1 * P[0][q] + x2 * P[1][q] + x3 * P[2][q];
        const double yq = y1 * P[0][q] + y2 * P[1][q] + y3 * P[2][q];
        d.w[q] = q < 3 ? W0 : W1;
        for (int k = 0; k < 3; ++k) d.lam[k][q] = a[k] + d.gx[k] * xq + d.gy[k] * yq;
    }
    auto dist = [&](int p, int q) {
        const double dx = xy[p][0] - xy[q][0], dy = xy[p][1] - xy[q][1];
        return sqrt(dx * dx + dy * dy);
    };
    for (int e = 0; e < 3; ++e) {
        d.ei[e] = lmap[e][0]; d.ej[e] = lmap[e][1];
        d.scale[e] = d.scale[e + 4] = dist(lmap[e][0], lmap[e][1]);
    }
    d.scale[3] = dist(0, 2);
    d.scale[7] = dist(0, 1);
    for (int z = 8; z < 14; ++z) d.scale[z] = 1.0;
    cx inv[3][3];
    ned2::matinv_ref(ur, inv);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) { d.Ms[i][j] = inv[i][j]; d.Mm[i][j] = er[i][j]; }
    d.Msz = inv[2][2];
    d.Mmz = er[2][2];
}

// tangential function a (0..7) at quadrature point q: value (nx, ny) and scalar curl
EMB_HD void tfun(const TriData& d, int a, int q, double& nx, double& ny, double& curl) {
    int i, j, m;             // N = sgn * lam_m * (lam_j grad lam_i - lam_i grad lam_j)
    double sgn = 1.0;
    if (a == 3) { i = 0; j = 2; m = 1; sgn = -1.0; }
    else if (a == 7) { i = 0; j = 1; m = 2; }
    else { const int e = a & 3; i = d.ei[e]; j = d.ej[e]; m = a < 4 ? i : j; }
    const double wx = d.lam[j][q] * d.gx[i] - d.lam[i][q] * d.gx[j];
    const double wy = d.lam[j][q] * d.gy[i] - d.lam[i][q] * d.gy[j];
    const double lm = d.lam[m][q];
    nx = sgn * lm * wx;
    ny = sgn * lm * wy;
    // curl(lam_m W) = grad lam_m x W + lam_m curl W,   curl W = -2 (grad lam_i x grad lam_j)
    const double cij = d.gx[i] * d.gy[j] - d.gy[i] * d.gx[j];
    curl = sgn * ((d.gx[m] * wy - d.gy[m] * wx) - 2.0 * lm * cij);
}
// axial function z (0..5): value and gradient
EMB_HD void zfun(const TriData& d, int z, int q, double& v, double& gx, double& gy) {
    if (z < 3) {
        const double l = d.lam[z][q];
        v = -l + 2.0 * l * l;
        gx = d.gx[z] * (4.0 * l - 1.0);
        gy = d.gy[z] * (4.0 * l - 1.0);
    } else {
        const int i = d.ei[z - 3], j = d.ej[z - 3];
        v = 4.0 * d.lam[i][q] * d.lam[j][q];
        gx = 4.0 * (d.gx[i] * d.lam[j][q] + d.gx[j] * d.lam[i][q]);
        gy = 4.0 * (d.gy[i] * d.lam[j][q] + d.gy[j] * d.lam[i][q]);
    }
}
// u^T T v for real 2-vectors and a complex 2x2 tensor
EMB_HD cx bil(const cx T[2][2], double ux, double uy, double vx, double vy) {
    return (ux * vx) * T[0][0] + (ux * vy) * T[0][1] + (uy * vx) * T[1][0] + (uy * vy) * T[1][1];
}

// row r (0..13: 8 tangential, 6 axial) of the 14 x 14 element matrices A and B
EMB_HD void element_row(const TriData& d, double k0, int r, cx* __restrict__ Arow, cx* __restrict__ Brow) {
    const double k2 = k0 * k0;
    for (int c = 0; c < 14; ++c) {
        cx av = mk(0.0), bv = mk(0.0);
        if (r < 8 && c < 8) {
            // face row against edge column: the reference stores the transposed entry (nedeleclegrange2.py:320-346)
            const bool mirror = ((r & 3) == 3) && ((c & 3) != 3);
            const int L = mirror ? c : r, R = mirror ? r : c;
            double scc = 0.0;
            cx sm = mk(0.0), ss = mk(0.0);
            for (int q = 0; q < 6; ++q) {
                double lx, ly, lc, rx, ry, rc;
                tfun(d, L, q, lx, ly, lc);
                tfun(d, R, q, rx, ry, rc);
                scc += d.w[q] * lc * rc;
                sm += d.w[q] * bil(d.Mm, lx, ly, rx, ry);
                ss += d.w[q] * bil(d.Ms, lx, ly, rx, ry);
            }
            av = scc * d.Msz - k2 * sm;
            // the reference never fills the face-face entries of the tangential block of B (D_tt[3|7][3|7] stay zero:
            // nedeleclegrange2.py:349-357 set them for A_tt and B_tt only) - reproduced
            bv = ((r & 3) == 3 && (c & 3) == 3) ? mk(0.0) : ss;
        } else if (r >= 8 && c >= 8) {
            cx s1 = mk(0.0);
            double s2 = 0.0;
            for (int q = 0; q < 6; ++q) {
                double lv, lgx, lgy, rv, rgx, rgy;
                zfun(d, r - 8, q, lv, lgx, lgy);
                zfun(d, c - 8, q, rv, rgx, rgy);
                s1 += d.w[q] * bil(d.Ms, lgx, lgy, rgx, rgy);
                s2 += d.w[q] * lv * rv;
            }
            bv = s1 - (k2 * s2) * d.Mmz;
        } else {
            // B_zt[z][b] = (grad L_z, Ms N_b); the tz block is its transpose (nedeleclegrange2.py:370-371)
            const int z = (r >= 8 ? r : c) - 8, b = r >= 8 ? c : r;
            cx s = mk(0.0);
            for (int q = 0; q < 6; ++q) {
                double v, gx, gy, nx, ny, cc;
                zfun(d, z, q, v, gx, gy);
                tfun(d, b, q, nx, ny, cc);
                s += d.w[q] * bil(d.Ms, gx, gy, nx, ny);
            }
            bv = s;
        }
        const double sc = d.scale[r] * d.scale[c] * d.area;
        Arow[c] = sc * av;
        Brow[c] = sc * bv;
    }
}

}  // namespace bma
