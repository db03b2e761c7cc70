// Second-order Nedelec tetrahedral element matrices, closed form, canonical-vertex formulation.
//
// Replaces the numba kernel ned2_tet_stiff_mass (reference fem/mth/tet.py:854-1088) and its helpers
// (fem/mth/optimized.py:251-260,343-369,428-461).  This is NOT a transliteration: the reference
// looks up, per tetrahedron, which local vertices form each globally-sorted edge/face
// (local_mapping, fem/mth/optimized.py:273-307) and indexes 5^4 coefficient tables at run time.
// Here the four vertices are first put in ascending global-id order.  In that order every edge is
// (a<b) and every face (a<b<c) of the fixed lists below, so all vertex labels, monomial-integral
// coefficients and signs are compile-time constants and each of the 2x400 entries is straight-line
// FP64 code on a handful of per-element vectors held in registers.  The element matrices are
// invariant under the relabelling (they only depend on even powers of the cofactor vectors).
//
// Functions, canonical index c in [0,20):  [6 edge-a | 4 face-a | 6 edge-b | 4 face-b]
//     N_c = s * l * lambda_X * w_PQ,   w_PQ = lambda_Q grad lambda_P - lambda_P grad lambda_Q
//     edge-a (A,B):   X=A P=A Q=B  l=|AB|      face-a (A,B,E): s=-1 X=B P=A Q=E l=|AE|
//     edge-b (A,B):   X=B P=A Q=B  l=|AB|      face-b (A,B,E):      X=E P=A Q=B l=|AB|
//   M_ij = ss ll/(6V)   [J(XiQiXjQj) g(Pi,Pj) - J(XiQiXjPj) g(Pi,Qj) - J(XiPiXjQj) g(Qi,Pj) + J(XiPiXjPj) g(Qi,Qj)]
//   K_ij = ss ll/(6V)^3 sum_ab k_a k_b I(v_a,v_b) C_a . Ms . C_b,
//          curl N = s l [lam_Q (G_X x G_P) - lam_P (G_X x G_Q) - 2 lam_X (G_P x G_Q)]/(6V)^2
//   g(p,q) = G_p . Mm . G_q,  G = cofactor vectors,  J = prod(m!)/7!,  I = (1+delta)/5!.
// Reference quirks reproduced on purpose:
//   * tet.py:1036 index typo: for face(i) x face-a(j) the coefficient of g(P_i,Q_j) is J(B_i,C_j,D_j,F_j);
//   * tet.py:997-1004: face-edge blocks are the transpose of the edge-face blocks (matters only for
//     non-symmetric tensors);
//   * optimized.py:441-461 matinv: diagonal tensors inverted exactly, full tensors return adj*det.
#pragma once
#include "emb_common.cuh"

namespace ned2 {

EMB_HD constexpr int eA(int e) { return e < 3 ? 0 : (e < 5 ? 1 : 2); }
EMB_HD constexpr int eB(int e) { return e == 0 ? 1 : e == 1 ? 2 : e == 2 ? 3 : e == 3 ? 2 : 3; }
EMB_HD constexpr int fA(int f) { return f == 3 ? 1 : 0; }
EMB_HD constexpr int fB(int f) { return f < 2 ? 1 : 2; }
EMB_HD constexpr int fE(int f) { return f == 0 ? 2 : 3; }
// index of the unordered vertex pair {p,q}: (0,1)=0 (0,2)=1 (0,3)=2 (1,2)=3 (1,3)=4 (2,3)=5
EMB_HD constexpr int pidx(int p, int q) {
    int a = p < q ? p : q, b = p < q ? q : p;
    return a == 0 ? b - 1 : (a == 1 ? b + 1 : 5);
}

struct Fn {
    int X, P, Q;   // N = s*l*lambda_X*w_PQ
    int la, lb;    // l = |v_la - v_lb|
    double s;
    bool face;
    int B;         // second vertex of the entity (used by the reproduced typo)
};
EMB_HD constexpr Fn fn(int c) {
    if (c < 6) return Fn{eA(c), eA(c), eB(c), eA(c), eB(c), 1.0, false, eB(c)};
    if (c < 10) return Fn{fB(c - 6), fA(c - 6), fE(c - 6), fA(c - 6), fE(c - 6), -1.0, true, fB(c - 6)};
    if (c < 16) return Fn{eB(c - 10), eA(c - 10), eB(c - 10), eA(c - 10), eB(c - 10), 1.0, false, eB(c - 10)};
    return Fn{fE(c - 16), fA(c - 16), fB(c - 16), fA(c - 16), fB(c - 16), 1.0, true, fB(c - 16)};
}

EMB_HD constexpr double fact(int n) { return n <= 1 ? 1.0 : n * fact(n - 1); }
EMB_HD constexpr int mult(int v, int a, int b, int c, int d) { return (a == v) + (b == v) + (c == v) + (d == v); }
// int lambda_a lambda_b lambda_c lambda_d dV / (6V)
EMB_HD constexpr double Jc(int a, int b, int c, int d) {
    return fact(mult(0, a, b, c, d)) * fact(mult(1, a, b, c, d)) * fact(mult(2, a, b, c, d)) *
           fact(mult(3, a, b, c, d)) / 5040.0;
}
EMB_HD constexpr double Ic(int a, int b) { return (a == b ? 2.0 : 1.0) / 120.0; }

// one term  kappa * lambda_vert * (sign * X[pair])  of (6V)^2/(s l) * curl N
struct Term {
    bool valid;
    double kappa;
    int vert, pair;
};
EMB_HD constexpr Term mkterm(double k, int vert, int u, int w) {
    return Term{u != w, u < w ? k : -k, vert, pidx(u, w)};
}
EMB_HD constexpr Term term(Fn f, int t) {
    if (!f.face) {   // edge functions collapse to -3 lambda_X (G_P x G_Q)
        return t == 0 ? Term{true, -3.0, f.X, pidx(f.P, f.Q)} : Term{false, 0.0, 0, 0};
    }
    return t == 0 ? mkterm(1.0, f.Q, f.X, f.P) : t == 1 ? mkterm(-1.0, f.P, f.X, f.Q) : mkterm(-2.0, f.X, f.P, f.Q);
}

// Per-element quantities every entry is built from.  The entry templates below are generic over a
// "view" D exposing X(pair,k), Y(pair,k), g(p,q), len(pair), kK(), kM(); TetData is the register/host
// view, the CUDA kernel uses a shared-memory structure-of-arrays view with the same interface.
struct TetData {
    double X_[6][3];   // G_p x G_q per vertex pair
    cx Y_[6][3];       // Ms . X
    cx g_[4][4];       // G_p . Mm . G_q
    double len_[6];    // |v_p - v_q|
    double kK_, kM_;   // 1/(6V)^3, 1/(6V)
    EMB_HD double X(int a, int k) const { return X_[a][k]; }
    EMB_HD cx Y(int a, int k) const { return Y_[a][k]; }
    EMB_HD cx g(int p, int q) const { return g_[p][q]; }
    EMB_HD double len(int a) const { return len_[a]; }
    EMB_HD double kK() const { return kK_; }
    EMB_HD double kM() const { return kM_; }
};

// reference matinv semantics (fem/mth/optimized.py:441-461)
EMB_HD void matinv_ref(const cx s[3][3], cx out[3][3]) {
    bool diag = true;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            if (i != j && (s[i][j].re != 0.0 || s[i][j].im != 0.0)) diag = false;
    if (diag) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) out[i][j] = (i == j) ? cdiv(mk(1.0), s[i][i]) : mk(0.0);
        return;
    }
    cx det = s[0][0] * s[1][1] * s[2][2] - s[0][0] * s[1][2] * s[2][1] - s[0][1] * s[1][0] * s[2][2] +
             s[0][1] * s[1][2] * s[2][0] + s[0][2] * s[1][0] * s[2][1] - s[0][2] * s[1][1] * s[2][0];
    out[0][0] = (s[1][1] * s[2][2] - s[1][2] * s[2][1]) * det;
    out[0][1] = (s[0][2] * s[2][1] - s[0][1] * s[2][2]) * det;
    out[0][2] = (s[0][1] * s[1][2] - s[0][2] * s[1][1]) * det;
    out[1][0] = (s[1][2] * s[2][0] - s[1][0] * s[2][2]) * det;
    out[1][1] = (s[0][0] * s[2][2] - s[0][2] * s[2][0]) * det;
    out[1][2] = (s[0][2] * s[1][0] - s[0][0] * s[1][2]) * det;
    out[2][0] = (s[1][0] * s[2][1] - s[1][1] * s[2][0]) * det;
    out[2][1] = (s[0][1] * s[2][0] - s[0][0] * s[2][1]) * det;
    out[2][2] = (s[0][0] * s[1][1] - s[0][1] * s[1][0]) * det;
}

// p: vertices in ascending global-id order; Ms = matinv(mu_r), Mm = eps_r
EMB_HD void tet_setup(const double p[4][3], const cx Ms[3][3], const cx Mm[3][3], TetData& d) {
    double e1[3], e2[3], e3[3], G[4][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        e1[k] = p[1][k] - p[0][k];
        e2[k] = p[2][k] - p[0][k];
        e3[k] = p[3][k] - p[0][k];
    }
    G[1][0] = e2[1] * e3[2] - e2[2] * e3[1]; G[1][1] = e2[2] * e3[0] - e2[0] * e3[2]; G[1][2] = e2[0] * e3[1] - e2[1] * e3[0];
    G[2][0] = e3[1] * e1[2] - e3[2] * e1[1]; G[2][1] = e3[2] * e1[0] - e3[0] * e1[2]; G[2][2] = e3[0] * e1[1] - e3[1] * e1[0];
    G[3][0] = e1[1] * e2[2] - e1[2] * e2[1]; G[3][1] = e1[2] * e2[0] - e1[0] * e2[2]; G[3][2] = e1[0] * e2[1] - e1[1] * e2[0];
#pragma unroll
    for (int k = 0; k < 3; ++k) G[0][k] = -(G[1][k] + G[2][k] + G[3][k]);
    double det = e1[0] * G[1][0] + e1[1] * G[1][1] + e1[2] * G[1][2];
    double V6 = det < 0 ? -det : det;
    d.kM_ = 1.0 / V6;
    d.kK_ = d.kM_ * d.kM_ * d.kM_;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) {
            const int k = pidx(a, b);
            d.X_[k][0] = G[a][1] * G[b][2] - G[a][2] * G[b][1];
            d.X_[k][1] = G[a][2] * G[b][0] - G[a][0] * G[b][2];
            d.X_[k][2] = G[a][0] * G[b][1] - G[a][1] * G[b][0];
            double dx = p[a][0] - p[b][0], dy = p[a][1] - p[b][1], dz = p[a][2] - p[b][2];
            d.len_[k] = sqrt(dx * dx + dy * dy + dz * dz);
        }
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int r = 0; r < 3; ++r)
            d.Y_[k][r] = d.X_[k][0] * Ms[r][0] + d.X_[k][1] * Ms[r][1] + d.X_[k][2] * Ms[r][2];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        cx H[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) H[r] = G[q][0] * Mm[r][0] + G[q][1] * Mm[r][1] + G[q][2] * Mm[r][2];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) d.g_[pp][q] = G[pp][0] * H[0] + G[pp][1] * H[1] + G[pp][2] * H[2];
    }
}

template <int I, int J, int TI, int TJ, class D>
EMB_HD void kterm(const D& d, cx& acc) {
    constexpr Term a = term(fn(I), TI), b = term(fn(J), TJ);
    if constexpr (a.valid && b.valid) {
        constexpr double coef = a.kappa * b.kappa * Ic(a.vert, b.vert);
        cx z = d.X(a.pair, 0) * d.Y(b.pair, 0) + d.X(a.pair, 1) * d.Y(b.pair, 1) + d.X(a.pair, 2) * d.Y(b.pair, 2);
        fma_r(acc, coef, z);
    }
}

// curl-curl entry as the reference's formula evaluates it for (left=I, right=J)
template <int I, int J, class D>
EMB_HD cx kform(const D& d) {
    constexpr Fn fi = fn(I), fj = fn(J);
    cx acc = mk(0.0);
    kterm<I, J, 0, 0, D>(d, acc); kterm<I, J, 0, 1, D>(d, acc); kterm<I, J, 0, 2, D>(d, acc);
    kterm<I, J, 1, 0, D>(d, acc); kterm<I, J, 1, 1, D>(d, acc); kterm<I, J, 1, 2, D>(d, acc);
    kterm<I, J, 2, 0, D>(d, acc); kterm<I, J, 2, 1, D>(d, acc); kterm<I, J, 2, 2, D>(d, acc);
    const double sc = (fi.s * fj.s) * d.len(pidx(fi.la, fi.lb)) * d.len(pidx(fj.la, fj.lb)) * d.kK();
    return sc * acc;
}

template <int I, int J, class D>
EMB_HD cx mform(const D& d) {
    constexpr Fn fi = fn(I), fj = fn(J);
    constexpr bool typo = fi.face && (J >= 6 && J < 10);
    constexpr double c1 = Jc(fi.X, fi.Q, fj.X, fj.Q);
    constexpr double c2 = typo ? Jc(fi.B, fj.P, fj.X, fj.Q) : Jc(fi.X, fi.Q, fj.X, fj.P);
    constexpr double c3 = Jc(fi.X, fi.P, fj.X, fj.Q);
    constexpr double c4 = Jc(fi.X, fi.P, fj.X, fj.P);
    cx acc = c1 * d.g(fi.P, fj.P);
    fma_r(acc, -c2, d.g(fi.P, fj.Q));
    fma_r(acc, -c3, d.g(fi.Q, fj.P));
    fma_r(acc, c4, d.g(fi.Q, fj.Q));
    const double sc = (fi.s * fj.s) * d.len(pidx(fi.la, fi.lb)) * d.len(pidx(fj.la, fj.lb)) * d.kM();
    return sc * acc;
}

// final entries: face-edge blocks mirror the edge-face blocks (tet.py:997-1004)
template <int I, int J, class D>
EMB_HD cx kentry(const D& d) {
    if constexpr (fn(I).face && !fn(J).face) return kform<J, I, D>(d);
    else return kform<I, J, D>(d);
}
template <int I, int J, class D>
EMB_HD cx mentry(const D& d) {
    if constexpr (fn(I).face && !fn(J).face) return mform<J, I, D>(d);
    else return mform<I, J, D>(d);
}

// ---- canonical <-> reference-local ordering ------------------------------------------------------
// Reference local order (fem/mesh3d.py:292,296): edges (1-2,1-3,1-4,2-3,4-2,3-4), faces (1-2-3,1-3-4,1-4-2,2-3-4).
// ord[k] = original local index (0..3) of the k-th smallest global vertex id.  Returns in ref[c] the
// reference-local slot (0..19) of canonical function c.
EMB_HD void canonical_to_ref(const int ord[4], int ref[20]) {
    // reference local edge index of the unordered pair of original local vertices
    const int edge_of[4][4] = {{-1, 0, 1, 2}, {0, -1, 3, 4}, {1, 3, -1, 5}, {2, 4, 5, -1}};
    // reference local face index by the original local vertex NOT in the face
    const int face_missing[4] = {3, 1, 2, 0};
    for (int e = 0; e < 6; ++e) {
        int r = edge_of[ord[eA(e)]][ord[eB(e)]];
        ref[e] = r;
        ref[10 + e] = 10 + r;
    }
    for (int f = 0; f < 4; ++f) {
        int miss = 3 - f;                       // canonical faces omit sorted position 3,2,1,0
        int r = face_missing[ord[miss]];
        ref[6 + f] = 6 + r;
        ref[16 + f] = 16 + r;
    }
}

}  // namespace ned2
