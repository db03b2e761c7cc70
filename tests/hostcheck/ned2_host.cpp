// Host-side instantiation of the product's element-matrix templates (emerge_b200/csrc/ned2_tet.cuh)
// so the closed-form math can be checked against the oracle on a machine without a GPU.
// TEST INFRASTRUCTURE: built by tests/test_host_element.py with g++; never part of the product library.
#include "../../emerge_b200/csrc/ned2_tet.cuh"
#include "../../emerge_b200/csrc/ned2_fused.cuh"
#include <cmath>
#include <utility>

template <int I, int... Js>
static void row(const ned2::TetData& d, cx* K, cx* M, std::integer_sequence<int, Js...>) {
    ((K[I * 20 + Js] = ned2::kentry<I, Js>(d), M[I * 20 + Js] = ned2::mentry<I, Js>(d)), ...);
}
template <int... Is>
static void all_rows(const ned2::TetData& d, cx* K, cx* M, std::integer_sequence<int, Is...>) {
    (row<Is>(d, K, M, std::make_integer_sequence<int, 20>{}), ...);
}

extern "C" void ned2_host_element(const double* p_orig, const long long* vid, const cx* ur, const cx* er,
                                  cx* Kref, cx* Mref) {
    int ord[4] = {0, 1, 2, 3};
    for (int i = 0; i < 4; ++i)
        for (int j = i + 1; j < 4; ++j)
            if (vid[ord[j]] < vid[ord[i]]) std::swap(ord[i], ord[j]);
    double p[4][3];
    for (int k = 0; k < 4; ++k)
        for (int c = 0; c < 3; ++c) p[k][c] = p_orig[ord[k] * 3 + c];
    cx mu[3][3], Ms[3][3], Mm[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { mu[i][j] = ur[i * 3 + j]; Mm[i][j] = er[i * 3 + j]; }
    ned2::matinv_ref(mu, Ms);
    ned2::TetData d;
    ned2::tet_setup(p, Ms, Mm, d);
    cx K[400], M[400];
    all_rows(d, K, M, std::make_integer_sequence<int, 20>{});
    int ref[20];
    ned2::canonical_to_ref(ord, ref);
    for (int i = 0; i < 20; ++i)
        for (int j = 0; j < 20; ++j) {
            Kref[ref[i] * 20 + ref[j]] = K[i * 20 + j];
            Mref[ref[i] * 20 + ref[j]] = M[i * 20 + j];
        }
}

// Same element matrices through the table-driven row form of the fused assembly kernel (ned2_fused.cuh).
static void fused_impl(const double* p_orig, const long long* vid, const cx* ur, const cx* er, cx* Kref, cx* Mref,
                       int pairs) {
    static const ned2f::Tables T{};
    int ord[4] = {0, 1, 2, 3};
    for (int i = 0; i < 4; ++i)
        for (int j = i + 1; j < 4; ++j)
            if (vid[ord[j]] < vid[ord[i]]) std::swap(ord[i], ord[j]);
    double p[4][3];
    for (int k = 0; k < 4; ++k)
        for (int c = 0; c < 3; ++c) p[k][c] = p_orig[ord[k] * 3 + c];
    cx mu[3][3], Ms[3][3], Mm[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { mu[i][j] = ur[i * 3 + j]; Mm[i][j] = er[i * 3 + j]; }
    ned2::matinv_ref(mu, Ms);
    ned2f::TetRec r;
    ned2f::make_record(p, Ms, Mm, r);
    int ref[20];
    ned2::canonical_to_ref(ord, ref);
    for (int i = 0; i < 20; ++i)
        for (int j = 0; j < 20; ++j) {
            cx K, M;
            ned2f::row_entry(T.f[i], T.f[j], T.mc[i][j], r.D, r.g, r.len, K, M);
            Kref[ref[i] * 20 + ref[j]] = K;
            Mref[ref[i] * 20 + ref[j]] = M;
        }
    if (!pairs) return;
    static const ned2f::KernTables KT{};
    // the forms that evaluate both rows of an entity at once (pairs == 2: the flat tables of the kernel)
    for (int i = 0; i < 10; ++i)
        for (int j = 0; j < 20; ++j) {
            cx Ka, Kb, Ma, Mb;
            if (pairs == 2)
                ned2f::row_pair_flat(KT, i, j, r.D, r.g, r.len, Ka, Kb, Ma, Mb);
            else
                ned2f::row_pair_entry(T.pt[i], T.f[i], T.f[i + 10], T.f[j], T.mc[i][j], T.mc[i + 10][j], r.D, r.g, r.len, Ka,
                                      Kb, Ma, Mb);
            Kref[ref[i] * 20 + ref[j]] = Ka;
            Kref[ref[i + 10] * 20 + ref[j]] = Kb;
            Mref[ref[i] * 20 + ref[j]] = Ma;
            Mref[ref[i + 10] * 20 + ref[j]] = Mb;
        }
}
extern "C" void ned2_host_fused(const double* p_orig, const long long* vid, const cx* ur, const cx* er, cx* Kref, cx* Mref) {
    fused_impl(p_orig, vid, ur, er, Kref, Mref, 0);
}
extern "C" void ned2_host_fused_pairs(const double* p_orig, const long long* vid, const cx* ur, const cx* er, cx* Kref,
                                      cx* Mref) {
    fused_impl(p_orig, vid, ur, er, Kref, Mref, 1);
}
extern "C" void ned2_host_fused_flat(const double* p_orig, const long long* vid, const cx* ur, const cx* er, cx* Kref,
                                     cx* Mref) {
    fused_impl(p_orig, vid, ur, er, Kref, Mref, 2);
}
