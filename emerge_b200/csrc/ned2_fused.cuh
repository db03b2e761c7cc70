// Table-driven form of the second-order Nedelec element matrices for the FUSED assembly kernel (assembly.cu,
// k_asm_rows): one matrix ROW (the row function I is only known at run time, the column function J is fixed per lane)
// is evaluated from a small per-tetrahedron record instead of 2x400 compile-time entries.
//
// Same mathematics and the same reference quirks as ned2_tet.cuh (reference fem/mth/tet.py:854-1088); everything is
// derived from ned2::fn / ned2::term / ned2::Jc at compile time, so there is one source of truth.  The host check
// tests/hostcheck/ned2_host.cpp compares row_entry() with ned2::kentry / ned2::mentry entry by entry.
//
// Per-tetrahedron record (896 B, written once by k_tet_records, read by the ~10 entities of the tetrahedron):
//   D[pa*6+pb] = (X_pa . Ms . X_pb) / (6V)^3     X_p = G_a x G_b of vertex pair p,  36 complex
//   g[p*4+q]   = (G_p . Mm . G_q) / (6V)                                           16 complex
//   len[6]     = edge lengths of the vertex pairs
// With  K_IJ = s_I s_J l_I l_J / 120 * sum_{ta,tb} kappa_ta kappa_tb (1 + [v_ta == v_tb]) D[pair_ta][pair_tb]
//       M_IJ = s_I s_J l_I l_J * (c1 g(Pl,Pr) - c2 g(Pl,Qr) - c3 g(Ql,Pr) + c4 g(Ql,Qr))
// where for a face row against an edge column the pair (left, right) is swapped (tet.py:997-1004: the face-edge blocks
// are filled as transposes of the edge-face blocks) and c2 carries the index typo of tet.py:1036.
#pragma once
#include "ned2_tet.cuh"

namespace ned2f {

struct TetRec {
    cx D[36];
    cx g[16];
    double len[6];
    double pad[2];
};
static_assert(sizeof(TetRec) == 896, "record layout");

struct FnTab {
    signed char P, Q, lp, face, nt;
    signed char tv[3], tp[3];
    double s;        // sign of the function
    double tk[3];    // kappa of the curl terms (sign of the pair orientation folded in)
};

// The two functions of one entity (rows ic and ic+10) use the same vertex pairs in their curl terms (an edge: its own
// pair; a face: its three edges), in different order and with different vertices / factors.  PairTab lists the pairs
// once so that a D entry is loaded once for both rows.
struct PairTab {
    signed char np, tp[3], va[3], vb[3];
    double ka[3], kb[3];
};

struct Tables {
    FnTab f[20];
    PairTab pt[10];
    double mc[20][20][4];   // {c1, -c2, -c3, c4} of row I, column J (mirror and typo resolved)
    constexpr Tables() : f{}, pt{}, mc{} {
        for (int c = 0; c < 20; ++c) {
            const ned2::Fn fc = ned2::fn(c);
            f[c].P = (signed char)fc.P;
            f[c].Q = (signed char)fc.Q;
            f[c].lp = (signed char)ned2::pidx(fc.la, fc.lb);
            f[c].face = fc.face ? 1 : 0;
            f[c].s = fc.s;
            int nt = 0;
            for (int t = 0; t < 3; ++t) {
                const ned2::Term tm = ned2::term(fc, t);
                f[c].tk[t] = 0.0;
                f[c].tv[t] = 0;
                f[c].tp[t] = 0;
                if (tm.valid) {
                    f[c].tk[nt] = tm.kappa;
                    f[c].tv[nt] = (signed char)tm.vert;
                    f[c].tp[nt] = (signed char)tm.pair;
                    ++nt;
                }
            }
            f[c].nt = (signed char)nt;
        }
        for (int c = 0; c < 10; ++c) {
            const FnTab &a = f[c], &b = f[c + 10];
            pt[c].np = a.nt;
            for (int t = 0; t < 3; ++t) {
                pt[c].tp[t] = 0; pt[c].va[t] = 0; pt[c].vb[t] = 0;
                pt[c].ka[t] = 0.0; pt[c].kb[t] = 0.0;
            }
            for (int t = 0; t < a.nt; ++t) {
                pt[c].tp[t] = a.tp[t];
                pt[c].va[t] = a.tv[t];
                pt[c].ka[t] = a.tk[t];
                for (int u = 0; u < b.nt; ++u)
                    if (b.tp[u] == a.tp[t]) {
                        pt[c].vb[t] = b.tv[u];
                        pt[c].kb[t] = b.tk[u];
                    }
            }
        }
        for (int I = 0; I < 20; ++I)
            for (int J = 0; J < 20; ++J) {
                const bool mirror = ned2::fn(I).face && !ned2::fn(J).face;
                const int L = mirror ? J : I, R = mirror ? I : J;
                const ned2::Fn fl = ned2::fn(L), fr = ned2::fn(R);
                const bool typo = fl.face && (R >= 6 && R < 10);
                mc[I][J][0] = ned2::Jc(fl.X, fl.Q, fr.X, fr.Q);
                mc[I][J][1] = -(typo ? ned2::Jc(fl.B, fr.P, fr.X, fr.Q) : ned2::Jc(fl.X, fl.Q, fr.X, fr.P));
                mc[I][J][2] = -ned2::Jc(fl.X, fl.P, fr.X, fr.Q);
                mc[I][J][3] = ned2::Jc(fl.X, fl.P, fr.X, fr.P);
            }
    }
};

// p: vertices in ascending global-id order; Ms = matinv(mu_r), Mm = eps_r
EMB_HD void make_record(const double p[4][3], const cx Ms[3][3], const cx Mm[3][3], TetRec& r) {
    ned2::TetData d;
    ned2::tet_setup(p, Ms, Mm, d);
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            cx z = d.X_[a][0] * d.Y_[b][0] + d.X_[a][1] * d.Y_[b][1] + d.X_[a][2] * d.Y_[b][2];
            r.D[a * 6 + b] = d.kK_ * z;
        }
#pragma unroll
    for (int pq = 0; pq < 16; ++pq) r.g[pq] = d.kM_ * d.g_[pq / 4][pq % 4];
#pragma unroll
    for (int k = 0; k < 6; ++k) r.len[k] = d.len_[k];
    r.pad[0] = r.pad[1] = 0.0;
}

// entries (I, J) of the curl-curl and mass element matrices from the record; mc = Tables::mc[I][J]
EMB_HD void row_entry(const FnTab& fi, const FnTab& fj, const double* mc, const cx* D, const cx* g, const double* len,
                      cx& K, cx& M) {
    const bool mirror = fi.face && !fj.face;
    cx acc = mk(0.0);
    for (int ta = 0; ta < fi.nt; ++ta)
        for (int tb = 0; tb < fj.nt; ++tb) {
            const double coef = fi.tk[ta] * fj.tk[tb] * (fi.tv[ta] == fj.tv[tb] ? 2.0 : 1.0);
            const int idx = mirror ? fj.tp[tb] * 6 + fi.tp[ta] : fi.tp[ta] * 6 + fj.tp[tb];
            fma_r(acc, coef, D[idx]);
        }
    const double sc = fi.s * fj.s * len[fi.lp] * len[fj.lp];
    K = (sc * (1.0 / 120.0)) * acc;
    const int lP = mirror ? fj.P : fi.P, lQ = mirror ? fj.Q : fi.Q;
    const int rP = mirror ? fi.P : fj.P, rQ = mirror ? fi.Q : fj.Q;
    cx m = mc[0] * g[lP * 4 + rP];
    fma_r(m, mc[1], g[lP * 4 + rQ]);
    fma_r(m, mc[2], g[lQ * 4 + rP]);
    fma_r(m, mc[3], g[lQ * 4 + rQ]);
    M = sc * m;
}

// the same for the two rows (ic, ic + 10) of one entity at once: what a lane of the fused kernel evaluates per tetrahedron
EMB_HD void row_pair_entry(const PairTab& pt, const FnTab& fa, const FnTab& fb, const FnTab& fj, const double* mca,
                           const double* mcb, const cx* D, const cx* g, const double* len, cx& Ka, cx& Kb, cx& Ma,
                           cx& Mb) {
    const bool mirror = fa.face && !fj.face;
    cx acca = mk(0.0), accb = mk(0.0);
    // fully unrolled with guards: the column function's tables stay in registers (compile-time indices)
#pragma unroll
    for (int ta = 0; ta < 3; ++ta) {
        if (ta < pt.np) {
            const int pa = pt.tp[ta], va = pt.va[ta], vb = pt.vb[ta];
            const double ka = pt.ka[ta], kb = pt.kb[ta];
#pragma unroll
            for (int tb = 0; tb < 3; ++tb) {
                if (tb < fj.nt) {
                    const int idx = mirror ? fj.tp[tb] * 6 + pa : pa * 6 + fj.tp[tb];
                    const cx d = D[idx];
                    const double kj = fj.tk[tb];
                    const int vj = fj.tv[tb];
                    fma_r(acca, ka * kj * (va == vj ? 2.0 : 1.0), d);
                    fma_r(accb, kb * kj * (vb == vj ? 2.0 : 1.0), d);
                }
            }
        }
    }
    const double lj = fj.s * len[fj.lp];
    const double sca = fa.s * len[fa.lp] * lj, scb = fb.s * len[fb.lp] * lj;
    Ka = (sca * (1.0 / 120.0)) * acca;
    Kb = (scb * (1.0 / 120.0)) * accb;
    {
        const int lP = mirror ? fj.P : fa.P, lQ = mirror ? fj.Q : fa.Q;
        const int rP = mirror ? fa.P : fj.P, rQ = mirror ? fa.Q : fj.Q;
        cx m = mca[0] * g[lP * 4 + rP];
        fma_r(m, mca[1], g[lP * 4 + rQ]);
        fma_r(m, mca[2], g[lQ * 4 + rP]);
        fma_r(m, mca[3], g[lQ * 4 + rQ]);
        Ma = sca * m;
    }
    {
        const int lP = mirror ? fj.P : fb.P, lQ = mirror ? fj.Q : fb.Q;
        const int rP = mirror ? fb.P : fj.P, rQ = mirror ? fb.Q : fj.Q;
        cx m = mcb[0] * g[lP * 4 + rP];
        fma_r(m, mcb[1], g[lP * 4 + rQ]);
        fma_r(m, mcb[2], g[lQ * 4 + rP]);
        fma_r(m, mcb[3], g[lQ * 4 + rQ]);
        Mb = scb * m;
    }
}

// ---- flat tables of the kernel ---------------------------------------------------------------------------------
// Everything that depends only on (entity slot ic, column function J) is tabulated, in a layout where the 20 column
// lanes of a warp read consecutive 16-byte words (conflict-free shared-memory loads), and padded to a uniform term
// count per entity type so the warp runs without divergence:
//   kc[ic][k][J]  = (ca, cb): coefficient of D[kidx(ic,J,k)] in rows ic and ic+10, signs and 1/120 folded in;
//                   k = ta*3 + tb < nk[ic] (3 for an edge, 9 for a face); edge columns have zero entries for tb > 0
//   kidx[ic][J]   = the nk D-indices, 6 bits each
//   mc[ic][r][h][J] = coefficients (c1, -c2 | -c3, c4) of row r (0: ic, 1: ic+10), signs folded in
//   gidx[ic][J]   = the 2 x 4 g-indices, 4 bits each
struct KernTables {
    double kc[10][9][20][2];
    double mc[10][2][2][20][2];
    unsigned long long kidx[10][20];
    unsigned int gidx[10][20];
    int nk[10];
    int lp[20];
    constexpr KernTables() : kc{}, mc{}, kidx{}, gidx{}, nk{}, lp{} {
        const Tables T{};
        for (int J = 0; J < 20; ++J) lp[J] = T.f[J].lp;
        for (int ic = 0; ic < 10; ++ic) {
            const FnTab &fa = T.f[ic], &fb = T.f[ic + 10];
            const PairTab& pt = T.pt[ic];
            nk[ic] = pt.np * 3;
            for (int J = 0; J < 20; ++J) {
                const FnTab& fj = T.f[J];
                const bool mirror = fa.face && !fj.face;
                unsigned long long packed = 0;
                for (int ta = 0; ta < pt.np; ++ta)
                    for (int tb = 0; tb < 3; ++tb) {
                        const int k = ta * 3 + tb;
                        double ca = 0.0, cb = 0.0;
                        int idx = 0;
                        if (tb < fj.nt) {
                            idx = mirror ? fj.tp[tb] * 6 + pt.tp[ta] : pt.tp[ta] * 6 + fj.tp[tb];
                            const double kj = fj.tk[tb] * fj.s * (1.0 / 120.0);
                            ca = fa.s * pt.ka[ta] * kj * (pt.va[ta] == fj.tv[tb] ? 2.0 : 1.0);
                            cb = fb.s * pt.kb[ta] * kj * (pt.vb[ta] == fj.tv[tb] ? 2.0 : 1.0);
                        }
                        kc[ic][k][J][0] = ca;
                        kc[ic][k][J][1] = cb;
                        packed |= (unsigned long long)idx << (6 * k);
                    }
                kidx[ic][J] = packed;
                unsigned int gp = 0;
                for (int r = 0; r < 2; ++r) {
                    const FnTab& fi = r == 0 ? fa : fb;
                    const int I = ic + 10 * r;
                    const int lP = mirror ? fj.P : fi.P, lQ = mirror ? fj.Q : fi.Q;
                    const int rP = mirror ? fi.P : fj.P, rQ = mirror ? fi.Q : fj.Q;
                    const int gi[4] = {lP * 4 + rP, lP * 4 + rQ, lQ * 4 + rP, lQ * 4 + rQ};
                    for (int q = 0; q < 4; ++q) {
                        gp |= (unsigned int)gi[q] << (4 * (r * 4 + q));
                        mc[ic][r][q / 2][J][q % 2] = fi.s * fj.s * T.mc[I][J][q];
                    }
                }
                gidx[ic][J] = gp;
            }
        }
    }
};

// rows (ic, ic + 10), column J of the element matrices from the record, through the flat tables (what a lane evaluates)
EMB_HD void row_pair_flat(const KernTables& T, int ic, int J, const cx* D, const cx* g, const double* len, cx& Ka, cx& Kb,
                          cx& Ma, cx& Mb) {
    cx acca = mk(0.0), accb = mk(0.0);
    unsigned long long ki = T.kidx[ic][J];
    const int nk = T.nk[ic];
    for (int k = 0; k < nk; ++k) {
        const cx d = D[(int)(ki & 63ull)];
        ki >>= 6;
        fma_r(acca, T.kc[ic][k][J][0], d);
        fma_r(accb, T.kc[ic][k][J][1], d);
    }
    const double lj = len[T.lp[J]];
    const double sa = len[T.lp[ic]] * lj, sb = len[T.lp[ic + 10]] * lj;
    Ka = sa * acca;
    Kb = sb * accb;
    const unsigned int gi = T.gidx[ic][J];
    cx ma = T.mc[ic][0][0][J][0] * g[gi & 15u];
    fma_r(ma, T.mc[ic][0][0][J][1], g[(gi >> 4) & 15u]);
    fma_r(ma, T.mc[ic][0][1][J][0], g[(gi >> 8) & 15u]);
    fma_r(ma, T.mc[ic][0][1][J][1], g[(gi >> 12) & 15u]);
    cx mb = T.mc[ic][1][0][J][0] * g[(gi >> 16) & 15u];
    fma_r(mb, T.mc[ic][1][0][J][1], g[(gi >> 20) & 15u]);
    fma_r(mb, T.mc[ic][1][1][J][0], g[(gi >> 24) & 15u]);
    fma_r(mb, T.mc[ic][1][1][J][1], g[(gi >> 28) & 15u]);
    Ma = sa * ma;
    Mb = sb * mb;
}

}  // namespace ned2f
