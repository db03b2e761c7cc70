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

}  // namespace ned2f
