// Surface (Robin) terms, Dirichlet elimination and A(f) formation.
//
// Reference path replaced: fem/physics/edm/assembler.py:63-144 (generate_points, compute_bc_entries,
// assemble_robin_bc_excited, assemble_robin_bc), fem/mth/tri.py:575-796 (ned2_tri_stiff_force, ned2_tri_stiff),
// fem/physics/edm/assembler.py:333-385 (K = E - k0^2 B + sum B_p, PEC ids, solve_ids) and the sub-matrix
// extraction of fem/solver.py:434.  The reference rebuilds a whole-mesh triangle COO (64 entries for EVERY
// mesh triangle) and converts it to CSR per port per frequency; here only the surface's own triangles are
// touched, the gamma-free matrix S (B_p = gamma*S) is computed once, and its scatter onto the global pattern
// is a fixed, sorted list (entries of one slot are summed in list order => deterministic).
//
// The 8 triangle functions with local vertices (0,1,2) = ascending global ids and local edges (0,1),(1,2),(0,2):
//   edge-a: l lam_A w_AB   face-a: -l_02 lam_1 w_02   edge-b: l lam_B w_AB   face-b: +l_01 lam_2 w_01
//   S_ij = ss ll A/(2A)^2 [At(XiQiXjQj) Gd(Pi,Pj) - At(XiQiXjPj) Gd(Pi,Qj) - At(XiPiXjQj) Gd(Qi,Pj) + At(XiPiXjPj) Gd(Qi,Qj)]
//   At = 2 prod(m!)/6!  (area_coeff, fem/mth/optimized.py:262-271),  Gd(p,q) = (b_p b_q + c_p c_q).
#include "context.cuh"
#include <algorithm>
#include <cub/cub.cuh>
#include <numeric>

namespace tri2 {
struct Fn {
    int X, P, Q;
    double s;
    int lk;   // 0,1,2: edge length slot (edges 01,12,02);  3: Lt1=|v2-v0| (local);  4: Lt2=|v1-v0| (local)
};
__host__ __device__ constexpr int tA(int e) { return e == 1 ? 1 : 0; }
__host__ __device__ constexpr int tB(int e) { return e == 0 ? 1 : 2; }
__host__ __device__ constexpr Fn fn(int c) {
    if (c < 3) return Fn{tA(c), tA(c), tB(c), 1.0, c};
    if (c == 3) return Fn{1, 0, 2, -1.0, 3};
    if (c < 7) return Fn{tB(c - 4), tA(c - 4), tB(c - 4), 1.0, c - 4};
    return Fn{2, 0, 1, 1.0, 4};
}
__host__ __device__ constexpr double fact(int n) { return n <= 1 ? 1.0 : n * fact(n - 1); }
__host__ __device__ constexpr int mult(int v, int a, int b, int c, int d) { return (a == v) + (b == v) + (c == v) + (d == v); }
__host__ __device__ constexpr double At(int a, int b, int c, int d) {
    return 2.0 * fact(mult(0, a, b, c, d)) * fact(mult(1, a, b, c, d)) * fact(mult(2, a, b, c, d)) / 720.0;
}
}  // namespace tri2

// local 2-D vertex coordinates of every surface triangle + true 3-D edge lengths
__global__ void k_surf_geom(int64_t n, const int* __restrict__ tri, const int* __restrict__ tris,
                            const double* __restrict__ nodes, int frame, const double* __restrict__ binv_org,
                            double* __restrict__ xy, double* __restrict__ len3) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int it = tri[i];
    double v[3][3];
    for (int k = 0; k < 3; ++k) {
        const double* q = nodes + (int64_t)tris[(int64_t)it * 3 + k] * 3;
        v[k][0] = q[0]; v[k][1] = q[1]; v[k][2] = q[2];
    }
    double x[3], y[3];
    if (frame == 0) {
        for (int k = 0; k < 3; ++k) {
            const double d0 = v[k][0] - binv_org[9], d1 = v[k][1] - binv_org[10], d2 = v[k][2] - binv_org[11];
            x[k] = binv_org[0] * d0 + binv_org[1] * d1 + binv_org[2] * d2;
            y[k] = binv_org[3] * d0 + binv_org[4] * d1 + binv_org[5] * d2;
        }
    } else {
        // fem/mth/tri.py:709-723 restated: columns ax1, ax2=-(axn x ax1), axn=ax1 x ax2' (axn NOT normalised);
        // coordinates = pinv(basis) @ v; the columns are orthogonal with norms (1,s,s) so pinv = diag(1,1/s^2,1/s^2) B^T.
        double a1[3], a2[3], an[3], b2[3];
        double n1 = 0, n2 = 0;
        for (int k = 0; k < 3; ++k) {
            a1[k] = v[1][k] - v[0][k]; a2[k] = v[2][k] - v[0][k];
            n1 += a1[k] * a1[k]; n2 += a2[k] * a2[k];
        }
        n1 = 1.0 / sqrt(n1); n2 = 1.0 / sqrt(n2);
        for (int k = 0; k < 3; ++k) { a1[k] *= n1; a2[k] *= n2; }
        an[0] = a1[1] * a2[2] - a1[2] * a2[1]; an[1] = a1[2] * a2[0] - a1[0] * a2[2]; an[2] = a1[0] * a2[1] - a1[1] * a2[0];
        b2[0] = -(an[1] * a1[2] - an[2] * a1[1]); b2[1] = -(an[2] * a1[0] - an[0] * a1[2]); b2[2] = -(an[0] * a1[1] - an[1] * a1[0]);
        const double s2 = b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2];
        for (int k = 0; k < 3; ++k) {
            x[k] = a1[0] * v[k][0] + a1[1] * v[k][1] + a1[2] * v[k][2];
            y[k] = (b2[0] * v[k][0] + b2[1] * v[k][1] + b2[2] * v[k][2]) / s2;
        }
    }
    for (int k = 0; k < 3; ++k) { xy[i * 6 + k] = x[k]; xy[i * 6 + 3 + k] = y[k]; }
    auto d3 = [&](int a, int b) {
        const double dx = v[a][0] - v[b][0], dy = v[a][1] - v[b][1], dz = v[a][2] - v[b][2];
        return sqrt(dx * dx + dy * dy + dz * dz);
    };
    len3[i * 3 + 0] = d3(0, 1); len3[i * 3 + 1] = d3(1, 2); len3[i * 3 + 2] = d3(0, 2);
}

struct TriGeom {
    double b[3], c[3], a[3], A, sA, L[5];
};
__device__ __forceinline__ void tri_geom(const double* __restrict__ xy, const double* __restrict__ len3, int frame, TriGeom& g) {
    const double x0 = xy[0], x1 = xy[1], x2 = xy[2], y0 = xy[3], y1 = xy[4], y2 = xy[5];
    g.a[0] = x1 * y2 - y1 * x2; g.a[1] = x2 * y0 - y2 * x0; g.a[2] = x0 * y1 - y0 * x1;
    g.b[0] = y1 - y2; g.b[1] = y2 - y0; g.b[2] = y0 - y1;
    g.c[0] = x2 - x1; g.c[1] = x0 - x2; g.c[2] = x1 - x0;
    g.sA = (x0 - x2) * (y1 - y0) - (x0 - x1) * (y2 - y0);
    g.A = 0.5 * fabs(g.sA);
    const double d01 = hypot(x0 - x1, y0 - y1), d12 = hypot(x1 - x2, y1 - y2), d02 = hypot(x0 - x2, y0 - y2);
    if (frame == 0) { g.L[0] = d01; g.L[1] = d12; g.L[2] = d02; }      // tri.py:634 (computed distances)
    else { g.L[0] = len3[0]; g.L[1] = len3[1]; g.L[2] = len3[2]; }      // tri.py:757 (table lengths)
    g.L[3] = d02; g.L[4] = d01;                                         // Lt1, Lt2 (tri.py:624)
}

template <int I, int J>
__device__ __forceinline__ double s_entry(const TriGeom& g, const double (&gd)[3][3]) {
    constexpr tri2::Fn fi = tri2::fn(I), fj = tri2::fn(J);
    constexpr double c1 = tri2::At(fi.X, fi.Q, fj.X, fj.Q), c2 = tri2::At(fi.X, fi.Q, fj.X, fj.P);
    constexpr double c3 = tri2::At(fi.X, fi.P, fj.X, fj.Q), c4 = tri2::At(fi.X, fi.P, fj.X, fj.P);
    const double v = c1 * gd[fi.P][fj.P] - c2 * gd[fi.P][fj.Q] - c3 * gd[fi.Q][fj.P] + c4 * gd[fi.Q][fj.Q];
    return (fi.s * fj.s) * g.L[fi.lk] * g.L[fj.lk] * v;
}
template <int I, int... Js>
__device__ __forceinline__ void s_row(const TriGeom& g, const double (&gd)[3][3], double sc, double* out,
                                      std::integer_sequence<int, Js...>) {
    ((out[I * 8 + Js] = sc * s_entry<I, Js>(g, gd)), ...);
}
template <int... Is>
__device__ __forceinline__ void s_all(const TriGeom& g, const double (&gd)[3][3], double sc, double* out,
                                      std::integer_sequence<int, Is...>) {
    (s_row<Is>(g, gd, sc, out, std::make_integer_sequence<int, 8>{}), ...);
}

__global__ void k_surf_S(int64_t n, const double* __restrict__ xy, const double* __restrict__ len3, int frame,
                         double* __restrict__ S) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    TriGeom g;
    tri_geom(xy + i * 6, len3 + i * 3, frame, g);
    double gd[3][3];
    for (int p = 0; p < 3; ++p)
        for (int q = 0; q < 3; ++q) gd[p][q] = g.b[p] * g.b[q] + g.c[p] * g.c[q];
    const double sc = g.A / ((2 * g.A) * (2 * g.A));     // AREA_COEFF*Area and COEFF/gamma (tri.py:630-631)
    s_all(g, gd, sc, S + i * 64, std::make_integer_sequence<int, 8>{});
}

// Dunavant degree-4, 6 points, in the order gaus_quad_tri(4) produces (fem/mth/optimized.py:28-29,77-108)
__constant__ double c_dw[6] = {0.223381589678011, 0.223381589678011, 0.223381589678011,
                               0.109951743655322, 0.109951743655322, 0.109951743655322};
__constant__ double c_dl[6][3] = {{0.108103018168070, 0.445948490915965, 0.445948490915965},
                                  {0.445948490915965, 0.445948490915965, 0.108103018168070},
                                  {0.445948490915965, 0.108103018168070, 0.445948490915965},
                                  {0.816847572980459, 0.091576213509771, 0.091576213509771},
                                  {0.091576213509771, 0.091576213509771, 0.816847572980459},
                                  {0.091576213509771, 0.816847572980459, 0.091576213509771}};

// xy_out (2,6,n): x then y of the 6 points of every triangle (generate_points, assembler.py:63-81)
__global__ void k_surf_points(int64_t n, const double* __restrict__ xy, double* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* v = xy + i * 6;
    for (int q = 0; q < 6; ++q) {
        out[(int64_t)q * n + i] = v[0] * c_dl[q][0] + v[1] * c_dl[q][1] + v[2] * c_dl[q][2];
        out[(int64_t)(6 + q) * n + i] = v[3] * c_dl[q][0] + v[4] * c_dl[q][1] + v[5] * c_dl[q][2];
    }
}

template <int I>
__device__ __forceinline__ cx f_entry(const TriGeom& g, const double (&lam)[3][6], const cx (&Ux)[6], const cx (&Uy)[6]) {
    constexpr tri2::Fn f = tri2::fn(I);
    const double i2A = 1.0 / (2 * g.A);
    cx acc = mk(0.0);
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const double wx = (lam[f.Q][q] * g.b[f.P] - lam[f.P][q] * g.b[f.Q]) * i2A;
        const double wy = (lam[f.Q][q] * g.c[f.P] - lam[f.P][q] * g.c[f.Q]) * i2A;
        const double nx = lam[f.X][q] * wx, ny = lam[f.X][q] * wy;
        acc += c_dw[q] * (nx * Ux[q] + ny * Uy[q]);
    }
    // lengths of the forcing are always computed distances (tri.py:634,624)
    const double L = g.L[f.lk];   // frame-0 geometry: all five are computed distances
    const double signA = g.sA > 0 ? -1.0 : (g.sA < 0 ? 1.0 : 0.0);
    return (f.s * L * signA * g.A) * acc;
}

// U (3,6,n) c128 -> per-triangle forcing bloc [n][8]   (ned2_tri_stiff_force, tri.py:612-613,670-693)
__global__ void k_surf_force(int64_t n, const double* __restrict__ xy, const double* __restrict__ len3,
                             const cx* __restrict__ U, cx* __restrict__ bloc) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    TriGeom g;
    tri_geom(xy + i * 6, len3 + i * 3, 0, g);
    double lam[3][6];
    cx Ux[6], Uy[6];
    const double* v = xy + i * 6;
    for (int q = 0; q < 6; ++q) {
        const double x = v[0] * c_dl[q][0] + v[1] * c_dl[q][1] + v[2] * c_dl[q][2];
        const double y = v[3] * c_dl[q][0] + v[4] * c_dl[q][1] + v[5] * c_dl[q][2];
        for (int k = 0; k < 3; ++k) lam[k][q] = (g.a[k] + g.b[k] * x + g.c[k] * y) / (2 * g.A);
        Ux[q] = U[(int64_t)q * n + i];
        Uy[q] = U[(int64_t)(6 + q) * n + i];
    }
    cx* o = bloc + i * 8;
    o[0] = f_entry<0>(g, lam, Ux, Uy); o[1] = f_entry<1>(g, lam, Ux, Uy); o[2] = f_entry<2>(g, lam, Ux, Uy);
    o[3] = f_entry<3>(g, lam, Ux, Uy); o[4] = f_entry<4>(g, lam, Ux, Uy); o[5] = f_entry<5>(g, lam, Ux, Uy);
    o[6] = f_entry<6>(g, lam, Ux, Uy); o[7] = f_entry<7>(g, lam, Ux, Uy);
}

// full-pattern slot of every (tri,i,j) and dof of every (tri,i)
__global__ void k_surf_slots(int64_t n, const int* __restrict__ tri, const int* __restrict__ tri2f, int64_t nTri,
                             const int64_t* __restrict__ rowptr, const int* __restrict__ col, int64_t* __restrict__ slot,
                             int* __restrict__ dofs) {
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= n * 64) return;
    const int64_t i = idx / 64;
    const int ij = (int)(idx % 64), a = ij / 8, b = ij % 8;
    const int it = tri[i];
    const int r = tri2f[(int64_t)a * nTri + it], cc = tri2f[(int64_t)b * nTri + it];
    int64_t lo = rowptr[r], hi = rowptr[r + 1] - 1;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (col[mid] < cc) lo = mid + 1; else hi = mid;
    }
    slot[idx] = (col[lo] == cc) ? lo : -1;
    if (b == 0) dofs[i * 8 + a] = r;
}

template <typename T>
__global__ void k_seg_sum(int64_t nseg, const int* __restrict__ segptr, const int* __restrict__ ent,
                          const T* __restrict__ vals, T* __restrict__ out) {
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u >= nseg) return;
    T acc = vals[ent[segptr[u]]];
    for (int k = segptr[u] + 1; k < segptr[u + 1]; ++k) acc = acc + vals[ent[k]];
    out[u] = acc;
}

template <typename KeyT>
static void group_by_key(const std::vector<KeyT>& keys, std::vector<KeyT>& ukeys, std::vector<int>& segptr,
                         std::vector<int>& ent) {
    const int n = (int)keys.size();
    ent.resize(n);
    std::iota(ent.begin(), ent.end(), 0);
    std::stable_sort(ent.begin(), ent.end(), [&](int a, int b) { return keys[a] < keys[b]; });
    ukeys.clear();
    segptr.clear();
    for (int k = 0; k < n; ++k)
        if (k == 0 || keys[ent[k]] != keys[ent[k - 1]]) {
            ukeys.push_back(keys[ent[k]]);
            segptr.push_back(k);
        }
    segptr.push_back(n);
}

extern "C" int emb_surface_define(emb_ctx* c, int sid, int64_t ntri, const int64_t* tri_ids, int frame,
                                  const double* basis_inv, const double* origin) {
    if (!c || sid < 0 || sid >= 16 || ntri <= 0 || !tri_ids || (frame == 0 && (!basis_inv || !origin))) {
        if (c) c->err = "emb_surface_define: bad argument";
        return EMB_ERR_ARG;
    }
    if (!c->have_pattern) {
        c->err = "emb_surface_define: needs emb_symbolic first";
        return EMB_ERR_STATE;
    }
    PhaseTimer pt(c, "surface");
    Surface& s = c->surf[sid];
    s.defined = false;
    s.has_rhs = false;
    s.frame = frame;
    s.ntri = ntri;
    std::vector<int> ht(ntri);
    for (int64_t i = 0; i < ntri; ++i) {
        if (tri_ids[i] < 0 || tri_ids[i] >= c->nTri) {
            c->err = "emb_surface_define: triangle id out of range";
            return EMB_ERR_ARG;
        }
        ht[i] = (int)tri_ids[i];
    }
    EMB_TRY(h2d(c, s.tri, ht.data(), (size_t)ntri));
    double hb[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    if (frame == 0) {
        memcpy(hb, basis_inv, 9 * sizeof(double));
        memcpy(hb + 9, origin, 3 * sizeof(double));
    }
    DevBuf<double> dB, len3;
    EMB_TRY(h2d(c, dB, hb, 12));
    EMB_TRY(dev_alloc(c, len3, (size_t)ntri * 3));
    EMB_TRY(dev_alloc(c, s.xy, (size_t)ntri * 6));
    EMB_TRY(dev_alloc(c, s.S, (size_t)ntri * 64));
    EMB_TRY(dev_alloc(c, s.bloc, (size_t)ntri * 8));
    k_surf_geom<<<blocks_for(ntri, 128), 128, 0, c->stream>>>(ntri, s.tri.p, c->tris.p, c->nodes.p, frame, dB.p, s.xy.p, len3.p);
    EMB_LAUNCH_CHECK(c);
    k_surf_S<<<blocks_for(ntri, 128), 128, 0, c->stream>>>(ntri, s.xy.p, len3.p, frame, s.S.p);
    EMB_LAUNCH_CHECK(c);
    DevBuf<int64_t> dslot;
    DevBuf<int> ddof;
    EMB_TRY(dev_alloc(c, dslot, (size_t)ntri * 64));
    EMB_TRY(dev_alloc(c, ddof, (size_t)ntri * 8));
    k_surf_slots<<<blocks_for(ntri * 64, 256), 256, 0, c->stream>>>(ntri, s.tri.p, c->tri2f.p, c->nTri, c->rowptr.p, c->col.p,
                                                                   dslot.p, ddof.p);
    EMB_LAUNCH_CHECK(c);
    std::vector<int64_t> hslot((size_t)ntri * 64);
    std::vector<int> hdof((size_t)ntri * 8);
    EMB_CUDA(c, cudaMemcpyAsync(hslot.data(), dslot.p, hslot.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaMemcpyAsync(hdof.data(), ddof.p, hdof.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (auto v : hslot)
        if (v < 0) {
            c->err = "emb_surface_define: surface triangle couples dofs outside the tet pattern";
            return EMB_ERR_ARG;
        }
    std::vector<int64_t> uslot;
    std::vector<int> seg, ent, udof, dseg, dent;
    group_by_key(hslot, uslot, seg, ent);
    group_by_key(hdof, udof, dseg, dent);
    s.nslot = (int64_t)uslot.size();
    s.ndof = (int64_t)udof.size();
    EMB_TRY(h2d(c, s.slot, uslot.data(), uslot.size()));
    EMB_TRY(h2d(c, s.segptr, seg.data(), seg.size()));
    EMB_TRY(h2d(c, s.ent, ent.data(), ent.size()));
    EMB_TRY(h2d(c, s.dof, udof.data(), udof.size()));
    EMB_TRY(h2d(c, s.dsegptr, dseg.data(), dseg.size()));
    EMB_TRY(h2d(c, s.dent, dent.data(), dent.size()));
    EMB_TRY(dev_alloc(c, s.Sval, (size_t)s.nslot));
    EMB_TRY(dev_alloc(c, s.bval, (size_t)s.ndof));
    k_seg_sum<double><<<blocks_for(s.nslot, 128), 128, 0, c->stream>>>(s.nslot, s.segptr.p, s.ent.p, s.S.p, s.Sval.p);
    EMB_LAUNCH_CHECK(c);
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    s.slot_s.release();
    dB.release(); len3.release(); dslot.release(); ddof.release();
    // keep len3? the forcing only uses frame-0 geometry (computed distances), so it is not needed again
    s.defined = true;
    return EMB_OK;
}

extern "C" int emb_surface_points(emb_ctx* c, int sid, double* xy) {
    if (!c || sid < 0 || sid >= 16 || !xy) return EMB_ERR_ARG;
    Surface& s = c->surf[sid];
    if (!s.defined) { c->err = "emb_surface_points: surface not defined"; return EMB_ERR_STATE; }
    DevBuf<double> out;
    EMB_TRY(dev_alloc(c, out, (size_t)s.ntri * 12));
    k_surf_points<<<blocks_for(s.ntri, 128), 128, 0, c->stream>>>(s.ntri, s.xy.p, out.p);
    EMB_LAUNCH_CHECK(c);
    EMB_CUDA(c, cudaMemcpyAsync(xy, out.p, (size_t)s.ntri * 12 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    out.release();
    return EMB_OK;
}

extern "C" int emb_surface_blocks(emb_ctx* c, int sid, double* S) {
    if (!c || sid < 0 || sid >= 16 || !S) return EMB_ERR_ARG;
    Surface& s = c->surf[sid];
    if (!s.defined) { c->err = "emb_surface_blocks: surface not defined"; return EMB_ERR_STATE; }
    EMB_CUDA(c, cudaMemcpyAsync(S, s.S.p, (size_t)s.ntri * 64 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    return EMB_OK;
}

extern "C" int emb_surface_set_U(emb_ctx* c, int sid, const emb_c128* U, emb_c128* b_full) {
    if (!c || sid < 0 || sid >= 16 || !U) return EMB_ERR_ARG;
    Surface& s = c->surf[sid];
    if (!s.defined) { c->err = "emb_surface_set_U: surface not defined"; return EMB_ERR_STATE; }
    DevBuf<cx> dU;
    DevBuf<double> len3;
    EMB_TRY(h2d(c, dU, reinterpret_cast<const cx*>(U), (size_t)s.ntri * 18));
    EMB_TRY(dev_alloc(c, len3, (size_t)s.ntri * 3));   // unused by frame-0 geometry
    k_surf_force<<<blocks_for(s.ntri, 128), 128, 0, c->stream>>>(s.ntri, s.xy.p, len3.p, dU.p, s.bloc.p);
    EMB_LAUNCH_CHECK(c);
    k_seg_sum<cx><<<blocks_for(s.ndof, 128), 128, 0, c->stream>>>(s.ndof, s.dsegptr.p, s.dent.p, s.bloc.p, s.bval.p);
    EMB_LAUNCH_CHECK(c);
    if (b_full) {
        std::vector<cx> hv((size_t)s.ndof);
        std::vector<int> hd((size_t)s.ndof);
        EMB_CUDA(c, cudaMemcpyAsync(hv.data(), s.bval.p, hv.size() * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
        EMB_CUDA(c, cudaMemcpyAsync(hd.data(), s.dof.p, hd.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        EMB_CUDA(c, cudaStreamSynchronize(c->stream));
        memset(b_full, 0, (size_t)c->N * sizeof(emb_c128));
        for (size_t k = 0; k < hv.size(); ++k) { b_full[hd[k]].re = hv[k].re; b_full[hd[k]].im = hv[k].im; }
    }
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    dU.release(); len3.release();
    s.has_rhs = true;
    return EMB_OK;
}

// ------------------------------------------------------------------------------------------------
// Dirichlet elimination: solve-space pattern = rows/cols of the non-PEC dofs
// ------------------------------------------------------------------------------------------------
__global__ void k_mark(int64_t n, const int64_t* __restrict__ ids, int64_t N, int* __restrict__ keep) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && ids[i] >= 0 && ids[i] < N) keep[ids[i]] = 0;
}
__global__ void k_fill_int(int* v, int64_t n, int val) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] = val;
}
// pair check: keep[d] == keep[d + H] for every d < H
__global__ void k_check_pairs(int64_t H, const int* __restrict__ keep, int* __restrict__ flag) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < H && keep[i] != keep[i + H]) atomicExch(flag, 1);
}
// H > 0: pair ordering (solve index = 2 * rank inside its half + half); H == 0: ascending dof order
__global__ void k_newid(int64_t N, int64_t H, int64_t Ns, const int* __restrict__ keep, const int* __restrict__ scan,
                        int* __restrict__ newid, int* __restrict__ solve_ids, int* __restrict__ sperm) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (!keep[i]) { newid[i] = -1; return; }
    int id = scan[i];
    if (H > 0) id = i < H ? 2 * scan[i] : 2 * (scan[i] - (int)(Ns / 2)) + 1;
    newid[i] = id;
    solve_ids[id] = (int)i;
    sperm[scan[i]] = id;
}
// one warp per kept row.  H > 0 (pair ordering): the kept columns of a row are its first-half entities followed by
// the same entities in the second half; they are emitted interleaved (2q, 2q+1) so the row stays sorted.
template <bool FILL>
__global__ void k_compact_rows(int64_t Ns, int64_t H, const int* __restrict__ solve_ids, const int* __restrict__ newid,
                               const int64_t* __restrict__ rowptr, const int* __restrict__ col, int64_t* __restrict__ rowlen,
                               const int64_t* __restrict__ rowptr_s, int* __restrict__ col_s, int64_t* __restrict__ src) {
    const int lane = threadIdx.x & 31;
    const int64_t rs = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (rs >= Ns) return;
    const int r = solve_ids[rs];
    const int64_t p0 = rowptr[r], p1 = rowptr[r + 1];
    int na = 0;                      // kept columns in the first half (pair ordering)
    if (FILL && H > 0)
        for (int64_t k0 = p0; k0 < p1; k0 += 32) {
            const int64_t k = k0 + lane;
            bool f = false;
            if (k < p1) { const int cc = col[k]; f = cc < H && newid[cc] >= 0; }
            na += __popc(__ballot_sync(0xffffffffu, f));
        }
    int base = 0;
    const int64_t o0 = FILL ? rowptr_s[rs] : 0;
    for (int64_t k0 = p0; k0 < p1; k0 += 32) {
        const int64_t k = k0 + lane;
        int nc = -1, cc = 0;
        if (k < p1) { cc = col[k]; nc = newid[cc]; }
        const unsigned m = __ballot_sync(0xffffffffu, nc >= 0);
        if (FILL && nc >= 0) {
            const int rank = base + __popc(m & ((1u << lane) - 1));
            int64_t o = o0 + rank;
            if (H > 0) o = o0 + (cc < H ? 2 * rank : 2 * (rank - na) + 1);
            col_s[o] = nc;
            src[o] = k;
        }
        base += __popc(m);
    }
    if (!FILL && lane == 0) rowlen[rs] = base;
}
// entity column of every 2x2 block: block q of block row j sits at entries rowptr_s[2j] + 2q (+1) of rows 2j, 2j+1
__global__ void k_blkcol(int64_t nblkrow, const int64_t* __restrict__ rowptr_s, const int* __restrict__ col_s,
                         int* __restrict__ blkcol, int* __restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (j >= nblkrow) return;
    const int64_t p0 = rowptr_s[2 * j], p1 = rowptr_s[2 * j + 1], p2 = rowptr_s[2 * j + 2];
    if (p1 - p0 != p2 - p1 || ((p1 - p0) & 1) || (p0 & 3)) { if (lane == 0) atomicExch(flag, 1); return; }
    const int64_t nb = (p1 - p0) >> 1;
    for (int64_t q = lane; q < nb; q += 32) {
        const int c0 = col_s[p0 + 2 * q], c1 = col_s[p0 + 2 * q + 1];
        if ((c0 & 1) || c1 != c0 + 1 || col_s[p1 + 2 * q] != c0 || col_s[p1 + 2 * q + 1] != c1) atomicExch(flag, 1);
        blkcol[(p0 >> 2) + q] = c0 >> 1;
    }
}

extern "C" int emb_set_dirichlet(emb_ctx* c, int64_t npec, const int64_t* pec_ids) {
    if (!c || npec < 0 || (npec > 0 && !pec_ids)) return EMB_ERR_ARG;
    if (!c->have_pattern) { c->err = "emb_set_dirichlet: needs emb_symbolic first"; return EMB_ERR_STATE; }
    const int64_t N = c->N;
    DevBuf<int> keep, scan;
    DevBuf<int64_t> dids, rowlen;
    DevBuf<char> tmp;
    EMB_TRY(dev_alloc(c, keep, (size_t)N));
    EMB_TRY(dev_alloc(c, scan, (size_t)N));
    EMB_TRY(dev_alloc(c, c->newid, (size_t)N));
    k_fill_int<<<blocks_for(N, 256), 256, 0, c->stream>>>(keep.p, N, 1);
    EMB_LAUNCH_CHECK(c);
    if (npec) {
        EMB_TRY(h2d(c, dids, pec_ids, (size_t)npec));
        k_mark<<<blocks_for(npec, 256), 256, 0, c->stream>>>(npec, dids.p, N, keep.p);
        EMB_LAUNCH_CHECK(c);
    }
    size_t tb = 0;
    EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tb, keep.p, scan.p, (int)N, c->stream));
    EMB_TRY(dev_alloc(c, tmp, tb));
    EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, tb, keep.p, scan.p, (int)N, c->stream));
    c->launches += 2;
    int last_scan = 0, last_keep = 0;
    EMB_CUDA(c, cudaMemcpyAsync(&last_scan, scan.p + N - 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaMemcpyAsync(&last_keep, keep.p + N - 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->Ns = (int64_t)last_scan + last_keep;
    if (c->Ns <= 0) { c->err = "emb_set_dirichlet: every dof is eliminated"; return EMB_ERR_ARG; }
    EMB_TRY(dev_alloc(c, c->solve_ids, (size_t)c->Ns));
    EMB_TRY(dev_alloc(c, c->sperm, (size_t)c->Ns));
    // pair ordering when both functions of every entity share their fate (fem/elements/nedelec2.py:46-62: function b of an
    // edge / face is dof a + nE + nTri)
    const int64_t Hh = c->nE + c->nTri;
    DevBuf<int> flag;
    EMB_TRY(dev_alloc(c, flag, 1));
    EMB_CUDA(c, cudaMemsetAsync(flag.p, 0, sizeof(int), c->stream));
    bool paired = (2 * Hh == N) && (c->Ns % 2 == 0) && !(getenv("EMB_NO_PAIRING") && atoi(getenv("EMB_NO_PAIRING")));
    if (paired) {
        k_check_pairs<<<blocks_for(Hh, 256), 256, 0, c->stream>>>(Hh, keep.p, flag.p);
        EMB_LAUNCH_CHECK(c);
        int hf = 0;
        EMB_CUDA(c, cudaMemcpyAsync(&hf, flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        EMB_CUDA(c, cudaStreamSynchronize(c->stream));
        paired = hf == 0;
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        const int64_t H = paired ? Hh : 0;
        k_newid<<<blocks_for(N, 256), 256, 0, c->stream>>>(N, H, c->Ns, keep.p, scan.p, c->newid.p, c->solve_ids.p, c->sperm.p);
        EMB_LAUNCH_CHECK(c);
        EMB_TRY(dev_alloc(c, rowlen, (size_t)c->Ns + 1));
        EMB_TRY(dev_alloc(c, c->rowptr_s, (size_t)c->Ns + 1));
        EMB_CUDA(c, cudaMemsetAsync(rowlen.p + c->Ns, 0, sizeof(int64_t), c->stream));
        k_compact_rows<false><<<blocks_for(c->Ns * 32, 256), 256, 0, c->stream>>>(c->Ns, H, c->solve_ids.p, c->newid.p, c->rowptr.p,
                                                                                 c->col.p, rowlen.p, nullptr, nullptr, nullptr);
        EMB_LAUNCH_CHECK(c);
        size_t tb2 = 0;
        EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tb2, rowlen.p, c->rowptr_s.p, (int)(c->Ns + 1), c->stream));
        if (tb2 > tmp.n) EMB_TRY(dev_alloc(c, tmp, tb2));
        EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, tb2, rowlen.p, c->rowptr_s.p, (int)(c->Ns + 1), c->stream));
        c->launches += 2;
        EMB_CUDA(c, cudaMemcpyAsync(&c->nnz_s, c->rowptr_s.p + c->Ns, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
        EMB_CUDA(c, cudaStreamSynchronize(c->stream));
        EMB_TRY(dev_alloc(c, c->col_s, (size_t)c->nnz_s));
        EMB_TRY(dev_alloc(c, c->src, (size_t)c->nnz_s));
        k_compact_rows<true><<<blocks_for(c->Ns * 32, 256), 256, 0, c->stream>>>(c->Ns, H, c->solve_ids.p, c->newid.p, c->rowptr.p,
                                                                                c->col.p, nullptr, c->rowptr_s.p, c->col_s.p, c->src.p);
        EMB_LAUNCH_CHECK(c);
        c->blkcol.release();
        if (!paired) break;
        // block structure (verified entry by entry; a mesh whose pattern is not pair-symmetric falls back to plain CSR)
        EMB_TRY(dev_alloc(c, c->blkcol, (size_t)(c->nnz_s / 4 + 1)));
        k_blkcol<<<blocks_for((c->Ns / 2) * 32, 256), 256, 0, c->stream>>>(c->Ns / 2, c->rowptr_s.p, c->col_s.p, c->blkcol.p, flag.p);
        EMB_LAUNCH_CHECK(c);
        int hf = 0;
        EMB_CUDA(c, cudaMemcpyAsync(&hf, flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        EMB_CUDA(c, cudaStreamSynchronize(c->stream));
        if (hf == 0) break;
        paired = false;
    }
    c->paired = paired;
    c->sell_ready = c->sell_active = c->sell_tried = false;      // the SELL layout of the inner operator follows the pattern
    c->sell_zeroed = nullptr;
    c->sell_rows.release(); c->sell_pos.release(); c->sell_bcol.release(); c->sell_sptr.release();
    flag.release();
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    keep.release(); scan.release(); dids.release(); rowlen.release(); tmp.release();
    for (auto& s : c->surf) { s.slot_s.release(); s.mv_slot.release(); s.mv_val.release(); }
    c->have_dirichlet = true;
    c->have_A = false;
    c->rc_n = 0; c->rc_nq = 0;
    if (c->rc_cap > 0) {   // vectors are sized by the solve space
        c->rcU.release(); c->rcU32.release(); c->rcQ.release(); c->rc_part.release(); c->rc_tmp.release(); c->rc_x0.release();
        c->rc_cap = 0;
    }
    // every solve-space buffer is sized by Ns
    for (auto& w : c->work) w.release();
    c->xs.release(); c->bs.release(); c->As.release(); c->As32.release(); c->have_As = false;
    return EMB_OK;
}

extern "C" int emb_get_solve_ids(emb_ctx* c, int64_t* out) {
    if (!c || !out) return EMB_ERR_ARG;
    if (!c->have_dirichlet) { c->err = "emb_get_solve_ids: emb_set_dirichlet not called"; return EMB_ERR_STATE; }
    // ascending dof order, as the reference builds it (assembler.py:385); position s is solve index sperm[s]
    std::vector<int> h((size_t)c->Ns), pm((size_t)c->Ns);
    EMB_CUDA(c, cudaMemcpyAsync(h.data(), c->solve_ids.p, h.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaMemcpyAsync(pm.data(), c->sperm.p, pm.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < h.size(); ++i) out[i] = h[(size_t)pm[i]];
    return EMB_OK;
}
// solve index of the s-th kept dof in ascending dof order (the order of emb_get_solve_ids); vectors and matrices that
// cross the ABI on the solve space (emb_get_csr which=2, emb_spmv_host, emb_aux_add*) are in SOLVE-INDEX order
extern "C" int emb_get_solve_perm(emb_ctx* c, int64_t* perm) {
    if (!c || !perm) return EMB_ERR_ARG;
    if (!c->have_dirichlet) { c->err = "emb_get_solve_perm: emb_set_dirichlet not called"; return EMB_ERR_STATE; }
    std::vector<int> pm((size_t)c->Ns);
    EMB_CUDA(c, cudaMemcpyAsync(pm.data(), c->sperm.p, pm.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < pm.size(); ++i) perm[i] = pm[i];
    return EMB_OK;
}
extern "C" int emb_is_paired(const emb_ctx* c) { return c && c->paired ? 1 : 0; }

// ------------------------------------------------------------------------------------------------
// A(f) = E - k0^2 B + sum gamma_s S_s  on the solve-space pattern
// ------------------------------------------------------------------------------------------------
// HBM-bound stream: per solve-space entry 8 B index + 2x16 B gathered values in, 16 B out.
__global__ void k_form_A(int64_t nnz_s, const int64_t* __restrict__ src, const cx* __restrict__ K,
                         const cx* __restrict__ M, double k02, cx* __restrict__ A) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= nnz_s) return;
    const int64_t p = src[k];
    const double2 a = *reinterpret_cast<const double2*>(K + p);
    const double2 b = *reinterpret_cast<const double2*>(M + p);
    *reinterpret_cast<double2*>(A + k) = make_double2(a.x - k02 * b.x, a.y - k02 * b.y);
}

// solve-space slot of each unique full-pattern slot of a surface (-1 if its row or column is eliminated)
__global__ void k_surf_slot_s(int64_t n, const int64_t* __restrict__ slot, const int64_t* __restrict__ rowptr,
                              const int* __restrict__ col, int64_t N, const int* __restrict__ newid,
                              const int64_t* __restrict__ rowptr_s, const int* __restrict__ col_s,
                              int64_t* __restrict__ slot_s) {
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u >= n) return;
    const int64_t p = slot[u];
    int64_t lo = 0, hi = N - 1;          // row of p: last r with rowptr[r] <= p
    while (lo < hi) {
        int64_t mid = (lo + hi + 1) >> 1;
        if (rowptr[mid] <= p) lo = mid; else hi = mid - 1;
    }
    const int rs = newid[lo], cs = newid[col[p]];
    int64_t out = -1;
    if (rs >= 0 && cs >= 0) {
        int64_t a = rowptr_s[rs], b = rowptr_s[rs + 1] - 1;
        while (a < b) {
            int64_t mid = (a + b) >> 1;
            if (col_s[mid] < cs) a = mid + 1; else b = mid;
        }
        out = a;
    }
    slot_s[u] = out;
}

__global__ void k_add_surface(int64_t n, const int64_t* __restrict__ slot_s, const double* __restrict__ Sval, cx gamma,
                              cx* __restrict__ A) {
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u >= n) return;
    const int64_t p = slot_s[u];
    if (p < 0) return;
    const double s = Sval[u];
    cx a = A[p];
    a.re += gamma.re * s;
    a.im += gamma.im * s;
    A[p] = a;
}

extern "C" int emb_form_A(emb_ctx* c, double k0, int nsurf, const int* sids, const emb_c128* gammas) {
    if (!c || nsurf < 0 || (nsurf > 0 && (!sids || !gammas))) return EMB_ERR_ARG;
    if (!c->have_KM || !c->have_dirichlet) {
        c->err = "emb_form_A: needs emb_assemble_KM and emb_set_dirichlet first";
        return EMB_ERR_STATE;
    }
    EMB_TRY(dev_alloc(c, c->A, (size_t)c->nnz_s));
    for (int i = 0; i < nsurf; ++i) {
        if (sids[i] < 0 || sids[i] >= 16 || !c->surf[sids[i]].defined) {
            c->err = "emb_form_A: undefined surface id";
            return EMB_ERR_ARG;
        }
        Surface& s = c->surf[sids[i]];
        if (!s.slot_s.p) {
            EMB_TRY(dev_alloc(c, s.slot_s, (size_t)s.nslot));
            k_surf_slot_s<<<blocks_for(s.nslot, 128), 128, 0, c->stream>>>(s.nslot, s.slot.p, c->rowptr.p, c->col.p, c->N,
                                                                          c->newid.p, c->rowptr_s.p, c->col_s.p, s.slot_s.p);
            EMB_LAUNCH_CHECK(c);
        }
    }
    {
        PhaseTimer pt(c, "form_A");
        k_form_A<<<blocks_for(c->nnz_s, 256), 256, 0, c->stream>>>(c->nnz_s, c->src.p, c->K.p, c->M.p, k0 * k0, c->A.p);
        EMB_LAUNCH_CHECK(c);
        for (int i = 0; i < nsurf; ++i) {   // stream order = caller's order: deterministic when surfaces overlap
            Surface& s = c->surf[sids[i]];
            k_add_surface<<<blocks_for(s.nslot, 128), 128, 0, c->stream>>>(s.nslot, s.slot_s.p, s.Sval.p,
                                                                          cx{gammas[i].re, gammas[i].im}, c->A.p);
            EMB_LAUNCH_CHECK(c);
        }
    }
    c->k0 = k0;
    c->have_A = true;
    c->have_As = false;
    // affine coefficients of this A(f) for the reduced-basis projection (recycle.cuh): K, M, surfaces in caller order
    c->aff_coef.assign(1, std::complex<double>(1.0, 0.0));
    c->aff_coef.push_back(std::complex<double>(-k0 * k0, 0.0));
    c->aff_sids.assign(sids, sids + nsurf);
    for (int i = 0; i < nsurf; ++i) c->aff_coef.push_back(std::complex<double>(gammas[i].re, gammas[i].im));
    return EMB_OK;
}
