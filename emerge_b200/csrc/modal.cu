// Port boundary-mode analysis on the device: element matrices of the mixed Nedelec / Lagrange eigenproblem
// (reference fem/physics/edm/nedeleclegrange2.py:223-417, assembler.py:246-308) and the shift-invert operator of its
// eigen-solve (reference fem/solver.py:311-357: dense scipy.linalg.eig or ARPACK eigsh with sigma = -target_kz^2).
#include "context.cuh"

// ---- port boundary-mode analysis: element matrices ----------------------------------------------------
// Replaces the numba prange loop _matrix_builder / generalized_matrix_GQ (reference fem/physics/edm/nedeleclegrange2.py:
// 223-417).  16 lanes per port triangle, lane = matrix row (14 live): each lane sets the triangle up (a few hundred flops)
// and integrates its row of both 14 x 14 matrices with the 6-point rule, 224-byte contiguous row stores.
#include "bma.cuh"

__global__ void __launch_bounds__(128) k_bma_elements(int64_t nT, const double* __restrict__ xy, int64_t nN,
                                                      const int64_t* __restrict__ tris, const int64_t* __restrict__ edges,
                                                      int64_t nE, const int64_t* __restrict__ t2e, const cx* __restrict__ er,
                                                      const cx* __restrict__ ur, double k0, cx* __restrict__ A,
                                                      cx* __restrict__ B, int* __restrict__ bad) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t t = g >> 4;
    const int r = (int)(g & 15);
    if (t >= nT || r >= 14) return;
    double p[3][2];
    int64_t v[3];
    for (int k = 0; k < 3; ++k) {
        v[k] = tris[k * nT + t];
        p[k][0] = xy[v[k]];
        p[k][1] = xy[nN + v[k]];
    }
    int lmap[3][2];
    for (int e = 0; e < 3; ++e) {
        const int64_t eid = t2e[e * nT + t];
        const int64_t g0 = edges[eid], g1 = edges[nE + eid];
        lmap[e][0] = g0 == v[0] ? 0 : (g0 == v[1] ? 1 : (g0 == v[2] ? 2 : -1));
        lmap[e][1] = g1 == v[0] ? 0 : (g1 == v[1] ? 1 : (g1 == v[2] ? 2 : -1));
        if (lmap[e][0] < 0 || lmap[e][1] < 0) { atomicExch(bad, 1); return; }
    }
    cx mu[3][3], ep[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { mu[i][j] = ur[(i * 3 + j) * nT + t]; ep[i][j] = er[(i * 3 + j) * nT + t]; }
    bma::TriData d;
    bma::tri_setup(p, lmap, mu, ep, d);
    bma::element_row(d, k0, r, A + (t * 14 + r) * 14, B + (t * 14 + r) * 14);
}

extern "C" int emb_bma_element_matrices(emb_ctx* c, int64_t n_tris, int64_t n_nodes, int64_t n_edges, const double* xy_2xn,
                                        const int64_t* tris_3xnt, const int64_t* edges_2xne, const int64_t* tri_to_edge_3xnt,
                                        const emb_c128* er_3x3xnt, const emb_c128* ur_3x3xnt, double k0, emb_c128* A_ntx14x14,
                                        emb_c128* B_ntx14x14) {
    if (!c || n_tris <= 0 || n_nodes <= 0 || n_edges <= 0 || !xy_2xn || !tris_3xnt || !edges_2xne || !tri_to_edge_3xnt ||
        !er_3x3xnt || !ur_3x3xnt || !A_ntx14x14 || !B_ntx14x14)
        return EMB_ERR_ARG;
    for (int64_t i = 0; i < 3 * n_tris; ++i)
        if (tris_3xnt[i] < 0 || tris_3xnt[i] >= n_nodes || tri_to_edge_3xnt[i] < 0 || tri_to_edge_3xnt[i] >= n_edges) {
            c->err = "emb_bma_element_matrices: vertex or edge index out of range";
            return EMB_ERR_ARG;
        }
    PhaseTimer pt(c, "bma_elements");
    DevBuf<double> dxy;
    DevBuf<int64_t> dt, de, dte;
    DevBuf<cx> der, dur, dA, dB;
    DevBuf<int> bad;
    EMB_TRY(h2d(c, dxy, xy_2xn, (size_t)n_nodes * 2));
    EMB_TRY(h2d(c, dt, tris_3xnt, (size_t)n_tris * 3));
    EMB_TRY(h2d(c, de, edges_2xne, (size_t)n_edges * 2));
    EMB_TRY(h2d(c, dte, tri_to_edge_3xnt, (size_t)n_tris * 3));
    EMB_TRY(h2d(c, der, reinterpret_cast<const cx*>(er_3x3xnt), (size_t)n_tris * 9));
    EMB_TRY(h2d(c, dur, reinterpret_cast<const cx*>(ur_3x3xnt), (size_t)n_tris * 9));
    EMB_TRY(dev_alloc(c, dA, (size_t)n_tris * 196));
    EMB_TRY(dev_alloc(c, dB, (size_t)n_tris * 196));
    EMB_TRY(dev_alloc(c, bad, 1));
    EMB_CUDA(c, cudaMemsetAsync(bad.p, 0, sizeof(int), c->stream));
    k_bma_elements<<<blocks_for(n_tris * 16, 128), 128, 0, c->stream>>>(n_tris, dxy.p, n_nodes, dt.p, de.p, n_edges, dte.p, der.p,
                                                                       dur.p, k0, dA.p, dB.p, bad.p);
    EMB_LAUNCH_CHECK(c);
    int hbad = 0;
    EMB_CUDA(c, cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaMemcpyAsync(A_ntx14x14, dA.p, (size_t)n_tris * 196 * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaMemcpyAsync(B_ntx14x14, dB.p, (size_t)n_tris * 196 * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    dxy.release(); dt.release(); de.release(); dte.release(); der.release(); dur.release(); dA.release(); dB.release(); bad.release();
    if (hbad) { c->err = "emb_bma_element_matrices: an edge of tri_to_edge is not an edge of its triangle"; return EMB_ERR_ARG; }
    return EMB_OK;
}

// ---- shift-invert operator of the port eigenproblem -----------------------------------------------------------------
// The reference's direct path hands dense A, B to scipy.linalg.eig (QZ, O(25 n^3) on one core, fem/solver.py:339) and its
// iterative path lets ARPACK factorise A - sigma B with SuperLU (eigsh(..., sigma), :356).  Here the operator
//     v -> (A - sigma B)^-1 B v
// lives on the device: S = A - sigma B is inverted in place by Gauss-Jordan elimination with partial (row) pivoting -
// three launches per elimination step (pivot search; row swap + pivot row + multiplier column; rank-1 update of the whole
// matrix, HBM-bound: 32 n^2 bytes per step) - and an application is two dense matrix-vector products.  The Krylov
// iteration itself (implicitly restarted Arnoldi on vectors of the port size) stays in the caller (emerge_b200/modal.py).
constexpr int GJ_T = 256;

__global__ void __launch_bounds__(GJ_T) k_gj_pivot(int n, int k, const cx* __restrict__ S, int* __restrict__ piv, cx* __restrict__ pval) {
    __shared__ double sv[GJ_T];
    __shared__ int si[GJ_T];
    double best = -1.0;
    int bi = k;
    for (int i = k + threadIdx.x; i < n; i += GJ_T) {
        const cx v = S[(size_t)i * n + k];
        const double m = v.re * v.re + v.im * v.im;
        if (m > best) { best = m; bi = i; }
    }
    sv[threadIdx.x] = best; si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = GJ_T / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double o = sv[threadIdx.x + s];
            const int oi = si[threadIdx.x + s];
            if (o > sv[threadIdx.x] || (o == sv[threadIdx.x] && oi < si[threadIdx.x])) { sv[threadIdx.x] = o; si[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { piv[k] = si[0]; *pval = S[(size_t)si[0] * n + k]; }
}
// thread j: moves row k into row p (column j), forms the scaled pivot row; thread i = j: multiplier of row i
__global__ void k_gj_rows(int n, int k, cx* __restrict__ S, const int* __restrict__ piv, const cx* __restrict__ pval,
                          cx* __restrict__ rowk, cx* __restrict__ f) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int p = piv[k];
    const cx ip = cdiv(mk(1.0), *pval);
    const cx a = S[(size_t)k * n + j], b = S[(size_t)p * n + j];
    // multiplier of row j (column k) BEFORE the swap touches column k: row p receives the old row k
    cx fj = S[(size_t)j * n + k];
    if (j == p) fj = S[(size_t)k * n + k];
    f[j] = fj;
    rowk[j] = (j == k) ? ip : b * ip;
    if (p != k) S[(size_t)p * n + j] = a;
}
__global__ void __launch_bounds__(256) k_gj_update(int n, int k, cx* __restrict__ S, const cx* __restrict__ rowk, const cx* __restrict__ f) {
    const int j = blockIdx.x * 64 + (threadIdx.x & 63);
    const int i0 = blockIdx.y * 16 + (threadIdx.x >> 6) * 4;
    if (j >= n) return;
    const cx r = rowk[j];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = i0 + u;
        if (i >= n) break;
        cx* s = S + (size_t)i * n + j;
        if (i == k) { *s = r; continue; }
        const cx fi = f[i];
        if (j == k) { *s = -(fi * r); continue; }
        *s = *s - fi * r;
    }
}
__global__ void k_gj_unpermute(int n, const cx* __restrict__ S, const int* __restrict__ colsrc, cx* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j < n) out[(size_t)i * n + j] = S[(size_t)i * n + colsrc[j]];
}
__global__ void k_shift(size_t nn, const cx* __restrict__ A, const cx* __restrict__ B, cx sigma, cx* __restrict__ S) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < nn) S[i] = A[i] - sigma * B[i];
}
// y = M x, one warp per row
__global__ void k_dense_matvec(int n, const cx* __restrict__ M, const cx* __restrict__ x, cx* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    double re = 0.0, im = 0.0;
    for (int j = lane; j < n; j += 32) {
        const cx m = M[(size_t)i * n + j], v = x[j];
        re += m.re * v.re - m.im * v.im;
        im += m.re * v.im + m.im * v.re;
    }
    for (int o = 16; o > 0; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
    if (lane == 0) y[i] = cx{re, im};
}

struct ShiftInvert {
    int n = 0;
    DevBuf<cx> Minv, B, v, w;
};
static void si_free(emb_ctx* c) {
    ShiftInvert* s = static_cast<ShiftInvert*>(c->shift_invert);
    if (!s) return;
    s->Minv.release(); s->B.release(); s->v.release(); s->w.release();
    delete s;
    c->shift_invert = nullptr;
}

extern "C" int emb_shift_invert_free(emb_ctx* c) {
    if (!c) return EMB_ERR_ARG;
    si_free(c);
    return EMB_OK;
}

extern "C" int emb_shift_invert_setup(emb_ctx* c, int64_t n64, const emb_c128* A_nxn, const emb_c128* B_nxn, double sigma_re,
                                      double sigma_im) {
    if (!c || n64 <= 0 || n64 > 46000 || !A_nxn || !B_nxn) return EMB_ERR_ARG;
    const int n = (int)n64;
    const size_t nn = (size_t)n * n;
    si_free(c);
    PhaseTimer pt(c, "shift_invert_setup");
    ShiftInvert* s = new ShiftInvert();
    c->shift_invert = s;
    s->n = n;
    DevBuf<cx> A, S, rowk, f, pval;
    DevBuf<int> piv, colsrc;
    EMB_TRY(h2d(c, A, reinterpret_cast<const cx*>(A_nxn), nn));
    EMB_TRY(h2d(c, s->B, reinterpret_cast<const cx*>(B_nxn), nn));
    EMB_TRY(dev_alloc(c, S, nn));
    EMB_TRY(dev_alloc(c, rowk, (size_t)n));
    EMB_TRY(dev_alloc(c, f, (size_t)n));
    EMB_TRY(dev_alloc(c, pval, 1));
    EMB_TRY(dev_alloc(c, piv, (size_t)n));
    EMB_TRY(dev_alloc(c, colsrc, (size_t)n));
    EMB_TRY(dev_alloc(c, s->v, (size_t)n));
    EMB_TRY(dev_alloc(c, s->w, (size_t)n));
    k_shift<<<blocks_for((int64_t)nn, 256), 256, 0, c->stream>>>(nn, A.p, s->B.p, cx{sigma_re, sigma_im}, S.p);
    EMB_LAUNCH_CHECK(c);
    const dim3 ug((unsigned)((n + 63) / 64), (unsigned)((n + 15) / 16));
    for (int k = 0; k < n; ++k) {
        k_gj_pivot<<<1, GJ_T, 0, c->stream>>>(n, k, S.p, piv.p, pval.p);
        k_gj_rows<<<blocks_for(n, 256), 256, 0, c->stream>>>(n, k, S.p, piv.p, pval.p, rowk.p, f.p);
        k_gj_update<<<ug, 256, 0, c->stream>>>(n, k, S.p, rowk.p, f.p);
    }
    EMB_LAUNCH_CHECK(c);
    c->launches += 3 * (int64_t)n;
    // the row swaps of the elimination become column swaps of the inverse, undone in reverse order: one gather
    std::vector<int> hp((size_t)n), src((size_t)n);
    EMB_CUDA(c, cudaMemcpyAsync(hp.data(), piv.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int j = 0; j < n; ++j) src[(size_t)j] = j;
    for (int k = n - 1; k >= 0; --k)
        if (hp[(size_t)k] != k) std::swap(src[(size_t)k], src[(size_t)hp[(size_t)k]]);
    EMB_CUDA(c, cudaMemcpyAsync(colsrc.p, src.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    A.release();
    EMB_TRY(dev_alloc(c, s->Minv, nn));
    k_gj_unpermute<<<dim3((unsigned)blocks_for(n, 256), (unsigned)n), 256, 0, c->stream>>>(n, S.p, colsrc.p, s->Minv.p);
    EMB_LAUNCH_CHECK(c);
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    S.release(); rowk.release(); f.release(); pval.release(); piv.release(); colsrc.release();
    return EMB_OK;          // a singular A - sigma B shows up as non-finite entries of the first application
}

extern "C" int emb_shift_invert_apply(emb_ctx* c, const emb_c128* x, emb_c128* y) {
    if (!c || !x || !y) return EMB_ERR_ARG;
    ShiftInvert* s = static_cast<ShiftInvert*>(c->shift_invert);
    if (!s) { c->err = "emb_shift_invert_apply: emb_shift_invert_setup not called"; return EMB_ERR_STATE; }
    const int n = s->n;
    EMB_CUDA(c, cudaMemcpyAsync(s->v.p, x, (size_t)n * sizeof(cx), cudaMemcpyHostToDevice, c->stream));
    k_dense_matvec<<<blocks_for((int64_t)n * 32, 256), 256, 0, c->stream>>>(n, s->B.p, s->v.p, s->w.p);
    k_dense_matvec<<<blocks_for((int64_t)n * 32, 256), 256, 0, c->stream>>>(n, s->Minv.p, s->w.p, s->v.p);
    EMB_LAUNCH_CHECK(c);
    c->launches += 2;
    EMB_CUDA(c, cudaMemcpyAsync(y, s->v.p, (size_t)n * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    return EMB_OK;
}
