// Reduced-basis recycling across the frequency points of a sweep.
//
// The sweep solves A(f) x = b_p(f) for a dense list of f.  A(f) = K - k0^2 M + sum_p gamma_p(f) S_p is AFFINE in T fixed
// matrices W_t (K, M and the port / absorbing-boundary matrices S_p), so for the solution directions U kept from
// earlier solves the products W_t U never change.  They are computed once per direction (T SpMVs) and held as
//      W_t U = Q R_t        Q: orthonormal columns in HBM (classical Gram-Schmidt, two passes, batched dots),
//                           R_t: small coefficient matrices on the host.
// At a new frequency  A(f) U = Q G(f)  with  G(f) = sum_t coef_t(f) R_t  - no SpMV and no orthogonalisation on the device.
// The start vector of a solve is the minimum-residual combination over span(U):
//      y = argmin || Q^H r0 - G(f) y ||  (Householder QR of the small matrix on the host),   x0 += U y,
// which costs one pass over Q (batched dots for all ports at once) and one pass over U.  The residual of x0 is then
// recomputed from A(f) itself in FP64 - the accuracy contract (relres <= rtol on the true operator) does not depend on
// anything in this file.  A point whose start vector already meets rtol is accepted without iterating; a point that
// iterates contributes its correction as a new direction.
// The reference has no counterpart: it factorises A(f) at every point (fem/solver.py:243-309, emfreq3d.py:658-694).
#pragma once
#include "krylov.cuh"
#include <cub/cub.cuh>
#include "lsq.hpp"
constexpr int RC_NP = 256;       // partials per dot product of the batched kernels (== VBLOCK)

// ---- W_t x on the solve space ------------------------------------------------------------------------------
// y = F_s x where F is K or M on the FULL pattern and src maps solve-space entries to full-pattern slots
__global__ void __launch_bounds__(256) k_spmv_src(int64_t n, const int64_t* __restrict__ rowptr, const int* __restrict__ col,
                                                  const int64_t* __restrict__ src, const cx* __restrict__ F,
                                                  const cx* __restrict__ x, cx* __restrict__ y) {
    constexpr int LPR = 8;
    const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t r = gt / LPR;
    const int sub = (int)(gt % LPR);
    double ar = 0, ai = 0;
    if (r < n)
        for (int64_t k = rowptr[r] + sub; k < rowptr[r + 1]; k += LPR) {
            const cx a = ldx(F + __ldg(src + k));
            const cx w = ldx(x + __ldg(col + k));
            ar += a.re * w.re - a.im * w.im;
            ai += a.re * w.im + a.im * w.re;
        }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
        ar += __shfl_down_sync(0xffffffffu, ar, o, LPR);
        ai += __shfl_down_sync(0xffffffffu, ai, o, LPR);
    }
    if (r < n && sub == 0) y[r] = cx{ar, ai};
}
// y += S x for a surface matrix held as (solve-pattern slot, real value) pairs SORTED by slot (eliminated entries, slot
// -1, first); y is zeroed by the caller.  The thread that owns the first entry of a row sums the whole row in list order (deterministic).
__device__ __forceinline__ int64_t row_of_slot(const int64_t* __restrict__ rowptr, int64_t n, int64_t p) {
    int64_t lo = 0, hi = n - 1;          // last r with rowptr[r] <= p
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (rowptr[mid] <= p) lo = mid; else hi = mid - 1;
    }
    return lo;
}
__global__ void k_surf_mv(int64_t nslot, const int64_t* __restrict__ slot_s, const double* __restrict__ Sval, int64_t n,
                          const int64_t* __restrict__ rowptr, const int* __restrict__ col, const cx* __restrict__ x,
                          cx* __restrict__ y) {
    const int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u >= nslot) return;
    const int64_t p = slot_s[u];
    if (p < 0) return;
    const int64_t r = row_of_slot(rowptr, n, p);
    for (int64_t w = u - 1; w >= 0; --w) {       // previous kept entry in the same row => not the owner
        const int64_t q = slot_s[w];
        if (q < 0) continue;
        if (q >= rowptr[r]) return;
        break;
    }
    const int64_t pend = rowptr[r + 1];
    double ar = 0, ai = 0;
    for (int64_t w = u; w < nslot; ++w) {
        const int64_t q = slot_s[w];
        if (q < 0) continue;
        if (q >= pend) break;
        const double s = Sval[w];
        const cx v = x[col[q]];
        ar += s * v.re;
        ai += s * v.im;
    }
    y[r] = cx{ar, ai};
}

// ---- batched dots and combinations ---------------------------------------------------------------------------
__device__ __forceinline__ cx block_sum1(cx v) {
    __shared__ double s_re[VBLOCK / 32], s_im[VBLOCK / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.re += __shfl_down_sync(0xffffffffu, v.re, o);
        v.im += __shfl_down_sync(0xffffffffu, v.im, o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { s_re[w] = v.re; s_im[w] = v.im; }
    __syncthreads();
    cx out = mk(0.0);
    if (threadIdx.x == 0)
        for (int i = 0; i < VBLOCK / 32; ++i) { out.re += s_re[i]; out.im += s_im[i]; }
    return out;   // valid in thread 0
}
// part[(j * RC_NP + blockIdx.x) * NV + v] = partial of <q_j, r_v> over this block's rows;  j = blockIdx.y;
// Q: contiguous columns of length n, r: NV interleaved columns
// CONJ = false: unconjugated products q_j^T r_v
template <int NV, bool CONJ = true, typename QT = cx>
__global__ void __launch_bounds__(VBLOCK) k_rc_dots(int64_t n, const QT* __restrict__ Q, const cx* __restrict__ r,
                                                    cx* __restrict__ part) {
    // blocks are scheduled x-fastest: remap so that consecutive blocks share the ROW RANGE p and differ in the column j -
    // the range of r is then fetched from HBM once and served from L2 to the other columns (launched as (RC_NP, ncols))
    const int lin = blockIdx.y * gridDim.x + blockIdx.x;
    const int bj = lin % (int)gridDim.y, bp = lin / (int)gridDim.y;
    const QT* q = Q + (int64_t)bj * n;
    cx acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = mk(0.0);
    const int64_t per = (n + RC_NP - 1) / RC_NP;
    const int64_t i0 = bp * per, i1 = (i0 + per < n) ? i0 + per : n;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += VBLOCK) {
        const cx u = ldv(q, i);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const cx w = r[i * NV + v];
            if (CONJ) {
                acc[v].re += u.re * w.re + u.im * w.im;
                acc[v].im += u.re * w.im - u.im * w.re;
            } else {
                fma_c(acc[v], u, w);
            }
        }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const cx t = block_sum1(acc[v]);
        if (threadIdx.x == 0) part[((int64_t)bj * RC_NP + bp) * NV + v] = t;
    }
}
// coef[j * NV + v] = sum of the RC_NP partials in fixed order (one block per column j, RC_NP == VBLOCK threads)
template <int NV>
__global__ void __launch_bounds__(VBLOCK) k_rc_coef(const cx* __restrict__ part, cx* __restrict__ coef) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const cx t = block_sum1(part[((int64_t)blockIdx.x * RC_NP + threadIdx.x) * NV + v]);
        if (threadIdx.x == 0) coef[blockIdx.x * NV + v] = t;
    }
}
// w -= sum_{j<m} coef[j] q_j      (contiguous w)
__global__ void __launch_bounds__(256) k_rc_sub(int64_t n, int m, const cx* __restrict__ coef, const cx* __restrict__ Q,
                                                cx* __restrict__ w) {
    extern __shared__ cx s_h[];
    for (int k = threadIdx.x; k < m; k += blockDim.x) s_h[k] = coef[k];
    __syncthreads();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    cx a = w[i];
    for (int k = 0; k < m; ++k) fma_c(a, -s_h[k], Q[(int64_t)k * n + i]);
    w[i] = a;
}
// The same two kernels for NB CONTIGUOUS columns w_v = W + v n at once (the T products of a new direction): Q is read
// once for all of them.   part[(j * RC_NP + p) * NB + v] = partial of <q_j, w_v>;   w_v -= sum_{j<m} coef[j * NB + v] q_j
template <int NB, bool CONJ = true>
__global__ void __launch_bounds__(VBLOCK) k_rc_dots_cols(int64_t n, const cx* __restrict__ Q, const cx* __restrict__ W,
                                                         cx* __restrict__ part) {
    const int lin = blockIdx.y * gridDim.x + blockIdx.x;
    const int bj = lin % (int)gridDim.y, bp = lin / (int)gridDim.y;       // consecutive blocks share the row range (see k_rc_dots)
    const cx* q = Q + (int64_t)bj * n;
    cx acc[NB];
#pragma unroll
    for (int v = 0; v < NB; ++v) acc[v] = mk(0.0);
    const int64_t per = (n + RC_NP - 1) / RC_NP;
    const int64_t i0 = bp * per, i1 = (i0 + per < n) ? i0 + per : n;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += VBLOCK) {
        const cx u = q[i];
#pragma unroll
        for (int v = 0; v < NB; ++v) {
            const cx w = W[(int64_t)v * n + i];
            if (CONJ) {
                acc[v].re += u.re * w.re + u.im * w.im;
                acc[v].im += u.re * w.im - u.im * w.re;
            } else {
                fma_c(acc[v], u, w);
            }
        }
    }
#pragma unroll
    for (int v = 0; v < NB; ++v) {
        const cx t = block_sum1(acc[v]);
        if (threadIdx.x == 0) part[((int64_t)bj * RC_NP + bp) * NB + v] = t;
    }
}
template <int NB>
__global__ void __launch_bounds__(256) k_rc_sub_cols(int64_t n, int m, const cx* __restrict__ coef, const cx* __restrict__ Q,
                                                     cx* __restrict__ W) {
    extern __shared__ cx s_h[];
    for (int k = threadIdx.x; k < m * NB; k += blockDim.x) s_h[k] = coef[k];
    __syncthreads();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    cx a[NB];
#pragma unroll
    for (int v = 0; v < NB; ++v) a[v] = W[(int64_t)v * n + i];
    for (int k = 0; k < m; ++k) {
        const cx q = Q[(int64_t)k * n + i];
#pragma unroll
        for (int v = 0; v < NB; ++v) fma_c(a[v], -s_h[k * NB + v], q);
    }
#pragma unroll
    for (int v = 0; v < NB; ++v) W[(int64_t)v * n + i] = a[v];
}
// x[i][v] += sum_{j<m} y[j * NV + v] u_j[i]
template <int NV, typename UT = cx>
__global__ void __launch_bounds__(256) k_rc_combine(int64_t n, int m, const cx* __restrict__ y, const UT* __restrict__ U,
                                                    cx* __restrict__ x) {
    extern __shared__ cx s_h[];
    for (int k = threadIdx.x; k < m * NV; k += blockDim.x) s_h[k] = y[k];
    __syncthreads();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    cx a[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) a[v] = x[i * NV + v];
    for (int k = 0; k < m; ++k) {
        const cx u = ldv(U, (int64_t)k * n + i);
#pragma unroll
        for (int v = 0; v < NV; ++v) fma_c(a[v], s_h[k * NV + v], u);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) x[i * NV + v] = a[v];
}
__global__ void k_scale_real(int64_t n, double s, cx* __restrict__ w) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) w[i] = s * w[i];
}

// ---- state ----------------------------------------------------------------------------------------------------
static inline int rc_T(const emb_ctx* c) { return 2 + (int)c->rc_terms.size(); }
static inline cx* rc_U(emb_ctx* c, int j) { return c->rcU.p + (int64_t)j * c->Ns; }
static inline cx* rc_Q(emb_ctx* c, int j) { return c->rcQ.p + (int64_t)j * c->Ns; }

static void rc_clear(emb_ctx* c) {
    c->rc_n = 0;
    c->rc_nq = 0;
    c->rc_version++;
    c->coarse_m = 0;
    for (auto& R : c->rc_R) std::fill(R.begin(), R.end(), zc(0.0, 0.0));
}

// (re)allocates U, Q and the coefficient matrices for the affine terms of the current A(f)
static int rc_prepare(emb_ctx* c) {
    if (c->rc_cap <= 0) return EMB_OK;
    const bool same_terms = c->rc_terms == c->aff_sids && c->rcU.p && c->rcQ.p;
    if (same_terms) return EMB_OK;
    c->rc_terms = c->aff_sids;
    const int T = rc_T(c);
    // the basis costs (1 + T) vectors per direction: never take more than half of the free HBM (5M tets: ~10 directions
    // fewer than asked for rather than an allocation failure in the middle of a sweep)
    c->rcU.release(); c->rcU32.release(); c->rcQ.release();
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
        const double per_dir = ((double)(1 + T) + (c->coarse_basis ? 0.5 : 0.0)) * (double)c->Ns * sizeof(cx);
        const int fit = (int)(0.5 * (double)free_b / per_dir);
        if (fit < c->rc_cap) c->rc_cap = fit > 0 ? fit : 0;
    }
    if (c->rc_cap <= 0) { c->rc_cap = 0; return EMB_OK; }
    c->rc_qcap = T * c->rc_cap;
    EMB_TRY(dev_alloc(c, c->rcU, (size_t)c->rc_cap * c->Ns));
    EMB_TRY(dev_alloc(c, c->rcQ, (size_t)c->rc_qcap * c->Ns));
    EMB_TRY(dev_alloc(c, c->rc_part, (size_t)c->rc_qcap * RC_NP * NVMAX + (size_t)(3 * c->rc_qcap + c->rc_cap + 8) * NVMAX));
    EMB_TRY(dev_alloc(c, c->rc_tmp, (size_t)c->Ns));
    // the coefficient area is copied to the host as a whole (rc_insert) while only the first nq entries are written:
    // defined contents keep compute-sanitizer --tool initcheck clean
    EMB_CUDA(c, cudaMemsetAsync(c->rc_part.p, 0, c->rc_part.n * sizeof(cx), c->stream));
    c->rc_R.assign((size_t)T, std::vector<zc>((size_t)c->rc_qcap * c->rc_cap, zc(0.0, 0.0)));
    c->rc_uscale.assign((size_t)c->rc_cap, 1.0);
    if (c->coarse_basis) {
        c->rc_UtQ.assign((size_t)c->rc_cap * c->rc_qcap, zc(0.0, 0.0));
        c->rc_UhU.assign((size_t)c->rc_cap * c->rc_cap, zc(0.0, 0.0));
        EMB_TRY(dev_alloc(c, c->rc_ceff, (size_t)c->rc_cap * c->rc_cap));
        EMB_TRY(dev_alloc(c, c->rcU32, (size_t)c->rc_cap * c->Ns * 2));
        EMB_TRY(dev_alloc(c, c->rc_ct, (size_t)2 * c->rc_cap * NVMAX));
    }
    rc_clear(c);
    return EMB_OK;
}
static inline cx* rc_coef_area(emb_ctx* c) { return c->rc_part.p + (size_t)c->rc_qcap * RC_NP * NVMAX; }

// w = W_t u   (t = 0: K, 1: M, >= 2: surface rc_terms[t-2]) on the solve space; contiguous vectors
static int rc_term_mv(emb_ctx* c, int t, const cx* u, cx* w) {
    const int64_t n = c->Ns;
    if (t < 2) {
        k_spmv_src<<<blocks_for(n * 8, 256), 256, 0, c->stream>>>(n, c->rowptr_s.p, c->col_s.p, c->src.p, t == 0 ? c->K.p : c->M.p, u, w);
        EMB_LAUNCH_CHECK(c);
    } else {
        Surface& s = c->surf[c->rc_terms[(size_t)t - 2]];
        if (!s.mv_slot.p) {      // once per surface: entries in solve-pattern order (the list is in full-pattern order)
            EMB_TRY(dev_alloc(c, s.mv_slot, (size_t)s.nslot));
            EMB_TRY(dev_alloc(c, s.mv_val, (size_t)s.nslot));
            DevBuf<char> tmp;
            size_t tb = 0;
            EMB_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tb, s.slot_s.p, s.mv_slot.p, s.Sval.p, s.mv_val.p, (int)s.nslot, 0, 64, c->stream));
            EMB_TRY(dev_alloc(c, tmp, tb));
            EMB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tb, s.slot_s.p, s.mv_slot.p, s.Sval.p, s.mv_val.p, (int)s.nslot, 0, 64, c->stream));
            EMB_CUDA(c, cudaStreamSynchronize(c->stream));
            tmp.release();
            c->launches += 3;
        }
        EMB_CUDA(c, cudaMemsetAsync(w, 0, (size_t)n * sizeof(cx), c->stream));
        k_surf_mv<<<blocks_for(s.nslot, 128), 128, 0, c->stream>>>(s.nslot, s.mv_slot.p, s.mv_val.p, n, c->rowptr_s.p, c->col_s.p, u, w);
        EMB_LAUNCH_CHECK(c);
    }
    c->rc_spmvs++;
    return EMB_OK;
}

// G(f) = sum_t coef_t R_t restricted to the first m directions (nq x m, column-major)
static void rc_build_G(const emb_ctx* c, int m, std::vector<zc>& G) {
    const int nq = c->rc_nq, T = rc_T(c);
    G.assign((size_t)nq * m, zc(0.0, 0.0));
    for (int t = 0; t < T; ++t) {
        const zc ct = c->aff_coef[(size_t)t];
        const std::vector<zc>& R = c->rc_R[(size_t)t];
        for (int j = 0; j < m; ++j)
            for (int i = 0; i < nq; ++i) G[(size_t)j * nq + i] += ct * R[(size_t)j * c->rc_qcap + i];
    }
}

// U[slot] holds a new direction: compute its T products, extend Q, fill column `slot` of every R_t.
// accept_test: keep the direction only if A(f) u is not (numerically) inside span(A(f) U[0..slot)).
static int rc_insert(emb_ctx* c, int slot, bool accept_test, bool* accepted) {
    const int64_t n = c->Ns;
    const int T = rc_T(c);
    const unsigned vb = blocks_for(n, 256);
    cx* part = c->rc_part.p;
    const int qc = c->rc_qcap;
    // coefficient area (12 qc + .. entries): block coefficients [qc * 4], in-block coefficients of the two passes [2][qc] and
    // a squared norm, squared norms of the raw products [T]
    cx* coef = rc_coef_area(c);
    cx* hB = coef;
    cx* hC = coef + (size_t)8 * qc;
    cx* nrm = coef + (size_t)10 * qc;
    const int nq0 = c->rc_nq;                    // columns of Q before this direction
    if (nq0 + T > qc) { c->err = "recycling: orthonormal basis is full"; return EMB_ERR_LIMIT; }
    // A: the T products W_t u, side by side behind the basis
    for (int t = 0; t < T; ++t) {
        cx* w = rc_Q(c, nq0 + t);
        EMB_TRY(rc_term_mv(c, t, rc_U(c, slot), w));
        k_dot<1, true><<<NPART, VBLOCK, 0, c->stream>>>(n, w, w, part); EMB_LAUNCH_CHECK(c);
        k_finish<1><<<1, VBLOCK, 0, c->stream>>>(part, nrm + t); EMB_LAUNCH_CHECK(c);
    }
    // Block classical Gram-Schmidt with re-orthogonalisation (BCGS2) of the T products against the basis:
    //   W = Qold H1 + W1,  W1 = Q1 R1 (in-block CGS2; a product inside the span of the others is dropped),
    //   Q1 = Qold H2 + Q1', Q1' = Qnew R2 (in-block CGS2)   =>   W = Qold (H1 + H2 R1) + Qnew (R2 R1).
    // Qold is read once per pass for up to four columns instead of twice per product.  (One block pass followed by the
    // in-block step only is NOT enough: K u and M u are nearly parallel, the in-block cancellation amplifies the rounding
    // left along Qold and the basis loses orthogonality within a few directions - measured, tools/rc_ab.py.)
    auto proj_old = [&](int first, int ncols, std::vector<zc>& H) -> int {      // H[v * nq0 + i] += <q_i, col(first + v)>
        for (int g0 = 0; g0 < ncols && nq0 > 0; g0 += 4) {
            const int gs = ncols - g0 < 4 ? ncols - g0 : 4;
            cx* W = rc_Q(c, first + g0);
            const dim3 grid(RC_NP, nq0);
            const size_t sh = (size_t)nq0 * gs * sizeof(cx);
            switch (gs) {
                case 1: k_rc_dots_cols<1><<<grid, VBLOCK, 0, c->stream>>>(n, c->rcQ.p, W, part); EMB_LAUNCH_CHECK(c);
                        k_rc_coef<1><<<nq0, VBLOCK, 0, c->stream>>>(part, hB); EMB_LAUNCH_CHECK(c);
                        k_rc_sub_cols<1><<<vb, 256, sh, c->stream>>>(n, nq0, hB, c->rcQ.p, W); EMB_LAUNCH_CHECK(c); break;
                case 2: k_rc_dots_cols<2><<<grid, VBLOCK, 0, c->stream>>>(n, c->rcQ.p, W, part); EMB_LAUNCH_CHECK(c);
                        k_rc_coef<2><<<nq0, VBLOCK, 0, c->stream>>>(part, hB); EMB_LAUNCH_CHECK(c);
                        k_rc_sub_cols<2><<<vb, 256, sh, c->stream>>>(n, nq0, hB, c->rcQ.p, W); EMB_LAUNCH_CHECK(c); break;
                case 3: k_rc_dots_cols<3><<<grid, VBLOCK, 0, c->stream>>>(n, c->rcQ.p, W, part); EMB_LAUNCH_CHECK(c);
                        k_rc_coef<3><<<nq0, VBLOCK, 0, c->stream>>>(part, hB); EMB_LAUNCH_CHECK(c);
                        k_rc_sub_cols<3><<<vb, 256, sh, c->stream>>>(n, nq0, hB, c->rcQ.p, W); EMB_LAUNCH_CHECK(c); break;
                default: k_rc_dots_cols<4><<<grid, VBLOCK, 0, c->stream>>>(n, c->rcQ.p, W, part); EMB_LAUNCH_CHECK(c);
                        k_rc_coef<4><<<nq0, VBLOCK, 0, c->stream>>>(part, hB); EMB_LAUNCH_CHECK(c);
                        k_rc_sub_cols<4><<<vb, 256, sh, c->stream>>>(n, nq0, hB, c->rcQ.p, W); EMB_LAUNCH_CHECK(c); break;
            }
            std::vector<cx> hh((size_t)nq0 * gs);
            EMB_CUDA(c, cudaMemcpyAsync(hh.data(), hB, hh.size() * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
            EMB_CUDA(c, cudaStreamSynchronize(c->stream));
            for (int v = 0; v < gs; ++v)
                for (int i = 0; i < nq0; ++i)
                    H[(size_t)(g0 + v) * nq0 + i] += zc(hh[(size_t)i * gs + v].re, hh[(size_t)i * gs + v].im);
        }
        return EMB_OK;
    };
    // in-block CGS2 of `ncols` columns (column v of the input sits at slot nq0 + v); accepted columns end up dense at
    // nq0, nq0 + 1, ...; R[v * T + k] = coefficient of input column v on accepted column k; n0 (may be null): norms the
    // drop test refers to.  Returns the number of accepted columns in *na.
    auto in_block = [&](int ncols, const double* n0, std::vector<zc>& R, int* na) -> int {
        int acc = 0;
        std::vector<cx> host((size_t)2 * qc + 1);
        for (int v = 0; v < ncols; ++v) {
            cx* w = rc_Q(c, nq0 + acc);
            if (acc != v)            // an earlier column was dropped: keep the columns dense
                EMB_CUDA(c, cudaMemcpyAsync(w, rc_Q(c, nq0 + v), (size_t)n * sizeof(cx), cudaMemcpyDeviceToDevice, c->stream));
            for (int pass = 0; pass < 2 && acc > 0; ++pass) {
                cx* h = hC + (size_t)pass * qc;
                k_rc_dots<1><<<dim3(RC_NP, acc), VBLOCK, 0, c->stream>>>(n, rc_Q(c, nq0), w, part); EMB_LAUNCH_CHECK(c);
                k_rc_coef<1><<<acc, VBLOCK, 0, c->stream>>>(part, h); EMB_LAUNCH_CHECK(c);
                k_rc_sub<<<vb, 256, acc * sizeof(cx), c->stream>>>(n, acc, h, rc_Q(c, nq0), w); EMB_LAUNCH_CHECK(c);
            }
            k_dot<1, true><<<NPART, VBLOCK, 0, c->stream>>>(n, w, w, part); EMB_LAUNCH_CHECK(c);
            k_finish<1><<<1, VBLOCK, 0, c->stream>>>(part, hC + (size_t)2 * qc); EMB_LAUNCH_CHECK(c);
            EMB_CUDA(c, cudaMemcpyAsync(host.data(), hC, host.size() * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
            EMB_CUDA(c, cudaStreamSynchronize(c->stream));
            const double n1 = std::sqrt(std::fabs(host[(size_t)2 * qc].re));
            for (int k = 0; k < acc; ++k)
                R[(size_t)v * T + k] = zc(host[(size_t)k].re + host[(size_t)qc + k].re, host[(size_t)k].im + host[(size_t)qc + k].im);
            const double ref = n0 ? n0[v] : 1.0;
            if (ref > 0 && n1 > 1e-11 * ref && n1 == n1) {
                k_scale_real<<<vb, 256, 0, c->stream>>>(n, 1.0 / n1, w); EMB_LAUNCH_CHECK(c);
                R[(size_t)v * T + acc] = zc(n1, 0.0);
                ++acc;
            }
        }
        *na = acc;
        return EMB_OK;
    };
    std::vector<cx> hn((size_t)T);
    EMB_CUDA(c, cudaMemcpyAsync(hn.data(), nrm, (size_t)T * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<double> n0((size_t)T);
    for (int t = 0; t < T; ++t) n0[(size_t)t] = std::sqrt(std::fabs(hn[(size_t)t].re));
    std::vector<zc> H1((size_t)T * nq0, zc(0.0, 0.0)), H2((size_t)T * nq0, zc(0.0, 0.0));
    std::vector<zc> R1((size_t)T * T, zc(0.0, 0.0)), R2((size_t)T * T, zc(0.0, 0.0));
    int a1 = 0, a2 = 0;
    EMB_TRY(proj_old(nq0, T, H1));
    EMB_TRY(in_block(T, n0.data(), R1, &a1));
    EMB_TRY(proj_old(nq0, a1, H2));
    EMB_TRY(in_block(a1, nullptr, R2, &a2));
    if (a2 != a1) { c->err = "recycling: re-orthogonalisation dropped a column"; return EMB_ERR_CUDA; }
    for (int t = 0; t < T; ++t) {
        std::vector<zc>& R = c->rc_R[(size_t)t];
        for (int i = 0; i < qc; ++i) R[(size_t)slot * qc + i] = zc(0.0, 0.0);
        for (int i = 0; i < nq0; ++i) {
            zc v = H1[(size_t)t * nq0 + i];
            for (int k = 0; k < a1; ++k) v += H2[(size_t)k * nq0 + i] * R1[(size_t)t * T + k];
            R[(size_t)slot * qc + i] = v;
        }
        for (int j = 0; j < a1; ++j) {          // (R2 R1)[j, t]: R2[k * T + j] = coefficient of pass-1 column k on final column j
            zc v(0.0, 0.0);
            for (int k = 0; k < a1; ++k) v += R2[(size_t)k * T + j] * R1[(size_t)t * T + k];
            R[(size_t)slot * qc + nq0 + j] = v;
        }
    }
    c->rc_nq = nq0 + a1;
    const int nq_before = nq0;
    // scale so that A(f) u has unit norm at its birth frequency; acceptance against the existing directions
    const int nq = c->rc_nq;
    std::vector<zc> cn((size_t)nq, zc(0.0, 0.0));
    for (int t = 0; t < T; ++t)
        for (int i = 0; i < nq; ++i) cn[(size_t)i] += c->aff_coef[(size_t)t] * c->rc_R[(size_t)t][(size_t)slot * qc + i];
    double cnorm = 0;
    for (auto& v : cn) cnorm += std::norm(v);
    cnorm = std::sqrt(cnorm);
    bool ok = cnorm > 0 && cnorm == cnorm;
    if (ok && accept_test && slot > 0) {
        std::vector<zc> G, y;
        std::vector<double> res;
        rc_build_G(c, slot, G);
        std::vector<zc> g = cn;
        ls_solve(nq, slot, G, 1, g, y, res);
        ok = res[0] > 1e-9 * cnorm;
    }
    if (ok) {
        const double s = 1.0 / cnorm;
        c->rc_uscale[(size_t)slot] = s;
        for (int t = 0; t < T; ++t)
            for (int i = 0; i < nq; ++i) c->rc_R[(size_t)t][(size_t)slot * qc + i] *= s;
    } else {
        for (int t = 0; t < T; ++t)
            for (int i = 0; i < qc; ++i) c->rc_R[(size_t)t][(size_t)slot * qc + i] = zc(0.0, 0.0);
    }
    if (ok && c->coarse_basis && c->rcU32.p) {       // the complex64 copy the coarse-space correction streams (a preconditioner)
        k_convert<cx, cf><<<vb, 256, 0, c->stream>>>(n, rc_U(c, slot), reinterpret_cast<cf*>(c->rcU32.p) + (int64_t)slot * n);
        EMB_LAUNCH_CHECK(c);
    }
    if (ok && c->coarse_basis && !c->rc_UtQ.empty()) {
        // bookkeeping of the experimental coarse space: u^T q for the new direction against every basis column, for the
        // new basis columns against every older direction, and the Hermitian Gram row of the new direction
        const int nq1 = c->rc_nq;
        auto fetch = [&](int cnt, std::vector<cx>& out) -> int {
            out.resize((size_t)cnt);
            EMB_CUDA(c, cudaMemcpyAsync(out.data(), coef, (size_t)cnt * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
            EMB_CUDA(c, cudaStreamSynchronize(c->stream));
            return EMB_OK;
        };
        std::vector<cx> hb;
        k_rc_dots<1, false><<<dim3(RC_NP, nq1), VBLOCK, 0, c->stream>>>(n, c->rcQ.p, rc_U(c, slot), part); EMB_LAUNCH_CHECK(c);
        k_rc_coef<1><<<nq1, VBLOCK, 0, c->stream>>>(part, coef); EMB_LAUNCH_CHECK(c);
        EMB_TRY(fetch(nq1, hb));
        for (int i = 0; i < qc; ++i) c->rc_UtQ[(size_t)slot * qc + i] = i < nq1 ? zc(hb[(size_t)i].re, hb[(size_t)i].im) : zc(0.0, 0.0);
        if (slot > 0)
            for (int i0 = nq_before; i0 < nq1; i0 += 4) {      // up to four new basis columns per pass over U
                const int gs = nq1 - i0 < 4 ? nq1 - i0 : 4;
                const dim3 grid(RC_NP, slot);
                switch (gs) {
                    case 1: k_rc_dots_cols<1, false><<<grid, VBLOCK, 0, c->stream>>>(n, c->rcU.p, rc_Q(c, i0), part); EMB_LAUNCH_CHECK(c);
                            k_rc_coef<1><<<slot, VBLOCK, 0, c->stream>>>(part, coef); EMB_LAUNCH_CHECK(c); break;
                    case 2: k_rc_dots_cols<2, false><<<grid, VBLOCK, 0, c->stream>>>(n, c->rcU.p, rc_Q(c, i0), part); EMB_LAUNCH_CHECK(c);
                            k_rc_coef<2><<<slot, VBLOCK, 0, c->stream>>>(part, coef); EMB_LAUNCH_CHECK(c); break;
                    case 3: k_rc_dots_cols<3, false><<<grid, VBLOCK, 0, c->stream>>>(n, c->rcU.p, rc_Q(c, i0), part); EMB_LAUNCH_CHECK(c);
                            k_rc_coef<3><<<slot, VBLOCK, 0, c->stream>>>(part, coef); EMB_LAUNCH_CHECK(c); break;
                    default: k_rc_dots_cols<4, false><<<grid, VBLOCK, 0, c->stream>>>(n, c->rcU.p, rc_Q(c, i0), part); EMB_LAUNCH_CHECK(c);
                            k_rc_coef<4><<<slot, VBLOCK, 0, c->stream>>>(part, coef); EMB_LAUNCH_CHECK(c); break;
                }
                EMB_TRY(fetch(slot * gs, hb));
                for (int j = 0; j < slot; ++j)
                    for (int v = 0; v < gs; ++v)
                        c->rc_UtQ[(size_t)j * qc + i0 + v] = zc(hb[(size_t)j * gs + v].re, hb[(size_t)j * gs + v].im);
            }
        k_rc_dots<1, true><<<dim3(RC_NP, slot + 1), VBLOCK, 0, c->stream>>>(n, c->rcU.p, rc_U(c, slot), part); EMB_LAUNCH_CHECK(c);
        k_rc_coef<1><<<slot + 1, VBLOCK, 0, c->stream>>>(part, coef); EMB_LAUNCH_CHECK(c);
        EMB_TRY(fetch(slot + 1, hb));
        const int cap = c->rc_cap;
        for (int j = 0; j <= slot; ++j) {                 // hb[j] = u_j^H u_slot
            c->rc_UhU[(size_t)j * cap + slot] = zc(hb[(size_t)j].re, hb[(size_t)j].im);
            c->rc_UhU[(size_t)slot * cap + j] = zc(hb[(size_t)j].re, -hb[(size_t)j].im);
        }
    }
    if (ok && getenv("EMB_RC_DEBUG")) {
        // W_t (s u) - Q R_t[:, slot] relative to W_t (s u), and the orthonormality of the new columns against the basis
        const int nq1 = c->rc_nq;
        for (int t = 0; t < T; ++t) {
            EMB_TRY(rc_term_mv(c, t, rc_U(c, slot), c->rc_tmp.p));
            c->rc_spmvs--;
            std::vector<cx> hc((size_t)nq1);
            const double sc = c->rc_uscale[(size_t)slot];
            for (int i = 0; i < nq1; ++i) {
                const zc r = c->rc_R[(size_t)t][(size_t)slot * qc + i] / sc;
                hc[(size_t)i] = cx{r.real(), r.imag()};
            }
            double nn[2] = {0, 0};
            for (int k = 0; k < 2; ++k) {
                k_dot<1, true><<<NPART, VBLOCK, 0, c->stream>>>(n, c->rc_tmp.p, c->rc_tmp.p, part); EMB_LAUNCH_CHECK(c);
                k_finish<1><<<1, VBLOCK, 0, c->stream>>>(part, nrm); EMB_LAUNCH_CHECK(c);
                cx hv;
                EMB_CUDA(c, cudaMemcpyAsync(&hv, nrm, sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
                EMB_CUDA(c, cudaStreamSynchronize(c->stream));
                nn[k] = std::sqrt(std::fabs(hv.re));
                if (k == 0) {
                    EMB_CUDA(c, cudaMemcpyAsync(hC, hc.data(), (size_t)nq1 * sizeof(cx), cudaMemcpyHostToDevice, c->stream));
                    k_rc_sub<<<vb, 256, nq1 * sizeof(cx), c->stream>>>(n, nq1, hC, c->rcQ.p, c->rc_tmp.p); EMB_LAUNCH_CHECK(c);
                }
            }
            if (t == 0) {       // orthonormality of the new columns against the whole basis
                double worst = 0;
                for (int i = nq_before; i < nq1; ++i) {
                    k_rc_dots<1><<<dim3(RC_NP, nq1), VBLOCK, 0, c->stream>>>(n, c->rcQ.p, rc_Q(c, i), part); EMB_LAUNCH_CHECK(c);
                    k_rc_coef<1><<<nq1, VBLOCK, 0, c->stream>>>(part, hB); EMB_LAUNCH_CHECK(c);
                    std::vector<cx> hq((size_t)nq1);
                    EMB_CUDA(c, cudaMemcpyAsync(hq.data(), hB, (size_t)nq1 * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
                    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
                    for (int j = 0; j < nq1; ++j) {
                        const double d = std::hypot(hq[(size_t)j].re - (j == i ? 1.0 : 0.0), hq[(size_t)j].im);
                        if (d > worst) worst = d;
                    }
                }
                fprintf(stderr, "[rc debug] slot %d: max |Q^H q_new - e| %.3e\n", slot, worst);
            }
            fprintf(stderr, "[rc debug] slot %d term %d: |W u| %.3e  |W u - Q R| / |W u| %.3e  (nq %d -> %d)\n", slot, t, nn[0],
                    nn[0] > 0 ? nn[1] / nn[0] : 0.0, nq_before, nq1);
        }
    }
    if (ok) c->rc_version++;
    if (accepted) *accepted = ok;
    return EMB_OK;
}

// drops the oldest `drop` directions and recomputes Q and R from the remaining ones
static int rc_compact(emb_ctx* c, int drop) {
    const int keep = c->rc_n - drop;
    for (int j = 0; j < keep; ++j)
        EMB_CUDA(c, cudaMemcpyAsync(rc_U(c, j), rc_U(c, j + drop), (size_t)c->Ns * sizeof(cx), cudaMemcpyDeviceToDevice, c->stream));
    rc_clear(c);
    for (int j = 0; j < keep; ++j) {
        if (c->rc_n != j)           // an earlier direction was dropped: keep the slots dense
            EMB_CUDA(c, cudaMemcpyAsync(rc_U(c, c->rc_n), rc_U(c, j), (size_t)c->Ns * sizeof(cx), cudaMemcpyDeviceToDevice, c->stream));
        bool ok = false;
        EMB_TRY(rc_insert(c, c->rc_n, false, &ok));
        if (ok) c->rc_n++;
    }
    c->rc_rebuilds++;
    return EMB_OK;
}

// add direction d (contiguous, solve space) as the NEWEST member; needs the A(f) it was computed with (aff_coef)
static int rc_append(emb_ctx* c, const cx* d) {
    EMB_TRY(rc_prepare(c));
    if (c->rc_cap <= 0) return EMB_OK;             // no room for a basis on this device
    const int T = rc_T(c);
    if (c->rc_n >= c->rc_cap || c->rc_nq + T > c->rc_qcap) {
        int drop = c->rc_cap / 4 > 0 ? c->rc_cap / 4 : 1;
        if (c->rc_n < c->rc_cap) drop = 0;
        if (drop > c->rc_n) drop = c->rc_n;
        EMB_TRY(rc_compact(c, drop));
        if (c->rc_n >= c->rc_cap) return EMB_OK;      // cap == 1 and nothing could be dropped
    }
    const int slot = c->rc_n;
    if (d != rc_U(c, slot))
        EMB_CUDA(c, cudaMemcpyAsync(rc_U(c, slot), d, (size_t)c->Ns * sizeof(cx), cudaMemcpyDeviceToDevice, c->stream));
    bool ok = false;
    EMB_TRY(rc_insert(c, slot, true, &ok));
    if (ok) {
        c->rc_n = slot + 1;
        c->rc_accepted_total++;
    }
    return EMB_OK;
}

// xs += U y,  y = argmin || r0 - A(f) U y ||  for NV interleaved right-hand sides
template <int NV>
static int rc_project(emb_ctx* c, const cx* r0, cx* xs) {
    const int64_t n = c->Ns;
    if (c->rc_n > 0 && c->rc_terms != c->aff_sids) rc_clear(c);      // A(f) has other affine terms than the stored products
    const int m = c->rc_n, nq = c->rc_nq;
    if (m == 0 || nq == 0) return EMB_OK;
    cx* part = c->rc_part.p;
    cx* coef = rc_coef_area(c);
    k_rc_dots<NV><<<dim3(RC_NP, nq), VBLOCK, 0, c->stream>>>(n, c->rcQ.p, r0, part); EMB_LAUNCH_CHECK(c);
    k_rc_coef<NV><<<nq, VBLOCK, 0, c->stream>>>(part, coef); EMB_LAUNCH_CHECK(c);
    std::vector<cx> h((size_t)nq * NV);
    EMB_CUDA(c, cudaMemcpyAsync(h.data(), coef, h.size() * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<zc> G, y, g((size_t)nq * NV);
    std::vector<double> res;
    rc_build_G(c, m, G);
    for (int v = 0; v < NV; ++v)
        for (int i = 0; i < nq; ++i) g[(size_t)v * nq + i] = zc(h[(size_t)i * NV + v].re, h[(size_t)i * NV + v].im);
    ls_solve(nq, m, G, NV, g, y, res);
    std::vector<cx> yd((size_t)m * NV);
    for (int j = 0; j < m; ++j)
        for (int v = 0; v < NV; ++v) {
            const zc t = y[(size_t)j * NV + v] * c->rc_uscale[(size_t)j];
            yd[(size_t)j * NV + v] = cx{t.real(), t.imag()};
        }
    cx* ydev = coef + (size_t)3 * c->rc_qcap * NVMAX;
    EMB_CUDA(c, cudaMemcpyAsync(ydev, yd.data(), yd.size() * sizeof(cx), cudaMemcpyHostToDevice, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));      // yd is a stack-lifetime host buffer
    k_rc_combine<NV><<<blocks_for(n, 256), 256, (size_t)m * NV * sizeof(cx), c->stream>>>(n, m, ydev, c->rcU.p, xs);
    EMB_LAUNCH_CHECK(c);
    return EMB_OK;
}

// ---- EXPERIMENTAL: the reduced basis as a coarse space of the preconditioner -----------------------------------------
// c[i * nv + v] = sum_j C[i * m + j] t[j * nv + v]   (one block)
__global__ void k_small_mm(int m, int nv, const cx* __restrict__ C, const cx* __restrict__ t, cx* __restrict__ out) {
    for (int e = threadIdx.x; e < m * nv; e += blockDim.x) {
        const int i = e / nv, v = e % nv;
        cx acc = mk(0.0);
        for (int j = 0; j < m; ++j) fma_c(acc, C[i * m + j], t[j * nv + v]);
        out[e] = acc;
    }
}

// Coefficient map of the coarse correction for the current A(f) and the current basis (host side, small matrices):
//   Gs  = sym( D UtQ G(f) ),  D = diag(uscale)            = (U D)^T As (U D)   from the stored products, no SpMV
//   H   = D UhU D = V L V^H  (Jacobi),  T = V_k L_k^-1/2 for the eigenvalues above 1e-10 of the largest
//   Ceff = D T (T^T Gs T)^-1 T^T D                          so that  z += U Ceff (U^T r)
static int rc_coarse_update(emb_ctx* c) {
    c->coarse_m = 0;
    if (!c->coarse_basis || c->rc_n <= 0 || c->rc_UtQ.empty() || c->rc_terms != c->aff_sids) return EMB_OK;
    const int m = c->rc_n, nq = c->rc_nq, qc = c->rc_qcap, cap = c->rc_cap;
    std::vector<zc> G;
    rc_build_G(c, m, G);                                   // nq x m, columns already scaled by uscale
    std::vector<zc> Gs((size_t)m * m);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            zc s(0.0, 0.0);
            for (int k = 0; k < nq; ++k) s += c->rc_UtQ[(size_t)i * qc + k] * G[(size_t)j * nq + k];
            Gs[(size_t)i * m + j] = s * c->rc_uscale[(size_t)i];
        }
    for (int i = 0; i < m; ++i)
        for (int j = i + 1; j < m; ++j) {
            const zc a = 0.5 * (Gs[(size_t)i * m + j] + Gs[(size_t)j * m + i]);
            Gs[(size_t)i * m + j] = Gs[(size_t)j * m + i] = a;
        }
    std::vector<zc> H((size_t)m * m), V;
    std::vector<double> lam;
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) H[(size_t)i * m + j] = c->rc_uscale[(size_t)i] * c->rc_uscale[(size_t)j] * c->rc_UhU[(size_t)i * cap + j];
    herm_eig_jacobi(m, H, V, lam);
    double lmax = 0;
    for (double l : lam) if (l > lmax) lmax = l;
    std::vector<int> keep;
    for (int k = 0; k < m; ++k) if (lam[(size_t)k] > 1e-10 * lmax) keep.push_back(k);
    const int r = (int)keep.size();
    if (r == 0) return EMB_OK;
    std::vector<zc> T((size_t)m * r);                      // row-major m x r
    for (int i = 0; i < m; ++i)
        for (int k = 0; k < r; ++k) T[(size_t)i * r + k] = V[(size_t)i * m + keep[(size_t)k]] / std::sqrt(lam[(size_t)keep[(size_t)k]]);
    // Gc = T^T Gs T (r x r, column-major for ls_solve), right-hand sides T^T (r x m)
    std::vector<zc> GsT((size_t)m * r, zc(0.0, 0.0));
    for (int i = 0; i < m; ++i)
        for (int k = 0; k < r; ++k)
            for (int j = 0; j < m; ++j) GsT[(size_t)i * r + k] += Gs[(size_t)i * m + j] * T[(size_t)j * r + k];
    std::vector<zc> Gc((size_t)r * r, zc(0.0, 0.0)), rhs((size_t)r * m), Y;
    for (int a = 0; a < r; ++a)
        for (int b = 0; b < r; ++b)
            for (int i = 0; i < m; ++i) Gc[(size_t)b * r + a] += T[(size_t)i * r + a] * GsT[(size_t)i * r + b];
    for (int j = 0; j < m; ++j)
        for (int a = 0; a < r; ++a) rhs[(size_t)j * r + a] = T[(size_t)j * r + a];
    std::vector<double> res;
    ls_solve(r, r, Gc, m, rhs, Y, res);                    // Y[a * m + j] = (Gc^-1 T^T)[a][j]
    std::vector<cx> Ceff((size_t)m * m);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            zc s(0.0, 0.0);
            for (int a = 0; a < r; ++a) s += T[(size_t)i * r + a] * Y[(size_t)a * m + j];
            s *= c->rc_uscale[(size_t)i] * c->rc_uscale[(size_t)j];
            if (!(s == s)) return EMB_OK;                   // NaN: leave the coarse space off for this operator
            Ceff[(size_t)i * m + j] = cx{s.real(), s.imag()};
        }
    EMB_CUDA(c, cudaMemcpyAsync(c->rc_ceff.p, Ceff.data(), Ceff.size() * sizeof(cx), cudaMemcpyHostToDevice, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->coarse_m = m;
    c->coarse_version = c->rc_version;
    c->coarse_k0 = c->k0;
    return EMB_OK;
}

// z += U Ceff (U^T r) for NV interleaved columns, in two halves so that the first one (the projection: a bandwidth-bound
// pass over U and the small product) runs on its own stream next to the latency-bound multilevel cycle of the same
// residual; the second half (z += U y) follows the kernel that writes z.  Capturable (fork / join by events).
// Both passes stream the complex64 copy of U (half the bytes; U32 Ceff U32^T is still complex symmetric, and the 6e-8
// rounding of a coarse space is irrelevant for a preconditioner); EMB_COARSE_U64=1 streams U itself.
static inline bool rc_coarse_u64() {
    static const bool u64 = getenv("EMB_COARSE_U64") && atoi(getenv("EMB_COARSE_U64")) != 0;
    return u64;
}
template <int NV>
static int rc_coarse_begin(emb_ctx* c, const cx* r) {
    const int m = c->coarse_m;
    if (m <= 0) return EMB_OK;
    const int64_t n = c->Ns;
    cudaStream_t s = c->stream;
    if (c->use_side_streams) {
        if (!c->coarse_stream) {
            int least = 0, greatest = 0;
            cudaDeviceGetStreamPriorityRange(&least, &greatest);
            EMB_CUDA(c, cudaStreamCreateWithPriority(&c->coarse_stream, cudaStreamNonBlocking, least));
            EMB_CUDA(c, cudaEventCreateWithFlags(&c->ev_coarse_fork, cudaEventDisableTiming));
            EMB_CUDA(c, cudaEventCreateWithFlags(&c->ev_coarse_done, cudaEventDisableTiming));
        }
        s = c->coarse_stream;
        EMB_CUDA(c, cudaEventRecord(c->ev_coarse_fork, c->stream));      // r is complete here
        EMB_CUDA(c, cudaStreamWaitEvent(s, c->ev_coarse_fork, 0));
    }
    cx* t = c->rc_ct.p;
    cx* cc = c->rc_ct.p + (size_t)c->rc_cap * NVMAX;
    const cf* U32 = reinterpret_cast<const cf*>(c->rcU32.p);
    if (rc_coarse_u64() || !U32) { k_rc_dots<NV, false><<<dim3(RC_NP, m), VBLOCK, 0, s>>>(n, c->rcU.p, r, c->rc_part.p); EMB_LAUNCH_CHECK(c); }
    else { k_rc_dots<NV, false, cf><<<dim3(RC_NP, m), VBLOCK, 0, s>>>(n, U32, r, c->rc_part.p); EMB_LAUNCH_CHECK(c); }
    k_rc_coef<NV><<<m, VBLOCK, 0, s>>>(c->rc_part.p, t); EMB_LAUNCH_CHECK(c);
    k_small_mm<<<1, 256, 0, s>>>(m, NV, c->rc_ceff.p, t, cc); EMB_LAUNCH_CHECK(c);
    if (s != c->stream) EMB_CUDA(c, cudaEventRecord(c->ev_coarse_done, s));
    return EMB_OK;
}
template <int NV>
static int rc_coarse_finish(emb_ctx* c, cx* z) {
    const int m = c->coarse_m;
    if (m <= 0) return EMB_OK;
    const int64_t n = c->Ns;
    if (c->use_side_streams && c->coarse_stream) EMB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_coarse_done, 0));
    const cx* cc = c->rc_ct.p + (size_t)c->rc_cap * NVMAX;
    const cf* U32 = reinterpret_cast<const cf*>(c->rcU32.p);
    const size_t sh = (size_t)m * NV * sizeof(cx);
    if (rc_coarse_u64() || !U32) { k_rc_combine<NV><<<blocks_for(n, 256), 256, sh, c->stream>>>(n, m, cc, c->rcU.p, z); EMB_LAUNCH_CHECK(c); }
    else { k_rc_combine<NV, cf><<<blocks_for(n, 256), 256, sh, c->stream>>>(n, m, cc, U32, z); EMB_LAUNCH_CHECK(c); }
    return EMB_OK;
}
