// V-cycle of a smoothed-aggregation hierarchy on the GPU: real CSR matrices acting on NV interleaved complex vectors.
//
// The hierarchy (aggregates, smoothed prolongators, Galerkin operators, coarsest dense inverse) is built once per mesh on
// the host (emerge_b200/amg.py) for the nodal auxiliary problems of the preconditioner; only the cycle runs here.
// The reference has no counterpart (sparse direct solves, fem/solver.py:243-309).
// Cycle (symmetric, so the preconditioner stays complex-symmetric for COCR):
//   x = w D^-1 b;  b_c = P^T (b - A x);  x += P V(b_c);  x += w D^-1 (b - A x);   coarsest: x = A^-1 b (dense).
// The matrices are shared by every auxiliary space that uses the hierarchy; the level vectors belong to the space
// (AmgWork), so the cycles of different spaces run concurrently on different streams.
// All kernels are short-row gathers: LPR lanes per row chosen from the average row length, 12 B per nonzero + gathers.
#pragma once
#include "krylov.cuh"

// MODE 0: y = A x        MODE 1: y = b - A x        MODE 2: y = x + w d (b - A x)   (out of place)
// MODE 3: y += A x       MODE 4 (pre-smoothing from a zero guess fused with the residual): y2 = w d b, y = b - A y2,
//                         called with x = b
// KPR x NV lanes per row: lane (ks, v) walks nonzeros ks, ks + KPR, ... for column v (one gather wavefront per nonzero).
template <int KPR, int MODE, int NV, typename VX>
__global__ void __launch_bounds__(256) k_rcsr(int64_t n, const int64_t* __restrict__ ptr, const int* __restrict__ col,
                                              const double* __restrict__ val, const VX* __restrict__ x, VX* __restrict__ y,
                                              const VX* __restrict__ b, const double* __restrict__ d, double w,
                                              VX* __restrict__ y2) {
    constexpr int LPR = KPR * NV;
    const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t r = gt / LPR;
    const int s = (int)(gt % LPR);
    const int v = s % NV, ks = s / NV;
    double ar = 0.0, ai = 0.0;
    if (r < n)
        for (int64_t k = ptr[r] + ks; k < ptr[r + 1]; k += KPR) {
            const int cc = col[k];
            double a = val[k];
            if (MODE == 4) a *= w * d[cc];
            const cx u = ldx(x + (int64_t)cc * NV + v);
            ar += a * u.re;
            ai += a * u.im;
        }
#pragma unroll
    for (int o = LPR / 2; o >= NV; o >>= 1) {
        ar += __shfl_down_sync(0xffffffffu, ar, o, LPR);
        ai += __shfl_down_sync(0xffffffffu, ai, o, LPR);
    }
    if (r < n && ks == 0) {
        const int64_t o = r * NV + v;
        if (MODE == 4) {
            const cx bb = ldv(b, o);
            const double sc = w * d[r];
            stv(y2, o, cx{sc * bb.re, sc * bb.im});
            stv(y, o, cx{bb.re - ar, bb.im - ai});
        } else
        if (MODE == 0) stv(y, o, cx{ar, ai});
        else if (MODE == 1) { const cx bb = ldv(b, o); stv(y, o, cx{bb.re - ar, bb.im - ai}); }
        else if (MODE == 2) {
            const cx bb = ldv(b, o), xx = ldv(x, o);
            const double sc = w * d[r];
            stv(y, o, cx{xx.re + sc * (bb.re - ar), xx.im + sc * (bb.im - ai)});
        } else {
            const cx yy = ldv(y, o);
            stv(y, o, cx{yy.re + ar, yy.im + ai});
        }
    }
}
// dense real matrix times NV interleaved complex vectors, one warp per row
template <int NV, typename VX>
__global__ void __launch_bounds__(256) k_dense_mv(int64_t n, const double* __restrict__ M, const VX* __restrict__ b,
                                                  VX* __restrict__ x) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    double ar[NV], ai[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) ar[v] = ai[v] = 0.0;
    for (int64_t k = lane; k < n; k += 32) {
        const double a = M[r * n + k];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const cx u = ldv(b, k * NV + v);
            ar[v] += a * u.re;
            ai[v] += a * u.im;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            ar[v] += __shfl_down_sync(0xffffffffu, ar[v], o);
            ai[v] += __shfl_down_sync(0xffffffffu, ai[v], o);
        }
    if (lane == 0)
#pragma unroll
        for (int v = 0; v < NV; ++v) stv(x, r * NV + v, cx{ar[v], ai[v]});
}

// lanes per row from the average row length (rows of P^T hold a whole aggregate neighbourhood: 50-200 entries)
static inline int pick_lpr(int64_t nnz, int64_t n) {
    const double avg = n > 0 ? (double)nnz / (double)n : 0.0;
    return avg <= 12.0 ? 4 : (avg <= 48.0 ? 8 : 32);
}
template <int MODE, int NV, typename VX>
static int rcsr_launch(emb_ctx* c, cudaStream_t s, int lpr, int64_t n, const int64_t* ptr, const int* col, const double* val,
                       const VX* x, VX* y, const VX* b, const double* d, double w, VX* y2 = nullptr) {
    constexpr int KMAX = 32 / NV;          // KPR * NV lanes must fit a warp
    if (lpr <= 4 || KMAX <= 4) k_rcsr<4, MODE, NV, VX><<<blocks_for(n * 4 * NV, 256), 256, 0, s>>>(n, ptr, col, val, x, y, b, d, w, y2);
    else if (lpr == 8 || KMAX == 8) k_rcsr<8, MODE, NV, VX><<<blocks_for(n * 8 * NV, 256), 256, 0, s>>>(n, ptr, col, val, x, y, b, d, w, y2);
    else k_rcsr<KMAX, MODE, NV, VX><<<blocks_for(n * KMAX * NV, 256), 256, 0, s>>>(n, ptr, col, val, x, y, b, d, w, y2);
    EMB_LAUNCH_CHECK(c);
    return EMB_OK;
}

// V-cycle on stream s with right-hand side in W.b[0]; returns the device pointer holding the result (a level-0 buffer).
// The level vectors are allocated as complex128 and used in the storage type VX of the inner iteration.
template <int NV, typename VX>
static int amg_vcycle(emb_ctx* c, cudaStream_t s, AmgHierarchy& H, AmgWork& W, VX** result) {
    const int L = (int)H.lev.size();
    std::vector<VX*> xs(L, nullptr);
    auto B = [&](int l) { return reinterpret_cast<VX*>(W.b[l].p); };
    auto XA = [&](int l) { return reinterpret_cast<VX*>(W.xa[l].p); };
    auto XB = [&](int l) { return reinterpret_cast<VX*>(W.xb[l].p); };
    auto T = [&](int l) { return reinterpret_cast<VX*>(W.t[l].p); };
    for (int l = 0; l < L; ++l) {
        AmgLevel& v = H.lev[l];
        if (l == L - 1) {
            k_dense_mv<NV, VX><<<blocks_for(v.n * 32, 256), 256, 0, s>>>(v.n, H.cinv.p, B(l), XA(l)); EMB_LAUNCH_CHECK(c);
            xs[l] = XA(l);
            break;
        }
        // xa = w D^-1 b and t = b - A xa in one kernel
        EMB_TRY((rcsr_launch<4, NV, VX>(c, s, v.lpr_a, v.n, v.aptr.p, v.acol.p, v.aval.p, B(l), T(l), B(l), v.dinv.p, v.omega, XA(l))));
        EMB_TRY((rcsr_launch<0, NV, VX>(c, s, v.lpr_t, v.nc, v.tptr.p, v.tcol.p, v.tval.p, T(l), B(l + 1), nullptr, nullptr, 0.0)));
        xs[l] = XA(l);
    }
    for (int l = L - 2; l >= 0; --l) {
        AmgLevel& v = H.lev[l];
        EMB_TRY((rcsr_launch<3, NV, VX>(c, s, v.lpr_p, v.n, v.pptr.p, v.pcol.p, v.pval.p, xs[l + 1], XA(l), nullptr, nullptr, 0.0)));
        EMB_TRY((rcsr_launch<2, NV, VX>(c, s, v.lpr_a, v.n, v.aptr.p, v.acol.p, v.aval.p, XA(l), XB(l), B(l), v.dinv.p, v.omega)));
        xs[l] = XB(l);
    }
    *result = xs[0];
    return EMB_OK;
}

static void amg_release(AmgHierarchy& H) {
    for (auto& v : H.lev) {
        v.aptr.release(); v.pptr.release(); v.tptr.release(); v.acol.release(); v.pcol.release(); v.tcol.release();
        v.aval.release(); v.pval.release(); v.tval.release(); v.dinv.release();
    }
    H.lev.clear();
    H.cinv.release();
}
static void amg_work_release(AmgWork& W) {
    for (auto& b : W.b) b.release();
    for (auto& b : W.xa) b.release();
    for (auto& b : W.xb) b.release();
    for (auto& b : W.t) b.release();
    W.b.clear(); W.xa.clear(); W.xb.clear(); W.t.clear();
}
static int amg_work_alloc(emb_ctx* c, const AmgHierarchy& H, AmgWork& W) {
    const size_t L = H.lev.size();
    W.b.resize(L); W.xa.resize(L); W.xb.resize(L); W.t.resize(L);
    for (size_t l = 0; l < L; ++l) {
        const size_t n = (size_t)H.lev[l].n * NVMAX;
        EMB_TRY(dev_alloc(c, W.b[l], n));
        EMB_TRY(dev_alloc(c, W.xa[l], n));
        if (H.lev[l].nc > 0) {
            EMB_TRY(dev_alloc(c, W.xb[l], n));
            EMB_TRY(dev_alloc(c, W.t[l], n));
        }
    }
    return EMB_OK;
}
