// V-cycle of a smoothed-aggregation hierarchy on the GPU: real CSR matrices acting on complex vectors.
//
// The hierarchy (aggregates, smoothed prolongators, Galerkin operators, coarsest dense inverse) is built once per mesh on
// the host (emerge_b200/amg.py) for the nodal auxiliary problems of the preconditioner; only the cycle runs here.
// The reference has no counterpart (sparse direct solves, fem/solver.py:243-309).
// Cycle (symmetric, so the preconditioner stays complex-symmetric for COCR):
//   x = w D^-1 b;  b_c = P^T (b - A x);  x += P V(b_c);  x += w D^-1 (b - A x);   coarsest: x = A^-1 b (dense).
// All kernels are short-row gathers (7-30 nonzeros per row): LPR lanes per row, 12 B per nonzero + 16 B gathers.
#pragma once
#include "context.cuh"

// MODE 0: y = A x        MODE 1: y = b - A x        MODE 2: y = x + w d (b - A x)   (out of place)
template <int LPR, int MODE>
__global__ void __launch_bounds__(256) k_rcsr(int64_t n, const int64_t* __restrict__ ptr, const int* __restrict__ col,
                                              const double* __restrict__ val, const cx* __restrict__ x, cx* __restrict__ y,
                                              const cx* __restrict__ b, const double* __restrict__ d, double w) {
    const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t r = gt / LPR;
    const int sub = (int)(gt % LPR);
    double ar = 0, ai = 0;
    if (r < n)
        for (int64_t k = ptr[r] + sub; k < ptr[r + 1]; k += LPR) {
            const double a = val[k];
            const double2 v = __ldg(reinterpret_cast<const double2*>(x + col[k]));
            ar += a * v.x;
            ai += a * v.y;
        }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
        ar += __shfl_down_sync(0xffffffffu, ar, o, LPR);
        ai += __shfl_down_sync(0xffffffffu, ai, o, LPR);
    }
    if (r < n && sub == 0) {
        if (MODE == 0) y[r] = cx{ar, ai};
        else if (MODE == 1) { const cx bb = b[r]; y[r] = cx{bb.re - ar, bb.im - ai}; }
        else {
            const cx bb = b[r], xx = x[r];
            const double s = w * d[r];
            y[r] = cx{xx.re + s * (bb.re - ar), xx.im + s * (bb.im - ai)};
        }
    }
}
// y += s * (A x), complex s
template <int LPR>
__global__ void __launch_bounds__(256) k_rcsr_add(int64_t n, const int64_t* __restrict__ ptr, const int* __restrict__ col,
                                                  const double* __restrict__ val, const cx* __restrict__ x, cx s,
                                                  cx* __restrict__ y) {
    const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t r = gt / LPR;
    const int sub = (int)(gt % LPR);
    double ar = 0, ai = 0;
    if (r < n)
        for (int64_t k = ptr[r] + sub; k < ptr[r + 1]; k += LPR) {
            const double a = val[k];
            const double2 v = __ldg(reinterpret_cast<const double2*>(x + col[k]));
            ar += a * v.x;
            ai += a * v.y;
        }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
        ar += __shfl_down_sync(0xffffffffu, ar, o, LPR);
        ai += __shfl_down_sync(0xffffffffu, ai, o, LPR);
    }
    if (r < n && sub == 0) {
        cx yy = y[r];
        fma_c(yy, s, cx{ar, ai});
        y[r] = yy;
    }
}
// x = w d b
__global__ void k_amg_smooth0(int64_t n, const double* __restrict__ d, double w, const cx* __restrict__ b, cx* __restrict__ x) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) { const double s = w * d[i]; const cx v = b[i]; x[i] = cx{s * v.re, s * v.im}; }
}
// dense real matrix times complex vector, one warp per row
__global__ void __launch_bounds__(256) k_dense_mv(int64_t n, const double* __restrict__ M, const cx* __restrict__ b, cx* __restrict__ x) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    double ar = 0, ai = 0;
    for (int64_t k = lane; k < n; k += 32) {
        const double a = M[r * n + k];
        const cx v = b[k];
        ar += a * v.re;
        ai += a * v.im;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ar += __shfl_down_sync(0xffffffffu, ar, o);
        ai += __shfl_down_sync(0xffffffffu, ai, o);
    }
    if (lane == 0) x[r] = cx{ar, ai};
}

constexpr int AMG_LPR = 4;

// V-cycle with right-hand side in lev[0].b; returns the device pointer holding the result (a level-0 buffer)
static int amg_vcycle(emb_ctx* c, AmgHierarchy& H, cx** result) {
    const int L = (int)H.lev.size();
    std::vector<cx*> xs(L, nullptr);
    for (int l = 0; l < L; ++l) {
        AmgLevel& v = H.lev[l];
        if (l == L - 1) {
            k_dense_mv<<<blocks_for(v.n * 32, 256), 256, 0, c->stream>>>(v.n, H.cinv.p, v.b.p, v.xa.p); EMB_LAUNCH_CHECK(c);
            xs[l] = v.xa.p;
            break;
        }
        AmgLevel& nx = H.lev[l + 1];
        k_amg_smooth0<<<blocks_for(v.n, 256), 256, 0, c->stream>>>(v.n, v.dinv.p, v.omega, v.b.p, v.xa.p); EMB_LAUNCH_CHECK(c);
        k_rcsr<AMG_LPR, 1><<<blocks_for(v.n * AMG_LPR, 256), 256, 0, c->stream>>>(v.n, v.aptr.p, v.acol.p, v.aval.p, v.xa.p, v.t.p,
                                                                                 v.b.p, nullptr, 0.0); EMB_LAUNCH_CHECK(c);
        k_rcsr<AMG_LPR, 0><<<blocks_for(v.nc * AMG_LPR, 256), 256, 0, c->stream>>>(v.nc, v.tptr.p, v.tcol.p, v.tval.p, v.t.p, nx.b.p,
                                                                                  nullptr, nullptr, 0.0); EMB_LAUNCH_CHECK(c);
        xs[l] = v.xa.p;
    }
    for (int l = L - 2; l >= 0; --l) {
        AmgLevel& v = H.lev[l];
        k_rcsr_add<AMG_LPR><<<blocks_for(v.n * AMG_LPR, 256), 256, 0, c->stream>>>(v.n, v.pptr.p, v.pcol.p, v.pval.p, xs[l + 1],
                                                                                  mk(1.0), v.xa.p); EMB_LAUNCH_CHECK(c);
        k_rcsr<AMG_LPR, 2><<<blocks_for(v.n * AMG_LPR, 256), 256, 0, c->stream>>>(v.n, v.aptr.p, v.acol.p, v.aval.p, v.xa.p, v.xb.p,
                                                                                 v.b.p, v.dinv.p, v.omega); EMB_LAUNCH_CHECK(c);
        xs[l] = v.xb.p;
    }
    *result = xs[0];
    return EMB_OK;
}

static void amg_release(AmgHierarchy& H) {
    for (auto& v : H.lev) {
        v.aptr.release(); v.pptr.release(); v.tptr.release(); v.acol.release(); v.pcol.release(); v.tcol.release();
        v.aval.release(); v.pval.release(); v.tval.release(); v.dinv.release();
        v.b.release(); v.xa.release(); v.xb.release(); v.t.release();
    }
    H.lev.clear();
    H.cinv.release();
}
