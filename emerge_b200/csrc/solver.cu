// Per-frequency linear solve on the solve space.
//
// Reference path replaced: fem/solver.py:405-469 (SolveRoutine.solve) with its direct solvers (:243-309) and the
// per-port loop of the sweep driver (fem/physics/edm/emfreq3d.py:683-694).
// The reference eliminates prescribed dofs per solve, optionally reorders, factorises (SuperLU/PARDISO) and
// back-substitutes per port.  Here the eliminated pattern is built once (operators.cu) and A(f) X = B is solved for
// ALL ports of the frequency point at once:
//   1. start vectors from the reduced basis of earlier solves (recycle.cuh); accepted as they are when the FP64
//      residual of A(f) already meets rtol;
//   2. otherwise block COCR (conjugate-orthogonal conjugate-residual, one operator application per iteration; the
//      ports share ONE Krylov space, krylov.cuh: interleaved vectors, the operator is read once per iteration, NV x NV
//      small solves on the device) on the complex-symmetric part As = (A + A^T)/2 stored in complex64, with the
//      additive multilevel preconditioner (precond.cuh), wrapped in defect correction on the true FP64 operator:
//      X += As^-1 (B - A X).  A(f) is not exactly symmetric because the reference's mass matrix is not
//      (fem/mth/tet.py:1036; relative asymmetry ~5e-5); that and the complex64 rounding of As are both
//      removed by the outer correction, which gains two digits per step.  Padded groups, empty right-hand sides and
//      a breakdown of the block recurrence (dependent columns, diverging residual) use independent recurrences in
//      lockstep.  The exit test is always ||b - A(f) x|| / ||b|| <= rtol in FP64 on A(f).
//   method 0: restarted GMRES(m) on A, method 1: BiCGStab on A (single right-hand side; cross-checks).
// An iteration is ~60 kernel launches (the multilevel cycle is latency-bound), so the loop body is captured once per
// solve into a CUDA graph whose independent auxiliary-space branches run concurrently; every 10th iteration is issued
// as plain launches instead, with CUDA events around the operator application (roofline sampling) and the
// convergence read-back.  No host synchronisation otherwise: all scalars live on the device.
#include "context.cuh"
#include "krylov.cuh"
#include "precond.cuh"
#include "recycle.cuh"
#include "sell.cuh"
#include <algorithm>
#include <type_traits>
#include <vector>

// inner operator application: SELL layout (sell.cuh) when As was built that way, else block-CSR / CSR (krylov.cuh)
template <int NV, typename VT, typename VX>
static int apply_inner(emb_ctx* c, const VT* As, const VX* x, VX* y) {
    if constexpr (std::is_same<VT, cf>::value && std::is_same<VX, cx>::value) {
        if (c->sell_active) return bsell_launch<NV>(c, As, x, y);
    }
    return spmv_inner<NV, VT, VX>(c, As, x, y);
}
// complex64 storage of the inner operator in the layout of this pattern: SELL-8-sigma on a pair-ordered solve space
// (EMB_SPMV_SELL=0: block layout), with the padding slots zeroed once per buffer
static int alloc_As32(emb_ctx* c, cf** out, int nv) {
    // SELL for one or two right-hand sides; four interleaved columns gather 128 contiguous bytes per block already and
    // were measured faster in the block layout (1.44 vs 1.59 ms at 1M tets)
    static const bool sell_env = !(getenv("EMB_SPMV_SELL") && atoi(getenv("EMB_SPMV_SELL")) == 0);
    const bool want_sell = sell_env && nv <= 2;
    if (want_sell && c->paired && !c->sell_tried) {
        c->sell_tried = true;
        EMB_TRY(sell_build(c));
    }
    if (want_sell && c->paired && c->sell_ready) {
        EMB_TRY(dev_alloc(c, c->As32, (size_t)c->sell_blocks * 8));
        if (c->sell_zeroed != c->As32.p || !c->sell_active) {
            EMB_CUDA(c, cudaMemsetAsync(c->As32.p, 0, c->As32.n * sizeof(float), c->stream));
            c->sell_zeroed = c->As32.p;
            c->have_As = false;
        }
        c->sell_active = true;
    } else {
        EMB_TRY(dev_alloc(c, c->As32, (size_t)c->nnz_s * 2));
        if (c->sell_active) c->have_As = false;
        c->sell_active = false;
    }
    *out = reinterpret_cast<cf*>(c->As32.p);
    return EMB_OK;
}
template <typename VT>
static int fill_As(emb_ctx* c, VT* As) {
    if constexpr (std::is_same<VT, cf>::value) {
        if (c->sell_active) {
            k_sym_part_sell<<<blocks_for((c->Ns / 2) * 32, 256), 256, 0, c->stream>>>(c->Ns / 2, c->rowptr_s.p, c->col_s.p, c->A.p,
                                                                                   c->sell_pos.p, c->sell_sptr.p, As);
            EMB_LAUNCH_CHECK(c);
            return EMB_OK;
        }
    }
    k_sym_part<VT><<<blocks_for(c->Ns * 32, 256), 256, 0, c->stream>>>(c->Ns, c->paired ? 1 : 0, c->rowptr_s.p, c->col_s.p, c->A.p, As);
    EMB_LAUNCH_CHECK(c);
    return EMB_OK;
}

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
static inline cx* red_part(emb_ctx* c) { return reinterpret_cast<cx*>(c->red.p); }
static inline cx* red_sc(emb_ctx* c) { return red_part(c) + (size_t)4 * NPART * NVMAX * NVMAX; }

// `count` work vectors of Ns * nv entries
static int ensure_work(emb_ctx* c, size_t count, int nv) {
    if (c->work.size() < count) c->work.resize(count);
    for (size_t i = 0; i < count; ++i)
        if (c->work[i].n < (size_t)c->Ns * nv) EMB_TRY(dev_alloc(c, c->work[i], (size_t)c->Ns * nv));
    EMB_TRY(dev_alloc(c, c->red, ((size_t)4 * NPART * NVMAX * NVMAX + SC_SIZE + 8 * NVMAX + 16) * 2));
    return EMB_OK;
}

static int spmv1(emb_ctx* c, const cx* val, const cx* x, cx* y) { return spmv<1, cx>(c, val, x, y); }

static int dot_host(emb_ctx* c, bool conj, const cx* a, const cx* b, cx* out) {
    cx* part = red_part(c);
    cx* sc = red_sc(c) + SC_SIZE;
    if (conj) k_dot<1, true><<<NPART, VBLOCK, 0, c->stream>>>(c->Ns, a, b, part);
    else k_dot<1, false><<<NPART, VBLOCK, 0, c->stream>>>(c->Ns, a, b, part);
    EMB_LAUNCH_CHECK(c);
    k_finish<1><<<1, VBLOCK, 0, c->stream>>>(part, sc);
    EMB_LAUNCH_CHECK(c);
    EMB_CUDA(c, cudaMemcpyAsync(out, sc, sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    return EMB_OK;
}
// out[k] = sum_i conj(a[i][k]) a[i][k] for NV interleaved columns
template <int NV>
static int norms2_host(emb_ctx* c, const cx* a, double* out) {
    cx* part = red_part(c);
    cx* sc = red_sc(c) + SC_SIZE;
    k_dot<NV, true><<<NPART, VBLOCK, 0, c->stream>>>(c->Ns, a, a, part);
    EMB_LAUNCH_CHECK(c);
    k_finish<NV><<<1, VBLOCK, 0, c->stream>>>(part, sc);
    EMB_LAUNCH_CHECK(c);
    cx h[NVMAX];
    EMB_CUDA(c, cudaMemcpyAsync(h, sc, NV * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < NV; ++k) out[k] = h[k].re;
    return EMB_OK;
}

// ------------------------------------------------------------------------------------------------
// COCR on As D = RHS for NV columns in lockstep, D starts at 0.  Stops when every column has |r_k| <= stop_abs[k]
// (or maxit).  Returns iterations in *its.
// ------------------------------------------------------------------------------------------------
// preconditioner of the inner iteration: the multilevel cycle, plus (experimental, off by default) the reduced basis as a
// coarse space
template <int NV, typename VX>
static int precond_inner(emb_ctx* c, int pmode, const VX* r, VX* z) {
    bool coarse = false;
    if constexpr (std::is_same<VX, cx>::value) {
        coarse = c->coarse_basis && c->coarse_m > 0;
        if (coarse) EMB_TRY(rc_coarse_begin<NV>(c, r));        // U^T r on its own stream, next to the multilevel cycle
    }
    EMB_TRY((precond_apply<NV, VX>(c, pmode, r, z)));
    if constexpr (std::is_same<VX, cx>::value)
        if (coarse) EMB_TRY(rc_coarse_finish<NV>(c, z));
    return EMB_OK;
}

template <int NV, typename VT, typename VX>
struct CocrBody {
    emb_ctx* c;
    int pmode, block;
    const VT* As;
    VX *d, *r, *z, *p, *Az, *Ap, *MAp;
    cx *partA, *partR, *partZ, *sc;
    // one iteration; `sample`: CUDA events around the operator application and the preconditioner (plain launches only)
    int run(bool sample) {
        const int64_t n = c->Ns;
        if (sample) cudaEventRecord(c->evp0, c->stream);
        EMB_TRY((precond_inner<NV, VX>(c, pmode, Ap, MAp)));
        if (sample) cudaEventRecord(c->evp1, c->stream);
        k_gram<NV, VX><<<NPART, VBLOCK, 0, c->stream>>>(n, Ap, MAp, partA); EMB_LAUNCH_CHECK(c);
        k_bcocr_update<NV, VX><<<NPART, VBLOCK, 0, c->stream>>>(n, block, partA, sc, p, Ap, MAp, d, r, z, partR); EMB_LAUNCH_CHECK(c);
        if (sample) cudaEventRecord(c->evs0, c->stream);
        EMB_TRY((apply_inner<NV, VT, VX>(c, As, z, Az)));
        if (sample) cudaEventRecord(c->evs1, c->stream);
        k_gram<NV, VX><<<NPART, VBLOCK, 0, c->stream>>>(n, z, Az, partZ); EMB_LAUNCH_CHECK(c);
        k_bcocr_dir<NV, VX><<<NPART, VBLOCK, 0, c->stream>>>(n, block, partZ, partR, sc, z, Az, p, Ap); EMB_LAUNCH_CHECK(c);
        k_bcocr_commit<NV><<<1, 32, 0, c->stream>>>(sc); EMB_LAUNCH_CHECK(c);
        return EMB_OK;
    }
};

// rhs: FP64 residual of the outer loop; d: correction in the storage type VX of the inner iteration
template <int NV, typename VT, typename VX>
static int cocr(emb_ctx* c, int pmode, int block, const VT* As, const cx* rhs, VX* d, const double* stop_abs, int maxit, int* its,
                int* spmvs, double* rnorm_out) {
    const int64_t n = c->Ns, nn = c->Ns * NV;
    cx* part = red_part(c);
    auto W = [&](int i) { return reinterpret_cast<VX*>(c->work[i].p); };
    CocrBody<NV, VT, VX> B{c, pmode, block, As, d, W(0), W(1), W(2), W(3), W(4), W(5),
                           part, part + (size_t)NPART * NVMAX * NVMAX, part + (size_t)2 * NPART * NVMAX * NVMAX, red_sc(c)};
    const unsigned vb = blocks_for(nn, 256);
    k_zero_v<VX><<<vb, 256, 0, c->stream>>>(nn, d); EMB_LAUNCH_CHECK(c);
    k_convert<cx, VX><<<vb, 256, 0, c->stream>>>(nn, rhs, B.r); EMB_LAUNCH_CHECK(c);
    EMB_TRY((precond_inner<NV, VX>(c, pmode, B.r, B.z)));
    k_convert<VX, VX><<<vb, 256, 0, c->stream>>>(nn, B.z, B.p); EMB_LAUNCH_CHECK(c);
    EMB_TRY((apply_inner<NV, VT, VX>(c, As, B.z, B.Az))); ++*spmvs;
    k_convert<VX, VX><<<vb, 256, 0, c->stream>>>(nn, B.Az, B.Ap); EMB_LAUNCH_CHECK(c);
    k_gram<NV, VX><<<NPART, VBLOCK, 0, c->stream>>>(n, B.z, B.Az, B.partZ); EMB_LAUNCH_CHECK(c);
    k_gram_finish<NV><<<1, VBLOCK, 0, c->stream>>>(B.partZ, B.sc + SC_RHO); EMB_LAUNCH_CHECK(c);

    // capture one iteration into a graph (side-stream branches of the preconditioner become parallel graph branches)
    static const bool use_graph = !(getenv("EMB_GRAPH") && atoi(getenv("EMB_GRAPH")) == 0);
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    int64_t body_launches = 0;
    if (use_graph && maxit > 2) {
        const int64_t l0 = c->launches;
        bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            const int rc = B.run(false);
            const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
            ok = rc == EMB_OK && e == cudaSuccess && graph != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&gexec, graph, 0) == cudaSuccess;
        body_launches = c->launches - l0;
        c->launches = l0;
        if (!ok) {
            cudaGetLastError();
            if (gexec) cudaGraphExecDestroy(gexec);
            if (graph) cudaGraphDestroy(graph);
            gexec = nullptr; graph = nullptr;
            static bool warned = false;
            if (!warned) { fprintf(stderr, "[emb] CUDA graph capture of the COCR iteration failed; using plain launches\n"); warned = true; }
        }
    }
    int it = 0, rc = EMB_OK;
    double rn[NVMAX], rbest[NVMAX];
    for (int k = 0; k < NV; ++k) rn[k] = rbest[k] = 1e300;
    const int check = 10;
    while (it < maxit) {
        const bool sample = ((it + 1) % check == 0) || !gexec;
        if (sample) {
            if ((rc = B.run((it + 1) % check == 0)) < 0) break;
        } else {
            if (cudaGraphLaunch(gexec, c->stream) != cudaSuccess) { c->err = "cudaGraphLaunch failed"; rc = EMB_ERR_CUDA; break; }
            c->launches += body_launches;
            c->graph_launches++;
        }
        ++*spmvs;
        ++it;
        if (it % check == 0 || it == maxit) {
            cx h[NVMAX];
            if (cudaMemcpyAsync(h, B.sc + SC_RR, NV * sizeof(cx), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                cudaStreamSynchronize(c->stream) != cudaSuccess) { c->err = "COCR: read-back failed"; rc = EMB_ERR_CUDA; break; }
            if (it % check == 0) {
                float sms = 0;
                if (cudaEventElapsedTime(&sms, c->evs0, c->evs1) == cudaSuccess) { c->spmv_ms_sum += sms; c->spmv_ms_cnt++; }
                if (cudaEventElapsedTime(&sms, c->evp0, c->evp1) == cudaSuccess) { c->prec_ms_sum += sms; c->prec_ms_cnt++; }
            }
            bool all = true, nan = false;
            for (int k = 0; k < NV; ++k) {
                rn[k] = sqrt(fabs(h[k].re));
                if (!(rn[k] == rn[k])) nan = true;
                if (!(rn[k] <= stop_abs[k])) all = false;
                // a residual that has grown a million times above its smallest value is a breakdown, not slow convergence
                if (rn[k] < rbest[k]) rbest[k] = rn[k];
                if (rn[k] > 1e6 * rbest[k] && rbest[k] > 0) nan = true;
            }
            if (nan) { c->err = "COCR breakdown (NaN or diverging residual)"; rc = EMB_NOT_CONVERGED; break; }
            if (all) break;
        }
    }
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
    *its = it;
    for (int k = 0; k < NV; ++k) rnorm_out[k] = rn[k];
    return rc;
}

// restarted GMRES on A with left... right preconditioning: A M^-1 u = b, x = M^-1 u
static int gmres(emb_ctx* c, int pmode, const cx* A, const cx* b, cx* x, double bnorm, const emb_solve_opts* o, int* its,
                 int* spmvs, double* relres) {
    const int64_t n = c->Ns;
    const int m = o->restart > 0 ? o->restart : 50;
    EMB_TRY(ensure_work(c, (size_t)m + 8, 1));
    cx *r = c->work[0].p, *w = c->work[1].p, *t = c->work[2].p;
    auto V = [&](int j) { return c->work[6 + j].p; };
    const unsigned vb = blocks_for(n, 256);
    std::vector<cx> H((size_t)(m + 1) * m), cs(m), sn(m), g(m + 1), y(m);
    cx* dsc = red_sc(c) + SC_SIZE + NVMAX + 2;
    int it = 0;
    double res = 1.0;
    auto set_scalar = [&](cx v) { return cudaMemcpyAsync(dsc, &v, sizeof(cx), cudaMemcpyHostToDevice, c->stream); };
    while (it < o->maxit) {
        // r = b - A x
        EMB_TRY(spmv1(c, A, x, r)); ++*spmvs;
        k_axpby<<<vb, 256, 0, c->stream>>>(n, nullptr, 1.0, b, nullptr, -1.0, r); EMB_LAUNCH_CHECK(c);
        cx rr;
        EMB_TRY(dot_host(c, true, r, r, &rr));
        double beta = sqrt(rr.re);
        res = beta / bnorm;
        if (res <= o->rtol) break;
        EMB_CUDA(c, set_scalar(mk(1.0 / beta)));
        k_axpby<<<vb, 256, 0, c->stream>>>(n, dsc, 1.0, r, nullptr, 0.0, V(0)); EMB_LAUNCH_CHECK(c);
        std::fill(g.begin(), g.end(), mk(0.0));
        g[0] = mk(beta);
        int j = 0;
        for (; j < m && it < o->maxit; ++j, ++it) {
            EMB_TRY((precond_apply<1, cx>(c, pmode, V(j), t)));
            EMB_TRY(spmv1(c, A, t, w)); ++*spmvs;
            // classical Gram-Schmidt, two passes
            for (int i = 0; i <= j; ++i) H[(size_t)i * m + j] = mk(0.0);
            for (int pass = 0; pass < 2; ++pass) {
                std::vector<cx> h(j + 1);
                for (int i = 0; i <= j; ++i) EMB_TRY(dot_host(c, true, V(i), w, &h[i]));
                for (int i = 0; i <= j; ++i) {
                    EMB_CUDA(c, set_scalar(-h[i]));
                    k_axpby<<<vb, 256, 0, c->stream>>>(n, dsc, 1.0, V(i), nullptr, 1.0, w); EMB_LAUNCH_CHECK(c);
                    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
                    H[(size_t)i * m + j] += h[i];
                }
            }
            cx ww;
            EMB_TRY(dot_host(c, true, w, w, &ww));
            const double hn = sqrt(ww.re);
            H[(size_t)(j + 1) * m + j] = mk(hn);
            if (j + 1 < m || true) {
                EMB_CUDA(c, set_scalar(mk(hn > 0 ? 1.0 / hn : 0.0)));
                k_axpby<<<vb, 256, 0, c->stream>>>(n, dsc, 1.0, w, nullptr, 0.0, V(j + 1)); EMB_LAUNCH_CHECK(c);
                EMB_CUDA(c, cudaStreamSynchronize(c->stream));
            }
            // Givens
            for (int i = 0; i < j; ++i) {
                cx a = H[(size_t)i * m + j], bq = H[(size_t)(i + 1) * m + j];
                H[(size_t)i * m + j] = conj(cs[i]) * a + conj(sn[i]) * bq;
                H[(size_t)(i + 1) * m + j] = -sn[i] * a + cs[i] * bq;
            }
            {
                cx a = H[(size_t)j * m + j], bq = H[(size_t)(j + 1) * m + j];
                double den = sqrt(norm2(a) + norm2(bq));
                if (den == 0) den = 1e-300;
                cs[j] = (1.0 / den) * a;
                sn[j] = (1.0 / den) * bq;
                H[(size_t)j * m + j] = conj(cs[j]) * a + conj(sn[j]) * bq;
                H[(size_t)(j + 1) * m + j] = mk(0.0);
                cx g0 = g[j];
                g[j] = conj(cs[j]) * g0;
                g[j + 1] = -sn[j] * g0;
            }
            res = sqrt(norm2(g[j + 1])) / bnorm;
            if (res <= o->rtol) { ++j; ++it; break; }
        }
        // solve upper triangular, x += M^-1 V y
        for (int i = j - 1; i >= 0; --i) {
            cx s = g[i];
            for (int k = i + 1; k < j; ++k) s -= H[(size_t)i * m + k] * y[k];
            y[i] = cdiv(s, H[(size_t)i * m + i]);
        }
        k_zero<<<vb, 256, 0, c->stream>>>(n, w); EMB_LAUNCH_CHECK(c);
        for (int i = 0; i < j; ++i) {
            EMB_CUDA(c, set_scalar(y[i]));
            k_axpby<<<vb, 256, 0, c->stream>>>(n, dsc, 1.0, V(i), nullptr, 1.0, w); EMB_LAUNCH_CHECK(c);
            EMB_CUDA(c, cudaStreamSynchronize(c->stream));
        }
        EMB_TRY((precond_apply<1, cx>(c, pmode, w, t)));
        k_axpby<<<vb, 256, 0, c->stream>>>(n, nullptr, 1.0, t, nullptr, 1.0, x); EMB_LAUNCH_CHECK(c);
    }
    *its = it;
    *relres = res;
    return EMB_OK;
}

static int bicgstab(emb_ctx* c, int pmode, const cx* A, const cx* b, cx* x, double bnorm, const emb_solve_opts* o, int* its,
                    int* spmvs, double* relres) {
    const int64_t n = c->Ns;
    cx *r = c->work[0].p, *r0 = c->work[1].p, *p = c->work[2].p, *v = c->work[3].p, *s = c->work[4].p, *t = c->work[5].p,
       *ph = c->work[6].p, *sh = c->work[7].p;
    const unsigned vb = blocks_for(n, 256);
    cx* dsc = red_sc(c) + SC_SIZE + NVMAX + 2;
    auto set_scalar = [&](cx val) { return cudaMemcpyAsync(dsc, &val, sizeof(cx), cudaMemcpyHostToDevice, c->stream); };
    auto axpy = [&](cx a, const cx* xx, double bfac, cx* yy) -> int {
        EMB_CUDA(c, set_scalar(a));
        k_axpby<<<vb, 256, 0, c->stream>>>(n, dsc, 1.0, xx, nullptr, bfac, yy);
        EMB_LAUNCH_CHECK(c);
        EMB_CUDA(c, cudaStreamSynchronize(c->stream));
        return EMB_OK;
    };
    EMB_TRY(spmv1(c, A, x, r)); ++*spmvs;
    k_axpby<<<vb, 256, 0, c->stream>>>(n, nullptr, 1.0, b, nullptr, -1.0, r); EMB_LAUNCH_CHECK(c);
    k_copy<<<vb, 256, 0, c->stream>>>(n, r, r0); EMB_LAUNCH_CHECK(c);
    k_zero<<<vb, 256, 0, c->stream>>>(n, p); EMB_LAUNCH_CHECK(c);
    k_zero<<<vb, 256, 0, c->stream>>>(n, v); EMB_LAUNCH_CHECK(c);
    cx rho = mk(1.0), alpha = mk(1.0), omega = mk(1.0);
    int it = 0;
    double res = 1.0;
    cx rr;
    EMB_TRY(dot_host(c, true, r, r, &rr));
    res = sqrt(rr.re) / bnorm;
    while (it < o->maxit && res > o->rtol) {
        cx rho1;
        EMB_TRY(dot_host(c, true, r0, r, &rho1));
        if (norm2(rho1) == 0) break;
        const cx beta = cdiv(rho1, rho) * cdiv(alpha, omega);
        // p = r + beta (p - omega v)
        EMB_TRY(axpy(-omega, v, 1.0, p));
        EMB_TRY(axpy(mk(1.0), r, 0.0, t));       // t = r (temp)
        EMB_TRY(axpy(beta, p, 1.0, t));          // t = r + beta p
        k_copy<<<vb, 256, 0, c->stream>>>(n, t, p); EMB_LAUNCH_CHECK(c);
        EMB_TRY((precond_apply<1, cx>(c, pmode, p, ph)));
        EMB_TRY(spmv1(c, A, ph, v)); ++*spmvs;
        cx r0v;
        EMB_TRY(dot_host(c, true, r0, v, &r0v));
        alpha = cdiv(rho1, r0v);
        k_copy<<<vb, 256, 0, c->stream>>>(n, r, s); EMB_LAUNCH_CHECK(c);
        EMB_TRY(axpy(-alpha, v, 1.0, s));
        EMB_TRY((precond_apply<1, cx>(c, pmode, s, sh)));
        EMB_TRY(spmv1(c, A, sh, t)); ++*spmvs;
        cx ts, tt;
        EMB_TRY(dot_host(c, true, t, s, &ts));
        EMB_TRY(dot_host(c, true, t, t, &tt));
        omega = (tt.re > 0) ? (1.0 / tt.re) * ts : mk(0.0);
        EMB_TRY(axpy(alpha, ph, 1.0, x));
        EMB_TRY(axpy(omega, sh, 1.0, x));
        k_copy<<<vb, 256, 0, c->stream>>>(n, s, r); EMB_LAUNCH_CHECK(c);
        EMB_TRY(axpy(-omega, t, 1.0, r));
        rho = rho1;
        EMB_TRY(dot_host(c, true, r, r, &rr));
        res = sqrt(rr.re) / bnorm;
        ++it;
        if (!(res == res)) break;
    }
    *its = it;
    *relres = res;
    return EMB_OK;
}



// ------------------------------------------------------------------------------------------------
// lockstep solve of NV right-hand sides: A xs = bs (device, solve space, interleaved); xs is in/out
// (initial guess when use_x0)
// ------------------------------------------------------------------------------------------------
// As (inner operator, block layout on a pair-ordered solve space) and the preconditioner, whose diagonal blocks and
// Galerkin diagonals are those of the symmetric part and are read straight from A(f): diag(R^T A R) = diag(R^T As R)
template <typename VT>
static int ensure_operator(emb_ctx* c, int precond, VT* As) {
    const int64_t n = c->Ns;
    if (!c->have_As) {
        EMB_TRY(fill_As<VT>(c, As));
        EMB_TRY(precond_setup<cx>(c, precond, c->A.p));
        c->have_As = true;
        c->As_precond = precond;
    } else if (c->As_precond != precond) {
        EMB_TRY(precond_setup<cx>(c, precond, c->A.p));
        c->As_precond = precond;
    }
    return EMB_OK;
}

template <int NV>
static int solve_device(emb_ctx* c, const emb_solve_opts* o, const cx* bs, cx* xs, emb_solve_info* info) {
    const int64_t n = c->Ns, nn = c->Ns * NV;
    const unsigned vb = blocks_for(nn, 256);
    EMB_TRY(ensure_work(c, 10, NV));
    if (!c->evp0) { EMB_CUDA(c, cudaEventCreate(&c->evp0)); EMB_CUDA(c, cudaEventCreate(&c->evp1)); }
    cudaEventRecord(c->ev0, c->stream);
    int its = 0, spmvs = 0, rc = EMB_OK;
    double relres[NVMAX], bnorm[NVMAX], b2[NVMAX];
    EMB_TRY(norms2_host<NV>(c, bs, b2));
    bool any_rhs = false;
    for (int k = 0; k < NV; ++k) { bnorm[k] = sqrt(b2[k]); relres[k] = bnorm[k] > 0 ? 1.0 : 0.0; any_rhs |= bnorm[k] > 0; }
    const bool recycle = c->rc_cap > 0 && any_rhs && o->method == 2;
    c->rc_last_proj_relres = -1;
    if (!o->use_x0 || !any_rhs) { k_zero<<<vb, 256, 0, c->stream>>>(nn, xs); EMB_LAUNCH_CHECK(c); }
    cx* rr = c->work[8].p;
    cx* dd = c->work[9].p;
    bool have_guess = o->use_x0 != 0;
    if (recycle) {
        EMB_TRY(dev_alloc(c, c->rc_x0, (size_t)c->Ns * NVMAX));
        if (c->rc_n > 0) {
            const cx* r0 = bs;
            if (o->use_x0) {    // residual of the caller's guess
                EMB_TRY((spmv_resid<NV, cx>(c, c->A.p, xs, bs, rr))); ++spmvs;
                r0 = rr;
            }
            EMB_TRY(rc_project<NV>(c, r0, xs));
            have_guess = true;
        }
        k_copy<<<vb, 256, 0, c->stream>>>(nn, xs, c->rc_x0.p); EMB_LAUNCH_CHECK(c);
    }
    const double target = recycle ? o->rtol * c->rc_snap : o->rtol;
    emb_solve_opts ot = *o;
    ot.rtol = target;
    if (!any_rhs) {
        // nothing to do
    } else if (o->method == 0 || o->method == 1) {
        if (NV != 1) { c->err = "GMRES / BiCGStab solve one right-hand side at a time"; return EMB_ERR_ARG; }
        EMB_TRY(precond_setup<cx>(c, o->precond, c->A.p));
        c->have_As = false;
        if (o->method == 0) EMB_TRY(gmres(c, o->precond, c->A.p, bs, xs, bnorm[0], &ot, &its, &spmvs, &relres[0]));
        else EMB_TRY(bicgstab(c, o->precond, c->A.p, bs, xs, bnorm[0], &ot, &its, &spmvs, &relres[0]));
        // true residual at exit
        EMB_TRY((spmv_resid<1, cx>(c, c->A.p, xs, bs, rr))); ++spmvs;
        double r2[NVMAX];
        EMB_TRY(norms2_host<1>(c, rr, r2));
        relres[0] = sqrt(r2[0]) / bnorm[0];
    } else {
        // defect correction on A with COCR on the symmetric part; the symmetric part and the preconditioner are only
        // built when the point really has to iterate
        static const bool as64 = getenv("EMB_AS_FP64") && atoi(getenv("EMB_AS_FP64")) != 0;
        const bool fp32 = c->as_fp32 && !as64;
        static const double inner_red = getenv("EMB_INNER") ? atof(getenv("EMB_INNER")) : 1e-2;
        static const bool verbose = getenv("EMB_VERBOSE") != nullptr;
        bool iterated = false, block_failed = false;
        for (int outer = 0; outer < 40 && its < o->maxit; ++outer) {
            double rn[NVMAX];
            if (outer > 0 || have_guess) {
                EMB_TRY((spmv_resid<NV, cx>(c, c->A.p, xs, bs, rr))); ++spmvs;
                double r2[NVMAX];
                EMB_TRY(norms2_host<NV>(c, rr, r2));
                for (int k = 0; k < NV; ++k) rn[k] = sqrt(r2[k]);
            } else {
                k_copy<<<vb, 256, 0, c->stream>>>(nn, bs, rr); EMB_LAUNCH_CHECK(c);
                for (int k = 0; k < NV; ++k) rn[k] = bnorm[k];
            }
            bool all_rtol = true, all_target = true;
            double worst = 0;
            for (int k = 0; k < NV; ++k) {
                relres[k] = bnorm[k] > 0 ? rn[k] / bnorm[k] : 0.0;
                if (!(relres[k] <= o->rtol)) all_rtol = false;
                if (!(relres[k] <= target)) all_target = false;
                if (relres[k] > worst || relres[k] != relres[k]) worst = relres[k];
            }
            if (outer == 0 && recycle && c->rc_n > 0) c->rc_last_proj_relres = worst;
            // start vectors that already meet rtol are accepted as they are; once the point has to iterate it feeds the
            // reduced basis and every column is run to the tighter snapshot tolerance
            if (outer == 0 ? all_rtol : all_target) break;
            double stop[NVMAX];
            for (int k = 0; k < NV; ++k) {
                // inner target: two digits below the current residual, never below what the outer loop needs
                stop[k] = inner_red * rn[k];
                const double need = 0.3 * target * bnorm[k];
                if (stop[k] < need) stop[k] = need;
            }
            int iit = 0;
            double irn[NVMAX];
            // one Krylov space for all columns (block COCR) unless a column is empty / already exact, the group is
            // padded, or the block recurrence broke down earlier in this solve
            bool blk = NV > 1 && c->block_krylov && !block_failed;
            for (int k = 0; k < NV; ++k)
                if (!(bnorm[k] > 0) || !(rn[k] > 0)) blk = false;
            if (c->coarse_basis && (c->coarse_version != c->rc_version || c->coarse_k0 != c->k0))
                EMB_TRY(rc_coarse_update(c));          // experimental: coefficient map of the coarse space for this A(f)
            if (fp32) {
                cf* As = nullptr;
                EMB_TRY(alloc_As32(c, &As, NV));
                EMB_TRY(ensure_operator<cf>(c, o->precond, As));
                rc = cocr<NV, cf, cx>(c, o->precond, blk ? 1 : 0, As, rr, dd, stop, o->maxit - its, &iit, &spmvs, irn);
            } else {
                EMB_TRY(dev_alloc(c, c->As, (size_t)c->nnz_s));
                EMB_TRY(ensure_operator<cx>(c, o->precond, c->As.p));
                rc = cocr<NV, cx, cx>(c, o->precond, blk ? 1 : 0, c->As.p, rr, dd, stop, o->maxit - its, &iit, &spmvs, irn);
            }
            if (rc == EMB_NOT_CONVERGED && blk) {      // breakdown of the block recurrence: redo this step column by column
                if (verbose) fprintf(stderr, "[emb] block COCR breakdown after %d iterations; lockstep recurrences from here\n", iit);
                block_failed = true;
                its += iit;
                rc = EMB_OK;
                continue;
            }
            its += iit;
            iterated = true;
            if (verbose)
                fprintf(stderr, "[emb] outer %d nv %d worst relres %.3e -> inner %d its, inner residual[0] %.3e\n", outer, NV, worst,
                        iit, irn[0] / (bnorm[0] > 0 ? bnorm[0] : 1.0));
            if (rc < 0) return rc;
            k_add_into<cx><<<vb, 256, 0, c->stream>>>(nn, dd, xs); EMB_LAUNCH_CHECK(c);
            if (rc == EMB_NOT_CONVERGED) break;
        }
        if (iterated) {
            EMB_TRY((spmv_resid<NV, cx>(c, c->A.p, xs, bs, rr))); ++spmvs;
            double r2[NVMAX];
            EMB_TRY(norms2_host<NV>(c, rr, r2));
            for (int k = 0; k < NV; ++k) relres[k] = bnorm[k] > 0 ? sqrt(r2[k]) / bnorm[k] : 0.0;
        }
        if (recycle && iterated) {
            EMB_TRY(dev_alloc(c, c->rc_tmp, (size_t)c->Ns));
            for (int k = 0; k < NV; ++k) {
                if (!(bnorm[k] > 0) || !(relres[k] <= 1e2 * o->rtol)) continue;
                k_extract_col<<<blocks_for(n, 256), 256, 0, c->stream>>>(n, xs, c->rc_x0.p, NV, k, c->rc_tmp.p); EMB_LAUNCH_CHECK(c);
                EMB_TRY(rc_append(c, c->rc_tmp.p));
            }
        }
    }
    cudaEventRecord(c->ev1, c->stream);
    cudaEventSynchronize(c->ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->ms["solve"] = ms;
    bool ok = true;
    for (int k = 0; k < NV; ++k) {
        if (info) { info[k].iters = its; info[k].relres = relres[k]; info[k].ms = ms / NV; info[k].spmvs = spmvs; }
        if (!(relres[k] <= o->rtol)) {
            ok = false;
            c->err = "solver did not reach rtol: relres=" + std::to_string(relres[k]) + " after " + std::to_string(its) + " iterations";
        }
    }
    return ok ? EMB_OK : EMB_NOT_CONVERGED;
}

static int solve_dispatch(emb_ctx* c, int nv, const emb_solve_opts* o, const cx* bs, cx* xs, emb_solve_info* info) {
    switch (nv) {
        case 1: return solve_device<1>(c, o, bs, xs, info);
        case 2: return solve_device<2>(c, o, bs, xs, info);
        case 4: return solve_device<4>(c, o, bs, xs, info);
    }
    c->err = "lockstep width must be 1, 2 or 4";
    return EMB_ERR_ARG;
}

// Asynchronous variant of the host copy: column k is scattered into its own staging vector on the compute stream; the
// copy stream moves it to the caller's buffer.  The staging vector is not rewritten before its previous copy has finished
// (the compute stream waits on that copy's event - no host synchronisation anywhere); emb_fields_sync() is the host's wait.
static int finish_solution_async(emb_ctx* c, int k, emb_c128* x_full) {
    DevBuf<cx>& st = c->xstage[k];
    EMB_TRY(dev_alloc(c, st, (size_t)c->N));
    if (!c->ev_stage_ready[k]) {
        EMB_CUDA(c, cudaEventCreateWithFlags(&c->ev_stage_ready[k], cudaEventDisableTiming));
        EMB_CUDA(c, cudaEventCreateWithFlags(&c->ev_stage_done[k], cudaEventDisableTiming));
    } else {
        EMB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_stage_done[k], 0));
    }
    EMB_CUDA(c, cudaMemsetAsync(st.p, 0, (size_t)c->N * sizeof(cx), c->stream));
    k_scatter_col<<<blocks_for(c->Ns, 256), 256, 0, c->stream>>>(c->Ns, c->solve_ids.p, c->xs.p, c->nsol, k, st.p);
    EMB_LAUNCH_CHECK(c);
    EMB_CUDA(c, cudaEventRecord(c->ev_stage_ready[k], c->stream));
    EMB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_stage_ready[k], 0));
    // in pieces: the scalar read-backs of the following point (projection coefficients, residual norms) share the D2H copy
    // engine with this transfer and would otherwise queue behind all of it (measured: no overlap with one 100 MB copy)
    const size_t total = (size_t)c->N * sizeof(cx), piece = (size_t)4 << 20;
    for (size_t off = 0; off < total; off += piece)
        EMB_CUDA(c, cudaMemcpyAsync(reinterpret_cast<char*>(x_full) + off, reinterpret_cast<const char*>(st.p) + off,
                                    std::min(piece, total - off), cudaMemcpyDeviceToHost, c->copy_stream));
    EMB_CUDA(c, cudaEventRecord(c->ev_stage_done[k], c->copy_stream));
    return EMB_OK;
}

// full-space copy of column k of the last solve (zeros at eliminated dofs), optionally to the host
static int finish_solution(emb_ctx* c, int k, emb_c128* x_full) {
    if (x_full && c->fields_async && c->copy_stream && k < 4) {
        EMB_TRY(finish_solution_async(c, k, x_full));
        if (k != 0) return EMB_OK;
        x_full = nullptr;                 // column 0 also becomes the device-resident "last solution" below
    }
    EMB_TRY(dev_alloc(c, c->xfull, (size_t)c->N));
    EMB_CUDA(c, cudaMemsetAsync(c->xfull.p, 0, (size_t)c->N * sizeof(cx), c->stream));
    k_scatter_col<<<blocks_for(c->Ns, 256), 256, 0, c->stream>>>(c->Ns, c->solve_ids.p, c->xs.p, c->nsol, k, c->xfull.p);
    EMB_LAUNCH_CHECK(c);
    if (x_full) EMB_CUDA(c, cudaMemcpyAsync(x_full, c->xfull.p, (size_t)c->N * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    return EMB_OK;
}

static const emb_solve_opts kDefaultOpts = {2, 2, 50, 100000, 1e-8, 0};

extern "C" int emb_solve_multi(emb_ctx* c, int nrhs, const int* sids, const emb_solve_opts* opts, emb_c128* const* x_full,
                               emb_solve_info* infos) {
    if (!c || nrhs < 1 || nrhs > NVMAX || !sids) return EMB_ERR_ARG;
    const emb_solve_opts* o = opts ? opts : &kDefaultOpts;
    for (int k = 0; k < nrhs; ++k)
        if (sids[k] < 0 || sids[k] >= 16) return EMB_ERR_ARG;
    for (int k = 0; k < nrhs; ++k)
        if (!c->have_A || !c->surf[sids[k]].defined || !c->surf[sids[k]].has_rhs) {
            c->err = "emb_solve: needs emb_form_A and emb_surface_set_U(sid) first";
            return EMB_ERR_STATE;
        }
    const int nv = nrhs == 3 ? 4 : nrhs;      // a group of three is padded with a zero column
    if (nv > 1 && o->method != 2) { c->err = "emb_solve_multi: GMRES / BiCGStab take one right-hand side"; return EMB_ERR_ARG; }
    EMB_TRY(dev_alloc(c, c->bs, (size_t)c->Ns * NVMAX));
    EMB_TRY(dev_alloc(c, c->xs, (size_t)c->Ns * NVMAX));
    k_zero<<<blocks_for(c->Ns * nv, 256), 256, 0, c->stream>>>(c->Ns * nv, c->bs.p); EMB_LAUNCH_CHECK(c);
    for (int k = 0; k < nrhs; ++k) {
        Surface& s = c->surf[sids[k]];
        k_scatter_rhs<<<blocks_for(s.ndof, 128), 128, 0, c->stream>>>(s.ndof, s.dof.p, s.bval.p, c->newid.p, nv, k, c->bs.p);
        EMB_LAUNCH_CHECK(c);
    }
    if (o->use_x0) {
        EMB_TRY(dev_alloc(c, c->xfull, (size_t)c->N));
        k_zero<<<blocks_for(c->Ns * nv, 256), 256, 0, c->stream>>>(c->Ns * nv, c->xs.p); EMB_LAUNCH_CHECK(c);
        for (int k = 0; k < nrhs; ++k) {
            if (!x_full || !x_full[k]) continue;
            EMB_CUDA(c, cudaMemcpyAsync(c->xfull.p, x_full[k], (size_t)c->N * sizeof(cx), cudaMemcpyHostToDevice, c->stream));
            k_gather_col<<<blocks_for(c->Ns, 256), 256, 0, c->stream>>>(c->Ns, c->solve_ids.p, c->xfull.p, nv, k, c->xs.p);
            EMB_LAUNCH_CHECK(c);
        }
    }
    emb_solve_info tmp[NVMAX];
    c->nsol = nv;
    int rc = solve_dispatch(c, nv, o, c->bs.p, c->xs.p, tmp);
    if (infos) for (int k = 0; k < nrhs; ++k) infos[k] = tmp[k];
    if (rc < 0) return rc;
    for (int k = nrhs - 1; k >= 0; --k)       // column 0 last: it is the device-resident "last solution"
        if ((x_full && x_full[k]) || k == 0) EMB_TRY(finish_solution(c, k, x_full ? x_full[k] : nullptr));
    return rc;
}

extern "C" int emb_solve(emb_ctx* c, int sid, const emb_solve_opts* opts, emb_c128* x_full, emb_solve_info* info) {
    emb_c128* xs[1] = {x_full};
    return emb_solve_multi(c, 1, &sid, opts, xs, info);
}

// Field output mode.  on = 1: the host copies of emb_solve_multi's solutions are issued on a copy stream and overlap the
// following work; the buffers (pinned memory, or the copy degenerates to a synchronous one) are valid after
// emb_fields_sync().  Counterpart of `data._fields[port] = solution` (emfreq3d.py:699), which the reference does on the host.
extern "C" int emb_fields_async(emb_ctx* c, int on) {
    if (!c) return EMB_ERR_ARG;
    if (on && !c->copy_stream) EMB_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    if (!on && c->copy_stream) EMB_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    c->fields_async = on != 0;
    return EMB_OK;
}
extern "C" int emb_fields_sync(emb_ctx* c) {
    if (!c) return EMB_ERR_ARG;
    if (c->copy_stream) EMB_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    return EMB_OK;
}

// makes column k of the last emb_solve_multi the device-resident solution read by emb_interp_last
extern "C" int emb_select_solution(emb_ctx* c, int k) {
    if (!c || k < 0 || k >= c->nsol || !c->xs.p) return EMB_ERR_ARG;
    return finish_solution(c, k, nullptr);
}

extern "C" int emb_solve_rhs(emb_ctx* c, const emb_c128* b_full, const emb_solve_opts* opts, emb_c128* x_full,
                             emb_solve_info* info) {
    if (!c || !b_full) return EMB_ERR_ARG;
    if (!c->have_A) { c->err = "emb_solve_rhs: needs emb_form_A first"; return EMB_ERR_STATE; }
    const emb_solve_opts* o = opts ? opts : &kDefaultOpts;
    DevBuf<cx> bf;
    EMB_TRY(h2d(c, bf, reinterpret_cast<const cx*>(b_full), (size_t)c->N));
    EMB_TRY(dev_alloc(c, c->bs, (size_t)c->Ns * NVMAX));
    EMB_TRY(dev_alloc(c, c->xs, (size_t)c->Ns * NVMAX));
    k_gather<<<blocks_for(c->Ns, 256), 256, 0, c->stream>>>(c->Ns, c->solve_ids.p, bf.p, c->bs.p);
    EMB_LAUNCH_CHECK(c);
    if (o->use_x0 && x_full) {
        EMB_TRY(dev_alloc(c, c->xfull, (size_t)c->N));
        EMB_CUDA(c, cudaMemcpyAsync(c->xfull.p, x_full, (size_t)c->N * sizeof(cx), cudaMemcpyHostToDevice, c->stream));
        k_gather<<<blocks_for(c->Ns, 256), 256, 0, c->stream>>>(c->Ns, c->solve_ids.p, c->xfull.p, c->xs.p);
        EMB_LAUNCH_CHECK(c);
    }
    c->nsol = 1;
    int rc = solve_device<1>(c, o, c->bs.p, c->xs.p, info);
    bf.release();
    if (rc < 0) return rc;
    EMB_TRY(finish_solution(c, 0, x_full));
    return rc;
}

// ---- recycling control --------------------------------------------------------------------------
extern "C" int emb_recycle_config(emb_ctx* c, int max_vectors, double snapshot_rtol_factor) {
    if (!c || max_vectors < 0 || max_vectors > 256) return EMB_ERR_ARG;
    if (max_vectors > 0 && !c->have_dirichlet) { c->err = "emb_recycle_config: needs emb_set_dirichlet first"; return EMB_ERR_STATE; }
    if (max_vectors != c->rc_cap) {       // vectors are allocated when the first direction arrives (recycle.cuh::rc_prepare)
        c->rcU.release(); c->rcU32.release(); c->rcQ.release(); c->rc_part.release();
        c->rc_terms.clear();
        c->rc_terms.push_back(-1);        // never equal to a real term list => rc_prepare reallocates
    }
    c->rc_cap = max_vectors;
    c->rc_snap = (snapshot_rtol_factor > 0 && snapshot_rtol_factor <= 1) ? snapshot_rtol_factor : 0.3;
    rc_clear(c);
    return EMB_OK;
}
extern "C" int emb_recycle_info(emb_ctx* c, int* n, int64_t* spmvs, double* last_proj_relres) {
    if (!c) return EMB_ERR_ARG;
    if (n) *n = c->rc_n;
    if (spmvs) *spmvs = c->rc_spmvs;
    if (last_proj_relres) *last_proj_relres = c->rc_last_proj_relres;
    return EMB_OK;
}
extern "C" int64_t emb_recycle_accepted(const emb_ctx* c) { return c ? c->rc_accepted_total : 0; }
// Device-to-device exchange of directions between the ranks of a sharded sweep (the host side moves the buffers with
// NCCL over NVLink).  d_dst / d_src are DEVICE pointers to Ns complex128 values.  j = 0 is the NEWEST direction.
extern "C" int emb_recycle_export(emb_ctx* c, int j, void* d_dst) {
    if (!c || !d_dst || j < 0 || j >= c->rc_n) return EMB_ERR_ARG;
    EMB_CUDA(c, cudaMemcpyAsync(d_dst, rc_U(c, c->rc_n - 1 - j), (size_t)c->Ns * sizeof(cx), cudaMemcpyDeviceToDevice, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    return EMB_OK;
}
extern "C" int emb_recycle_import(emb_ctx* c, const void* d_src) {
    if (!c || !d_src) return EMB_ERR_ARG;
    if (c->rc_cap <= 0 || !c->have_A) { c->err = "emb_recycle_import: recycling not configured or no A(f)"; return EMB_ERR_STATE; }
    EMB_TRY(ensure_work(c, 10, 1));
    return rc_append(c, reinterpret_cast<const cx*>(d_src));
}

extern "C" int emb_spmv_host(emb_ctx* c, const emb_c128* x, emb_c128* y) {
    if (!c || !x || !y) return EMB_ERR_ARG;
    if (!c->have_A) { c->err = "emb_spmv_host: needs emb_form_A first"; return EMB_ERR_STATE; }
    DevBuf<cx> dx, dy;
    EMB_TRY(h2d(c, dx, reinterpret_cast<const cx*>(x), (size_t)c->Ns));
    EMB_TRY(dev_alloc(c, dy, (size_t)c->Ns));
    {
        PhaseTimer pt(c, "spmv");
        EMB_TRY(spmv1(c, c->A.p, dx.p, dy.p));
    }
    EMB_CUDA(c, cudaMemcpyAsync(y, dy.p, (size_t)c->Ns * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    dx.release(); dy.release();
    return EMB_OK;
}

// times `reps` operator applications on resident vectors.  nv: interleaved right-hand sides (1, 2, 4);
// fp32: the complex64 symmetric-part operator of the inner iteration instead of A(f) in complex128
extern "C" int emb_spmv_bench_ex(emb_ctx* c, int reps, int nv, int fp32, double* ms_per_spmv) {
    if (!c || reps <= 0 || !ms_per_spmv || (nv != 1 && nv != 2 && nv != 4)) return EMB_ERR_ARG;
    if (!c->have_A) { c->err = "emb_spmv_bench: needs emb_form_A first"; return EMB_ERR_STATE; }
    DevBuf<cx> dx, dy;
    EMB_TRY(dev_alloc(c, dx, (size_t)c->Ns * nv));
    EMB_TRY(dev_alloc(c, dy, (size_t)c->Ns * nv));
    EMB_CUDA(c, cudaMemsetAsync(dx.p, 0, (size_t)c->Ns * nv * sizeof(cx), c->stream));
    EMB_CUDA(c, cudaMemsetAsync(dy.p, 0, (size_t)c->Ns * nv * sizeof(cx), c->stream));
    cf* A32 = nullptr;
    if (fp32) {
        EMB_TRY(alloc_As32(c, &A32, nv));
        EMB_TRY(fill_As<cf>(c, A32));
        c->have_As = false;
    }
    auto one = [&](const cx* x, cx* y) -> int {
        if (fp32) {
            if (nv == 1) return apply_inner<1, cf, cx>(c, A32, x, y);
            if (nv == 2) return apply_inner<2, cf, cx>(c, A32, x, y);
            return apply_inner<4, cf, cx>(c, A32, x, y);
        }
        if (nv == 1) return spmv<1, cx>(c, c->A.p, x, y);
        if (nv == 2) return spmv<2, cx>(c, c->A.p, x, y);
        return spmv<4, cx>(c, c->A.p, x, y);
    };
    for (int i = 0; i < 3; ++i) EMB_TRY(one(dx.p, dy.p));
    {
        PhaseTimer pt(c, "spmv");
        for (int i = 0; i < reps; ++i) EMB_TRY(one((i & 1) ? dy.p : dx.p, (i & 1) ? dx.p : dy.p));
    }
    *ms_per_spmv = c->ms["spmv"] / reps;
    c->ms["spmv"] = *ms_per_spmv;
    if (getenv("EMB_SPMV_CHECK")) {      // tuning probe: weighted checksum of y = Op x for a fixed pseudo-random x (emb_last_ms "spmv_check_*")
        const size_t n = (size_t)c->Ns * nv;
        std::vector<cx> hx(n), hy(n);
        uint64_t st = 0x9e3779b97f4a7c15ull;
        auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (double)(st >> 11) / 9007199254740992.0 - 0.5; };
        for (size_t i = 0; i < n; ++i) hx[i] = cx{rnd(), rnd()};
        EMB_CUDA(c, cudaMemcpyAsync(dx.p, hx.data(), n * sizeof(cx), cudaMemcpyHostToDevice, c->stream));
        EMB_TRY(one(dx.p, dy.p));
        EMB_CUDA(c, cudaMemcpyAsync(hy.data(), dy.p, n * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
        EMB_CUDA(c, cudaStreamSynchronize(c->stream));
        double sr = 0, si = 0, sa = 0;
        for (size_t i = 0; i < n; ++i) {
            const double w = 0.5 + (double)((i * 2654435761ull) & 1023u) / 1024.0;
            sr += w * hy[i].re; si += w * hy[i].im; sa += hy[i].re * hy[i].re + hy[i].im * hy[i].im;
        }
        c->ms["spmv_check_re"] = sr; c->ms["spmv_check_im"] = si; c->ms["spmv_check_abs2"] = sa;
        // against the complex128 operator A(f) on the same x (As is its symmetric part rounded to complex64: ~1e-4 apart)
        {
            std::vector<cx> hz(n);
            int rc2 = nv == 1 ? spmv<1, cx>(c, c->A.p, dx.p, dy.p) : nv == 2 ? spmv<2, cx>(c, c->A.p, dx.p, dy.p) : spmv<4, cx>(c, c->A.p, dx.p, dy.p);
            if (rc2 < 0) return rc2;
            EMB_CUDA(c, cudaMemcpyAsync(hz.data(), dy.p, n * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
            EMB_CUDA(c, cudaStreamSynchronize(c->stream));
            double num = 0, den = 0;
            for (size_t i = 0; i < n; ++i) {
                num += (hy[i].re - hz[i].re) * (hy[i].re - hz[i].re) + (hy[i].im - hz[i].im) * (hy[i].im - hz[i].im);
                den += hz[i].re * hz[i].re + hz[i].im * hz[i].im;
            }
            c->ms["spmv_check_vs_A"] = den > 0 ? sqrt(num / den) : -1.0;
        }
    }
    dx.release(); dy.release();
    return EMB_OK;
}
extern "C" int emb_spmv_bench(emb_ctx* c, int reps, double* ms_per_spmv) { return emb_spmv_bench_ex(c, reps, 1, 0, ms_per_spmv); }

// ------------------------------------------------------------------------------------------------
// field evaluation at points with known host tets: per-point part of ned2_tet_interp (fem/mth/tet.py:371-497)
// ------------------------------------------------------------------------------------------------
__global__ void k_interp(int64_t npts, const int* __restrict__ tet, const double* __restrict__ xyz, const int* __restrict__ tetc,
                         const int* __restrict__ gid, const double* __restrict__ nodes, const cx* __restrict__ xfull,
                         cx* __restrict__ E) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= npts) return;
    const int t = tet[k];
    double p[4][3];
    for (int v = 0; v < 4; ++v) {
        const double* q = nodes + (int64_t)tetc[(int64_t)t * 4 + v] * 3;
        p[v][0] = q[0]; p[v][1] = q[1]; p[v][2] = q[2];
    }
    double e1[3], e2[3], e3[3], G[4][3];
    for (int a = 0; a < 3; ++a) { e1[a] = p[1][a] - p[0][a]; e2[a] = p[2][a] - p[0][a]; e3[a] = p[3][a] - p[0][a]; }
    G[1][0] = e2[1] * e3[2] - e2[2] * e3[1]; G[1][1] = e2[2] * e3[0] - e2[0] * e3[2]; G[1][2] = e2[0] * e3[1] - e2[1] * e3[0];
    G[2][0] = e3[1] * e1[2] - e3[2] * e1[1]; G[2][1] = e3[2] * e1[0] - e3[0] * e1[2]; G[2][2] = e3[0] * e1[1] - e3[1] * e1[0];
    G[3][0] = e1[1] * e2[2] - e1[2] * e2[1]; G[3][1] = e1[2] * e2[0] - e1[0] * e2[2]; G[3][2] = e1[0] * e2[1] - e1[1] * e2[0];
    for (int a = 0; a < 3; ++a) G[0][a] = -(G[1][a] + G[2][a] + G[3][a]);
    const double det = e1[0] * G[1][0] + e1[1] * G[1][1] + e1[2] * G[1][2];
    const double idet = 1.0 / det;
    double lam[4], grad[4][3];
    const double dx = xyz[k] - p[0][0], dy = xyz[npts + k] - p[0][1], dz = xyz[2 * npts + k] - p[0][2];
    for (int v = 1; v < 4; ++v) lam[v] = (G[v][0] * dx + G[v][1] * dy + G[v][2] * dz) * idet;
    lam[0] = 1.0 - lam[1] - lam[2] - lam[3];
    for (int v = 0; v < 4; ++v)
        for (int a = 0; a < 3; ++a) grad[v][a] = G[v][a] * idet;
    cx Ex = mk(0.0), Ey = mk(0.0), Ez = mk(0.0);
    // canonical functions: N = s l lam_X (lam_Q grad_P - lam_P grad_Q)
    const int eA[6] = {0, 0, 0, 1, 1, 2}, eB[6] = {1, 2, 3, 2, 3, 3};
    const int fA[4] = {0, 0, 0, 1}, fB[4] = {1, 1, 2, 2}, fE[4] = {2, 3, 3, 3};
    auto dist = [&](int a, int b) {
        const double x = p[a][0] - p[b][0], y = p[a][1] - p[b][1], z = p[a][2] - p[b][2];
        return sqrt(x * x + y * y + z * z);
    };
    auto add = [&](int cidx, double s, double l, int X, int P, int Q) {
        const cx c = xfull[gid[(int64_t)t * 20 + cidx]];
        const double f = s * l * lam[X];
        const double wx = lam[Q] * grad[P][0] - lam[P] * grad[Q][0];
        const double wy = lam[Q] * grad[P][1] - lam[P] * grad[Q][1];
        const double wz = lam[Q] * grad[P][2] - lam[P] * grad[Q][2];
        fma_r(Ex, f * wx, c); fma_r(Ey, f * wy, c); fma_r(Ez, f * wz, c);
    };
    for (int e = 0; e < 6; ++e) {
        const double l = dist(eA[e], eB[e]);
        add(e, 1.0, l, eA[e], eA[e], eB[e]);
        add(10 + e, 1.0, l, eB[e], eA[e], eB[e]);
    }
    for (int f = 0; f < 4; ++f) {
        add(6 + f, -1.0, dist(fA[f], fE[f]), fB[f], fA[f], fE[f]);
        add(16 + f, 1.0, dist(fA[f], fB[f]), fE[f], fA[f], fB[f]);
    }
    E[k] = Ex; E[npts + k] = Ey; E[2 * npts + k] = Ez;
}

static int interp_impl(emb_ctx* c, const cx* dxfull, int64_t npts, const int64_t* tet_ids, const double* xyz, emb_c128* E) {
    std::vector<int>& ht = c->itp_host;
    ht.resize((size_t)npts);
    for (int64_t i = 0; i < npts; ++i) {
        if (tet_ids[i] < 0 || tet_ids[i] >= c->nT) { c->err = "emb_interp: tet id out of range"; return EMB_ERR_ARG; }
        ht[i] = (int)tet_ids[i];
    }
    if (c->itp_tet.n < (size_t)npts) {
        EMB_TRY(dev_alloc(c, c->itp_tet, (size_t)npts));
        EMB_TRY(dev_alloc(c, c->itp_xyz, (size_t)npts * 3));
        EMB_TRY(dev_alloc(c, c->itp_E, (size_t)npts * 3));
    }
    EMB_CUDA(c, cudaMemcpyAsync(c->itp_tet.p, ht.data(), (size_t)npts * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    EMB_CUDA(c, cudaMemcpyAsync(c->itp_xyz.p, xyz, (size_t)npts * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_interp<<<blocks_for(npts, 128), 128, 0, c->stream>>>(npts, c->itp_tet.p, c->itp_xyz.p, c->tetc.p, c->gid.p, c->nodes.p,
                                                           dxfull, c->itp_E.p);
    EMB_LAUNCH_CHECK(c);
    EMB_CUDA(c, cudaMemcpyAsync(E, c->itp_E.p, (size_t)npts * 3 * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    return EMB_OK;
}

extern "C" int emb_interp(emb_ctx* c, const emb_c128* x_full, int64_t npts, const int64_t* tet_ids, const double* xyz,
                          emb_c128* E) {
    if (!c || !x_full || npts <= 0 || !tet_ids || !xyz || !E) return EMB_ERR_ARG;
    if (!c->have_mesh) { c->err = "emb_interp: mesh not uploaded"; return EMB_ERR_STATE; }
    DevBuf<cx> dx;
    EMB_TRY(h2d(c, dx, reinterpret_cast<const cx*>(x_full), (size_t)c->N));
    int rc = interp_impl(c, dx.p, npts, tet_ids, xyz, E);
    dx.release();
    return rc;
}

extern "C" int emb_interp_last(emb_ctx* c, int64_t npts, const int64_t* tet_ids, const double* xyz, emb_c128* E) {
    if (!c || npts <= 0 || !tet_ids || !xyz || !E) return EMB_ERR_ARG;
    if (!c->xfull.p) { c->err = "emb_interp_last: no solution on the device"; return EMB_ERR_STATE; }
    return interp_impl(c, c->xfull.p, npts, tet_ids, xyz, E);
}


extern "C" int emb_aux_clear(emb_ctx* c) {
    if (!c) return EMB_ERR_ARG;
    for (auto& a : c->aux) {
        a.rptr.release(); a.tptr.release(); a.rcol.release(); a.tcol.release(); a.rval.release(); a.tval.release();
        a.dinv.release(); a.tmp.release(); a.traw.release();
        amg_work_release(a.wk);
    }
    c->aux.clear();
    for (auto& st : c->side)          // stream priorities follow the shape of the tree: recreated at the next setup
        if (st) { cudaStreamDestroy(st); st = nullptr; }
    if (c->ev_fork) { cudaEventDestroy(c->ev_fork); c->ev_fork = nullptr; }
    for (auto& h : c->amg) amg_release(h);
    c->amg.clear();
    c->have_As = false;
    return EMB_OK;
}

extern "C" int emb_aux_add_ex(emb_ctx* c, int64_t nrow, int64_t ncol, const int64_t* Rp, const int32_t* Ri, const double* Rv,
                              const int64_t* Tp, const int32_t* Ti, const double* Tv, int parent, int solver, int hid,
                              int scale_mode) {
    if (!c || ncol <= 0 || nrow <= 0 || !Rp || !Ri || !Rv || !Tp || !Ti || !Tv) return EMB_ERR_ARG;
    if (!c->have_dirichlet) { c->err = "emb_aux_add: needs emb_set_dirichlet first (solve-space rows)"; return EMB_ERR_STATE; }
    const int na = (int)c->aux.size();
    if (na >= 16) { c->err = "emb_aux_add: at most 16 auxiliary spaces"; return EMB_ERR_LIMIT; }
    if (parent >= na || (parent < 0 && nrow != c->Ns) || (parent >= 0 && nrow != c->aux[parent].ncol)) {
        c->err = "emb_aux_add: row count does not match the parent space";
        return EMB_ERR_ARG;
    }
    if (solver == 0 && parent >= 0) { c->err = "emb_aux_add: the diagonal solver needs a top-level space"; return EMB_ERR_ARG; }
    if (solver == 1 && (hid < 0 || hid >= (int)c->amg.size() || c->amg[hid].lev.empty() || c->amg[hid].lev[0].n != ncol ||
                        !c->amg[hid].cinv.p)) {
        c->err = "emb_aux_add: AMG hierarchy missing, incomplete or of the wrong size";
        return EMB_ERR_ARG;
    }
    const int64_t nnz = Rp[nrow];
    if (Tp[ncol] != nnz) { c->err = "emb_aux_add: R and R^T disagree on nnz"; return EMB_ERR_ARG; }
    c->aux.emplace_back();
    AuxSpace& a = c->aux.back();
    a.ncol = ncol; a.nrow = nrow; a.nnz = nnz;
    a.parent = parent; a.solver = solver; a.hid = hid; a.scale_mode = scale_mode;
    EMB_TRY(h2d(c, a.rptr, Rp, (size_t)nrow + 1));
    EMB_TRY(h2d(c, a.rcol, reinterpret_cast<const int*>(Ri), (size_t)nnz));
    EMB_TRY(h2d(c, a.rval, Rv, (size_t)nnz));
    EMB_TRY(h2d(c, a.tptr, Tp, (size_t)ncol + 1));
    EMB_TRY(h2d(c, a.tcol, reinterpret_cast<const int*>(Ti), (size_t)nnz));
    EMB_TRY(h2d(c, a.tval, Tv, (size_t)nnz));
    EMB_TRY(dev_alloc(c, a.dinv, (size_t)ncol));
    EMB_TRY(dev_alloc(c, a.tmp, (size_t)ncol * NVMAX));
    EMB_TRY(dev_alloc(c, a.traw, (size_t)ncol * NVMAX));
    if (solver == 1) EMB_TRY(amg_work_alloc(c, c->amg[hid], a.wk));
    if (parent >= 0) c->aux[parent].has_children = true;
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->have_As = false;
    return EMB_OK;
}
extern "C" int emb_aux_add(emb_ctx* c, int64_t ncol, const int64_t* Rp, const int32_t* Ri, const double* Rv, const int64_t* Tp,
                           const int32_t* Ti, const double* Tv) {
    if (!c) return EMB_ERR_ARG;
    return emb_aux_add_ex(c, c->Ns, ncol, Rp, Ri, Rv, Tp, Ti, Tv, -1, 0, -1, 0);
}

extern "C" int emb_amg_create(emb_ctx* c, int* hid) {
    if (!c || !hid) return EMB_ERR_ARG;
    c->amg.emplace_back();
    *hid = (int)c->amg.size() - 1;
    return EMB_OK;
}
extern "C" int emb_amg_add_level(emb_ctx* c, int hid, int64_t n, const int64_t* Ap, const int32_t* Ai, const double* Av,
                                 const double* dinv, double omega, int64_t ncoarse, const int64_t* Pp, const int32_t* Pi,
                                 const double* Pv, const int64_t* Tp, const int32_t* Ti, const double* Tv) {
    if (!c || hid < 0 || hid >= (int)c->amg.size() || n <= 0) return EMB_ERR_ARG;
    AmgHierarchy& H = c->amg[hid];
    if (!H.lev.empty() && H.lev.back().nc != n) { c->err = "emb_amg_add_level: size does not match the previous level"; return EMB_ERR_ARG; }
    H.lev.emplace_back();
    AmgLevel& v = H.lev.back();
    v.n = n; v.nc = ncoarse; v.omega = omega;
    if (ncoarse > 0) {
        if (!Ap || !Ai || !Av || !dinv || !Pp || !Pi || !Pv || !Tp || !Ti || !Tv) return EMB_ERR_ARG;
        EMB_TRY(h2d(c, v.aptr, Ap, (size_t)n + 1));
        EMB_TRY(h2d(c, v.acol, reinterpret_cast<const int*>(Ai), (size_t)Ap[n]));
        EMB_TRY(h2d(c, v.aval, Av, (size_t)Ap[n]));
        EMB_TRY(h2d(c, v.dinv, dinv, (size_t)n));
        EMB_TRY(h2d(c, v.pptr, Pp, (size_t)n + 1));
        EMB_TRY(h2d(c, v.pcol, reinterpret_cast<const int*>(Pi), (size_t)Pp[n]));
        EMB_TRY(h2d(c, v.pval, Pv, (size_t)Pp[n]));
        EMB_TRY(h2d(c, v.tptr, Tp, (size_t)ncoarse + 1));
        EMB_TRY(h2d(c, v.tcol, reinterpret_cast<const int*>(Ti), (size_t)Tp[ncoarse]));
        EMB_TRY(h2d(c, v.tval, Tv, (size_t)Tp[ncoarse]));
        v.lpr_a = pick_lpr(Ap[n], n);
        v.lpr_p = pick_lpr(Pp[n], n);
        v.lpr_t = pick_lpr(Tp[ncoarse], ncoarse);
    }
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    return EMB_OK;
}
extern "C" int emb_amg_set_coarse_inverse(emb_ctx* c, int hid, int64_t n, const double* Ainv) {
    if (!c || hid < 0 || hid >= (int)c->amg.size() || n <= 0 || !Ainv) return EMB_ERR_ARG;
    AmgHierarchy& H = c->amg[hid];
    if (H.lev.empty() || H.lev.back().n != n || H.lev.back().nc != 0) {
        c->err = "emb_amg_set_coarse_inverse: the last level must be the coarsest (ncoarse = 0) and of size n";
        return EMB_ERR_ARG;
    }
    EMB_TRY(h2d(c, H.cinv, Ainv, (size_t)n * n));
    H.ncinv = n;
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    return EMB_OK;
}

extern "C" int emb_spmv_sampled(emb_ctx* c, double* avg_ms, int64_t* count) {
    if (!c || !avg_ms || !count) return EMB_ERR_ARG;
    *count = c->spmv_ms_cnt;
    *avg_ms = c->spmv_ms_cnt ? c->spmv_ms_sum / c->spmv_ms_cnt : 0.0;
    c->spmv_ms_sum = 0;
    c->spmv_ms_cnt = 0;
    return EMB_OK;
}
// same for the preconditioner applications sampled next to them
extern "C" int emb_precond_sampled(emb_ctx* c, double* avg_ms, int64_t* count) {
    if (!c || !avg_ms || !count) return EMB_ERR_ARG;
    *count = c->prec_ms_cnt;
    *avg_ms = c->prec_ms_cnt ? c->prec_ms_sum / c->prec_ms_cnt : 0.0;
    c->prec_ms_sum = 0;
    c->prec_ms_cnt = 0;
    return EMB_OK;
}
// solver switches (no reference counterpart): complex64 storage of the inner operator, concurrent auxiliary spaces
extern "C" int emb_solver_config(emb_ctx* c, int inner_fp32, int side_streams) {
    if (!c) return EMB_ERR_ARG;
    c->as_fp32 = inner_fp32 != 0;
    c->use_side_streams = side_streams != 0;
    c->have_As = false;
    return EMB_OK;
}
extern "C" int emb_solver_block(emb_ctx* c, int on) {
    if (!c) return EMB_ERR_ARG;
    c->block_krylov = on != 0;
    return EMB_OK;
}
// EXPERIMENTAL: the reduced basis as an extra coarse space of the preconditioner (default off).  Takes effect for the
// directions added after the call (emb_recycle_config clears the basis).
extern "C" int emb_solver_coarse_basis(emb_ctx* c, int on) {
    if (!c) return EMB_ERR_ARG;
    c->coarse_basis = on != 0;
    c->coarse_m = 0;
    c->coarse_version = -1;
    c->rc_terms.clear();
    c->rc_terms.push_back(-1);          // forces rc_prepare to (re)allocate with the bookkeeping matrices
    rc_clear(c);
    return EMB_OK;
}
extern "C" int64_t emb_graph_launch_count(const emb_ctx* c) { return c ? c->graph_launches : 0; }
