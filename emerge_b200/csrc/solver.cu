// Complex CSR SpMV and Krylov solvers on the solve space.
//
// Reference path replaced: fem/solver.py:405-469 (SolveRoutine.solve) with its direct solvers (:243-309).
// The reference eliminates prescribed dofs per solve, optionally reorders, factorises (SuperLU/PARDISO).
// Here the eliminated pattern is built once (operators.cu) and A(f) x = b is solved iteratively:
//   method 2 (default): COCR (conjugate-orthogonal conjugate-residual, one SpMV per iteration) on the
//       complex-symmetric part As = (A + A^T)/2, wrapped in defect correction on the true A
//       (x += As^-1 (b - A x)).  A(f) is not exactly symmetric because the reference's mass matrix is not
//       (fem/mth/tet.py:1036; relative asymmetry ~5e-5), so each correction gains ~3 digits.
//   method 0: restarted GMRES(m) on A (classical Gram-Schmidt with re-orthogonalisation).
//   method 1: BiCGStab on A.
// Preconditioners: Jacobi, or 2x2 block-Jacobi over the two functions of each edge / face.
// Kernels are HBM-bound: SpMV moves 20 B per nonzero + 36 B per row; vector updates are fused so an
// iteration of COCR is 4 launches and no host synchronisation (scalars live on the device, block partial
// sums are reduced in a fixed order by the consumer kernel => bitwise reproducible).
#include "context.cuh"
#include "amg.cuh"
#include <vector>

constexpr int NPART = 1024;          // block partials per reduction (fixed => deterministic)
constexpr int VBLOCK = 256;

// ------------------------------------------------------------------------------------------------
// SpMV: LPR lanes per row
// ------------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(256) k_spmv(int64_t n, const int64_t* __restrict__ rowptr, const int* __restrict__ col,
                                              const cx* __restrict__ val, const cx* __restrict__ x, cx* __restrict__ y) {
    const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t r = gt / LPR;
    const int sub = (int)(gt % LPR);
    double ar = 0, ai = 0;
    if (r < n) {
        const int64_t p0 = rowptr[r], p1 = rowptr[r + 1];
        for (int64_t k = p0 + sub; k < p1; k += LPR) {
            const double2 a = __ldg(reinterpret_cast<const double2*>(val + k));
            const double2 v = __ldg(reinterpret_cast<const double2*>(x + __ldg(col + k)));
            ar += a.x * v.x - a.y * v.y;
            ai += a.x * v.y + a.y * v.x;
        }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
        ar += __shfl_down_sync(0xffffffffu, ar, o, LPR);
        ai += __shfl_down_sync(0xffffffffu, ai, o, LPR);
    }
    if (r < n && sub == 0) *reinterpret_cast<double2*>(y + r) = make_double2(ar, ai);
}

static int spmv(emb_ctx* c, const cx* val, const cx* x, cx* y) {
    constexpr int LPR = 8;
    k_spmv<LPR><<<blocks_for(c->Ns * LPR, 256), 256, 0, c->stream>>>(c->Ns, c->rowptr_s.p, c->col_s.p, val, x, y);
    EMB_LAUNCH_CHECK(c);
    return EMB_OK;
}

// ------------------------------------------------------------------------------------------------
// deterministic reductions: each block writes one partial; consumers sum the NPART partials in order
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ cx block_sum(cx v) {
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

// sum NPART partials (all threads of the block get the result); fixed order
__device__ __forceinline__ cx sum_partials(const cx* __restrict__ part) {
    __shared__ cx s_tot;
    cx v = mk(0.0);
    for (int i = threadIdx.x; i < NPART; i += blockDim.x) v += part[i];
    cx t = block_sum(v);
    if (threadIdx.x == 0) s_tot = t;
    __syncthreads();
    return s_tot;
}

// partial[blockIdx] = sum over this block's grid-stride range of a_i * b_i (unconjugated) or conj(a_i)*b_i
template <bool CONJ>
__global__ void __launch_bounds__(VBLOCK) k_dot(int64_t n, const cx* __restrict__ a, const cx* __restrict__ b,
                                                cx* __restrict__ part) {
    cx acc = mk(0.0);
    const int64_t per = (n + NPART - 1) / NPART;
    const int64_t i0 = blockIdx.x * per, i1 = (i0 + per < n) ? i0 + per : n;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += VBLOCK) {
        const cx u = a[i], v = b[i];
        if (CONJ) { acc.re += u.re * v.re + u.im * v.im; acc.im += u.re * v.im - u.im * v.re; }
        else fma_c(acc, u, v);
    }
    cx t = block_sum(acc);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
}
__global__ void __launch_bounds__(VBLOCK) k_finish(const cx* __restrict__ part, cx* __restrict__ out) {
    cx t = sum_partials(part);
    if (threadIdx.x == 0) *out = t;
}

// ------------------------------------------------------------------------------------------------
// small vector kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_copy(int64_t n, const cx* __restrict__ a, cx* __restrict__ b) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) b[i] = a[i];
}
__global__ void k_zero(int64_t n, cx* __restrict__ a) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = cx{0, 0};
}
// y = a*x + b*y with device scalars sa[0]*fa, sb[0]*fb (null => 1)
__global__ void k_axpby(int64_t n, const cx* __restrict__ sa, double fa, const cx* __restrict__ x, const cx* __restrict__ sb,
                        double fb, cx* __restrict__ y) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const cx a = sa ? fa * (*sa) : mk(fa);
    if (!sb && fb == 0.0) { y[i] = a * x[i]; return; }      // do not read an uninitialised y
    const cx b = sb ? fb * (*sb) : mk(fb);
    y[i] = a * x[i] + b * y[i];
}
__global__ void k_gather(int64_t ns, const int* __restrict__ ids, const cx* __restrict__ full, cx* __restrict__ sub) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < ns) sub[i] = full[ids[i]];
}
__global__ void k_scatter(int64_t ns, const int* __restrict__ ids, const cx* __restrict__ sub, cx* __restrict__ full) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < ns) full[ids[i]] = sub[i];
}
__global__ void k_scatter_rhs(int64_t nd, const int* __restrict__ dof, const cx* __restrict__ bval, const int* __restrict__ newid,
                              cx* __restrict__ bs) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nd) return;
    const int s = newid[dof[i]];
    if (s >= 0) bs[s] = bval[i];
}

// ------------------------------------------------------------------------------------------------
// symmetric part and preconditioner setup
// ------------------------------------------------------------------------------------------------
// one warp per row: As[k] = (A[k] + A[k^T])/2; the pattern is structurally symmetric
__global__ void k_sym_part(int64_t n, const int64_t* __restrict__ rowptr, const int* __restrict__ col, const cx* __restrict__ A,
                           cx* __restrict__ As) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    for (int64_t k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32) {
        const int j = col[k];
        int64_t lo = rowptr[j], hi = rowptr[j + 1] - 1;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (col[mid] < (int)r) lo = mid + 1; else hi = mid;
        }
        cx a = A[k];
        if (col[lo] == (int)r) { const cx b = A[lo]; a = cx{0.5 * (a.re + b.re), 0.5 * (a.im + b.im)}; }
        As[k] = a;
    }
}

// mate[s] = solve-space index of the other function of the same edge/face (or -1)
__global__ void k_pairmate(int64_t ns, const int* __restrict__ solve_ids, const int* __restrict__ newid, int64_t nE, int64_t nTri,
                           int* __restrict__ mate) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= ns) return;
    const int64_t d = solve_ids[i];
    int64_t m;
    if (d < nE) m = d + nE + nTri;                    // edge-a -> edge-b   (fem/elements/nedelec2.py:46-50)
    else if (d < nE + nTri) m = d + nE + nTri;        // face-a -> face-b
    else m = d - nE - nTri;                           // b -> a
    mate[i] = newid[m];
}

__device__ __forceinline__ cx csr_get(const int64_t* rowptr, const int* col, const cx* val, int r, int cidx) {
    int64_t lo = rowptr[r], hi = rowptr[r + 1] - 1;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (col[mid] < cidx) lo = mid + 1; else hi = mid;
    }
    return (lo <= hi && col[lo] == cidx) ? val[lo] : cx{0, 0};
}

// dinv[2i], dinv[2i+1]: row i of the inverse 2x2 block (acting on (x_i, x_mate));  Jacobi: (1/a_ii, 0)
__global__ void k_precond_setup(int64_t ns, int mode, const int64_t* __restrict__ rowptr, const int* __restrict__ col,
                                const cx* __restrict__ val, const int* __restrict__ mate, cx* __restrict__ dinv) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= ns) return;
    const cx aii = csr_get(rowptr, col, val, (int)i, (int)i);
    const int m = (mode == 2) ? mate[i] : -1;
    if (mode == 0) { dinv[2 * i] = mk(1.0); dinv[2 * i + 1] = mk(0.0); return; }
    if (m < 0) { dinv[2 * i] = cdiv(mk(1.0), aii); dinv[2 * i + 1] = mk(0.0); return; }
    const cx aim = csr_get(rowptr, col, val, (int)i, m), ami = csr_get(rowptr, col, val, m, (int)i);
    const cx amm = csr_get(rowptr, col, val, m, m);
    const cx det = aii * amm - aim * ami;
    dinv[2 * i] = cdiv(amm, det);
    dinv[2 * i + 1] = cdiv(-aim, det);
}
__global__ void k_precond_apply(int64_t ns, const cx* __restrict__ dinv, const int* __restrict__ mate, const cx* __restrict__ r,
                                cx* __restrict__ z) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= ns) return;
    cx v = dinv[2 * i] * r[i];
    const int m = mate ? mate[i] : -1;
    if (m >= 0) fma_c(v, dinv[2 * i + 1], r[m]);
    z[i] = v;
}


// ------------------------------------------------------------------------------------------------
// auxiliary-space corrections  z += R diag(R^T A R)^-1 R^T r   (R real, CSR; R^T CSR)
// ------------------------------------------------------------------------------------------------
// one warp per aux column k: d_k = sum_{i,j in supp(k)} R_ik A_ij R_jk, supp(k) = row k of R^T (sorted by i)
__global__ void __launch_bounds__(256) k_aux_diag(int64_t ncol, const int64_t* __restrict__ tptr, const int* __restrict__ tcol,
                                                  const double* __restrict__ tval, const int64_t* __restrict__ rowptr,
                                                  const int* __restrict__ col, const cx* __restrict__ A, cx* __restrict__ dinv) {
    const int lane = threadIdx.x & 31;
    const int64_t k = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (k >= ncol) return;
    const int64_t p0 = tptr[k], p1 = tptr[k + 1];
    double ar = 0, ai = 0;
    for (int64_t m = p0; m < p1; ++m) {
        const int i = tcol[m];
        const double ri = tval[m];
        for (int64_t e = rowptr[i] + lane; e < rowptr[i + 1]; e += 32) {
            const int j = col[e];
            int64_t lo = p0, hi = p1 - 1;
            while (lo < hi) {
                int64_t mid = (lo + hi) >> 1;
                if (tcol[mid] < j) lo = mid + 1; else hi = mid;
            }
            if (tcol[lo] == j) {
                const double w = ri * tval[lo];
                const cx a = A[e];
                ar += w * a.re;
                ai += w * a.im;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ar += __shfl_down_sync(0xffffffffu, ar, o);
        ai += __shfl_down_sync(0xffffffffu, ai, o);
    }
    if (lane == 0) {
        const double n2 = ar * ar + ai * ai;
        dinv[k] = n2 > 0 ? cdiv(mk(1.0), cx{ar, ai}) : mk(0.0);
    }
}
// s = sum_i RT[k,i] r[i];  traw[k] = s (if traw);  t[k] = dinv ? dinv[k]*s : s     (8 lanes per aux column)
__global__ void __launch_bounds__(256) k_aux_restrict(int64_t ncol, const int64_t* __restrict__ tptr, const int* __restrict__ tcol,
                                                      const double* __restrict__ tval, const cx* __restrict__ dinv,
                                                      const cx* __restrict__ r, cx* __restrict__ t, cx* __restrict__ traw) {
    const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t k = gt >> 3;
    const int sub = (int)(gt & 7);
    double ar = 0, ai = 0;
    if (k < ncol)
        for (int64_t m = tptr[k] + sub; m < tptr[k + 1]; m += 8) {
            const double w = tval[m];
            const double2 v = __ldg(reinterpret_cast<const double2*>(r + tcol[m]));
            ar += w * v.x;
            ai += w * v.y;
        }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        ar += __shfl_down_sync(0xffffffffu, ar, o, 8);
        ai += __shfl_down_sync(0xffffffffu, ai, o, 8);
    }
    if (k < ncol && sub == 0) {
        if (traw) traw[k] = cx{ar, ai};
        if (t) t[k] = dinv ? dinv[k] * cx{ar, ai} : cx{ar, ai};
    }
}
// z[i] += s * sum_k R[i,k] t[k]     (thread per row; rows of R are short)
__global__ void __launch_bounds__(256) k_aux_prolong(int64_t n, const int64_t* __restrict__ rptr, const int* __restrict__ rcol,
                                                     const double* __restrict__ rval, const cx* __restrict__ t, cx s,
                                                     cx* __restrict__ z) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t m0 = rptr[i], m1 = rptr[i + 1];
    if (m0 == m1) return;
    double ar = 0, ai = 0;
    for (int64_t m = m0; m < m1; ++m) {
        const double w = rval[m];
        const double2 v = __ldg(reinterpret_cast<const double2*>(t + rcol[m]));
        ar += w * v.x;
        ai += w * v.y;
    }
    cx zi = z[i];
    fma_c(zi, s, cx{ar, ai});
    z[i] = zi;
}

// ------------------------------------------------------------------------------------------------
// fused COCR kernels (device scalars: sc[0]=zAz, sc[1]=alpha, sc[2]=beta, sc[3]=|r|^2, sc[4]=zAz_new)
// ------------------------------------------------------------------------------------------------
// alpha = zAz / sum(partA);  x += alpha p; r -= alpha Ap; z -= alpha MAp;  partial |r|^2
__global__ void __launch_bounds__(VBLOCK) k_cocr_update(int64_t n, const cx* __restrict__ partA, cx* __restrict__ sc,
                                                        const cx* __restrict__ p, const cx* __restrict__ Ap,
                                                        const cx* __restrict__ MAp, cx* __restrict__ x, cx* __restrict__ r,
                                                        cx* __restrict__ z, cx* __restrict__ partR) {
    const cx den = sum_partials(partA);
    const cx alpha = cdiv(sc[0], den);
    if (blockIdx.x == 0 && threadIdx.x == 0) sc[1] = alpha;
    cx acc = mk(0.0);
    const int64_t per = (n + NPART - 1) / NPART;
    const int64_t i0 = blockIdx.x * per, i1 = (i0 + per < n) ? i0 + per : n;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += VBLOCK) {
        cx xi = x[i], ri = r[i], zi = z[i];
        fma_c(xi, alpha, p[i]);
        const cx na = -alpha;
        fma_c(ri, na, Ap[i]);
        fma_c(zi, na, MAp[i]);
        x[i] = xi; r[i] = ri; z[i] = zi;
        acc.re += ri.re * ri.re + ri.im * ri.im;
    }
    cx t = block_sum(acc);
    if (threadIdx.x == 0) partR[blockIdx.x] = t;
}
// beta = sum(partZ)/zAz; zAz = sum(partZ); p = z + beta p; Ap = Az + beta Ap; also finishes |r|^2
__global__ void __launch_bounds__(VBLOCK) k_cocr_dir(int64_t n, const cx* __restrict__ partZ, const cx* __restrict__ partR,
                                                     cx* __restrict__ sc, const cx* __restrict__ z, const cx* __restrict__ Az,
                                                     cx* __restrict__ p, cx* __restrict__ Ap) {
    const cx znew = sum_partials(partZ);
    const cx beta = cdiv(znew, sc[0]);
    const cx rr = sum_partials(partR);
    const int64_t per = (n + NPART - 1) / NPART;
    const int64_t i0 = blockIdx.x * per, i1 = (i0 + per < n) ? i0 + per : n;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += VBLOCK) {
        cx pi = z[i], api = Az[i];
        fma_c(pi, beta, p[i]);
        fma_c(api, beta, Ap[i]);
        p[i] = pi; Ap[i] = api;
    }
    __syncthreads();
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { sc[2] = beta; sc[3] = rr; sc[4] = znew; }
}
// all blocks must have read sc[0] before it is overwritten -> separate tiny kernel
__global__ void k_cocr_commit(cx* sc) { sc[0] = sc[4]; }

// ------------------------------------------------------------------------------------------------
// host drivers
// ------------------------------------------------------------------------------------------------
struct Work {
    emb_ctx* c;
    int64_t n;
    std::vector<cx*> v;
};

static int ensure_work(emb_ctx* c, size_t count) {
    if (c->work.size() < count) c->work.resize(count);
    for (size_t i = 0; i < count; ++i) EMB_TRY(dev_alloc(c, c->work[i], (size_t)c->Ns));
    EMB_TRY(dev_alloc(c, c->red, (size_t)(4 * NPART + 16) * 2));
    return EMB_OK;
}

static int dot_host(emb_ctx* c, bool conj, const cx* a, const cx* b, cx* out) {
    cx* part = reinterpret_cast<cx*>(c->red.p);
    cx* sc = part + 4 * NPART;
    if (conj) k_dot<true><<<NPART, VBLOCK, 0, c->stream>>>(c->Ns, a, b, part);
    else k_dot<false><<<NPART, VBLOCK, 0, c->stream>>>(c->Ns, a, b, part);
    EMB_LAUNCH_CHECK(c);
    k_finish<<<1, VBLOCK, 0, c->stream>>>(part, sc + 8);
    EMB_LAUNCH_CHECK(c);
    EMB_CUDA(c, cudaMemcpyAsync(out, sc + 8, sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    return EMB_OK;
}

static int precond_setup(emb_ctx* c, int mode_in, const cx* val) {
    const int mode = mode_in == 3 ? 2 : mode_in;
    if (mode_in == 3) {
        if (c->aux.empty()) { c->err = "precond=3 needs auxiliary spaces (emb_aux_add)"; return EMB_ERR_STATE; }
        for (auto& a : c->aux) {
            if (a.solver != 0) continue;
            k_aux_diag<<<blocks_for(a.ncol * 32, 256), 256, 0, c->stream>>>(a.ncol, a.tptr.p, a.tcol.p, a.tval.p, c->rowptr_s.p,
                                                                            c->col_s.p, val, a.dinv.p);
            EMB_LAUNCH_CHECK(c);
        }
    }
    EMB_TRY(dev_alloc(c, c->dinv, (size_t)c->Ns * 2));
    if (mode == 2 && !c->pairmate.p) {
        EMB_TRY(dev_alloc(c, c->pairmate, (size_t)c->Ns));
        k_pairmate<<<blocks_for(c->Ns, 256), 256, 0, c->stream>>>(c->Ns, c->solve_ids.p, c->newid.p, c->nE, c->nTri, c->pairmate.p);
        EMB_LAUNCH_CHECK(c);
    }
    k_precond_setup<<<blocks_for(c->Ns, 256), 256, 0, c->stream>>>(c->Ns, mode, c->rowptr_s.p, c->col_s.p, val,
                                                                  mode == 2 ? c->pairmate.p : nullptr, c->dinv.p);
    EMB_LAUNCH_CHECK(c);
    return EMB_OK;
}
// z = M^-1 r.  mode 3: block-Jacobi on the solve space plus the tree of auxiliary spaces (additive):
// restrict down the tree (parents before children), solve every space (diagonal, or AMG V-cycle), prolong up.
static int precond_apply(emb_ctx* c, int mode, const cx* r, cx* z) {
    k_precond_apply<<<blocks_for(c->Ns, 256), 256, 0, c->stream>>>(c->Ns, c->dinv.p, mode >= 2 ? c->pairmate.p : nullptr, r, z);
    EMB_LAUNCH_CHECK(c);
    if (mode != 3) return EMB_OK;
    const int na = (int)c->aux.size();
    for (int i = 0; i < na; ++i) {
        AuxSpace& a = c->aux[i];
        const cx* src = a.parent < 0 ? r : c->aux[a.parent].traw.p;
        cx* traw = (a.has_children || a.solver == 1) ? a.traw.p : nullptr;
        cx* t = a.solver == 0 ? a.tmp.p : nullptr;
        k_aux_restrict<<<blocks_for(a.ncol * 8, 256), 256, 0, c->stream>>>(a.ncol, a.tptr.p, a.tcol.p, a.tval.p,
                                                                           a.solver == 0 ? a.dinv.p : nullptr, src, t, traw);
        EMB_LAUNCH_CHECK(c);
    }
    for (int i = na - 1; i >= 0; --i) {
        AuxSpace& a = c->aux[i];
        const cx* x = a.tmp.p;
        cx scale = mk(1.0);
        if (a.solver == 1) {
            AmgHierarchy& H = c->amg[a.hid];
            EMB_CUDA(c, cudaMemcpyAsync(H.lev[0].b.p, a.traw.p, (size_t)a.ncol * sizeof(cx), cudaMemcpyDeviceToDevice, c->stream));
            cx* res = nullptr;
            EMB_TRY(amg_vcycle(c, H, &res));
            x = res;
            if (a.scale_mode == 1) scale = mk(-1.0 / (c->k0 * c->k0));
        }
        cx* dst = a.parent < 0 ? z : c->aux[a.parent].tmp.p;
        k_aux_prolong<<<blocks_for(a.nrow, 256), 256, 0, c->stream>>>(a.nrow, a.rptr.p, a.rcol.p, a.rval.p, x, scale, dst);
        EMB_LAUNCH_CHECK(c);
    }
    return EMB_OK;
}

// COCR on As d = rhs, d starts at 0.  Stops when |r| <= stop_abs or maxit.  Returns iterations in *its.
static int cocr(emb_ctx* c, int pmode, const cx* As, const cx* rhs, cx* d, double stop_abs, int maxit, int* its,
                int* spmvs, double* rnorm_out) {
    const int64_t n = c->Ns;
    cx *r = c->work[0].p, *z = c->work[1].p, *p = c->work[2].p, *Az = c->work[3].p, *Ap = c->work[4].p, *MAp = c->work[5].p;
    cx* part = reinterpret_cast<cx*>(c->red.p);
    cx *partA = part, *partR = part + NPART, *partZ = part + 2 * NPART, *sc = part + 4 * NPART;
    const unsigned vb = blocks_for(n, 256);
    k_zero<<<vb, 256, 0, c->stream>>>(n, d); EMB_LAUNCH_CHECK(c);
    k_copy<<<vb, 256, 0, c->stream>>>(n, rhs, r); EMB_LAUNCH_CHECK(c);
    EMB_TRY(precond_apply(c, pmode, r, z));
    k_copy<<<vb, 256, 0, c->stream>>>(n, z, p); EMB_LAUNCH_CHECK(c);
    EMB_TRY(spmv(c, As, z, Az)); ++*spmvs;
    k_copy<<<vb, 256, 0, c->stream>>>(n, Az, Ap); EMB_LAUNCH_CHECK(c);
    k_dot<false><<<NPART, VBLOCK, 0, c->stream>>>(n, z, Az, partZ); EMB_LAUNCH_CHECK(c);
    k_finish<<<1, VBLOCK, 0, c->stream>>>(partZ, sc); EMB_LAUNCH_CHECK(c);
    int it = 0;
    double rn = 1e300;
    const int check = 10;
    while (it < maxit) {
        EMB_TRY(precond_apply(c, pmode, Ap, MAp));
        k_dot<false><<<NPART, VBLOCK, 0, c->stream>>>(n, Ap, MAp, partA); EMB_LAUNCH_CHECK(c);
        k_cocr_update<<<NPART, VBLOCK, 0, c->stream>>>(n, partA, sc, p, Ap, MAp, d, r, z, partR); EMB_LAUNCH_CHECK(c);
        const bool sample = ((it + 1) % check == 0);
        if (sample) cudaEventRecord(c->evs0, c->stream);
        EMB_TRY(spmv(c, As, z, Az)); ++*spmvs;
        if (sample) cudaEventRecord(c->evs1, c->stream);
        k_dot<false><<<NPART, VBLOCK, 0, c->stream>>>(n, z, Az, partZ); EMB_LAUNCH_CHECK(c);
        k_cocr_dir<<<NPART, VBLOCK, 0, c->stream>>>(n, partZ, partR, sc, z, Az, p, Ap); EMB_LAUNCH_CHECK(c);
        k_cocr_commit<<<1, 1, 0, c->stream>>>(sc); EMB_LAUNCH_CHECK(c);
        ++it;
        if (it % check == 0 || it == maxit) {
            cx h[5];
            EMB_CUDA(c, cudaMemcpyAsync(h, sc, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
            EMB_CUDA(c, cudaStreamSynchronize(c->stream));
            if (it % check == 0) {
                float sms = 0;
                if (cudaEventElapsedTime(&sms, c->evs0, c->evs1) == cudaSuccess) { c->spmv_ms_sum += sms; c->spmv_ms_cnt++; }
            }
            rn = sqrt(fabs(h[3].re));
            if (!(rn == rn)) { c->err = "COCR breakdown (NaN)"; *its = it; *rnorm_out = rn; return EMB_NOT_CONVERGED; }
            if (rn <= stop_abs) break;
        }
    }
    *its = it;
    *rnorm_out = rn;
    return EMB_OK;
}

// restarted GMRES on A with left... right preconditioning: A M^-1 u = b, x = M^-1 u
static int gmres(emb_ctx* c, int pmode, const cx* A, const cx* b, cx* x, double bnorm, const emb_solve_opts* o, int* its,
                 int* spmvs, double* relres) {
    const int64_t n = c->Ns;
    const int m = o->restart > 0 ? o->restart : 50;
    EMB_TRY(ensure_work(c, (size_t)m + 8));
    cx *r = c->work[0].p, *w = c->work[1].p, *t = c->work[2].p;
    auto V = [&](int j) { return c->work[6 + j].p; };
    const unsigned vb = blocks_for(n, 256);
    std::vector<cx> H((size_t)(m + 1) * m), cs(m), sn(m), g(m + 1), y(m);
    cx* dsc = reinterpret_cast<cx*>(c->red.p) + 4 * NPART + 10;
    int it = 0;
    double res = 1.0;
    auto set_scalar = [&](cx v) { return cudaMemcpyAsync(dsc, &v, sizeof(cx), cudaMemcpyHostToDevice, c->stream); };
    while (it < o->maxit) {
        // r = b - A x
        EMB_TRY(spmv(c, A, x, r)); ++*spmvs;
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
            EMB_TRY(precond_apply(c, pmode, V(j), t));
            EMB_TRY(spmv(c, A, t, w)); ++*spmvs;
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
        EMB_TRY(precond_apply(c, pmode, w, t));
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
    cx* dsc = reinterpret_cast<cx*>(c->red.p) + 4 * NPART + 10;
    auto set_scalar = [&](cx val) { return cudaMemcpyAsync(dsc, &val, sizeof(cx), cudaMemcpyHostToDevice, c->stream); };
    auto axpy = [&](cx a, const cx* xx, double bfac, cx* yy) -> int {
        EMB_CUDA(c, set_scalar(a));
        k_axpby<<<vb, 256, 0, c->stream>>>(n, dsc, 1.0, xx, nullptr, bfac, yy);
        EMB_LAUNCH_CHECK(c);
        EMB_CUDA(c, cudaStreamSynchronize(c->stream));
        return EMB_OK;
    };
    EMB_TRY(spmv(c, A, x, r)); ++*spmvs;
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
        EMB_TRY(precond_apply(c, pmode, p, ph));
        EMB_TRY(spmv(c, A, ph, v)); ++*spmvs;
        cx r0v;
        EMB_TRY(dot_host(c, true, r0, v, &r0v));
        alpha = cdiv(rho1, r0v);
        k_copy<<<vb, 256, 0, c->stream>>>(n, r, s); EMB_LAUNCH_CHECK(c);
        EMB_TRY(axpy(-alpha, v, 1.0, s));
        EMB_TRY(precond_apply(c, pmode, s, sh));
        EMB_TRY(spmv(c, A, sh, t)); ++*spmvs;
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
// Subspace recycling across frequency points (the sweep solves A(f) x = b(f) for a dense list of f).
// U holds previous Krylov corrections (any port); once per frequency C = A(f) U is re-formed and (C, U) are
// orthonormalised together (modified Gram-Schmidt on C, same column operations on U, so C = A U keeps holding).
// For a right-hand side b the start vector x0 = sum_i <c_i, b - A x> u_i minimises the residual over span(U);
// the Krylov method then only has to remove what the recycled space cannot represent.  The true residual is
// always recomputed from A afterwards, so the accuracy contract (relres <= rtol in FP64) is unchanged.
// The reference has no counterpart (it refactorises at every frequency, fem/solver.py:243-309).
// ------------------------------------------------------------------------------------------------
constexpr int RC_NP = 256;       // partials per dot product of the batched Gram-Schmidt kernels

// part[k][blockIdx.x] = partial of <c_k, v> over this block's range, k = blockIdx.y (column k = slot (head+k)%cap)
__global__ void __launch_bounds__(VBLOCK) k_rc_dots(int64_t n, const cx* __restrict__ Cb, int head, int cap,
                                                    const cx* __restrict__ v, cx* __restrict__ part) {
    const cx* ck = Cb + (int64_t)((head + blockIdx.y) % cap) * n;
    cx acc = mk(0.0);
    const int64_t per = (n + RC_NP - 1) / RC_NP;
    const int64_t i0 = blockIdx.x * per, i1 = (i0 + per < n) ? i0 + per : n;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += VBLOCK) {
        const cx u = ck[i], w = v[i];
        acc.re += u.re * w.re + u.im * w.im;
        acc.im += u.re * w.im - u.im * w.re;
    }
    cx t = block_sum(acc);
    if (threadIdx.x == 0) part[(int64_t)blockIdx.y * RC_NP + blockIdx.x] = t;
}
// coef[k] = sum of the RC_NP partials of column k in fixed order (one block per column)
__global__ void __launch_bounds__(VBLOCK) k_rc_coef(const cx* __restrict__ part, cx* __restrict__ coef) {
    cx v = part[(int64_t)blockIdx.x * RC_NP + threadIdx.x];     // RC_NP == VBLOCK
    cx t = block_sum(v);
    if (threadIdx.x == 0) coef[blockIdx.x] = t;
}
// cj -= sum_k coef[k] c_k ; uj -= sum_k coef[k] u_k   (k < m, columns counted from `head`)
__global__ void __launch_bounds__(256) k_rc_sub(int64_t n, int m, const cx* __restrict__ coef, const cx* __restrict__ Cb,
                                                const cx* __restrict__ Ub, int head, int cap, cx* __restrict__ cj,
                                                cx* __restrict__ uj) {
    extern __shared__ cx s_h[];
    for (int k = threadIdx.x; k < m; k += blockDim.x) s_h[k] = coef[k];
    __syncthreads();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    cx a = cj[i], b = uj[i];
    for (int k = 0; k < m; ++k) {
        const int64_t off = (int64_t)((head + k) % cap) * n + i;
        const cx h = -s_h[k];
        fma_c(a, h, Cb[off]);
        fma_c(b, h, Ub[off]);
    }
    cj[i] = a; uj[i] = b;
}
// x += sum_k coef[k] u_k
__global__ void __launch_bounds__(256) k_rc_combine(int64_t n, int m, const cx* __restrict__ coef, const cx* __restrict__ Ub,
                                                    int head, int cap, cx* __restrict__ x) {
    extern __shared__ cx s_h[];
    for (int k = threadIdx.x; k < m; k += blockDim.x) s_h[k] = coef[k];
    __syncthreads();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    cx a = x[i];
    for (int k = 0; k < m; ++k) fma_c(a, s_h[k], Ub[(int64_t)((head + k) % cap) * n + i]);
    x[i] = a;
}
// cj, uj *= 1/sqrt(sum(part).re); the norm goes to norm_out[0] (0 => vector zeroed)
__global__ void __launch_bounds__(VBLOCK) k_rc_scale(int64_t n, const cx* __restrict__ part, cx* __restrict__ cj,
                                                     cx* __restrict__ uj, double* __restrict__ norm_out) {
    const cx t = sum_partials(part);
    const double nrm = sqrt(fabs(t.re));
    const double s = nrm > 0 ? 1.0 / nrm : 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out) *norm_out = nrm;
    const int64_t per = (n + NPART - 1) / NPART;
    const int64_t i0 = blockIdx.x * per, i1 = (i0 + per < n) ? i0 + per : n;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += VBLOCK) {
        cj[i] = s * cj[i];
        uj[i] = s * uj[i];
    }
}

static inline cx* rc_U(emb_ctx* c, int j) { return c->rcU.p + (int64_t)((c->rc_head + j) % c->rc_cap) * c->Ns; }
static inline cx* rc_C(emb_ctx* c, int j) { return c->rcC.p + (int64_t)((c->rc_head + j) % c->rc_cap) * c->Ns; }

static void rc_clear(emb_ctx* c) {
    c->rc_n = 0;
    c->rc_head = 0;
    c->rc_C_valid = false;
}

// classical Gram-Schmidt of column j of (C, U) against columns [0, j) (batched: one pass over the 2j vectors),
// `passes` times, then normalisation; *norm_dev receives the norm
static int rc_orth_column(emb_ctx* c, int j, int passes, double* norm_dev) {
    const int64_t n = c->Ns;
    cx* part = c->rc_part.p;
    cx* coef = part + (int64_t)c->rc_cap * RC_NP;
    cx *Cj = rc_C(c, j), *Uj = rc_U(c, j);
    if (j > 0)
        for (int pass = 0; pass < passes; ++pass) {
            k_rc_dots<<<dim3(RC_NP, j), VBLOCK, 0, c->stream>>>(n, c->rcC.p, c->rc_head, c->rc_cap, Cj, part); EMB_LAUNCH_CHECK(c);
            k_rc_coef<<<j, VBLOCK, 0, c->stream>>>(part, coef); EMB_LAUNCH_CHECK(c);
            k_rc_sub<<<blocks_for(n, 256), 256, j * sizeof(cx), c->stream>>>(n, j, coef, c->rcC.p, c->rcU.p, c->rc_head, c->rc_cap, Cj, Uj);
            EMB_LAUNCH_CHECK(c);
        }
    cx* pn = reinterpret_cast<cx*>(c->red.p) + 3 * NPART;
    k_dot<true><<<NPART, VBLOCK, 0, c->stream>>>(n, Cj, Cj, pn); EMB_LAUNCH_CHECK(c);
    k_rc_scale<<<NPART, VBLOCK, 0, c->stream>>>(n, pn, Cj, Uj, norm_dev); EMB_LAUNCH_CHECK(c);
    return EMB_OK;
}

// C = A(f) U, orthonormalised, newest direction first (once per frequency)
static int rc_rebuild(emb_ctx* c) {
    for (int j = 0; j < c->rc_n; ++j) {
        const bool sample = (j == c->rc_n / 2) && !c->rc_sample_pending;     // one timed SpMV per frequency point
        if (sample) cudaEventRecord(c->evr0, c->stream);
        EMB_TRY(spmv(c, c->A.p, rc_U(c, j), rc_C(c, j)));
        if (sample) { cudaEventRecord(c->evr1, c->stream); c->rc_sample_pending = true; }
        c->rc_spmvs++;
        EMB_TRY(rc_orth_column(c, j, 1, nullptr));
    }
    c->rc_C_valid = true;
    return EMB_OK;
}

// xs += U C^H r
static int rc_project(emb_ctx* c, const cx* r, cx* xs) {
    const int64_t n = c->Ns;
    const int m = c->rc_n;
    cx* part = c->rc_part.p;
    cx* coef = part + (int64_t)c->rc_cap * RC_NP;
    k_rc_dots<<<dim3(RC_NP, m), VBLOCK, 0, c->stream>>>(n, c->rcC.p, c->rc_head, c->rc_cap, r, part); EMB_LAUNCH_CHECK(c);
    k_rc_coef<<<m, VBLOCK, 0, c->stream>>>(part, coef); EMB_LAUNCH_CHECK(c);
    k_rc_combine<<<blocks_for(n, 256), 256, m * sizeof(cx), c->stream>>>(n, m, coef, c->rcU.p, c->rc_head, c->rc_cap, xs);
    EMB_LAUNCH_CHECK(c);
    return EMB_OK;
}

// add direction d (or d - minus) to the recycled space as its NEWEST member.  It is orthogonalised against the
// current members first, so the member dropped when the ring is full (the last in Gram-Schmidt order) only carries
// what was unique to the oldest direction.
static int rc_append(emb_ctx* c, const cx* d, const cx* minus) {
    const int64_t n = c->Ns;
    const unsigned vb = blocks_for(n, 256);
    if (!c->rc_C_valid && c->rc_n > 0) EMB_TRY(rc_rebuild(c));
    const int m = c->rc_n < c->rc_cap ? c->rc_n : c->rc_cap - 1;     // members kept
    // stage the candidate in the slot just before the head (it becomes column 0 if accepted); when the ring is
    // full that slot is the oldest member's, which is the one being replaced
    const int slot = (c->rc_head - 1 + c->rc_cap) % c->rc_cap;
    cx *Un = c->rcU.p + (int64_t)slot * n, *Cn = c->rcC.p + (int64_t)slot * n;
    if (minus) {
        k_copy<<<vb, 256, 0, c->stream>>>(n, minus, Un); EMB_LAUNCH_CHECK(c);
        k_axpby<<<vb, 256, 0, c->stream>>>(n, nullptr, 1.0, d, nullptr, -1.0, Un); EMB_LAUNCH_CHECK(c);
    } else {
        k_copy<<<vb, 256, 0, c->stream>>>(n, d, Un); EMB_LAUNCH_CHECK(c);
    }
    EMB_TRY(spmv(c, c->A.p, Un, Cn));
    c->rc_spmvs++;
    double* nd = c->red.p + (size_t)(4 * NPART + 16) * 2 - 4;     // scalar area: |A d|^2 (cx) then the final norm
    cx* pn = reinterpret_cast<cx*>(c->red.p) + 3 * NPART;
    k_dot<true><<<NPART, VBLOCK, 0, c->stream>>>(n, Cn, Cn, pn); EMB_LAUNCH_CHECK(c);
    k_finish<<<1, VBLOCK, 0, c->stream>>>(pn, reinterpret_cast<cx*>(nd)); EMB_LAUNCH_CHECK(c);
    // orthogonalise against members 0..m-1 (two passes: the candidate may be nearly inside the space)
    {
        cx* part = c->rc_part.p;
        cx* coef = part + (int64_t)c->rc_cap * RC_NP;
        if (m > 0)
            for (int pass = 0; pass < 2; ++pass) {
                k_rc_dots<<<dim3(RC_NP, m), VBLOCK, 0, c->stream>>>(n, c->rcC.p, c->rc_head, c->rc_cap, Cn, part); EMB_LAUNCH_CHECK(c);
                k_rc_coef<<<m, VBLOCK, 0, c->stream>>>(part, coef); EMB_LAUNCH_CHECK(c);
                k_rc_sub<<<vb, 256, m * sizeof(cx), c->stream>>>(n, m, coef, c->rcC.p, c->rcU.p, c->rc_head, c->rc_cap, Cn, Un);
                EMB_LAUNCH_CHECK(c);
            }
        k_dot<true><<<NPART, VBLOCK, 0, c->stream>>>(n, Cn, Cn, pn); EMB_LAUNCH_CHECK(c);
        k_rc_scale<<<NPART, VBLOCK, 0, c->stream>>>(n, pn, Cn, Un, nd + 2); EMB_LAUNCH_CHECK(c);
    }
    double h[3];
    EMB_CUDA(c, cudaMemcpyAsync(h, nd, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    const double before = sqrt(fabs(h[0])), after = h[2];
    if (!(after > 1e-9 * before) || !(after == after)) {     // numerically inside the space: not kept
        if (c->rc_n == c->rc_cap) c->rc_n = c->rc_cap - 1;     // its slot was overwritten by the candidate
        return EMB_OK;
    }
    c->rc_head = slot;
    c->rc_n = m + 1;
    return EMB_OK;
}

// solves A xs = bs (device, solve space); xs is in/out (initial guess when use_x0)
static int solve_device(emb_ctx* c, const emb_solve_opts* o, const cx* bs, cx* xs, emb_solve_info* info) {
    const int64_t n = c->Ns;
    const unsigned vb = blocks_for(n, 256);
    EMB_TRY(ensure_work(c, 10));
    cudaEventRecord(c->ev0, c->stream);
    int its = 0, spmvs = 0;
    double relres = 1.0;
    cx bb;
    EMB_TRY(dot_host(c, true, bs, bs, &bb));
    const double bnorm = sqrt(bb.re);
    int rc = EMB_OK;
    const bool recycle = c->rc_cap > 0 && bnorm > 0;
    c->rc_last_proj_relres = -1;
    if (bnorm > 0 && !o->use_x0) { k_zero<<<vb, 256, 0, c->stream>>>(n, xs); EMB_LAUNCH_CHECK(c); }
    if (recycle) {
        EMB_TRY(dev_alloc(c, c->rc_x0, (size_t)n));
        if (c->rc_n > 0) {
            if (!c->rc_C_valid) EMB_TRY(rc_rebuild(c));
            const cx* r0 = bs;
            if (o->use_x0) {    // residual of the caller's guess
                cx* t = c->work[0].p;
                EMB_TRY(spmv(c, c->A.p, xs, t)); ++spmvs;
                k_axpby<<<vb, 256, 0, c->stream>>>(n, nullptr, 1.0, bs, nullptr, -1.0, t); EMB_LAUNCH_CHECK(c);
                r0 = t;
            }
            EMB_TRY(rc_project(c, r0, xs));
        }
        k_copy<<<vb, 256, 0, c->stream>>>(n, xs, c->rc_x0.p); EMB_LAUNCH_CHECK(c);
    }
    const bool have_guess = o->use_x0 || (recycle && c->rc_n > 0);
    const double target = recycle ? o->rtol * c->rc_snap : o->rtol;
    emb_solve_opts ot = *o;
    ot.rtol = target;
    if (bnorm == 0) {
        k_zero<<<vb, 256, 0, c->stream>>>(n, xs); EMB_LAUNCH_CHECK(c);
        relres = 0;
    } else if (o->method == 0) {
        EMB_TRY(precond_setup(c, o->precond, c->A.p));
        EMB_TRY(gmres(c, o->precond, c->A.p, bs, xs, bnorm, &ot, &its, &spmvs, &relres));
    } else if (o->method == 1) {
        EMB_TRY(precond_setup(c, o->precond, c->A.p));
        EMB_TRY(bicgstab(c, o->precond, c->A.p, bs, xs, bnorm, &ot, &its, &spmvs, &relres));
    } else {
        // defect correction on A with COCR on the symmetric part
        struct { cx* p; } rr{c->work[8].p}, dd{c->work[9].p};      // persistent workspace (no per-solve cudaMalloc)
        DevBuf<cx>& As = c->As;
        // the symmetric part and the preconditioner are only built when a point really has to iterate
        auto ensure_operator = [&]() -> int {
            if (!c->have_As) {
                EMB_TRY(dev_alloc(c, As, (size_t)c->nnz_s));
                k_sym_part<<<blocks_for(n * 32, 256), 256, 0, c->stream>>>(n, c->rowptr_s.p, c->col_s.p, c->A.p, As.p);
                EMB_LAUNCH_CHECK(c);
                EMB_TRY(precond_setup(c, o->precond, As.p));
                c->have_As = true;
                c->As_precond = o->precond;
            } else if (c->As_precond != o->precond) {
                EMB_TRY(precond_setup(c, o->precond, As.p));
                c->As_precond = o->precond;
            }
            return EMB_OK;
        };
        double prev = 1e300;
        for (int outer = 0; outer < 30 && its < o->maxit; ++outer) {
            double rn = bnorm;
            if (outer > 0 || have_guess) {
                EMB_TRY(spmv(c, c->A.p, xs, rr.p)); ++spmvs;
                k_axpby<<<vb, 256, 0, c->stream>>>(n, nullptr, 1.0, bs, nullptr, -1.0, rr.p); EMB_LAUNCH_CHECK(c);
                cx r2;
                EMB_TRY(dot_host(c, true, rr.p, rr.p, &r2));
                rn = sqrt(r2.re);
            } else {
                k_copy<<<vb, 256, 0, c->stream>>>(n, bs, rr.p); EMB_LAUNCH_CHECK(c);
            }
            relres = rn / bnorm;
            if (outer == 0 && recycle && c->rc_n > 0) c->rc_last_proj_relres = relres;
            // a start vector that already meets rtol is accepted as is; once the solve has to iterate it feeds the
            // recycled space and is run to the tighter snapshot tolerance
            if (relres <= (outer == 0 ? o->rtol : target)) break;
            if (outer > 2 && rn > 0.5 * prev) { /* stagnation of the correction: keep going, but it is visible in info */ }
            prev = rn;
            // inner target: two digits below the current residual, never below what the outer loop needs
            static const double inner_red = getenv("EMB_INNER") ? atof(getenv("EMB_INNER")) : 1e-2;
            double stop = inner_red * rn;
            const double need = 0.3 * target * bnorm;
            if (stop < need) stop = need;
            int iit = 0;
            double irn = 0;
            EMB_TRY(ensure_operator());
            rc = cocr(c, o->precond, As.p, rr.p, dd.p, stop, o->maxit - its, &iit, &spmvs, &irn);
            its += iit;
            static const bool verbose = getenv("EMB_VERBOSE") != nullptr;
            if (verbose) fprintf(stderr, "[emb] outer %d relres %.3e -> inner %d its, inner residual %.3e (target %.3e)\n", outer, relres, iit, irn / bnorm, stop / bnorm);
            if (rc < 0) break;
            k_axpby<<<vb, 256, 0, c->stream>>>(n, nullptr, 1.0, dd.p, nullptr, 1.0, xs); EMB_LAUNCH_CHECK(c);
            if (rc == EMB_NOT_CONVERGED) break;
        }
        if (rc >= 0) {
            EMB_TRY(spmv(c, c->A.p, xs, rr.p)); ++spmvs;
            k_axpby<<<vb, 256, 0, c->stream>>>(n, nullptr, 1.0, bs, nullptr, -1.0, rr.p); EMB_LAUNCH_CHECK(c);
            cx r2;
            EMB_TRY(dot_host(c, true, rr.p, rr.p, &r2));
            relres = sqrt(r2.re) / bnorm;
        }
        if (rc < 0) return rc;
    }
    if (o->method != 2 && bnorm > 0) {   // true residual at exit
        cx* t = c->work[0].p;
        EMB_TRY(spmv(c, c->A.p, xs, t)); ++spmvs;
        k_axpby<<<vb, 256, 0, c->stream>>>(n, nullptr, 1.0, bs, nullptr, -1.0, t); EMB_LAUNCH_CHECK(c);
        cx r2;
        EMB_TRY(dot_host(c, true, t, t, &r2));
        relres = sqrt(r2.re) / bnorm;
    }
    if (recycle && its > 0 && relres <= 1e2 * o->rtol) EMB_TRY(rc_append(c, xs, c->rc_x0.p));
    cudaEventRecord(c->ev1, c->stream);
    cudaEventSynchronize(c->ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->ms["solve"] = ms;
    if (c->rc_sample_pending) {
        float sms = 0;
        if (cudaEventElapsedTime(&sms, c->evr0, c->evr1) == cudaSuccess) { c->spmv_ms_sum += sms; c->spmv_ms_cnt++; }
        c->rc_sample_pending = false;
    }
    if (info) { info->iters = its; info->relres = relres; info->ms = ms; info->spmvs = spmvs; }
    if (!(relres <= o->rtol)) {
        c->err = "solver did not reach rtol: relres=" + std::to_string(relres) + " after " + std::to_string(its) + " iterations";
        return EMB_NOT_CONVERGED;
    }
    return EMB_OK;
}

static int finish_solution(emb_ctx* c, emb_c128* x_full) {
    EMB_TRY(dev_alloc(c, c->xfull, (size_t)c->N));
    EMB_CUDA(c, cudaMemsetAsync(c->xfull.p, 0, (size_t)c->N * sizeof(cx), c->stream));
    k_scatter<<<blocks_for(c->Ns, 256), 256, 0, c->stream>>>(c->Ns, c->solve_ids.p, c->xs.p, c->xfull.p);
    EMB_LAUNCH_CHECK(c);
    if (x_full) EMB_CUDA(c, cudaMemcpyAsync(x_full, c->xfull.p, (size_t)c->N * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    return EMB_OK;
}

static const emb_solve_opts kDefaultOpts = {2, 2, 50, 100000, 1e-8, 0};

extern "C" int emb_solve(emb_ctx* c, int sid, const emb_solve_opts* opts, emb_c128* x_full, emb_solve_info* info) {
    if (!c || sid < 0 || sid >= 16) return EMB_ERR_ARG;
    if (!c->have_A || !c->surf[sid].defined || !c->surf[sid].has_rhs) {
        c->err = "emb_solve: needs emb_form_A and emb_surface_set_U(sid) first";
        return EMB_ERR_STATE;
    }
    const emb_solve_opts* o = opts ? opts : &kDefaultOpts;
    Surface& s = c->surf[sid];
    DevBuf<cx>& bs = c->bs;
    EMB_TRY(dev_alloc(c, bs, (size_t)c->Ns));
    EMB_TRY(dev_alloc(c, c->xs, (size_t)c->Ns));
    k_zero<<<blocks_for(c->Ns, 256), 256, 0, c->stream>>>(c->Ns, bs.p); EMB_LAUNCH_CHECK(c);
    k_scatter_rhs<<<blocks_for(s.ndof, 128), 128, 0, c->stream>>>(s.ndof, s.dof.p, s.bval.p, c->newid.p, bs.p);
    EMB_LAUNCH_CHECK(c);
    if (o->use_x0 && x_full) {
        EMB_TRY(dev_alloc(c, c->xfull, (size_t)c->N));
        EMB_CUDA(c, cudaMemcpyAsync(c->xfull.p, x_full, (size_t)c->N * sizeof(cx), cudaMemcpyHostToDevice, c->stream));
        k_gather<<<blocks_for(c->Ns, 256), 256, 0, c->stream>>>(c->Ns, c->solve_ids.p, c->xfull.p, c->xs.p);
        EMB_LAUNCH_CHECK(c);
    }
    int rc = solve_device(c, o, bs.p, c->xs.p, info);
    if (rc < 0) return rc;
    EMB_TRY(finish_solution(c, x_full));
    return rc;
}

extern "C" int emb_solve_rhs(emb_ctx* c, const emb_c128* b_full, const emb_solve_opts* opts, emb_c128* x_full,
                             emb_solve_info* info) {
    if (!c || !b_full) return EMB_ERR_ARG;
    if (!c->have_A) { c->err = "emb_solve_rhs: needs emb_form_A first"; return EMB_ERR_STATE; }
    const emb_solve_opts* o = opts ? opts : &kDefaultOpts;
    DevBuf<cx> bf;
    DevBuf<cx>& bs = c->bs;
    EMB_TRY(h2d(c, bf, reinterpret_cast<const cx*>(b_full), (size_t)c->N));
    EMB_TRY(dev_alloc(c, bs, (size_t)c->Ns));
    EMB_TRY(dev_alloc(c, c->xs, (size_t)c->Ns));
    k_gather<<<blocks_for(c->Ns, 256), 256, 0, c->stream>>>(c->Ns, c->solve_ids.p, bf.p, bs.p);
    EMB_LAUNCH_CHECK(c);
    if (o->use_x0 && x_full) {
        EMB_TRY(dev_alloc(c, c->xfull, (size_t)c->N));
        EMB_CUDA(c, cudaMemcpyAsync(c->xfull.p, x_full, (size_t)c->N * sizeof(cx), cudaMemcpyHostToDevice, c->stream));
        k_gather<<<blocks_for(c->Ns, 256), 256, 0, c->stream>>>(c->Ns, c->solve_ids.p, c->xfull.p, c->xs.p);
        EMB_LAUNCH_CHECK(c);
    }
    int rc = solve_device(c, o, bs.p, c->xs.p, info);
    bf.release();
    if (rc < 0) return rc;
    EMB_TRY(finish_solution(c, x_full));
    return rc;
}


// ---- recycling control --------------------------------------------------------------------------
extern "C" int emb_recycle_config(emb_ctx* c, int max_vectors, double snapshot_rtol_factor) {
    if (!c || max_vectors < 0 || max_vectors > 256) return EMB_ERR_ARG;
    if (max_vectors > 0 && !c->have_dirichlet) { c->err = "emb_recycle_config: needs emb_set_dirichlet first"; return EMB_ERR_STATE; }
    c->rcU.release(); c->rcC.release(); c->rc_part.release();
    c->rc_cap = max_vectors;
    c->rc_snap = (snapshot_rtol_factor > 0 && snapshot_rtol_factor <= 1) ? snapshot_rtol_factor : 0.1;
    rc_clear(c);
    if (max_vectors > 0) {
        EMB_TRY(dev_alloc(c, c->rcU, (size_t)max_vectors * c->Ns));
        EMB_TRY(dev_alloc(c, c->rcC, (size_t)max_vectors * c->Ns));
        EMB_TRY(dev_alloc(c, c->rc_part, (size_t)max_vectors * (RC_NP + 1)));
        EMB_TRY(ensure_work(c, 10));
    }
    return EMB_OK;
}
extern "C" int emb_recycle_info(emb_ctx* c, int* n, int64_t* spmvs, double* last_proj_relres) {
    if (!c) return EMB_ERR_ARG;
    if (n) *n = c->rc_n;
    if (spmvs) *spmvs = c->rc_spmvs;
    if (last_proj_relres) *last_proj_relres = c->rc_last_proj_relres;
    return EMB_OK;
}
// Device-to-device exchange of recycled directions between the ranks of a sharded sweep (the host side moves the
// buffers with NCCL over NVLink).  d_dst / d_src are DEVICE pointers to Ns complex128 values.
extern "C" int emb_recycle_export(emb_ctx* c, int j, void* d_dst) {
    if (!c || !d_dst || j < 0 || j >= c->rc_n) return EMB_ERR_ARG;
    EMB_CUDA(c, cudaMemcpyAsync(d_dst, rc_U(c, j), (size_t)c->Ns * sizeof(cx), cudaMemcpyDeviceToDevice, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    return EMB_OK;
}
extern "C" int emb_recycle_import(emb_ctx* c, const void* d_src) {
    if (!c || !d_src) return EMB_ERR_ARG;
    if (c->rc_cap <= 0 || !c->have_A) { c->err = "emb_recycle_import: recycling not configured or no A(f)"; return EMB_ERR_STATE; }
    return rc_append(c, reinterpret_cast<const cx*>(d_src), nullptr);
}

extern "C" int emb_spmv_host(emb_ctx* c, const emb_c128* x, emb_c128* y) {
    if (!c || !x || !y) return EMB_ERR_ARG;
    if (!c->have_A) { c->err = "emb_spmv_host: needs emb_form_A first"; return EMB_ERR_STATE; }
    DevBuf<cx> dx, dy;
    EMB_TRY(h2d(c, dx, reinterpret_cast<const cx*>(x), (size_t)c->Ns));
    EMB_TRY(dev_alloc(c, dy, (size_t)c->Ns));
    {
        PhaseTimer pt(c, "spmv");
        EMB_TRY(spmv(c, c->A.p, dx.p, dy.p));
    }
    EMB_CUDA(c, cudaMemcpyAsync(y, dy.p, (size_t)c->Ns * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    dx.release(); dy.release();
    return EMB_OK;
}

extern "C" int emb_spmv_bench(emb_ctx* c, int reps, double* ms_per_spmv) {
    if (!c || reps <= 0 || !ms_per_spmv) return EMB_ERR_ARG;
    if (!c->have_A) { c->err = "emb_spmv_bench: needs emb_form_A first"; return EMB_ERR_STATE; }
    DevBuf<cx> dx, dy;
    EMB_TRY(dev_alloc(c, dx, (size_t)c->Ns));
    EMB_TRY(dev_alloc(c, dy, (size_t)c->Ns));
    EMB_CUDA(c, cudaMemsetAsync(dx.p, 0, (size_t)c->Ns * sizeof(cx), c->stream));
    for (int i = 0; i < 3; ++i) EMB_TRY(spmv(c, c->A.p, dx.p, dy.p));
    {
        PhaseTimer pt(c, "spmv");
        for (int i = 0; i < reps; ++i) EMB_TRY(spmv(c, c->A.p, (i & 1) ? dy.p : dx.p, (i & 1) ? dx.p : dy.p));
    }
    *ms_per_spmv = c->ms["spmv"] / reps;
    c->ms["spmv"] = *ms_per_spmv;
    dx.release(); dy.release();
    return EMB_OK;
}

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
    std::vector<int> ht((size_t)npts);
    for (int64_t i = 0; i < npts; ++i) {
        if (tet_ids[i] < 0 || tet_ids[i] >= c->nT) { c->err = "emb_interp: tet id out of range"; return EMB_ERR_ARG; }
        ht[i] = (int)tet_ids[i];
    }
    DevBuf<int> dt;
    DevBuf<double> dp;
    DevBuf<cx> dE;
    EMB_TRY(h2d(c, dt, ht.data(), (size_t)npts));
    EMB_TRY(h2d(c, dp, xyz, (size_t)npts * 3));
    EMB_TRY(dev_alloc(c, dE, (size_t)npts * 3));
    k_interp<<<blocks_for(npts, 128), 128, 0, c->stream>>>(npts, dt.p, dp.p, c->tetc.p, c->gid.p, c->nodes.p, dxfull, dE.p);
    EMB_LAUNCH_CHECK(c);
    EMB_CUDA(c, cudaMemcpyAsync(E, dE.p, (size_t)npts * 3 * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    dt.release(); dp.release(); dE.release();
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
    }
    c->aux.clear();
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
    EMB_TRY(dev_alloc(c, a.tmp, (size_t)ncol));
    EMB_TRY(dev_alloc(c, a.traw, (size_t)ncol));
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
    EMB_TRY(dev_alloc(c, v.b, (size_t)n));
    EMB_TRY(dev_alloc(c, v.xa, (size_t)n));
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
        EMB_TRY(dev_alloc(c, v.xb, (size_t)n));
        EMB_TRY(dev_alloc(c, v.t, (size_t)n));
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
