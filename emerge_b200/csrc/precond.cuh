// Additive multilevel preconditioner of the complex-symmetric inner operator As (symmetric part of A(f)):
//
//     M^-1 = D_blk^-1 + sum_k R_k S_k R_k^T        (tree of auxiliary spaces, emerge_b200/sweep.py::_setup_aux_spaces)
//
// D_blk: 2x2 blocks over the two functions of an edge / face; S_k: diagonal of R^T As R, or a V-cycle of a
// smoothed-aggregation hierarchy (amg.cuh).  The spaces are independent given the residual, so each one runs on its
// own side stream (fork / join with events); the contributions are added into z in a FIXED order on the parent's
// stream, so the result is bitwise identical to the serial schedule.  Vectors hold NV interleaved columns.
// The reference has no counterpart (it factorises, fem/solver.py:243-309).
#pragma once
#include "amg.cuh"

// one warp per row: As = (A + A^T)/2; the pattern is structurally symmetric.  blk: block layout of the pair-ordered solve
// space (entry e of row r goes to rowptr[2 (r/2)] + 4 (e/2) + 2 (r%2) + e%2, krylov.cuh::k_bspmv), else CSR position.
template <typename VT>
__global__ void k_sym_part(int64_t n, int blk, const int64_t* __restrict__ rowptr, const int* __restrict__ col,
                           const cx* __restrict__ A, VT* __restrict__ As) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const int64_t p0 = rowptr[r], pb = rowptr[r & ~(int64_t)1];
    for (int64_t k = p0 + lane; k < rowptr[r + 1]; k += 32) {
        const int j = col[k];
        int64_t lo = rowptr[j], hi = rowptr[j + 1] - 1;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (col[mid] < (int)r) lo = mid + 1; else hi = mid;
        }
        cx a = A[k];
        if (col[lo] == (int)r) { const cx b = A[lo]; a = cx{0.5 * (a.re + b.re), 0.5 * (a.im + b.im)}; }
        const int64_t e = k - p0;
        stval(As, blk ? pb + 4 * (e >> 1) + 2 * (r & 1) + (e & 1) : k, a);
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

template <typename VT>
__device__ __forceinline__ cx csr_get(const int64_t* rowptr, const int* col, const VT* val, int r, int cidx) {
    int64_t lo = rowptr[r], hi = rowptr[r + 1] - 1;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (col[mid] < cidx) lo = mid + 1; else hi = mid;
    }
    return (lo <= hi && col[lo] == cidx) ? ldval(val, lo) : cx{0, 0};
}

// dinv[2i], dinv[2i+1]: row i of the inverse 2x2 block (acting on (x_i, x_mate));  Jacobi: (1/a_ii, 0)
template <typename VT>
__global__ void k_precond_setup(int64_t ns, int mode, const int64_t* __restrict__ rowptr, const int* __restrict__ col,
                                const VT* __restrict__ val, const int* __restrict__ mate, cx* __restrict__ dinv) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= ns) return;
    const cx aii = csr_get(rowptr, col, val, (int)i, (int)i);
    const int m = (mode == 2) ? mate[i] : -1;
    if (mode == 0) { dinv[2 * i] = mk(1.0); dinv[2 * i + 1] = mk(0.0); return; }
    if (m < 0) { dinv[2 * i] = cdiv(mk(1.0), aii); dinv[2 * i + 1] = mk(0.0); return; }
    // the block of the SYMMETRIC part (val may be A(f) itself)
    const cx a1 = csr_get(rowptr, col, val, (int)i, m), a2 = csr_get(rowptr, col, val, m, (int)i);
    const cx aim = cx{0.5 * (a1.re + a2.re), 0.5 * (a1.im + a2.im)}, ami = aim;
    const cx amm = csr_get(rowptr, col, val, m, m);
    const cx det = aii * amm - aim * ami;
    dinv[2 * i] = cdiv(amm, det);
    dinv[2 * i + 1] = cdiv(-aim, det);
}
// z = D_blk^-1 r   (flat over ns * nv entries)
template <typename VX>
__global__ void k_precond_apply(int64_t ns, int nv, const cx* __restrict__ dinv, const int* __restrict__ mate,
                                const VX* __restrict__ r, VX* __restrict__ z) {
    int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (f >= ns * nv) return;
    const int64_t i = f / nv;
    const int k = (int)(f % nv);
    cx v = dinv[2 * i] * ldv(r, f);
    const int m = mate ? mate[i] : -1;
    if (m >= 0) fma_c(v, dinv[2 * i + 1], ldv(r, (int64_t)m * nv + k));
    stv(z, f, v);
}

// one warp per aux column k: d_k = sum_{i,j in supp(k)} R_ik A_ij R_jk, supp(k) = row k of R^T (sorted by i)
template <typename VT>
__global__ void __launch_bounds__(256) k_aux_diag(int64_t ncol, const int64_t* __restrict__ tptr, const int* __restrict__ tcol,
                                                  const double* __restrict__ tval, const int64_t* __restrict__ rowptr,
                                                  const int* __restrict__ col, const VT* __restrict__ A, cx* __restrict__ dinv) {
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
                const cx a = ldval(A, e);
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
// s = sum_i RT[k,i] r[i];  traw[k] = s (if traw);  t[k] = dinv ? dinv[k]*s : s     (KPR x NV lanes per aux column;
// the rows of R^T are short (5-30 entries), so few lanes with several independent loads each beat many idle lanes)
template <int NV, int KPR, typename VX>
__global__ void __launch_bounds__(256) k_aux_restrict(int64_t ncol, const int64_t* __restrict__ tptr, const int* __restrict__ tcol,
                                                      const double* __restrict__ tval, const cx* __restrict__ dinv,
                                                      const VX* __restrict__ r, VX* __restrict__ t, VX* __restrict__ traw) {
    constexpr int LPR = KPR * NV;
    const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t k = gt / LPR;
    const int s = (int)(gt % LPR);
    const int v = s % NV, ks = s / NV;
    double ar = 0.0, ai = 0.0;
    if (k < ncol) {
        const int64_t m1 = tptr[k + 1];
#pragma unroll 4
        for (int64_t m = tptr[k] + ks; m < m1; m += KPR) {
            const double w = __ldg(tval + m);
            const cx u = ldx(r + (int64_t)__ldg(tcol + m) * NV + v);
            ar += w * u.re;
            ai += w * u.im;
        }
    }
#pragma unroll
    for (int o = LPR / 2; o >= NV; o >>= 1) {
        ar += __shfl_down_sync(0xffffffffu, ar, o, LPR);
        ai += __shfl_down_sync(0xffffffffu, ai, o, LPR);
    }
    if (k < ncol && ks == 0) {
        const cx sum = cx{ar, ai};
        if (traw) stv(traw, k * NV + v, sum);
        if (t) stv(t, k * NV + v, dinv ? dinv[k] * sum : sum);
    }
}
template <int NV, typename VX>
static int aux_restrict_launch(emb_ctx* c, cudaStream_t s, const AuxSpace& a, const cx* dinv, const VX* src, VX* t, VX* traw) {
    const double avg = a.ncol > 0 ? (double)a.nnz / (double)a.ncol : 0.0;
    if (avg <= 16.0)
        k_aux_restrict<NV, 2, VX><<<blocks_for(a.ncol * 2 * NV, 256), 256, 0, s>>>(a.ncol, a.tptr.p, a.tcol.p, a.tval.p, dinv, src, t, traw);
    else if (avg <= 48.0 || NV == 4)
        k_aux_restrict<NV, 4, VX><<<blocks_for(a.ncol * 4 * NV, 256), 256, 0, s>>>(a.ncol, a.tptr.p, a.tcol.p, a.tval.p, dinv, src, t, traw);
    else
        k_aux_restrict<NV, 8, VX><<<blocks_for(a.ncol * 8 * NV, 256), 256, 0, s>>>(a.ncol, a.tptr.p, a.tcol.p, a.tval.p, dinv, src, t, traw);
    EMB_LAUNCH_CHECK(c);
    return EMB_OK;
}
// z[i] += s * sum_k R[i,k] t[k]     (NV threads per row; rows of R are short)
template <int NV, typename VX>
__global__ void __launch_bounds__(256) k_aux_prolong(int64_t n, const int64_t* __restrict__ rptr, const int* __restrict__ rcol,
                                                     const double* __restrict__ rval, const VX* __restrict__ t, cx s,
                                                     VX* __restrict__ z) {
    const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (f >= n * NV) return;
    const int64_t i = f / NV;
    const int v = (int)(f % NV);
    const int64_t m0 = rptr[i], m1 = rptr[i + 1];
    if (m0 == m1) return;
    double ar = 0.0, ai = 0.0;
    for (int64_t m = m0; m < m1; ++m) {
        const double w = rval[m];
        const cx u = ldx(t + (int64_t)rcol[m] * NV + v);
        ar += w * u.re;
        ai += w * u.im;
    }
    cx zi = ldv(z, f);
    fma_c(zi, s, cx{ar, ai});
    stv(z, f, zi);
}

// Top-level spaces whose prolongations are fused with the block-Jacobi term into the single pass that writes z
struct TopSpaces {
    int n;
    const int64_t* rptr[4];
    const int* rcol[4];
    const double* rval[4];
    const void* t[4];             // vectors in the storage type of the inner iteration
    cx scale[4];
};
// z = D_blk^-1 r + sum_s scale_s R_s t_s   (flat over ns * NV entries; terms added in the order of the list)
template <int NV, typename VX>
__global__ void __launch_bounds__(256) k_precond_final(int64_t ns, const cx* __restrict__ dinv, const int* __restrict__ mate,
                                                       const VX* __restrict__ r, TopSpaces T, VX* __restrict__ z) {
    const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (f >= ns * NV) return;
    const int64_t i = f / NV;
    const int v = (int)(f % NV);
    cx acc = dinv[2 * i] * ldv(r, f);
    const int m = mate ? mate[i] : -1;
    if (m >= 0) fma_c(acc, dinv[2 * i + 1], ldv(r, (int64_t)m * NV + v));
    for (int s = 0; s < T.n; ++s) {
        const int64_t m0 = T.rptr[s][i], m1 = T.rptr[s][i + 1];
        if (m0 == m1) continue;
        const VX* ts = static_cast<const VX*>(T.t[s]);
        double ar = 0.0, ai = 0.0;
        for (int64_t q = m0; q < m1; ++q) {
            const double w = __ldg(T.rval[s] + q);
            const cx u = ldx(ts + (int64_t)__ldg(T.rcol[s] + q) * NV + v);
            ar += w * u.re;
            ai += w * u.im;
        }
        fma_c(acc, T.scale[s], cx{ar, ai});
    }
    stv(z, f, acc);
}

// dst[i] += sum_s scale_s R_s t_s   (the children of one auxiliary space in a single pass over its vector)
template <int NV, typename VX>
__global__ void __launch_bounds__(256) k_aux_prolong_multi(int64_t n, TopSpaces T, VX* __restrict__ dst) {
    const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (f >= n * NV) return;
    const int64_t i = f / NV;
    const int v = (int)(f % NV);
    cx acc = ldv(dst, f);
    bool any = false;
    for (int s = 0; s < T.n; ++s) {
        const int64_t m0 = T.rptr[s][i], m1 = T.rptr[s][i + 1];
        if (m0 == m1) continue;
        const VX* ts = static_cast<const VX*>(T.t[s]);
        double ar = 0.0, ai = 0.0;
        for (int64_t q = m0; q < m1; ++q) {
            const double w = __ldg(T.rval[s] + q);
            const cx u = ldx(ts + (int64_t)__ldg(T.rcol[s] + q) * NV + v);
            ar += w * u.re;
            ai += w * u.im;
        }
        fma_c(acc, T.scale[s], cx{ar, ai});
        any = true;
    }
    if (any) stv(dst, f, acc);
}

// Side streams, one per auxiliary space (modulo NSIDE).  The chain restrict -> V-cycle -> prolong of the nodal spaces is a
// sequence of ~12 small latency-bound kernels; it gets the highest stream priority so that its blocks are scheduled ahead
// of the bulk restrictions of the childless top-level spaces (tens of thousands of blocks), which fill the machine around it.
static int precond_streams(emb_ctx* c) {
    if (c->side[0]) return EMB_OK;
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    for (int i = 0; i < emb_ctx::NSIDE; ++i) {
        bool bulk = true;           // every space mapped to this stream is a childless diagonal top-level space
        bool used = false;
        for (size_t k = (size_t)i; k < c->aux.size(); k += emb_ctx::NSIDE) {
            const AuxSpace& a = c->aux[k];
            used = true;
            if (a.parent >= 0 || a.has_children || a.solver != 0) bulk = false;
        }
        const int prio = (used && bulk) ? least : greatest;
        EMB_CUDA(c, cudaStreamCreateWithPriority(&c->side[i], cudaStreamNonBlocking, prio));
    }
    EMB_CUDA(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    return EMB_OK;
}
static int precond_events(emb_ctx* c) {
    const size_t na = c->aux.size();
    while (c->ev_restr.size() < na) {
        cudaEvent_t a = nullptr, b = nullptr;
        EMB_CUDA(c, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        EMB_CUDA(c, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        c->ev_restr.push_back(a);
        c->ev_done.push_back(b);
    }
    return EMB_OK;
}

template <typename VT>
static int precond_setup(emb_ctx* c, int mode_in, const VT* val) {
    const int mode = mode_in == 3 ? 2 : mode_in;
    if (mode_in == 3) {
        if (c->aux.empty()) { c->err = "precond=3 needs auxiliary spaces (emb_aux_add)"; return EMB_ERR_STATE; }
        for (auto& a : c->aux) {
            if (a.solver != 0) continue;
            k_aux_diag<VT><<<blocks_for(a.ncol * 32, 256), 256, 0, c->stream>>>(a.ncol, a.tptr.p, a.tcol.p, a.tval.p, c->rowptr_s.p,
                                                                                c->col_s.p, val, a.dinv.p);
            EMB_LAUNCH_CHECK(c);
        }
        EMB_TRY(precond_streams(c));
        EMB_TRY(precond_events(c));
    }
    EMB_TRY(dev_alloc(c, c->dinv, (size_t)c->Ns * 2));
    if (mode == 2 && !c->pairmate.p) {
        EMB_TRY(dev_alloc(c, c->pairmate, (size_t)c->Ns));
        k_pairmate<<<blocks_for(c->Ns, 256), 256, 0, c->stream>>>(c->Ns, c->solve_ids.p, c->newid.p, c->nE, c->nTri, c->pairmate.p);
        EMB_LAUNCH_CHECK(c);
    }
    k_precond_setup<VT><<<blocks_for(c->Ns, 256), 256, 0, c->stream>>>(c->Ns, mode, c->rowptr_s.p, c->col_s.p, val,
                                                                      mode == 2 ? c->pairmate.p : nullptr, c->dinv.p);
    EMB_LAUNCH_CHECK(c);
    return EMB_OK;
}

// z = M^-1 r on NV interleaved columns.  mode 3: block-Jacobi on the solve space plus the tree of auxiliary spaces
// (additive): restrict down the tree (parents before children), solve every space, prolong up; the prolongations of
// the top-level spaces are fused with the block-Jacobi term into the one pass that writes z.
template <int NV, typename VX>
static int precond_apply(emb_ctx* c, int mode, const VX* r, VX* z) {
    cudaStream_t main = c->stream;
    const int na = (mode == 3) ? (int)c->aux.size() : 0;
    const bool par = c->use_side_streams && na > 0;
    const int* mate = mode >= 2 ? c->pairmate.p : nullptr;
    if (na == 0) {
        k_precond_apply<VX><<<blocks_for(c->Ns * NV, 256), 256, 0, main>>>(c->Ns, NV, c->dinv.p, mate, r, z);
        EMB_LAUNCH_CHECK(c);
        return EMB_OK;
    }
    auto V = [](DevBuf<cx>& b) { return reinterpret_cast<VX*>(b.p); };      // complex128 allocations used as VX
    if (par) EMB_CUDA(c, cudaEventRecord(c->ev_fork, main));       // r is complete here
    auto strm = [&](int i) { return par ? c->side[i % emb_ctx::NSIDE] : main; };
    std::vector<VX*> xres((size_t)na, nullptr);
    for (int i = 0; i < na; ++i) {
        AuxSpace& a = c->aux[i];
        cudaStream_t s = strm(i);
        if (par) EMB_CUDA(c, cudaStreamWaitEvent(s, a.parent < 0 ? c->ev_fork : c->ev_restr[a.parent], 0));
        const VX* src = a.parent < 0 ? r : V(c->aux[a.parent].traw);
        VX* traw = (a.has_children || a.solver == 1) ? V(a.traw) : nullptr;
        VX* t = a.solver == 0 ? V(a.tmp) : nullptr;
        // solver 1 restricts straight into the right-hand side of its V-cycle
        if (a.solver == 1 && !a.has_children) traw = V(a.wk.b[0]);
        EMB_TRY((aux_restrict_launch<NV, VX>(c, s, a, a.solver == 0 ? a.dinv.p : nullptr, src, t, traw)));
        if (par) EMB_CUDA(c, cudaEventRecord(c->ev_restr[i], s));
        xres[i] = V(a.tmp);
        if (a.solver == 1) {
            AmgHierarchy& H = c->amg[a.hid];
            if (a.has_children)
                EMB_CUDA(c, cudaMemcpyAsync(a.wk.b[0].p, a.traw.p, (size_t)a.ncol * NV * sizeof(VX), cudaMemcpyDeviceToDevice, s));
            VX* res = nullptr;
            EMB_TRY((amg_vcycle<NV, VX>(c, s, H, a.wk, &res)));
            xres[i] = res;
        }
    }
    TopSpaces T;
    T.n = 0;
    int ntop = 0;
    for (int i = 0; i < na; ++i) ntop += c->aux[i].parent < 0;
    const bool fuse = ntop <= 4;
    if (!fuse) {
        k_precond_apply<VX><<<blocks_for(c->Ns * NV, 256), 256, 0, main>>>(c->Ns, NV, c->dinv.p, mate, r, z);
        EMB_LAUNCH_CHECK(c);
    }
    std::vector<TopSpaces> kids((size_t)na);
    for (auto& k : kids) k.n = 0;
    auto flush_kids = [&](int p) -> int {       // one pass over the parent's vector for up to 4 children
        TopSpaces& K = kids[(size_t)p];
        if (K.n == 0) return EMB_OK;
        AuxSpace& pa = c->aux[p];
        k_aux_prolong_multi<NV, VX><<<blocks_for(pa.ncol * NV, 256), 256, 0, strm(p)>>>(pa.ncol, K, V(pa.tmp));
        EMB_LAUNCH_CHECK(c);
        K.n = 0;
        return EMB_OK;
    };
    for (int i = na - 1; i >= 0; --i) {
        AuxSpace& a = c->aux[i];
        EMB_TRY(flush_kids(i));                  // every child of i has a larger index and was visited already
        cx scale = mk(1.0);
        if (a.solver == 1 && a.scale_mode == 1) scale = mk(-1.0 / (c->k0 * c->k0));
        cudaStream_t dst_s = a.parent < 0 ? main : strm(a.parent);
        if (par && dst_s != strm(i)) {
            EMB_CUDA(c, cudaEventRecord(c->ev_done[i], strm(i)));
            EMB_CUDA(c, cudaStreamWaitEvent(dst_s, c->ev_done[i], 0));
        }
        if (a.parent < 0 && fuse) {
            const int q = T.n++;
            T.rptr[q] = a.rptr.p; T.rcol[q] = a.rcol.p; T.rval[q] = a.rval.p; T.t[q] = xres[i]; T.scale[q] = scale;
            continue;
        }
        if (a.parent >= 0 && c->aux[a.parent].solver == 0) {
            TopSpaces& K = kids[(size_t)a.parent];
            const int q = K.n++;
            K.rptr[q] = a.rptr.p; K.rcol[q] = a.rcol.p; K.rval[q] = a.rval.p; K.t[q] = xres[i]; K.scale[q] = scale;
            if (K.n == 4) EMB_TRY(flush_kids(a.parent));
            continue;
        }
        VX* dst = a.parent < 0 ? z : V(c->aux[a.parent].tmp);
        k_aux_prolong<NV, VX><<<blocks_for(a.nrow * NV, 256), 256, 0, dst_s>>>(a.nrow, a.rptr.p, a.rcol.p, a.rval.p, xres[i], scale, dst);
        EMB_LAUNCH_CHECK(c);
    }
    if (fuse) {
        k_precond_final<NV, VX><<<blocks_for(c->Ns * NV, 256), 256, 0, main>>>(c->Ns, c->dinv.p, mate, r, T, z);
        EMB_LAUNCH_CHECK(c);
    }
    return EMB_OK;
}
