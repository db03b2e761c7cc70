// Top-level auxiliary spaces of the multilevel preconditioner built on the device: the transfer matrices
//   G (solve space x kept P2 nodal functions: gradients of the quadratic Lagrange space = kernel of the curl-curl matrix)
//   P (solve space x kept edges: the lowest-order Whitney space)
// restricted to the kept dofs, rows in solve-index order, columns touched by an eliminated dof dropped, plus their
// transposes - what emerge_b200/auxspace.py::build_aux_spaces_paired assembles with numpy (3-4 s at 1M tets, the critical
// path of the host-side setup).  Same closed-form entries (auxspace.py header: small integers / edge lengths, no
// quadrature) in the same arithmetic, so the two builders agree bit for bit (tests/test_gpu_auxbuild.py).
// The reference has no counterpart (it factorises A(f), fem/solver.py:243-309).
// Kernels: mark the columns eliminated rows touch; CUB scan -> compact column numbering; count / fill the rows (<= 6
// entries each, sorted in registers); transpose by one stable radix sort of (column, row) keys.  All integer / streaming
// work, ~30 M entries at 1M tets.
#include "context.cuh"
#include <cub/cub.cuh>

constexpr int AUXB_NVMAX = 4;     // = NVMAX of krylov.cuh (widest lockstep group): columns of the per-space work vectors
struct AuxTab {             // (u, w) pairs of auxspace._face_tables(): vertex, edge (P2) and Whitney targets
    double vt[3][2], et[3][2], wt[3][2];
};
struct AuxRow {
    int n;
    int col[6];
    double val[6];
};

__device__ __forceinline__ double edge_len(const double* __restrict__ nodes, int a, int b) {
    const double dx = nodes[3 * (int64_t)b] - nodes[3 * (int64_t)a], dy = nodes[3 * (int64_t)b + 1] - nodes[3 * (int64_t)a + 1],
                 dz = nodes[3 * (int64_t)b + 2] - nodes[3 * (int64_t)a + 2];
    return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}
// row of G (which = 0) or P (which = 1) of full dof d, columns in the un-compacted numbering (G: nodes then edges)
__device__ __forceinline__ void aux_row(int which, int64_t d, int64_t nE, int64_t nTri, int64_t nN, const int64_t* __restrict__ edges,
                                        const int* __restrict__ tris, const int* __restrict__ tri2f, const double* __restrict__ nodes,
                                        const AuxTab& tb, AuxRow& r) {
    const int64_t H = nE + nTri;
    const int fb = d >= H ? 1 : 0;              // second function of the entity
    const int64_t e = d - (fb ? H : 0);
    r.n = 0;
    if (e < nE) {
        const int A = (int)edges[e], B = (int)edges[nE + e];
        const double il = 1.0 / edge_len(nodes, A, B);
        if (which == 0) {
            r.n = 3;
            r.col[0] = A; r.col[1] = B; r.col[2] = (int)(nN + e);
            if (!fb) { r.val[0] = 3 * il; r.val[1] = il; r.val[2] = -4 * il; }
            else { r.val[0] = -il; r.val[1] = -3 * il; r.val[2] = 4 * il; }
        } else {
            r.n = 1;
            r.col[0] = (int)e;
            r.val[0] = il;
        }
        return;
    }
    const int64_t f = e - nE;
    const int v[3] = {tris[3 * f], tris[3 * f + 1], tris[3 * f + 2]};
    const int te[3] = {tri2f[f], tri2f[nTri + f], tri2f[2 * nTri + f]};      // edges (A,B), (B,E), (A,E)
    // the length that scales function a is |A E|, function b |A B| (auxspace.py: u = c_fa l_AE, w = c_fb l_AB)
    const double li = 1.0 / (fb ? edge_len(nodes, (int)edges[te[0]], (int)edges[nE + te[0]])
                                : edge_len(nodes, (int)edges[te[2]], (int)edges[nE + te[2]]));
    if (which == 0) {
        r.n = 6;
        for (int k = 0; k < 3; ++k) {
            r.col[k] = v[k];            r.val[k] = tb.vt[k][fb] * li;
            r.col[3 + k] = (int)(nN + te[k]); r.val[3 + k] = tb.et[k][fb] * li;
        }
    } else {
        r.n = 3;
        for (int k = 0; k < 3; ++k) { r.col[k] = te[k]; r.val[k] = tb.wt[k][fb] * li; }
    }
}

__global__ void k_auxb_mark(int64_t N, int which, const int* __restrict__ newid, int64_t nE, int64_t nTri, int64_t nN,
                            const int64_t* __restrict__ edges, const int* __restrict__ tris, const int* __restrict__ tri2f,
                            const double* __restrict__ nodes, AuxTab tb, int* __restrict__ good) {
    const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (d >= N || newid[d] >= 0) return;
    AuxRow r;
    aux_row(which, d, nE, nTri, nN, edges, tris, tri2f, nodes, tb, r);
    for (int k = 0; k < r.n; ++k)
        if (r.val[k] != 0.0) good[r.col[k]] = 0;
}
template <bool FILL>
__global__ void k_auxb_rows(int64_t Ns, int which, const int* __restrict__ solve_ids, int64_t nE, int64_t nTri, int64_t nN,
                            const int64_t* __restrict__ edges, const int* __restrict__ tris, const int* __restrict__ tri2f,
                            const double* __restrict__ nodes, AuxTab tb, const int* __restrict__ good, const int* __restrict__ newcol,
                            int64_t* __restrict__ cnt, const int64_t* __restrict__ rptr, int* __restrict__ rcol, double* __restrict__ rval,
                            unsigned long long* __restrict__ tkey, int* __restrict__ colcnt) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= Ns) return;
    AuxRow r;
    aux_row(which, solve_ids[i], nE, nTri, nN, edges, tris, tri2f, nodes, tb, r);
    int c[6];
    double v[6];
    int n = 0;
    for (int k = 0; k < r.n; ++k)
        if (r.val[k] != 0.0 && good[r.col[k]]) { c[n] = newcol[r.col[k]]; v[n] = r.val[k]; ++n; }
    if (!FILL) { cnt[i] = n; return; }
    for (int a = 1; a < n; ++a) {               // ascending columns
        const int cc = c[a];
        const double vv = v[a];
        int b = a - 1;
        while (b >= 0 && c[b] > cc) { c[b + 1] = c[b]; v[b + 1] = v[b]; --b; }
        c[b + 1] = cc; v[b + 1] = vv;
    }
    const int64_t p = rptr[i];
    for (int k = 0; k < n; ++k) {
        rcol[p + k] = c[k];
        rval[p + k] = v[k];
        tkey[p + k] = ((unsigned long long)(unsigned)c[k] << 32) | (unsigned long long)(unsigned)i;
        atomicAdd(colcnt + c[k], 1);
    }
}
__global__ void k_auxb_tcol(int64_t nnz, const unsigned long long* __restrict__ skey, int* __restrict__ tcol) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < nnz) tcol[k] = (int)(skey[k] & 0xffffffffull);
}
__global__ void k_auxb_i2l(int64_t n, const int* __restrict__ in, int64_t* __restrict__ out) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) out[k] = in[k];
}
__global__ void k_auxb_fill(int* __restrict__ v, int64_t n, int val) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) v[k] = val;
}
__global__ void k_auxb_flags(int64_t n, const int* __restrict__ good, unsigned char* __restrict__ bad) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) bad[k] = good[k] ? 0 : 1;
}

// which: 0 = G, 1 = P.  Appends the space as a top-level space with the diagonal solver (as emb_aux_add does) when it has
// at least one column.  bad_out (optional): 1 for every dropped column of the un-compacted numbering (G: nN + nE, P: nE).
extern "C" int emb_aux_build_top(emb_ctx* c, int which, const int64_t* edges_2xnE, const double* face_tables_18, int64_t* ncol_out,
                                 unsigned char* bad_out) {
    if (!c || which < 0 || which > 1 || !edges_2xnE || !face_tables_18 || !ncol_out) return EMB_ERR_ARG;
    if (!c->have_dirichlet || !c->have_mesh) { c->err = "emb_aux_build_top: needs emb_upload_mesh and emb_set_dirichlet first"; return EMB_ERR_STATE; }
    if ((int)c->aux.size() >= 16) { c->err = "emb_aux_build_top: at most 16 auxiliary spaces"; return EMB_ERR_LIMIT; }
    PhaseTimer pt(c, which ? "aux_build_P" : "aux_build_G");
    const int64_t nE = c->nE, nTri = c->nTri, nN = c->nN, N = c->N, Ns = c->Ns;
    const int64_t ncol0 = which ? nE : nN + nE;
    AuxTab tb;
    memcpy(&tb, face_tables_18, sizeof(AuxTab));
    DevBuf<int64_t> edges, cnt;
    DevBuf<int> good, newcol, colcnt;
    DevBuf<char> tmp;
    EMB_TRY(h2d(c, edges, edges_2xnE, (size_t)nE * 2));
    EMB_TRY(dev_alloc(c, good, (size_t)ncol0 + 1));
    EMB_TRY(dev_alloc(c, newcol, (size_t)ncol0 + 1));
    k_auxb_fill<<<blocks_for(ncol0 + 1, 256), 256, 0, c->stream>>>(good.p, ncol0 + 1, 1);
    EMB_LAUNCH_CHECK(c);
    k_auxb_mark<<<blocks_for(N, 256), 256, 0, c->stream>>>(N, which, c->newid.p, nE, nTri, nN, edges.p, c->tris.p, c->tri2f.p,
                                                          c->nodes.p, tb, good.p);
    EMB_LAUNCH_CHECK(c);
    size_t tb1 = 0;
    EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tb1, good.p, newcol.p, (int)(ncol0 + 1), c->stream));
    EMB_TRY(dev_alloc(c, tmp, tb1));
    EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, tb1, good.p, newcol.p, (int)(ncol0 + 1), c->stream));
    int ncol = 0;
    EMB_CUDA(c, cudaMemcpyAsync(&ncol, newcol.p + ncol0, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (bad_out) {
        DevBuf<unsigned char> bad;
        EMB_TRY(dev_alloc(c, bad, (size_t)ncol0));
        k_auxb_flags<<<blocks_for(ncol0, 256), 256, 0, c->stream>>>(ncol0, good.p, bad.p);
        EMB_LAUNCH_CHECK(c);
        EMB_CUDA(c, cudaMemcpyAsync(bad_out, bad.p, (size_t)ncol0, cudaMemcpyDeviceToHost, c->stream));
        EMB_CUDA(c, cudaStreamSynchronize(c->stream));
        bad.release();
    }
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    *ncol_out = ncol;
    if (ncol == 0) { edges.release(); good.release(); newcol.release(); tmp.release(); return EMB_OK; }
    c->aux.emplace_back();
    AuxSpace& a = c->aux.back();
    a.ncol = ncol; a.nrow = Ns; a.parent = -1; a.solver = 0; a.hid = -1; a.scale_mode = 0;
    EMB_TRY(dev_alloc(c, cnt, (size_t)Ns + 1));
    EMB_TRY(dev_alloc(c, a.rptr, (size_t)Ns + 1));
    EMB_CUDA(c, cudaMemsetAsync(cnt.p + Ns, 0, sizeof(int64_t), c->stream));
    k_auxb_rows<false><<<blocks_for(Ns, 256), 256, 0, c->stream>>>(Ns, which, c->solve_ids.p, nE, nTri, nN, edges.p, c->tris.p,
                                                                 c->tri2f.p, c->nodes.p, tb, good.p, newcol.p, cnt.p, nullptr, nullptr,
                                                                 nullptr, nullptr, nullptr);
    EMB_LAUNCH_CHECK(c);
    size_t tb2 = 0;
    EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tb2, cnt.p, a.rptr.p, (int)(Ns + 1), c->stream));
    if (tb2 > tmp.n) EMB_TRY(dev_alloc(c, tmp, tb2));
    EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, tb2, cnt.p, a.rptr.p, (int)(Ns + 1), c->stream));
    int64_t nnz = 0;
    EMB_CUDA(c, cudaMemcpyAsync(&nnz, a.rptr.p + Ns, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    a.nnz = nnz;
    if (nnz >= ((int64_t)1 << 31)) { c->aux.pop_back(); c->err = "emb_aux_build_top: more than 2^31 entries"; return EMB_ERR_LIMIT; }
    DevBuf<unsigned long long> tkey, skey;
    DevBuf<int> tcnt;
    EMB_TRY(dev_alloc(c, a.rcol, (size_t)nnz));
    EMB_TRY(dev_alloc(c, a.rval, (size_t)nnz));
    EMB_TRY(dev_alloc(c, a.tcol, (size_t)nnz));
    EMB_TRY(dev_alloc(c, a.tval, (size_t)nnz));
    EMB_TRY(dev_alloc(c, a.tptr, (size_t)ncol + 1));
    EMB_TRY(dev_alloc(c, tkey, (size_t)nnz));
    EMB_TRY(dev_alloc(c, skey, (size_t)nnz));
    EMB_TRY(dev_alloc(c, colcnt, (size_t)ncol + 1));
    EMB_TRY(dev_alloc(c, tcnt, (size_t)ncol + 1));
    EMB_CUDA(c, cudaMemsetAsync(colcnt.p, 0, ((size_t)ncol + 1) * sizeof(int), c->stream));
    k_auxb_rows<true><<<blocks_for(Ns, 256), 256, 0, c->stream>>>(Ns, which, c->solve_ids.p, nE, nTri, nN, edges.p, c->tris.p,
                                                                c->tri2f.p, c->nodes.p, tb, good.p, newcol.p, nullptr, a.rptr.p,
                                                                a.rcol.p, a.rval.p, tkey.p, colcnt.p);
    EMB_LAUNCH_CHECK(c);
    // R^T: entries ordered by (column, row) - one stable radix sort of 64-bit keys carrying the values
    int cbits = 1;
    while (((int64_t)1 << cbits) < ncol) ++cbits;
    size_t tb3 = 0;
    EMB_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tb3, tkey.p, skey.p, a.rval.p, a.tval.p, (int)nnz, 0, 32 + cbits, c->stream));
    if (tb3 > tmp.n) EMB_TRY(dev_alloc(c, tmp, tb3));
    EMB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tb3, tkey.p, skey.p, a.rval.p, a.tval.p, (int)nnz, 0, 32 + cbits, c->stream));
    k_auxb_tcol<<<blocks_for(nnz, 256), 256, 0, c->stream>>>(nnz, skey.p, a.tcol.p);
    EMB_LAUNCH_CHECK(c);
    size_t tb4 = 0;
    EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tb4, colcnt.p, tcnt.p, ncol + 1, c->stream));
    if (tb4 > tmp.n) EMB_TRY(dev_alloc(c, tmp, tb4));
    EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, tb4, colcnt.p, tcnt.p, ncol + 1, c->stream));
    k_auxb_i2l<<<blocks_for(ncol + 1, 256), 256, 0, c->stream>>>(ncol + 1, tcnt.p, a.tptr.p);
    EMB_LAUNCH_CHECK(c);
    c->launches += 12;
    EMB_TRY(dev_alloc(c, a.dinv, (size_t)ncol));
    EMB_TRY(dev_alloc(c, a.tmp, (size_t)ncol * AUXB_NVMAX));
    EMB_TRY(dev_alloc(c, a.traw, (size_t)ncol * AUXB_NVMAX));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    edges.release(); cnt.release(); good.release(); newcol.release(); colcnt.release(); tcnt.release(); tkey.release();
    skey.release(); tmp.release();
    c->have_As = false;
    return EMB_OK;
}

// host copy of an auxiliary space's transfer matrix (CSR of R, rows in solve-index order): test / inspection access
extern "C" int emb_aux_get(emb_ctx* c, int idx, int64_t* nrow, int64_t* ncol, int64_t* nnz, int64_t* rptr, int32_t* rcol, double* rval) {
    if (!c || idx < 0 || idx >= (int)c->aux.size() || !nrow || !ncol || !nnz) return EMB_ERR_ARG;
    AuxSpace& a = c->aux[idx];
    *nrow = a.nrow; *ncol = a.ncol; *nnz = a.nnz;
    if (rptr) EMB_CUDA(c, cudaMemcpyAsync(rptr, a.rptr.p, ((size_t)a.nrow + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    if (rcol) EMB_CUDA(c, cudaMemcpyAsync(rcol, a.rcol.p, (size_t)a.nnz * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (rval) EMB_CUDA(c, cudaMemcpyAsync(rval, a.rval.p, (size_t)a.nnz * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    return EMB_OK;
}
