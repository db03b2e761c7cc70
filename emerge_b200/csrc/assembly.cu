// Mesh upload, symbolic phase (pattern), numeric phase (element kernel + deterministic reduction).
//
// Reference path replaced: fem/physics/edm/optimized_assembly.py:43-116 (_matrix_builder + coo->csr)
// and the element kernel fem/mth/tet.py:854-1088.
//
// Data layout in HBM
//   nodes  [nN][3] f64           tetc [nT][4] i32 (vertices ascending)     gid [nT][20] i32 (canonical dofs)
//   er/ur  [9][nT] c128 (component-major: a warp of tets reads 32 consecutive values per component)
//   cooK/cooM [nT_chunk][20][20] c128 scratch: row (t,i) is 320 contiguous bytes = 10 full sectors
//   adj    dof -> ascending list of (tet*20+i)   rowptr/col canonical CSR   K,M [nnz] c128
// Kernels (default numeric phase, "fused": K and M values are written exactly once, no COO intermediate)
//   k_tet_records: one thread per tetrahedron -> 896-B record (ned2_fused.cuh): everything an element-matrix row needs.
//   k_asm_rows   : one warp per ENTITY (edge / face = rows e and e + nE + nTri, which share adjacency and column list);
//                  walks the entity's tetrahedra in ascending order, lanes 0..19 = the 20 columns of the element-matrix
//                  rows, evaluates the 2 x 20 K and M entries from the record (staged in shared memory, next record
//                  prefetched into registers), accumulates into the row held in shared memory (fixed order => bitwise
//                  reproducible), then streams the four finished rows to HBM.  Persistent grid, entities ordered by
//                  their first tetrahedron so that records are re-read from L2.
// Kernels (alternative numeric phase, emb_assemble_mode(1): element kernel -> COO scratch -> row reduction)
//   tet_kernel   : one thread per tetrahedron, FP64, all indices compile-time (ned2_tet.cuh); algorithmic
//                  traffic 16+96+288 B in, 12.8 KB out per tet.  Store-bound.
//   reduce_rows  : one warp per CSR row; walks the row's adjacency in ascending tet order (fixed summation
//                  order => bitwise reproducible), lanes 0..19 fetch one 320-B COO row + the tet's 20 dof ids,
//                  locate the column by binary search in the row's sorted column list held in shared memory.
#include "context.cuh"
#include "ned2_tet.cuh"
#include "ned2_fused.cuh"
#include <cub/cub.cuh>

// ------------------------------------------------------------------------------------------------
// upload
// ------------------------------------------------------------------------------------------------
__global__ void k_cast_i64_i32(const int64_t* __restrict__ in, int* __restrict__ out, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)in[i];
}

// per tet: sort vertex ids, canonical dof list
__global__ void k_canon(int64_t nT, const int64_t* __restrict__ tets, const int64_t* __restrict__ ttf,
                        int* __restrict__ tetc, int* __restrict__ tetord, int* __restrict__ gid) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nT) return;
    int64_t v[4];
    int ord[4] = {0, 1, 2, 3};
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = tets[t * 4 + k];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i + 1; j < 4; ++j)
            if (v[j] < v[i]) {
                int64_t tv = v[i]; v[i] = v[j]; v[j] = tv;
                int to = ord[i]; ord[i] = ord[j]; ord[j] = to;
            }
    int4 tc = make_int4((int)v[0], (int)v[1], (int)v[2], (int)v[3]);
    reinterpret_cast<int4*>(tetc)[t] = tc;
    tetord[t] = ord[0] | (ord[1] << 2) | (ord[2] << 4) | (ord[3] << 6);
    int ref[20];
    ned2::canonical_to_ref(ord, ref);
#pragma unroll
    for (int c = 0; c < 20; ++c) gid[t * 20 + c] = (int)ttf[(int64_t)ref[c] * nT + t];
}

extern "C" int emb_upload_mesh(emb_ctx* c, int64_t nN, int64_t nT, int64_t nE, int64_t nTri, const double* nodes,
                               const int64_t* tets, const int64_t* tris, const int64_t* ttf, const int64_t* tri2f) {
    if (!c || !nodes || !tets || !tris || !ttf || !tri2f || nT <= 0) {
        if (c) c->err = "emb_upload_mesh: null/empty argument";
        return EMB_ERR_ARG;
    }
    if (2 * nE + 2 * nTri >= (int64_t)1 << 31 || 20 * nT >= (int64_t)1 << 31) {
        c->err = "emb_upload_mesh: mesh exceeds int32 dof/adjacency range";
        return EMB_ERR_LIMIT;
    }
    c->nN = nN; c->nT = nT; c->nE = nE; c->nTri = nTri; c->N = 2 * nE + 2 * nTri;
    c->have_pattern = c->have_KM = c->have_dirichlet = c->have_A = false;
    EMB_TRY(h2d(c, c->nodes, nodes, (size_t)nN * 3));
    DevBuf<int64_t> t64, f64, r64, g64;
    EMB_TRY(h2d(c, t64, tets, (size_t)nT * 4));
    EMB_TRY(h2d(c, f64, ttf, (size_t)nT * 20));
    EMB_TRY(h2d(c, r64, tris, (size_t)nTri * 3));
    EMB_TRY(h2d(c, g64, tri2f, (size_t)nTri * 8));
    EMB_TRY(dev_alloc(c, c->tetc, (size_t)nT * 4));
    EMB_TRY(dev_alloc(c, c->tetord, (size_t)nT));
    EMB_TRY(dev_alloc(c, c->gid, (size_t)nT * 20));
    EMB_TRY(dev_alloc(c, c->tris, (size_t)nTri * 3));
    EMB_TRY(dev_alloc(c, c->tri2f, (size_t)nTri * 8));
    k_canon<<<blocks_for(nT, 128), 128, 0, c->stream>>>(nT, t64.p, f64.p, c->tetc.p, c->tetord.p, c->gid.p);
    EMB_LAUNCH_CHECK(c);
    k_cast_i64_i32<<<blocks_for(nTri * 3, 256), 256, 0, c->stream>>>(r64.p, c->tris.p, nTri * 3);
    EMB_LAUNCH_CHECK(c);
    k_cast_i64_i32<<<blocks_for(nTri * 8, 256), 256, 0, c->stream>>>(g64.p, c->tri2f.p, nTri * 8);
    EMB_LAUNCH_CHECK(c);
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    t64.release(); f64.release(); r64.release(); g64.release();
    c->have_mesh = true;
    return EMB_OK;
}

// (3,3,nT) C-order is already component-major [9][nT]
extern "C" int emb_upload_materials(emb_ctx* c, const emb_c128* er, const emb_c128* ur) {
    if (!c || !c->have_mesh || !er || !ur) {
        if (c) c->err = "emb_upload_materials: mesh not uploaded or null argument";
        return c ? EMB_ERR_STATE : EMB_ERR_ARG;
    }
    EMB_TRY(h2d(c, c->er, reinterpret_cast<const cx*>(er), (size_t)9 * c->nT));
    EMB_TRY(h2d(c, c->ur, reinterpret_cast<const cx*>(ur), (size_t)9 * c->nT));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->have_mat = true;
    c->have_KM = c->have_A = false;
    return EMB_OK;
}

// ------------------------------------------------------------------------------------------------
// symbolic phase
// ------------------------------------------------------------------------------------------------
__global__ void k_iota(int* v, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int)i;
}

// keys sorted ascending; ptr[d] = first position with key >= d
__global__ void k_segptr(const int* __restrict__ keys, int64_t n, int64_t N, int64_t* __restrict__ ptr) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int k = keys[i];
    if (i == 0) {
        for (int d = 0; d <= k; ++d) ptr[d] = 0;
    } else {
        int kp = keys[i - 1];
        for (int d = kp + 1; d <= k; ++d) ptr[d] = i;
    }
    if (i == n - 1)
        for (int64_t d = (int64_t)k + 1; d <= N; ++d) ptr[d] = n;
}

__global__ void k_key_class(int n, const unsigned* __restrict__ key, int* __restrict__ cls) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cls[i] = (int)(key[i] >> 28);
}

static int build_entity_order(emb_ctx* c);   // work list of the fused numeric phase (below)

constexpr int CANDCAP = 2048;   // candidates per row held in shared memory: up to 102 tets around a dof
constexpr int PWARPS = 4;

template <bool FILL>
__global__ void __launch_bounds__(PWARPS * 32) k_row_pattern(int64_t N, const int64_t* __restrict__ adjptr,
                                                             const int* __restrict__ adj, const int* __restrict__ gid,
                                                             int64_t* __restrict__ rowlen, const int64_t* __restrict__ rowptr,
                                                             int* __restrict__ col, int* __restrict__ errflag) {
    __shared__ int s_cand[PWARPS][CANDCAP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = blockIdx.x * (int64_t)PWARPS + warp;
    if (r >= N) return;
    int* cand = s_cand[warp];
    const int64_t a0 = adjptr[r];
    const int deg = (int)(adjptr[r + 1] - a0);
    const int ncand = deg * 20;
    if (ncand > CANDCAP) {
        if (lane == 0) atomicExch(errflag, 1);
        if (!FILL && lane == 0) rowlen[r] = 0;
        return;
    }
    int n2 = 32;
    while (n2 < ncand) n2 <<= 1;
    for (int k = lane; k < n2; k += 32) {
        int v = 0x7fffffff;
        if (k < ncand) {
            int a = adj[a0 + k / 20];
            v = gid[(int64_t)(a / 20) * 20 + (k % 20)];
        }
        cand[k] = v;
    }
    __syncwarp();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < n2; i += 32) {
                int ixj = i ^ j;
                if (ixj > i) {
                    int a = cand[i], b = cand[ixj];
                    bool asc = (i & k) == 0;
                    if ((a > b) == asc) { cand[i] = b; cand[ixj] = a; }
                }
            }
            __syncwarp();
        }
    int base = 0;
    const int64_t out0 = FILL ? rowptr[r] : 0;
    for (int k0 = 0; k0 < ncand; k0 += 32) {
        int k = k0 + lane;
        bool flag = k < ncand && (k == 0 || cand[k] != cand[k - 1]);
        unsigned m = __ballot_sync(0xffffffffu, flag);
        if (FILL && flag) col[out0 + base + __popc(m & ((1u << lane) - 1))] = cand[k];
        base += __popc(m);
    }
    if (!FILL && lane == 0) rowlen[r] = base;
}

extern "C" int emb_symbolic(emb_ctx* c) {
    if (!c || !c->have_mesh) {
        if (c) c->err = "emb_symbolic: mesh not uploaded";
        return c ? EMB_ERR_STATE : EMB_ERR_ARG;
    }
    PhaseTimer pt(c, "symbolic");
    const int64_t n = 20 * c->nT, N = c->N;
    DevBuf<int> vals_in, keys_out;
    DevBuf<char> tmp;
    EMB_TRY(dev_alloc(c, vals_in, (size_t)n));
    EMB_TRY(dev_alloc(c, keys_out, (size_t)n));
    EMB_TRY(dev_alloc(c, c->adj, (size_t)n));
    EMB_TRY(dev_alloc(c, c->adjptr, (size_t)N + 1));
    k_iota<<<blocks_for(n, 256), 256, 0, c->stream>>>(vals_in.p, n);
    EMB_LAUNCH_CHECK(c);
    int bits = 1;
    while (((int64_t)1 << bits) < N) ++bits;
    size_t tb = 0;
    EMB_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tb, c->gid.p, keys_out.p, vals_in.p, c->adj.p, (int)n, 0, bits, c->stream));
    EMB_TRY(dev_alloc(c, tmp, tb));
    EMB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tb, c->gid.p, keys_out.p, vals_in.p, c->adj.p, (int)n, 0, bits, c->stream));
    c->launches += 8;
    k_segptr<<<blocks_for(n, 256), 256, 0, c->stream>>>(keys_out.p, n, N, c->adjptr.p);
    EMB_LAUNCH_CHECK(c);
    vals_in.release();
    keys_out.release();

    DevBuf<int64_t> rowlen;
    DevBuf<int> errflag;
    EMB_TRY(dev_alloc(c, rowlen, (size_t)N + 1));
    EMB_TRY(dev_alloc(c, errflag, 1));
    EMB_TRY(dev_alloc(c, c->rowptr, (size_t)N + 1));
    EMB_CUDA(c, cudaMemsetAsync(errflag.p, 0, sizeof(int), c->stream));
    EMB_CUDA(c, cudaMemsetAsync(rowlen.p + N, 0, sizeof(int64_t), c->stream));
    k_row_pattern<false><<<blocks_for(N, PWARPS), PWARPS * 32, 0, c->stream>>>(N, c->adjptr.p, c->adj.p, c->gid.p, rowlen.p,
                                                                              nullptr, nullptr, errflag.p);
    EMB_LAUNCH_CHECK(c);
    size_t tb2 = 0;
    EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tb2, rowlen.p, c->rowptr.p, (int)(N + 1), c->stream));
    if (tb2 > tmp.n) EMB_TRY(dev_alloc(c, tmp, tb2));
    EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, tb2, rowlen.p, c->rowptr.p, (int)(N + 1), c->stream));
    c->launches += 2;
    int herr = 0;
    int64_t nnz = 0;
    EMB_CUDA(c, cudaMemcpyAsync(&herr, errflag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaMemcpyAsync(&nnz, c->rowptr.p + N, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (herr) {
        c->err = "emb_symbolic: a dof is shared by more than " + std::to_string(CANDCAP / 20) + " tetrahedra";
        return EMB_ERR_LIMIT;
    }
    c->nnz = nnz;
    EMB_TRY(dev_alloc(c, c->col, (size_t)nnz));
    k_row_pattern<true><<<blocks_for(N, PWARPS), PWARPS * 32, 0, c->stream>>>(N, c->adjptr.p, c->adj.p, c->gid.p, nullptr,
                                                                             c->rowptr.p, c->col.p, errflag.p);
    EMB_LAUNCH_CHECK(c);
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    rowlen.release(); errflag.release(); tmp.release();
    c->have_pattern = true;
    c->have_KM = c->have_dirichlet = c->have_A = false;
    c->asm_items.release();
    c->asm_chunk = 0;
    {
        PhaseTimer pt2(c, "symbolic_entities");
        EMB_TRY(build_entity_order(c));
    }
    return EMB_OK;
}

// ------------------------------------------------------------------------------------------------
// numeric phase: element kernel
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st16(cx* p, cx v) {
    *reinterpret_cast<double2*>(p) = make_double2(v.re, v.im);
}
// two consecutive entries as ONE 32-byte store (full sector; STG.E.ENL2.256 on sm_100a)
__device__ __forceinline__ void st32(cx* p, cx a, cx b) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a.re), "d"(a.im), "d"(b.re), "d"(b.im) : "memory");
}

// Shared-memory structure-of-arrays copy of the per-element quantities of the block's 32 tetrahedra.
// Lane = tetrahedron, so every access below is conflict-free; indices are compile-time after inlining.
struct TetSmem {
    double X[6][3][32];
    double2 Y[6][3][32];
    double2 g[4][4][32];
    double len[6][32];
    double kK[32], kM[32];
};
struct TetView {
    const TetSmem* s;
    int lane;
    __device__ __forceinline__ double X(int a, int k) const { return s->X[a][k][lane]; }
    __device__ __forceinline__ cx Y(int a, int k) const { double2 v = s->Y[a][k][lane]; return cx{v.x, v.y}; }
    __device__ __forceinline__ cx g(int p, int q) const { double2 v = s->g[p][q][lane]; return cx{v.x, v.y}; }
    __device__ __forceinline__ double len(int a) const { return s->len[a][lane]; }
    __device__ __forceinline__ double kK() const { return s->kK[lane]; }
    __device__ __forceinline__ double kM() const { return s->kM[lane]; }
};

template <int I, int... J2s>
__device__ __forceinline__ void emit_row(const TetView& d, cx* __restrict__ K, cx* __restrict__ M,
                                         std::integer_sequence<int, J2s...>) {
    ((st32(K + I * 20 + 2 * J2s, ned2::kentry<I, 2 * J2s>(d), ned2::kentry<I, 2 * J2s + 1>(d)),
      st32(M + I * 20 + 2 * J2s, ned2::mentry<I, 2 * J2s>(d), ned2::mentry<I, 2 * J2s + 1>(d))), ...);
}

__device__ __forceinline__ void load_tensor(const cx* __restrict__ base, int64_t nT, int64_t t, cx m[3][3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double2 a = __ldg(reinterpret_cast<const double2*>(base + (int64_t)(i * 3 + j) * nT + t));
            m[i][j] = cx{a.x, a.y};
        }
}

// piece W of the per-element setup: W<6 -> vertex pair W (X, Y=Ms.X, length); W=6..9 -> column W-6 of g
template <int W>
__device__ __forceinline__ void setup_piece(TetSmem& sm, int lane, int64_t t, int64_t nT, const double (&p)[4][3],
                                            const double (&G)[4][3], double V6, const cx* __restrict__ er,
                                            const cx* __restrict__ ur) {
    if constexpr (W < 6) {
        constexpr int a = ned2::eA(W), b = ned2::eB(W);
        cx mu[3][3], Ms[3][3];
        load_tensor(ur, nT, t, mu);
        ned2::matinv_ref(mu, Ms);
        double X[3];
        X[0] = G[a][1] * G[b][2] - G[a][2] * G[b][1];
        X[1] = G[a][2] * G[b][0] - G[a][0] * G[b][2];
        X[2] = G[a][0] * G[b][1] - G[a][1] * G[b][0];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            sm.X[W][r][lane] = X[r];
            cx y = X[0] * Ms[r][0] + X[1] * Ms[r][1] + X[2] * Ms[r][2];
            sm.Y[W][r][lane] = make_double2(y.re, y.im);
        }
        const double dx = p[a][0] - p[b][0], dy = p[a][1] - p[b][1], dz = p[a][2] - p[b][2];
        sm.len[W][lane] = sqrt(dx * dx + dy * dy + dz * dz);
        if constexpr (W == 0) {
            const double kM = 1.0 / V6;
            sm.kM[lane] = kM;
            sm.kK[lane] = kM * kM * kM;
        }
    } else {
        constexpr int q = W - 6;
        cx Mm[3][3], H[3];
        load_tensor(er, nT, t, Mm);
#pragma unroll
        for (int r = 0; r < 3; ++r) H[r] = G[q][0] * Mm[r][0] + G[q][1] * Mm[r][1] + G[q][2] * Mm[r][2];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
            cx v = G[pp][0] * H[0] + G[pp][1] * H[1] + G[pp][2] * H[2];
            sm.g[pp][q][lane] = make_double2(v.re, v.im);
        }
    }
}

// Element kernel.  Block = 32 tetrahedra (lane) x 10 warps; warp w owns the two functions of one entity
// (rows w and w+10: edge w, or face w-6) and emits their 2x20 K and M entries as 32-byte stores.
// COO chunk layout: [t - t0][20][20] in canonical order.
constexpr int TWARPS = 10;
__global__ void __launch_bounds__(TWARPS * 32) k_tet(int64_t t0, int64_t t1, int64_t nT, const int* __restrict__ tetc,
                                                     const double* __restrict__ nodes, const cx* __restrict__ er,
                                                     const cx* __restrict__ ur, cx* __restrict__ cooK,
                                                     cx* __restrict__ cooM) {
    __shared__ TetSmem sm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t t = t0 + blockIdx.x * (int64_t)32 + lane;
    const bool valid = t < t1;
    if (valid) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(tetc) + t);
        const int vi[4] = {v.x, v.y, v.z, v.w};
        double p[4][3], G[4][3], e1[3], e2[3], e3[3];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double* q = nodes + (int64_t)vi[k] * 3;
            p[k][0] = __ldg(q); p[k][1] = __ldg(q + 1); p[k][2] = __ldg(q + 2);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            e1[k] = p[1][k] - p[0][k];
            e2[k] = p[2][k] - p[0][k];
            e3[k] = p[3][k] - p[0][k];
        }
        G[1][0] = e2[1] * e3[2] - e2[2] * e3[1]; G[1][1] = e2[2] * e3[0] - e2[0] * e3[2]; G[1][2] = e2[0] * e3[1] - e2[1] * e3[0];
        G[2][0] = e3[1] * e1[2] - e3[2] * e1[1]; G[2][1] = e3[2] * e1[0] - e3[0] * e1[2]; G[2][2] = e3[0] * e1[1] - e3[1] * e1[0];
        G[3][0] = e1[1] * e2[2] - e1[2] * e2[1]; G[3][1] = e1[2] * e2[0] - e1[0] * e2[2]; G[3][2] = e1[0] * e2[1] - e1[1] * e2[0];
#pragma unroll
        for (int k = 0; k < 3; ++k) G[0][k] = -(G[1][k] + G[2][k] + G[3][k]);
        const double det = e1[0] * G[1][0] + e1[1] * G[1][1] + e1[2] * G[1][2];
        const double V6 = fabs(det);
        switch (warp) {
            case 0: setup_piece<0>(sm, lane, t, nT, p, G, V6, er, ur); break;
            case 1: setup_piece<1>(sm, lane, t, nT, p, G, V6, er, ur); break;
            case 2: setup_piece<2>(sm, lane, t, nT, p, G, V6, er, ur); break;
            case 3: setup_piece<3>(sm, lane, t, nT, p, G, V6, er, ur); break;
            case 4: setup_piece<4>(sm, lane, t, nT, p, G, V6, er, ur); break;
            case 5: setup_piece<5>(sm, lane, t, nT, p, G, V6, er, ur); break;
            case 6: setup_piece<6>(sm, lane, t, nT, p, G, V6, er, ur); break;
            case 7: setup_piece<7>(sm, lane, t, nT, p, G, V6, er, ur); break;
            case 8: setup_piece<8>(sm, lane, t, nT, p, G, V6, er, ur); break;
            default: setup_piece<9>(sm, lane, t, nT, p, G, V6, er, ur); break;
        }
    }
    __syncthreads();
    if (!valid) return;
    const TetView d{&sm, lane};
    cx* K = cooK + (t - t0) * 400;
    cx* M = cooM + (t - t0) * 400;
    constexpr auto seq = std::make_integer_sequence<int, 10>{};
    switch (warp) {
        case 0: emit_row<0>(d, K, M, seq); emit_row<10>(d, K, M, seq); break;
        case 1: emit_row<1>(d, K, M, seq); emit_row<11>(d, K, M, seq); break;
        case 2: emit_row<2>(d, K, M, seq); emit_row<12>(d, K, M, seq); break;
        case 3: emit_row<3>(d, K, M, seq); emit_row<13>(d, K, M, seq); break;
        case 4: emit_row<4>(d, K, M, seq); emit_row<14>(d, K, M, seq); break;
        case 5: emit_row<5>(d, K, M, seq); emit_row<15>(d, K, M, seq); break;
        case 6: emit_row<6>(d, K, M, seq); emit_row<16>(d, K, M, seq); break;
        case 7: emit_row<7>(d, K, M, seq); emit_row<17>(d, K, M, seq); break;
        case 8: emit_row<8>(d, K, M, seq); emit_row<18>(d, K, M, seq); break;
        default: emit_row<9>(d, K, M, seq); emit_row<19>(d, K, M, seq); break;
    }
}

// ------------------------------------------------------------------------------------------------
// numeric phase: deterministic reduction COO -> CSR
// ------------------------------------------------------------------------------------------------
constexpr int RWARPS = 4;
constexpr int ROWCAP = 192;    // row entries cached in shared memory per warp; longer rows use global memory

__global__ void __launch_bounds__(RWARPS * 32) k_reduce_rows(int64_t N, int64_t t0, int64_t t1, int first_launch,
                                                             const unsigned long long* __restrict__ items,
                                                             const int64_t* __restrict__ rowptr, const int* __restrict__ col,
                                                             const int64_t* __restrict__ adjptr, const int* __restrict__ adj,
                                                             const int* __restrict__ gid, const cx* __restrict__ cooK,
                                                             const cx* __restrict__ cooM, cx* __restrict__ K, cx* __restrict__ M) {
    __shared__ int s_col[RWARPS][ROWCAP];
    __shared__ double2 s_K[RWARPS][ROWCAP];
    __shared__ double2 s_M[RWARPS][ROWCAP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work unit: a whole row (items == nullptr, N rows), or the part of a row that lies in the tet chunk [t0, t1)
    // (items: N packed (row << 32 | first adjacency index of the chunk) entries, see build_chunk_items)
    const int64_t w = blockIdx.x * (int64_t)RWARPS + warp;
    if (w >= N) return;
    int64_t r = w, ai0 = -1;
    if (items) {
        const unsigned long long it = items[w];
        r = (int64_t)(it >> 32);
        ai0 = (int64_t)(it & 0xffffffffull);
    }
    const int64_t a0 = adjptr[r], a1 = adjptr[r + 1];
    const bool first = items ? (ai0 == a0) : (first_launch != 0);
    if (!items) ai0 = a0;
    const int64_t p0 = rowptr[r];
    const int len = (int)(rowptr[r + 1] - p0);
    const bool in_smem = len <= ROWCAP;
    const int* cols = col + p0;
    if (in_smem) {
        for (int k = lane; k < len; k += 32) {
            s_col[warp][k] = cols[k];
            double2 zk = make_double2(0, 0), zm = zk;
            if (!first) {
                zk = *reinterpret_cast<const double2*>(K + p0 + k);
                zm = *reinterpret_cast<const double2*>(M + p0 + k);
            }
            s_K[warp][k] = zk;
            s_M[warp][k] = zm;
        }
        cols = s_col[warp];
    } else if (first) {
        for (int k = lane; k < len; k += 32) {
            st16(K + p0 + k, cx{0, 0});
            st16(M + p0 + k, cx{0, 0});
        }
    }
    __syncwarp();
    for (int64_t ai = ai0; ai < a1; ++ai) {
        const int a = adj[ai];            // tet*20 + canonical local row, ascending in tet
        const int t = a / 20;
        if (t < t0) continue;
        if (t >= t1) break;
        if (lane < 20) {
            const int cj = gid[(int64_t)t * 20 + lane];
            const int64_t src = ((int64_t)(a - t0 * 20)) * 20 + lane;
            const double2 vk = *reinterpret_cast<const double2*>(cooK + src);
            const double2 vm = *reinterpret_cast<const double2*>(cooM + src);
            int lo = 0, hi = len - 1;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (cols[mid] < cj) lo = mid + 1; else hi = mid;
            }
            if (in_smem) {
                s_K[warp][lo].x += vk.x; s_K[warp][lo].y += vk.y;
                s_M[warp][lo].x += vm.x; s_M[warp][lo].y += vm.y;
            } else {
                double2* pk = reinterpret_cast<double2*>(K + p0 + lo);
                double2* pm = reinterpret_cast<double2*>(M + p0 + lo);
                double2 ck = *pk, cm = *pm;
                ck.x += vk.x; ck.y += vk.y; cm.x += vm.x; cm.y += vm.y;
                *pk = ck; *pm = cm;
            }
        }
        __syncwarp();
    }
    if (in_smem)
        for (int k = lane; k < len; k += 32) {
            *reinterpret_cast<double2*>(K + p0 + k) = s_K[warp][k];
            *reinterpret_cast<double2*>(M + p0 + k) = s_M[warp][k];
        }
}


// ------------------------------------------------------------------------------------------------
// numeric phase, fused: per-tet records + one warp per entity
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_tet_records(int64_t nT, const int* __restrict__ tetc, const double* __restrict__ nodes,
                                                     const cx* __restrict__ er, const cx* __restrict__ ur,
                                                     ned2f::TetRec* __restrict__ recs) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nT) return;
    const int4 v = __ldg(reinterpret_cast<const int4*>(tetc) + t);
    const int vi[4] = {v.x, v.y, v.z, v.w};
    double p[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double* q = nodes + (int64_t)vi[k] * 3;
        p[k][0] = __ldg(q); p[k][1] = __ldg(q + 1); p[k][2] = __ldg(q + 2);
    }
    cx mu[3][3], Ms[3][3], Mm[3][3];
    load_tensor(ur, nT, t, mu);
    ned2::matinv_ref(mu, Ms);
    load_tensor(er, nT, t, Mm);
    ned2f::TetRec r;
    ned2f::make_record(p, Ms, Mm, r);
    cx* out = reinterpret_cast<cx*>(recs + t);
    const cx* in = reinterpret_cast<const cx*>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(ned2f::TetRec) / 32); ++k) st32(out + 2 * k, in[2 * k], in[2 * k + 1]);
}

// the mesh tables have the structure the fused kernel relies on: gid[t][c+10] = gid[t][c] + nE + nTri
__global__ void k_check_pairs(int64_t nT, int nEF, const int* __restrict__ gid, int* __restrict__ bad) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nT * 10) return;
    const int64_t t = i / 10;
    const int c = (int)(i % 10);
    const int a = gid[t * 20 + c], b = gid[t * 20 + c + 10];
    if (a >= nEF || b != a + nEF) atomicExch(bad, 1);
}

// processing class of an entity and its sort key (class << 28 | first tet)
// row-length caps of the two shared-memory classes (faces and edges with <= 6 tetrahedra | edges with <= 13); class 2 =
// accumulators in global memory.  One launch per class; inside a class entities are ordered by their first tetrahedron.
constexpr int ACLS0 = 80, ACLS1 = 160;
constexpr int AWARPS0 = 20, AWARPS1 = 12;
__global__ void k_entity_keys(int nEF, const int64_t* __restrict__ adjptr, const int* __restrict__ adj,
                              const int64_t* __restrict__ rowptr, unsigned* __restrict__ key, int* __restrict__ val,
                              int* __restrict__ bad) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nEF) return;
    const int64_t a0 = adjptr[e];
    const int deg = (int)(adjptr[e + 1] - a0);
    const int len = (int)(rowptr[e + 1] - rowptr[e]);
    const int len2 = (int)(rowptr[e + nEF + 1] - rowptr[e + nEF]);
    const int deg2 = (int)(adjptr[e + nEF + 1] - adjptr[e + nEF]);
    if (len2 != len || deg2 != deg) atomicExch(bad, 1);
    int cls = len <= ACLS0 ? 0 : len <= ACLS1 ? 1 : 2;
    if (deg > 32) cls = 2;
    const unsigned t0 = deg > 0 ? (unsigned)(adj[a0] / 20) : 0u;
    key[e] = ((unsigned)cls << 28) | t0;
    val[e] = e;
}

template <int ROWCAP>
struct AsmWarpBuf {
    double2 acc[4][ROWCAP];      // K row a, K row b, M row a, M row b
    double2 rec[56];             // the current tetrahedron's record
    int cols[ROWCAP];
};

__device__ const ned2f::KernTables d_ktab{};

// per-item metadata, identical in every lane (broadcast loads), prefetched one item ahead
struct AsmMeta {
    int64_t a0, p0, p1;
    int deg, len;
};
__device__ __forceinline__ AsmMeta asm_load_meta(int e, int nEF, const int64_t* __restrict__ adjptr,
                                                 const int64_t* __restrict__ rowptr) {
    AsmMeta m;
    m.a0 = __ldg(adjptr + e);
    m.deg = (int)(__ldg(adjptr + e + 1) - m.a0);
    m.p0 = __ldg(rowptr + e);
    m.len = (int)(__ldg(rowptr + e + 1) - m.p0);
    m.p1 = __ldg(rowptr + e + nEF);
    return m;
}

template <int ROWCAP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_asm_rows(int64_t nitems, const int* __restrict__ items, int nEF,
                                                         const int64_t* __restrict__ adjptr, const int* __restrict__ adj,
                                                         const int* __restrict__ gid, const int64_t* __restrict__ rowptr,
                                                         const int* __restrict__ col,
                                                         const ned2f::TetRec* __restrict__ recs, cx* __restrict__ K,
                                                         cx* __restrict__ M) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ned2f::KernTables* T = reinterpret_cast<ned2f::KernTables*>(smem_raw);
    {
        const double* src = reinterpret_cast<const double*>(&d_ktab);
        double* dst = reinterpret_cast<double*>(T);
        for (int i = threadIdx.x; i < (int)(sizeof(ned2f::KernTables) / sizeof(double)); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    constexpr size_t TOFF = (sizeof(ned2f::KernTables) + 15) / 16 * 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    AsmWarpBuf<ROWCAP>& wb = reinterpret_cast<AsmWarpBuf<ROWCAP>*>(smem_raw + TOFF)[warp];
    const bool colLane = lane < 20;
    const int J = colLane ? lane : 0;
    const int lpj = T->lp[J];
    const double2* recD = wb.rec;                 // D[36] | g[16] | len[6] (+pad) as 16-byte words
    const double2* recG = wb.rec + 36;
    const double* recLen = reinterpret_cast<const double*>(wb.rec + 52);

    // Software pipeline over the items of this warp (every stage is a dependent global load of the previous one):
    //   item w+3: entity id          item w+2: metadata (row / adjacency offsets)
    //   item w+1: adjacency list and column list (registers)          item w: records, tet by tet, one ahead
    // so that no item starts with an exposed chain of DRAM round trips.
    constexpr int CREG = (ROWCAP + 31) / 32;
    const int64_t stride = (int64_t)gridDim.x * WARPS;
    int64_t w = blockIdx.x * (int64_t)WARPS + warp;
    if (w >= nitems) return;
    auto item_at = [&](int64_t i) { return i < nitems ? __ldg(items + i) : -1; };
    auto load_adj = [&](const AsmMeta& mm) { return lane < mm.deg ? __ldg(adj + mm.a0 + lane) : 0; };
    int e1 = item_at(w + stride), e2 = item_at(w + 2 * stride);
    AsmMeta m = asm_load_meta(__ldg(items + w), nEF, adjptr, rowptr);
    AsmMeta m1 = m;
    if (e1 >= 0) m1 = asm_load_meta(e1, nEF, adjptr, rowptr);
    int myadj = load_adj(m);
    int colr[CREG];
#pragma unroll
    for (int i = 0; i < CREG; ++i) colr[i] = (lane + 32 * i < m.len) ? __ldg(col + m.p0 + lane + 32 * i) : 0;
    double2 r0 = make_double2(0, 0), r1 = r0;
    int cj = 0;
    auto fetch_rec = [&](int t) {
        const double2* rp = reinterpret_cast<const double2*>(recs + t);
        r0 = __ldg(rp + lane);
        if (lane < 24) r1 = __ldg(rp + 32 + lane);
        if (colLane) cj = __ldg(gid + (int64_t)t * 20 + lane);
    };
    if (m.deg > 0) fetch_rec(__shfl_sync(0xffffffffu, myadj, 0) / 20);
    for (; w < nitems; w += stride) {
        // stage loads of the following items
        int adj1 = 0, colr1[CREG];
#pragma unroll
        for (int i = 0; i < CREG; ++i) colr1[i] = 0;
        if (e1 >= 0) {
            adj1 = load_adj(m1);
#pragma unroll
            for (int i = 0; i < CREG; ++i) colr1[i] = (lane + 32 * i < m1.len) ? __ldg(col + m1.p0 + lane + 32 * i) : 0;
        }
        AsmMeta m2 = m1;
        if (e2 >= 0) m2 = asm_load_meta(e2, nEF, adjptr, rowptr);
        const int e3 = item_at(w + 3 * stride);
        const int deg = m.deg, len = m.len;
#pragma unroll
        for (int i = 0; i < CREG; ++i) {
            const int k = lane + 32 * i;
            if (k < len) {
                wb.cols[k] = colr[i];
                const double2 z = make_double2(0.0, 0.0);
                wb.acc[0][k] = z; wb.acc[1][k] = z; wb.acc[2][k] = z; wb.acc[3][k] = z;
            }
        }
        bool next_fetched = false;
        for (int n = 0; n < deg; ++n) {
            const int ic = __shfl_sync(0xffffffffu, myadj, n) % 20;      // canonical local row of the entity's first function
            wb.rec[lane] = r0;
            if (lane < 24) wb.rec[32 + lane] = r1;
            const int c = cj;
            __syncwarp();
            // next record: of this entity, or the first one of the next entity
            if (n + 1 < deg) {
                fetch_rec(__shfl_sync(0xffffffffu, myadj, n + 1) / 20);
            } else if (e1 >= 0 && m1.deg > 0) {
                fetch_rec(__shfl_sync(0xffffffffu, adj1, 0) / 20);
                next_fetched = true;
            }
            if (colLane) {
                // position of column c in the row's sorted column list (present by construction)
                int lo = 0, mm = len;
                while (mm > 1) {
                    const int half = mm >> 1;
                    lo = (wb.cols[lo + half] <= c) ? lo + half : lo;
                    mm -= half;
                }
                // curl-curl rows: nk uniform terms (3: edge, 9: face), coefficient pairs and D indices from the tables
                double kar = 0, kai = 0, kbr = 0, kbi = 0;
                unsigned long long ki = T->kidx[ic][J];
                const double2* kc = reinterpret_cast<const double2*>(&T->kc[ic][0][J][0]);
                const int nk = T->nk[ic];
#pragma unroll 3
                for (int k = 0; k < nk; ++k) {
                    const double2 d = recD[(int)(ki & 63ull)];
                    ki >>= 6;
                    const double2 cc = kc[k * 20];
                    kar += cc.x * d.x; kai += cc.x * d.y;
                    kbr += cc.y * d.x; kbi += cc.y * d.y;
                }
                const double lj = recLen[lpj];
                const double sa = recLen[T->lp[ic]] * lj, sb = recLen[T->lp[ic + 10]] * lj;
                // mass rows: four g terms each
                const unsigned gi = T->gidx[ic][J];
                const double2 ca0 = *reinterpret_cast<const double2*>(&T->mc[ic][0][0][J][0]);
                const double2 ca1 = *reinterpret_cast<const double2*>(&T->mc[ic][0][1][J][0]);
                const double2 cb0 = *reinterpret_cast<const double2*>(&T->mc[ic][1][0][J][0]);
                const double2 cb1 = *reinterpret_cast<const double2*>(&T->mc[ic][1][1][J][0]);
                double2 g0 = recG[gi & 15u], g1 = recG[(gi >> 4) & 15u], g2 = recG[(gi >> 8) & 15u], g3 = recG[(gi >> 12) & 15u];
                double mar = ca0.x * g0.x, mai = ca0.x * g0.y;
                mar += ca0.y * g1.x; mai += ca0.y * g1.y;
                mar += ca1.x * g2.x; mai += ca1.x * g2.y;
                mar += ca1.y * g3.x; mai += ca1.y * g3.y;
                g0 = recG[(gi >> 16) & 15u]; g1 = recG[(gi >> 20) & 15u]; g2 = recG[(gi >> 24) & 15u]; g3 = recG[(gi >> 28) & 15u];
                double mbr = cb0.x * g0.x, mbi = cb0.x * g0.y;
                mbr += cb0.y * g1.x; mbi += cb0.y * g1.y;
                mbr += cb1.x * g2.x; mbi += cb1.x * g2.y;
                mbr += cb1.y * g3.x; mbi += cb1.y * g3.y;
                double2 v;
                v = wb.acc[0][lo]; v.x += sa * kar; v.y += sa * kai; wb.acc[0][lo] = v;
                v = wb.acc[1][lo]; v.x += sb * kbr; v.y += sb * kbi; wb.acc[1][lo] = v;
                v = wb.acc[2][lo]; v.x += sa * mar; v.y += sa * mai; wb.acc[2][lo] = v;
                v = wb.acc[3][lo]; v.x += sb * mbr; v.y += sb * mbi; wb.acc[3][lo] = v;
            }
            __syncwarp();
        }
        if (!next_fetched && e1 >= 0 && m1.deg > 0) fetch_rec(__shfl_sync(0xffffffffu, adj1, 0) / 20);
        double2* Ka = reinterpret_cast<double2*>(K + m.p0);
        double2* Kb = reinterpret_cast<double2*>(K + m.p1);
        double2* Ma = reinterpret_cast<double2*>(M + m.p0);
        double2* Mb = reinterpret_cast<double2*>(M + m.p1);
#pragma unroll
        for (int i = 0; i < CREG; ++i) {
            const int k = lane + 32 * i;
            if (k < len) { Ka[k] = wb.acc[0][k]; Kb[k] = wb.acc[1][k]; Ma[k] = wb.acc[2][k]; Mb[k] = wb.acc[3][k]; }
        }
        __syncwarp();
        m = m1; m1 = m2; e1 = e2; e2 = e3; myadj = adj1;
#pragma unroll
        for (int i = 0; i < CREG; ++i) colr[i] = colr1[i];
    }
}

// entities whose rows do not fit the shared-memory classes (or with more than 32 tetrahedra): accumulators in K / M
// themselves, column search in global memory.  Same summation order.
__global__ void __launch_bounds__(128) k_asm_rows_big(int64_t nitems, const int* __restrict__ items, int nEF,
                                                      const int64_t* __restrict__ adjptr, const int* __restrict__ adj,
                                                      const int* __restrict__ gid, const int64_t* __restrict__ rowptr,
                                                      const int* __restrict__ col, const ned2f::TetRec* __restrict__ recs,
                                                      cx* __restrict__ K, cx* __restrict__ M) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t w = blockIdx.x * (int64_t)4 + warp;
    if (w >= nitems) return;
    const ned2f::KernTables& T = d_ktab;
    const int e = items[w];
    const int64_t a0 = adjptr[e], a1 = adjptr[e + 1];
    const int64_t p0 = rowptr[e], p1 = rowptr[e + nEF];
    const int len = (int)(rowptr[e + 1] - p0);
    for (int k = lane; k < len; k += 32) {
        st16(K + p0 + k, cx{0, 0}); st16(K + p1 + k, cx{0, 0});
        st16(M + p0 + k, cx{0, 0}); st16(M + p1 + k, cx{0, 0});
    }
    __syncwarp();
    for (int64_t ai = a0; ai < a1; ++ai) {
        const int a = adj[ai];
        const int t = a / 20, ic = a % 20;
        if (lane < 20) {
            const int c = gid[(int64_t)t * 20 + lane];
            int lo = 0, m = len;
            while (m > 1) {
                const int half = m >> 1;
                lo = (col[p0 + lo + half] <= c) ? lo + half : lo;
                m -= half;
            }
            const cx* D = reinterpret_cast<const cx*>(recs + t);
            cx Ka, Kb, Ma, Mb;
            ned2f::row_pair_flat(T, ic, lane, D, D + 36, reinterpret_cast<const double*>(D + 52), Ka, Kb, Ma, Mb);
            st16(K + p0 + lo, K[p0 + lo] + Ka); st16(K + p1 + lo, K[p1 + lo] + Kb);
            st16(M + p0 + lo, M[p0 + lo] + Ma); st16(M + p1 + lo, M[p1 + lo] + Mb);
        }
        __syncwarp();
    }
}

template <int ROWCAP, int WARPS>
static int launch_asm_rows(emb_ctx* c, int64_t i0, int64_t i1, const ned2f::TetRec* recs) {
    if (i1 <= i0) return EMB_OK;
    constexpr size_t TOFF = (sizeof(ned2f::KernTables) + 15) / 16 * 16;
    const size_t smem = TOFF + (size_t)WARPS * sizeof(AsmWarpBuf<ROWCAP>);
    auto kern = k_asm_rows<ROWCAP, WARPS>;
    EMB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, nsm = 0;
    EMB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    EMB_CUDA(c, cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)nsm * per_sm;
    const int64_t need = (i1 - i0 + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, WARPS * 32, smem, c->stream>>>(i1 - i0, c->asm_ent.p + i0, (int)(c->nE + c->nTri), c->adjptr.p,
                                                        c->adj.p, c->gid.p, c->rowptr.p, c->col.p, recs, c->K.p, c->M.p);
    EMB_LAUNCH_CHECK(c);
    return EMB_OK;
}

// one-time (per pattern) work list of the fused numeric phase
static int build_entity_order(emb_ctx* c) {
    const int nEF = (int)(c->nE + c->nTri);
    c->asm_pairs_ok = false;
    DevBuf<int> bad, val;
    DevBuf<unsigned> key, key2;
    DevBuf<char> tmp;
    EMB_TRY(dev_alloc(c, bad, 1));
    EMB_CUDA(c, cudaMemsetAsync(bad.p, 0, sizeof(int), c->stream));
    k_check_pairs<<<blocks_for(c->nT * 10, 256), 256, 0, c->stream>>>(c->nT, nEF, c->gid.p, bad.p);
    EMB_LAUNCH_CHECK(c);
    EMB_TRY(dev_alloc(c, key, (size_t)nEF));
    EMB_TRY(dev_alloc(c, key2, (size_t)nEF));
    EMB_TRY(dev_alloc(c, val, (size_t)nEF));
    EMB_TRY(dev_alloc(c, c->asm_ent, (size_t)nEF));
    k_entity_keys<<<blocks_for(nEF, 256), 256, 0, c->stream>>>(nEF, c->adjptr.p, c->adj.p, c->rowptr.p, key.p, val.p, bad.p);
    EMB_LAUNCH_CHECK(c);
    size_t tb = 0;
    EMB_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, key2.p, val.p, c->asm_ent.p, nEF, 0, 30, c->stream));
    EMB_TRY(dev_alloc(c, tmp, tb));
    EMB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.p, key2.p, val.p, c->asm_ent.p, nEF, 0, 30, c->stream));
    c->launches += 4;
    // class boundaries: first position whose key is >= cls << 28
    DevBuf<int64_t> ptr;
    DevBuf<int> cls;
    EMB_TRY(dev_alloc(c, ptr, 5));
    EMB_TRY(dev_alloc(c, cls, (size_t)nEF));
    k_key_class<<<blocks_for(nEF, 256), 256, 0, c->stream>>>(nEF, key2.p, cls.p);
    EMB_LAUNCH_CHECK(c);
    k_segptr<<<blocks_for(nEF, 256), 256, 0, c->stream>>>(cls.p, nEF, 3, ptr.p);
    EMB_LAUNCH_CHECK(c);
    int hbad = 0;
    EMB_CUDA(c, cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaMemcpyAsync(c->asm_cls_ptr, ptr.p, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    bad.release(); val.release(); key.release(); key2.release(); tmp.release(); ptr.release(); cls.release();
    c->asm_pairs_ok = hbad == 0;
    return EMB_OK;
}

static int assemble_fused(emb_ctx* c) {
    DevBuf<ned2f::TetRec> recs;
    EMB_TRY(dev_alloc(c, recs, (size_t)c->nT));
    {
        PhaseTimer pt(c, "tet_kernel");
        k_tet_records<<<blocks_for(c->nT, 128), 128, 0, c->stream>>>(c->nT, c->tetc.p, c->nodes.p, c->er.p, c->ur.p, recs.p);
        EMB_LAUNCH_CHECK(c);
    }
    const double ms_rec = c->ms["tet_kernel"];
    {
        PhaseTimer pt(c, "reduce");
        const int64_t* cp = c->asm_cls_ptr;
        static const int warps_env = getenv("EMB_ASM_WARPS") ? atoi(getenv("EMB_ASM_WARPS")) : 0;     // tuning probe
        if (warps_env == 16) EMB_TRY((launch_asm_rows<ACLS0, 16>(c, cp[0], cp[1], recs.p)));
        else if (warps_env == 24) EMB_TRY((launch_asm_rows<ACLS0, 24>(c, cp[0], cp[1], recs.p)));
        else EMB_TRY((launch_asm_rows<ACLS0, AWARPS0>(c, cp[0], cp[1], recs.p)));
        EMB_TRY((launch_asm_rows<ACLS1, AWARPS1>(c, cp[1], cp[2], recs.p)));
        if (cp[3] > cp[2]) {
            k_asm_rows_big<<<blocks_for(cp[3] - cp[2], 4), 128, 0, c->stream>>>(cp[3] - cp[2], c->asm_ent.p + cp[2],
                                                                              (int)(c->nE + c->nTri), c->adjptr.p, c->adj.p,
                                                                              c->gid.p, c->rowptr.p, c->col.p, recs.p, c->K.p,
                                                                              c->M.p);
            EMB_LAUNCH_CHECK(c);
        }
    }
    c->ms["assemble"] = ms_rec + c->ms["reduce"];
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    recs.release();
    return EMB_OK;
}

// ---- chunk work lists ------------------------------------------------------------------------------------
// The numeric phase runs chunk by chunk over the tetrahedra so that the COO scratch of one chunk (12.8 KB per tet)
// stays resident in the 126 MB L2 between the element kernel that writes it and the reduction that reads it: HBM then
// only sees the inputs and the K/M rows.  A row is touched by every chunk that owns one of its tets; the partial sums
// are carried in K/M themselves, chunks run in ascending tet order, so the additions happen in exactly the order of
// the single-pass reduction (bitwise identical result).
__global__ void k_count_items(int64_t N, int64_t CT, const int64_t* __restrict__ adjptr, const int* __restrict__ adj,
                              int* __restrict__ cnt) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= N) return;
    int n = 0;
    int64_t prev = -1;
    for (int64_t ai = adjptr[r]; ai < adjptr[r + 1]; ++ai) {
        const int64_t ch = (adj[ai] / 20) / CT;
        if (ch != prev) { ++n; prev = ch; }
    }
    cnt[r] = n;
}
__global__ void k_fill_items(int64_t N, int64_t CT, const int64_t* __restrict__ adjptr, const int* __restrict__ adj,
                             const int* __restrict__ off, int* __restrict__ key, unsigned long long* __restrict__ val) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= N) return;
    int o = off[r];
    int64_t prev = -1;
    for (int64_t ai = adjptr[r]; ai < adjptr[r + 1]; ++ai) {
        const int64_t ch = (adj[ai] / 20) / CT;
        if (ch != prev) {
            key[o] = (int)ch;
            val[o] = ((unsigned long long)r << 32) | (unsigned long long)(unsigned)ai;
            ++o;
            prev = ch;
        }
    }
}

static int build_chunk_items(emb_ctx* c, int64_t CT) {
    if (c->asm_chunk == CT && c->asm_items.p) return EMB_OK;
    const int64_t N = c->N;
    const int64_t nch = (c->nT + CT - 1) / CT;
    DevBuf<int> cnt, off, key, key2;
    DevBuf<unsigned long long> val;
    DevBuf<char> tmp;
    EMB_TRY(dev_alloc(c, cnt, (size_t)N + 1));
    EMB_TRY(dev_alloc(c, off, (size_t)N + 1));
    EMB_CUDA(c, cudaMemsetAsync(cnt.p + N, 0, sizeof(int), c->stream));
    k_count_items<<<blocks_for(N, 256), 256, 0, c->stream>>>(N, CT, c->adjptr.p, c->adj.p, cnt.p);
    EMB_LAUNCH_CHECK(c);
    size_t tb = 0;
    EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, off.p, (int)(N + 1), c->stream));
    EMB_TRY(dev_alloc(c, tmp, tb));
    EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.p, off.p, (int)(N + 1), c->stream));
    int nitems = 0;
    EMB_CUDA(c, cudaMemcpyAsync(&nitems, off.p + N, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    EMB_TRY(dev_alloc(c, key, (size_t)nitems));
    EMB_TRY(dev_alloc(c, key2, (size_t)nitems));
    EMB_TRY(dev_alloc(c, val, (size_t)nitems));
    EMB_TRY(dev_alloc(c, c->asm_items, (size_t)nitems));
    k_fill_items<<<blocks_for(N, 256), 256, 0, c->stream>>>(N, CT, c->adjptr.p, c->adj.p, off.p, key.p, val.p);
    EMB_LAUNCH_CHECK(c);
    int bits = 1;
    while (((int64_t)1 << bits) < nch) ++bits;
    // stable sort by chunk: inside a chunk the items stay in ascending row order
    size_t tb2 = 0;
    EMB_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tb2, key.p, key2.p, val.p, c->asm_items.p, nitems, 0, bits, c->stream));
    if (tb2 > tmp.n) EMB_TRY(dev_alloc(c, tmp, tb2));
    EMB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tb2, key.p, key2.p, val.p, c->asm_items.p, nitems, 0, bits, c->stream));
    DevBuf<int64_t> ptr;
    EMB_TRY(dev_alloc(c, ptr, (size_t)nch + 1));
    k_segptr<<<blocks_for(nitems, 256), 256, 0, c->stream>>>(key2.p, nitems, nch, ptr.p);
    EMB_LAUNCH_CHECK(c);
    c->launches += 6;
    c->asm_chunk_ptr.resize((size_t)nch + 1);
    EMB_CUDA(c, cudaMemcpyAsync(c->asm_chunk_ptr.data(), ptr.p, (size_t)(nch + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    cnt.release(); off.release(); key.release(); key2.release(); val.release(); tmp.release(); ptr.release();
    c->asm_chunk = CT;
    return EMB_OK;
}

extern "C" int emb_assemble_mode(emb_ctx* c, int mode) {
    if (!c || mode < 0 || mode > 1) return EMB_ERR_ARG;
    c->asm_mode = mode;
    return EMB_OK;
}

extern "C" int emb_assemble_config(emb_ctx* c, int64_t chunk_tets, int persist_l2) {
    if (!c || chunk_tets < 0) return EMB_ERR_ARG;
    c->asm_chunk_req = chunk_tets;
    c->asm_persist = persist_l2 != 0;
    return EMB_OK;
}

extern "C" int emb_assemble_KM(emb_ctx* c) {
    if (!c || !c->have_pattern || !c->have_mat) {
        if (c) c->err = "emb_assemble_KM: needs emb_symbolic and emb_upload_materials first";
        return c ? EMB_ERR_STATE : EMB_ERR_ARG;
    }
    EMB_TRY(dev_alloc(c, c->K, (size_t)c->nnz));
    EMB_TRY(dev_alloc(c, c->M, (size_t)c->nnz));
    int mode = c->asm_mode;
    if (const char* e = getenv("EMB_ASM_MODE")) mode = atoi(e);
    if (mode == 0 && c->asm_pairs_ok) {
        EMB_TRY(assemble_fused(c));
        c->have_KM = true;
        c->have_A = false;
        return EMB_OK;
    }
    // chunk size in tets: 0 (default) = single pass, the COO scratch makes one round trip through HBM.  Chunks of one
    // wave of element-kernel blocks (32 tets x 148 SMs = 60 MB of COO, pinned in L2) were measured on a B200 and LOSE
    // (1M tets: 16.6 ms vs 14.4 ms, profiles/r1_asm_chunk_sweep.json): one block per SM leaves the element kernel
    // latency-bound and the per-chunk reductions re-read partially summed rows.  The option stays for experiments.
    int64_t CT = c->asm_chunk_req;
    if (const char* e = getenv("EMB_ASM_CHUNK")) CT = atoll(e);
    if (CT <= 0) CT = c->nT;
    const int64_t max_chunk = (int64_t)(48.0e9 / 12800.0);      // single-pass scratch stays below ~48 GB
    if (CT > max_chunk) CT = max_chunk;
    const bool chunked = CT < c->nT;
    const int64_t chunk = chunked ? CT : c->nT;
    DevBuf<cx> coo;                                              // [K | M] of one chunk, one allocation (one L2 window)
    EMB_TRY(dev_alloc(c, coo, (size_t)chunk * 800));
    cx* cooK = coo.p;
    cx* cooM = coo.p + chunk * 400;
    if (!chunked) {
        double ms_tet = 0, ms_red = 0;
        {
            PhaseTimer pt(c, "tet_kernel");
            k_tet<<<blocks_for(c->nT, 32), TWARPS * 32, 0, c->stream>>>(0, c->nT, c->nT, c->tetc.p, c->nodes.p, c->er.p, c->ur.p,
                                                                       cooK, cooM);
            EMB_LAUNCH_CHECK(c);
        }
        ms_tet = c->ms["tet_kernel"];
        {
            PhaseTimer pt(c, "reduce");
            k_reduce_rows<<<blocks_for(c->N, RWARPS), RWARPS * 32, 0, c->stream>>>(
                c->N, 0, c->nT, 1, nullptr, c->rowptr.p, c->col.p, c->adjptr.p, c->adj.p, c->gid.p, cooK, cooM, c->K.p, c->M.p);
            EMB_LAUNCH_CHECK(c);
        }
        ms_red = c->ms["reduce"];
        c->ms["assemble"] = ms_tet + ms_red;
    } else {
        EMB_TRY(build_chunk_items(c, CT));
        bool window = false;
        bool persist = c->asm_persist;
        if (const char* e = getenv("EMB_ASM_PERSIST")) persist = atoi(e) != 0;
        if (persist) {
            cudaDeviceProp prop;
            if (cudaGetDeviceProperties(&prop, c->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
                const size_t bytes = (size_t)chunk * 800 * sizeof(cx);
                const size_t carve = bytes < (size_t)prop.persistingL2CacheMaxSize ? bytes : (size_t)prop.persistingL2CacheMaxSize;
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
                cudaStreamAttrValue av;
                memset(&av, 0, sizeof(av));
                av.accessPolicyWindow.base_ptr = coo.p;
                av.accessPolicyWindow.num_bytes = bytes < (size_t)prop.accessPolicyMaxWindowSize ? bytes : (size_t)prop.accessPolicyMaxWindowSize;
                av.accessPolicyWindow.hitRatio = (float)((double)carve / (double)av.accessPolicyWindow.num_bytes);
                if (av.accessPolicyWindow.hitRatio > 1.0f) av.accessPolicyWindow.hitRatio = 1.0f;
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                window = cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
                cudaGetLastError();
            }
        }
        {
            PhaseTimer pt(c, "assemble");
            const int64_t nch = (c->nT + CT - 1) / CT;
            for (int64_t ch = 0; ch < nch; ++ch) {
                const int64_t t0 = ch * CT, t1 = (t0 + CT < c->nT) ? t0 + CT : c->nT;
                k_tet<<<blocks_for(t1 - t0, 32), TWARPS * 32, 0, c->stream>>>(t0, t1, c->nT, c->tetc.p, c->nodes.p, c->er.p,
                                                                             c->ur.p, cooK, cooM);
                EMB_LAUNCH_CHECK(c);
                const int64_t i0 = c->asm_chunk_ptr[(size_t)ch], i1 = c->asm_chunk_ptr[(size_t)ch + 1];
                if (i1 > i0) {
                    k_reduce_rows<<<blocks_for(i1 - i0, RWARPS), RWARPS * 32, 0, c->stream>>>(
                        i1 - i0, t0, t1, 0, c->asm_items.p + i0, c->rowptr.p, c->col.p, c->adjptr.p, c->adj.p, c->gid.p, cooK,
                        cooM, c->K.p, c->M.p);
                    EMB_LAUNCH_CHECK(c);
                }
            }
        }
        if (window) {
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof(av));
            cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &av);
            cudaCtxResetPersistingL2Cache();
            cudaGetLastError();
        }
        c->ms.erase("tet_kernel");
        c->ms.erase("reduce");
    }
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    coo.release();
    c->have_KM = true;
    c->have_A = false;
    return EMB_OK;
}

// element matrices in the reference's local order and slot layout (parity/debug)
__global__ void k_to_ref_order(int64_t n, int64_t t0, const int* __restrict__ tetord, const cx* __restrict__ canK,
                               const cx* __restrict__ canM, cx* __restrict__ refK, cx* __restrict__ refM) {
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= n * 400) return;
    int64_t t = idx / 400;
    int ij = (int)(idx % 400), i = ij / 20, j = ij % 20;
    int po = tetord[t0 + t];
    int ord[4] = {po & 3, (po >> 2) & 3, (po >> 4) & 3, (po >> 6) & 3};
    int ref[20];
    ned2::canonical_to_ref(ord, ref);
    refK[t * 400 + ref[i] * 20 + ref[j]] = canK[idx];
    refM[t * 400 + ref[i] * 20 + ref[j]] = canM[idx];
}

extern "C" int emb_element_matrices(emb_ctx* c, int64_t t0, int64_t t1, emb_c128* E400, emb_c128* B400) {
    if (!c || !c->have_mesh || !c->have_mat || t0 < 0 || t1 > c->nT || t1 <= t0) {
        if (c) c->err = "emb_element_matrices: bad state or range";
        return c ? EMB_ERR_STATE : EMB_ERR_ARG;
    }
    const int64_t n = t1 - t0;
    DevBuf<cx> cK, cM, rK, rM;
    EMB_TRY(dev_alloc(c, cK, (size_t)n * 400));
    EMB_TRY(dev_alloc(c, cM, (size_t)n * 400));
    EMB_TRY(dev_alloc(c, rK, (size_t)n * 400));
    EMB_TRY(dev_alloc(c, rM, (size_t)n * 400));
    k_tet<<<blocks_for(n, 32), TWARPS * 32, 0, c->stream>>>(t0, t1, c->nT, c->tetc.p, c->nodes.p, c->er.p, c->ur.p, cK.p, cM.p);
    EMB_LAUNCH_CHECK(c);
    k_to_ref_order<<<blocks_for(n * 400, 256), 256, 0, c->stream>>>(n, t0, c->tetord.p, cK.p, cM.p, rK.p, rM.p);
    EMB_LAUNCH_CHECK(c);
    EMB_CUDA(c, cudaMemcpyAsync(E400, rK.p, (size_t)n * 400 * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaMemcpyAsync(B400, rM.p, (size_t)n * 400 * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    cK.release(); cM.release(); rK.release(); rM.release();
    return EMB_OK;
}
