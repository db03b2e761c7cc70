// Inner operator As (complex64, 2x2 blocks) in a sliced-ELL layout: SELL-8-sigma over BLOCK-ROWS.
//
// Why: in block-CSR (krylov.cuh::k_bspmv) every lane group walks its own block-row, so one warp-wide load of values or
// column indices touches 8 different cache lines and uses 16 of each 32-byte sector; ncu (profiles/r1_*): l1tex 80 %,
// 86 % long-scoreboard stalls, 50 % of the HBM roofline although DRAM traffic is only 1.08 x the algorithmic bytes.  Here
// the block-rows are grouped in slices of 8 with the blocks of a slice interleaved (slot = sptr[s] + 8 q + r for block q
// of slice-row r), so the same warp-wide load reads 8 consecutive 32-byte blocks (two full lines) and 8 consecutive
// column indices (one sector); what is left for the load/store unit are the x gathers.  Slices are padded to their
// longest block-row; block-rows are sorted by length inside windows of sigma rows first (waveguide mesh: padding 11 %
// unsorted, 1.7 % with sigma = 64, 0.5 % with 256, 0.2 % with 1024; measured 0.818 / 0.853 / 0.804 / 0.816 ms for sigma =
// 8 / 64 / 256 / 1024 with the staged kernel at 1M tets, profiles/r2_spmv_sell_tuning.txt), which only permutes which rows
// a warp owns: the numbering of x and y is untouched.
// Block storage is column-major, [h][r]: the two values a lane needs (rows 2j, 2j+1 of column 2c+h) are one 16-byte load.
// The reference has no counterpart (it factorises, fem/solver.py:243-309).
#pragma once
#include "krylov.cuh"
#include <cub/cub.cuh>

constexpr int SELL_C = 8;
constexpr int SELL_SIGMA_DEFAULT = 256;     // sorting window in block-rows (EMB_SELL_SIGMA overrides; profiles/r2_spmv_sell_tuning.txt)
constexpr int SELL_LENBITS = 12;

__global__ void k_sell_keys(int nbr, int sigma, const int64_t* __restrict__ rowptr_s, unsigned* __restrict__ key,
                            int* __restrict__ val, int* __restrict__ flag) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nbr) return;
    const int64_t len = (rowptr_s[2 * j + 1] - rowptr_s[2 * j]) >> 1;
    if (len >= (1 << SELL_LENBITS)) atomicExch(flag, 1);
    key[j] = ((unsigned)(j / sigma) << SELL_LENBITS) | (unsigned)(len & ((1 << SELL_LENBITS) - 1));
    val[j] = j;
}
// rows[R] = block-row of slice-row R (or -1), pos[j] = R, slen[s] = 8 * (longest block-row of slice s)
__global__ void k_sell_slices(int nbr, int nslices, const int* __restrict__ sorted, const int64_t* __restrict__ rowptr_s,
                              int* __restrict__ rows, int* __restrict__ pos, int64_t* __restrict__ slen, int* __restrict__ empty) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > nslices) return;
    if (s == nslices) { slen[s] = 0; return; }
    int64_t mx = 0;
    for (int r = 0; r < SELL_C; ++r) {
        const int R = s * SELL_C + r;
        int j = -1;
        if (R < nbr) {
            j = sorted[R];
            pos[j] = R;
            const int64_t len = (rowptr_s[2 * j + 1] - rowptr_s[2 * j]) >> 1;
            mx = len > mx ? len : mx;
        }
        rows[R] = j;
    }
    slen[s] = mx * SELL_C;
    if (mx == 0) atomicExch(empty, 1);
}
__global__ void k_sell_bcol(int nslices, const int* __restrict__ rows, const int64_t* __restrict__ sptr,
                            const int64_t* __restrict__ rowptr_s, const int* __restrict__ blkcol, int* __restrict__ bcol) {
    const int lane = threadIdx.x & 31;
    const int64_t R = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (R >= (int64_t)nslices * SELL_C) return;
    const int s = (int)(R / SELL_C), r = (int)(R % SELL_C);
    const int64_t base = sptr[s];
    const int nb = (int)((sptr[s + 1] - base) / SELL_C);
    const int j = rows[R];
    int64_t p0 = 0;
    int len = 0;
    if (j >= 0) { p0 = rowptr_s[2 * j]; len = (int)((rowptr_s[2 * j + 1] - p0) >> 1); }
    const int padcol = len > 0 ? blkcol[p0 >> 2] : 0;        // padding blocks carry zero values: any valid column will do
    for (int q = lane; q < nb; q += 32) bcol[base + (int64_t)q * SELL_C + r] = q < len ? blkcol[(p0 >> 2) + q] : padcol;
}

static int sell_build(emb_ctx* c) {
    c->sell_ready = false;
    if (!c->paired) return EMB_OK;
    const int nbr = (int)(c->Ns / 2);
    const int nslices = (nbr + SELL_C - 1) / SELL_C;
    DevBuf<unsigned> key, key2;
    DevBuf<int> val, sorted, flag;
    DevBuf<int64_t> slen;
    DevBuf<char> tmp;
    EMB_TRY(dev_alloc(c, key, (size_t)nbr));
    EMB_TRY(dev_alloc(c, key2, (size_t)nbr));
    EMB_TRY(dev_alloc(c, val, (size_t)nbr));
    EMB_TRY(dev_alloc(c, sorted, (size_t)nbr));
    EMB_TRY(dev_alloc(c, flag, 2));
    EMB_CUDA(c, cudaMemsetAsync(flag.p, 0, 2 * sizeof(int), c->stream));
    int sigma = getenv("EMB_SELL_SIGMA") ? atoi(getenv("EMB_SELL_SIGMA")) : SELL_SIGMA_DEFAULT;
    if (sigma < SELL_C) sigma = SELL_C;
    k_sell_keys<<<blocks_for(nbr, 256), 256, 0, c->stream>>>(nbr, sigma, c->rowptr_s.p, key.p, val.p, flag.p);
    EMB_LAUNCH_CHECK(c);
    int wbits = 1;
    while (((int64_t)1 << wbits) < (nbr + sigma - 1) / sigma + 1) ++wbits;
    if (SELL_LENBITS + wbits > 32) { c->err = "sell_build: sorting window too small for this many rows"; return EMB_ERR_LIMIT; }
    size_t tb = 0;
    EMB_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, key2.p, val.p, sorted.p, nbr, 0, SELL_LENBITS + wbits, c->stream));
    EMB_TRY(dev_alloc(c, tmp, tb));
    EMB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.p, key2.p, val.p, sorted.p, nbr, 0, SELL_LENBITS + wbits, c->stream));
    EMB_TRY(dev_alloc(c, c->sell_rows, (size_t)nslices * SELL_C));
    EMB_TRY(dev_alloc(c, c->sell_pos, (size_t)nbr));
    EMB_TRY(dev_alloc(c, slen, (size_t)nslices + 1));
    EMB_TRY(dev_alloc(c, c->sell_sptr, (size_t)nslices + 1));
    k_sell_slices<<<blocks_for(nslices + 1, 256), 256, 0, c->stream>>>(nbr, nslices, sorted.p, c->rowptr_s.p, c->sell_rows.p,
                                                                      c->sell_pos.p, slen.p, flag.p + 1);
    EMB_LAUNCH_CHECK(c);
    size_t tb2 = 0;
    EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tb2, slen.p, c->sell_sptr.p, nslices + 1, c->stream));
    if (tb2 > tmp.n) EMB_TRY(dev_alloc(c, tmp, tb2));
    EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, tb2, slen.p, c->sell_sptr.p, nslices + 1, c->stream));
    c->launches += 4;
    int hflags[2] = {0, 0};
    int64_t total = 0;
    EMB_CUDA(c, cudaMemcpyAsync(hflags, flag.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaMemcpyAsync(&total, c->sell_sptr.p + nslices, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    key.release(); key2.release(); val.release(); sorted.release(); flag.release(); slen.release(); tmp.release();
    const int hflag = hflags[0];
    c->sell_has_empty = hflags[1] != 0;
    if (hflag || total >= ((int64_t)1 << 31)) {           // a block-row of >= 4096 blocks / int32 slots: keep block-CSR
        c->sell_rows.release(); c->sell_pos.release(); c->sell_sptr.release();
        return EMB_OK;
    }
    c->sell_nslices = nslices;
    c->sell_blocks = total;
    EMB_TRY(dev_alloc(c, c->sell_bcol, (size_t)total));
    k_sell_bcol<<<blocks_for((int64_t)nslices * SELL_C * 32, 256), 256, 0, c->stream>>>(nslices, c->sell_rows.p, c->sell_sptr.p,
                                                                                       c->rowptr_s.p, c->blkcol.p, c->sell_bcol.p);
    EMB_LAUNCH_CHECK(c);
    c->sell_ready = true;
    return EMB_OK;
}

// As = (A + A^T) / 2 in complex64, written straight into the SELL layout.  One warp per block-row; the transposed
// entry is found by binary search in the column list of the other row (the pattern is structurally symmetric).
__global__ void k_sym_part_sell(int64_t nbr, const int64_t* __restrict__ rowptr, const int* __restrict__ col,
                                const cx* __restrict__ A, const int* __restrict__ pos, const int64_t* __restrict__ sptr,
                                cf* __restrict__ As) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (j >= nbr) return;
    const int R = pos[j];
    const int64_t base = sptr[R / SELL_C];
    const int r = R % SELL_C;
    const int64_t p0 = rowptr[2 * j], p1 = rowptr[2 * j + 1];
    const int n2 = (int)(p1 - p0);                      // entries per row = 2 x blocks
    for (int e2 = lane; e2 < 2 * n2; e2 += 32) {
        const int rr = e2 >= n2 ? 1 : 0, e = e2 - rr * n2;
        const int64_t k = (rr ? p1 : p0) + e;
        const int row = (int)(2 * j + rr);
        const int cj = col[k];
        int64_t lo = rowptr[cj], hi = rowptr[cj + 1] - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (col[mid] < row) lo = mid + 1; else hi = mid;
        }
        cx a = A[k];
        if (col[lo] == row) { const cx b = A[lo]; a = cx{0.5 * (a.re + b.re), 0.5 * (a.im + b.im)}; }
        const int q = e >> 1, h = e & 1;
        stval(As, (base + (int64_t)q * SELL_C + r) * 4 + h * 2 + rr, a);
    }
}

// y = As x on NV interleaved right-hand sides.  Thread (R, u): slice-row R = block-row rows[R], u = (h, k) = column half
// and right-hand side; per block one column index (broadcast over the 2 NV lanes of the row), one 16-byte piece of x, one
// 16-byte pair of values.  NV = 2: a warp is exactly one slice.
template <int NV>
__global__ void __launch_bounds__(256) k_bsell(int64_t nR, const int* __restrict__ rows, const int64_t* __restrict__ sptr,
                                               const int* __restrict__ bcol, const float4* __restrict__ val,
                                               const cx* __restrict__ x, cx* __restrict__ y) {
    constexpr int LPB = 2 * NV;
    const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t R = gt / LPB;
    const int u = (int)(gt % LPB);
    const int h = u / NV;
    const bool live = R < nR;
    if (!live) R = nR - 1;
    const int64_t s = R / SELL_C;
    const int r = (int)(R % SELL_C);
    const int64_t base = __ldg(sptr + s);
    const int nb = live ? (int)((__ldg(sptr + s + 1) - base) / SELL_C) : 0;
    const int* bc = bcol + base + r;
    const float4* vv = val + (base + r) * 2 + h;
    double a0r = 0.0, a0i = 0.0, a1r = 0.0, a1i = 0.0;
#pragma unroll 4
    for (int q = 0; q < nb; ++q) {
        const int cb = __ldg(bc + q * SELL_C);
        const float4 e = __ldg(vv + q * (SELL_C * 2));
        const cx w = ldx(x + (int64_t)cb * LPB + u);
        a0r += (double)e.x * w.re - (double)e.y * w.im;
        a0i += (double)e.x * w.im + (double)e.y * w.re;
        a1r += (double)e.z * w.re - (double)e.w * w.im;
        a1i += (double)e.z * w.im + (double)e.w * w.re;
    }
    a0r += __shfl_xor_sync(0xffffffffu, a0r, NV);
    a0i += __shfl_xor_sync(0xffffffffu, a0i, NV);
    a1r += __shfl_xor_sync(0xffffffffu, a1r, NV);
    a1i += __shfl_xor_sync(0xffffffffu, a1i, NV);
    const int j = live ? __ldg(rows + R) : -1;
    if (j >= 0 && h == 0) {
        stv(y, (int64_t)(2 * j) * NV + u, cx{a0r, a0i});
        stv(y, (int64_t)(2 * j + 1) * NV + u, cx{a1r, a1i});
    }
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- the same product with the slice streams staged by the TMA engine -------------------------------------------------
// The column indices and values of a slice are two CONTIGUOUS byte ranges.  One elected lane per warp copies them chunk by
// chunk (Q block columns) into the warp's own shared-memory ring with cp.async.bulk (SASS UBLKCP) completing on an
// mbarrier, DEPTH - 1 chunks ahead of the arithmetic; the lanes then read indices and values from shared memory and have
// the Q x gathers of a chunk in flight at once - the load/store unit only sees the gathers.  Persistent warps: warp g of
// the grid owns slices g, g + G, ...  (NV = 2: a warp = the 8 block-rows of a slice x 4 lanes = column half x right-hand side.)
//   Q      block columns per chunk (gathers a lane keeps in flight), WARPS warps per CTA (one CTA per SM)
//   DEPTH  chunk buffers per warp
//   HINT   bit 0: the value / index streams are copied with an L2 evict-first policy (they are read exactly once; x, which
//          is re-read by neighbouring rows, keeps the L2), bit 1: the x gathers carry an evict-last policy, bit 2: y is
//          written with streaming stores
// No dependent global load is left on a slice boundary: the [begin, end) offsets of the warp's next 32 slices are fetched
// by ONE load per lane and handed out by shuffles, the block-column count of every slice the copy cursor enters is parked
// in a 4-entry shared-memory ring for the arithmetic cursor, and the row ids of a slice are loaded when the slice starts
// (used when it ends).
// Measured at 1M tets (profiles/r2_spmv_sell_tuning.txt): first generation (Q = 16, 16 warps, offsets and row ids loaded
// on the boundary, no cache policy) 0.826 ms; boundary loads removed 0.72; + evict-first streams 0.70; Q = 8 with 24
// warps (80 registers) 0.596 ms = 4.7 TB/s; deeper rings, 28 / 32 warps, evict-last gathers, streaming stores: no gain.
// Rejected: accumulating the block Gram matrix x^T y (rho = Z^T A Z of block COCR) in the epilogue - the rows of x re-read
// at every slice end cost 0.18 ms, the separate k_gram pass it would replace 0.08 ms; gathers of the next chunk issued
// before the arithmetic of the current one (register-held, 1.0 ms); a Z-order numbering of the mesh; a blocked slice
// distribution; gathering from a complex64 copy of x (costs iterations).
constexpr int STG_Q = 8, STG_WARPS = 24, STG_DEPTH = 2, STG_HINT = 1;      // shipped configuration
template <int Q, int DEPTH>
struct __align__(128) SellStageQ {
    float4 val[DEPTH][Q * SELL_C * 2];
    int col[DEPTH][Q * SELL_C];
    unsigned long long bar[DEPTH];
    int nbq[4];
};
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, unsigned bytes, unsigned long long* bar,
                                              unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ cx ldx_hint(const cx* __restrict__ p, unsigned long long pol) {
    double a, b;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(a), "=d"(b) : "l"(p), "l"(pol));
    return cx{a, b};
}
__device__ __forceinline__ void st_stream(cx* p, cx v) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.re), "d"(v.im) : "memory");
}

template <int Q, int WARPS, int DEPTH, int HINT>
__global__ void __launch_bounds__(WARPS * 32, 1) k_bsell_tma(int nslices, const int* __restrict__ rows,
                                                             const int64_t* __restrict__ sptr, const int* __restrict__ bcol,
                                                             const float4* __restrict__ val, const cx* __restrict__ x,
                                                             cx* __restrict__ y) {
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char stage_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    SellStageQ<Q, DEPTH>& st = reinterpret_cast<SellStageQ<Q, DEPTH>*>(stage_raw)[warp];
    if (lane == 0)
        for (int b = 0; b < DEPTH; ++b) mbar_init(&st.bar[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int u = lane & 3, h = u >> 1, rl = lane >> 2;
    const int gw = blockIdx.x * WARPS + warp, G = gridDim.x * WARPS;
    const int n_my = gw < nslices ? (nslices - gw + G - 1) / G : 0;        // this warp owns slices gw, gw + G, ...
    unsigned long long pol = 0, polx = 0;
    if constexpr ((HINT & 1) != 0) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if constexpr ((HINT & 2) != 0) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(polx));
    long long mb = 0, me = 0;                    // offsets of the warp's slice number (batch * 32 + lane)
    int q_i = 0, nb_i = 0, i_i = -1, b_i = 0;    // copy cursor: slice number, next block column, block columns, next buffer
    int64_t base_i = 0;
    auto issue_next = [&]() {                    // all lanes call; lane 0 talks to the TMA engine
        if (i_i >= n_my) return;
        if (q_i >= nb_i) {                       // enter the warp's next slice
            ++i_i;
            if (i_i >= n_my) return;
            if ((i_i & 31) == 0) {
                const int k = i_i + lane;
                if (k < n_my) {
                    const int64_t s = gw + (int64_t)k * G;
                    mb = __ldg(sptr + s);
                    me = __ldg(sptr + s + 1);
                }
            }
            base_i = __shfl_sync(FULL, mb, i_i & 31);
            nb_i = (int)((__shfl_sync(FULL, me, i_i & 31) - base_i) / SELL_C);
            q_i = 0;
            if (lane == 0) st.nbq[i_i & 3] = nb_i;
        }
        const int nq = min(Q, nb_i - q_i);
        if (lane == 0) {
            const int64_t blk0 = base_i + (int64_t)q_i * SELL_C;
            const unsigned nblk = (unsigned)(nq * SELL_C);
            mbar_expect_tx(&st.bar[b_i], nblk * 36u);
            if constexpr ((HINT & 1) != 0) {
                bulk_g2s_hint(st.val[b_i], val + blk0 * 2, nblk * 32u, &st.bar[b_i], pol);
                bulk_g2s_hint(st.col[b_i], bcol + blk0, nblk * 4u, &st.bar[b_i], pol);
            } else {
                bulk_g2s(st.val[b_i], val + blk0 * 2, nblk * 32u, &st.bar[b_i]);
                bulk_g2s(st.col[b_i], bcol + blk0, nblk * 4u, &st.bar[b_i]);
            }
        }
        q_i += nq;
        b_i = (b_i + 1 == DEPTH) ? 0 : b_i + 1;
    };
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d) issue_next();
    __syncwarp();
    int i_c = 0, q_c = 0, nb_c = n_my > 0 ? st.nbq[0] : 0;      // arithmetic cursor
    int j = n_my > 0 ? __ldg(rows + (int64_t)gw * SELL_C + rl) : -1;
    int buf = 0;
    unsigned phases = 0u;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    while (i_c < n_my) {
        issue_next();                            // refills the buffer the previous iteration released
        mbar_wait(&st.bar[buf], (phases >> buf) & 1u);
        phases ^= 1u << buf;
        const int nq = min(Q, nb_c - q_c);
        cx wr[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q)
            if (q < nq) {
                const cx* px = x + (int64_t)st.col[buf][q * SELL_C + rl] * 4 + u;
                if constexpr ((HINT & 2) != 0) wr[q] = ldx_hint(px, polx);
                else wr[q] = ldx(px);
            }
#pragma unroll
        for (int q = 0; q < Q; ++q)
            if (q < nq) {
                const float4 e = st.val[buf][(q * SELL_C + rl) * 2 + h];
                const cx w = wr[q];
                a0 += (double)e.x * w.re - (double)e.y * w.im;
                a1 += (double)e.x * w.im + (double)e.y * w.re;
                a2 += (double)e.z * w.re - (double)e.w * w.im;
                a3 += (double)e.z * w.im + (double)e.w * w.re;
            }
        __syncwarp();                            // every lane is done with this buffer before it is refilled
        q_c += nq;
        buf = (buf + 1 == DEPTH) ? 0 : buf + 1;
        if (q_c >= nb_c) {                       // slice finished: reduce over the column halves, write the two rows
            a0 += __shfl_xor_sync(FULL, a0, 2);
            a1 += __shfl_xor_sync(FULL, a1, 2);
            a2 += __shfl_xor_sync(FULL, a2, 2);
            a3 += __shfl_xor_sync(FULL, a3, 2);
            if (j >= 0 && h == 0) {
                if constexpr ((HINT & 4) != 0) {
                    st_stream(y + (int64_t)(2 * j) * 2 + u, cx{a0, a1});
                    st_stream(y + (int64_t)(2 * j + 1) * 2 + u, cx{a2, a3});
                } else {
                    stv(y, (int64_t)(2 * j) * 2 + u, cx{a0, a1});
                    stv(y, (int64_t)(2 * j + 1) * 2 + u, cx{a2, a3});
                }
            }
            a0 = a1 = a2 = a3 = 0.0;
            ++i_c;
            q_c = 0;
            if (i_c < n_my) {                    // the copy cursor is at least one chunk ahead: it has entered this slice
                nb_c = st.nbq[i_c & 3];
                j = __ldg(rows + ((int64_t)gw + (int64_t)i_c * G) * SELL_C + rl);
            }
        }
    }
}

template <int Q, int WARPS, int DEPTH, int HINT>
static int bsell_tma_launch(emb_ctx* c, const cf* val, const cx* x, cx* y, int nsm) {
    const size_t smem = (size_t)WARPS * sizeof(SellStageQ<Q, DEPTH>);
    auto kern = k_bsell_tma<Q, WARPS, DEPTH, HINT>;
    EMB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<nsm, WARPS * 32, smem, c->stream>>>((int)c->sell_nslices, c->sell_rows.p, c->sell_sptr.p, c->sell_bcol.p,
                                               reinterpret_cast<const float4*>(val), x, y);
    EMB_LAUNCH_CHECK(c);
    return EMB_OK;
}
// EMB_SELL_VARIANT = "Q:WARPS:DEPTH:HINT": the alternatives kept for tools/sell_variants.py (tuning probe; read per launch)
static int bsell_tma_dispatch(emb_ctx* c, const char* v, const cf* val, const cx* x, cx* y, int nsm) {
    int Q = 0, W = 0, D = 0, H = 0;
    if (sscanf(v, "%d:%d:%d:%d", &Q, &W, &D, &H) == 4) {
#define EMB_V(q, w, d, hh) if (Q == q && W == w && D == d && H == hh) return bsell_tma_launch<q, w, d, hh>(c, val, x, y, nsm)
        EMB_V(8, 24, 2, 1); EMB_V(8, 24, 2, 0); EMB_V(8, 24, 3, 1); EMB_V(8, 24, 2, 5); EMB_V(8, 24, 2, 3);
        EMB_V(6, 28, 2, 1); EMB_V(5, 32, 2, 1); EMB_V(10, 20, 2, 1); EMB_V(16, 16, 2, 1); EMB_V(16, 16, 2, 0);
#undef EMB_V
    }
    c->err = std::string("EMB_SELL_VARIANT: no such instantiation: ") + v;
    return EMB_ERR_ARG;
}

template <int NV>
static int bsell_launch(emb_ctx* c, const cf* val, const cx* x, cx* y) {
    static const int mode = getenv("EMB_SPMV_TMA") ? atoi(getenv("EMB_SPMV_TMA")) : 1;
    if constexpr (NV == 2) {               // NV = 4 in two passes over the staged chunk was measured slower than k_bsell
        if (mode && !c->sell_has_empty) {      // (a slice without any block is never visited by the staged kernel)
            int nsm = 0;
            EMB_CUDA(c, cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device));
            if (const char* v = getenv("EMB_SELL_VARIANT")) return bsell_tma_dispatch(c, v, val, x, y, nsm);
            return bsell_tma_launch<STG_Q, STG_WARPS, STG_DEPTH, STG_HINT>(c, val, x, y, nsm);
        }
    }
    const int64_t nR = (int64_t)c->sell_nslices * SELL_C;
    k_bsell<NV><<<blocks_for(nR * 2 * NV, 256), 256, 0, c->stream>>>(nR, c->sell_rows.p, c->sell_sptr.p, c->sell_bcol.p,
                                                                    reinterpret_cast<const float4*>(val), x, y);
    EMB_LAUNCH_CHECK(c);
    return EMB_OK;
}
