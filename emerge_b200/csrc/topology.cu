// Mesh topology tables on the device: what Mesh3D.update() (reference fem/mesh3d.py:224-355) and Nedelec2.__init__
// (fem/elements/nedelec2.py:32-62) build with Python sets, dicts and per-tetrahedron loops (21 s at 0.5 M tets).
//
// Input: vertex coordinates and the tetrahedra.  Output, in the reference's layouts:
//   edges (2,nE), tris (3,nTri)  ascending vertex ids per column
//   tet_to_edge (6,nT)  local edges (1-2,1-3,1-4,2-3,4-2,3-4)      tet_to_tri (4,nT)  local faces (1-2-3,1-3-4,1-4-2,2-3-4)
//   tri_to_edge (3,nTri) edges (1-2,2-3,1-3)                       tri_to_tet (2,nTri) ascending tet ids, -1 padded
//   edge_lengths (nE)    tet_to_field (20,nT)  tri_to_field (8,nTri)  edge_to_field (2,nE)
// Numbering: the reference numbers edges / triangles in CPython-set iteration order (mesh3d.py:252-271), which only
// that interpreter can produce.  When the caller passes the reference's `edges` / `tris` they are kept verbatim and only
// the lookups are built (so every downstream dof number is the reference's); otherwise the numbering is lexicographic
// in the sorted vertex tuple (what emerge_b200.synthmesh.mesh_tables does on the host with np.unique).
// Method: 64-bit keys of the sorted vertex tuples, one stable radix sort per entity kind, head flags + scan for the
// unique ids, binary search for the triangle edges.  Everything is integer / byte work bound by HBM; ~2 ms at 1M tets.
#include "context.cuh"
#include <cub/cub.cuh>

typedef unsigned long long u64;

__device__ __forceinline__ void sort2(long long& a, long long& b) { if (b < a) { long long t = a; a = b; b = t; } }

// keys of the 6 local edges and 4 local faces of every tetrahedron
__global__ void k_topo_keys(int64_t nT, int64_t nN, const int64_t* __restrict__ tets, u64* __restrict__ ekey, int* __restrict__ epos,
                            u64* __restrict__ fkey, int* __restrict__ fpos) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nT) return;
    long long v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = tets[t * 4 + k];
    const int le[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {3, 1}, {2, 3}};
    const int lf[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 1}, {1, 2, 3}};
#pragma unroll
    for (int e = 0; e < 6; ++e) {
        long long a = v[le[e][0]], b = v[le[e][1]];
        sort2(a, b);
        ekey[t * 6 + e] = (u64)a * (u64)nN + (u64)b;
        epos[t * 6 + e] = (int)(t * 6 + e);
    }
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        long long a = v[lf[f][0]], b = v[lf[f][1]], c = v[lf[f][2]];
        sort2(a, b); sort2(b, c); sort2(a, b);
        fkey[t * 4 + f] = ((u64)a * (u64)nN + (u64)b) * (u64)nN + (u64)c;
        fpos[t * 4 + f] = (int)(t * 4 + f);
    }
}
__global__ void k_topo_heads(int64_t n, const u64* __restrict__ skey, int* __restrict__ head) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || skey[i] != skey[i - 1]) ? 1 : 0;
}
// ukey[uid] = key of unique entity uid (sorted order); uid_of_pos[original position] = uid
__global__ void k_topo_unique(int64_t n, const u64* __restrict__ skey, const int* __restrict__ spos, const int* __restrict__ head,
                              const int* __restrict__ scan, u64* __restrict__ ukey, int* __restrict__ uid_of_pos) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int uid = scan[i] + head[i] - 1;          // inclusive count of heads - 1
    if (head[i]) ukey[uid] = skey[i];
    uid_of_pos[spos[i]] = uid;
}
__global__ void k_topo_given_keys(int64_t n, int64_t nN, int arity, const int64_t* __restrict__ ent, u64* __restrict__ key,
                                  int* __restrict__ idx) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 k = (u64)ent[i] * (u64)nN + (u64)ent[n + i];
    if (arity == 3) k = k * (u64)nN + (u64)ent[2 * n + i];
    key[i] = k;
    idx[i] = (int)i;
}
// perm[uid] = caller's index of unique entity uid; the sorted given keys must equal the mesh's unique keys
__global__ void k_topo_match(int64_t n, const u64* __restrict__ ukey, const u64* __restrict__ gkey_sorted,
                             const int* __restrict__ gidx_sorted, int* __restrict__ perm, int* __restrict__ bad) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (ukey[i] != gkey_sorted[i]) atomicExch(bad, 1);
    perm[i] = gidx_sorted[i];
}
__global__ void k_topo_decode(int64_t n, int64_t nN, int arity, const u64* __restrict__ ukey, int64_t* __restrict__ ent) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 k = ukey[i];
    if (arity == 3) { ent[2 * n + i] = (int64_t)(k % (u64)nN); k /= (u64)nN; }
    ent[n + i] = (int64_t)(k % (u64)nN);
    ent[i] = (int64_t)(k / (u64)nN);
}
// out[l][t] = perm[uid_of_pos[t * L + l]]  (+ optional identity perm)
__global__ void k_topo_tet_table(int64_t nT, int L, const int* __restrict__ uid_of_pos, const int* __restrict__ perm,
                                 int64_t* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nT * L) return;
    const int64_t t = i / L;
    const int l = (int)(i % L);
    const int uid = uid_of_pos[i];
    out[(int64_t)l * nT + t] = perm ? perm[uid] : uid;
}
__device__ __forceinline__ int64_t topo_find(const u64* __restrict__ ukey, int64_t n, u64 k) {
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (ukey[mid] < k) lo = mid + 1; else hi = mid;
    }
    return ukey[lo] == k ? lo : -1;
}
__global__ void k_topo_tri_edges(int64_t nTri, int64_t nE, int64_t nN, const int64_t* __restrict__ tris, const u64* __restrict__ uek,
                                 const int* __restrict__ eperm, int64_t* __restrict__ tri_to_edge, int* __restrict__ bad) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nTri) return;
    const u64 a = (u64)tris[i], b = (u64)tris[nTri + i], c = (u64)tris[2 * nTri + i];
    const u64 keys[3] = {a * (u64)nN + b, b * (u64)nN + c, a * (u64)nN + c};      // (1-2, 2-3, 1-3): mesh3d.py:330-335
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int64_t u = topo_find(uek, nE, keys[k]);
        if (u < 0) { atomicExch(bad, 1); tri_to_edge[k * nTri + i] = -1; continue; }
        tri_to_edge[k * nTri + i] = eperm ? eperm[u] : u;
    }
}
// sorted face keys: the (up to two) tets of every unique triangle in ascending tet order (stable sort keeps positions ascending)
__global__ void k_topo_tri_tets(int64_t n, int64_t nTri, const int* __restrict__ spos, const int* __restrict__ head,
                                const int* __restrict__ scan, const int* __restrict__ fperm, int64_t* __restrict__ tri_to_tet,
                                int* __restrict__ bad) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int uid = scan[i] + head[i] - 1;
    const int64_t tri = fperm ? fperm[uid] : uid;
    const int64_t tet = spos[i] / 4;
    if (head[i]) {
        tri_to_tet[tri] = tet;
        if (i + 1 >= n || head[i + 1]) tri_to_tet[nTri + tri] = -1;
    } else {
        if (!head[i - 1]) atomicExch(bad, 1);          // a face shared by more than two tetrahedra
        tri_to_tet[nTri + tri] = tet;
    }
}
__global__ void k_topo_lengths(int64_t nE, const int64_t* __restrict__ edges, const double* __restrict__ nodes, double* __restrict__ len) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nE) return;
    const double* a = nodes + edges[i] * 3;
    const double* b = nodes + edges[nE + i] * 3;
    const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    // mesh3d.py:352 with numpy's evaluation order and no FMA contraction: bit-identical to the reference's edge_lengths
    len[i] = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}
// dof tables of Nedelec2 (nedelec2.py:46-62)
__global__ void k_topo_fields(int64_t nT, int64_t nE, int64_t nTri, const int64_t* __restrict__ t2e, const int64_t* __restrict__ t2t,
                              const int64_t* __restrict__ tri2e, int64_t* __restrict__ ttf, int64_t* __restrict__ trf,
                              int64_t* __restrict__ etf) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < nT) {
        for (int e = 0; e < 6; ++e) {
            const int64_t v = t2e[e * nT + i];
            ttf[e * nT + i] = v;
            ttf[(10 + e) * nT + i] = v + nTri + nE;
        }
        for (int f = 0; f < 4; ++f) {
            const int64_t v = t2t[f * nT + i];
            ttf[(6 + f) * nT + i] = v + nE;
            ttf[(16 + f) * nT + i] = v + nTri + 2 * nE;
        }
    }
    if (i < nTri) {
        for (int e = 0; e < 3; ++e) {
            const int64_t v = tri2e[e * nTri + i];
            trf[e * nTri + i] = v;
            trf[(4 + e) * nTri + i] = v + nE + nTri;
        }
        trf[3 * nTri + i] = i + nE;
        trf[7 * nTri + i] = i + 2 * nE + nTri;
    }
    if (i < nE) {
        etf[i] = i;
        etf[nE + i] = i + nTri + nE;
    }
}

struct TopoState {
    int64_t nN = 0, nT = 0, nE = 0, nTri = 0;
    DevBuf<int64_t> edges, tris, t2e, t2t, tri2e, tri2t, ttf, trf, etf;
    DevBuf<double> len;
    bool ready = false;
    void release() {
        edges.release(); tris.release(); t2e.release(); t2t.release(); tri2e.release(); tri2t.release(); ttf.release();
        trf.release(); etf.release(); len.release();
        ready = false;
    }
};
static std::map<emb_ctx*, TopoState>& topo_states() {
    static std::map<emb_ctx*, TopoState> m;
    return m;
}

void topology_release(emb_ctx* c) {
    auto it = topo_states().find(c);
    if (it == topo_states().end()) return;
    it->second.release();
    topo_states().erase(it);
}

// unique entities of one kind: sorts (key, pos), returns sorted keys / positions, head flags, exclusive scan, count
static int topo_unique(emb_ctx* c, int64_t n, DevBuf<u64>& key, DevBuf<int>& pos, DevBuf<u64>& skey, DevBuf<int>& spos,
                       DevBuf<int>& head, DevBuf<int>& scan, DevBuf<char>& tmp, int key_bits, int64_t* count) {
    EMB_TRY(dev_alloc(c, skey, (size_t)n));
    EMB_TRY(dev_alloc(c, spos, (size_t)n));
    EMB_TRY(dev_alloc(c, head, (size_t)n));
    EMB_TRY(dev_alloc(c, scan, (size_t)n));
    size_t tb = 0;
    EMB_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, skey.p, pos.p, spos.p, (int)n, 0, key_bits, c->stream));
    if (tb > tmp.n) EMB_TRY(dev_alloc(c, tmp, tb));
    EMB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.p, skey.p, pos.p, spos.p, (int)n, 0, key_bits, c->stream));
    k_topo_heads<<<blocks_for(n, 256), 256, 0, c->stream>>>(n, skey.p, head.p);
    EMB_LAUNCH_CHECK(c);
    size_t tb2 = 0;
    EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tb2, head.p, scan.p, (int)n, c->stream));
    if (tb2 > tmp.n) EMB_TRY(dev_alloc(c, tmp, tb2));
    EMB_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, tb2, head.p, scan.p, (int)n, c->stream));
    c->launches += 4;
    int last_scan = 0, last_head = 0;
    EMB_CUDA(c, cudaMemcpyAsync(&last_scan, scan.p + n - 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaMemcpyAsync(&last_head, head.p + n - 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    *count = (int64_t)last_scan + last_head;
    return EMB_OK;
}

static int bits_for(double maxkey) {
    int b = 1;
    while (b < 64 && ldexp(1.0, b) <= maxkey) ++b;
    return b;
}

// given_edges / given_tris: the caller's numbering ((2,nE) / (3,nTri) int64, C order) or NULL for lexicographic numbering
extern "C" int emb_topology_build(emb_ctx* c, int64_t nN, int64_t nT, const double* nodes_n3, const int64_t* tets_n4,
                                  int64_t nE_given, const int64_t* given_edges, int64_t nTri_given, const int64_t* given_tris,
                                  int64_t* nE_out, int64_t* nTri_out) {
    if (!c || !nodes_n3 || !tets_n4 || nN <= 0 || nT <= 0 || !nE_out || !nTri_out) {
        if (c) c->err = "emb_topology_build: null/empty argument";
        return EMB_ERR_ARG;
    }
    if ((double)nN * (double)nN * (double)nN >= 1.8e19 || 6 * nT >= ((int64_t)1 << 31)) {
        c->err = "emb_topology_build: mesh exceeds the 64-bit face key / int32 position range";
        return EMB_ERR_LIMIT;
    }
    PhaseTimer pt(c, "topology");
    TopoState& S = topo_states()[c];
    S.release();
    S.nN = nN; S.nT = nT;
    DevBuf<double> dnodes;
    DevBuf<int64_t> dtets;
    EMB_TRY(h2d(c, dnodes, nodes_n3, (size_t)nN * 3));
    EMB_TRY(h2d(c, dtets, tets_n4, (size_t)nT * 4));
    DevBuf<u64> ekey, fkey, sek, sfk, uek, ufk;
    DevBuf<int> epos, fpos, sep, sfp, ehead, fhead, escan, fscan, euid, fuid, eperm, fperm, bad;
    DevBuf<char> tmp;
    EMB_TRY(dev_alloc(c, ekey, (size_t)nT * 6));
    EMB_TRY(dev_alloc(c, epos, (size_t)nT * 6));
    EMB_TRY(dev_alloc(c, fkey, (size_t)nT * 4));
    EMB_TRY(dev_alloc(c, fpos, (size_t)nT * 4));
    EMB_TRY(dev_alloc(c, bad, 1));
    EMB_CUDA(c, cudaMemsetAsync(bad.p, 0, sizeof(int), c->stream));
    k_topo_keys<<<blocks_for(nT, 256), 256, 0, c->stream>>>(nT, nN, dtets.p, ekey.p, epos.p, fkey.p, fpos.p);
    EMB_LAUNCH_CHECK(c);
    const int ebits = bits_for((double)nN * (double)nN), fbits = bits_for((double)nN * (double)nN * (double)nN);
    int64_t nE = 0, nTri = 0;
    EMB_TRY(topo_unique(c, nT * 6, ekey, epos, sek, sep, ehead, escan, tmp, ebits, &nE));
    EMB_TRY(topo_unique(c, nT * 4, fkey, fpos, sfk, sfp, fhead, fscan, tmp, fbits, &nTri));
    if ((given_edges && nE_given != nE) || (given_tris && nTri_given != nTri)) {
        c->err = "emb_topology_build: the given edges / tris do not match the mesh (count)";
        return EMB_ERR_ARG;
    }
    S.nE = nE; S.nTri = nTri;
    EMB_TRY(dev_alloc(c, uek, (size_t)nE));
    EMB_TRY(dev_alloc(c, ufk, (size_t)nTri));
    EMB_TRY(dev_alloc(c, euid, (size_t)nT * 6));
    EMB_TRY(dev_alloc(c, fuid, (size_t)nT * 4));
    k_topo_unique<<<blocks_for(nT * 6, 256), 256, 0, c->stream>>>(nT * 6, sek.p, sep.p, ehead.p, escan.p, uek.p, euid.p);
    EMB_LAUNCH_CHECK(c);
    k_topo_unique<<<blocks_for(nT * 4, 256), 256, 0, c->stream>>>(nT * 4, sfk.p, sfp.p, fhead.p, fscan.p, ufk.p, fuid.p);
    EMB_LAUNCH_CHECK(c);
    EMB_TRY(dev_alloc(c, S.edges, (size_t)nE * 2));
    EMB_TRY(dev_alloc(c, S.tris, (size_t)nTri * 3));
    // numbering: the caller's, or lexicographic
    auto adopt = [&](int64_t n, int arity, const int64_t* given, DevBuf<int64_t>& ent, DevBuf<u64>& ukey, DevBuf<int>& perm, int kbits) -> int {
        if (!given) {
            k_topo_decode<<<blocks_for(n, 256), 256, 0, c->stream>>>(n, nN, arity, ukey.p, ent.p);
            EMB_LAUNCH_CHECK(c);
            return EMB_OK;
        }
        EMB_CUDA(c, cudaMemcpyAsync(ent.p, given, (size_t)n * arity * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        DevBuf<u64> gk, gks;
        DevBuf<int> gi, gis;
        EMB_TRY(dev_alloc(c, gk, (size_t)n));
        EMB_TRY(dev_alloc(c, gks, (size_t)n));
        EMB_TRY(dev_alloc(c, gi, (size_t)n));
        EMB_TRY(dev_alloc(c, gis, (size_t)n));
        EMB_TRY(dev_alloc(c, perm, (size_t)n));
        k_topo_given_keys<<<blocks_for(n, 256), 256, 0, c->stream>>>(n, nN, arity, ent.p, gk.p, gi.p);
        EMB_LAUNCH_CHECK(c);
        size_t tb = 0;
        EMB_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tb, gk.p, gks.p, gi.p, gis.p, (int)n, 0, kbits, c->stream));
        if (tb > tmp.n) EMB_TRY(dev_alloc(c, tmp, tb));
        EMB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tb, gk.p, gks.p, gi.p, gis.p, (int)n, 0, kbits, c->stream));
        c->launches += 1;
        k_topo_match<<<blocks_for(n, 256), 256, 0, c->stream>>>(n, ukey.p, gks.p, gis.p, perm.p, bad.p);
        EMB_LAUNCH_CHECK(c);
        EMB_CUDA(c, cudaStreamSynchronize(c->stream));
        gk.release(); gks.release(); gi.release(); gis.release();
        return EMB_OK;
    };
    EMB_TRY(adopt(nE, 2, given_edges, S.edges, uek, eperm, ebits));
    EMB_TRY(adopt(nTri, 3, given_tris, S.tris, ufk, fperm, fbits));
    EMB_TRY(dev_alloc(c, S.t2e, (size_t)nT * 6));
    EMB_TRY(dev_alloc(c, S.t2t, (size_t)nT * 4));
    EMB_TRY(dev_alloc(c, S.tri2e, (size_t)nTri * 3));
    EMB_TRY(dev_alloc(c, S.tri2t, (size_t)nTri * 2));
    EMB_TRY(dev_alloc(c, S.len, (size_t)nE));
    EMB_TRY(dev_alloc(c, S.ttf, (size_t)nT * 20));
    EMB_TRY(dev_alloc(c, S.trf, (size_t)nTri * 8));
    EMB_TRY(dev_alloc(c, S.etf, (size_t)nE * 2));
    k_topo_tet_table<<<blocks_for(nT * 6, 256), 256, 0, c->stream>>>(nT, 6, euid.p, eperm.p, S.t2e.p);
    EMB_LAUNCH_CHECK(c);
    k_topo_tet_table<<<blocks_for(nT * 4, 256), 256, 0, c->stream>>>(nT, 4, fuid.p, fperm.p, S.t2t.p);
    EMB_LAUNCH_CHECK(c);
    k_topo_tri_edges<<<blocks_for(nTri, 256), 256, 0, c->stream>>>(nTri, nE, nN, S.tris.p, uek.p, eperm.p, S.tri2e.p, bad.p);
    EMB_LAUNCH_CHECK(c);
    k_topo_tri_tets<<<blocks_for(nT * 4, 256), 256, 0, c->stream>>>(nT * 4, nTri, sfp.p, fhead.p, fscan.p, fperm.p, S.tri2t.p, bad.p);
    EMB_LAUNCH_CHECK(c);
    k_topo_lengths<<<blocks_for(nE, 256), 256, 0, c->stream>>>(nE, S.edges.p, dnodes.p, S.len.p);
    EMB_LAUNCH_CHECK(c);
    const int64_t nmax = nT > nTri ? (nT > nE ? nT : nE) : (nTri > nE ? nTri : nE);
    k_topo_fields<<<blocks_for(nmax, 256), 256, 0, c->stream>>>(nT, nE, nTri, S.t2e.p, S.t2t.p, S.tri2e.p, S.ttf.p, S.trf.p, S.etf.p);
    EMB_LAUNCH_CHECK(c);
    int hbad = 0;
    EMB_CUDA(c, cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    dnodes.release(); dtets.release(); ekey.release(); fkey.release(); sek.release(); sfk.release(); uek.release(); ufk.release();
    epos.release(); fpos.release(); sep.release(); sfp.release(); ehead.release(); fhead.release(); escan.release(); fscan.release();
    euid.release(); fuid.release(); eperm.release(); fperm.release(); bad.release(); tmp.release();
    if (hbad) {
        S.release();
        c->err = "emb_topology_build: the given edges / tris do not match the mesh, or a face belongs to more than two tetrahedra";
        return EMB_ERR_ARG;
    }
    S.ready = true;
    *nE_out = nE;
    *nTri_out = nTri;
    return EMB_OK;
}

// copies the tables of the last emb_topology_build out (any pointer may be NULL) and releases the device copies
extern "C" int emb_topology_get(emb_ctx* c, int64_t* edges_2xnE, int64_t* tris_3xnTri, int64_t* tet_to_edge_6xnT,
                                int64_t* tet_to_tri_4xnT, int64_t* tri_to_edge_3xnTri, int64_t* tri_to_tet_2xnTri,
                                double* edge_lengths, int64_t* tet_to_field_20xnT, int64_t* tri_to_field_8xnTri,
                                int64_t* edge_to_field_2xnE) {
    if (!c) return EMB_ERR_ARG;
    auto it = topo_states().find(c);
    if (it == topo_states().end() || !it->second.ready) { c->err = "emb_topology_get: emb_topology_build not called"; return EMB_ERR_STATE; }
    TopoState& S = it->second;
    auto out = [&](void* dst, const void* src, size_t bytes) -> int {
        if (dst) EMB_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
        return EMB_OK;
    };
    EMB_TRY(out(edges_2xnE, S.edges.p, (size_t)S.nE * 2 * 8));
    EMB_TRY(out(tris_3xnTri, S.tris.p, (size_t)S.nTri * 3 * 8));
    EMB_TRY(out(tet_to_edge_6xnT, S.t2e.p, (size_t)S.nT * 6 * 8));
    EMB_TRY(out(tet_to_tri_4xnT, S.t2t.p, (size_t)S.nT * 4 * 8));
    EMB_TRY(out(tri_to_edge_3xnTri, S.tri2e.p, (size_t)S.nTri * 3 * 8));
    EMB_TRY(out(tri_to_tet_2xnTri, S.tri2t.p, (size_t)S.nTri * 2 * 8));
    EMB_TRY(out(edge_lengths, S.len.p, (size_t)S.nE * 8));
    EMB_TRY(out(tet_to_field_20xnT, S.ttf.p, (size_t)S.nT * 20 * 8));
    EMB_TRY(out(tri_to_field_8xnTri, S.trf.p, (size_t)S.nTri * 8 * 8));
    EMB_TRY(out(edge_to_field_2xnE, S.etf.p, (size_t)S.nE * 2 * 8));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    S.release();
    topo_states().erase(it);
    return EMB_OK;
}
