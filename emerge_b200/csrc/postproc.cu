// Field post-processing on the device: point location and E / curl E interpolation of a solution at
// arbitrary points.  Replaces EMDataSet.interpolate (reference fem/physics/edm/emdata.py:181-199) -> FEMBasis.interpolate /
// interpolate_curl (fem/elements/nedelec2.py:72-86) -> ned2_tet_interp / ned2_tet_interp_curl (fem/mth/tet.py:371-626),
// which loop over ALL tetrahedra for ALL points on the host (O(n_tets x n_points) in numba, single thread).
//
// Semantics kept: a point belongs to the LAST tetrahedron (highest index) whose test passes - the reference overwrites as
// it walks the tets (tet.py:393-497); the test is on the local coordinates w.r.t. (v2-v1, v3-v1, v4-v1) of the tet's own
// vertex order: l1+l2+l3 <= 1.00000001 and l_i >= -1e-6 (tet.py:425).  Points in no tetrahedron get zeros.
// Kernels: k_locate_prep (thread per tet: the 3x4 affine map to local coordinates, 96 B per tet), k_locate (thread per
// point, walks the tets from the last one down, every load is a warp-wide broadcast; exits when the whole warp has
// found its tets), k_interp_eh (thread per point: the 20 basis functions and their curls in the canonical-vertex form
// of ned2_tet.cuh).
#include "context.cuh"

__global__ void k_locate_prep(int64_t nT, const int* __restrict__ tetc, const int* __restrict__ tetord,
                              const double* __restrict__ nodes, double* __restrict__ maps) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nT) return;
    // original vertex order: ord[k] = original local index of the k-th smallest vertex id
    const int po = tetord[t];
    double v[4][3];
    for (int k = 0; k < 4; ++k) {
        const int o = (po >> (2 * k)) & 3;
        const double* q = nodes + (int64_t)tetc[t * 4 + k] * 3;
        v[o][0] = q[0]; v[o][1] = q[1]; v[o][2] = q[2];
    }
    double b1[3], b2[3], b3[3];
    for (int a = 0; a < 3; ++a) { b1[a] = v[1][a] - v[0][a]; b2[a] = v[2][a] - v[0][a]; b3[a] = v[3][a] - v[0][a]; }
    double r[3][3];
    r[0][0] = b2[1] * b3[2] - b2[2] * b3[1]; r[0][1] = b2[2] * b3[0] - b2[0] * b3[2]; r[0][2] = b2[0] * b3[1] - b2[1] * b3[0];
    r[1][0] = b3[1] * b1[2] - b3[2] * b1[1]; r[1][1] = b3[2] * b1[0] - b3[0] * b1[2]; r[1][2] = b3[0] * b1[1] - b3[1] * b1[0];
    r[2][0] = b1[1] * b2[2] - b1[2] * b2[1]; r[2][1] = b1[2] * b2[0] - b1[0] * b2[2]; r[2][2] = b1[0] * b2[1] - b1[1] * b2[0];
    const double det = b1[0] * r[0][0] + b1[1] * r[0][1] + b1[2] * r[0][2];
    const double id = 1.0 / det;
    double* m = maps + t * 12;
    for (int i = 0; i < 3; ++i) {
        const double mx = r[i][0] * id, my = r[i][1] * id, mz = r[i][2] * id;
        m[4 * i] = mx; m[4 * i + 1] = my; m[4 * i + 2] = mz;
        m[4 * i + 3] = -(mx * v[0][0] + my * v[0][1] + mz * v[0][2]);
    }
}

__global__ void __launch_bounds__(128) k_locate(int64_t npts, int64_t nT, const double* __restrict__ xyz,
                                                const double* __restrict__ maps, int* __restrict__ tet_of) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool live = k < npts;
    const double x = live ? xyz[k] : 0.0, y = live ? xyz[npts + k] : 0.0, z = live ? xyz[2 * npts + k] : 0.0;
    int found = live ? -1 : 0;
    for (int64_t t = nT - 1; t >= 0; --t) {
        const double2* m = reinterpret_cast<const double2*>(maps + t * 12);        // same address in every lane: broadcast
        const double2 a0 = __ldg(m), a1 = __ldg(m + 1), b0 = __ldg(m + 2), b1 = __ldg(m + 3), c0 = __ldg(m + 4), c1 = __ldg(m + 5);
        const double l1 = a0.x * x + a0.y * y + a1.x * z + a1.y;
        const double l2 = b0.x * x + b0.y * y + b1.x * z + b1.y;
        const double l3 = c0.x * x + c0.y * y + c1.x * z + c1.y;
        if (found < 0 && (l1 + l2 + l3) <= 1.00000001 && l1 >= -1e-6 && l2 >= -1e-6 && l3 >= -1e-6) found = (int)t;
        if ((t & 63) == 0 && __all_sync(0xffffffffu, found >= 0)) break;
    }
    if (live) tet_of[k] = found;
}

// E (and curl E * cc[tet]) at point k in tetrahedron tet[k] (< 0: zeros)
__global__ void k_interp_eh(int64_t npts, const int* __restrict__ tet, const double* __restrict__ xyz, const int* __restrict__ tetc,
                            const int* __restrict__ gid, const double* __restrict__ nodes, const cx* __restrict__ xfull,
                            const cx* __restrict__ cc, cx* __restrict__ E, cx* __restrict__ Hc) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= npts) return;
    const int t = tet[k];
    cx Ex = mk(0.0), Ey = mk(0.0), Ez = mk(0.0), Cx = mk(0.0), Cy = mk(0.0), Cz = mk(0.0);
    if (t >= 0) {
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
        const int eA[6] = {0, 0, 0, 1, 1, 2}, eB[6] = {1, 2, 3, 2, 3, 3};
        const int fA[4] = {0, 0, 0, 1}, fB[4] = {1, 1, 2, 2}, fE[4] = {2, 3, 3, 3};
        auto dist = [&](int a, int b) {
            const double u = p[a][0] - p[b][0], v = p[a][1] - p[b][1], w = p[a][2] - p[b][2];
            return sqrt(u * u + v * v + w * w);
        };
        // canonical functions N = s l lam_X w_PQ, w_PQ = lam_Q grad_P - lam_P grad_Q;
        // curl N = s l [grad_X x w_PQ - 2 lam_X grad_P x grad_Q]
        auto add = [&](int cidx, double s, double l, int X, int P, int Q) {
            const cx c = xfull[gid[(int64_t)t * 20 + cidx]];
            const double sl = s * l;
            double w[3], pq[3], xw[3];
            for (int a = 0; a < 3; ++a) w[a] = lam[Q] * grad[P][a] - lam[P] * grad[Q][a];
            pq[0] = grad[P][1] * grad[Q][2] - grad[P][2] * grad[Q][1];
            pq[1] = grad[P][2] * grad[Q][0] - grad[P][0] * grad[Q][2];
            pq[2] = grad[P][0] * grad[Q][1] - grad[P][1] * grad[Q][0];
            xw[0] = grad[X][1] * w[2] - grad[X][2] * w[1];
            xw[1] = grad[X][2] * w[0] - grad[X][0] * w[2];
            xw[2] = grad[X][0] * w[1] - grad[X][1] * w[0];
            const double f = sl * lam[X];
            fma_r(Ex, f * w[0], c); fma_r(Ey, f * w[1], c); fma_r(Ez, f * w[2], c);
            fma_r(Cx, sl * (xw[0] - 2.0 * lam[X] * pq[0]), c);
            fma_r(Cy, sl * (xw[1] - 2.0 * lam[X] * pq[1]), c);
            fma_r(Cz, sl * (xw[2] - 2.0 * lam[X] * pq[2]), c);
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
        if (Hc) { const cx s = cc[t]; Cx = Cx * s; Cy = Cy * s; Cz = Cz * s; }
    }
    E[k] = Ex; E[npts + k] = Ey; E[2 * npts + k] = Ez;
    if (Hc) { Hc[k] = Cx; Hc[npts + k] = Cy; Hc[2 * npts + k] = Cz; }
}

static int locate_device(emb_ctx* c, int64_t npts, const double* dxyz, int* dtet) {
    DevBuf<double> maps;
    EMB_TRY(dev_alloc(c, maps, (size_t)c->nT * 12));
    k_locate_prep<<<blocks_for(c->nT, 128), 128, 0, c->stream>>>(c->nT, c->tetc.p, c->tetord.p, c->nodes.p, maps.p);
    EMB_LAUNCH_CHECK(c);
    k_locate<<<blocks_for(npts, 128), 128, 0, c->stream>>>(npts, c->nT, dxyz, maps.p, dtet);
    EMB_LAUNCH_CHECK(c);
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    maps.release();
    return EMB_OK;
}

extern "C" int emb_locate_points(emb_ctx* c, int64_t npts, const double* xyz_3xnpts, int64_t* tet_ids) {
    if (!c || npts <= 0 || !xyz_3xnpts || !tet_ids) return EMB_ERR_ARG;
    if (!c->have_mesh) { c->err = "emb_locate_points: mesh not uploaded"; return EMB_ERR_STATE; }
    PhaseTimer pt(c, "locate");
    DevBuf<double> dp;
    DevBuf<int> dt;
    EMB_TRY(h2d(c, dp, xyz_3xnpts, (size_t)npts * 3));
    EMB_TRY(dev_alloc(c, dt, (size_t)npts));
    EMB_TRY(locate_device(c, npts, dp.p, dt.p));
    std::vector<int> ht((size_t)npts);
    EMB_CUDA(c, cudaMemcpyAsync(ht.data(), dt.p, (size_t)npts * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int64_t i = 0; i < npts; ++i) tet_ids[i] = ht[(size_t)i];
    dp.release(); dt.release();
    return EMB_OK;
}

extern "C" int emb_interp_fields(emb_ctx* c, const emb_c128* x_full, int64_t npts, const double* xyz_3xnpts,
                                 const int64_t* tet_ids, const emb_c128* curl_const_nT, emb_c128* E_3xnpts,
                                 emb_c128* H_3xnpts) {
    if (!c || npts <= 0 || !xyz_3xnpts || !E_3xnpts || (H_3xnpts && !curl_const_nT)) return EMB_ERR_ARG;
    if (!c->have_mesh) { c->err = "emb_interp_fields: mesh not uploaded"; return EMB_ERR_STATE; }
    if (!x_full && !c->xfull.p) { c->err = "emb_interp_fields: no solution on the device"; return EMB_ERR_STATE; }
    PhaseTimer pt(c, "interp_fields");
    DevBuf<cx> dx, dcc, dE, dH;
    DevBuf<double> dp;
    DevBuf<int> dt;
    const cx* xf = c->xfull.p;
    if (x_full) { EMB_TRY(h2d(c, dx, reinterpret_cast<const cx*>(x_full), (size_t)c->N)); xf = dx.p; }
    EMB_TRY(h2d(c, dp, xyz_3xnpts, (size_t)npts * 3));
    EMB_TRY(dev_alloc(c, dt, (size_t)npts));
    if (tet_ids) {
        std::vector<int> ht((size_t)npts);
        for (int64_t i = 0; i < npts; ++i) {
            if (tet_ids[i] >= c->nT) { c->err = "emb_interp_fields: tet id out of range"; return EMB_ERR_ARG; }
            ht[(size_t)i] = tet_ids[i] < 0 ? -1 : (int)tet_ids[i];
        }
        EMB_CUDA(c, cudaMemcpyAsync(dt.p, ht.data(), (size_t)npts * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    } else {
        EMB_TRY(locate_device(c, npts, dp.p, dt.p));
    }
    EMB_TRY(dev_alloc(c, dE, (size_t)npts * 3));
    if (H_3xnpts) {
        EMB_TRY(h2d(c, dcc, reinterpret_cast<const cx*>(curl_const_nT), (size_t)c->nT));
        EMB_TRY(dev_alloc(c, dH, (size_t)npts * 3));
    }
    k_interp_eh<<<blocks_for(npts, 128), 128, 0, c->stream>>>(npts, dt.p, dp.p, c->tetc.p, c->gid.p, c->nodes.p, xf, dcc.p, dE.p,
                                                            H_3xnpts ? dH.p : nullptr);
    EMB_LAUNCH_CHECK(c);
    EMB_CUDA(c, cudaMemcpyAsync(E_3xnpts, dE.p, (size_t)npts * 3 * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    if (H_3xnpts) EMB_CUDA(c, cudaMemcpyAsync(H_3xnpts, dH.p, (size_t)npts * 3 * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    dx.release(); dcc.release(); dE.release(); dH.release(); dp.release(); dt.release();
    return EMB_OK;
}

// ---- Stratton-Chu far field ---------------------------------------------------------------------------
// Replaces stratton_chu_ff (reference fem/physics/edm/sc.py:27-142): E_far(r) = Q r x SUM_j [ (n x E_j - Z0 r x (n x H_j))
// exp(j k0 r . v_j) ],  Q = -j k0 / 4 pi,  H_far = r x E_far / Z0,  r = (cos th cos ph, cos th sin ph, sin th) (the
// reference's convention, sc.py:78-80), over the surface samples j (edge midpoints v_j with area-weighted normals n_j)
// whose |E_j| exceeds LR = 1e-3 of the largest one (sc.py:57-74).  The reference works in complex64 / float32 with
// fastmath; here the inputs are rounded to float32 exactly as its .astype() calls do (sc.py:172-178) and the sums run in
// FP64, so results agree to float32 round-off.  Mapping: a CTA per (direction, source chunk): every thread forms the three
// tangential-current components of its sources times the phase factor, fixed-order tree reduction in shared memory, the
// chunk partials are summed in ascending order by k_sc_finish (bitwise reproducible).
constexpr int SC_THREADS = 256;
constexpr int SC_CHUNK = 8192;            // sources per CTA

__global__ void k_sc_prepare(int64_t n, const cx* __restrict__ E, const cx* __restrict__ H, const double* __restrict__ vis,
                             const double* __restrict__ wns, float* __restrict__ src, float* __restrict__ emag) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    float e[6], h[6], nv[3];
    for (int a = 0; a < 3; ++a) {
        e[2 * a] = (float)E[a * n + j].re; e[2 * a + 1] = (float)E[a * n + j].im;
        h[2 * a] = (float)H[a * n + j].re; h[2 * a + 1] = (float)H[a * n + j].im;
        nv[a] = (float)wns[a * n + j];
    }
    float* s = src + j * 18;            // one record per sample: E, H (re, im pairs), weighted normal, position
    for (int a = 0; a < 6; ++a) { s[a] = e[a]; s[6 + a] = h[a]; }
    s[12] = nv[0]; s[13] = nv[1]; s[14] = nv[2];
    s[15] = (float)vis[j]; s[16] = (float)vis[n + j]; s[17] = (float)vis[2 * n + j];
    const double m = sqrt((double)e[0] * e[0] + (double)e[1] * e[1] + (double)e[2] * e[2] + (double)e[3] * e[3] +
                          (double)e[4] * e[4] + (double)e[5] * e[5]);
    emag[j] = (float)m;
}
__global__ void k_sc_max(int64_t n, const float* __restrict__ emag, unsigned* __restrict__ mx) {
    __shared__ float sh[SC_THREADS];
    float m = 0.f;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, emag[j]);
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int s = SC_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] = fmaxf(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicMax(mx, __float_as_uint(sh[0]));        // non-negative floats order like their bit patterns
}
__global__ void __launch_bounds__(SC_THREADS) k_sc_partial(int64_t n, int64_t nout, const float* __restrict__ src,
                                                           const float* __restrict__ emag, const unsigned* __restrict__ mx,
                                                           const double* __restrict__ theta, const double* __restrict__ phi,
                                                           double k0, double* __restrict__ part) {
    __shared__ double sh[6][SC_THREADS];
    const int64_t d = blockIdx.x;
    const int chunk = blockIdx.y;
    const float th = (float)theta[d], ph = (float)phi[d];
    const float kf = (float)k0;
    const double rx = (double)(cosf(th) * cosf(ph)), ry = (double)(cosf(th) * sinf(ph)), rz = (double)sinf(th);
    const double kx = (double)kf * rx, ky = (double)kf * ry, kz = (double)kf * rz;
    const double Z0 = (double)376.73031366857f;
    const float level = __uint_as_float(*mx) * 0.001f;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    const int64_t j1 = min(n, (int64_t)(chunk + 1) * SC_CHUNK);
    for (int64_t j = (int64_t)chunk * SC_CHUNK + threadIdx.x; j < j1; j += SC_THREADS) {
        if (!(emag[j] > level)) continue;
        const float* s = src + j * 18;
        const double nx = s[12], ny = s[13], nz = s[14];
        double sn, cs;
        sincos(kx * (double)s[15] + ky * (double)s[16] + kz * (double)s[17], &sn, &cs);
#pragma unroll
        for (int p = 0; p < 2; ++p) {            // real and imaginary parts of the (linear) current terms
            const double Ex = s[0 + p], Ey = s[2 + p], Ez = s[4 + p], Hx = s[6 + p], Hy = s[8 + p], Hz = s[10 + p];
            const double nHx = ny * Hz - nz * Hy, nHy = nz * Hx - nx * Hz, nHz = nx * Hy - ny * Hx;
            const double nEx = ny * Ez - nz * Ey, nEy = nz * Ex - nx * Ez, nEz = nx * Ey - ny * Ex;
            const double tx = nEx - Z0 * (ry * nHz - rz * nHy);
            const double ty = nEy - Z0 * (rz * nHx - rx * nHz);
            const double tz = nEz - Z0 * (rx * nHy - ry * nHx);
            // (t_re + j t_im)(cs + j sn): p = 0 adds t_re * (cs, sn), p = 1 adds t_im * (-sn, cs)
            const double a = p == 0 ? cs : -sn, b = p == 0 ? sn : cs;
            acc[0] += tx * a; acc[1] += tx * b;
            acc[2] += ty * a; acc[3] += ty * b;
            acc[4] += tz * a; acc[5] += tz * b;
        }
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) sh[q][threadIdx.x] = acc[q];
    __syncthreads();
    for (int s2 = SC_THREADS / 2; s2 > 0; s2 >>= 1) {
        if (threadIdx.x < s2)
#pragma unroll
            for (int q = 0; q < 6; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + s2];
        __syncthreads();
    }
    if (threadIdx.x < 6) part[((int64_t)chunk * nout + d) * 6 + threadIdx.x] = sh[threadIdx.x][0];
}
__global__ void k_sc_finish(int64_t nout, int nchunk, const double* __restrict__ part, const double* __restrict__ theta,
                            const double* __restrict__ phi, double k0, cx* __restrict__ Eout, cx* __restrict__ Hout) {
    const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (d >= nout) return;
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int c = 0; c < nchunk; ++c)
        for (int q = 0; q < 6; ++q) s[q] += part[((int64_t)c * nout + d) * 6 + q];
    const float th = (float)theta[d], ph = (float)phi[d];
    const double rx = (double)(cosf(th) * cosf(ph)), ry = (double)(cosf(th) * sinf(ph)), rz = (double)sinf(th);
    const double Z0 = (double)376.73031366857f;
    const double q = -(double)(float)k0 / (double)(4.0f * 3.14159265358979323846f);   // Q = j q
    const cx ix{s[0], s[1]}, iy{s[2], s[3]}, iz{s[4], s[5]};
    auto jmul = [q](double re, double im) { return cx{-q * im, q * re}; };             // (j q)(re + j im)
    const cx ex = jmul(ry * iz.re - rz * iy.re, ry * iz.im - rz * iy.im);
    const cx ey = jmul(rz * ix.re - rx * iz.re, rz * ix.im - rx * iz.im);
    const cx ez = jmul(rx * iy.re - ry * ix.re, rx * iy.im - ry * ix.im);
    Eout[d] = ex; Eout[nout + d] = ey; Eout[2 * nout + d] = ez;
    Hout[d] = cx{(ry * ez.re - rz * ey.re) / Z0, (ry * ez.im - rz * ey.im) / Z0};
    Hout[nout + d] = cx{(rz * ex.re - rx * ez.re) / Z0, (rz * ex.im - rx * ez.im) / Z0};
    Hout[2 * nout + d] = cx{(rx * ey.re - ry * ex.re) / Z0, (rx * ey.im - ry * ex.im) / Z0};
}

extern "C" int emb_stratton_chu(emb_ctx* c, int64_t nsrc, const emb_c128* E_3xn, const emb_c128* H_3xn, const double* pos_3xn,
                                const double* wnormal_3xn, int64_t nout, const double* theta, const double* phi, double k0,
                                emb_c128* Eout_3xnout, emb_c128* Hout_3xnout) {
    if (!c || nsrc <= 0 || nout <= 0 || !E_3xn || !H_3xn || !pos_3xn || !wnormal_3xn || !theta || !phi || !Eout_3xnout ||
        !Hout_3xnout)
        return EMB_ERR_ARG;
    PhaseTimer pt(c, "stratton_chu");
    DevBuf<cx> dE, dH, dEo, dHo;
    DevBuf<double> dv, dn, dth, dph, part;
    DevBuf<float> src, emag;
    DevBuf<unsigned> mx;
    EMB_TRY(h2d(c, dE, reinterpret_cast<const cx*>(E_3xn), (size_t)nsrc * 3));
    EMB_TRY(h2d(c, dH, reinterpret_cast<const cx*>(H_3xn), (size_t)nsrc * 3));
    EMB_TRY(h2d(c, dv, pos_3xn, (size_t)nsrc * 3));
    EMB_TRY(h2d(c, dn, wnormal_3xn, (size_t)nsrc * 3));
    EMB_TRY(h2d(c, dth, theta, (size_t)nout));
    EMB_TRY(h2d(c, dph, phi, (size_t)nout));
    const int nchunk = (int)((nsrc + SC_CHUNK - 1) / SC_CHUNK);
    if (nchunk > 65535) { c->err = "emb_stratton_chu: more than 65535 x 8192 surface samples"; return EMB_ERR_LIMIT; }
    EMB_TRY(dev_alloc(c, src, (size_t)nsrc * 18));
    EMB_TRY(dev_alloc(c, emag, (size_t)nsrc));
    EMB_TRY(dev_alloc(c, mx, 1));
    EMB_TRY(dev_alloc(c, part, (size_t)nchunk * nout * 6));
    EMB_TRY(dev_alloc(c, dEo, (size_t)nout * 3));
    EMB_TRY(dev_alloc(c, dHo, (size_t)nout * 3));
    EMB_CUDA(c, cudaMemsetAsync(mx.p, 0, sizeof(unsigned), c->stream));
    k_sc_prepare<<<blocks_for(nsrc, 256), 256, 0, c->stream>>>(nsrc, dE.p, dH.p, dv.p, dn.p, src.p, emag.p);
    EMB_LAUNCH_CHECK(c);
    k_sc_max<<<(unsigned)std::min<int64_t>(blocks_for(nsrc, SC_THREADS), 1024), SC_THREADS, 0, c->stream>>>(nsrc, emag.p, mx.p);
    EMB_LAUNCH_CHECK(c);
    k_sc_partial<<<dim3((unsigned)nout, (unsigned)nchunk), SC_THREADS, 0, c->stream>>>(nsrc, nout, src.p, emag.p, mx.p, dth.p, dph.p,
                                                                                       k0, part.p);
    EMB_LAUNCH_CHECK(c);
    k_sc_finish<<<blocks_for(nout, 128), 128, 0, c->stream>>>(nout, nchunk, part.p, dth.p, dph.p, k0, dEo.p, dHo.p);
    EMB_LAUNCH_CHECK(c);
    EMB_CUDA(c, cudaMemcpyAsync(Eout_3xnout, dEo.p, (size_t)nout * 3 * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaMemcpyAsync(Hout_3xnout, dHo.p, (size_t)nout * 3 * sizeof(cx), cudaMemcpyDeviceToHost, c->stream));
    EMB_CUDA(c, cudaStreamSynchronize(c->stream));
    dE.release(); dH.release(); dv.release(); dn.release(); dth.release(); dph.release(); part.release(); src.release();
    emag.release(); mx.release(); dEo.release(); dHo.release();
    return EMB_OK;
}
