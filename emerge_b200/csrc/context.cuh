// Device-side state of one emerge_b200 context (one GPU, one stream, one owner thread).
#pragma once
#include "emb_common.cuh"
#include "emerge_b200.h"
#include <complex>
#include <map>
#include <vector>

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

struct Surface {
    bool defined = false;
    int frame = 0;
    int64_t ntri = 0;
    DevBuf<int> tri;          // global triangle ids
    DevBuf<double> xy;        // local 2-D vertex coordinates [ntri][6] (x0,x1,x2,y0,y1,y2)
    DevBuf<double> S;         // gamma-free blocks [ntri][64]
    // matrix scatter lists on the FULL pattern (deterministic: entries of one slot summed in list order)
    int64_t nslot = 0;
    DevBuf<int64_t> slot;     // unique full-pattern slots, ascending
    DevBuf<int> segptr;       // [nslot+1] into ent
    DevBuf<int> ent;          // entry ids (tri*64 + i*8 + j)
    DevBuf<double> Sval;      // summed value per unique slot
    DevBuf<int64_t> slot_s;   // same slots mapped to the solve-space pattern (-1 if eliminated)
    DevBuf<int64_t> mv_slot;  // slot_s sorted ascending with the matching values (matrix-vector product S x, recycle.cuh)
    DevBuf<double> mv_val;
    // forcing vector
    int64_t ndof = 0;
    DevBuf<int> dof;          // unique dofs, ascending
    DevBuf<int> dsegptr;      // [ndof+1]
    DevBuf<int> dent;         // entry ids (tri*8 + i)
    DevBuf<cx> bloc;          // per-triangle forcing [ntri][8]
    DevBuf<cx> bval;          // summed forcing per unique dof
    bool has_rhs = false;
};

// One level of a smoothed-aggregation hierarchy (set up on the host, emerge_b200/amg.py): real SPD matrix A, prolongator
// P to the next coarser level and its transpose, damped-Jacobi data, complex work vectors of the V-cycle.
struct AmgLevel {
    int64_t n = 0, nc = 0;
    DevBuf<int64_t> aptr, pptr, tptr;
    DevBuf<int> acol, pcol, tcol;
    DevBuf<double> aval, pval, tval, dinv;
    double omega = 1.0;
    int lpr_a = 4, lpr_p = 4, lpr_t = 4;      // lanes per row of the A, P and P^T kernels (from the average row length)
};
struct AmgHierarchy {
    std::vector<AmgLevel> lev;
    DevBuf<double> cinv;            // dense inverse of the coarsest matrix, row-major [n][n]
    int64_t ncinv = 0;
};
// level vectors of one user of a hierarchy (NVMAX interleaved columns each)
struct AmgWork {
    std::vector<DevBuf<cx>> b, xa, xb, t;
};

// auxiliary space of the additive multilevel preconditioner: real transfer matrix R (rows of the parent space x ncol)
// and R^T, both CSR.  parent < 0: the parent is the solve space.  solver 0: diagonal of R^T A R (top-level spaces only,
// recomputed per frequency); solver 1: V-cycle of hierarchy `hid` times scale (scale_mode 0: 1, 1: -1/k0^2).
struct AuxSpace {
    int64_t ncol = 0, nnz = 0, nrow = 0;
    int parent = -1, solver = 0, hid = -1, scale_mode = 0;
    bool has_children = false;
    DevBuf<int64_t> rptr, tptr;     // R rows [nrow+1], R^T rows [ncol+1]
    DevBuf<int> rcol, tcol;
    DevBuf<double> rval, tval;
    DevBuf<cx> dinv, tmp, traw;     // 1/diag(R^T A R) [ncol]; correction x, raw restricted residual [ncol][NVMAX]
    AmgWork wk;                     // V-cycle vectors (solver 1)
};

struct emb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    std::map<std::string, double> ms;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // mesh
    int64_t nN = 0, nT = 0, nE = 0, nTri = 0, N = 0;
    DevBuf<double> nodes;     // [nN][3]
    DevBuf<int> tris;         // [nTri][3]
    DevBuf<int> tri2f;        // [8][nTri]
    DevBuf<int> tetc;         // [nT][4] vertex ids in ascending order
    DevBuf<int> tetord;       // [nT] packed original local index of k-th smallest vertex (2 bits each)
    DevBuf<int> gid;          // [nT][20] global dof ids in canonical function order
    DevBuf<cx> er, ur;        // [9][nT]
    bool have_mesh = false, have_mat = false;

    // adjacency dof -> (tet*20 + canonical local index), ascending
    DevBuf<int64_t> adjptr;   // [N+1]
    DevBuf<int> adj;          // [20 nT]
    // full pattern
    DevBuf<int64_t> rowptr;   // [N+1]
    DevBuf<int> col;          // [nnz]
    int64_t nnz = 0;
    bool have_pattern = false;
    DevBuf<cx> K, M;          // [nnz]
    bool have_KM = false;
    // chunked numeric phase (assembly.cu): (row, chunk) work items sorted by chunk, host copy of the chunk offsets
    int64_t asm_chunk_req = 0, asm_chunk = 0;
    bool asm_persist = true;
    DevBuf<unsigned long long> asm_items;
    std::vector<int64_t> asm_chunk_ptr;
    // fused numeric phase (assembly.cu, k_asm_rows): entities (edge / face = the pair of rows e, e + nE + nTri) grouped
    // by row-length class and, inside a class, ordered by their first tetrahedron (L2 locality of the per-tet records)
    int asm_mode = 0;                   // 0: fused (default), 1: element kernel -> COO scratch -> row reduction
    bool asm_pairs_ok = false;          // the mesh tables have the pair structure the fused kernel relies on
    DevBuf<int> asm_ent;                // [nE + nTri] entity ids in processing order
    int64_t asm_cls_ptr[5] = {0, 0, 0, 0, 0};

    // solve space
    int64_t Ns = 0, nnz_s = 0;
    DevBuf<int> newid;        // [N] full dof -> solve index or -1
    DevBuf<int> solve_ids;    // [Ns]
    DevBuf<int64_t> rowptr_s; // [Ns+1]
    DevBuf<int> col_s;        // [nnz_s]
    DevBuf<int64_t> src;      // [nnz_s] full-pattern slot of each solve-space entry
    // Pair ordering of the solve space: when both functions of every kept edge / face are kept (PEC elimination always
    // does that), solve index 2j is the first and 2j+1 the second function of kept entity j.  Rows 2j and 2j+1 then share
    // one column list made of pairs (2c, 2c+1): the operator is a block-CSR matrix of 2x2 blocks whose column index is
    // stored once per block (blkcol), and the two x entries of a block are adjacent in memory.
    bool paired = false;
    DevBuf<int> sperm;        // [Ns] ascending-dof position -> solve index (identity when not paired)
    DevBuf<int> blkcol;       // [nnz_s / 4] entity column of each 2x2 block (paired only)
    // SELL-8-sigma layout of the inner operator As (complex64), sell.cuh: slice-row -> block-row, block-row -> slice-row,
    // slice offsets (in blocks), column of every padded block slot
    bool sell_ready = false, sell_active = false, sell_tried = false, sell_has_empty = false;
    const void* sell_zeroed = nullptr;        // As32 buffer whose padding slots are known to be zero
    int64_t sell_nslices = 0, sell_blocks = 0;
    DevBuf<int> sell_rows, sell_pos, sell_bcol;
    DevBuf<int64_t> sell_sptr;
    DevBuf<cx> A;             // [nnz_s]
    bool have_dirichlet = false, have_A = false;
    double k0 = 0;

    Surface surf[16];

    // solver workspace
    DevBuf<cx> xs;            // last solution (solve space)
    DevBuf<cx> bs;            // right-hand side of the current solve (solve space)
    DevBuf<cx> xfull;         // last solution (full space)
    // scratch of the point evaluations (emb_interp*): grown on demand and kept, six calls per frequency point otherwise
    // paid three cudaMalloc / cudaFree each
    DevBuf<int> itp_tet;
    DevBuf<double> itp_xyz;
    DevBuf<cx> itp_E;
    std::vector<int> itp_host;
    void* shift_invert = nullptr;     // dense shift-invert operator of the port eigenproblem (modal.cu)
    // asynchronous field output (emb_fields_async): per-column staging copies of the full-space solutions, moved to the
    // caller's (pinned) buffers by a copy stream while the next point is being solved
    bool fields_async = false;
    cudaStream_t copy_stream = nullptr;
    DevBuf<cx> xstage[4];
    cudaEvent_t ev_stage_ready[4] = {}, ev_stage_done[4] = {};
    std::vector<DevBuf<cx>> work;
    DevBuf<cx> dinv;          // Jacobi / block-Jacobi inverse blocks
    DevBuf<int> pairmate;     // solve-space index of the paired dof (block-Jacobi) or -1
    DevBuf<double> red;       // reduction scratch
    std::vector<AuxSpace> aux;
    std::vector<AmgHierarchy> amg;
    // inner operator of COCR: symmetric part of A(f), stored complex64 (default) or complex128; cached between the
    // solves of one frequency
    DevBuf<cx> As;
    DevBuf<float> As32;       // [nnz_s][2]
    bool as_fp32 = true;
    bool block_krylov = true;  // the ports of a lockstep group share one Krylov space (block COCR)
    bool have_As = false;
    int As_precond = -1;
    // side streams of the additive preconditioner (independent auxiliary spaces run concurrently) and their events
    static constexpr int NSIDE = 6;
    cudaStream_t side[NSIDE] = {};
    cudaEvent_t ev_fork = nullptr;
    std::vector<cudaEvent_t> ev_restr, ev_done;
    bool use_side_streams = true;
    // Reduced-basis recycling across the frequency points of a sweep (recycle.cuh).  U: solution directions of earlier
    // solves (any port, any frequency).  A(f) = sum_t coef_t(f) W_t is affine in T = 2 + nsurf fixed matrices
    // (K, M, S_p), so W_t u is computed ONCE per direction and kept in an orthonormal basis Q with small host-side
    // coefficient matrices R_t:  W_t U = Q R_t,  A(f) U = Q G(f),  G(f) = sum_t coef_t(f) R_t.
    int rc_cap = 0, rc_n = 0;           // directions held: slots 0..rc_n-1, oldest first
    int rc_nq = 0, rc_qcap = 0;         // columns of Q in use / allocated
    DevBuf<cx> rcU, rcQ;                // [rc_cap][Ns], [rc_qcap][Ns]
    std::vector<int> rc_terms;          // surface ids of the affine terms 2.. (terms 0, 1 are K and M)
    std::vector<std::vector<std::complex<double>>> rc_R;   // [T] column-major rc_qcap x rc_cap
    std::vector<double> rc_uscale;      // x0 = sum_j y_j * uscale_j * U_j
    std::vector<std::complex<double>> aff_coef;            // coefficients of the last emb_form_A: 1, -k0^2, gamma_p...
    std::vector<int> aff_sids;
    double rc_snap = 0.3;               // solves that feed the recycled space run to rc_snap * rtol
    DevBuf<cx> rc_x0;                   // start vectors of the current solve (new direction = x - x0)
    DevBuf<cx> rc_part;                 // dot partials + coefficients of the batched Gram-Schmidt / projection kernels
    DevBuf<cx> rc_tmp;                  // one contiguous vector
    int64_t rc_spmvs = 0;               // SpMVs spent on W_t u so far (reported by emb_recycle_info)
    int64_t rc_rebuilds = 0;
    int64_t rc_accepted_total = 0;      // directions accepted since the context was created (monotonic: survives compaction / reset)
    double rc_last_proj_relres = -1;    // relative residual left by the projection in the last solve
    int nsol = 0;                       // columns of the last lockstep solve held in xs
    // EXPERIMENTAL (off by default, not yet measured on the GPU): the reduced basis as an extra coarse space of the
    // preconditioner, M^-1 += U Ceff U^T with Ceff = T (T^T U^T As U T)^-1 T^T and T a truncated orthonormalisation of U
    // (recycle.cuh::rc_coarse_update; CPU prototype tools/proto_coarse_basis.py: 326 -> 211 iterations with 4 vectors)
    bool coarse_basis = false;
    std::vector<std::complex<double>> rc_UtQ;      // [rc_cap][rc_qcap]  u_j^T q_i (unconjugated)
    std::vector<std::complex<double>> rc_UhU;      // [rc_cap][rc_cap]   u_i^H u_j
    DevBuf<cx> rc_ceff, rc_ct;                     // [rc_cap][rc_cap] coefficient map; [2][rc_cap][NVMAX] small vectors
    cudaStream_t coarse_stream = nullptr;          // the projection U^T r of the coarse-space correction overlaps the multilevel cycle
    cudaEvent_t ev_coarse_fork = nullptr, ev_coarse_done = nullptr;
    DevBuf<float> rcU32;                           // [rc_cap][Ns][2] complex64 copy of U: what the coarse-space correction streams
    int coarse_m = 0;                              // directions in the coarse space of the current operator (0 = none)
    int rc_version = 0, coarse_version = -1;       // basis change counter / the one the coefficient map was built for
    double coarse_k0 = -1;
    double spmv_ms_sum = 0;   // sampled SpMV timings inside solves (CUDA events)
    int64_t spmv_ms_cnt = 0;
    double prec_ms_sum = 0;   // sampled preconditioner applications
    int64_t prec_ms_cnt = 0;
    cudaEvent_t evs0 = nullptr, evs1 = nullptr, evt0 = nullptr, evt1 = nullptr, evr0 = nullptr, evr1 = nullptr;
    cudaEvent_t evp0 = nullptr, evp1 = nullptr;
    int64_t graph_launches = 0;
};

void topology_release(emb_ctx* c);      // topology.cu: device tables of emb_topology_build not yet fetched

template <typename T>
static int dev_alloc(emb_ctx* c, DevBuf<T>& b, size_t n) {
    if (b.n == n && b.p) return EMB_OK;
    b.release();
    if (n == 0) return EMB_OK;
    cudaError_t e = cudaMalloc((void**)&b.p, n * sizeof(T));
    if (e != cudaSuccess) {
        c->err = std::string("cudaMalloc(") + std::to_string(n * sizeof(T)) + " B): " + cudaGetErrorString(e);
        b.p = nullptr;
        return EMB_ERR_CUDA;
    }
    b.n = n;
    return EMB_OK;
}

template <typename T>
static int h2d(emb_ctx* c, DevBuf<T>& b, const T* h, size_t n) {
    EMB_TRY(dev_alloc(c, b, n));
    if (n) EMB_CUDA(c, cudaMemcpyAsync(b.p, h, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    return EMB_OK;
}

struct PhaseTimer {
    emb_ctx* c;
    const char* name;
    PhaseTimer(emb_ctx* c_, const char* n) : c(c_), name(n) { cudaEventRecord(c->ev0, c->stream); }
    ~PhaseTimer() {
        cudaEventRecord(c->ev1, c->stream);
        cudaEventSynchronize(c->ev1);
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        c->ms[name] = ms;
    }
};

#define EMB_LAUNCH_CHECK(ctx)                                                        \
    do {                                                                             \
        (ctx)->launches++;                                                           \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            (ctx)->err = std::string("kernel launch: ") + cudaGetErrorString(e__) +  \
                         " at " + __FILE__ + ":" + std::to_string(__LINE__);         \
            return EMB_ERR_CUDA;                                                     \
        }                                                                            \
    } while (0)

static inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }
