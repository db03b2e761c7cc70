/* emerge_b200 — C ABI of the B200-native EMerge frequency-domain hot path.
 *
 * Plain pointers and sizes only; every pointer is a HOST pointer unless the name starts with d_.
 * The library copies what it needs during the call and never retains host pointers.  All device
 * state lives in an opaque context bound to one GPU and one CUDA stream; a context is owned by one
 * host thread.  Return value: 0 ok, <0 error (emb_last_error gives the text), >0 non-convergence.
 *
 * Each entry point names the reference interface it replaces (paths relative to the EMerge tree).
 * Arrays use the reference's own memory layouts so the Python seam can pass the
 * numpy buffers it already holds:
 *   nodes   (nN,3) f64 row-major  = memory of mesh.nodes (3,nN) F-order view   fem/mesh3d.py:227-230
 *   tets    (nT,4) i64 row-major  = memory of mesh.tets (4,nT) view             fem/mesh3d.py:245
 *   tris    (nTri,3) i64 row-major                                              fem/mesh3d.py:271
 *   tet_to_field (20,nT) i64 C-order, tri_to_field (8,nTri) i64 C-order         fem/elements/nedelec2.py:46-62
 *   er, ur  (3,3,nT) c128 C-order                                               fem/mesh3d.py:358-378
 */
#ifndef EMERGE_B200_H
#define EMERGE_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct emb_ctx emb_ctx;
typedef struct { double re, im; } emb_c128;

/* ---- context ------------------------------------------------------------------------------ */
int emb_create(int device, emb_ctx** out);
void emb_destroy(emb_ctx* ctx);
const char* emb_last_error(const emb_ctx* ctx);
const char* emb_version(void);
/* number of kernels this library launched since the context was created (bench.py gpu_launches) */
int64_t emb_launch_count(const emb_ctx* ctx);
/* device time (ms, CUDA events on the context stream) of the last call of the named phase:
 * "symbolic", "tet_kernel", "reduce", "form_A", "spmv", "solve", "surface" */
double emb_last_ms(const emb_ctx* ctx, const char* phase);

/* CUDA-event stopwatch on the context's stream (bench.py times whole steps with it: torch.cuda.Event only sees
 * torch's own stream).  stop returns the elapsed device-timeline milliseconds since start. */
int emb_timer_start(emb_ctx* ctx);
int emb_timer_stop(emb_ctx* ctx, double* ms);
/* cudaProfilerStart (on=1) / cudaProfilerStop (on=0) after a stream sync: window for `ncu --profile-from-start off` */
int emb_profiler(emb_ctx* ctx, int on);

/* ---- mesh topology tables ---------------------------------------------------------------------- */
/* Builds on the device what Mesh3D.update() (fem/mesh3d.py:224-355) and Nedelec2.__init__ (fem/elements/
 * nedelec2.py:32-62) build with Python sets and per-tetrahedron loops: edges, tris, tet_to_edge, tet_to_tri, tri_to_edge,
 * tri_to_tet, edge_lengths, tet_to_field, tri_to_field, edge_to_field, all int64 in the reference's layouts.
 * given_edges (2,nE) / given_tris (3,nTri) int64 C-order: the caller's numbering (e.g. the reference's CPython-set order,
 * mesh3d.py:252-271) is kept verbatim and verified against the mesh; NULL: lexicographic numbering of the sorted
 * vertex tuples.  emb_topology_get copies the tables out (any pointer may be NULL) and frees the device copies. */
int emb_topology_build(emb_ctx* ctx, int64_t nN, int64_t nT, const double* nodes_n3, const int64_t* tets_n4,
                       int64_t nE_given, const int64_t* given_edges, int64_t nTri_given, const int64_t* given_tris,
                       int64_t* nE_out, int64_t* nTri_out);
int emb_topology_get(emb_ctx* ctx, int64_t* edges_2xnE, int64_t* tris_3xnTri, int64_t* tet_to_edge_6xnT,
                     int64_t* tet_to_tri_4xnT, int64_t* tri_to_edge_3xnTri, int64_t* tri_to_tet_2xnTri, double* edge_lengths,
                     int64_t* tet_to_field_20xnT, int64_t* tri_to_field_8xnTri, int64_t* edge_to_field_2xnE);

/* ---- mesh + DOF tables (input contract of Nedelec2 / Mesh3D; consumed, never renumbered) ---- */
/* replaces the array gathering at fem/physics/edm/optimized_assembly.py:47-57 */
int emb_upload_mesh(emb_ctx* ctx, int64_t nN, int64_t nT, int64_t nE, int64_t nTri,
                    const double* nodes_n3, const int64_t* tets_n4, const int64_t* tris_n3,
                    const int64_t* tet_to_field_20xnT, const int64_t* tri_to_field_8xnTri);
/* er/ur per tet (fem/physics/edm/emfreq3d.py:615-616) */
int emb_upload_materials(emb_ctx* ctx, const emb_c128* er_3x3xnT, const emb_c128* ur_3x3xnT);

/* ---- assembly ------------------------------------------------------------------------------- */
/* One-time symbolic phase: canonical CSR pattern of E/B (sorted columns, duplicates merged, explicit
 * zeros kept) = the pattern coo_matrix(...).tocsr() produces at optimized_assembly.py:61-62. */
int emb_symbolic(emb_ctx* ctx);
/* Numeric phase: element kernel + deterministic reduction -> E (curl-curl) and B (mass) values.
 * Replaces tet_mass_stiffness_matrices(field, er, ur) (optimized_assembly.py:43-64). */
int emb_assemble_KM(emb_ctx* ctx);
/* Numeric-phase algorithm (no reference counterpart; results agree to rounding, each is bitwise reproducible):
 *   0 (default) fused: per-tetrahedron records + one warp per edge / face that evaluates its element-matrix rows and
 *     writes the finished CSR rows once (no COO intermediate); needs the Nedelec2 table structure
 *     tet_to_field[c+10] = tet_to_field[c] + nE + nTri (fem/elements/nedelec2.py:46-50), else mode 1 is used;
 *   1 element kernel -> COO scratch [nT][20][20] -> deterministic row reduction (the reference's own two-step shape,
 *     optimized_assembly.py:76-116 + coo.tocsr()). */
int emb_assemble_mode(emb_ctx* ctx, int mode);
/* Mode-1 tuning: tets per chunk of the element kernel -> reduction pipeline
 * (0 = default = single pass through HBM; e.g. 32 x SM count keeps the COO scratch of a chunk resident in L2)
 * and whether the scratch is pinned in L2 with a persisting access-policy window.  Results are bitwise independent
 * of both settings. */
int emb_assemble_config(emb_ctx* ctx, int64_t chunk_tets, int persist_l2);
int64_t emb_n_field(const emb_ctx* ctx);
int64_t emb_nnz(const emb_ctx* ctx);
/* copy the CSR out (parity checks, and the scipy csr_matrix the Assembler seam must return).
 * which: 0=E (K), 1=B (M), 2=A(f) of the last emb_form_A (solve-space pattern).  Any pointer may be NULL. */
int emb_get_csr(emb_ctx* ctx, int which, int64_t* indptr, int32_t* indices, emb_c128* data);
int64_t emb_csr_rows(const emb_ctx* ctx, int which);
int64_t emb_csr_nnz(const emb_ctx* ctx, int which);
/* element matrices of tets [t0,t1) in the reference's slot order p=400(t-t0)+20i+j
 * (optimized_assembly.py:109-115); debug/parity only */
int emb_element_matrices(emb_ctx* ctx, int64_t t0, int64_t t1, emb_c128* E400, emb_c128* B400);

/* ---- surface (Robin) terms --------------------------------------------------------------------- */
/* Define surface `sid` (0..15) on global triangle ids.  frame=0: port frame, local = basis_inv @ (x - origin)
 * as assemble_robin_bc_excited does (assembler.py:100-123); frame=1: per-triangle frame of ned2_tri_stiff
 * (fem/mth/tri.py:709-723) as assemble_robin_bc does (assembler.py:125-144).  Computes the gamma-free
 * matrix S (B_p = gamma*S) on the global pattern once. */
int emb_surface_define(emb_ctx* ctx, int sid, int64_t ntri, const int64_t* tri_ids, int frame,
                       const double* basis_inv_3x3, const double* origin_3);
/* Dunavant-4 points of the surface in its local frame, xy (2,6,ntri): generate_points (assembler.py:63-81) */
int emb_surface_points(emb_ctx* ctx, int sid, double* xy_2x6xntri);
/* forcing vector from incident field samples U (3,6,ntri) c128 (compute_bc_entries, assembler.py:83-98);
 * result kept on the device as the RHS of surface sid and optionally copied to b_full (length n_field). */
int emb_surface_set_U(emb_ctx* ctx, int sid, const emb_c128* U_3x6xntri, emb_c128* b_full);
/* B_p = gamma*S as COO over the surface's triangles in the reference's order (64 per triangle); parity only */
int emb_surface_blocks(emb_ctx* ctx, int sid, double* S_ntrix64);

/* ---- Dirichlet elimination + A(f) ----------------------------------------------------------- */
/* PEC dof ids (assembler.py:348-359,384-385).  Builds the solve-space pattern (rows/cols of solve_ids only),
 * i.e. A[np.ix_(solve_ids, solve_ids)] of fem/solver.py:434, once instead of per solve. */
int emb_set_dirichlet(emb_ctx* ctx, int64_t npec, const int64_t* pec_ids);
int64_t emb_n_solve(const emb_ctx* ctx);
/* kept dofs in ascending order = the reference's solve_ids (assembler.py:385) */
int emb_get_solve_ids(emb_ctx* ctx, int64_t* solve_ids);
/* Internal numbering of the solve space.  When both functions of every kept edge / face are kept (PEC elimination always
 * does), the library orders the solve space by pairs (2j, 2j+1 = the two functions of kept entity j) so that A(f) is a
 * block-CSR matrix of 2x2 blocks.  perm[s] = solve index of solve_ids[s].  Everything that crosses this ABI on the solve
 * space - emb_get_csr(which=2), emb_spmv_host, the rows of top-level emb_aux_add* matrices, the device vectors of
 * emb_recycle_export/import - is in SOLVE-INDEX order; full-space vectors (x_full, b_full) are unaffected. */
int emb_get_solve_perm(emb_ctx* ctx, int64_t* perm);
int emb_is_paired(const emb_ctx* ctx);
/* A(f) = E - k0^2 B + sum_s gamma[s] S_s on the solve-space pattern (assembler.py:333,383) */
int emb_form_A(emb_ctx* ctx, double k0, int nsurf, const int* sids, const emb_c128* gammas);

/* ---- linear algebra ------------------------------------------------------------------------------ */
/* y = A x on the solve space (host vectors of length n_solve); parity + SpMV benchmark */
int emb_spmv_host(emb_ctx* ctx, const emb_c128* x, emb_c128* y);
/* times `reps` SpMVs on resident device vectors, returns avg ms per SpMV */
int emb_spmv_bench(emb_ctx* ctx, int reps, double* ms_per_spmv);

/* Auxiliary spaces of the additive multilevel preconditioner (precond = 3):
 *   M^-1 = D^-1 + sum_k R_k diag(R_k^T A R_k)^-1 R_k^T
 * R_k is a real sparse transfer matrix on the SOLVE space (n_solve x ncol), given as CSR together with its transpose
 * (gradients of the quadratic Lagrange space, the Whitney space, and P1 gradients: emerge_b200/auxspace.py).
 * The reference has no counterpart (it solves directly, fem/solver.py:243-309). */
int emb_aux_clear(emb_ctx* ctx);
int emb_aux_add(emb_ctx* ctx, int64_t ncol, const int64_t* R_indptr, const int32_t* R_indices, const double* R_data,
                const int64_t* RT_indptr, const int32_t* RT_indices, const double* RT_data);
/* Tree of auxiliary spaces with multilevel solvers.  A space is given by its real transfer matrix R (nrow x ncol) to its
 * parent (parent < 0: the solve space, else the index of an earlier emb_aux_add* call, nrow = that space's ncol).
 * solver 0: diagonal of R^T A R (top-level only); solver 1: V-cycle of AMG hierarchy hid on an ncol-sized real SPD matrix,
 * times 1 (scale_mode 0) or -1/k0^2 (scale_mode 1: the gradient spaces, on which A(f) = -k0^2 (grad, eps grad)).
 * emb_aux_clear also frees the hierarchies. */
int emb_aux_add_ex(emb_ctx* ctx, int64_t nrow, int64_t ncol, const int64_t* R_indptr, const int32_t* R_indices,
                   const double* R_data, const int64_t* RT_indptr, const int32_t* RT_indices, const double* RT_data,
                   int parent, int solver, int hid, int scale_mode);
/* Top-level auxiliary spaces built on the device from the uploaded mesh tables (no reference counterpart): which = 0 the
 * gradients of the P2 Lagrange space (G), which = 1 the Whitney space (P), restricted to the solve space, rows in
 * solve-index order, columns touched by an eliminated dof dropped; appended like emb_aux_add (diagonal solver) when at
 * least one column remains.  edges (2,nE) as fem/mesh3d.py holds them; face_tables: the 18 numbers of
 * emerge_b200/auxspace.py::_face_tables() (vertex, edge, Whitney targets x 3 x (u, w)).  ncol_out: columns kept;
 * bad_out (optional, nN + nE or nE bytes): 1 for every dropped column.  emb_aux_get copies a space's CSR to the host. */
int emb_aux_build_top(emb_ctx* ctx, int which, const int64_t* edges_2xnE, const double* face_tables_18, int64_t* ncol_out,
                      unsigned char* bad_out);
int emb_aux_get(emb_ctx* ctx, int idx, int64_t* nrow, int64_t* ncol, int64_t* nnz, int64_t* rptr, int32_t* rcol, double* rval);
/* Smoothed-aggregation hierarchy, built on the host (emerge_b200/amg.py), finest level first.  Every level but the last
 * carries its matrix A (real CSR), 1/diag(A), the Jacobi damping omega, the prolongator P (n x ncoarse) and P^T; the last
 * level has ncoarse = 0, null matrices, and gets the dense inverse of its matrix (row-major n x n). */
int emb_amg_create(emb_ctx* ctx, int* hid);
int emb_amg_add_level(emb_ctx* ctx, int hid, int64_t n, const int64_t* A_indptr, const int32_t* A_indices,
                      const double* A_data, const double* dinv, double omega, int64_t ncoarse, const int64_t* P_indptr,
                      const int32_t* P_indices, const double* P_data, const int64_t* PT_indptr, const int32_t* PT_indices,
                      const double* PT_data);
int emb_amg_set_coarse_inverse(emb_ctx* ctx, int hid, int64_t n, const double* Ainv_nxn);
/* average device time (ms) of the operator applications sampled with CUDA events inside the solves since the last
 * call, and their count; same for the preconditioner applications sampled next to them */
int emb_spmv_sampled(emb_ctx* ctx, double* avg_ms, int64_t* count);
int emb_precond_sampled(emb_ctx* ctx, double* avg_ms, int64_t* count);
/* operator application on resident vectors: nv = 1, 2, 4 interleaved right-hand sides; fp32 = 1: the complex64
 * symmetric-part operator of the inner iteration, 0: A(f) in complex128 */
int emb_spmv_bench_ex(emb_ctx* ctx, int reps, int nv, int fp32, double* ms_per_spmv);
/* solver switches (no reference counterpart): complex64 storage of the inner operator (default 1), auxiliary spaces of
 * the preconditioner on concurrent streams (default 1).  Results of the second are bitwise independent of it. */
int emb_solver_config(emb_ctx* ctx, int inner_fp32, int side_streams);
/* 1 (default): the right-hand sides of an emb_solve_multi group share one Krylov space (block COCR); 0: independent
 * recurrences in lockstep.  Padded groups, empty right-hand sides and a breakdown of the block recurrence use 0. */
int emb_solver_block(emb_ctx* ctx, int on);
/* The reduced basis as an extra coarse space of the preconditioner, M^-1 += U Ceff U^T (csrc/recycle.cuh::
 * rc_coarse_update).  Library default 0; the sweep driver (emerge_b200/sweep.py) switches it on: measured on the 1M-tet
 * waveguide, first 20 points of the sweep, 5,130 -> 3,420 block iterations (profiles/r2_coarse_basis_1M.json).
 * Clears the basis. */
int emb_solver_coarse_basis(emb_ctx* ctx, int on);
/* iterations replayed from the captured CUDA graph so far (the kernels inside are counted by emb_launch_count) */
int64_t emb_graph_launch_count(const emb_ctx* ctx);

typedef struct {
    int method;        /* 0 = GMRES(restart), 1 = BiCGStab, 2 = COCR on the symmetric part + defect correction */
    int precond;       /* 0 = none, 1 = Jacobi, 2 = block-Jacobi (2x2 edge/face pairs), 3 = block-Jacobi + auxiliary spaces */
    int restart;       /* GMRES restart length */
    int maxit;
    double rtol;       /* ||b-Ax||/||b|| */
    int use_x0;        /* 1: x_full holds an initial guess */
} emb_solve_opts;
typedef struct {
    int iters;
    double relres;     /* true relative residual recomputed at exit */
    double ms;         /* device time */
    int spmvs;
} emb_solve_info;
/* Solve A x = b_sid (RHS of surface sid from emb_surface_set_U) on the solve space; x_full has n_field
 * entries with zeros at Dirichlet DOFs, as SolveRoutine.solve returns it (fem/solver.py:405-469). */
int emb_solve(emb_ctx* ctx, int sid, const emb_solve_opts* opts, emb_c128* x_full, emb_solve_info* info);
/* All ports of a frequency point in LOCKSTEP (the per-port loop of emfreq3d.py:683-694): nrhs = 1..4 surfaces whose
 * right-hand sides are solved together on interleaved vectors, so A(f) and the preconditioner are read once per
 * iteration for all of them.  x_full: NULL, or nrhs host pointers (each NULL or n_field entries; initial guesses when
 * use_x0).  infos: nrhs entries (iters / ms are those of the group).  Column 0 becomes the device-resident "last
 * solution"; emb_select_solution(k) switches emb_interp_last to column k. */
int emb_solve_multi(emb_ctx* ctx, int nrhs, const int* sids, const emb_solve_opts* opts, emb_c128* const* x_full,
                    emb_solve_info* infos);
int emb_select_solution(emb_ctx* ctx, int k);
/* Field output mode (emfreq3d.py:699 `data._fields[port] = solution`).  on = 1: the host copies of emb_solve_multi's
 * solutions are issued on a copy stream from per-column staging vectors and overlap the next point's work; the caller's
 * buffers (pinned memory) are valid after emb_fields_sync().  on = 0 (default): the copy has finished on return. */
int emb_fields_async(emb_ctx* ctx, int on);
int emb_fields_sync(emb_ctx* ctx);
/* same with an explicit host RHS of length n_field (the b + port_vectors[p] of emfreq3d.py:691) */
int emb_solve_rhs(emb_ctx* ctx, const emb_c128* b_full, const emb_solve_opts* opts, emb_c128* x_full,
                  emb_solve_info* info);

/* ---- subspace recycling across the frequency points of a sweep ------------------------------------------
 * The reference refactorises A(f) at every frequency (fem/solver.py:264-275, emfreq3d.py:658-694).  Here every
 * solve that iterated leaves its correction in a set of at most max_vectors directions U (shared by all ports).
 * A(f) is affine in K, M and the surface matrices, so the products W_t U are formed once per direction and kept in an
 * orthonormal basis; at a new frequency A(f) U is a small host-side matrix combination and each solve starts from the
 * minimum-residual combination over span(U) (csrc/recycle.cuh).  The exit test is unchanged (true residual of A(f) in
 * FP64 <= rtol).
 * A solve that does have to iterate is run to snapshot_rtol_factor * rtol (0 < factor <= 1, default 0.3) so that the
 * direction it leaves is accurate enough for neighbouring points to be accepted without iterating (0.3 is what four
 * steps of the defect correction reach; 0.1 costs a fifth step - 13 % more iterations on the 1M-tet sweep - and accepts
 * the same 190 of 201 points).
 * max_vectors = 0 switches it off and frees the vectors ((1 + T) * max_vectors * n_solve * 16 B of HBM, T = 2 + number
 * of surfaces in A(f)). */
int emb_recycle_config(emb_ctx* ctx, int max_vectors, double snapshot_rtol_factor);
/* n: directions held; spmvs: matrix-vector products spent on W_t u so far; last_proj_relres: relative residual of the recycled
 * start vector in the last solve (-1 if none) */
int emb_recycle_info(emb_ctx* ctx, int* n, int64_t* spmvs, double* last_proj_relres);
/* directions accepted into the basis since the context was created.  Monotonic: unlike the `n` of emb_recycle_info it
 * does not drop when the ring is compacted or emb_recycle_config resets the basis, so "how many did the last solve
 * add" is a difference of two readings. */
int64_t emb_recycle_accepted(const emb_ctx* ctx);
/* exchange of directions between the ranks of a frequency-sharded sweep; d_dst / d_src are DEVICE pointers to
 * n_solve complex128 values (the host side moves them with NCCL).  j = 0 is the NEWEST direction. */
int emb_recycle_export(emb_ctx* ctx, int j, void* d_dst);
int emb_recycle_import(emb_ctx* ctx, const void* d_src);

/* ---- field evaluation (S-parameter extraction) -------------------------------------------------- */
/* E-field of solution x_full (host, n_field) at npts points, point k lying in tet tet_ids[k]:
 * the per-point part of ned2_tet_interp (fem/mth/tet.py:371-497).  E_3xnpts is (3,npts) c128. */
int emb_interp(emb_ctx* ctx, const emb_c128* x_full, int64_t npts, const int64_t* tet_ids,
               const double* xyz_3xnpts, emb_c128* E_3xnpts);
/* same, using the device-resident solution of the last emb_solve (no x upload) */
int emb_interp_last(emb_ctx* ctx, int64_t npts, const int64_t* tet_ids, const double* xyz_3xnpts,
                    emb_c128* E_3xnpts);


/* ---- field post-processing ------------------------------------------------------------------------ */
/* Tetrahedron of each point as the reference's interpolation picks it: the LAST tet whose local-coordinate test passes
 * (fem/mth/tet.py:393-497, test at :425); -1 for points outside the mesh.  xyz is (3,npts) C-order. */
int emb_locate_points(emb_ctx* ctx, int64_t npts, const double* xyz_3xnpts, int64_t* tet_ids);
/* E (3,npts) and optionally H = curl(E) * curl_const[tet] (3,npts) of solution x_full (n_field entries; NULL: the
 * device-resident solution of the last solve) at arbitrary points: EMDataSet.interpolate (fem/physics/edm/emdata.py:181-
 * 199) = ned2_tet_interp + ned2_tet_interp_curl (fem/mth/tet.py:371-626).  tet_ids NULL: located on the device.
 * curl_const (nT complex; the reference passes 1/(-j w mu0 mu_r[0,0,tet]), emdata.py:193) and H may be NULL. */
int emb_interp_fields(emb_ctx* ctx, const emb_c128* x_full, int64_t npts, const double* xyz_3xnpts, const int64_t* tet_ids,
                      const emb_c128* curl_const_nT, emb_c128* E_3xnpts, emb_c128* H_3xnpts);
/* Stratton-Chu far field of surface samples: stratton_chu_ff (fem/physics/edm/sc.py:27-142).  E, H (3,nsrc) tangential-field
 * samples at pos (3,nsrc) with area-weighted outward normals wnormal (3,nsrc) (built by stratton_chu, sc.py:144-166, from
 * a SurfaceMesh); directions theta, phi (nout each; r = (cos th cos ph, cos th sin ph, sin th), sc.py:78-80); outputs
 * (3,nout).  Inputs are rounded to float32 as the reference does (sc.py:172-178), samples below 1e-3 of the largest |E|
 * are skipped (sc.py:57-74), the sums run in FP64. */
int emb_stratton_chu(emb_ctx* ctx, int64_t nsrc, const emb_c128* E_3xn, const emb_c128* H_3xn, const double* pos_3xn,
                     const double* wnormal_3xn, int64_t nout, const double* theta, const double* phi, double k0,
                     emb_c128* Eout_3xnout, emb_c128* Hout_3xnout);

/* ---- port boundary-mode analysis ------------------------------------------------------------------- */
/* The 14 x 14 element matrices (A, B) of the generalised eigenproblem A e = -beta^2 B e on the port triangles:
 * generalized_matrix_GQ / _matrix_builder (fem/physics/edm/nedeleclegrange2.py:223-417).  xy (2,n_nodes): port-local node
 * coordinates (rows 0, 1 of pinv(basis) @ nodes, :43); tris (3,nt), edges (2,ne), tri_to_edge (3,nt): the SurfaceMesh
 * tables (fem/mesh3d.py:458-575); er, ur (3,3,nt) as assemble_bma_matrices slices them (assembler.py:262-263).  Outputs
 * (nt,14,14) row-major in the local order of NedelecLegrange2.tri_to_field (fem/elements/nedleg2.py:57-65): 3 edge-a, face-a,
 * 3 edge-b, face-b, 3 vertices, 3 edges. */
int emb_bma_element_matrices(emb_ctx* ctx, int64_t n_tris, int64_t n_nodes, int64_t n_edges, const double* xy_2xn,
                             const int64_t* tris_3xnt, const int64_t* edges_2xne, const int64_t* tri_to_edge_3xnt,
                             const emb_c128* er_3x3xnt, const emb_c128* ur_3x3xnt, double k0, emb_c128* A_ntx14x14,
                             emb_c128* B_ntx14x14);
/* Shift-invert operator of that eigenproblem, v -> (A - sigma B)^-1 B v, for dense n x n row-major A, B restricted to
 * the solve ids: the device-side replacement of the factorisations inside SolverLAPACK.eig / SolverARPACK.eig
 * (fem/solver.py:311-357; the reference's shift is sigma = -target_kz^2, :355).  setup inverts A - sigma B on the device
 * (Gauss-Jordan, partial pivoting), apply is two dense matrix-vector products; one operator per context. */
int emb_shift_invert_setup(emb_ctx* ctx, int64_t n, const emb_c128* A_nxn, const emb_c128* B_nxn, double sigma_re, double sigma_im);
int emb_shift_invert_apply(emb_ctx* ctx, const emb_c128* x_n, emb_c128* y_n);
int emb_shift_invert_free(emb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
