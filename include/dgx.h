/*
 * dgx.h -- C ABI of the B200-native matrix-free DG operator / multigrid library (libdgx.so).
 *
 * This is the drop-in boundary for the one hot path of the deal.II code gallery's
 * NavierStokes_TRBDF2_DG program: the operator evaluation behind
 * `NavierStokesProjectionOperator : MatrixFreeOperators::Base<dim, Vec>`
 * (NavierStokes_TRBDF2_DG/navier_stokes_TRBDF2_DG.cc:242-2060, "NS.cc" below) and the deal.II
 * solver objects wired around it in diffusion_step / projection_step / project_grad
 * (NS.cc:2424-2577).  The reference has no FFI: its "plugin API" is C++ duck typing against
 * deal.II class templates.  Each entry point below therefore cites the deal.II/NS.cc member it
 * replaces; `include/dgx_dealii_shim.h` (header-only C++) re-exposes them under the reference's
 * own class and method names so that drivers written like NS.cc:2424-2577 compile unchanged.
 *
 * Conventions
 *   - every function returns 0 on success, a negative DGX_ERR_* code on failure; the message is
 *     available from dgx_last_error() (replaces C++ exceptions: AssertThrow NS.cc:2247,
 *     SolverControl::NoConvergence -> DGX_ERR_NO_CONVERGENCE; `main` maps them to exit code 1,
 *     NS.cc:3055-3077).
 *   - handles are opaque; the library owns device memory; it never frees caller memory.
 *   - one host thread per context; all work of a context is ordered on its CUDA stream.
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - vectors use deal.II's DG layout: global dof = cell * dofs_per_cell + comp * n^dim + lex(i,j,k),
 *     cells in refine_global (Morton) order; local storage is [owned dofs | ghost dofs] like
 *     LinearAlgebra::distributed::Vector.
 */
#ifndef DGX_H
#define DGX_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dgx_ctx dgx_ctx;
typedef struct dgx_mesh dgx_mesh;
typedef struct dgx_mf dgx_mf;
typedef struct dgx_vec dgx_vec;
typedef struct dgx_op dgx_op;
typedef struct dgx_mg dgx_mg;
typedef struct dgx_cheb dgx_cheb;

enum { DGX_F64 = 0, DGX_F32 = 1 };
enum { DGX_OP_VELOCITY = 1, DGX_OP_PRESSURE = 2, DGX_OP_MASS = 3 }; /* = NS_stage, NS.cc:432-443 */
enum {
  DGX_OK = 0,
  DGX_ERR_INVALID = -1,
  DGX_ERR_CUDA = -2,
  DGX_ERR_NCCL = -3,
  DGX_ERR_NO_CONVERGENCE = -4,
  DGX_ERR_UNSUPPORTED = -5
};
#define DGX_NCCL_ID_BYTES 128

/* ---- error / context ---------------------------------------------------------------------- */
const char *dgx_last_error(void);
/* build tag, e.g. "dgx 0.1 sm_100a" */
const char *dgx_version(void);
/* number of CUDA kernels this library has launched since load (all contexts) */
long long dgx_kernel_launch_count(void);

/* replaces MPI_InitFinalize + MPI_COMM_WORLD (NS.cc:3038, 2216): one context per GPU / rank.
 * nccl_id: DGX_NCCL_ID_BYTES from dgx_nccl_unique_id() of rank 0 (may be NULL when nranks == 1). */
int dgx_nccl_unique_id(void *out_id);
int dgx_ctx_create(int device, int rank, int nranks, const void *nccl_id, dgx_ctx **out);
int dgx_ctx_destroy(dgx_ctx *ctx);
/* run all work of this context on a caller-owned stream (cudaStream_t); NULL restores the own stream */
int dgx_ctx_set_stream(dgx_ctx *ctx, void *cuda_stream);
void *dgx_ctx_get_stream(dgx_ctx *ctx);
int dgx_ctx_synchronize(dgx_ctx *ctx);
/* CUDA-event stopwatch on the context's stream (TimerOutput sections NS.cc:2191, 2425, 2471 measure
 * wall time per solve; this measures device time). stop returns milliseconds. */
int dgx_ctx_timer_start(dgx_ctx *ctx);
int dgx_ctx_timer_stop(dgx_ctx *ctx, double *ms);

/* ---- mesh: GridGenerator::subdivided_hyper_rectangle + refine_global (NS.cc:2261-2270) ----- */
/* periodic_mask bit d: periodic in direction d; boundary_ids[2*dim]: id of face 2d (lower), 2d+1 (upper) */
int dgx_mesh_create_cartesian(int dim, const int *cells0, int n_refine, const double *lo, const double *hi,
                              int periodic_mask, const int *boundary_ids, dgx_mesh **out);
int dgx_mesh_destroy(dgx_mesh *mesh);
/* Non-Cartesian geometry on the same topology (what GridTools::transform / a manifold give the reference's triangulation before
 * MatrixFree::reinit(MappingQ1, ...), NS.cc:2264, 2329): positions of the (cells + 1)^dim vertices of `level` (-1: finest), lattice
 * order with x fastest, dim doubles per vertex.  Every cell of that level becomes the multilinear (MappingQ1) image of the unit
 * cell; MatrixFree objects created afterwards carry the metric at every quadrature point (the pressure operator, its diagonal,
 * Chebyshev / multigrid / CG on top of it; the other operators return DGX_ERR_UNSUPPORTED on such a level).  Periodic directions
 * need vertex positions that are periodic up to a translation.  Set the vertices of every level a multigrid hierarchy uses. */
int dgx_mesh_set_vertices(dgx_mesh *mesh, int level, const double *coords, long long n_values);
int dgx_mesh_is_general(const dgx_mesh *mesh, int level);
int dgx_mesh_n_levels(const dgx_mesh *mesh); /* Triangulation::n_global_levels, NS.cc:2351 */
long long dgx_mesh_n_cells(const dgx_mesh *mesh, int level);
/* integer coordinates (x,y,z) of cells [first, first+count) of `level` in deal.II cell order; out[count*dim] */
int dgx_mesh_cell_coords(const dgx_mesh *mesh, int level, long long first, long long count, int *out);
/* global face-neighbour table: out[count*2*dim]; entry = neighbour cell index, or -1-boundary_id */
int dgx_mesh_neighbors(const dgx_mesh *mesh, int level, long long first, long long count, long long *out);

/* ---- MatrixFree<dim,Number>::reinit (NS.cc:2329 fine/double, NS.cc:2374-2375 per level/float) - */
/* level = -1: active (finest) cells; partition = contiguous chunk `rank` of `nranks` of the cell order */
int dgx_mf_create(dgx_ctx *ctx, dgx_mesh *mesh, int level, int dtype, dgx_mf **out);
int dgx_mf_destroy(dgx_mf *mf);
long long dgx_mf_n_owned_cells(const dgx_mf *mf);
long long dgx_mf_n_ghost_cells(const dgx_mf *mf);
long long dgx_mf_first_owned_cell(const dgx_mf *mf);
/* local face-neighbour table out[2*dim][n_owned] (SoA as stored in HBM): local cell index
 * (>= n_owned: ghost slot) or -1-boundary_id */
int dgx_mf_local_neighbors(const dgx_mf *mf, int *out);
/* global cell index of every ghost slot, out[n_ghost] */
int dgx_mf_ghost_cells(const dgx_mf *mf, long long *out);
/* bytes of metric data in HBM (J^-1 J^-T JxW per cell point; J^-1 n JxW of both sides and the penalty per face point) that a
 * degree-`degree` operator on a non-Cartesian level streams per application; 0 before its first application / on Cartesian levels */
long long dgx_mf_general_metric_bytes(dgx_mf *mf, int degree);
/* frees the metric tables of a non-Cartesian level (100+ bytes per dof); they are rebuilt by the next operator application */
int dgx_mf_release_general_metric(dgx_mf *mf);
/* the same tables in double precision on the host (introspection / tests; pure host logic, no device needed), n = degree + 1:
 *   cell_tables[cell][dim (dim+1)/2 + 1][n^dim]: (J^-1 J^-T)_{ee'} JxW in the order 00, 11, (22,) 01, (02, 12), then JxW
 *   face_tables[cell][2 dim faces][2 dim + 1][n^(dim-1)]: (J_own^-1 n)_e JxW_f, (J_nbr^-1 n)_e JxW_f, sigma JxW_f   (n = own outward normal)
 * NULL pointers only query the sizes. */
int dgx_mf_general_metric_tables(const dgx_mf *mf, int degree, double *cell_tables, long long cell_capacity, double *face_tables, long long face_capacity,
                                 long long *n_cell_values, long long *n_face_values);
/* ghost-exchange plan (Utilities::MPI::Partitioner import / export lists): peers in ascending rank order */
int dgx_mf_n_peers(const dgx_mf *mf);
int dgx_mf_peer_info(const dgx_mf *mf, int i, int *rank, long long *n_send, long long *recv_first, long long *recv_count);
int dgx_mf_peer_send_cells(const dgx_mf *mf, int i, int *out_local_cells);

/* unit support points of FE_DGQ(degree) in 1-D (Gauss-Lobatto nodes in [0,1], FE_DGQ NS.cc:2158-2159): out[degree+1];
 * dof (i,j,k) of a cell sits at the tensor product of these (VectorTools::interpolate, NS.cc:2393-2396) */
int dgx_fe_unit_support_points(int degree, double *out);

/* ---- LinearAlgebra::distributed::Vector<Number> (NS.cc:2102-2115, 2330-2341) ---------------- */
/* MatrixFree::initialize_dof_vector for FESystem(FE_DGQ(degree), ncomp); zero-initialised */
int dgx_vec_create(dgx_mf *mf, int degree, int ncomp, dgx_vec **out);
int dgx_vec_destroy(dgx_vec *v);
long long dgx_vec_size(const dgx_vec *v);       /* locally_owned_size */
long long dgx_vec_ghost_size(const dgx_vec *v);
long long dgx_vec_global_size(const dgx_vec *v); /* size() */
int dgx_vec_dtype(const dgx_vec *v);
void *dgx_vec_data(dgx_vec *v);                  /* device pointer to the owned range */
int dgx_vec_copy_from_host(dgx_vec *v, const void *host, long long n);
int dgx_vec_copy_to_host(const dgx_vec *v, void *host, long long n);
int dgx_vec_set(dgx_vec *v, double value);                                /* operator=(Number) */
int dgx_vec_copy(dgx_vec *dst, const dgx_vec *src);                       /* operator=(Vector), also double<->float */
int dgx_vec_equ(dgx_vec *dst, double a, const dgx_vec *src);              /* equ  */
int dgx_vec_add(dgx_vec *dst, double a, const dgx_vec *src);              /* add(a, V) */
int dgx_vec_sadd(dgx_vec *dst, double s, double a, const dgx_vec *src);   /* sadd(s, a, V) */
int dgx_vec_scale(dgx_vec *dst, double a);                                /* operator*= */
int dgx_vec_scale_by(dgx_vec *dst, const dgx_vec *factors);               /* scale(V): entrywise */
int dgx_vec_dot(const dgx_vec *a, const dgx_vec *b, double *out);         /* operator*, all-reduced */
int dgx_vec_l2_norm(const dgx_vec *a, double *out);
int dgx_vec_linfty_norm(const dgx_vec *a, double *out);
int dgx_vec_mean_value(const dgx_vec *a, double *out);
int dgx_vec_add_and_dot(dgx_vec *dst, double a, const dgx_vec *v, const dgx_vec *w, double *out);
/* the vector work of one SolverCG iteration in one pass: x += a*p ; r += b*v ; *out = r . r */
int dgx_vec_cg_update(dgx_vec *x, double a, const dgx_vec *p, dgx_vec *r, double b, const dgx_vec *v, double *out);
/* interpolate_velocity (NS.cc:2403-2418): dst = a*x - b*y in one pass, bit-identical to equ / equ / operator-= */
int dgx_vec_interpolate(dgx_vec *dst, double a, const dgx_vec *x, double b, const dgx_vec *y);
/* one (classical) Gram-Schmidt sweep: h[j] = w . V[j] (j < k <= 32, one pass + one all-reduce), then w -= sum_j h[j] V[j] */
int dgx_vec_orthogonalize(dgx_vec *w, dgx_vec *const *V, int k, double *h);
/* same sweep (k <= 31); *w_dot_w = w . w BEFORE the sweep, reduced in the same pass: SolverGMRES re-orthogonalises only when the
 * norm after the sweep is below 10 sqrt(eps) times the norm before (Kelley 1995, the test deal.II's SolverGMRES applies) */
int dgx_vec_orthogonalize_checked(dgx_vec *w, dgx_vec *const *V, int k, double *h, double *w_dot_w);
int dgx_vec_update_ghost_values(dgx_vec *v);
int dgx_vec_zero_out_ghost_values(dgx_vec *v);
/* compress(VectorOperation::add): a no-op by construction (cell-centric kernels never write ghosts) */
int dgx_vec_compress_add(dgx_vec *v);

/* ---- NavierStokesProjectionOperator (NS.cc:242-2060) ----------------------------------------- */
/* kind = NS_stage; degree = degree of the space the operator acts on; Re, dt as Data_Storage */
int dgx_op_create(dgx_mf *mf, int kind, int degree, double Re, double dt, dgx_op **out);
int dgx_op_destroy(dgx_op *op);
int dgx_op_set_dt(dgx_op *op, double dt);                 /* NS.cc:415 */
int dgx_op_set_TR_BDF2_stage(dgx_op *op, int stage);      /* NS.cc:424 */
int dgx_op_set_u_extr(dgx_op *op, const dgx_vec *src);    /* NS.cc:448: copies + ghost update */
long long dgx_op_m(const dgx_op *op);                     /* Base::m() */
int dgx_op_vmult(dgx_op *op, dgx_vec *dst, dgx_vec *src);      /* Base::vmult: dst = A src */
int dgx_op_vmult_add(dgx_op *op, dgx_vec *dst, dgx_vec *src);  /* Base::vmult_add -> apply_add NS.cc:1398 */
int dgx_op_Tvmult(dgx_op *op, dgx_vec *dst, dgx_vec *src);     /* Base::Tvmult = apply_add (NS.cc:281) */
int dgx_op_Tvmult_add(dgx_op *op, dgx_vec *dst, dgx_vec *src);
/* dst = M^-1 src for the DG mass operator (NS_stage 3): exact tensor-product inverse, an opt-in replacement of the CG solve
 * of project_grad (NS.cc:2574-2576) */
int dgx_op_mass_inverse(dgx_op *op, dgx_vec *dst, dgx_vec *src);
int dgx_op_compute_diagonal(dgx_op *op);                  /* NS.cc:2018-2060 (stores the INVERSE) */
int dgx_op_get_matrix_diagonal_inverse(dgx_op *op, dgx_vec **out); /* borrowed handle, NS.cc:2515 */
int dgx_op_precondition_Jacobi(dgx_op *op, dgx_vec *dst, const dgx_vec *src, double omega);
/* Base::el(i, i): the diagonal entry of locally owned row i (= 1 / inverse diagonal; needs compute_diagonal first) */
int dgx_op_el(dgx_op *op, long long i, double *value);
/* right-hand-side linear forms (zero_dst_vector = true); every source gets update_ghost_values first (NS.cc:789-790, 916-917) */
int dgx_op_vmult_rhs_velocity(dgx_op *op, dgx_vec *dst, dgx_vec *const *srcs, int n_srcs);   /* NS.cc:788; srcs as NS.cc:2439 / 2444 */
int dgx_op_vmult_rhs_pressure(dgx_op *op, dgx_vec *dst, dgx_vec *const *srcs, int n_srcs);   /* NS.cc:915; srcs = {u_star, pres_old} */
int dgx_op_vmult_grad_p_projection(dgx_op *op, dgx_vec *dst, const dgx_vec *src);            /* NS.cc:1363 */
/* kernel variant selection for measurement: 0 = auto, 1 = generic, 2 = blocked fast path */
int dgx_op_set_variant(dgx_op *op, int variant);
/* end-to-end entry: host buffers in the DG layout, H2D + vmult + D2H on the context stream */
int dgx_op_vmult_host(dgx_op *op, void *dst_host, const void *src_host, long long n);
/* schedule of the pipelined form of dgx_op_vmult_host for vectors with bytes_per_cell bytes per cell (introspection / tests; pure
 * host logic): upload segments (cells [up_begin, up_end), in upload order) and compute groups (cells [grp_begin, grp_end), in
 * compute order, each computable once upload segment grp_need is in).  chunk_log2 = 0: automatic brick size.  *n_up = 0 means the
 * call is not pipelined (fewer than 16 bricks, non-Cartesian or replicated level).  With several ranks the cells other ranks need are
 * gathered on the host, uploaded and exchanged first; the schedule of the owned cells is the same.  Arrays may be NULL; at most `capacity` entries are written. */
int dgx_mf_host_pipeline_plan(const dgx_mf *mf, long long bytes_per_cell, int chunk_log2, int capacity, long long *up_begin, long long *up_end,
                              long long *grp_begin, long long *grp_end, int *grp_need, int *n_up, int *n_grp);

/* ---- MGTransferMatrixFree<dim,float> (NS.cc:2490-2491): DG embedding of a parent cell into its children ---- */
/* dst_fine += P src_coarse ; dst_coarse += P^T src_fine (vectors on consecutive levels of one mesh) */
int dgx_transfer_prolongate_and_add(dgx_vec *dst_fine, const dgx_vec *src_coarse);
int dgx_transfer_restrict_and_add(dgx_vec *dst_coarse, const dgx_vec *src_fine);
/* (scalar or vector-valued spaces: cells are component-blocked)
 * SolutionTransfer::interpolate for the globally refined case of refine_mesh (NS.cc:2781-2822): dst_fine = embedding of src_coarse into
 * the children of every cell (exact for DG), vectors on consecutive levels of one mesh, any number of components */
int dgx_solution_transfer_refine(dgx_vec *dst_fine, const dgx_vec *src_coarse);

/* ---- PreconditionChebyshev<LevelOp, Vector<float>> with the inverse diagonal inside (NS.cc:2493-2516) ------- */
int dgx_cheb_create(dgx_op *op, int degree, double smoothing_range, int eig_cg_n_iterations, dgx_cheb **out);
int dgx_cheb_destroy(dgx_cheb *c);
int dgx_cheb_estimate(dgx_cheb *c);   /* eigenvalue estimate (lazy in deal.II; needs compute_diagonal first) */
int dgx_cheb_get_estimates(const dgx_cheb *c, double *min_eig, double *max_eig, double *theta, double *delta);
int dgx_cheb_vmult(dgx_cheb *c, dgx_vec *dst, dgx_vec *src);   /* zero initial guess (pre-smoothing) */
int dgx_cheb_step(dgx_cheb *c, dgx_vec *x, dgx_vec *b);        /* non-zero initial guess (post-smoothing) */

/* ---- Multigrid V-cycle + PreconditionMG + MGCoarseGridIterativeSolver (NS.cc:2519-2537) ---------------------- */
/* builds one MatrixFree / pressure operator / Chebyshev(3, 15, 10) smoother per level 0..n_levels-1 (NS.cc:2351-2379,
 * 2501-2517); level 0 is solved by plain CG with the tolerance given to dgx_mg_setup (the outer SolverControl,
 * NS.cc:2486, 2520) */
int dgx_mg_create(dgx_ctx *ctx, dgx_mesh *mesh, int degree, int level_dtype, double Re, double dt, dgx_mg **out);
int dgx_mg_destroy(dgx_mg *mg);
int dgx_mg_n_levels(const dgx_mg *mg);
/* smoother_data[level] (PreconditionChebyshev::AdditionalData, NS.cc:2501-2516); defaults are the reference's 3 / 15.0 / 10.
 * Level 0 is solved by the coarse solver: its data (NS.cc:2510-2512) is accepted and ignored, as Multigrid never uses it. */
int dgx_mg_set_smoother(dgx_mg *mg, int level, int degree, double smoothing_range, int eig_cg_n_iterations);
int dgx_mg_set_dt(dgx_mg *mg, double dt);
int dgx_mg_set_TR_BDF2_stage(dgx_mg *mg, int stage);
int dgx_mg_setup(dgx_mg *mg, double coarse_abs_tol, int coarse_max_it);   /* diagonals + eigenvalue estimates */
/* the Chebyshev eigenvalue estimates depend only on (TR_BDF2 stage, dt): by default a setup with an unchanged pair reuses them
 * (enable = 0 restores the reference's recomputation per solve, NS.cc:2501-2517; enable < 0 only queries) */
int dgx_mg_cache_estimates(dgx_mg *mg, int enable, int *was_cached);
int dgx_mg_vmult(dgx_mg *mg, dgx_vec *dst, dgx_vec *src);                 /* PreconditionMG::vmult (fp64 in/out) */
int dgx_mg_coarse_iterations(const dgx_mg *mg, int *out, int capacity, int *count);
int dgx_mg_level_op(dgx_mg *mg, int level, dgx_op **op);                  /* borrowed */

/* ---- SolverCG / SolverGMRES (NS.cc:2451-2463, 2487, 2542-2546, 2575-2576) -------------------------------------- */
/* precond_kind: 0 PreconditionIdentity, 1 PreconditionJacobi (precond = dgx_op* or NULL for A itself),
 * 2 PreconditionMG (precond = dgx_mg*).  abs_tol as SolverControl(max_it, tol).  Returns DGX_ERR_NO_CONVERGENCE
 * (SolverControl::NoConvergence) when max_it is reached; iters / final_res are filled either way. */
int dgx_solve_cg(dgx_op *A, dgx_vec *x, dgx_vec *b, int precond_kind, void *precond, int max_it, double abs_tol, int *iters, double *final_res);
int dgx_solve_gmres(dgx_op *A, dgx_vec *x, dgx_vec *b, int precond_kind, void *precond, int max_it, double abs_tol, int max_basis, int *iters, double *final_res);

/* ---- diagnostics and output of the time loop, evaluated on the device (NS.cc:2608-2708) ---------------------------------- */
/* NavierStokesProjection::compute_lift_and_drag (NS.cc:2642-2708): sum over the boundary faces with id `boundary_id` (the
 * reference: 4) of (1/Re grad u - p I) (-n) JxW at the points of QGauss<dim-1>(degree_p + 2); drag = component 0, lift = component 1,
 * summed over all ranks (Utilities::MPI::sum).  u, p: vectors of one MatrixFree. */
int dgx_compute_lift_and_drag(const dgx_vec *u, const dgx_vec *p, double Re, int boundary_id, double *lift, double *drag);
/* refine_mesh's cell-wise error indicator (NS.cc:2727-2752), evaluated on the device: out_host[owned cell] = diameter(K)^2 *
 * int_K (d u_1 / d x_0 - d u_0 / d x_1)^2 dx with QGauss<dim>(n_q_points_1d) (the reference: degree_p + 2; it takes this vorticity
 * component in 3-D as well), float like Vector<float> estimated_error_per_cell; Cartesian and MappingQ1 levels */
int dgx_estimate_vorticity_indicator(const dgx_vec *u, int n_q_points_1d, float *out_host, long long capacity);
/* GridRefinement::refine_and_coarsen_fixed_number (NS.cc:2771) on one rank, pure host logic: flags[i] = +1 (refine) for the
 * top_fraction * n cells with the largest indicators, -1 (coarsen) for the bottom_fraction * n smallest, else 0.  The refinement
 * itself (hanging-node faces, SolutionTransfer) is not part of this library yet: the flags are what refine_mesh would execute. */
int dgx_refine_and_coarsen_fixed_number(const float *indicators, long long n, double top_fraction, double bottom_fraction, signed char *flags);
/* fields per patch point of output_results: v (dim), p, v_old (dim), vorticity (1 in 2-D, 3 in 3-D) */
int dgx_output_n_fields(int dim);
/* the data DataOut::build_patches(MappingQ1, 1) collects for output_results (NS.cc:2608-2630): one patch per owned cell, its 2^dim
 * vertices in deal.II vertex order (x fastest), out_host[(cell * 2^dim + vertex) * n_fields + field]; the vorticity is
 * PostprocessorVorticity::evaluate_vector_field (NS.cc:131-158) on the velocity gradients at the vertex */
int dgx_output_patch_data(const dgx_vec *u_n, const dgx_vec *pres_n, const dgx_vec *u_n_minus_1, double *out_host, long long capacity);
/* one <Piece> as an UnstructuredGrid .vtu file (ascii arrays; every patch owns its points like DataOutBase::write_vtu):
 * points[n_cells * 2^dim * dim], patch_data as above */
int dgx_write_vtu_piece(const char *path, int dim, long long n_cells, const double *points, const double *patch_data);
/* output_results (NS.cc:2608-2633): patch data from the device, written to `path` (.vtu); with several ranks every rank writes
 * <stem>.<rank>.vtu and rank 0 the index <stem>.pvtu (the reference: one file through MPI I/O, write_vtu_in_parallel) */
int dgx_output_results(const dgx_vec *u_n, const dgx_vec *pres_n, const dgx_vec *u_n_minus_1, const char *path);

#ifdef __cplusplus
}
#endif
#endif /* DGX_H */
