/* gfgpu.h -- C ABI of the B200-native generic weak-form assembly path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  A GetFEM
 * maintainer binds these from ga_workspace::assembly() (see INTEGRATION.md and
 * getfem_b200/shim/gfgpu_getfem_shim.cc); the Python host mirror binds the same symbols
 * through ctypes (getfem_b200/capi.py).
 *
 * Every entry point names the reference interface it replaces; paths are relative to the
 * GetFEM source tree (getfem/getfem v5.5), "C&E.cc" = src/getfem_generic_assembly_compile_and_exec.cc.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; gfgpu_last_error() returns
 *     the message of the last failure on the calling thread (the C++ shim turns it into
 *     GMM_ASSERT1 / gmm::gmm_error, gmm_except.h:55-163);
 *   - "_host" pointers are host memory, "_dev" pointers are device memory of the context's GPU;
 *   - tensors follow the reference: column-major, first index fastest (bgeot_tensor.h:200-230);
 *     local dof of a vector fem = node*Q + q (C&E.cc:5008-5020);
 *   - the tangent is produced column-compressed like gmm::csc_matrix (gmm_matrix.h:506-566):
 *     jc[ndof+1], ir[nnz] (rows ascending inside each column), pr[nnz].  K(r,c): r = Test_ dof,
 *     c = Test2_ dof (C&E.cc:4837-4843).  jc is int64 because nnz exceeds 2^32 at benchmark size.
 *   - there is no CPU fallback anywhere behind this interface.
 */
#ifndef GFGPU_H
#define GFGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gfgpu_ctx gfgpu_ctx;     /* one GPU + one stream */
typedef struct gfgpu_mesh gfgpu_mesh;   /* SoA node coordinates + connectivity */
typedef struct gfgpu_fem gfgpu_fem;     /* element->dof table (mesh_fem) */
typedef struct gfgpu_tables gfgpu_tables; /* reference tables at the quadrature points */
typedef struct gfgpu_term gfgpu_term;   /* one compiled weak-form term + its pattern */
typedef struct gfgpu_matrix gfgpu_matrix; /* the workspace-level tangent: sum of terms, resident on the device */

/* geometric transformation kinds (bgeot_geometric_trans.cc:600-672) */
enum { GFGPU_GT_PK = 0 /* affine simplex, is_linear() */, GFGPU_GT_QK = 1 /* multilinear, K per Gauss point */ };
/* Lagrange fem kinds (getfem_fem.cc:719-843, 1010-1047) */
enum { GFGPU_FEM_PK = 0, GFGPU_FEM_QK = 1 };

/* Expression families = the compiled ga_instruction chains that are replaced:
 *   LAPLACE      "a*Grad_u.Grad_Test_u" / "a*Grad_u:Grad_Test_u"     params = {a}
 *                (add_generic_elliptic_brick, getfem_models.cc:3943-3997)
 *   ELASTICITY   "(Div_u*(lambda*Id(meshdim))+(2*mu)*Sym(Grad_u)):Grad_Test_u"  params = {lambda, mu}
 *                (add_isotropic_linearized_elasticity_brick, getfem_models.cc:6102-6136)
 *   SVK / NEOHOOKEAN_CIARLET / NEOHOOKEAN_BONET
 *                "((Id(meshdim)+Grad_u)*(<law>_PK2(Grad_u,params))):Grad_Test_u" params = {lambda, mu}
 *                (add_finite_strain_elasticity_brick, getfem_nonlinear_elasticity.cc:2301-2325;
 *                 laws :1945-1994, :612-702)
 *   MASS         "a*u.Test_u"                                        params = {a}
 *   SOURCE       "F.Test_u" (F a constant of qdim components; "-f*Test_u" passes F = -f)   params = {F_0 .. F_{qdim-1}}
 *                (add_source_term_brick, getfem_models.cc:4124-): an order-1 term.  It has no order-2 tree: the
 *                TANGENT bit is accepted and leaves an empty matrix (nnz = 0), as ga_workspace::assembly(2) does.
 *   NORMAL_SOURCE "(Reshape(A,qdim(u),meshdim)*Normal).Test_u" / "((g).Normal)*Test_u": A(b,n) = params[b + qdim*n]
 *                (add_normal_source_term_brick, getfem_models.cc:4280-4299): order 1, boundary-face regions only
 *                (the unit normal exists on faces, C&E.cc:8836-8847).
 * Every family can be integrated over a mesh region (gfgpu_term_set_region): MASS on boundary faces is the Robin /
 * Dirichlet-penalisation matrix, SOURCE on faces the Neumann load.
 */
enum {
  GFGPU_LAPLACE = 0,
  GFGPU_ELASTICITY = 1,
  GFGPU_SVK = 2,
  GFGPU_NEOHOOKEAN_CIARLET = 3,
  GFGPU_NEOHOOKEAN_BONET = 4,
  GFGPU_MASS = 5,
  GFGPU_SOURCE = 6,
  GFGPU_NORMAL_SOURCE = 7,
  /* the other laws of add_finite_strain_elasticity_brick (getfem_nonlinear_elasticity.cc:2276-2290), same expression
   * "((Id(meshdim)+Grad_u)*(<law>_PK2(Grad_u,params))):Grad_Test_u", 3D vector fields:
   *   MOONEY_RIVLIN     Compressible_Mooney_Rivlin_PK2 (:503-607), params (C10, C01, D1)
   *   CIARLET_GEYMONAT  Ciarlet_Geymonat_PK2 (:817-888), params (lambda, mu, a)
   *   BLATZ_KO          Generalized_Blatz_Ko_PK2 (:706-815), params (a, b, c, d, n) */
  GFGPU_MOONEY_RIVLIN = 8,
  GFGPU_CIARLET_GEYMONAT = 9,
  GFGPU_BLATZ_KO = 10,
  GFGPU_JIT = 11 /* gfgpu_term_create_jit: the integrand is given as source and compiled at run time (NVRTC) */
};
#define GFGPU_MAX_PARAMS 12 /* parameters kept per term (NORMAL_SOURCE: qdim x dim <= 9) */
#define GFGPU_MAX_FACES 6   /* faces of a reference element (simplices: dim+1, parallelepipeds: 2*dim) */

/* order_mask bits of gfgpu_term_assemble_*: ga_workspace::assembly(1) and assembly(2)
 * (getfem_generic_assembly_workspace.cc:791-936) */
enum { GFGPU_RESIDUAL = 1, GFGPU_TANGENT = 2 };

/* numeric strategies; AUTO picks RECOMPUTE for affine geometry + constant
 * coefficient bilinear forms and STAGED otherwise */
enum { GFGPU_STRATEGY_AUTO = 0, GFGPU_STRATEGY_STAGED = 1, GFGPU_STRATEGY_RECOMPUTE = 2 };

const char *gfgpu_last_error(void);
/* library/ABI version, and the number of kernels launched by this process so far
 * (bench.py's gpu_launches claim is read from here) */
int gfgpu_version(void);
int64_t gfgpu_launch_count(void);

/* ---- context.  `stream` is a cudaStream_t passed as void* (NULL = the context creates its own). */
int gfgpu_ctx_create(int device, void *stream, gfgpu_ctx **out);
int gfgpu_ctx_destroy(gfgpu_ctx *ctx);
int gfgpu_ctx_synchronize(gfgpu_ctx *ctx);
/* bytes currently allocated by the library on this context */
int64_t gfgpu_ctx_bytes_in_use(gfgpu_ctx *ctx);
/* measured fp64 FMA peak of the device (TFLOP/s, FMA = 2 flops; register-resident DFMA chains on every SM, best of 3
 * after a warm-up): the denominator of the fp64 roofline bench.py reports (MEASURED_PEAKS.json has no fp64 figure) */
int gfgpu_ctx_measure_fp64_peak(gfgpu_ctx *ctx, double *tflops);
/* the same for the fp64 tensor-core path (mma.sync m8n8k4 f64 chains): the evidence behind keeping the sum-factorised
 * Q4 contraction on the FMA pipe */
int gfgpu_ctx_measure_dmma_peak(gfgpu_ctx *ctx, double *tflops);

/* ---- mesh.  Replaces basic_mesh::points_of_convex(cv,G) (getfem/bgeot_mesh.h:94) and
 * mesh_structure::ind_points_of_convex (bgeot_mesh_structure.h:106): points in point-id order
 * (npts x dim, row-major), connectivity in convex order (ne x ng, the geometric nodes in
 * pgt->geometric_nodes() order).  Stored on device as SoA x[],y[],z[] + int32 conn. */
int gfgpu_mesh_create(gfgpu_ctx *ctx, int dim, int64_t npts, const double *pts_host, int64_t ne, int ng,
                      const int32_t *conn_host, int gt_kind, gfgpu_mesh **out);
int gfgpu_mesh_destroy(gfgpu_mesh *m);

/* ---- fem.  Replaces mesh_fem::ind_scalar_basic_dof_of_element (getfem_mesh_fem.h:459-461).
 * elem_dof_host: ne x nd, global dof of COMPONENT 0 of local node i (components are consecutive),
 * or NULL: the first-touch numbering of mesh_fem::enumerate_dof (getfem_mesh_fem.cc:320-446) is
 * then re-derived on the device from the connectivity (classical Lagrange PK/QK of `degree`). */
int gfgpu_fem_create(gfgpu_ctx *ctx, gfgpu_mesh *mesh, int fem_kind, int degree, int qdim, int nd,
                     const int64_t *elem_dof_host, int64_t ndof, gfgpu_fem **out);
int64_t gfgpu_fem_nb_dof(gfgpu_fem *f);
int gfgpu_fem_get_elem_dof(gfgpu_fem *f, int64_t *elem_dof_host /* ne x nd */);
int gfgpu_fem_destroy(gfgpu_fem *f);

/* ---- reference tables at the nq volume quadrature points.  Replaces geotrans_precomp_::grad
 * (bgeot_geometric_trans.h:292-345), fem_precomp_::val/grad (getfem_fem.h:653-680) and
 * approx_integration::coeff (getfem_integration.h:155-222).
 *   w[nq]; gt_grad[nq][ng][dim]; phi[nq][nd]; gphi[nq][nd][dim]   (row-major as written) */
int gfgpu_tables_create(gfgpu_ctx *ctx, int dim, int nq, int ng, int nd, const double *w_host,
                        const double *gt_grad_host, const double *phi_host, const double *gphi_host,
                        gfgpu_tables **out);
/* Tables at the points of the element FACES, for integration over boundary regions.  Replaces the face part of
 * approx_integration (ind_first_point_on_face / nb_points_on_face, getfem_integration.h:170-186: the points of face f
 * follow the volume points, face after face, getfem_integration.cc:353-368) and pgt->normals()
 * (bgeot_geometric_trans.h:141).  nf faces with nqf points each (the classical rules carry the same method on every face):
 *   normals[nf][dim] (reference normals as the reference stores them, not necessarily unit);
 *   w[nf][nqf]; gt_grad[nf][nqf][ng][dim]; phi[nf][nqf][nd]; gphi[nf][nqf][nd][dim] */
/* optional: values of the geometric transformation's shape functions at the volume points, [nq][ng] (geotrans_precomp_::val,
 * bgeot_geometric_trans.h), and -- NULL or, after gfgpu_tables_set_faces, [nf][nqf][ng] -- at the face points: the position
 * X = sum_g G_g N_g(q) of a Gauss point, for JIT integrands that mention X (gmm::mult(G, pgp->val(ii)), C&E.cc ga_instruction_X) */
int gfgpu_tables_set_gt_values(gfgpu_tables *t, const double *gt_val_host, const double *face_gt_val_host);
int gfgpu_tables_set_faces(gfgpu_tables *t, int nf, int nqf, const double *normals_host, const double *w_host,
                           const double *gt_grad_host, const double *phi_host, const double *gphi_host);
int gfgpu_tables_destroy(gfgpu_tables *t);

/* ---- term.  Replaces ga_compile + ga_exec for one expression of a recognised family
 * (C&E.cc:7910-8643, 8750-9047).  alpha multiplies every contribution (factor_of_variable,
 * C&E.cc:5052,8196-8198). */
int gfgpu_term_create(gfgpu_ctx *ctx, gfgpu_mesh *mesh, gfgpu_fem *fem, gfgpu_tables *tab, int family,
                      const double *params_host, int nparams, double alpha, int strategy, gfgpu_term **out);
int gfgpu_term_destroy(gfgpu_term *t);

/* Integrate over a mesh region (getfem::mesh_region, getfem_mesh_region.h; walked by mr_visitor in ga_exec,
 * C&E.cc:8789-8866): n_items items (cv_host[k], face_host[k]) in the visitor's order (ascending convex, then face).
 * face_host == NULL or face -1: the whole convex; face >= 0: that face of the convex -- the weight becomes
 * J * |B n_ref| * w_q and the unit normal B n_ref / |B n_ref| (C&E.cc:8836-8847); needs gfgpu_tables_set_faces.
 * A region is either all convexes or all faces.  Each item is one "element" of the reference's loop: its element
 * matrix goes through the drop rule on its own.  Face regions use strategy STAGED.
 * n_items = 0 with cv_host == NULL removes the region (all convexes again). */
int gfgpu_term_set_region(gfgpu_term *t, int64_t n_items, const int32_t *cv_host, const int32_t *face_host);

/* Fem-data coefficients: the leading `nfields` parameters of the family become FIELDS on a data mesh_fem instead of
 * constants -- ga_workspace::add_fem_constant (generic_assembly.h:465), evaluated at every Gauss point by
 * ga_instruction_val on the data fem (C&E.cc:636-690): "a*Grad_u.Grad_Test_u" with a heterogeneous a(x), lambda(x) / mu(x)
 * of the elasticity brick, the distributed load F(x) of add_source_term_brick (getfem_models.cc:4124-).
 *   LAPLACE, MASS: 1 field (a), scalar data fem; ELASTICITY: 1 or 2 (lambda[, mu]), scalar data fem;
 *   SOURCE: 1 field of qdim components (data fem of the variable's qdim), vals0 already carrying the sign of the expression;
 *   JIT terms: 1 or 2 scalar fields, the integrand's fld[0], fld[1]; or one vector field (data fem of qdim = mesh dimension), vfld.
 *   data_fem: a gfgpu_fem on the same mesh; phi_host[nq][nd_d]: its basis at the volume quadrature points
 *   (fem_precomp_::val); phi_faces_host[nf][nqf][nd_d] (or NULL): the same at the face points, for regions of faces;
 *   vals*_host: nodal values (data_fem ndof).  nfields = 0 removes the fields.  Such terms use strategy STAGED.
 * update_field replaces the nodal values of field k (same data fem), e.g. between load steps. */
int gfgpu_term_set_fields(gfgpu_term *t, int nfields, gfgpu_fem *data_fem, const double *phi_host,
                          const double *phi_faces_host, const double *vals0_host, const double *vals1_host);
int gfgpu_term_update_field(gfgpu_term *t, int k, const double *vals_host);

/* Restrict the term to the block [e0, e1) of its elements -- of its region's items when a region is set -- (per-rank
 * element partition, getfem_mesh_region.cc:145-185).  Default: everything. */
int gfgpu_term_set_element_range(gfgpu_term *t, int64_t e0, int64_t e1);

/* Assemble with the state vector resident on the device (U_dev may be NULL: zero state).
 * Tangent: builds/validates the value-dependent pattern (add_elem_matrix drop rule,
 * C&E.cc:4853-4936, threshold 1e-14*max|K_e| from :5380-5402,:5441-5465) then the
 * deterministic gather-sum.  Residual: ga_instruction_vector_assembly_mf (C&E.cc:4669-4735). */
int gfgpu_term_assemble_dev(gfgpu_term *t, const double *U_dev, int order_mask);
/* Same through host buffers (the drop-in call): copies U host->device, assembles, copies the
 * tangent values (nnz doubles, may be NULL) and the residual (ndof doubles, may be NULL) back. */
int gfgpu_term_assemble_host(gfgpu_term *t, const double *U_host, int order_mask, double *pr_host,
                             double *R_host);

/* Order 0: the scalar ga_workspace::assembly(0) accumulates into assembled_potential() (workspace.cc:791-803,
 * ga_instruction_scalar_assembly, C&E.cc:4628-4640) for the term's POTENTIAL:
 *   LAPLACE / ELASTICITY / MASS   1/2 u^T K u   (the quadratic form whose first variation is the family's residual)
 *   SOURCE / NORMAL_SOURCE        the linear form itself, F.u
 *   finite-strain families        int W(E(Grad_u)) with the law's strain energy ("<law>_potential(Grad_u,params)",
 *                                 AHL_wrapper_potential, getfem_nonlinear_elasticity.cc:1841-1928; 1e200 where det F <= 0)
 * Synchronises the stream; the value comes back through E_host. */
int gfgpu_term_potential_dev(gfgpu_term *t, const double *U_dev, double *E_host);
int gfgpu_term_potential_host(gfgpu_term *t, const double *U_host, double *E_host);

/* JIT terms (the NVRTC route, csrc/jit.cu): a scalar variable -- or a vector variable of the mesh dimension, for which u / tv
 * are vec, gu / tg are mat (m[c][k] = d u_c / d x_k) and the helpers trace, transp, sym, skew, deviator, ddot, outer, mkmat apply --
 * and an integrand that is none of the families above.  After the
 * reference's analysis and symbolic differentiation the order-1 tree of an expression is linear in the test function and the
 * order-2 tree bilinear in (Test, Test2) (ga_exec interprets exactly those trees, C&E.cc:8750-9047); the caller hands them over
 * as C expressions in the identifiers
 *     u (double), gu (vec, Grad_u), X (vec, the position: needs gfgpu_tables_set_gt_values), Normal (vec, the unit outward normal:
 *     needs a region of faces, gfgpu_term_set_region; J and Normal as C&E.cc:8836-8847), par[k] (the term's parameters),
 *     fld[0], fld[1] (scalar fem-data coefficients at the Gauss point: gfgpu_term_set_fields with 1 or 2 fields),
 *     vfld (vec: ONE vector-valued fem-data field of the mesh dimension, e.g. an advection velocity: gfgpu_term_set_fields
 *     with a data fem of qdim = mesh dimension),
 *     tv / tg (Test_u / Grad_Test_u), t2v / t2g (Test2)
 * with the helpers dot(a,b), normsqr(v), gnorm(v), mkvec(a,b,c), sqr, pos_part, neg_part, Heaviside, sign and the CUDA math
 * library.  form1 must be linear in (tv, tg), form2 bilinear in (tv, tg) x (t2v, t2g): the kernel extracts their coefficients
 * with unit probes.  Example, "(1+sqr(u))*Grad_u.Grad_Test_u + sin(u)*Test_u":
 *     form1 = "(1.0+sqr(u))*dot(gu,tg) + sin(u)*tv"
 *     form2 = "(2.0*u*t2v)*dot(gu,tg) + (1.0+sqr(u))*dot(t2g,tg) + cos(u)*t2v*tv"
 * The kernel is compiled on first use for sm_100a; a form that does not compile raises with the NVRTC log.  value_dependent != 0:
 * the keep masks are re-derived at every assembly (the tangent's pattern may move with u).  Output, pattern, gather, regions of
 * convexes, gfgpu_matrix_add_term: as for every other STAGED term.  Not handled: faces, fem-data fields, order 0.
 * gfgpu_jit_check compiles the two forms without a GPU and returns 0 if they compile (the log goes to gfgpu_last_error). */
int gfgpu_term_create_jit(gfgpu_ctx *ctx, gfgpu_mesh *mesh, gfgpu_fem *fem, gfgpu_tables *tab, const char *form1, const char *form2,
                          const double *params, int nparams, double alpha, int value_dependent, gfgpu_term **out);
int gfgpu_jit_check(int dim, int qdim, const char *form1, const char *form2);
/* order 0 of a JIT term: the scalar integrand of the potential in the same identifiers, without test functions
 * (ga_workspace::assembly(0), workspace.cc:791-803); gfgpu_term_potential_* then returns its integral. */
int gfgpu_term_set_jit_potential(gfgpu_term *t, const char *form0);
/* new values of par[] for the next assemblies of a JIT term (constants of the expression may change between calls) */
int gfgpu_term_set_params(gfgpu_term *t, const double *params, int nparams);

/* Device durations (ms, CUDA events on the context's stream) of the kernels of the LAST assemble call:
 * out[0] generic element kernel, out[1] tangent gather-sum (STAGED), out[2] residual gather-sum,
 * out[3] pattern (re)build (0 when reused), out[4] per-nonzero tangent kernel (RECOMPUTE), out[5..7] 0.
 * Synchronises the stream. */
int gfgpu_term_last_timings(gfgpu_term *t, float *out8);
/* the strategy the term actually uses (GFGPU_STRATEGY_STAGED or GFGPU_STRATEGY_RECOMPUTE) */
int gfgpu_term_strategy(gfgpu_term *t);
/* which tangent kernel the last plan of the term selected: 0 = generic element kernel + gather-sum (STAGED, or RECOMPUTE not
 * planned yet), 1 = general per-nonzero tile kernel, 2 = column kernel (low-order scalar forms), 3 = class-uniform tile kernel
 * (meshes with translated structure), 4 = STAGED in direct mode (scalar sum-factorised element kernel under a fixed pattern:
 * entries go straight from the element kernel to their CSC slots, no element matrix in HBM).  Diagnostic: the choice never
 * changes results beyond round-off. */
int gfgpu_term_kernel_kind(gfgpu_term *t);

int64_t gfgpu_term_nnz(gfgpu_term *t);
int64_t gfgpu_term_nb_dof(gfgpu_term *t);
/* number of times the pattern has been (re)built; a Newton loop can watch it */
int64_t gfgpu_term_pattern_generation(gfgpu_term *t);
/* device views (valid until the next assemble/destroy): gmm::csc_matrix layout */
int gfgpu_term_csc_view(gfgpu_term *t, const int64_t **jc_dev, const int32_t **ir_dev, const double **pr_dev);
int gfgpu_term_residual_view(gfgpu_term *t, const double **R_dev);
/* host export of the pattern / values / residual (any pointer may be NULL) */
int gfgpu_term_export_csc_host(gfgpu_term *t, int64_t *jc_host, int32_t *ir_host, double *pr_host);
int gfgpu_term_export_residual_host(gfgpu_term *t, double *R_host);

/* ---- workspace-level tangent on the device.  ga_workspace::assembly(2) adds every order-2 tree
 * into ONE gmm::col_matrix<rsvector> at the variables' intervals (getfem_generic_assembly_workspace.cc:791-936;
 * add_elem_matrix C&E.cc:4853-4936); a model sums brick matrices into its tangent and forms the residual of linear bricks
 * as K*u (getfem_models.cc:2536-2620, 2753-2900).  gfgpu_matrix is that container, kept in HBM:
 *   add_term   K(row_off + r, col_off + c) += alpha * term(r, c) over the term's stored entries.  The pattern is the UNION
 *              of what was added (an entry stays stored even if a later term cancels it, like rsvector); values are summed
 *              in the order of the calls; no atomics, bitwise reproducible.  The term must have an assembled tangent.
 *   clear      keep_pattern != 0: zero the values (next Newton iteration); 0: forget the pattern as well
 *   mult       y = beta*y + alpha*K x (transposed == 0) or alpha*K^T x (transposed != 0), fixed summation order
 *   export / view: gmm::csc_matrix layout (gmm_matrix.h:545-566), like the term-level calls. */
int gfgpu_matrix_create(gfgpu_ctx *ctx, int64_t nrows, int64_t ncols, gfgpu_matrix **out);
int gfgpu_matrix_destroy(gfgpu_matrix *m);
int gfgpu_matrix_clear(gfgpu_matrix *m, int keep_pattern);
int gfgpu_matrix_add_term(gfgpu_matrix *m, gfgpu_term *t, double alpha, int64_t row_off, int64_t col_off);
int64_t gfgpu_matrix_nnz(gfgpu_matrix *m);
int64_t gfgpu_matrix_pattern_generation(gfgpu_matrix *m);
int gfgpu_matrix_csc_view(gfgpu_matrix *m, const int64_t **jc_dev, const int32_t **ir_dev, const double **pr_dev);
int gfgpu_matrix_export_csc_host(gfgpu_matrix *m, int64_t *jc_host, int32_t *ir_host, double *pr_host);
int gfgpu_matrix_mult_dev(gfgpu_matrix *m, int transposed, double alpha, const double *x_dev, double beta, double *y_dev);
int gfgpu_matrix_mult_host(gfgpu_matrix *m, int transposed, double alpha, const double *x_host, double beta, double *y_host);
/* Model-level algebra on the resident tangent, so that assemble -> constrain -> solve needs no copy of K to the host.
 *   apply_dof_constraints  Dirichlet conditions "with simplification" (model::real_dof_constraints,
 *              getfem_models.cc:2806-2871) for the n dofs dof_host[] with prescribed values go_host[]:
 *              rhs_dev != NULL (BUILD_RHS): linear + symmetric model: rhs -= K(:, SI) go (with K as it is before the
 *              clearing, refused without GFGPU_BUILD_MATRIX like models.cc:2843-2845), then rhs[SI] = go; nonlinear
 *              model: rhs[SI] += go - pr_host (present values of those dofs);
 *              GFGPU_BUILD_MATRIX: rows SI cleared, columns SI too for a symmetric model, K(i, i) = 1.  The entries stay
 *              in the pattern as explicit zeros (the pattern generation moves only if a diagonal slot was missing).
 *   export_csr_dev  hand-off to a device solver: row pointers (nrows + 1), columns and values in row-major order, ascending
 *              columns inside a row, written into caller-owned DEVICE buffers (any may be NULL); needs nnz < 2^31.
 *   cg_dev     Jacobi-preconditioned conjugate gradient for a symmetric positive definite K: x_dev holds the initial guess
 *              and receives the solution, stops at |r| <= rtol |b| or max_iter; deterministic reductions.
 *   gfgpu_term_residual_add_dev  rhs_dev[row_off + i] += alpha R_i: the model's rrhs accumulated on the device
 *              (getfem_models.cc:2553-2570). */
enum { GFGPU_MODEL_LINEAR = 1, GFGPU_MODEL_SYMMETRIC = 2, GFGPU_BUILD_MATRIX = 4 };
int gfgpu_matrix_apply_dof_constraints(gfgpu_matrix *m, int64_t n, const int64_t *dof_host, const double *go_host,
                                       const double *pr_host, double *rhs_dev, int flags);
int gfgpu_matrix_export_csr_dev(gfgpu_matrix *m, int64_t *rowptr_dev, int32_t *col_dev, double *val_dev);
int gfgpu_matrix_cg_dev(gfgpu_matrix *m, const double *b_dev, double *x_dev, double rtol, int max_iter, int *iters_out,
                        double *relres_out);
int gfgpu_term_residual_add_dev(gfgpu_term *t, double alpha, double *rhs_dev, int64_t row_off);
/* y = beta y + alpha K^T x on the term's own CSC (warp per column, fixed order): for the symmetric families K^T x = K x, which is
 * how linear bricks form their residual (getfem_models.cc:2536-2620); bench.py uses it for its full-size property checks
 * (R = K U, K t = 0 for rigid translations, x2.K x1 = x1.K x2). */
int gfgpu_term_tmult_dev(gfgpu_term *t, double alpha, const double *x_dev, double beta, double *y_dev);

/* ---- reduced mesh_fem (mesh_fem::is_reduced(): partial_mesh_fem -- the multiplier spaces of the Dirichlet bricks --, periodic
 * or enriched spaces).  The reference assembles such a variable on its BASIC dofs into unreduced containers and projects them
 * with the extension matrix E (nb_basic_dof x nb_dof, mesh_fem::extension_matrix()): K(I1, I2) += E1^T K_basic E2 and
 * V(I1) += E1^T V_basic; the state of the variable is extended first, U_basic = E U (workspace.cc:861-935,
 * getfem_mesh_fem.h extend_vector).  Device terms are therefore created on the basic dof table (gfgpu_fem_create with
 * ndof = nb_basic_dof) and their results pass through a gfgpu_reduction:
 *   create            E in CSR: rowptr[n_basic + 1], ascending columns inside a row (gmm::csr_matrix layout)
 *   extend            y[n_basic] = E x[n_dof]
 *   restrict_add      y[n_dof] += alpha E^T x[n_basic]              (ordered gather, no atomics)
 *   gfgpu_matrix_add_term_reduced / add_rect_reduced
 *                     K(row_off.., col_off..) += alpha Er^T S Ec, one expand-sort-compress pass per reduced side, products
 *                     summed in ascending inner index (the order of gmm's sparse products), exact zeros not stored
 *                     (rsvector::w); a NULL extension matrix = that side is not reduced. */
typedef struct gfgpu_reduction gfgpu_reduction;
int gfgpu_reduction_create(gfgpu_ctx *ctx, int64_t n_basic, int64_t n_dof, const int64_t *rowptr_host, const int32_t *col_host,
                           const double *val_host, gfgpu_reduction **out);
int gfgpu_reduction_destroy(gfgpu_reduction *E);
int gfgpu_reduction_extend_host(gfgpu_reduction *E, const double *x_host, double *y_host);
int gfgpu_reduction_restrict_add_host(gfgpu_reduction *E, double alpha, const double *x_host, double *y_host);
int gfgpu_matrix_add_term_reduced(gfgpu_matrix *m, gfgpu_term *t, gfgpu_reduction *E, double alpha, int64_t row_off, int64_t col_off);

/* ---- coupled bilinear terms: Test on one fem (rows), Test2 on another (columns) -- the off-diagonal blocks of mixed
 * formulations.  ga_workspace::assembly(2) adds an order-2 tree whose test functions belong to two variables into the block
 * (interval of Test's variable) x (interval of Test2's variable), element matrix by element matrix through add_elem_matrix and
 * its drop rule (C&E.cc:4853-4936, 5380-5402); the incompressibility bricks add "-p*Div_Test_u - Test_p*Div_u"
 * (getfem_models.cc, add_linear_incompressibility).
 *   create      both fems on `mesh`, both table sets at the SAME quadrature points (gfgpu_tables_create twice with the same
 *               w / gt_grad).  GFGPU_RECT_DIV_PRESSURE: block(row (i,a), column j) = alpha * coef * int psi_j d(phi_i)/dx_a,
 *               rows on a vector fem (qdim = mesh dimension), columns on a scalar fem.
 *   assemble    element matrices, drop rule per element matrix, pattern on the first call (an entry exists iff one of its
 *               contributions is kept), ordered sums; bitwise reproducible.
 *   export      gmm::csc_matrix layout of the block (transposed == 0: nrows = row fem dofs) or of its transpose (the tree
 *               with the test functions swapped).
 *   mult        y = beta y + alpha B x (transposed == 0) or alpha B^T x: the residual parts R_u = B p, R_p = B^T u.
 *   gfgpu_matrix_add_rect   K(row_off.., col_off..) += alpha * block (or its transpose), union pattern like add_term. */
enum { GFGPU_RECT_DIV_PRESSURE = 0,
       /* block((i,a), (j,b)) = alpha * coef * delta_ab int phi_i psi_j: "Test_u1:Test2_u2" on two fems of the same qdim --
        * asm_mass_matrix(M, mim, mf1, mf2, rg) (getfem_assembling.h:743-755), the constraint matrix of the Dirichlet bricks
        * with multipliers (getfem_models.cc:4386-4421), usually on a region of faces */
       GFGPU_RECT_MASS = 1 };
typedef struct gfgpu_rect gfgpu_rect;
int gfgpu_rect_create(gfgpu_ctx *ctx, gfgpu_mesh *mesh, gfgpu_fem *fem_rows, gfgpu_tables *tab_rows, gfgpu_fem *fem_cols,
                      gfgpu_tables *tab_cols, int family, double coef, double alpha, gfgpu_rect **out);
/* restricts the coupled term to a mesh region: items in mr_visitor order, face_host NULL / -1 for convexes, or all faces (both
 * table sets then need gfgpu_tables_set_faces); cv_host NULL = back to all convexes.  Same meaning as gfgpu_term_set_region. */
int gfgpu_rect_set_region(gfgpu_rect *r, int64_t n_items, const int32_t *cv_host, const int32_t *face_host);
int gfgpu_rect_destroy(gfgpu_rect *r);
int gfgpu_rect_assemble_dev(gfgpu_rect *r);
int64_t gfgpu_rect_nnz(gfgpu_rect *r);
int gfgpu_rect_export_csc_host(gfgpu_rect *r, int transposed, int64_t *jc_host, int32_t *ir_host, double *pr_host);
int gfgpu_rect_mult_dev(gfgpu_rect *r, int transposed, double alpha, const double *x_dev, double beta, double *y_dev);
int gfgpu_rect_mult_host(gfgpu_rect *r, int transposed, double alpha, const double *x_host, double beta, double *y_host);
int gfgpu_matrix_add_rect(gfgpu_matrix *m, gfgpu_rect *r, int transposed, double alpha, int64_t row_off, int64_t col_off);
int gfgpu_matrix_add_rect_reduced(gfgpu_matrix *m, gfgpu_rect *r, int transposed, gfgpu_reduction *E_rows, gfgpu_reduction *E_cols,
                                  double alpha, int64_t row_off, int64_t col_off);

/* ---- multi-GPU: element blocks per rank, column-owned CSC slabs, one halo exchange per assembly.
 * Replaces the reference's MPI scheme (per-rank partial matrices summed with MPI_SUM_SPARSE_MATRIX /
 * MPI_SUM_VECTOR, getfem_generic_assembly_workspace.cc:855-858, getfem_models.cc:586,2572,2608) by owned slabs:
 * dofs are owned by the rank whose element block touches them first (first-touch numbering, getfem_mesh_fem.cc:
 * 320-446, makes these contiguous ranges [own_lo, own_hi)); a rank's columns below own_lo are ghosts.
 *   symbolic, once (host mediated, any transport):
 *     halo_begin        local structure + pattern of the element block; returns the dof range it touches
 *     halo_ghost_pairs  the (column node J, row node I, keep mask) pairs of MY ghost columns in [dof_lo, dof_hi),
 *                       sorted by (J, I), to be sent to the owner of that range (call with NULL arrays for n)
 *     halo_add_source   pairs announced BY rank src for columns I own (ascending src), and the dof range of
 *                       the residual slice it will send
 *     halo_commit       declares the owned range; the next assemble rebuilds the pattern with the announced
 *                       pairs merged in (reference drop rule = OR of the keep masks)
 *   numeric, every assembly (device pointers, e.g. ncclSend / ncclRecv on them):
 *     assemble_dev      local contributions; owned columns are written in the merged layout
 *     halo_send_view    my partial values of the ghost columns [dof_lo, dof_hi): a contiguous slice of pr,
 *                       and the matching slice of the residual
 *     halo_recv_view    where the values of source rank src must land
 *     halo_accumulate   owner adds the received parts in ascending source rank (fixed order, no atomics)
 * After halo_accumulate the columns [own_lo, own_hi) of the CSC view and R[own_lo, own_hi) are complete. */
int gfgpu_term_halo_begin(gfgpu_term *t, const double *U_dev, int64_t *touched_lo, int64_t *touched_hi);
int gfgpu_term_halo_ghost_pairs(gfgpu_term *t, int64_t dof_lo, int64_t dof_hi, int64_t *n, int32_t *J_host,
                                int32_t *I_host, uint16_t *mask_host);
int gfgpu_term_halo_add_source(gfgpu_term *t, int src_rank, int64_t n, const int32_t *J_host, const int32_t *I_host,
                               const uint16_t *mask_host, int64_t r_lo, int64_t r_hi);
int gfgpu_term_halo_commit(gfgpu_term *t, int64_t own_lo, int64_t own_hi);
int gfgpu_term_halo_send_view(gfgpu_term *t, int64_t dof_lo, int64_t dof_hi, const double **pr_dev, int64_t *count,
                              const double **R_dev);
int gfgpu_term_halo_recv_view(gfgpu_term *t, int src_rank, double **pr_recv_dev, int64_t *count, double **R_recv_dev,
                              int64_t *r_count);
int gfgpu_term_halo_accumulate(gfgpu_term *t, int order_mask);
int gfgpu_term_owned_range(gfgpu_term *t, int64_t *own_lo, int64_t *own_hi);

/* ---- the exchange inside the library: NCCL over NVLink, one communicator per process (one process per GPU).
 * Replaces MPI_SUM_SPARSE_MATRIX / MPI_SUM_VECTOR of src/getfem/getfem_config.h:214-341 (used by
 * getfem_generic_assembly_workspace.cc:855-858) with point-to-point slices between neighbouring element blocks.
 *   gfgpu_comm_unique_id   rank 0 creates the rendezvous id (GFGPU_COMM_ID_BYTES bytes) and hands it to the other ranks by any
 *                          means (MPI_Bcast, torch.distributed, a file): ncclGetUniqueId
 *   gfgpu_comm_create      collective over the ranks: ncclCommInitRank on the context's device
 *   gfgpu_term_halo_add_send  after halo_commit: this rank's ghost columns [dof_lo, dof_hi) belong to rank `owner`; it will
 *                          send them and the residual slice [r_lo, dof_hi) at every exchange
 *   gfgpu_term_halo_exchange  after gfgpu_term_assemble_dev, on the same stream, asynchronous: ONE ncclGroup with the sends to
 *                          the owners and the receives from the sources, then halo_accumulate.  The owned slab and residual
 *                          slice are complete when the stream reaches the end of it.
 * NCCL is loaded at run time (libnccl.so.2); without it these calls fail with a message, nothing else is affected. */
#define GFGPU_COMM_ID_BYTES 128
typedef struct gfgpu_comm gfgpu_comm;
int gfgpu_comm_unique_id(char *id_out, int capacity);
int gfgpu_comm_create(gfgpu_ctx *ctx, int nranks, int rank, const char *id, gfgpu_comm **out);
int gfgpu_comm_destroy(gfgpu_comm *c);
int gfgpu_comm_rank(gfgpu_comm *c);
int gfgpu_comm_size(gfgpu_comm *c);
int gfgpu_term_halo_add_send(gfgpu_term *t, int owner_rank, int64_t dof_lo, int64_t dof_hi, int64_t r_lo);
int gfgpu_term_halo_exchange(gfgpu_term *t, gfgpu_comm *c, int order_mask);

#ifdef __cplusplus
}
#endif
#endif /* GFGPU_H */
