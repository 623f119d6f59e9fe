/* bloch_b200.h - C ABI of the B200-native Bloch Maxwell eigen path.
 *
 * The reference (mlstowell/mfem-bravais) has no FFI; its seam for this path is the C++ class
 * mfem::bloch::MaxwellBlochWaveEquation (maxwell/maxwell_bloch.hpp:140-234).  Each entry
 * point below names the member function it replaces.  All arrays are caller-allocated, plain
 * pointers + sizes, no exceptions cross the boundary.  Every function returns 0 on success
 * and a negative code on failure; bloch_last_error() gives the message (thread-local).
 *
 * Vector layout at the boundary is the reference's: one real vector of length 2N per complex
 * field, [re(N); im(N)] (block offsets [0,N,2N], maxwell_bloch.cpp:112-120); nvec vectors
 * are stored one after the other (vector v starts at x + v*2N).
 *
 * One handle = one eigenproblem object bound to one CUDA device and stream.  A handle is
 * not thread-safe; different handles are fully independent (the k-point sweep relies on it).
 * There is no CPU fallback: bloch_create fails if no CUDA device is usable.
 */
#ifndef BLOCH_B200_H
#define BLOCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bloch_handle_s *bloch_handle;

/* error codes */
#define BLOCH_OK 0
#define BLOCH_ERR_ARG (-1)
#define BLOCH_ERR_CUDA (-2)
#define BLOCH_ERR_STATE (-3)
#define BLOCH_ERR_NOCONV (-4)
#define BLOCH_ERR_INTERNAL (-5)

/* lattice types: values of bravais::BRAVAIS_LATTICE_TYPE (lib/bravais.hpp:24-50) */
#define BLOCH_LATTICE_CUB 7
#define BLOCH_LATTICE_FCC 8
#define BLOCH_LATTICE_BCC 9
#define BLOCH_LATTICE_HEX 16   /* PRIMITIVE_HEXAGONAL_PRISM, parameters a and c */

const char *bloch_last_error(void);
int bloch_version(void);
/* number of usable CUDA devices (0 if none); never fails */
int bloch_device_count(void);

/* ---- lattice / k-path API (lib/bravais.hpp:64-181, factory :1175-1239) ---- */
typedef struct bloch_lattice_s *bloch_lattice;
int bloch_lattice_create(bloch_lattice *out, int lattice_type, double a, double b, double c,
                         double alpha, double beta, double gamma);     /* BravaisLatticeFactory */
int bloch_lattice_destroy(bloch_lattice lat);
int bloch_lattice_label(bloch_lattice lat, char *buf, int buflen);     /* GetLatticeTypeLabel */
double bloch_lattice_volume(bloch_lattice lat);                         /* GetUnitCellVolume */
int bloch_lattice_vectors(bloch_lattice lat, double lat9[9], double rec9[9]); /* Get(Reciprocal)LatticeVectors */
int bloch_lattice_num_translations(bloch_lattice lat);
int bloch_lattice_translations(bloch_lattice lat, double *trn, double *face_radii); /* GetTranslationVectors, GetFaceRadii */
int bloch_lattice_num_symmetry_points(bloch_lattice lat);               /* GetNumberSymmetryPoints */
int bloch_lattice_symmetry_point(bloch_lattice lat, int i, double kappa[3], char *label, int buflen); /* GetSymmetryPoint(+Label): 2 pi sp */
int bloch_lattice_symmetry_point_index(bloch_lattice lat, const char *label); /* GetSymmetryPointIndex; -1 if unknown */
int bloch_lattice_num_paths(bloch_lattice lat);                         /* GetNumberPaths */
int bloch_lattice_num_path_segments(bloch_lattice lat, int p);          /* GetNumberPathSegments */
int bloch_lattice_path_segment(bloch_lattice lat, int p, int s, int *e0, int *e1); /* GetPathSegmentEndPointIndices */
int bloch_lattice_intermediate_point(bloch_lattice lat, int p, int s, double kappa[3], char *label, int buflen); /* GetIntermediatePoint(+Label) */
int bloch_lattice_map_to_primitive_cell(bloch_lattice lat, const double pt[3], double ipt[3]); /* MapToPrimitiveCell; returns 1 if mapped */

/* ---- eigenproblem object ---- */

/* MaxwellBlochWaveEquation(ParMesh&, int order) (maxwell_bloch.hpp:143, .cpp:34-140) on the
 * periodic Wigner-Seitz hex mesh of `lat` with every coarse hex subdivided n_sub^3 times
 * (n_sub = 2^r reproduces r uniform refinements, maxwell_dispersion.cpp:366-388).
 * device == -1 selects the current CUDA device.  device == BLOCH_DEVICE_NONE builds a
 * topology-only handle (mesh, dof maps and sizes can be queried without a GPU; every compute
 * entry point fails with BLOCH_ERR_ARG) - used by the host-logic tests, never a fallback. */
#define BLOCH_DEVICE_NONE (-2)
int bloch_create(bloch_handle *out, bloch_lattice lat, int n_sub, int order, int device);
/* same, for an arbitrary periodic mesh of parallelepiped hexes: vertices xyz[3*n_vert], hexes
 * hex[8*n_hex] in MFEM vertex order, each subdivided n_sub^3; rec9 = reciprocal vectors (rows) */
int bloch_create_from_hexes(bloch_handle *out, int n_vert, const double *xyz, int n_hex,
                            const int *hex, const double rec9[9], int n_sub, int order, int device);
int bloch_destroy(bloch_handle h);                                      /* ~MaxwellBlochWaveEquation */
/* run on an externally owned stream (a cudaStream_t cast to void*); NULL = handle's own stream */
int bloch_set_stream(bloch_handle h, void *cuda_stream);

/* sizes: GetHCurlFESpace()->GlobalTrueVSize() etc. (maxwell_bloch.hpp:205-206) */
int bloch_num_elements(bloch_handle h, int64_t *n_elem, int *n_class);
int bloch_num_dofs(bloch_handle h, int64_t *n_nd, int64_t *n_rt, int64_t *n_h1);
int bloch_mesh_counts(bloch_handle h, int64_t *n_vert, int64_t *n_edge, int64_t *n_face, double *volume);
int bloch_element_centers(bloch_handle h, double *xyz /* 3*n_elem */);
/* geometry of the affine elements: x0[3*n_elem], cls[n_elem], J[9*n_class] (row-major dx_i/dxhat_j) */
int bloch_element_geometry(bloch_handle h, double *x0, int *cls, double *J);
/* element -> signed 1-based global dof ids (+-(gid+1)) in the natural local order documented
 * in DESIGN.md; space: 0 = H1, 1 = ND, 2 = RT.  out has n_elem * local_size entries. */
int bloch_local_size(bloch_handle h, int space);
int bloch_get_dofmap(bloch_handle h, int space, int32_t *out);

/* SetMassCoef / SetStiffnessCoef (maxwell_bloch.hpp:164-165): one value per element, copied */
int bloch_set_eps(bloch_handle h, const double *eps_per_elem);
int bloch_set_muinv(bloch_handle h, const double *muinv_per_elem);
/* SetKappa (maxwell_bloch.hpp:152; beta = |kappa|, zeta = kappa/beta, .cpp:200-210) */
int bloch_set_kappa(bloch_handle h, const double kappa[3]);
/* k-point batch: nk Bloch vectors kappa[3*nk] are set at once and the next bloch_solve iterates their nk
 * INDEPENDENT eigenproblems together - same mesh, maps and coefficients, one set of kernel launches on block
 * vectors of nk * block columns (the k-loop of maxwell_dispersion.cpp:475-531 run nk points at a time; results are
 * those of nk separate GetEigenvalues calls).  Needs an even n_sub (multigrid projector).  bloch_set_kappa is the
 * batch of one.  bloch_select_kpoint chooses which k-point of the batch the getters below (eigenvalues,
 * eigenvectors, stats, field averages) refer to; bloch_solve resets the selection to 0.  Operator applies on a
 * batched handle take nvec = nk * c vectors, vector v using kappa[v / c]. */
#define BLOCH_MAX_BATCH 16
int bloch_set_kappa_batch(bloch_handle h, int nk, const double *kappa /* [nk][3] */);
int bloch_batch_size(bloch_handle h);
int bloch_select_kpoint(bloch_handle h, int k);
/* SetNumEigs counts REAL modes in the reference (2 per complex band); this takes complex bands */
int bloch_set_num_bands(bloch_handle h, int n_complex_bands);
/* SetAbsoluteTolerance (default 1e-6, .cpp:53) and lobpcg_->SetMaxIter(2000) (.cpp:543) */
int bloch_set_tol(bloch_handle h, double abs_tol, int max_iter);
/* Setup() (maxwell_bloch.hpp:167, .cpp:337-620): uploads the kappa-dependent class tables and
 * (re)builds the preconditioner data.  Idempotent; only what changed is rebuilt. */
int bloch_setup(bloch_handle h);
/* SetInitialVectors (maxwell_bloch.hpp:169): m vectors of length 2N [re;im]; NULL or m == 0
 * restores the built-in initial guess */
int bloch_set_initial_vectors(bloch_handle h, int m, const double *vecs);
/* Solve() (maxwell_bloch.hpp:176, .cpp:809-825) */
int bloch_solve(bloch_handle h);
/* GetEigenvalues (maxwell_bloch.hpp:179, .cpp:1052-1076): ascending lambda = omega^2, one per
 * complex band (the C++ wrapper duplicates them to the reference's real-mode count) */
int bloch_get_eigenvalues(bloch_handle h, double *lambda, int n);
/* GetEigenvectorE / GetEigenvectorB (maxwell_bloch.hpp:187-197, .cpp:1371-1458): copy-out */
int bloch_get_eigenvector_E(bloch_handle h, int i, double *re, double *im);
int bloch_get_eigenvector_B(bloch_handle h, int i, double *re, double *im);

typedef struct {
  int iterations;          /* outer LOBPCG iterations of the last solve */
  int converged_bands;
  int inner_iterations;    /* total projector CG iterations */
  double solve_seconds;    /* device time of the last Solve() */
  double max_residual;     /* max_j || A x_j - lambda_j M x_j ||_2 over the wanted bands */
  int64_t applies_A;       /* number of single-vector operator applications */
  int64_t kernel_launches; /* kernels launched by this handle since creation */
} bloch_stats;
int bloch_get_stats(bloch_handle h, bloch_stats *st);                   /* GetSolverStats */

/* Measurement hook (GetSolverStats has only wall times, maxwell_bloch.hpp:232-234): with profiling on, the next
 * bloch_solve brackets its phases with CUDA events on the handle's stream.  ms[0] whole solve, [1] ND operator-apply
 * kernels outside the preconditioner (incl. clearing y), [2] divergence projector (S0 solves + G / G^H M applies),
 * [3] Chebyshev preconditioner (its ND applies are [7]), [4] Gram + Rayleigh-Ritz rotation kernels, [5] host
 * Rayleigh-Ritz incl. the copies around it (wall clock), [6] extra work of the lifted operator (V-cycle, G, G^H M, M). */
int bloch_set_profile(bloch_handle h, int on);
int bloch_get_profile(bloch_handle h, double *ms, int n /* <= 8 */);

/* GetAOperator()/GetMOperator()->Mult (maxwell_bloch.hpp:199-200) on nvec vectors of length
 * 2N [re;im], HOST pointers (copies inside) */
int bloch_apply_A(bloch_handle h, const double *x, double *y, int nvec);
int bloch_apply_M(bloch_handle h, const double *x, double *y, int nvec);
/* GetSubSpaceProjector()->Mult (maxwell_bloch.hpp:203, .cpp:2280-2290) */
int bloch_apply_projector(bloch_handle h, const double *x, double *y, int nvec);
/* C = [[T12, beta Z12], [-beta Z12, T12]] apply (maxwell_bloch.cpp:471-490): ND 2N -> RT 2N_rt */
int bloch_apply_C(bloch_handle h, const double *x, double *y, int nvec);

/* Device-resident variants: x, y are DEVICE pointers to the handle's internal block layout,
 * interleaved complex [N][nvec] (element (dof i, vector v) at 2*(i*nvec+v), re then im).
 * They launch on the handle's stream and do not synchronise. */
int bloch_apply_A_device(bloch_handle h, const double *d_x, double *d_y, int nvec);
int bloch_apply_M_device(bloch_handle h, const double *d_x, double *d_y, int nvec);
/* layout conversion helpers between the boundary layout and the block layout (device ptrs) */
int bloch_pack_device(bloch_handle h, const double *d_reim, double *d_block, int nvec);
int bloch_unpack_device(bloch_handle h, const double *d_block, double *d_reim, int nvec);

/* Host-side dense solver of the Rayleigh-Ritz step (csrc/dense.hpp; the reference: hypre LOBPCG's LAPACK calls /
 * dsygv, meta_material_solver.cpp:3285): lowest m eigenpairs of GA c = lambda GM c, n x n Hermitian row-major
 * interleaved (re, im); c_reim is n x m.  values_only != 0: the pivoted-Cholesky variant of the reduced-basis sweep
 * (tolerates nearly dependent bases).  Test hook - needs no GPU. */
int bloch_debug_hegv(int n, int m, const double *ga_reim, const double *gm_reim, double *lambda, double *c_reim,
                     int values_only);

/* The same dense problem solved by the DEVICE Rayleigh-Ritz kernel the eigensolver uses (csrc/rr_device.cu: one CTA
 * per pencil, Cholesky + parallel Jacobi in shared memory): nk pencils of size n <= 63 stored one after the other,
 * lowest m <= 32 pairs each; act (nk x m bytes, may be NULL = all) and use_p select the basis columns like the solver
 * does (column i >= m takes part only if act[i % m]; columns >= 2m only if use_p); info[k] = 0 ok, 1 = basis shrunk,
 * -1 = failed.  Test hook: parity against bloch_debug_hegv. */
int bloch_debug_hegv_device(int n, int m, int nk, const double *ga_reim, const double *gm_reim, const unsigned char *act,
                            int use_p, double *lambda, double *c_reim, int *info);

/* Assembled operators for interchange (the reference's -wm dump of Ar / Ai / M, maxwell_dispersion.cpp:553-590).
 * which = 0: A = S1 - i beta DKZ (complex Hermitian; Re = the reference's Ar block, Im = its (1,0) block with the
 * coefficient folded in), which = 1: M = M1(eps) (real).  CSR over the ND dofs of bloch_get_dofmap numbering.
 * Two calls: assemble (returns nnz), then copy out (im may be NULL).  Host-side merge - a debug path. */
int bloch_assemble_matrix(bloch_handle h, int which, int64_t *nnz);
int bloch_get_matrix(bloch_handle h, int64_t *rowptr, int32_t *col, double *re, double *im);

/* Multilevel warm start (MaxwellBlochWaveSolver::GetEigenfrequencies, meta-material/meta_material_solver.cpp:
 * 2731-2881): `fine` must be the uniform refinement of `coarse` (n_sub doubled, same cell, same order, same
 * device).  Interpolates the coarse handle's eigenvectors onto the fine mesh (the reference's
 * GetUpdateOperator()->Mult, :2829-2849) and installs them as the fine handle's starting block, so the next
 * bloch_solve(fine) starts from them (the reference's SetInitialVectors, :2853). */
int bloch_prolong_eigenvectors(bloch_handle coarse, bloch_handle fine);

/* GetFieldAverages (maxwell/maxwell_bloch.cpp:1550-1632): cell integrals of the full Bloch fields
 * e^{i kappa.x}(Er + i Ei) etc. of band i, out24 = Er[3], Ei[3], Br[3], Bi[3], Dr[3], Di[3], Hr[3], Hi[3]
 * (D = eps E, H = mu^-1 B, B as returned by bloch_get_eigenvector_B; not divided by the cell volume, like
 * the reference).  Quadrature: p + 1 Gauss points per direction (MFEM's default for these linear forms). */
int bloch_get_field_averages(bloch_handle h, int i, double out24[24]);

/* ---- reduced-basis k-sweep: MaxwellDispersion::buildRawBasis / approxEigenfrequencies
 * (meta-material/meta_material_solver.cpp:3132-3305).  Full solves are done only at symmetry (and mid)
 * points and their eigenvectors appended to a raw basis kept on the device; bloch_rb_approx then, for any
 * kappa, projects that basis with the kappa's divergence projector, forms the reduced Gram matrices
 * <A p_i, p_j>, <M p_i, p_j> and solves the small dense pencil (the reference: LAPACK dsygv, :3285-3299).
 * Returns APPROXIMATE eigenvalues (ascending), exact at the k-points whose eigenvectors are in the basis. */
int bloch_rb_clear(bloch_handle h);
int bloch_rb_append(bloch_handle h);        /* append the bands of the last bloch_solve */
int bloch_rb_size(bloch_handle h);
int bloch_rb_approx(bloch_handle h, const double kappa[3], double *lambda, int n);

/* ---- scalar variant: ScalarFloquetWaveEquation of misc/scalar3d.cpp:662-818 ----
 * (G - i Z_kappa)^T M1(k) (G - i Z_kappa) u = lambda M0(m) u on H1_p (orders 1..4), same handle, same mesh.
 * kappa is set with bloch_set_kappa; the reference's phase shift beta is in DEGREES:
 * kappa = (beta * pi / 180) * zeta (scalar3d.cpp:733,784-785).  Coefficients: one value per element
 * (stiffness_coef / mass_coef, scalar3d.cpp:564-588).  No null space, hence no projector. */
int bloch_scalar_set_coefs(bloch_handle h, const double *stiffness_k, const double *mass_m);
int bloch_scalar_set_num_modes(bloch_handle h, int n_complex_modes);      /* lobpcg nev (default 5, :427) */
int bloch_scalar_solve(bloch_handle h);
int bloch_scalar_get_eigenvalues(bloch_handle h, double *lambda, int n);
/* which = 0: y = A u (block [[S0, b DKZ],[-b DKZ, S0]]), 1: y = M u (diag(M0, M0)); vectors [re;im] of length 2 N_h1 */
int bloch_scalar_apply(bloch_handle h, int which, const double *x, double *y, int nvec);

/* Test hook for the H1 <-> ND operators inside the projector (MaxwellBlochWaveProjector::Setup,
 * maxwell_bloch.cpp:2053-2158), host vectors [re;im]:
 *   mode 0: y(2 N_h1) = S0 x(2 N_h1), S0 = G^T M G;  mode 1: y(2 N) = G x(2 N_h1);
 *   mode 2: y(2 N_h1) = G^T M x(2 N) */
int bloch_debug_apply_h1op(bloch_handle h, int mode, const double *x, double *y, int nvec);
/* Test hook for the auxiliary-space part of the preconditioner (what HypreAMS contributes in the reference,
 * maxwell_bloch.cpp:492-517): (H1)^3 vectors hold the Cartesian components as blocks, [re(3 N_h1); im(3 N_h1)].
 *   mode 0: y(2 N) = Pi u   (nodal interpolation (H1)^3 -> ND);  mode 1: y(2 * 3 N_h1) = Pi^T x(2 N);
 *   mode 2: y = B u, one multigrid V-cycle per component for (grad + i kappa)^H mu^-1 (grad + i kappa) */
int bloch_debug_apply_aux(bloch_handle h, int mode, const double *x, double *y, int nvec);
/* The matrix Pi itself in CSR form (rows = ND dofs, column d * N_h1 + node, real weights): host code only, also on
 * BLOCH_DEVICE_NONE handles, so that it can be compared with an assembled interpolation without a GPU.
 * rowptr == NULL: only *nnz is returned; otherwise rowptr[N + 1], col[nnz], val[nnz] are filled. */
int bloch_debug_pi_matrix(bloch_handle h, int64_t *nnz, int64_t *rowptr, int32_t *col, double *val);
/* Test hook for the nested-mesh transfers of the projector's H1 multigrid (no counterpart in the reference, whose
 * MINRES has no hierarchy): level 0 (the handle's mesh) <-> level 1 (n_sub / 2) in each implementation -
 * variant 0 explicit CSR, 1 sum-factorised parent kernels, 2 element-wise kernels; dir 0: y(2 N_h1) = P x(2 n_coarse),
 * dir 1: y(2 n_coarse) = P^T x(2 N_h1).  x == y == NULL only returns n_coarse. */
int bloch_debug_mg_transfer(bloch_handle h, int variant, int dir, const double *x, double *y, int nvec, int64_t *n_coarse);
/* Measurement hook: fp64 FMA throughput of the handle's device in TFLOP/s (roofline denominator
 * of the flop-bound side of the element kernels; not part of the reference interface). */
int bloch_debug_fp64_peak(bloch_handle h, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* BLOCH_B200_H */
