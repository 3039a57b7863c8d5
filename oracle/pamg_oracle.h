/*
 * pamg_oracle.h -- CPU oracle for the P-A_multigrids hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain fp64 restatement of the reference's serial Fortran algorithm.  It is
 * NOT part of the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product path
 * (libpamg_cuda.so) never links, loads or calls anything in oracle/.
 *
 * PARITY: the reference is Fortran 90 with no tests and cannot be compiled in this image (no Fortran
 * compiler), so the oracle cannot be run against the reference binary.  It is pinned against the ONE
 * output of the reference binary that the repository ships: the dump DG-rectangular_structured of
 * trans_rec (main.F90:19), which orc_trans_rec reproduces to the reference's single precision
 * (tests/test_oracle_known_answers.py::test_trans_rec_reproduces_the_reference_output_file; that path
 * shares the shape-function / face-geometry / upwind-flux / local-solve / time-loop structure with the
 * triangle path).  For the triangle multigrid path itself no reference output exists: PARITY UNPINNED
 * there; it is pinned instead against the known answers the reference formulas imply (SURVEY.md
 * appendix C), the reconstructible dump DG-rectangular_structured_analytical, the manufactured
 * solution sin(x+y), the erfc boundary-layer gate of Check_thermal_analytical_validation.py
 * (tests/test_erfc_boundary_layer.py), geometric / algebraic invariants (tests/test_oracle_*.py) and
 * ANALYTIC identities that any correct implementation of the discretisation must satisfy: the operator,
 * the implicit matrix and the explicit step applied to a continuous linear field have closed forms
 * (tests/test_oracle_operator_consistency.py, tests/test_oracle_solver_properties.py), the direct
 * solution of the assembled system is the fixed point of the smoothers and the limit of the V-cycles; every entry
 * of the level-1 matrix and right-hand side equals an independent exact-integration assembly in numpy that sees only
 * the child coordinates (tests/test_oracle_independent_assembly.py); the theta != 1 branches converge to the exact
 * semi-discrete solution with the order of theta (tests/test_oracle_theta.py).
 *
 * Every function cites the reference file:line it follows.  Two behaviours exist
 * where the reference is work-in-progress:
 *   LITERAL  - what HEAD computes (face block commented out, in-place M*src, 3-child
 *              mean restriction, mixed prolongation)
 *   INTENDED - the mathematically consistent composition (face block on, b-Ax,
 *              P1 interpolation and its transpose)
 * selected by fields of orc_params, never silently.
 */
#ifndef PAMG_ORACLE_H
#define PAMG_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int32_t n_split;        /* transport_tri_semi.F90:118 */
  int32_t multi_levels;   /* main.F90:46 (last-but-one argument) */
  int32_t face_terms;     /* 0: face loop body commented out (HEAD, :619-688); 1: face block on */
  int32_t literal_source; /* 1: get_RHS overwrites source in place while summing (:456); 0: plain M*src */
  int32_t transfer;       /* 0: literal restrictor/prolongator (splitting.F90:10-91); 1: P1 interp / transpose */
  int32_t residual_sign;  /* +1: r = A x - b (HEAD :869);  -1: r = b - A x */
  int32_t halo_rule;      /* 0: literal Dir/Nside reversal table (splitting.F90:1256-1391); 1: geometric */
  int32_t coarse_bc_zero; /* 1: Dirichlet data is 0 on levels > 1 (error equation); 0: sin(x+y) on all levels (HEAD) */
  double theta;           /* :117 */
  double dt;              /* :133 */
  double k;               /* :136 */
  double omega;           /* :140 */
  double u_x, u_y;        /* :208-212 */
  double source_coef;     /* source = source_coef * sin(x+y); HEAD: -2k (:593) */
} orc_params;

/* ---- tables and numbering (pure functions) ---------------------------------- */
void orc_tables(double* n9, double* nlx18, double* weight3, double* sn4, double* snlx4, double* sweight2);
void orc_get_str_info(int n_split, int ele, int* irow, int* ipos, int* orientation);
void orc_get_splitting(const double* X6, int n_split, int str_ele, double* x6);
void orc_str_neig(int n, int32_t* str_neig /* 3*4^n, (f,ele) -> [ (ele-1)*3 + f-1 ] */);
void orc_surf_ele(int n, int32_t* surf_ele /* 3*2^n, (i,f) -> [ (f-1)*2^n + i-1 ] */);
void orc_element_conversion(int coarse_ele, int i_split_coarse, int32_t* fin4);
void orc_tri_det_nlx(const double* x6, double* nx18 /* (g,d,i) -> [(g*2+d)*3+i] */, double* detwei3);
void orc_face_geometry(const double* x6, int iface_gmsh, double* sdetwei2, double* snorm4 /* (s,d) */);
int  orc_findinv(const double* A, double* Ainv, int n); /* row-major n x n; returns errorflag 0 / -1 */

/* ---- mesh reader (ReadMSH + CheckNeig + getNeigDataMesh) ---------------------- */
/* returns number of triangles (<0 on error); arrays sized max_tri */
int orc_read_msh(const char* path, int max_tri, double* X /* [u][node][dim] */, int32_t* neig,
                 int32_t* dir, int32_t* region);
void orc_neig_data(int U, const int32_t* neig, const int32_t* dir, int32_t* fneig, int32_t* snodes /* [u][f][2] */);

/* ---- semi-structured multigrid problem -------------------------------------- */
typedef struct orc_semi orc_semi;
enum { ORC_TNEW = 0, ORC_TOLD = 1, ORC_RHS = 2, ORC_RES = 3, ORC_SRC = 4, ORC_TNONLIN = 5 };
orc_semi* orc_semi_create(const orc_params* p, int U, const double* X, const int32_t* neig,
                          const int32_t* fneig, const int32_t* dir);
void orc_semi_destroy(orc_semi*);
/* Dirichlet data per parent face [U][3] (gmsh face order), used where Neig == 0: kind 0 = sin(x+y) (HEAD,
 * splitting.F90:1246-1252), 1 = the constant value[] (update_overlaps' t_bc argument, :1210), 2 = open face (no data,
 * no penalty term; exterior trace = interior trace).  Clears the halo strips. */
void orc_semi_set_boundary(orc_semi*, const int32_t* kind, const double* value);
double* orc_semi_field(orc_semi*, int field, int level);      /* (3, 4^s, U) Fortran order */
double* orc_semi_overlap(orc_semi*, int level, int old);     /* [u][face][3*2^s] */
int64_t orc_semi_ndof(orc_semi*, int level);
void orc_semi_update_overlaps(orc_semi*, int level);
/* solver: 1 Jacobi, 2 Richardson, 3 GS lexicographic (reference order), 4 GS two-colour (down then up) */
void orc_semi_smooth(orc_semi*, int level, int solver, int nsweeps);
void orc_semi_build_rhs(orc_semi*);                           /* level-1 RHS from told + source */
void orc_semi_residual(orc_semi*, int level, double* l2, double* linf);
double orc_semi_convergence(orc_semi*, int level);            /* get_convergence :876-889 */
void orc_semi_restrict(orc_semi*, int fine_level);
void orc_semi_prolong(orc_semi*, int fine_level);
/* INTENDED V-cycle (pre-smooth, r=b-Ax, restrict, recurse, prolong, post-smooth); returns cycles used */
int orc_semi_vcycle_solve(orc_semi*, int solver, int nu1, int nu2, int ncoarse, int max_cycles,
                          double tol, double* relres_hist);
/* LITERAL HEAD time loop body (:316-379) for one time step */
void orc_semi_literal_timestep(orc_semi*, int solver, int n_multigrid, int n_smooth);
void orc_semi_set_threads(int nthreads);                      /* OpenMP threads for Jacobi/residual */

/* ---- unstructured explicit DG step (transport_tri_unstr.F90:588-795) --------- */
void orc_unstr_explicit(int E, const double* X, const int32_t* neig, const int32_t* fneig,
                        const int32_t* dir, double u_x, double u_y, double dt, int ntime, int nits,
                        int njac_its, int exact_minv, int use_dir, double t_bc, double* tnew /* (3,E) in/out */);

/* ---- unstructured implicit assembly + dense solve (transport_tri_unstr.F90:214-387); dense (3E)^2, small E only */
void orc_unstr_implicit_assemble(int E, const double* X, const int32_t* neig, const int32_t* fneig, double u_x,
                                 double u_y, double dt, int use_dir, double* A, double* Mdt);
/* the same + the diffusion operator of the iterative path (volume term and face penalty, get_A_x transport_tri_semi.F90:412-448) */
void orc_unstr_implicit_assemble_diff(int E, const double* X, const int32_t* neig, const int32_t* fneig, double u_x,
                                      double u_y, double k, double dt, int use_dir, double* A, double* Mdt);
int orc_unstr_implicit(int E, const double* X, const int32_t* neig, const int32_t* fneig, double u_x, double u_y,
                       double dt, int ntime, int nits, int use_dir, double* tnew);

/* ---- Petrov-Galerkin stabilisation (transport_tri_unstr.F90:239-267,278): diff_coe (3,E), stab [E][3][3] row-major */
void orc_unstr_stab(int E, const double* X, const double* tnew, const double* told, double u_x, double u_y, double dt,
                    double* diff_coe, double* stab);
/* INTENDED: lhs + flux + blockdiag(stab(tnew_nonlin, told)) re-solved nits times per step (dense, small E only) */
int orc_unstr_implicit_stab(int E, const double* X, const int32_t* neig, const int32_t* fneig, double u_x, double u_y,
                            double dt, int ntime, int nits, int use_dir, double* tnew);

/* ---- trans_rec (transport_rect.F90:7-380): explicit Q1 DG on a structured rectangular grid; returns ntime.
 * volume_term 0 = HEAD (tnew_gi never set; the version that wrote DG-rectangular_structured), 1 = intended.
 * x_all [totele][4][2], tnew_out [totele][4] */
int orc_trans_rec(double CFL, int no_ele_row, int no_ele_col, double x_length, double y_length, double u_x, double u_y,
                  double time, int nits, int njac_its, int direct_solver, int volume_term, double* x_all, double* tnew_out);

/* ---- analytic cases ----------------------------------------------------------- */
/* transport_rect.F90:83,101-105,337-344 : fills x[800], t[800] */
void orc_rect_analytical(double CFL, int no_ele_row, double x_length, double u_x, double time,
                         double* x_out, double* t_out);
/* Check_thermal_analytical_validation.py:34-43 */
double orc_thermal_analytical(double x, double t, double u, double gamma);

#ifdef __cplusplus
}
#endif
#endif
