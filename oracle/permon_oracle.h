/*
 * permon_oracle.h -- CPU restatement of PERMON's QPSMPGP / QPSSMALXE hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or as
 * the timed CPU baseline.  The product path (permon_b200/) never links,
 * imports or calls it.
 *
 * Parity status: PINNED.  The restatement reproduces the reference's own
 * golden outputs exactly (iteration / Hessian-mult / step-kind counts of
 * src/tutorials/output/ex1_{1,opt,optapprox,bb,projcg}.out,
 * ex2_1_infinite-{false,true}.out, and the 11-digit per-iteration traces and
 * alpha of jbearing2_{4,5,6}.out); see tests/test_oracle_golden.py.  Also
 * ex3_1.out (MPGP on the dualised problem of ex3, restated with a dense K^-1
 * instead of MUMPS: counts and printed KKT digits agree) and, for SMALXE,
 * ex3_nullspace.out: the dual QP with a zero-row equality constraint runs
 * QPSSolve_SMALXE + QPSConverged_Inner_SMALXE to "1 outer iteration, inner
 * CONVERGED_HAPPY_BREAKDOWN after 46 iterations, 74/18/27/1" exactly.  The
 * SMALXE update rules for a non-empty B (M1 / rho updates) have no reference
 * output reachable without QPTDualize of FETI problems; they are covered by
 * hand-checkable cases.
 * Parity UNPINNED for the "next" rows added later:
 * orc_cg_solve (QPSKSP = PETSc's KSPCG recurrence), orc_pcpg_solve (QPSPCPG),
 * orc_orth_rows / orc_homogenize / the projected operator (QPTOrthonormalizeEq,
 * QPTEnforceEqByProjector).  The reference has no test or golden output that
 * reaches them without QPTDualize + MUMPS; they are checked against their
 * defining identities and scipy (tests/test_oracle_linear.py,
 * tests/test_oracle_transforms.py).
 *
 * The reference (permon/permon) is plain C on PETSc.  PETSc is an un-vendored
 * third-party dependency (requires >= 3.17, uses 3.23/3.24-era API; no lock
 * file) that carries the arithmetic (MatMult, VecDot, VecAXPY ...).  Its
 * published semantics are restated here: CSR SpMV in row order with a running
 * sum, VecDot = plain sum in index order (per rank, then over ranks in rank
 * order), VecNorm_2 = sqrt(sum x^2), VecAXPY(y,a,x): y += a*x,
 * VecAYPX(y,a,x): y = x + a*y, VecWAXPY(w,a,x,y): w = a*x + y.
 *
 * All file:line citations are into /root/reference.
 */
#ifndef PERMON_ORACLE_H
#define PERMON_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_DECIDE (-1.0) /* PETSC_DECIDE */

/* PETSc: PETSC_INFINITY = PETSC_MAX_REAL/4, PETSC_NINFINITY = -PETSC_INFINITY */
#define ORC_INFINITY (1.7976931348623157e+308 / 4.0)
#define ORC_NINFINITY (-ORC_INFINITY)
#define ORC_EPS 2.2204460492503131e-16 /* PETSC_MACHINE_EPSILON (double) */

/* KSPConvergedReason values used by the path (petscksp.h) */
enum {
  ORC_CONVERGED_ITERATING       = 0,
  ORC_CONVERGED_RTOL            = 2,
  ORC_CONVERGED_ATOL            = 3,
  ORC_CONVERGED_ITS             = 4,
  ORC_CONVERGED_HAPPY_BREAKDOWN = 7,
  ORC_DIVERGED_ITS              = -3,
  ORC_DIVERGED_DTOL             = -4,
  ORC_DIVERGED_BREAKDOWN        = -5,
  ORC_DIVERGED_NANORINF         = -9,
  ORC_DIVERGED_INDEFINITE_MAT   = -10
};

/* include/permonqps.h:100-114 */
enum { ORC_EXP_STD = 0, ORC_EXP_PROJCG, ORC_EXP_GF, ORC_EXP_G, ORC_EXP_GFGR, ORC_EXP_GGR };
enum { ORC_LEN_FIXED = 0, ORC_LEN_OPT, ORC_LEN_OPTAPPROX, ORC_LEN_BB };

/* A linear operator y = A x, row-partition free (one address space).
 * kind 0: CSR matrix (n x n)
 * kind 1: product  A = M1 * M2  with M1 (n x d) and M2 (d x n) both CSR
 *         (PERMON MatCreateProd semantics, src/mat/impls/composite/matprod.c:42)
 * Optional penalty term (MatMult_Penalized, src/qp/utils/matpenalized.c:12-22):
 *   y = rho * B^T (B x) + A x, B dense m x n row-major (m small).          */
typedef struct {
  int           kind;
  int           n;
  const int    *ia, *ja;
  const double *a;
  int           d; /* inner dimension for kind 1 */
  const int    *ia2, *ja2;
  const double *a2;
  double       *twork; /* length d, kind 1 */
  int           m;     /* # penalty rows (0 = none) */
  const double *B;     /* m x n */
  double        rho;
  double       *bwork; /* length m */
  /* projected operator (QPTEnforceEqByProjector, qptransform.c:283-296): y = P A P x (pmode 2) or y = P A x (pmode 1, only
   * equality constraints present) with P = I - G^T (G G^T)^{-1} G; 0 = no projector.  The plain / product part above is "A". */
  int           pmode;
  int           pm;     /* rows of G */
  const double *PG;     /* pm x n */
  int           porth;  /* G has orthonormal rows (QPPFApplyQ then skips the coarse solve, qppf.c:454-502) */
  double       *pw1, *pw2; /* two work vectors of length n */
  /* the penalty rows are orthonormal IMPLICITLY (QPTOrthonormalizeEq with MAT_ORTH_IMPLICIT, qptransform.c:566-636): QPPFApplyGtG
   * routes through QPPFApplyQ (qppf.c:586-589), the penalty term is rho * B^T (B B^T)^{-1} (B x)                                  */
  int           bimplicit;
} orc_op;

/* Box constraint (QPC_Box, src/qpc/impls/box/qpcboximpl.h:5-10) optionally
 * restricted to an index set (QPCSetIS; QPCGetSubvector qpc.c:416-437).    */
typedef struct {
  int           n;    /* full vector length */
  int           nis;  /* -1: no IS (all components); else IS length */
  const int    *is;   /* indices into the full vector */
  const double *lb;   /* NULL or length (nis<0 ? n : nis) */
  const double *ub;
  double        astol; /* qpc->astol = 10*eps, qpc.c:28 */
} orc_box;

typedef struct {
  /* QPS base (src/qps/interface/qps.c:73-76) */
  double rtol, atol, divtol;
  int    max_it;
  /* MPGP (src/qps/impls/mpgp/mpgp.c:827-843) */
  double alpha_user;   /* ORC_DECIDE -> 2.0 */
  int    alpha_direct; /* QPS_ARG_DIRECT */
  double gamma;
  double maxeig;       /* ORC_DECIDE -> power method */
  double maxeig_tol;   /* ORC_DECIDE -> 1e-4 */
  int    maxeig_iter;  /* -1 -> 50 */
  int    exptype, explengthtype;
  int    resetalpha;
  int    fallback, fallback2;
  int    nthreads;     /* OpenMP threads standing in for MPI ranks; <=0: omp default */
} orc_mpgp_opts;

typedef struct {
  int    its, reason;
  int    nmv, ncg, nexp, nprop, nfinc, nfall;
  double rnorm, alpha, maxeig;
  int    maxeig_its;
  double norm_rhs, ttol;
  double seconds; /* wall time of the iteration loop only */
} orc_mpgp_result;

/* optional per-iteration trace = what QPSMonitorDefault_MPGP prints, mpgp.c:21-34 */
typedef struct {
  int     cap, len;
  char   *step;
  double *rnorm, *gfnorm, *gcnorm, *alpha;
} orc_trace;

typedef struct {
  double rtol, atol, divtol;
  int    max_it; /* outer, default 100 (smalxe.c:1203) */
  double M1_user;  int M1_direct;  double M1_update;
  double rho_user; int rho_direct; double rho_update, rho_update_late;
  double eta_user; int eta_direct;
  double rtol_E, update_threshold;
  double maxeig, maxeig_tol; int maxeig_iter;
  int    inject_maxeig, inject_maxeig_set;
  int    inner_iter_min, inner_no_gtol_stop;
  int    knoll, get_lambda;
  orc_mpgp_opts inner; /* options of the inner MPGP ("smalxe_" prefix) */
  /* QPTOrthonormalizeEq(MAT_ORTH_IMPLICIT) in front of SMALXE: B_E becomes a dummy without MatMult, so ||B u|| is updated by
   * QPSSMALXEUpdateNormBu_SMALXEON (smalxe.c:265-285) or, with lag_enabled, by ..._Lag_SMALXEON (:289-370; defaults :1190-1200) */
  int    implicit_orth;
  int    lag_enabled, lag_offset, Jstart, Jstep, Jend;
  double lag_lower, lag_upper;
} orc_smalxe_opts;

typedef struct {
  int    outer_its, reason, inner_its_accu, inner_reason_last;
  int    M1_updates, M1_hits, eta_hits, rho_updates, state;
  int    nmv, ncg, nexp, nprop; /* inner MPGP counters since reset */
  double rnorm, normBu, M1, rho, maxeig, maxeig_inner, alpha_inner, eta;
  double seconds;
} orc_smalxe_result;

/* linear solvers of the "next" rows: QPSKSP (= PETSc KSPCG, PCNONE, unpreconditioned norm, non-zero
 * initial guess: src/qps/impls/ksp/qpsksp.c:232-253) and QPSPCPG (src/qps/impls/pcpg/pcpg.c:49-131) */
typedef struct {
  double rtol, atol, divtol;
  int    max_it;
  int    nthreads;
} orc_lin_opts;
typedef struct {
  int    its, reason;
  double rnorm, norm_rhs;
  double seconds;
} orc_lin_result;
/* "next" row rank 2: pieces of QPTOrthonormalizeEq / QPTHomogenizeEq / QPTEnforceEqByProjector.
 * orth types as MatOrthType (permonmat.h:160-167): 1 = MAT_ORTH_GS, 3 = MAT_ORTH_CHOLESKY (explicit forms).            */
int  orc_orth_rows(int n, int m, const double *B, const double *c, int type, double *TB, double *Tc, double *T);
int  orc_rows_orthonormal(int n, int m, const double *B);
void orc_apply_P(int n, int m, const double *G, const double *v, double *Pv);
/* QPTHomogenizeEq (qptransform.c:437-527): xtilde = G^T (G G^T)^{-1} c, b_h = b - A xtilde, bounds shifted by xtilde */
void orc_homogenize(const orc_op *op, const double *b, const orc_box *bx, int m, const double *G, const double *c, double *xtilde, double *b_h,
                    double *lb_h, double *ub_h);
void orc_default_lin_opts(orc_lin_opts *o);
int  orc_cg_solve(const orc_op *op, const double *b, double *x, const orc_lin_opts *opts, orc_lin_result *res);
/* min 1/2 x'Ax - b'x  s.t.  G x = c  (G dense row-major m x n, c may be NULL = 0) */
int  orc_pcpg_solve(const orc_op *op, const double *b, int m, const double *G, const double *c, double *x, const orc_lin_opts *opts,
                    orc_lin_result *res);

void orc_default_mpgp_opts(orc_mpgp_opts *o);
void orc_default_smalxe_opts(orc_smalxe_opts *o);
void orc_set_threads(int nthreads);
int  orc_get_max_threads(void);

/* building blocks (each also used by the GPU unit-parity tests) */
void   orc_spmv(int n, const int *ia, const int *ja, const double *a, const double *x, double *y);
void   orc_op_apply(const orc_op *op, const double *x, double *y);
double orc_dot(int n, const double *x, const double *y);
double orc_norm2(int n, const double *x);
void   orc_qpc_project(const orc_box *bx, const double *x, double *Px);
void   orc_qpc_grads(const orc_box *bx, const double *x, const double *g, double *gf, double *gc);
void   orc_qpc_gradreduced(const orc_box *bx, const double *x, const double *gf, double alpha, double *gr);
double orc_qpc_feas(const orc_box *bx, const double *x, const double *d);
int    orc_max_eigenvalue(const orc_op *op, double tol, int maxits, double *lambda_out);
double orc_objective(const orc_op *op, const double *b, const double *x);
/* QPComputeMissingBoxMultipliers (qp.c:829-890); Bt_lambda may be NULL. llb/lub full-IS length. */
void   orc_box_multipliers(const orc_op *op, const double *b, const double *Bt_lambda, const orc_box *bx,
                           const double *x, double *llb, double *lub);
/* the r = ... numbers of QPViewKKT (qp.c:245-370) + QPCViewKKT_Box (qpcbox.c:333-427):
 * out[0]=||b||, out[1]=||A x - b [+Bt_lambda] - llb + lub||, out[2]=||min(x-lb,0)||,
 * out[3]=||min(llb,0)||, out[4]=|llb'(lb-x)|, out[5]=||max(x-ub,0)||, out[6]=||min(lub,0)||,
 * out[7]=|lub'(x-ub)|  (entries that do not apply are left NaN)               */
void   orc_kkt(const orc_op *op, const double *b, const double *Bt_lambda, const orc_box *bx, const double *x,
               const double *llb, const double *lub, double *out);

/* QPSSetup_MPGP + QPSSolve_MPGP with QPSConvergedDefault.  x: in = initial guess, out = solution. */
int orc_mpgp_solve(const orc_op *op, const double *b, const orc_box *bx, double *x, const orc_mpgp_opts *opts,
                   orc_mpgp_result *res, orc_trace *trace);

/* QPSSetUp_SMALXE + QPSSolve_SMALXE with inner MPGP.  Equality B x = c, B dense m x n (op->m/B/rho are
 * overwritten internally); c may be NULL.  Bt_lambda (n) and lambda (m) optional outputs. */
int orc_smalxe_solve(orc_op *op, const double *b, const orc_box *bx, int m, const double *B, const double *c, double *x,
                     const orc_smalxe_opts *opts, orc_smalxe_result *res, double *Bt_lambda_out, double *lambda_out);

#ifdef __cplusplus
}
#endif
#endif
