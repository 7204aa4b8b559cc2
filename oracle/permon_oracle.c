/*
 * permon_oracle.c -- CPU restatement of PERMON's QPSMPGP / QPSSMALXE hot path.
 * TEST INFRASTRUCTURE ONLY (see permon_oracle.h).  Parity status: PINNED against the
 * reference's golden outputs (tests/test_oracle_golden.py).
 *
 * The operation ORDER follows the reference statement by statement and is deliberately
 * un-fused (one loop per PETSc Vec call), so that (i) branch decisions reproduce the golden
 * iteration counts and (ii) timed with OpenMP threads standing in for MPI ranks it is a fair
 * "reference CPU path" for bench.py.  Compile with -ffp-contract=off.
 *
 * All file:line citations are into /root/reference.
 */
#include "permon_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;

void orc_set_threads(int nthreads)
{
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  g_threads = nthreads;
  omp_set_num_threads(nthreads);
#else
  (void)nthreads;
  g_threads = 1;
#endif
}

int orc_get_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static double now_seconds(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

#define ORC_MIN(a, b) (((a) < (b)) ? (a) : (b)) /* PetscMin */
#define ORC_MAX(a, b) (((a) < (b)) ? (b) : (a)) /* PetscMax */

/* ------------------------------------------------------------------------------------------ */
/* PETSc Vec semantics (third-party; restated from the published API documentation)            */
/* ------------------------------------------------------------------------------------------ */

#define PFOR _Pragma("omp parallel for schedule(static) if (g_threads > 1)")

static void v_copy(int n, const double *x, double *y) /* VecCopy(x,y) */
{
  if (x == y) return;
  PFOR for (int i = 0; i < n; i++) y[i] = x[i];
}
static void v_set(int n, double *x, double a) /* VecSet */
{
  PFOR for (int i = 0; i < n; i++) x[i] = a;
}
static void v_scale(int n, double *x, double a) /* VecScale */
{
  PFOR for (int i = 0; i < n; i++) x[i] *= a;
}
static void v_axpy(int n, double *y, double a, const double *x) /* VecAXPY: y += a x */
{
  PFOR for (int i = 0; i < n; i++) y[i] += a * x[i];
}
static void v_aypx(int n, double *y, double a, const double *x) /* VecAYPX: y = x + a y */
{
  PFOR for (int i = 0; i < n; i++) y[i] = x[i] + a * y[i];
}
static void v_waxpy(int n, double *w, double a, const double *x, const double *y) /* VecWAXPY: w = a x + y */
{
  if (a == 1.0) {
    PFOR for (int i = 0; i < n; i++) w[i] = y[i] + x[i];
  } else if (a == -1.0) {
    PFOR for (int i = 0; i < n; i++) w[i] = y[i] - x[i];
  } else {
    PFOR for (int i = 0; i < n; i++) w[i] = a * x[i] + y[i];
  }
}

/* VecDot: plain running sum in index order on each rank, partial sums added in rank order
 * (MPI_Allreduce(SUM)); threads stand in for ranks with a static contiguous row partition. */
double orc_dot(int n, const double *x, const double *y)
{
  int nt = g_threads;
  if (nt <= 1 || n < 4 * nt) {
    double s = 0.0;
    for (int i = 0; i < n; i++) s += x[i] * y[i];
    return s;
  }
  double part[256];
  if (nt > 256) nt = 256;
#pragma omp parallel num_threads(nt)
  {
#ifdef _OPENMP
    int t = omp_get_thread_num();
#else
    int t = 0;
#endif
    long   lo = (long)n * t / nt, hi = (long)n * (t + 1) / nt;
    double s = 0.0;
    for (long i = lo; i < hi; i++) s += x[i] * y[i];
    part[t] = s;
  }
  double s = 0.0;
  for (int t = 0; t < nt; t++) s += part[t];
  return s;
}

double orc_norm2(int n, const double *x) /* VecNorm(NORM_2) */
{
  return sqrt(orc_dot(n, x, x));
}

/* ------------------------------------------------------------------------------------------ */
/* Operators                                                                                  */
/* ------------------------------------------------------------------------------------------ */

/* MatMult_SeqAIJ semantics: y_i = sum_j a_ij x_j, running sum from 0 in storage order. */
void orc_spmv(int n, const int *ia, const int *ja, const double *a, const double *x, double *y)
{
  PFOR for (int i = 0; i < n; i++)
  {
    double s = 0.0;
    for (int k = ia[i]; k < ia[i + 1]; k++) s += a[k] * x[ja[k]];
    y[i] = s;
  }
}

/* MatMultAdd_SeqAIJ semantics with y == z: running sum starts from y_i. */
static void spmv_add(int n, const int *ia, const int *ja, const double *a, const double *x, double *y)
{
  PFOR for (int i = 0; i < n; i++)
  {
    double s = y[i];
    for (int k = ia[i]; k < ia[i + 1]; k++) s += a[k] * x[ja[k]];
    y[i] = s;
  }
}

static void op_apply_plain(const orc_op *op, const double *x, double *y)
{
  if (op->kind == 0) {
    orc_spmv(op->n, op->ia, op->ja, op->a, x, y);
  } else {
    /* MAT_COMPOSITE_MULTIPLICATIVE: mats[1]*(mats[0]*x); here A = M1*(M2 x) (matprod.c:42-48) */
    orc_spmv(op->d, op->ia2, op->ja2, op->a2, x, op->twork);
    orc_spmv(op->n, op->ia, op->ja, op->a, op->twork, y);
  }
}

/* y = B^T (B x) for dense row-major B (m x n): QPPFApplyGtG, src/qppf/interface/qppf.c:580-605
 * (G_left = G v; GtGv = Gt G_left).  When G has orthonormal rows the reference routes through
 * QPPFApplyQ (qppf.c:586-589, 454-502) which computes the same G^T (G v) product.              */
static void gtg_apply(int n, int m, const double *B, const double *x, double *y, double *bwork)
{
  for (int j = 0; j < m; j++) bwork[j] = orc_dot(n, B + (size_t)j * n, x);
  if (m == 1) {
    /* MatMultTranspose_OneRow: VecCopy(a,z); VecScale(z,xval)  (onerow.c:41-57) */
    const double t = bwork[0];
    PFOR for (int i = 0; i < n; i++) y[i] = B[i] * t;
  } else {
    PFOR for (int i = 0; i < n; i++)
    {
      double s = 0.0;
      for (int j = 0; j < m; j++) s += B[(size_t)j * n + i] * bwork[j];
      y[i] = s;
    }
  }
}

static int  ggt_solve(int n, int m, const double *B, const double *r, double *y);
/* y = B^T (B B^T)^{-1} (B x): QPPFApplyQ (qppf.c:454-502) for rows that are NOT explicitly orthonormal -- what QPPFApplyGtG computes when
 * the rows are orthonormal implicitly (qppf.c:586-589)                                                                                  */
static void q_apply(int n, int m, const double *B, const double *x, double *y, double *bwork)
{
  double s[64];
  for (int j = 0; j < m; j++) bwork[j] = orc_dot(n, B + (size_t)j * n, x); /* G_left = G v */
  ggt_solve(n, m, B, bwork, s);                                             /* Gt_right = (G G^T)^{-1} G_left */
  PFOR for (int i = 0; i < n; i++)
  {
    double t = 0.0;
    for (int j = 0; j < m; j++) t += B[(size_t)j * n + i] * s[j];
    y[i] = t;
  }
}
static void pf_apply_P(int n, int m, const double *G, int orth, const double *v, double *Pv, double *gl, double *y);

/* the Hessian without the penalty term: A, or the MatProd P*A*P / P*A of QPTEnforceEqByProjector (applied right to left,
 * matprod.c:42-48, each P through QPPFApplyP)                                                                            */
static void op_apply_hessian(const orc_op *op, const double *x, double *y)
{
  if (!op->pmode) {
    op_apply_plain(op, x, y);
    return;
  }
  double gl[8], yy[8];
  const double *in = x;
  if (op->pmode == 2) {
    pf_apply_P(op->n, op->pm, op->PG, op->porth, x, op->pw1, gl, yy);
    in = op->pw1;
  }
  op_apply_plain(op, in, op->pw2);
  pf_apply_P(op->n, op->pm, op->PG, op->porth, op->pw2, y, gl, yy);
}

/* MatMult for the (possibly penalised) Hessian.
 * MatMult_Penalized (src/qp/utils/matpenalized.c:12-22): y = BtB x; y *= rho; y += A x.         */
void orc_op_apply(const orc_op *op, const double *x, double *y)
{
  if (op->m <= 0) {
    op_apply_hessian(op, x, y);
    return;
  }
  if (op->bimplicit) q_apply(op->n, op->m, op->B, x, y, op->bwork);
  else gtg_apply(op->n, op->m, op->B, x, y, op->bwork);
  v_scale(op->n, y, op->rho);
  if (op->pmode) {
    double *w = (double *)malloc(sizeof(double) * (size_t)op->n);
    op_apply_hessian(op, x, w);
    v_axpy(op->n, y, 1.0, w);
    free(w);
  } else if (op->kind == 0) {
    spmv_add(op->n, op->ia, op->ja, op->a, x, y);
  } else {
    /* composite operator: MatMultAdd = MatMult into a work vector + add */
    double *w = (double *)malloc(sizeof(double) * (size_t)op->n);
    op_apply_plain(op, x, w);
    v_axpy(op->n, y, 1.0, w);
    free(w);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* QPC box: src/qpc/interface/qpc.c + src/qpc/impls/box/qpcbox.c                              */
/* ------------------------------------------------------------------------------------------ */

#define NSUB(bx) ((bx)->nis < 0 ? (bx)->n : (bx)->nis)
#define IDX(bx, k) ((bx)->nis < 0 ? (k) : (bx)->is[(k)])

/* QPCProject (qpc.c:466-491): VecCopy(x,Px) then QPCProject_Box on the sub-vectors
 * (qpcbox.c:290-305): lb ? (Px=max(x,lb); ub ? Px=min(Px,ub)) : Px=min(x,ub).               */
void orc_qpc_project(const orc_box *bx, const double *x, double *Px)
{
  const int ns = NSUB(bx);
  v_copy(bx->n, x, Px);
  if (bx->lb) {
    PFOR for (int k = 0; k < ns; k++)
    {
      int i = IDX(bx, k);
      Px[i] = ORC_MAX(x[i], bx->lb[k]);
    }
    if (bx->ub) {
      PFOR for (int k = 0; k < ns; k++)
      {
        int i = IDX(bx, k);
        Px[i] = ORC_MIN(Px[i], bx->ub[k]);
      }
    }
  } else if (bx->ub) {
    PFOR for (int k = 0; k < ns; k++)
    {
      int i = IDX(bx, k);
      Px[i] = ORC_MIN(x[i], bx->ub[k]);
    }
  }
}

/* QPCGrads (qpc.c:540-569): gf = g; gc = 0; then QPCGrads_Box (qpcbox.c:21-64). */
void orc_qpc_grads(const orc_box *bx, const double *x, const double *g, double *gf, double *gc)
{
  const int     ns = NSUB(bx);
  const double *lb = bx->lb, *ub = bx->ub;
  const double  astol = bx->astol;
  v_copy(bx->n, g, gf);
  v_set(bx->n, gc, 0.0);
  PFOR for (int k = 0; k < ns; k++)
  {
    int i = IDX(bx, k);
    if (lb && fabs(x[i] - lb[k]) <= astol) {
      gf[i] = 0.0;
      gc[i] = ORC_MIN(g[i], 0.0);
    } else if (ub && fabs(x[i] - ub[k]) <= astol) {
      gf[i] = 0.0;
      gc[i] = ORC_MAX(g[i], 0.0);
    } else {
      gf[i] = g[i];
    }
  }
}

/* QPCGradReduced (qpc.c:589-615): gr = gf; then QPCGradReduced_Box (qpcbox.c:68-100). */
void orc_qpc_gradreduced(const orc_box *bx, const double *x, const double *gf, double alpha, double *gr)
{
  const int     ns = NSUB(bx);
  const double *lb = bx->lb, *ub = bx->ub;
  v_copy(bx->n, gf, gr);
  PFOR for (int k = 0; k < ns; k++)
  {
    int i = IDX(bx, k);
    if (lb && gf[i] > 0.0) {
      gr[i] = ORC_MIN(gf[i], (x[i] - lb[k]) / alpha);
    } else if (ub && gf[i] < 0.0) {
      gr[i] = ORC_MAX(gf[i], (x[i] - ub[k]) / alpha);
    }
  }
}

/* QPCFeas (qpc.c:503-527) + QPCFeas_Box (qpcbox.c:104-146); the MPI_Allreduce(MIN) of qpc.c:521
 * is the min over thread-local results.                                                        */
double orc_qpc_feas(const orc_box *bx, const double *x, const double *d)
{
  const int     ns = NSUB(bx);
  const double *lb = bx->lb, *ub = bx->ub;
  double        alpha = ORC_INFINITY;
#pragma omp parallel for schedule(static) reduction(min : alpha) if (g_threads > 1)
  for (int k = 0; k < ns; k++) {
    int    i = IDX(bx, k);
    double ai;
    if (d[i] > 0. && lb && lb[k] > ORC_NINFINITY) {
      ai = x[i] - lb[k];
      ai = ai / d[i];
      if (ai < alpha) alpha = ai;
    }
    if (d[i] < 0. && ub && ub[k] < ORC_INFINITY) {
      ai = x[i] - ub[k];
      ai = ai / d[i];
      if (ai < alpha) alpha = ai;
    }
  }
  return alpha;
}

/* ------------------------------------------------------------------------------------------ */
/* MatGetMaxEigenvalue: src/mat/interface/permonmatutils.c:442-522                            */
/* ------------------------------------------------------------------------------------------ */
int orc_max_eigenvalue(const orc_op *op, double tol, int maxits, double *lambda_out)
{
  const int n = op->n;
  double   *v = (double *)malloc(sizeof(double) * (size_t)n);
  double   *Av = (double *)malloc(sizeof(double) * (size_t)n);
  double    lambda = 0.0, lambda0, err, relerr, vAv, vv;
  int       i;

  if (tol == ORC_DECIDE || tol == -2.0) tol = 1e-4; /* :473 */
  if (maxits == -1 || maxits == -2) maxits = 50;    /* :474 */
  v_set(n, v, 1.0);                                 /* :477 */
  for (i = 1; i <= maxits; i++) {                   /* :484 */
    lambda0 = lambda;
    orc_op_apply(op, v, Av);       /* :487 */
    vAv    = orc_dot(n, Av, v);    /* VecMDot(v,2,{Av,v}) :491 */
    vv     = orc_dot(n, v, v);
    lambda = vAv / vv;             /* :492 */
    if (lambda < 2.2204460492503131e-16) { /* :493-502 A v fell into the null space: Av <- VecSetRandom(PETSCRAND48), which becomes the next v */
      /* PetscRandomCreate seeds with 0x12345678 (+ 76543 * rank, rank 0 here); rand48 draws drand48() per entry, in order:
         X <- 0x5DEECE66D X + 0xB (mod 2^48) from (seed << 16) | 0x330E, value X / 2^48 */
      unsigned long long X = (((unsigned long long)0x12345678u) << 16) | 0x330Eull;
      for (int k = 0; k < n; k++) {
        X     = (0x5DEECE66Dull * X + 0xBull) & 0xFFFFFFFFFFFFull;
        Av[k] = (double)X * (1.0 / 281474976710656.0);
      }
    }
    err    = fabs(lambda - lambda0); /* :504 */
    relerr = err / fabs(lambda);
    if (relerr < tol) break;       /* :506 */
    v_copy(n, Av, v);              /* :509 */
    v_scale(n, v, 1.0 / sqrt(vv)); /* :510  (divides by ||v||, not ||Av||) */
  }
  *lambda_out = lambda;
  free(v);
  free(Av);
  return i;
}

/* QPComputeObjective: src/qp/interface/qp.c:913-927   f = -x'(b - 1/2 A x) */
double orc_objective(const orc_op *op, const double *b, const double *x)
{
  const int n = op->n;
  double   *w = (double *)malloc(sizeof(double) * (size_t)n);
  orc_op_apply(op, x, w);  /* xwork = A x */
  v_aypx(n, w, -0.5, b);   /* xwork = b - 0.5 xwork */
  double f = -orc_dot(n, x, w);
  free(w);
  return f;
}

/* QPComputeObjectiveFromGradient: qp.c:981-996   f = x'(g - b)/2 */
static double objective_from_gradient(int n, const double *b, const double *x, const double *g, double *xwork)
{
  v_waxpy(n, xwork, -1.0, b, g);
  return .5 * orc_dot(n, x, xwork);
}

/* QPComputeMissingBoxMultipliers (qp.c:829-890) via QPComputeLagrangianGradient (qp.c:668-775) on a QP
 * whose QPC was removed: r = A x - b [+ Bt_lambda]; llb = r; lub = -r; both clamped at 0 iff both exist. */
void orc_box_multipliers(const orc_op *op, const double *b, const double *Bt_lambda, const orc_box *bx, const double *x,
                         double *llb, double *lub)
{
  const int n = op->n, ns = NSUB(bx);
  double   *r = (double *)malloc(sizeof(double) * (size_t)n);
  orc_op_apply(op, x, r);
  v_axpy(n, r, -1.0, b);
  if (Bt_lambda) v_axpy(n, r, 1.0, Bt_lambda);
  if (bx->lb) {
    for (int k = 0; k < ns; k++) llb[k] = r[IDX(bx, k)];
  }
  if (bx->ub) {
    for (int k = 0; k < ns; k++) lub[k] = r[IDX(bx, k)];
    for (int k = 0; k < ns; k++) lub[k] *= -1.0;
  }
  if (bx->lb && bx->ub) {
    for (int k = 0; k < ns; k++) llb[k] = ORC_MAX(llb[k], 0.0);
    for (int k = 0; k < ns; k++) lub[k] = ORC_MAX(lub[k], 0.0);
  }
  free(r);
}

/* the numbers printed by QPViewKKT (qp.c:245-370) and QPCViewKKT_Box (qpcbox.c:333-427) */
void orc_kkt(const orc_op *op, const double *b, const double *Bt_lambda, const orc_box *bx, const double *x,
             const double *llb, const double *lub, double *out)
{
  const int n = op->n, ns = NSUB(bx);
  double   *r = (double *)malloc(sizeof(double) * (size_t)n);
  double   *t = (double *)malloc(sizeof(double) * (size_t)n);
  for (int k = 0; k < 8; k++) out[k] = NAN;
  out[0] = orc_norm2(n, b);
  /* QPComputeLagrangianGradient: r = A x - b - llb + lub + Bt_lambda */
  orc_op_apply(op, x, r);
  v_axpy(n, r, -1.0, b);
  if (bx->lb) for (int k = 0; k < ns; k++) r[IDX(bx, k)] += -1.0 * llb[k];
  if (bx->ub) for (int k = 0; k < ns; k++) r[IDX(bx, k)] += 1.0 * lub[k];
  if (Bt_lambda) v_axpy(n, r, 1.0, Bt_lambda);
  out[1] = orc_norm2(n, r);
  /* NB: QPCViewKKT is handed the full x; with an IS the reference's sub-vector logic is bypassed
   * there (qp.c:368) -- we evaluate on the constrained components.                              */
  if (bx->lb) {
    for (int k = 0; k < ns; k++) t[k] = ORC_MIN(x[IDX(bx, k)] - bx->lb[k], 0.0);
    out[2] = orc_norm2(ns, t);
    for (int k = 0; k < ns; k++) t[k] = ORC_MIN(llb[k], 0.0);
    out[3] = orc_norm2(ns, t);
    for (int k = 0; k < ns; k++) {
      t[k] = bx->lb[k] - x[IDX(bx, k)];
      if (bx->lb[k] <= ORC_NINFINITY) t[k] = -1.0;
    }
    out[4] = fabs(orc_dot(ns, llb, t));
  }
  if (bx->ub) {
    for (int k = 0; k < ns; k++) t[k] = ORC_MAX(x[IDX(bx, k)] - bx->ub[k], 0.0);
    out[5] = orc_norm2(ns, t);
    for (int k = 0; k < ns; k++) t[k] = ORC_MIN(lub[k], 0.0);
    out[6] = orc_norm2(ns, t);
    for (int k = 0; k < ns; k++) {
      t[k] = x[IDX(bx, k)] - bx->ub[k];
      if (bx->ub[k] >= ORC_INFINITY) t[k] = 1.0;
    }
    out[7] = fabs(orc_dot(ns, lub, t));
  }
  free(r);
  free(t);
}

/* ------------------------------------------------------------------------------------------ */
/* defaults                                                                                   */
/* ------------------------------------------------------------------------------------------ */
void orc_default_mpgp_opts(orc_mpgp_opts *o)
{
  o->rtol = 1e-5; o->atol = 1e-50; o->divtol = 1e4; o->max_it = 10000; /* qps.c:73-76 */
  o->alpha_user = ORC_DECIDE; o->alpha_direct = 0; o->gamma = 1.0;      /* mpgp.c:827-829 */
  o->maxeig = ORC_DECIDE; o->maxeig_tol = ORC_DECIDE; o->maxeig_iter = -1; /* :830-832 */
  o->exptype = ORC_EXP_STD; o->explengthtype = ORC_LEN_FIXED;           /* :836-837 */
  o->resetalpha = 0; o->fallback = 0; o->fallback2 = 0;                 /* :840-843 */
  o->nthreads = 1;
}

void orc_default_smalxe_opts(orc_smalxe_opts *o)
{
  o->rtol = 1e-5; o->atol = 1e-50; o->divtol = 1e4; o->max_it = 100; /* qps.c:73-76, smalxe.c:1203 */
  o->M1_user = 1e2; o->M1_direct = 0; o->M1_update = 2.0;              /* smalxe.c:1159-1163 */
  o->rtol_E = 1e-0;                                                    /* :1165 */
  o->rho_user = 1.1; o->rho_direct = 0; o->rho_update = 1.0; o->rho_update_late = 2.0; /* :1167-1170 */
  o->eta_user = 1e-1; o->eta_direct = 0;                               /* :1172-1173 */
  o->update_threshold = 0.0;                                           /* :1176 */
  o->maxeig = ORC_DECIDE; o->maxeig_tol = ORC_DECIDE; o->maxeig_iter = -1; /* :1178-1180 */
  o->inject_maxeig = 0; o->inject_maxeig_set = 0;                      /* :1181-1182 */
  o->inner_iter_min = 1; o->inner_no_gtol_stop = 0;                    /* :1205-1206 */
  o->knoll = 0; o->get_lambda = 0;
  o->implicit_orth = 0;
  o->lag_enabled = 0; o->lag_offset = 2; o->Jstart = 10; o->Jstep = 5; o->Jend = 20; o->lag_lower = 0.1; o->lag_upper = 1.1; /* :1190-1200 */
  orc_default_mpgp_opts(&o->inner);
}

/* ------------------------------------------------------------------------------------------ */
/* MPGP: src/qps/impls/mpgp/mpgp.c                                                            */
/* ------------------------------------------------------------------------------------------ */

struct mpgp;
typedef int (*conv_fn)(struct mpgp *s, void *ctx); /* sets s->reason; (*qps->convergencetest)() */

typedef struct mpgp {
  const orc_op  *op;
  const double  *b;
  const orc_box *bx;
  double        *x;
  int            n;
  /* QPS base */
  double rtol, atol, divtol, rnorm;
  int    max_it, iteration, reason;
  /* QPSConvergedDefaultCtx (include/permon/private/qpsimpl.h:73-76) */
  double norm_rhs, ttol, norm_rhs_div;
  int    cctx_setup_called;
  /* QPS_MPGP (src/qps/impls/mpgp/mpgpimpl.h:5-38) */
  double alpha, alpha_user, gamma, maxeig, maxeig_tol;
  int    alpha_direct, maxeig_iter, maxeig_its;
  int    exptype, explengthtype, expproject, resetalpha, fallback, fallback2;
  int    nmv, ncg, nexp, nprop, nfinc, nfall;
  char   currentStepType;
  double gfnorm, gcnorm;
  /* work vectors (mpgp.c:6-17) */
  double *gP, *gf, *gc, *g, *p, *Ap, *gr, *w7, *w8, *w9, *xwork;
  double *expdirection, *explengthvec, *explengthvecold, *xold;
  conv_fn conv;
  void   *conv_ctx;
  orc_trace *trace;
} mpgp_t;

/* QPSConvergedDefaultSetUp: src/qps/interface/qps.c:718-731 */
static void converged_default_setup(mpgp_t *s)
{
  if (s->cctx_setup_called) return;
  s->norm_rhs          = orc_norm2(s->n, s->b);
  s->ttol              = ORC_MAX(s->rtol * s->norm_rhs, s->atol);
  s->norm_rhs_div      = s->norm_rhs;
  s->cctx_setup_called = 1;
}

/* QPSConvergedDefault: src/qps/interface/qps.c:675-714 */
static int converged_default(mpgp_t *s, void *ctx)
{
  (void)ctx;
  const int    i = s->iteration;
  const double rnorm = s->rnorm;
  s->reason = ORC_CONVERGED_ITERATING;
  if (!s->cctx_setup_called) converged_default_setup(s);
  if (i > s->max_it) { /* :688 */
    s->reason = ORC_DIVERGED_ITS;
    return 0;
  }
  if (isnan(rnorm) || isinf(rnorm)) { /* :696 */
    s->reason = ORC_DIVERGED_NANORINF;
  } else if (rnorm <= s->ttol) { /* :699 */
    if (rnorm < s->atol) s->reason = ORC_CONVERGED_ATOL;
    else s->reason = ORC_CONVERGED_RTOL;
  } else if (rnorm >= s->divtol * s->norm_rhs_div) { /* :708 */
    s->reason = ORC_DIVERGED_DTOL;
  }
  return 0;
}

/* MPGPGrads: mpgp.c:198-223 */
static void mpgp_grads(mpgp_t *s, const double *x, const double *g)
{
  orc_qpc_grads(s->bx, x, g, s->gf, s->gc);                 /* :219 */
  orc_qpc_gradreduced(s->bx, x, s->gf, s->alpha, s->gr);     /* :220 */
  v_waxpy(s->n, s->gP, 1.0, s->gf, s->gc);                   /* :221 */
}

/* MPGPExpansionLength: mpgp.c:233-287 */
static void mpgp_expansion_length(mpgp_t *s)
{
  double dots[2];
  switch (s->explengthtype) {
  case ORC_LEN_FIXED:
    break;
  case ORC_LEN_OPT:
    orc_op_apply(s->op, s->explengthvec, s->Ap); /* :250 */
    s->nmv++;                                    /* :251 */
    dots[0] = orc_dot(s->n, s->g, s->explengthvec);  /* VecMDot(v,2,{g,Ap}) :252 */
    dots[1] = orc_dot(s->n, s->Ap, s->explengthvec);
    if (dots[1] == .0 && s->resetalpha) s->alpha = s->alpha / s->maxeig; /* :253-254 */
    else s->alpha = s->alpha_user * (dots[0] / dots[1]);                 /* :256 */
    break;
  case ORC_LEN_OPTAPPROX:
    if (s->g != s->explengthvec) { /* :262 */
      dots[0] = orc_dot(s->n, s->g, s->explengthvec);
      dots[1] = orc_dot(s->n, s->explengthvec, s->explengthvec);
      s->alpha = s->alpha_user * (dots[0] / dots[1]); /* :264 */
    } else {
      s->alpha = s->alpha_user; /* :266 */
    }
    s->alpha = s->alpha / s->maxeig; /* :268 */
    break;
  case ORC_LEN_BB:
    v_aypx(s->n, s->explengthvecold, -1.0, s->explengthvec); /* :274 */
    v_aypx(s->n, s->xold, -1.0, s->x);                       /* :275 */
    dots[0] = orc_dot(s->n, s->explengthvecold, s->explengthvecold); /* VecMDot(vecs[0],2,vecs) :276 */
    dots[1] = orc_dot(s->n, s->xold, s->explengthvecold);
    if (dots[1] == .0 && s->resetalpha) s->alpha = s->alpha / s->maxeig; /* :277-278 */
    else s->alpha = s->alpha_user * (dots[0] / dots[1]);                 /* :280 */
    break;
  }
}

/* MPGPExpansion_Std: mpgp.c:299-323 */
static void mpgp_expansion_std(mpgp_t *s, double afeas, double acg)
{
  (void)acg;
  v_axpy(s->n, s->x, -afeas, s->p);             /* :316 */
  v_axpy(s->n, s->g, -afeas, s->Ap);            /* :317 */
  mpgp_grads(s, s->x, s->g);                    /* :318 */
  mpgp_expansion_length(s);                     /* :320 */
  v_axpy(s->n, s->x, -s->alpha, s->expdirection); /* :321 */
}

/* MPGPExpansion_ProjCG: mpgp.c:335-349 */
static void mpgp_expansion_projcg(mpgp_t *s, double afeas, double acg)
{
  (void)afeas;
  v_axpy(s->n, s->x, -acg, s->p); /* :347 */
}

/* QPSSetup_MPGP: mpgp.c:359-428 */
static void mpgp_setup(mpgp_t *s)
{
  s->expproject = 1; /* QPSCreate_MPGP :839 */
  switch (s->exptype) {
  case ORC_EXP_STD:
    s->expdirection = s->gr; s->explengthvec = s->gr;
    if (s->explengthtype == ORC_LEN_FIXED) s->expproject = 0; /* :388 */
    break;
  case ORC_EXP_GF:   s->expdirection = s->gf; s->explengthvec = s->gf; break;
  case ORC_EXP_G:    s->expdirection = s->g;  s->explengthvec = s->g;  break;
  case ORC_EXP_GFGR: s->expdirection = s->gf; s->explengthvec = s->gr; break;
  case ORC_EXP_GGR:  s->expdirection = s->g;  s->explengthvec = s->gr; break;
  case ORC_EXP_PROJCG: s->expdirection = s->gf; s->explengthvec = s->gf; break; /* :406-411 */
  }
  if (!s->alpha_direct) { /* :417-425 */
    if (s->maxeig == ORC_DECIDE) s->maxeig_its = orc_max_eigenvalue(s->op, s->maxeig_tol, s->maxeig_iter, &s->maxeig);
    if (s->alpha_user == ORC_DECIDE) s->alpha_user = 2.0;
    s->alpha = s->alpha_user / s->maxeig;
  } else {
    s->alpha = s->alpha_user;
  }
}

static void trace_push(mpgp_t *s)
{
  orc_trace *t = s->trace;
  if (!t || t->len >= t->cap) return;
  int k = t->len++;
  t->step[k]   = s->currentStepType;
  t->rnorm[k]  = s->rnorm;
  t->gfnorm[k] = s->gfnorm;
  t->gcnorm[k] = s->gcnorm;
  t->alpha[k]  = s->alpha;
}

/* QPSSolve_MPGP: mpgp.c:438-650 */
static void mpgp_solve(mpgp_t *s)
{
  const int n = s->n;
  double   *x = s->x, *g = s->g, *p = s->p, *Ap = s->Ap, *gf = s->gf, *gc = s->gc, *gP = s->gP;
  const double *b = s->b;
  double   *gold = NULL;
  double    gamma2, acg, bcg, afeas, pAp, gcTgc, gfTgf, f, fold;
  int       nmv = 0, ncg = 0, nprop = 0, nexp = 0, nfinc = 0, nfall = 0;

  if (s->explengthtype == ORC_LEN_BB) { /* :479-486 */
    s->explengthvecold = s->w7;
    s->xold            = s->w8;
    if (s->fallback || s->fallback2) gold = s->w9;
  } else if (s->fallback || s->fallback2) {
    s->xold = s->w7;
    gold    = s->w8;
  }
  gamma2 = s->gamma * s->gamma; /* :489 */

  orc_qpc_project(s->bx, x, x); /* :497 */
  orc_op_apply(s->op, x, g);    /* :500 */
  nmv++;
  v_axpy(n, g, -1.0, b);        /* :502 */
  mpgp_grads(s, x, g);          /* :504 */
  v_copy(n, gf, p);             /* :507 */

  s->currentStepType = ' ';
  s->iteration       = 0;
  while (1) {
    s->rnorm = orc_norm2(n, gP);  /* :514 */
    gcTgc    = orc_dot(n, gc, gc); /* :517 */
    gfTgf    = orc_dot(n, gf, gf); /* :521 */
    s->gfnorm = sqrt(gfTgf);
    s->gcnorm = sqrt(gcTgc);
    trace_push(s);                /* QPSMonitor :524-528 */

    s->conv(s, s->conv_ctx);      /* :531 */
    if (s->reason != ORC_CONVERGED_ITERATING) break;

    if (gcTgc <= gamma2 * gfTgf) { /* :535 */
      orc_op_apply(s->op, p, Ap);  /* :537 */
      nmv++;
      pAp   = orc_dot(n, p, Ap);   /* :541 */
      acg   = orc_dot(n, g, p);    /* :542 */
      acg   = acg / pAp;           /* :543 */
      afeas = orc_qpc_feas(s->bx, x, p); /* :544 */

      if (acg <= afeas) { /* :547 */
        ncg++;
        s->currentStepType = 'c';
        v_axpy(n, x, -acg, p);   /* :553 */
        v_axpy(n, g, -acg, Ap);  /* :554 */
        mpgp_grads(s, x, g);     /* :555 */
        bcg = orc_dot(n, Ap, gf); /* :558 */
        bcg = bcg / pAp;         /* :559 */
        v_aypx(n, p, -bcg, gf);  /* :560 */
      } else {
        nexp++;
        s->currentStepType = 'e';
        if (s->explengthtype == ORC_LEN_BB || s->fallback || s->fallback2) { /* :568-571 */
          v_copy(n, x, s->xold);
          if (s->explengthtype == ORC_LEN_BB) v_copy(n, s->explengthvec, s->explengthvecold);
        }
        if (s->exptype == ORC_EXP_PROJCG) mpgp_expansion_projcg(s, afeas, acg); /* :573 */
        else mpgp_expansion_std(s, afeas, acg);
        if (s->expproject) orc_qpc_project(s->bx, x, x); /* :574 */

        if (s->fallback || s->fallback2) v_copy(n, g, gold); /* :577 */
        orc_op_apply(s->op, x, g); /* :578 */
        nmv++;
        v_axpy(n, g, -1.0, b);     /* :580 */

        if (s->fallback || s->fallback2) { /* :582-611 */
          fold = objective_from_gradient(n, b, s->xold, gold, s->xwork);
          f    = objective_from_gradient(n, b, x, g, s->xwork);
          if (f > fold) {
            nfinc++;
            if (s->fallback2) {
              mpgp_grads(s, x, g);
              gcTgc = orc_dot(n, gc, gc);
              gfTgf = orc_dot(n, gf, gf);
              if (gcTgc <= gamma2 * gfTgf) s->fallback = 0;
              else s->fallback = 1;
            }
            if (s->fallback) {
              nfall++;
              s->currentStepType = 'f';
              v_copy(n, s->xold, x);
              v_copy(n, gold, g);
              if (s->fallback2) mpgp_grads(s, s->xold, gold);
              mpgp_expansion_std(s, afeas, acg);
              orc_qpc_project(s->bx, x, x);
              orc_op_apply(s->op, x, g);
              nmv++;
              v_axpy(n, g, -1.0, b);
            }
          }
        }
        mpgp_grads(s, x, g); /* :613 */
        v_copy(n, gf, p);    /* :615 */
      }
    } else {
      nprop++;
      s->currentStepType = 'p';
      v_copy(n, gc, p);           /* :623 */
      orc_op_apply(s->op, p, Ap); /* :624 */
      nmv++;
      pAp = orc_dot(n, p, Ap);    /* :628 */
      acg = orc_dot(n, g, p);     /* :629 */
      acg = acg / pAp;            /* :630 */
      v_axpy(n, x, -acg, p);      /* :633 */
      v_axpy(n, g, -acg, Ap);     /* :634 */
      mpgp_grads(s, x, g);        /* :635 */
      v_copy(n, gf, p);           /* :638 */
    }
    s->iteration++; /* :640 */
  }
  s->ncg += ncg; s->nexp += nexp; s->nmv += nmv; s->nprop += nprop; s->nfinc += nfinc; s->nfall += nfall; /* :643-648 */
}

static void mpgp_alloc(mpgp_t *s, const orc_op *op, const double *b, const orc_box *bx, double *x, const orc_mpgp_opts *o)
{
  const size_t nb = sizeof(double) * (size_t)op->n;
  memset(s, 0, sizeof(*s));
  s->op = op; s->b = b; s->bx = bx; s->x = x; s->n = op->n;
  s->rtol = o->rtol; s->atol = o->atol; s->divtol = o->divtol; s->max_it = o->max_it;
  s->alpha_user = o->alpha_user; s->alpha_direct = o->alpha_direct; s->gamma = o->gamma;
  s->maxeig = o->maxeig; s->maxeig_tol = o->maxeig_tol; s->maxeig_iter = o->maxeig_iter;
  s->exptype = o->exptype; s->explengthtype = o->explengthtype; s->resetalpha = o->resetalpha;
  s->fallback = o->fallback; s->fallback2 = o->fallback2;
  if (s->fallback2) s->fallback = 0; /* mpgp.c:744 */
  s->gP = malloc(nb); s->gf = malloc(nb); s->gc = malloc(nb); s->g = malloc(nb); s->p = malloc(nb);
  s->Ap = malloc(nb); s->gr = malloc(nb); s->w7 = malloc(nb); s->w8 = malloc(nb); s->w9 = malloc(nb);
  s->xwork = malloc(nb);
  /* first touch by the owning thread */
  v_set(s->n, s->gP, 0); v_set(s->n, s->gf, 0); v_set(s->n, s->gc, 0); v_set(s->n, s->g, 0); v_set(s->n, s->p, 0);
  v_set(s->n, s->Ap, 0); v_set(s->n, s->gr, 0); v_set(s->n, s->w7, 0); v_set(s->n, s->w8, 0); v_set(s->n, s->w9, 0);
  v_set(s->n, s->xwork, 0);
  s->conv = converged_default;
  s->conv_ctx = NULL;
}

static void mpgp_free(mpgp_t *s)
{
  free(s->gP); free(s->gf); free(s->gc); free(s->g); free(s->p); free(s->Ap); free(s->gr);
  free(s->w7); free(s->w8); free(s->w9); free(s->xwork);
}

int orc_mpgp_solve(const orc_op *op, const double *b, const orc_box *bx, double *x, const orc_mpgp_opts *opts,
                   orc_mpgp_result *res, orc_trace *trace)
{
  mpgp_t s;
  orc_set_threads(opts->nthreads <= 0 ? 0 : opts->nthreads);
  mpgp_alloc(&s, op, b, bx, x, opts);
  s.trace = trace;
  if (trace) trace->len = 0;
  mpgp_setup(&s);
  double t0 = now_seconds();
  mpgp_solve(&s);
  double t1 = now_seconds();
  res->its = s.iteration; res->reason = s.reason; res->nmv = s.nmv; res->ncg = s.ncg; res->nexp = s.nexp;
  res->nprop = s.nprop; res->nfinc = s.nfinc; res->nfall = s.nfall; res->rnorm = s.rnorm; res->alpha = s.alpha;
  res->maxeig = s.maxeig; res->maxeig_its = s.maxeig_its; res->norm_rhs = s.norm_rhs; res->ttol = s.ttol;
  res->seconds = t1 - t0;
  mpgp_free(&s);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* SMALXE: src/qps/impls/smalxe/smalxe.c                                                      */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  /* outer QPS */
  double rtol, atol, divtol, rnorm;
  int    max_it, iteration, reason;
  /* outer QPSConvergedDefaultCtx */
  double norm_rhs, ttol, norm_rhs_div;
  int    cctx_setup_called;
  const double *b_outer;
  /* QPS_SMALXE (smalxeimpl.h:13-67) */
  double M1, M1_initial, M1_update, rtol_E, rho_update, rho_update_late, eta, update_threshold;
  int    M1_updates, M1_hits, eta_hits, rho_updates, state, inner_iter_accu, inner_iter_min, inner_no_gtol_stop;
  double normBu, normBu_old, normBu_prev, enorm;
  /* QPSConvergedCtx_Inner_SMALXE (smalxeimpl.h:5-11) */
  double gtol, norm_rhs_outer, ttol_outer, MNormBu;
  /* data */
  int           n, m;
  const double *B, *c; /* homogenised: c == NULL */
  double       *Bu;
  orc_op       *op_inner; /* penalised operator (shares A with the outer op) */
  mpgp_t       *inner;
  /* implicit orthonormalisation: SMALXEON norm updates (smalxe.c:265-370) */
  int     implicit, lag_enabled, lag_offset, Jstart, Jstep, Jend;
  double  lag_lower, lag_upper;
  double *BtBu;            /* work[0] */
  double  lag_normBu0;     /* the function statics of :291-293 */
  int     lag_II, lag_J, lag_neval, lag_niter;
} smalxe_t;

/* QPSConvergedDefault for the OUTER solver (same code as converged_default, other struct) */
static void outer_converged_default(smalxe_t *o)
{
  const int    i = o->iteration;
  const double rnorm = o->rnorm;
  o->reason = ORC_CONVERGED_ITERATING;
  if (!o->cctx_setup_called) { /* qps.c:718-731 */
    o->norm_rhs          = orc_norm2(o->n, o->b_outer);
    o->ttol              = ORC_MAX(o->rtol * o->norm_rhs, o->atol);
    o->norm_rhs_div      = o->norm_rhs;
    o->cctx_setup_called = 1;
  }
  if (i > o->max_it) { o->reason = ORC_DIVERGED_ITS; return; }
  if (isnan(rnorm) || isinf(rnorm)) o->reason = ORC_DIVERGED_NANORINF;
  else if (rnorm <= o->ttol) o->reason = (rnorm < o->atol) ? ORC_CONVERGED_ATOL : ORC_CONVERGED_RTOL;
  else if (rnorm >= o->divtol * o->norm_rhs_div) o->reason = ORC_DIVERGED_DTOL;
}

/* QPSSMALXEUpdateNormBu_SMALXEON: smalxe.c:265-285 */
static void update_normBu_on(smalxe_t *o, const double *u, double *normBu, double *enorm)
{
  q_apply(o->n, o->m, o->B, u, o->BtBu, o->op_inner->bwork); /* BtBu = B'*B*u with BtB = the penalised term (Q) */
  const double dot = orc_dot(o->n, u, o->BtBu);
  *normBu = sqrt(dot);
  *enorm  = *normBu / o->rtol_E;
}

/* QPSSMALXEUpdateNormBu_Lag_SMALXEON: smalxe.c:289-370 */
static void update_normBu_lag_on(smalxe_t *o, const double *u, double *normBu, double *enorm)
{
  double normBu_approx, normBu_exact, enorm_exact, rdiff;
  if (o->inner->iteration <= o->lag_offset) { /* :312-319 */
    update_normBu_on(o, u, &normBu_exact, &enorm_exact);
    o->lag_neval++;
    o->lag_normBu0 = normBu_exact;
    normBu_approx  = o->lag_normBu0;
    o->lag_J       = o->Jstart;
    o->lag_II      = 0;
  } else {
    if (o->lag_II == 0) { /* :321-339 */
      update_normBu_on(o, u, &normBu_exact, &enorm_exact);
      o->lag_neval++;
      rdiff = fabs(normBu_exact / o->lag_normBu0);
      if (rdiff >= o->lag_upper) { o->lag_II = 0; o->lag_J = o->Jstart; }
      else if (rdiff < o->lag_lower) { o->lag_II = 0; o->lag_J = o->Jstart; }
      else o->lag_II++;
      o->lag_normBu0 = normBu_exact;
    } else {
      o->lag_II++;
    }
    normBu_approx = o->lag_normBu0;
  }
  o->lag_niter++;
  if (o->lag_II == o->lag_J) { /* :345-348 */
    o->lag_II = 0;
    if (o->lag_J < o->Jend) o->lag_J += o->Jstep;
  }
  *normBu = normBu_approx;
  *enorm  = *normBu / o->rtol_E;
}

/* QPSSMALXEUpdateNormBu_SMALXE: smalxe.c:247-261; the variant is chosen at set-up (:878-886) */
static void update_normBu(smalxe_t *o, const double *u, double *normBu, double *enorm)
{
  if (o->implicit) {
    if (o->lag_enabled) update_normBu_lag_on(o, u, normBu, enorm);
    else update_normBu_on(o, u, normBu, enorm);
    return;
  }
  for (int j = 0; j < o->m; j++) o->Bu[j] = orc_dot(o->n, o->B + (size_t)j * o->n, u); /* Bu = B u */
  if (o->c) for (int j = 0; j < o->m; j++) o->Bu[j] += -1.0 * o->c[j];
  *normBu = orc_norm2(o->m, o->Bu);
  *enorm  = *normBu / o->rtol_E;
}

/* QPSConverged_Inner_SMALXE: smalxe.c:610-692 */
static int converged_inner_smalxe(mpgp_t *in, void *ctx)
{
  smalxe_t    *o = (smalxe_t *)ctx;
  const int    i = in->iteration;
  const double gnorm = in->rnorm;

  in->reason = ORC_CONVERGED_ITERATING;
  update_normBu(o, in->x, &o->normBu, &o->enorm); /* :625 */
  o->rnorm   = ORC_MAX(o->enorm, gnorm);          /* :626 */
  o->MNormBu = o->M1 * o->normBu;                 /* :627 */
  in->atol   = ORC_MIN(o->MNormBu, o->eta);       /* :628 */

  if (i > in->max_it - o->inner_iter_accu) { /* :633 */
    in->reason = ORC_DIVERGED_ITS;
    o->reason  = ORC_DIVERGED_BREAKDOWN;
    return 0;
  }
  if (isnan(gnorm) || isinf(gnorm)) { /* :641 */
    in->reason = ORC_DIVERGED_NANORINF;
    o->reason  = ORC_DIVERGED_BREAKDOWN;
    return 0;
  }
  outer_converged_default(o); /* :648 */
  if (o->reason) {            /* :650-659 */
    in->reason = (o->reason > 0) ? ORC_CONVERGED_HAPPY_BREAKDOWN : ORC_DIVERGED_BREAKDOWN;
    return 0;
  }
  if (gnorm < in->atol) { /* :661-671 */
    in->reason = ORC_CONVERGED_ATOL;
    if (o->MNormBu < o->eta) o->M1_hits++;
    else o->eta_hits++;
    return 0;
  }
  if (o->state == 3 && (i < o->inner_iter_min || o->inner_no_gtol_stop)) return 0; /* :673 */
  if (gnorm <= o->gtol) { /* :675-690 */
    if (in->rnorm > o->enorm) {
      /* skipping gtol criterion because G > E */
    } else {
      if (o->inner_no_gtol_stop < 2) in->reason = ORC_CONVERGED_RTOL;
      if (o->state != 3) o->state = 3;
    }
  }
  return 0;
}

static int rows_orthonormal(int n, int m, const double *B)
{
  /* stands in for MatHasOrthonormalRows(G, PETSC_SMALL, 3 random trials) (permonmatorth.c:551-565):
   * here G G^T is formed explicitly (m is small) and compared with I to PETSC_SMALL = 1e-10.     */
  for (int i = 0; i < m; i++)
    for (int j = 0; j <= i; j++) {
      double d = orc_dot(n, B + (size_t)i * n, B + (size_t)j * n);
      if (fabs(d - (i == j ? 1.0 : 0.0)) > 1e-10) return 0;
    }
  return 1;
}

/* solve (B B^T) y = r by Cholesky: stands in for MatMult_Inv (KSPPREONLY + PCCHOLESKY,
 * src/mat/impls/inv/matinv.c:487-488,734-743).  m is small.                                */
static int ggt_solve(int n, int m, const double *B, const double *r, double *y)
{
  double *L = (double *)calloc((size_t)m * m, sizeof(double));
  for (int i = 0; i < m; i++)
    for (int j = 0; j <= i; j++) L[i * m + j] = orc_dot(n, B + (size_t)i * n, B + (size_t)j * n);
  for (int j = 0; j < m; j++) {
    double d = L[j * m + j];
    for (int k = 0; k < j; k++) d -= L[j * m + k] * L[j * m + k];
    if (d <= 0.0) { free(L); return 1; }
    d = sqrt(d);
    L[j * m + j] = d;
    for (int i = j + 1; i < m; i++) {
      double v = L[i * m + j];
      for (int k = 0; k < j; k++) v -= L[i * m + k] * L[j * m + k];
      L[i * m + j] = v / d;
    }
  }
  for (int i = 0; i < m; i++) {
    double v = r[i];
    for (int k = 0; k < i; k++) v -= L[i * m + k] * y[k];
    y[i] = v / L[i * m + i];
  }
  for (int i = m - 1; i >= 0; i--) {
    double v = y[i];
    for (int k = i + 1; k < m; k++) v -= L[k * m + i] * y[k];
    y[i] = v / L[i * m + i];
  }
  free(L);
  return 0;
}

int orc_smalxe_solve(orc_op *op, const double *b_user, const orc_box *bx_user, int m, const double *B, const double *c,
                     double *x_user, const orc_smalxe_opts *opts, orc_smalxe_result *res, double *Bt_lambda_out,
                     double *lambda_out)
{
  const int    n = op->n;
  const size_t nb = sizeof(double) * (size_t)n;
  smalxe_t     o;
  mpgp_t       in;
  orc_op       op_inner;
  orc_box      bx = *bx_user;
  double      *xtilde = NULL, *b_h = NULL, *lb_h = NULL, *ub_h = NULL, *x = x_user;
  const double *b = b_user;
  double      *Btmu = (double *)calloc((size_t)n, sizeof(double));
  double      *b_inner = (double *)malloc(nb);
  double      *BtBu = (double *)malloc(nb);
  double      *bw = (double *)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
  double       maxeig = opts->maxeig, rho, maxeig_inner, Lag, Lag_old;
  int          i, orth;

  orc_set_threads(opts->inner.nthreads <= 0 ? 0 : opts->inner.nthreads);
  memset(&o, 0, sizeof(o));
  op->m = 0; /* the outer operator is the plain Hessian */

  /* --- QPSSetUp_SMALXE: smalxe.c:772-888 --- */
  if (c) { /* QPTHomogenizeEq: src/qp/interface/qptransform.c:437-527 */
    const int ns = NSUB(&bx);
    double   *y = (double *)malloc(sizeof(double) * (size_t)m);
    xtilde = (double *)malloc(nb);
    orth   = rows_orthonormal(n, m, B);
    /* QPPFApplyHalfQTranspose (qppf.c:535-568): xtilde = G^T (G G^T)^{-1} c */
    if (!orth) ggt_solve(n, m, B, c, y);
    else memcpy(y, c, sizeof(double) * (size_t)m);
    for (int k = 0; k < n; k++) {
      double s = 0.0;
      for (int j = 0; j < m; j++) s += B[(size_t)j * n + k] * y[j];
      xtilde[k] = s;
    }
    b_h = (double *)malloc(nb);
    orc_op_apply(op, xtilde, b_h);     /* :468 */
    v_aypx(n, b_h, -1.0, b_user);      /* :469  b_bar = b - A xtilde */
    b = b_h;
    if (bx.lb) { /* :498-501 */
      lb_h = (double *)malloc(sizeof(double) * (size_t)ns);
      for (int k = 0; k < ns; k++) lb_h[k] = bx_user->lb[k] - xtilde[IDX(&bx, k)];
      bx.lb = lb_h;
    }
    if (bx.ub) { /* :503-506 */
      ub_h = (double *)malloc(sizeof(double) * (size_t)ns);
      for (int k = 0; k < ns; k++) ub_h[k] = bx_user->ub[k] - xtilde[IDX(&bx, k)];
      bx.ub = ub_h;
    }
    /* child->x is destroyed (:515); QPInitializeInitialVector_Private (qp.c:23-43) then gives the child a
     * COPY of the parent's x (un-shifted) as its initial guess */
    x = (double *)malloc(nb);
    memcpy(x, x_user, nb);
    free(y);
  }
  o.n = n; o.m = m; o.B = B; o.c = NULL; o.Bu = bw; o.b_outer = b;
  o.rtol = opts->rtol; o.atol = opts->atol; o.divtol = opts->divtol; o.max_it = opts->max_it;
  o.M1_update = opts->M1_update; o.rtol_E = opts->rtol_E; o.rho_update = opts->rho_update;
  o.rho_update_late = opts->rho_update_late; o.update_threshold = opts->update_threshold;
  o.inner_iter_min = opts->inner_iter_min; o.inner_no_gtol_stop = opts->inner_no_gtol_stop;
  o.state = 1; o.normBu = NAN; o.enorm = NAN;

  o.eta = opts->eta_user; /* :806-811 */
  if (!opts->eta_direct) o.eta *= orc_norm2(n, b);
  o.M1_initial = opts->M1_user; /* :814-818 */
  if (!opts->M1_direct) {
    if (maxeig == ORC_DECIDE) orc_max_eigenvalue(op, opts->maxeig_tol, opts->maxeig_iter, &maxeig);
    o.M1_initial *= maxeig;
  }
  if (!opts->rho_direct) { /* :821-826 */
    if (maxeig == ORC_DECIDE) orc_max_eigenvalue(op, opts->maxeig_tol, opts->maxeig_iter, &maxeig);
    rho = opts->rho_user * maxeig;
  } else {
    rho = opts->rho_user;
  }
  orth = rows_orthonormal(n, m, B); /* QPPFSetUp :834 -> qppf.c:394 */
  o.implicit = opts->implicit_orth; o.lag_enabled = opts->lag_enabled; o.lag_offset = opts->lag_offset;
  o.Jstart = opts->Jstart; o.Jstep = opts->Jstep; o.Jend = opts->Jend; o.lag_lower = opts->lag_lower; o.lag_upper = opts->lag_upper;
  o.BtBu = BtBu;

  /* QPTEnforceEqByPenalty(qp, rho, direct): qptransform.c:329-410; A_rho shell = matpenalized.c:212-243 */
  op_inner = *op;
  op_inner.m = m; op_inner.B = B; op_inner.rho = rho; op_inner.bimplicit = opts->implicit_orth;
  op_inner.bwork = (double *)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
  o.op_inner = &op_inner;
  v_copy(n, b, b_inner); /* :850-853 */

  mpgp_alloc(&in, &op_inner, b_inner, &bx, x, &opts->inner);
  o.inner = &in;
  maxeig_inner = ORC_MAX(rho, maxeig); /* :865 */
  {
    int inject = opts->inject_maxeig;
    if (!opts->inject_maxeig_set) inject = orth || opts->implicit_orth; /* :866; QPPFGetGHasOrthonormalRows: explicitly or implicitly (qppf.c:732-740) */
    if (inject) in.maxeig = maxeig_inner;        /* :868 */
  }
  mpgp_setup(&in); /* QPSSetUp(inner) :871 */
  in.conv = converged_inner_smalxe; /* :874-875 */
  in.conv_ctx = &o;

  /* --- QPSSolve_SMALXE: smalxe.c:893-997 --- */
  double t0 = now_seconds();
  o.M1 = o.M1_initial;
  v_set(n, Btmu, 0.0); /* :935 */
  if (opts->knoll) { /* :938-943  u = P b = b - G^T (G G^T)^{-1} G b */
    double *gl = (double *)malloc(sizeof(double) * (size_t)m), *y = (double *)malloc(sizeof(double) * (size_t)m);
    for (int j = 0; j < m; j++) gl[j] = orc_dot(n, B + (size_t)j * n, b);
    if (!orth) ggt_solve(n, m, B, gl, y);
    else memcpy(y, gl, sizeof(double) * (size_t)m);
    for (int k = 0; k < n; k++) {
      double s = 0.0;
      for (int j = 0; j < m; j++) s += B[(size_t)j * n + k] * y[j];
      x[k] = b[k] + -1.0 * s; /* VecAYPX(Pv,-1,v) qppf.c:572 */
    }
    free(gl); free(y);
  }
  Lag_old = orc_objective(&op_inner, b_inner, x); /* :946 */
  update_normBu(&o, x, &o.normBu_old, &o.enorm);  /* :949 */
  o.normBu_prev = o.normBu_old;
  o.iteration = 0; o.inner_iter_accu = 0; o.reason = ORC_CONVERGED_ITERATING;
  in.ncg = in.nexp = in.nmv = in.nprop = 0; /* QPSResetStatistics(inner) :955 */

  for (i = 0; i < o.max_it; i++) { /* :957 */
    /* QPSSMALXEUpdateLambda_SMALXE :402-435: Btmu += rho * BtB u */
    if (opts->implicit_orth) q_apply(n, m, B, x, BtBu, op_inner.bwork);
    else gtg_apply(n, m, B, x, BtBu, op_inner.bwork);
    v_axpy(n, Btmu, rho, BtBu);
    if (o.reason) break; /* :962 */
    v_waxpy(n, b_inner, -1.0, Btmu, b); /* :965 */
    in.divtol = o.divtol;               /* :968 */
    /* QPSConvergedSetUp_Inner_SMALXE :537-557 */
    o.norm_rhs_outer = orc_norm2(n, b);
    o.gtol           = o.rtol * o.norm_rhs_outer;
    o.ttol_outer     = ORC_MAX(o.rtol * o.norm_rhs_outer, o.atol);
    o.norm_rhs_div   = orc_norm2(n, b_inner); /* QPSConvergedDefaultSetRhsForDivergence qps.c:736-744 */
    mpgp_solve(&in); /* :970 */
    o.inner_iter_accu += in.iteration; /* :972 */
    o.iteration = i + 1;
    update_normBu(&o, x, &o.normBu, &o.enorm); /* :976 */
    rho = op_inner.rho;                        /* :979 */
    Lag = orc_objective(&op_inner, b_inner, x); /* :982 */
    { /* QPSSMALXEUpdate_SMALXE :439-488 */
      double t  = 0.5 * rho * o.normBu * o.normBu;
      double t2 = Lag - (Lag_old + t);
      int flag  = (t2 < o.update_threshold);
      if (flag && o.M1_update != 1.0) {
        if (in.reason == ORC_CONVERGED_ATOL) {
          o.M1 = o.M1 / o.M1_update;
          o.M1_updates++;
        }
      }
      if (!(in.rnorm > o.enorm)) { /* :482 */
        /* QPSSMALXEUpdateRho_SMALXE :373-398 */
        double rho_update = (o.state == 3) ? o.rho_update_late : o.rho_update;
        int    lagflag    = (o.state == 3) ? 1 : flag;
        if (lagflag && rho_update != 1.0) {
          op_inner.rho *= rho_update; /* MatPenalizedUpdatePenalty */
          /* QPSMPGPUpdateMaxEigenvalue: mpgp.c:119-134 */
          in.maxeig = in.maxeig * rho_update;
          if (!in.alpha_direct) in.alpha = in.alpha / rho_update;
          o.rho_updates++;
        }
      }
    }
    Lag_old      = Lag;
    o.normBu_old = o.normBu;
  }
  if (i == o.max_it && !o.reason) o.reason = ORC_DIVERGED_ITS; /* :986-989 */
  double t1 = now_seconds();

  if (lambda_out && opts->get_lambda) { /* QPPFApplyHalfQ(pf, Bt_lambda, lambda) :994 */
    double *gl = (double *)malloc(sizeof(double) * (size_t)m);
    for (int j = 0; j < m; j++) gl[j] = orc_dot(n, B + (size_t)j * n, Btmu);
    if (!orth) ggt_solve(n, m, B, gl, lambda_out);
    else memcpy(lambda_out, gl, sizeof(double) * (size_t)m);
    free(gl);
  }
  if (Bt_lambda_out) v_copy(n, Btmu, Bt_lambda_out);

  res->outer_its = o.iteration; res->reason = o.reason; res->inner_its_accu = o.inner_iter_accu;
  res->inner_reason_last = in.reason; res->M1_updates = o.M1_updates; res->M1_hits = o.M1_hits;
  res->eta_hits = o.eta_hits; res->rho_updates = o.rho_updates; res->state = o.state;
  res->nmv = in.nmv; res->ncg = in.ncg; res->nexp = in.nexp; res->nprop = in.nprop;
  res->rnorm = o.rnorm; res->normBu = o.normBu; res->M1 = o.M1; res->rho = op_inner.rho; res->maxeig = maxeig;
  res->maxeig_inner = in.maxeig; res->alpha_inner = in.alpha; res->eta = o.eta;
  res->seconds = t1 - t0;

  /* post-solve (qpchain.c:200-275): penalised child shares x; homogenised child: x_parent = x_child + xtilde */
  if (c) {
    v_waxpy(n, x_user, 1.0, x, xtilde); /* qptransform.c:419 */
    free(x); free(xtilde); free(b_h); free(lb_h); free(ub_h);
  }
  mpgp_free(&in);
  free(op_inner.bwork);
  free(Btmu); free(b_inner); free(BtBu); free(bw);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* QPSKSP and QPSPCPG                                                      */
/* ------------------------------------------------------------------------------------------ */
void orc_default_lin_opts(orc_lin_opts *o)
{
  o->rtol = 1e-5; o->atol = 1e-50; o->divtol = 1e4; o->max_it = 10000; /* qps.c:73-76 */
  o->nthreads = 1;
}

/* QPSConvergedDefault (qps.c:675-714) on plain scalars */
static int lin_converged(int i, double rnorm, const orc_lin_opts *o, double norm_rhs)
{
  const double ttol = ORC_MAX(o->rtol * norm_rhs, o->atol);
  if (i > o->max_it) return ORC_DIVERGED_ITS;
  if (isnan(rnorm) || isinf(rnorm)) return ORC_DIVERGED_NANORINF;
  if (rnorm <= ttol) return rnorm < o->atol ? ORC_CONVERGED_ATOL : ORC_CONVERGED_RTOL;
  if (rnorm >= o->divtol * norm_rhs) return ORC_DIVERGED_DTOL;
  return ORC_CONVERGED_ITERATING;
}

/* QPSSolve_KSP (qpsksp.c:137-153) hands the problem to PETSc's KSPCG configured in QPSCreate_KSP (qpsksp.c:232-253):
 * KSPCG, PCNONE, KSP_NORM_UNPRECONDITIONED, non-zero initial guess, convergence through QPSKSPConverged_KSP -> the QPS test
 * (qpsksp.c:5-14).  PETSc is not in the reference tree: what follows restates the published KSPSolve_CG recurrence
 * (petsc src/ksp/ksp/impls/cg/cg.c, single-reduction off, real symmetric) -- parity UNPINNED, the reference has no test or
 * golden output that runs QPSKSP.                                                                                       */
int orc_cg_solve(const orc_op *op, const double *b, double *x, const orc_lin_opts *opts, orc_lin_result *res)
{
  const int n = op->n;
  double   *R = (double *)malloc(sizeof(double) * (size_t)n), *P = (double *)malloc(sizeof(double) * (size_t)n),
         *W = (double *)malloc(sizeof(double) * (size_t)n);
  double beta, betaold = 1.0, dpi, a, dp, bb;
  int    i = 0, reason;
  orc_set_threads(opts->nthreads <= 0 ? 0 : opts->nthreads);
  const double norm_rhs = orc_norm2(n, b);
  const double t0 = now_seconds();
  orc_op_apply(op, x, R);      /* r = b - A x  (initial guess non-zero) */
  v_aypx(n, R, -1.0, b);
  dp     = orc_norm2(n, R);    /* z = r (PCNONE); unpreconditioned norm */
  reason = lin_converged(0, dp, opts, norm_rhs);
  beta   = orc_dot(n, R, R);   /* beta = z'r */
  if (!reason) {
    do {
      if (beta == 0.0) { reason = ORC_CONVERGED_ATOL; break; }
      if (!i) {
        v_copy(n, R, P);       /* p = z */
      } else {
        bb = beta / betaold;
        v_aypx(n, P, bb, R);   /* p = z + b p */
      }
      orc_op_apply(op, P, W);  /* w = A p */
      dpi     = orc_dot(n, P, W);
      betaold = beta;
      if (!(dpi > 0.0)) { reason = ORC_DIVERGED_INDEFINITE_MAT; break; }
      a = beta / dpi;
      v_axpy(n, x, a, P);      /* x += a p */
      v_axpy(n, R, -a, W);     /* r -= a w */
      dp     = orc_norm2(n, R);
      reason = lin_converged(i + 1, dp, opts, norm_rhs);
      i++;
      if (reason) break;
      beta = orc_dot(n, R, R);
    } while (i < opts->max_it);
    if (!reason) reason = ORC_DIVERGED_ITS;
  }
  res->its = i; res->reason = reason; res->rnorm = dp; res->norm_rhs = norm_rhs; res->seconds = now_seconds() - t0;
  free(R); free(P); free(W);
  return 0;
}

/* P v = v - G^T (G G^T)^{-1} G v : QPPFApplyP = QPPFApplyQ + VecAYPX (qppf.c:454-502,560-575) */
static void pf_apply_P(int n, int m, const double *G, int orth, const double *v, double *Pv, double *gl, double *y)
{
  for (int j = 0; j < m; j++) gl[j] = orc_dot(n, G + (size_t)j * n, v);
  if (!orth) ggt_solve(n, m, G, gl, y);
  else memcpy(y, gl, sizeof(double) * (size_t)m);
  for (int k = 0; k < n; k++) {
    double s = 0.0;
    for (int j = 0; j < m; j++) s += G[(size_t)j * n + k] * y[j];
    Pv[k] = v[k] + -1.0 * s;
  }
}

/* QPSSolve_PCPG: src/qps/impls/pcpg/pcpg.c:49-131 with PCNONE (y = w); QPSSetup_PCPG (:31-41) homogenises G x = c first
 * (QPTHomogenizeEq, qptransform.c:437-527: xtilde = G^T (G G^T)^{-1} c, b_bar = b - A xtilde, the child starts from a copy
 * of the parent's x, and the post-solve adds xtilde back).                                                              */
int orc_pcpg_solve(const orc_op *op, const double *b_user, int m, const double *G, const double *c, double *x_user, const orc_lin_opts *opts,
                   orc_lin_result *res)
{
  const int    n = op->n;
  const size_t nb = sizeof(double) * (size_t)n;
  double      *p = (double *)malloc(nb), *r = (double *)malloc(nb), *w = (double *)malloc(nb), *Ap = (double *)malloc(nb);
  double      *gl = (double *)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1)), *yy = (double *)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
  double      *xtilde = NULL, *bh = NULL, *x = x_user;
  const double *b = b_user;
  double       alpha, alpha1, beta, beta1 = 0.0, beta2, rnorm = 0.0;
  int          it = 0, reason = ORC_CONVERGED_ITERATING;
  orc_set_threads(opts->nthreads <= 0 ? 0 : opts->nthreads);
  const int orth = rows_orthonormal(n, m, G);
  if (c) {
    xtilde = (double *)malloc(nb);
    bh     = (double *)malloc(nb);
    if (!orth) ggt_solve(n, m, G, c, yy);
    else memcpy(yy, c, sizeof(double) * (size_t)m);
    for (int k = 0; k < n; k++) {
      double s = 0.0;
      for (int j = 0; j < m; j++) s += G[(size_t)j * n + k] * yy[j];
      xtilde[k] = s;
    }
    orc_op_apply(op, xtilde, bh);
    v_aypx(n, bh, -1.0, b_user);
    b = bh;
    x = (double *)malloc(nb);
    memcpy(x, x_user, nb);
  }
  const double norm_rhs = orc_norm2(n, b);
  const double t0 = now_seconds();
  orc_op_apply(op, x, r);   /* :95-96  r = b - A lm */
  v_aypx(n, r, -1.0, b);
  do {
    pf_apply_P(n, m, G, orth, r, w, gl, yy);             /* :100 */
    rnorm  = orc_norm2(n, w);                            /* :103 */
    reason = lin_converged(it, rnorm, opts, norm_rhs);   /* :104 */
    if (reason) break;
    beta2 = beta1;                                       /* :113  (y = w) */
    beta1 = orc_dot(n, w, w);
    if (!it) {
      beta = 0;
      v_copy(n, w, p);                                   /* :117 */
    } else {
      beta = beta1 / beta2;
      v_aypx(n, p, beta, w);                             /* :120 */
    }
    orc_op_apply(op, p, Ap);                             /* :122 */
    alpha1 = orc_dot(n, p, Ap);
    alpha  = beta1 / alpha1;
    v_axpy(n, x, alpha, p);                              /* :125 */
    v_axpy(n, r, -alpha, Ap);
    it++;
  } while (it < opts->max_it);                           /* :129 */
  if (c) { /* QPTHomogenizeEqPostSolve: x = x_child + xtilde */
    for (int k = 0; k < n; k++) x_user[k] = x[k] + xtilde[k];
    free(x); free(xtilde); free(bh);
  }
  res->its = it; res->reason = reason; res->rnorm = rnorm; res->norm_rhs = norm_rhs; res->seconds = now_seconds() - t0;
  free(p); free(r); free(w); free(Ap); free(gl); free(yy);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* QPTOrthonormalizeEq / QPTHomogenizeEq / projector pieces                 */
/* ------------------------------------------------------------------------------------------ */
int orc_rows_orthonormal(int n, int m, const double *B) { return rows_orthonormal(n, m, B); }

void orc_apply_P(int n, int m, const double *G, const double *v, double *Pv)
{
  double gl[8], yy[8];
  pf_apply_P(n, m, G, rows_orthonormal(n, m, G), v, Pv, gl, yy);
}

/* MatOrthRows (src/mat/interface/permonmatorth.c:494-520) = MatOrthColumns on the transpose; explicit forms.
 * type 1: MatOrthColumns_GS_Default (:196-234), iterated classical Gram-Schmidt with re-orthogonalisation while the norm drops
 *         below alpha = 0.5 of its previous value; S accumulates the same operations on the identity, T = S^T.
 * type 3: MatOrthColumns_Cholesky_Default (:33-150): G G^T = L L^T, TB = L^{-1} G (forward solve of every column), T = L^{-1}.
 * T is m x m row-major; Tc = T c (qptransform.c:610-613).                                                                   */
int orc_orth_rows(int n, int m, const double *B, const double *c, int type, double *TB, double *Tc, double *T)
{
  memcpy(TB, B, sizeof(double) * (size_t)m * n);
  for (int i = 0; i < m; i++)
    for (int j = 0; j < m; j++) T[i * m + j] = (i == j) ? 1.0 : 0.0;
  if (type == 1) {
    double dots[8];
    for (int i = 0; i < m; i++) {
      double *q = TB + (size_t)i * n;
      double  norm = orc_norm2(n, q), norm_last;
      do {
        norm_last = norm;
        for (int j = 0; j < i; j++) dots[j] = -orc_dot(n, q, TB + (size_t)j * n); /* VecMDot, negated :220-221 */
        for (int j = 0; j < i; j++) v_axpy(n, q, dots[j], TB + (size_t)j * n);    /* VecMAXPY :222 */
        for (int j = 0; j < i; j++)
          for (int k = 0; k < m; k++) T[i * m + k] += dots[j] * T[j * m + k];    /* the same on s[i] :223 */
        norm = orc_norm2(n, q);
        if (norm < 1e2 * 2.220446049250313e-16) return 1;                          /* :227 */
      } while (norm <= 0.5 * norm_last);
      v_scale(n, q, 1.0 / norm);
      for (int k = 0; k < m; k++) T[i * m + k] *= 1.0 / norm;
    }
  } else if (type == 3) {
    double L[64];
    for (int i = 0; i < m; i++)
      for (int j = 0; j <= i; j++) L[i * m + j] = orc_dot(n, B + (size_t)i * n, B + (size_t)j * n);
    for (int j = 0; j < m; j++) {
      double d = L[j * m + j];
      for (int k = 0; k < j; k++) d -= L[j * m + k] * L[j * m + k];
      if (d <= 0.0) return 1;
      d = sqrt(d);
      L[j * m + j] = d;
      for (int i = j + 1; i < m; i++) {
        double v = L[i * m + j];
        for (int k = 0; k < j; k++) v -= L[i * m + k] * L[j * m + k];
        L[i * m + j] = v / d;
      }
    }
    /* forward solve L Y = G (column by column of G) and L T = I */
    for (int col = 0; col < n; col++)
      for (int i = 0; i < m; i++) {
        double v = B[(size_t)i * n + col];
        for (int k = 0; k < i; k++) v -= L[i * m + k] * TB[(size_t)k * n + col];
        TB[(size_t)i * n + col] = v / L[i * m + i];
      }
    for (int col = 0; col < m; col++)
      for (int i = 0; i < m; i++) {
        double v = (i == col) ? 1.0 : 0.0;
        for (int k = 0; k < i; k++) v -= L[i * m + k] * T[k * m + col];
        T[i * m + col] = v / L[i * m + i];
      }
  } else {
    return 2;
  }
  if (c && Tc)
    for (int i = 0; i < m; i++) {
      double s = 0.0;
      for (int k = 0; k < m; k++) s += T[i * m + k] * c[k];
      Tc[i] = s;
    }
  return 0;
}

void orc_homogenize(const orc_op *op, const double *b, const orc_box *bx, int m, const double *G, const double *c, double *xtilde, double *b_h,
                    double *lb_h, double *ub_h)
{
  const int n = op->n;
  double    y[8];
  if (!rows_orthonormal(n, m, G)) ggt_solve(n, m, G, c, y); /* QPPFApplyHalfQTranspose qppf.c:535-568 */
  else memcpy(y, c, sizeof(double) * (size_t)m);
  for (int k = 0; k < n; k++) {
    double s = 0.0;
    for (int j = 0; j < m; j++) s += G[(size_t)j * n + k] * y[j];
    xtilde[k] = s;
  }
  orc_op_apply(op, xtilde, b_h); /* qptransform.c:468-469 */
  v_aypx(n, b_h, -1.0, b);
  if (bx) {
    const int ns = NSUB(bx);
    if (bx->lb && lb_h)
      for (int k = 0; k < ns; k++) lb_h[k] = bx->lb[k] - xtilde[IDX(bx, k)]; /* :498-501 */
    if (bx->ub && ub_h)
      for (int k = 0; k < ns; k++) ub_h[k] = bx->ub[k] - xtilde[IDX(bx, k)]; /* :503-506 */
  }
}
