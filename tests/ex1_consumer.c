/* TEST INFRASTRUCTURE: the reference tutorial src/tutorials/ex1.c:58-157 as a plain C program against include/permonqps.h, linked to
 * libpermon_b200.so (the drop-in boundary).  The PERMON calls are the tutorial's, in the tutorial's order; only the PETSc
 * matrix / vector ASSEMBLY (MatSetValues / VecSetValue, which this library does not stand in for) is replaced by filling host arrays and
 * handing them to MatCreateSeqAIJWithArrays / VecCreateSeqWithArray.  Run as the reference's test does:
 *     ex1_consumer -n 100 -qps_view_convergence -qp_chain_view_kkt [-qps_mpgp_expansion_type gf ...] | grep -e CONVERGED -e number -e "r ="  */
#include <math.h>
#include <permonqps.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHK(call)                                                                                                    \
  do {                                                                                                               \
    PetscErrorCode ierr_ = (call);                                                                                   \
    if (ierr_) {                                                                                                     \
      fprintf(stderr, "error %d in %s: %s\n", (int)ierr_, #call, PermonB200GetLastErrorMessage());                   \
      return (int)ierr_;                                                                                             \
    }                                                                                                                \
  } while (0)

static PetscReal fobst(PetscInt i, PetscInt n)
{   /* ex1.c:24-28 */
  PetscReal h = 1. / (n - 1);
  return sin(4 * M_PI * i * h - M_PI / 6.) / 2 - 2;
}

int main(int argc, char **args)
{
  Vec       b, c, x;
  Mat       A;
  QP        qp;
  QPS       qps;
  PetscInt  i, n = 10;
  PetscBool converged;

  CHK(PermonInitialize(&argc, &args, (char *)0, NULL));
  for (int k = 1; k + 1 < argc; k++)
    if (!strcmp(args[k], "-n")) n = atoi(args[k + 1]);

  /* ex1.c:65-106: tridiag(-1, 2, -1) with identity first / last rows and the couplings to them dropped */
  const PetscScalar h = 1. / (n - 1);
  PetscInt    *ia = malloc(sizeof(PetscInt) * (size_t)(n + 1)), *ja = malloc(sizeof(PetscInt) * (size_t)(3 * n));
  PetscScalar *a = malloc(sizeof(PetscScalar) * (size_t)(3 * n)), *bh = calloc((size_t)n, sizeof(PetscScalar)), *ch = calloc((size_t)n, sizeof(PetscScalar)),
              *xh = calloc((size_t)n, sizeof(PetscScalar));
  PetscInt nz = 0;
  for (i = 0; i < n; i++) {
    ia[i] = nz;
    if (i == 0 || i == n - 1) {
      ja[nz] = i; a[nz++] = 1.0;
      continue;
    }
    if (i != 1) { ja[nz] = i - 1; a[nz++] = -1.0; }
    ja[nz] = i; a[nz++] = 2.0;
    if (i != n - 2) { ja[nz] = i + 1; a[nz++] = -1.0; }
    bh[i] = -15 * h * h * 2;
    ch[i] = fobst(i, n);
  }
  ia[n] = nz;
  CHK(MatCreateSeqAIJWithArrays(PETSC_COMM_WORLD, n, n, ia, ja, a, &A));
  CHK(VecCreateSeqWithArray(PETSC_COMM_WORLD, 1, n, bh, &b));
  CHK(VecCreateSeqWithArray(PETSC_COMM_WORLD, 1, n, ch, &c));
  CHK(VecCreateSeqWithArray(PETSC_COMM_WORLD, 1, n, xh, &x));

  /* ex1.c:108-141, verbatim */
  CHK(QPCreate(PETSC_COMM_WORLD, &qp));
  CHK(QPSetOperator(qp, A));
  CHK(QPSetRhs(qp, b));
  CHK(QPSetInitialVector(qp, x));
  CHK(QPSetBox(qp, NULL, c, NULL));
  CHK(QPSetFromOptions(qp));
  CHK(QPSCreate(PETSC_COMM_WORLD, &qps));
  CHK(QPSSetQP(qps, qp));
  CHK(QPSSetFromOptions(qps));
  CHK(QPSSolve(qps));
  CHK(QPIsSolved(qp, &converged));
  if (!converged) printf("QPS did not converge!\n");
  {
    const PetscScalar *sol;
    PetscScalar        s = 0.0;
    CHK(VecGetArrayRead(x, &sol));   /* the user's x is the solution storage: sol == xh */
    for (i = 0; i < n; i++) s += sol[i];
    fprintf(stderr, "sum(x) = %.15e  same_storage = %d\n", (double)s, (int)(sol == xh));
    CHK(VecRestoreArrayRead(x, &sol));
  }
  CHK(QPSDestroy(&qps));
  CHK(QPDestroy(&qp));
  CHK(VecDestroy(&x));
  CHK(VecDestroy(&c));
  CHK(VecDestroy(&b));
  CHK(MatDestroy(&A));
  CHK(PermonFinalize());
  free(ia); free(ja); free(a); free(bh); free(ch); free(xh);
  return 0;
}
