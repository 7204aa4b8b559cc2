/* TEST INFRASTRUCTURE: "PERMON user code" on the mock PETSc (adapters/mock) that solves a QP through adapters/qpsb200.c, i.e. through
 * QPSRegister -> QPSSetType("mpgp" | "smalxe") -> _QPSOps::setup / solve -> dlopen(libpermon_b200.so).  The call sequence is the reference
 * tutorial's (src/tutorials/ex1.c:108-157).  usage: adapter_driver problem.bin x_out.bin type [-option value ...] */
#include "../adapters/mock/permon_mock.h"

PetscErrorCode PermonB200RegisterQPS(void);

static void *rd(FILE *f, size_t bytes)
{
  void *p = malloc(bytes ? bytes : 1);
  if (bytes && fread(p, 1, bytes, f) != bytes) exit(3);
  return p;
}

int main(int argc, char **argv)
{
  if (argc < 4) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  int hdr[5];   /* n, nnz, has_lb, has_ub, has_eq */
  if (fread(hdr, sizeof(int), 5, f) != 5) return 3;
  const int n = hdr[0], nnz = hdr[1];
  int    *ia = rd(f, sizeof(int) * (size_t)(n + 1)), *ja = rd(f, sizeof(int) * (size_t)nnz);
  double *a = rd(f, 8 * (size_t)nnz), *b = rd(f, 8 * (size_t)n), *x = rd(f, 8 * (size_t)n);
  double *lb = hdr[2] ? rd(f, 8 * (size_t)n) : NULL, *ub = hdr[3] ? rd(f, 8 * (size_t)n) : NULL, *be = hdr[4] ? rd(f, 8 * (size_t)n) : NULL;
  fclose(f);
  for (int k = 4; k + 1 < argc; k += 2) PetscOptionsSetValue(NULL, argv[k], argv[k + 1]);

  Mat A, BE = NULL;
  Vec vb, vx, vlb = NULL, vub = NULL, vbe = NULL;
  QP  qp;
  QPS qps;
  PetscCall(PermonB200RegisterQPS());   /* overrides "mpgp" / "smalxe" (qpsregis.c:29-36) */
  PetscCall(MatCreateSeqAIJWithArrays(PETSC_COMM_WORLD, n, n, ia, ja, a, &A));
  PetscCall(VecCreateSeqWithArray(PETSC_COMM_WORLD, 1, n, b, &vb));
  PetscCall(VecCreateSeqWithArray(PETSC_COMM_WORLD, 1, n, x, &vx));
  if (lb) PetscCall(VecCreateSeqWithArray(PETSC_COMM_WORLD, 1, n, lb, &vlb));
  if (ub) PetscCall(VecCreateSeqWithArray(PETSC_COMM_WORLD, 1, n, ub, &vub));
  PetscCall(QPCreate(PETSC_COMM_WORLD, &qp));
  PetscCall(QPSetOperator(qp, A));
  PetscCall(QPSetRhs(qp, vb));
  PetscCall(QPSetInitialVector(qp, vx));
  PetscCall(QPSetBox(qp, NULL, vlb, vub));
  if (be) {
    PetscCall(VecCreateSeqWithArray(PETSC_COMM_WORLD, 1, n, be, &vbe));
    PetscCall(MatCreateOneRow(vbe, &BE));
    PetscCall(QPSetEq(qp, BE, NULL));
  }
  PetscCall(QPSCreate(PETSC_COMM_WORLD, &qps));
  PetscCall(QPSSetType(qps, argv[3]));
  PetscCall(QPSSetQP(qps, qp));
  PetscCall(QPSSolve(qps));
  printf("reason %d iterations %d rnorm %.10e\n", (int)qps->reason, (int)qps->iteration, (double)qps->rnorm);
  PetscCall(QPSViewConvergence(qps, PETSC_VIEWER_STDOUT_WORLD));
  f = fopen(argv[2], "wb");
  if (!f || fwrite(x, 8, (size_t)n, f) != (size_t)n) return 4;
  fclose(f);
  PetscCall(QPSDestroy(&qps));
  PetscCall(QPDestroy(&qp));
  PetscCall(MatDestroy(&A));
  if (BE) PetscCall(MatDestroy(&BE));
  PetscCall(VecDestroy(&vb)); PetscCall(VecDestroy(&vx)); PetscCall(VecDestroy(&vlb)); PetscCall(VecDestroy(&vub)); PetscCall(VecDestroy(&vbe));
  return 0;
}
