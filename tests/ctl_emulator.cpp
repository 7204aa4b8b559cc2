// TEST INFRASTRUCTURE (CPU): drives the product's control logic -- permon_b200/csrc/mpgp_ctl.h, the very functions the CUDA kernels run in
// their prologues -- through the fused K_A -> K_B -> K_A' -> K_C schedule with the big kernels replaced by plain loops over host arrays.
// It checks, without a GPU, that the step selection / stopping / direction-mode protocol of the device-driven iteration reproduces the
// reference's golden counts.  Not part of the product and never loaded by it.
#include <cmath>
#include <cstring>
#include <vector>

#include "../permon_b200/csrc/mpgp_ctl.h"

namespace {
const double PINF = 1.7976931348623157e+308 / 4.0;
struct Prob {
  int           n;
  const int    *ia, *ja;
  const double *a, *b, *lb, *ub;
  double        astol;
};
void spmv(const Prob &P, const double *x, double *y)
{
  for (int r = 0; r < P.n; r++) {
    double s = 0.0;
    for (int k = P.ia[r]; k < P.ia[r + 1]; k++) s += P.a[k] * x[P.ja[k]];
    y[r] = s;
  }
}
void split(const Prob &P, int r, double x, double g, double &gf, double &gc)
{   // QPCGrads_Box qpcbox.c:41-55
  gf = g;
  gc = 0.0;
  if (P.lb && fabs(x - P.lb[r]) <= P.astol) {
    gf = 0.0;
    gc = g < 0.0 ? g : 0.0;
  } else if (P.ub && fabs(x - P.ub[r]) <= P.astol) {
    gf = 0.0;
    gc = g > 0.0 ? g : 0.0;
  }
}
double reduced(const Prob &P, int r, double x, double gf, double alpha)
{   // QPCGradReduced_Box qpcbox.c:86-92
  if (P.lb && gf > 0.0) {
    const double t = (x - P.lb[r]) / alpha;
    return gf < t ? gf : t;
  }
  if (P.ub && gf < 0.0) {
    const double t = (x - P.ub[r]) / alpha;
    return gf < t ? t : gf;
  }
  return gf;
}
}   // namespace

extern "C" int ctl_emulate(int n, const int *ia, const int *ja, const double *a, const double *b, const double *lb, const double *ub, double *x, double rtol,
                           double atol, double divtol, int max_it, double alpha, int *out)
{
  Prob                P{n, ia, ja, a, b, lb, ub, 10 * 2.220446049250313e-16};
  std::vector<double> g(n), p(n), Ap(n), gfv(n);
  MpgpCtl             S;
  memset(&S, 0, sizeof S);
  double nb = 0.0;
  for (int r = 0; r < n; r++) nb += b[r] * b[r];
  nb = sqrt(nb);
  S.max_it = max_it; S.nranks = 1; S.m = 0; S.gamma2 = 1.0; S.alpha = alpha;
  S.rtol = rtol; S.atol = atol; S.divtol = divtol; S.ttol = (rtol * nb > atol) ? rtol * nb : atol; S.norm_rhs_div = nb;
  S.step = ' '; S.init = 1;
  double ra[PB_NRED], rb[PB_NRED];
  // ---- initial phase: x = P(x); K_A': g = A x - b, split, p = gf ; ctrl_B ; K_C
  for (int r = 0; r < n; r++) {
    if (lb && x[r] < lb[r]) x[r] = lb[r];
    if (ub && x[r] > ub[r]) x[r] = ub[r];
  }
  auto second_spmv = [&]() {
    spmv(P, x, g.data());
    memset(rb, 0, sizeof rb);
    for (int r = 0; r < n; r++) {
      g[r] -= b[r];
      double gf, gc;
      split(P, r, x[r], g[r], gf, gc);
      p[r] = gf;
      const double gP = gf + gc;
      rb[RB_GP2] += gP * gP; rb[RB_GC2] += gc * gc; rb[RB_GF2] += gf * gf;
    }
  };
  double gp_from_C = 0.0;   // K_C's sum of g.p (the K_A record carries it)
  auto step_C = [&]() {
    if (S.reason != 0) return;
    gp_from_C = 0.0;
    if (S.pmode == 1) {
      for (int r = 0; r < n; r++) {
        p[r] = gfv[r] - S.bcg * p[r];
        gp_from_C += g[r] * p[r];
      }
    } else if (S.pmode == 2) {
      for (int r = 0; r < n; r++) {
        double gf, gc;
        split(P, r, x[r], g[r], gf, gc);
        p[r] = gc;
      }
    }
  };
  second_spmv();
  mpgp_ctrl_E(&S, rb);
  mpgp_ctrl_B(&S, rb);
  step_C();
  // ---- main loop
  while (S.reason == 0) {
    // K_A
    spmv(P, p.data(), Ap.data());
    memset(ra, 0, sizeof ra);
    ra[RA_FEAS] = HUGE_VAL;
    for (int r = 0; r < n; r++) {
      ra[RA_PAP] += p[r] * Ap[r];
      if (p[r] > 0. && lb && lb[r] > -PINF) {
        const double t = (x[r] - lb[r]) / p[r];
        if (t < ra[RA_FEAS]) ra[RA_FEAS] = t;
      }
      if (p[r] < 0. && ub && ub[r] < PINF) {
        const double t = (x[r] - ub[r]) / p[r];
        if (t < ra[RA_FEAS]) ra[RA_FEAS] = t;
      }
    }
    ra[RA_GP] = gp_from_C;
    mpgp_ctrl_A(&S, ra);
    // K_B
    memset(rb, 0, sizeof rb);
    if (S.step == 'e') {
      for (int r = 0; r < n; r++) {   // MPGPExpansion_Std mpgp.c:316-321
        const double xh = x[r] - S.afeas * p[r], gh = g[r] - S.afeas * Ap[r];
        double       gf, gc;
        split(P, r, xh, gh, gf, gc);
        x[r] = xh - S.alpha * reduced(P, r, xh, gf, S.alpha);
      }
      mpgp_ctrl_E(&S, rb);
      second_spmv();
    } else {
      for (int r = 0; r < n; r++) {
        x[r] -= S.acg * p[r];
        g[r] -= S.acg * Ap[r];
        double gf, gc;
        split(P, r, x[r], g[r], gf, gc);
        if (S.step == 'c') {
          rb[RB_APGF] += Ap[r] * gf;
          gfv[r] = gf;
        } else {
          p[r] = gf;
        }
        const double gP = gf + gc;
        rb[RB_GP2] += gP * gP; rb[RB_GC2] += gc * gc; rb[RB_GF2] += gf * gf;
      }
    }
    mpgp_ctrl_B(&S, rb);
    step_C();
  }
  out[0] = S.iteration; out[1] = S.nmv; out[2] = S.ncg; out[3] = S.nexp; out[4] = S.nprop; out[5] = S.reason;
  return 0;
}
