// mpgp_ctl.h -- scalar control state of the device-resident MPGP iteration and the step-selection /
// stopping logic that runs on it.  The functions are __host__ __device__: on the device they run in the
// one-thread ctrl kernels between the big kernels (no host round trip per iteration); the host uses the
// very same code when a user callback (monitor / custom convergence test) forces a per-iteration sync.
//
// Reference behaviour restated here (file:line into permon/permon):
//   step selection            src/qps/impls/mpgp/mpgp.c:535-547,617-621
//   QPSConvergedDefault       src/qps/interface/qps.c:675-714
//   QPSConverged_Inner_SMALXE src/qps/impls/smalxe/smalxe.c:610-692
#pragma once
#include <math.h>

#ifndef __CUDACC__
#define PB_HD
#else
#define PB_HD __host__ __device__
#endif

#define PB_MAXEQ 4       // equality rows handled by the fused rank-m update (accumulators in registers)
#define PB_MAXEQ_ALL 64  // equality rows handled at all: beyond PB_MAXEQ the un-fused route runs (dense rows, 8 per reduction launch)
#define PB_NRED 8    // doubles per reduction record
#define PB_MAXRANKS 16

// reduction record after the SpMV kernel K_A:   [pAp, g.p, Bp_0..3, afeas(min), unused]
//   g.p is not computed by K_A itself: whoever produces the direction p knows it.  After p = gf (initial gradient, expansion,
//   proportioning update) g.p = |gf|^2 and after p = gc it is |gc|^2, term by term (gf_i is g_i or 0) -- ctrl_B keeps that
//   value (gp_known); after p = gf - beta p (CG) K_C accumulates g.p while it writes p and K_A's record carries that partial sum.
// reduction record after K_B / K_A':            [gP2, gc2, gf2, Ap.gf, Bu_0..3]
enum { RA_PAP = 0, RA_GP = 1, RA_BP = 2, RA_FEAS = 6 };
enum { RB_GP2 = 0, RB_GC2 = 1, RB_GF2 = 2, RB_APGF = 3, RB_BU = 4 };

enum { PB_REASON_ITERATING = 0, PB_RTOL = 2, PB_ATOL = 3, PB_HAPPY = 7, PB_DIV_ITS = -3, PB_DIV_DTOL = -4, PB_DIV_BREAKDOWN = -5, PB_DIV_NAN = -9 };

struct MpgpCtl {
  // ---- configuration, written by the host before a solve ------------------------------------------
  int    max_it;
  int    inner_mode;      // 0: QPSConvergedDefault, 1: QPSConverged_Inner_SMALXE
  int    host_conv;       // 1: the host runs monitors / the convergence test (ctrl_B leaves `reason` alone)
  int    nranks, m;       // ranks in the rank-ordered reductions; equality rows in the fused rank-m update
  int    serp;            // 1: serpentine sweeps (consecutive kernels walk the rows in opposite directions)
  double gamma2, alpha, rho;
  double rtol, atol, ttol, divtol, norm_rhs_div;          // QPSConvergedDefaultCtx (qpsimpl.h:73-76)
  // SMALXE inner test (smalxeimpl.h:5-11 and the outer QPS fields it reads)
  double M1, eta, rtol_E, gtol;
  double outer_ttol, outer_atol, outer_divtol, outer_norm_rhs_div;
  int    outer_max_it, outer_iteration, smalxe_state, inner_iter_min, inner_no_gtol_stop, inner_iter_accu;
  // ---- status ---------------------------------------------------------------------------------------
  int    iteration, reason, outer_reason;
  int    step;            // ' ', 'c', 'e', 'p'
  int    do_prop;         // the coming step is a proportioning step
  int    pmode;           // what K_C does: 0 nothing, 1 p = gf - bcg p, 2 p = gc
  int    init;            // the initial gradient evaluation is in flight
  int    gp_known;        // g.p of the coming K_A is gp_next (p = gf or p = gc), not the record sum
  int    sweep;           // direction of the coming K_A sweep (0 ascending rows); K_B runs the other way, K_A' / K_C this way again
  int    nmv, ncg, nexp, nprop, M1_hits, eta_hits;
  double rnorm, gP2, gc2, gf2, Apgf, pAp, gp, afeas, acg, bcg, gp_next;
  double normBu, enorm, outer_rnorm, MNormBu;
  double Bp[PB_MAXEQ], Bu[PB_MAXEQ];
};

// ctrl_A: runs after the SpMV + dots kernel.  `ra` holds one record per rank, in rank order.
PB_HD inline void mpgp_ctrl_A(MpgpCtl *S, const double *ra)
{
  if (S->reason != PB_REASON_ITERATING) return;
  double pAp = 0.0, gp = 0.0, afeas = HUGE_VAL, bp[PB_MAXEQ];
  for (int j = 0; j < PB_MAXEQ; j++) bp[j] = 0.0;
  for (int r = 0; r < S->nranks; r++) {
    const double *q = ra + r * PB_NRED;
    pAp += q[RA_PAP];
    gp += q[RA_GP];   // K_C's partial sums of g.p (CG direction update), carried by the K_A records
    for (int j = 0; j < S->m; j++) bp[j] += q[RA_BP + j];
    if (q[RA_FEAS] < afeas) afeas = q[RA_FEAS];
  }
  for (int j = 0; j < S->m; j++) {  // p'(A + rho B'B)p = p'Ap + rho |Bp|^2   (matpenalized.c:12-22)
    pAp += S->rho * bp[j] * bp[j];
    S->Bp[j] = bp[j];
  }
  if (S->gp_known) gp = S->gp_next;   // p = gf / p = gc: VecDot(g,p) is |gf|^2 / |gc|^2
  S->pAp   = pAp;
  S->gp    = gp;
  S->afeas = afeas;
  S->acg   = gp / pAp;   // mpgp.c:543,630
  S->nmv++;
  if (S->do_prop) {      // mpgp.c:617-621
    S->step = 'p';
    S->nprop++;
  } else if (S->acg <= afeas) {  // mpgp.c:547
    S->step = 'c';
    S->ncg++;
  } else {
    S->step = 'e';
    S->nexp++;
  }
}

// ctrl_E: after the expansion half of K_B, before the second SpMV: B u of the new iterate.
PB_HD inline void mpgp_ctrl_E(MpgpCtl *S, const double *rb)
{
  if (S->reason != PB_REASON_ITERATING) return;
  if (!(S->step == 'e' || S->init)) return;
  for (int j = 0; j < S->m; j++) {
    double s = 0.0;
    for (int r = 0; r < S->nranks; r++) s += rb[r * PB_NRED + RB_BU + j];
    S->Bu[j] = s;
  }
}

PB_HD inline bool pb_isnanorinf(double v) { return !(v == v) || v > 1.7976931348623157e308 || v < -1.7976931348623157e308; }

// QPSConvergedDefault (qps.c:675-714) evaluated on explicit numbers
PB_HD inline int pb_converged_default(int i, double rnorm, int max_it, double ttol, double atol, double divtol, double norm_rhs_div)
{
  if (i > max_it) return PB_DIV_ITS;
  if (pb_isnanorinf(rnorm)) return PB_DIV_NAN;
  if (rnorm <= ttol) return (rnorm < atol) ? PB_ATOL : PB_RTOL;
  if (rnorm >= divtol * norm_rhs_div) return PB_DIV_DTOL;
  return PB_REASON_ITERATING;
}

// QPSConverged_Inner_SMALXE (smalxe.c:610-692)
PB_HD inline void pb_converged_inner_smalxe(MpgpCtl *S)
{
  const int    i = S->iteration;
  const double gnorm = S->rnorm;
  double       nb = 0.0;
  for (int j = 0; j < S->m; j++) nb += S->Bu[j] * S->Bu[j];   // VecNorm(Bu) :256-258 (c = 0 after homogenisation)
  S->normBu      = sqrt(nb);
  S->enorm       = S->normBu / S->rtol_E;
  S->outer_rnorm = (S->enorm < gnorm) ? gnorm : S->enorm;     // :626
  S->MNormBu     = S->M1 * S->normBu;                         // :627
  S->atol        = (S->MNormBu < S->eta) ? S->MNormBu : S->eta;  // :628
  S->reason      = PB_REASON_ITERATING;
  if (i > S->max_it - S->inner_iter_accu) {  // :633
    S->reason       = PB_DIV_ITS;
    S->outer_reason = PB_DIV_BREAKDOWN;
    return;
  }
  if (pb_isnanorinf(gnorm)) {  // :641
    S->reason       = PB_DIV_NAN;
    S->outer_reason = PB_DIV_BREAKDOWN;
    return;
  }
  S->outer_reason = pb_converged_default(S->outer_iteration, S->outer_rnorm, S->outer_max_it, S->outer_ttol, S->outer_atol, S->outer_divtol, S->outer_norm_rhs_div);  // :648
  if (S->outer_reason) {  // :650-659
    S->reason = (S->outer_reason > 0) ? PB_HAPPY : PB_DIV_BREAKDOWN;
    return;
  }
  if (gnorm < S->atol) {  // :661-671
    S->reason = PB_ATOL;
    if (S->MNormBu < S->eta) S->M1_hits++;
    else S->eta_hits++;
    return;
  }
  if (S->smalxe_state == 3 && (i < S->inner_iter_min || S->inner_no_gtol_stop)) return;  // :673
  if (gnorm <= S->gtol) {  // :675-690
    if (!(gnorm > S->enorm)) {
      if (S->inner_no_gtol_stop < 2) S->reason = PB_RTOL;
      if (S->smalxe_state != 3) S->smalxe_state = 3;
    }
  }
}

// ctrl_B: runs after K_B (steps 'c','p') or after the second SpMV K_A' (step 'e' and the initial gradient).
PB_HD inline void mpgp_ctrl_B(MpgpCtl *S, const double *rb)
{
  if (S->reason != PB_REASON_ITERATING) return;
  double gP2 = 0.0, gc2 = 0.0, gf2 = 0.0, apgf = 0.0;
  for (int r = 0; r < S->nranks; r++) {
    const double *q = rb + r * PB_NRED;
    gP2 += q[RB_GP2];
    gc2 += q[RB_GC2];
    gf2 += q[RB_GF2];
    apgf += q[RB_APGF];
  }
  if (!(S->step == 'e' || S->init)) {   // Bu of the new iterate came with this record ('e'/init: ctrl_E had it)
    for (int j = 0; j < S->m; j++) {
      double s = 0.0;
      for (int r = 0; r < S->nranks; r++) s += rb[r * PB_NRED + RB_BU + j];
      S->Bu[j] = s;
    }
  }
  if (S->step == 'e' || S->init) S->nmv++;   // mpgp.c:501,579
  S->gP2 = gP2; S->gc2 = gc2; S->gf2 = gf2; S->Apgf = apgf;
  S->rnorm = sqrt(gP2);                     // mpgp.c:514
  S->bcg   = apgf / S->pAp;                 // mpgp.c:558-559
  if (!S->init) S->iteration++;             // mpgp.c:640
  const int was_cg = (S->step == 'c' && !S->init);
  S->init = 0;
  if (!S->host_conv) {
    if (S->inner_mode) pb_converged_inner_smalxe(S);
    else S->reason = pb_converged_default(S->iteration, S->rnorm, S->max_it, S->ttol, S->atol, S->divtol, S->norm_rhs_div);
  }
  S->do_prop = !(gc2 <= S->gamma2 * gf2);   // mpgp.c:535
  S->pmode   = S->do_prop ? 2 : (was_cg ? 1 : 0);
  if (S->reason != PB_REASON_ITERATING) S->pmode = 0;
  // VecDot(g, p) of the coming iteration (mpgp.c:541): p = gc -> sum g_i gc_i = |gc|^2 ; p = gf -> |gf|^2 ; CG direction: K_C sums it
  S->gp_known = (S->pmode != 1);
  S->gp_next  = (S->pmode == 2) ? gc2 : gf2;
  S->sweep ^= S->serp;
}
