#!/bin/bash
# compute-sanitizer target (race / memory checking): the smoke solve (reference tutorial ex1 + a 2-D obstacle problem with CG,
# expansion and proportioning steps, through the fused device-driven driver) and a small SMALXE + generic-driver solve under
#   memcheck  (out-of-bounds / misaligned global, shared and local accesses, leaks of device memory are not reported: the pool keeps buffers)
#   racecheck (shared-memory hazards: block reductions, mbarrier rings of the TMA kernels)
#   synccheck (invalid __syncthreads / __syncwarp / mbarrier usage)
# usage: tests/run_sanitizer.sh [outdir=gpurun_out]      -> <outdir>/sanitizer_<tool>.log, exit code 0 only if every tool reports 0 errors
out=${1:-gpurun_out}; mkdir -p $out; rc=0
cat > /tmp/sanitizer_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as G
from permon_b200 import api as P, problems as PR
G.smoke()
# packed dictionary tiles (variable coefficients, both bounds) + SMALXE with one equality row (fused) + the generic driver
pr = PR.varcoef3d(12)
r = P.solve_problem(pr, "mpgp", "-qps_rtol 1e-6")
print("varcoef3d", r.its, r.reason)
pr = PR.obstacle2d(24); n = pr.n
pr.B = np.full((1, n), 1.0 / np.sqrt(n)); pr.c = np.array([-0.05 * np.sqrt(n)])
r = P.solve_problem(pr, "smalxe", "-qps_rtol 1e-6")
print("smalxe", r.its, r.reason)
r = P.solve_problem(PR.obstacle2d(32, -100.0), "mpgp", "-qps_rtol 1e-6 -qps_mpgp_b200_driver generic -qps_mpgp_expansion_type projcg")
print("generic projcg", r.its, r.reason)
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --target-processes all python /tmp/sanitizer_case.py > $out/sanitizer_$tool.log 2>&1
  e=$?
  tail -4 $out/sanitizer_$tool.log | sed "s/^/[$tool] /"
  [ $e -ne 0 ] && rc=1
done
exit $rc
