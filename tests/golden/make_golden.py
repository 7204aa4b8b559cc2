#!/usr/bin/env python
"""Extract the reference's golden numbers for the MPGP path into tests/golden/reference_golden.json.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Source files: /root/reference/src/tutorials/output/{ex1_*,ex2_1_*,ex3_*,jbearing2_{4,5,6}}.out, produced by the
reference's own test harness (test specs: src/tutorials/ex1.c:161-184, ex2.c:163-169, jbearing2.c:586-598).
Only numbers are extracted (counts, KKT residual magnitudes, per-iteration monitor values) -- no source code.
"""
import json
import os
import re

REF = "/root/reference/src/tutorials/output"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.json")


def parse_counts(text):
    d = {}
    m = re.search(r"CONVERGED due to (\w+), KSPReason=(-?\d+), required (\d+) iterations", text)
    d["reason_name"], d["reason"], d["its"] = m.group(1), int(m.group(2)), int(m.group(3))
    for key, pat in (("nmv", "Hessian multiplications"), ("ncg", "CG steps"), ("nexp", "expansion steps"),
                     ("nprop", "proportioning steps")):
        d[key] = int(re.search(r"number of %s (\d+)" % pat, text).group(1))
    kkt = re.findall(r"^r = (.*?)\s*= (\S+)\s+r[O]?/\|\|b\|\| = (\S+)", text, flags=re.M)
    d["kkt"] = [dict(name=n.strip(), r=float(r), rel=float(q)) for n, r, q in kkt]
    return d


def parse_trace(text):
    rows = []
    for m in re.finditer(r"^\s*(\d+) MPGP \[(.)\] \|\|gp\|\|=(\S+),\s+\|\|gf\|\|=(\S+),\s+\|\|gc\|\|=(\S+),\s+alpha=(\S+)", text, flags=re.M):
        rows.append(dict(it=int(m.group(1)), step=m.group(2), gp=float(m.group(3)), gf=float(m.group(4)),
                         gc=float(m.group(5)), alpha=float(m.group(6))))
    return rows


def main():
    g = {"_source": "permon/permon src/tutorials/output/*.out (reference test harness golden outputs)"}
    cases = {
        "ex1_1": dict(problem="ex1", n=100, args={}),
        "ex1_opt": dict(problem="ex1", n=100, args=dict(exptype="gf", explengthtype="opt")),
        "ex1_optapprox": dict(problem="ex1", n=100, args=dict(exptype="g", explengthtype="optapprox")),
        "ex1_bb": dict(problem="ex1", n=100, args=dict(exptype="gfgr", explengthtype="bb")),
        "ex1_projcg": dict(problem="ex1", n=100, args=dict(exptype="projcg")),
        "ex2_1_infinite-false": dict(problem="ex2", n=100, infinite=False, args={}),
        "ex2_1_infinite-true": dict(problem="ex2", n=100, infinite=True, args={}),
    }
    for name, spec in cases.items():
        text = open(os.path.join(REF, name + ".out")).read()
        g[name] = dict(spec, **parse_counts(text))
    for name, (mx, my) in {"jbearing2_4": (8, 12), "jbearing2_5": (10, 16), "jbearing2_6": (30, 30)}.items():
        text = open(os.path.join(REF, name + ".out")).read()
        d = parse_counts(text)
        d.update(problem="jbearing2", mx=mx, my=my, args=dict(rtol=1e-6, atol=1e-8), trace=parse_trace(text))
        m = re.search(r"Norm of difference of results from TAO and QP = (\S+) <= (\S+) = tolerance", text)
        d["tao_diff"], d["tao_diff_tol"] = float(m.group(1)), float(m.group(2))
        g[name] = d
    # ex3 (dualised problem, src/tutorials/ex3.c:166-182): MPGP on the dual QP, and SMALXE around it when an empty null space
    # matrix makes the dual carry a zero-row equality constraint
    text = open(os.path.join(REF, "ex3_1.out")).read()
    g["ex3_1"] = dict(problem="ex3dual", n=100, args={}, **parse_counts(text))
    text = open(os.path.join(REF, "ex3_nullspace.out")).read()
    d = parse_counts(text)               # counts of the inner MPGP; the first CONVERGED line is the outer SMALXE
    inner = re.findall(r"CONVERGED due to (\w+), KSPReason=(-?\d+), required (\d+) iterations", text)
    d.update(problem="ex3dual", n=100, outer_reason=int(inner[0][1]), outer_its=int(inner[0][2]), inner_reason=int(inner[1][1]),
             inner_its=int(inner[1][2]), total_inner=int(re.search(r"Total number of inner iterations (\d+)", text).group(1)))
    g["ex3_nullspace"] = d
    # the golden TEXT itself (test outputs of the reference's harness, not source code): the GPU tests diff the lines the library prints
    # for -qps_view_convergence / -qp_chain_view_kkt / the MPGP monitor against these files byte for byte
    import shutil
    outdir = os.path.join(os.path.dirname(OUT), "out")
    os.makedirs(outdir, exist_ok=True)
    for name in list(cases) + ["jbearing2_4", "jbearing2_5", "jbearing2_6", "ex3_1", "ex3_nullspace"]:
        shutil.copyfile(os.path.join(REF, name + ".out"), os.path.join(outdir, name + ".out"))
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print("wrote", OUT, {k: (v["its"], v["nmv"]) for k, v in g.items() if isinstance(v, dict)})


if __name__ == "__main__":
    main()
