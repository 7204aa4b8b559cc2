"""Instruction counts per kernel of permon_b200/libpermon_b200.so (cuobjdump -sass): UBLKCP / UBLKPF = cp.async.bulk (TMA engine, 1-D bulk
copies / L2 prefetch), SYNCS = mbarrier, DFMA/DMUL/DADD = fp64 pipe.  usage: python profiles/r2_sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "permon_b200", "libpermon_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
cols = ["UBLKCP", "UBLKPF", "SYNCS", "DFMA", "DMUL", "DADD", "LDG", "STG", "LDS", "STS", "SHFL", "ATOMG", "RED", "MUFU", "BAR"]
print("SASS of permon_b200/libpermon_b200.so (cuobjdump -sass, sm_100a), instruction counts per kernel.")
print("UBLKCP = cp.async.bulk (TMA engine, 1-D bulk copies), UBLKPF = its L2-prefetch form, SYNCS = mbarrier operations, DFMA/DMUL/DADD = fp64 pipe,")
print("LDG/STG = global, LDS/STS = shared.  The direct stencil kernels (k_spmv_sd*) deliberately contain no UBLKCP: see DESIGN.md section 3.\n")
print(f"{'kernel':110s}" + "".join(f"{c:>8s}" for c in cols))
blocks = re.split(r"\n\s*Function : ", sass)[1:]
tot = collections.Counter()
for name, blk in zip(names, blocks):
    cnt = collections.Counter()
    for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", blk, re.M):
        op = m.group(1)
        for c in cols:
            if op == c or op.startswith(c):
                cnt[c] += 1
                break
    short = re.sub(r"\(.*", "", name.replace("pb::", "")).replace("void ", "")[:108]
    print(f"{short:110s}" + "".join(f"{cnt[c]:8d}" for c in cols))
    tot.update(cnt)
print(f"\n{'TOTAL':110s}" + "".join(f"{tot[c]:8d}" for c in cols))
