"""Collect the DRAM traffic of the fused kernels from `ncu --set full` raw-page CSV exports (profiles/r2_ncu_csv.sh) into
profiles/ncu_traffic.json, which bench.py quotes as `roofline.traffic` (never measured inside a bench run).
usage: python profiles/r2_ncu_traffic.py <workload> <csv> [<csv> ...]"""
import csv
import json
import os
import sys

FAM = {"k_spmv": "K_A spmv+dots+feas", "k_update_B": "K_B update+split", "k_direction_C": "K_C direction"}


def main():
    wl, files = sys.argv[1], sys.argv[2:]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
    db = json.load(open(out)) if os.path.exists(out) else {}
    ent = db.setdefault(wl, dict(source="", dram_bytes_per_launch={}, ncu_duration_us={}, kernels={}))
    for f in files:
        rows = list(csv.reader(open(f)))
        hdr, units = rows[0], rows[1]
        best = {}
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            name = d["Kernel Name"]
            fam = next((v for k, v in FAM.items() if k in name), None)
            if not fam:
                continue
            if "EpiA2T" in name:
                fam = "K_A' spmv+grad+split"
            elif "k_spmv" in name and "EpiAT" not in name:
                continue
            tscale = {"s": 1e6, "ms": 1e3, "us": 1.0, "ns": 1e-3}[units[hdr.index("gpu__time_duration.sum")]]
            dur = float(d["gpu__time_duration.sum"]) * tscale      # microseconds
            if dur < 20.0:          # early exits
                continue
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            rd = float(d["dram__bytes_read.sum"]) * scale[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(d["dram__bytes_write.sum"]) * scale[units[hdr.index("dram__bytes_write.sum")]]
            if fam not in best or dur < best[fam][0]:
                best[fam] = (dur, rd + wr, name)
        for fam, (dur, byts, name) in best.items():
            ent["dram_bytes_per_launch"][fam] = int(byts)
            ent["ncu_duration_us"][fam] = round(dur, 1)
            ent["kernels"][fam] = name[:120]
    ent["source"] = "ncu --set full --clock-control none, one launch each, exported with --page raw --csv: " + ", ".join("profiles/" + os.path.basename(f) for f in files)
    json.dump(db, open(out, "w"), indent=1)
    print(json.dumps(ent, indent=1))


if __name__ == "__main__":
    main()
