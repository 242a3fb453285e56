"""Instruction / DRAM-byte counts of the dominant (last-level, score-only) join launches from an ncu metrics CSV, merged into
profiles/r2_counts.json - the static side of bench.py's roofline (the time side is measured live with CUDA events).

    ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active \
        --clock-control none -k regex:join_ -c 400 --csv --log-file X.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e [workload flags]
    python profiles/make_counts.py X.csv <patients> <edges> <permutations> <path_length> [label]

Per method the launch with the longest duration among the score-only (KEEP = 0) join kernels is the last-level join."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    src, patients, edges, perms, plen = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    label = sys.argv[6] if len(sys.argv) > 6 else ""
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ii, ki, mi, vi, ui = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        d = launches.setdefault(r[ii], {"name": r[ki]})
        v = float(r[vi].replace(",", ""))
        if r[mi] == "gpu__time_duration.sum":
            v = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        if r[mi].startswith("dram__bytes"):
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1)
        d[r[mi]] = v
    out = {}
    for method in (1, 2):
        best = None
        for d in launches.values():
            m = re.search(r"join_sparse(?:_sc)?_kernel<\(?(?:int\))?(\d), \(?(?:bool\))?(\d)", d["name"]) or re.search(r"join_dense_kernel<\(?(?:int\))?(\d), \(?(?:bool\))?(\d)", d["name"])
            if not m or int(m.group(1)) != method or int(m.group(2)) != 0:
                continue
            if best is None or d.get("gpu__time_duration.sum", 0) > best.get("gpu__time_duration.sum", 0):
                best = d
        if best:
            out[f"method{method}"] = {"kernel": best["name"].split("(gcre::")[0].replace("void gcre::", ""),
                                      "warp_instructions": int(best.get("smsp__inst_executed.sum", 0)),
                                      "dram_bytes": int(best.get("dram__bytes_read.sum", 0) + best.get("dram__bytes_write.sum", 0)),
                                      "ncu_ms": round(best.get("gpu__time_duration.sum", 0.0), 4),
                                      "issue_active_pct": round(best.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0), 2),
                                      "alu_pipe_pct": round(best.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 0.0), 2)}
    path = os.path.join(ROOT, "profiles", "r2_counts.json")
    db = json.load(open(path)) if os.path.exists(path) else {"_comment": "per last-level join launch: smsp__inst_executed.sum and dram__bytes_read.sum + dram__bytes_write.sum from ncu "
                                                                       "(profiles/make_counts.py); bench.py divides them by the kernel time it measures live", "workloads": []}
    cfg = {"patients": patients, "edges": edges, "permutations": perms, "path_length": plen}
    db["workloads"] = [x for x in db["workloads"] if x["config"] != cfg] + [{"config": cfg, "label": label, "source": os.path.basename(src), "launches": out}]
    json.dump(db, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
