#!/usr/bin/env python
"""Density sweep of the kernel choice: level-4 join at n = 10,000 patients with gene carrier frequencies up
to max_freq in {0.01, 0.05, 0.2, 0.5} (no all-zero genes), dense and sparse kernels forced and AUTO, both methods.

    python tools/density_sweep.py > profiles/r2_density_sweep.json

Prints one JSON object: per (max_freq, method) the mean carriers per level-4 operand row, the kernel time of the level-4 join
per kernel choice, and what AUTO picked.  AUTO must never be > 10 % slower than the better forced kernel."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from geneticscre_b200 import _lib, api, build, schedule, synth

    build.build()
    out = {"patients": 10000, "permutations": 1000, "genes": 2500, "edges": 14000, "rows": []}
    for max_freq in (0.01, 0.05, 0.2, 0.5):
        w = synth.make_workload(5000, 5000, out["genes"], out["edges"], out["permutations"], seed=777, max_path_length=4, real_table=True,
                                max_freq=max_freq, zero_frac=0.0)
        lv = w.net.levels["4"]
        for method in ("method1", "method2"):
            row = {"max_freq": max_freq, "method": method, "level4_pairs": lv.n_pairs}
            ref = None
            for name, kernel in (("dense", _lib.KERNEL_DENSE), ("sparse", _lib.KERNEL_SPARSE), ("auto", _lib.KERNEL_AUTO)):
                ex = api.JoinExec(method, w.n_cases, w.n_ctrls, w.n_perms)
                ex.kernel = kernel
                ex.top_k = 10
                ex.setValueTable(w.value_table)
                ex.setPermutedMasks(w.perm_masks)
                best = None
                for rep in range(3):
                    res, kept = schedule.replay_levels(ex, api.UidRelSet, w, 4, only=("4",))
                    r = res["4"]
                    best = r.info["kernel_ms"] if best is None else min(best, r.info["kernel_ms"])
                if name == "dense":
                    p3 = kept["paths3"].to_numpy()
                    p2 = kept["paths2"].to_numpy()
                    row["mean_carriers_paths3"] = float(np.unpackbits(p3.view(np.uint8), axis=1).sum(axis=1).mean())
                    row["mean_carriers_paths2"] = float(np.unpackbits(p2.view(np.uint8), axis=1).sum(axis=1).mean())
                key = ([(s.score, s.src, s.trg, s.cases, s.ctrls) for s in r.scores], r.permuted_scores.view(np.uint64).tolist())
                if ref is None:
                    ref = key
                row[name + "_ms"] = best
                row[name + "_equal_to_dense"] = key == ref
                if name == "auto":
                    row["auto_picked"] = {1: "dense", 2: "sparse"}.get(r.info["kernel"]) + ("+screen" if r.info.get("screened") else "")
                del res, kept
                ex.close()
            row["auto_vs_best"] = row["auto_ms"] / min(row["dense_ms"], row["sparse_ms"])
            out["rows"].append(row)
            sys.stderr.write(json.dumps(row) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
