"""Dev: where the wall time of a vignette-sized run (BASELINE config 1) goes, call by call."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import helpers
from geneticscre_b200 import api, schedule

for name in ("vignette_m1_len2_p100.npz", "vignette_m2_len3_p1000.npz"):
    w, method, L, k, want, _ = helpers.load_golden(name)
    for rep in range(4):
        T = []
        t0 = t = time.perf_counter()
        def tick(what):
            global t
            now = time.perf_counter(); T.append((what, (now - t) * 1e3)); t = now
        ex = api.JoinExec(method, w.n_cases, w.n_ctrls, w.n_perms); ex.top_k = k; tick("exec")
        ex.setValueTable(w.value_table); tick("table")
        ex.setPermutedMasks(w.perm_masks); tick("masks")
        seen = []
        res, kept = schedule.replay_levels(ex, api.UidRelSet, w, L, on_level=lambda n, r: seen.append((n, (time.perf_counter() - t0) * 1e3)))
        tick("schedule")
        del res, kept; ex.close(); tick("close")
        print(name, rep, "total %.2f ms:" % ((time.perf_counter() - t0) * 1e3), " ".join("%s=%.2f" % x for x in T), "| level done at", " ".join("%s@%.2f" % x for x in seen))
