"""Dev: wall time of BASELINE configs 1 and 2 (vignette-sized golden fixtures) on the engine vs the reference build."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers
from geneticscre_b200 import api
from oracle import pyoracle as po
for name in ("vignette_m1_len2_p100.npz", "vignette_m2_len3_p1000.npz"):
    w, method, L, k, want, _ = helpers.load_golden(name)
    for label, cls, ucls, kw in (("engine", api.JoinExec, api.UidRelSet, {}), ("reference(1 thread)", po.RefExec, po.UidRelSet, {"use_int_matrices": True})):
        ts = []
        for rep in range(4):
            t = time.perf_counter(); helpers.run_schedule(cls, ucls, w, method, L, k, **kw); ts.append((time.perf_counter() - t) * 1e3)
        pairs = sum(w.net.levels[l].n_pairs for l in want)
        print(name, label, "ms per full schedule:", [round(x, 2) for x in ts], "pairs", pairs, "perms", w.n_perms)
