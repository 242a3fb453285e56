"""Host-side timing breakdown of one resident schedule step (dev tool)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from geneticscre_b200 import api, synth, _lib
import bench
class A: pass
a = A(); a.__dict__.update(bench.WORKLOAD); a.table = "auto"
if len(sys.argv) > 2: a.n_edges = int(sys.argv[2])
if len(sys.argv) > 3: a.n_perms = int(sys.argv[3])
w, _ = bench.make_workload(a)
lv = w.net.levels
kernel = {"dense": 1, "sparse": 2}[sys.argv[1] if len(sys.argv) > 1 else "sparse"]
for method in ("method1", "method2"):
    ex = api.JoinExec(method, w.n_cases, w.n_ctrls, w.n_perms); ex.kernel = kernel; ex.top_k = 10
    ex.setValueTable(w.value_table); ex.setPermutedMasks(w.perm_masks)
    d1 = ex.createPathSet(w.gene_bits.shape[0]); d1.load_bits(w.gene_bits)
    uid = {k: api.UidRelSet(lv[k].path_length, lv[k].src, lv[k].trg, lv[k].count, lv[k].location, lv[k].signs) for k in lv}
    for rep in range(3):
        T = {}
        def tick(name, t0):
            torch.cuda.synchronize(); T[name] = T.get(name, 0) + (time.perf_counter() - t0) * 1e3
        t = time.perf_counter(); zero = ex.createPathSet(0); p1 = ex.createPathSet(lv["1a"].n_pairs); z1 = ex.createPathSet(lv["1a"].n_uids); tick("create", t)
        t = time.perf_counter(); s1 = d1.select(w.net.data_idx["1a"]); tick("select", t)
        t = time.perf_counter(); r = ex.join(uid["1a"], z1, s1, p1); tick("join1a", t); k1 = r.info["kernel_ms"]
        t = time.perf_counter(); s2 = d1.select(w.net.data_idx["2"]); p2 = ex.createPathSet(lv["2"].n_pairs); tick("select+create2", t)
        t = time.perf_counter(); r = ex.join(uid["2"], p1, s2, p2); tick("join2", t); k2 = r.info["kernel_ms"]
        t = time.perf_counter(); s3 = d1.select(w.net.data_idx["3"]); p3 = ex.createPathSet(lv["3"].n_pairs); tick("select+create3", t)
        t = time.perf_counter(); r = ex.join(uid["3"], p2, s3, p3); tick("join3", t); k3 = r.info["kernel_ms"]
        t = time.perf_counter(); r = ex.join(uid["4"], p3, p2, zero); tick("join4", t); k4 = r.info["kernel_ms"]
        t = time.perf_counter(); del zero, p1, z1, s1, s2, p2, s3, p3; tick("free", t)
        print(method, rep, {k: round(v, 2) for k, v in T.items()}, "kernel_ms", [round(x, 2) for x in (k1, k2, k3, k4)], "launches", r.info["launches"])
