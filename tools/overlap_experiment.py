"""Dev experiment: the two methods' schedules run concurrently from two host threads (each exec on its own stream)."""
import sys, time, os, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from geneticscre_b200 import api, synth, _lib, schedule
import bench
class A: pass
a = A(); a.__dict__.update(bench.WORKLOAD); a.table = "auto"
w, _ = bench.make_workload(a)
lv = w.net.levels
st = {}
for method in ("method1", "method2"):
    ex = api.JoinExec(method, w.n_cases, w.n_ctrls, w.n_perms); ex.top_k = 10
    ex.setValueTable(w.value_table); ex.setPermutedMasks(w.perm_masks)
    d1 = ex.createPathSet(w.gene_bits.shape[0]); d1.load_bits(w.gene_bits)
    uid = {k: api.UidRelSet(lv[k].path_length, lv[k].src, lv[k].trg, lv[k].count, lv[k].location, lv[k].signs).make_resident(ex) for k in lv}
    st[method] = (ex, d1, uid)
def run(method):
    ex, d1, uid = st[method]
    zero = ex.createPathSet(0); p1 = ex.createPathSet(lv["1a"].n_pairs)
    ex.join(uid["1a"], ex.createPathSet(lv["1a"].n_uids), d1.select(w.net.data_idx["1a"]), p1)
    p2 = ex.createPathSet(lv["2"].n_pairs); ex.join(uid["2"], p1, d1.select(w.net.data_idx["2"]), p2)
    p3 = ex.createPathSet(lv["3"].n_pairs); ex.join(uid["3"], p2, d1.select(w.net.data_idx["3"]), p3)
    ex.join(uid["4"], p3, p2, zero)
for mode in ("serial", "threads", "serial", "threads"):
    ts = []
    for rep in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if mode == "serial":
            run("method1"); run("method2")
        else:
            th = [threading.Thread(target=run, args=(m,)) for m in ("method1", "method2")]
            [t.start() for t in th]; [t.join() for t in th]
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(mode, [round(t, 1) for t in ts])
