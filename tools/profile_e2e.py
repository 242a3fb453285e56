"""Host-side timing breakdown of the e2e step and of the resident step under different stream choices (dev tool)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from geneticscre_b200 import api, synth, _lib
import bench
class A: pass
a = A(); a.__dict__.update(bench.WORKLOAD); a.table = "auto"
w, _ = bench.make_workload(a)
lv = w.net.levels; n = w.n_patients
import os as _os; _os.environ.setdefault("GCRE_TRACE_OFF","1")
def T(label, t0):
    torch.cuda.synchronize(); print(f"  {label:28s} {(time.perf_counter()-t0)*1e3:9.2f} ms"); return time.perf_counter()
data1_i = torch.from_numpy(synth.unpack_bits(w.gene_bits, n)).pin_memory().numpy()
bits = np.unpackbits(w.perm_masks.view(np.uint8), axis=1, bitorder="little")[:, :n].astype(bool)
is_case = np.zeros(n, dtype=bool); is_case[: w.n_cases] = True
perm_i = torch.from_numpy((bits == is_case[None, :]).astype(np.int32)).pin_memory().numpy()
table = torch.from_numpy(np.ascontiguousarray(w.value_table)).pin_memory().numpy()
uid = {k: api.UidRelSet(lv[k].path_length, lv[k].src, lv[k].trg, lv[k].count, lv[k].location, lv[k].signs) for k in lv}
for rep in range(3):
    print("e2e rep", rep)
    t = time.perf_counter()
    ex = api.JoinExec("method1", w.n_cases, w.n_ctrls, w.n_perms); t = T("JoinExec()", t)
    ex.top_k = 10
    ex.setValueTable(table); t = T("setValueTable", t)
    ex.setPermutedCases(perm_i); t = T("setPermutedCases", t)
    d1 = ex.createPathSet(data1_i.shape[0]); d1.load(data1_i); t = T("load int matrix", t)
    zero = ex.createPathSet(0); p1 = ex.createPathSet(lv["1a"].n_pairs)
    ex.join(uid["1a"], ex.createPathSet(lv["1a"].n_uids), d1.select(w.net.data_idx["1a"]), p1); t = T("level 1a", t)
    p2 = ex.createPathSet(lv["2"].n_pairs); ex.join(uid["2"], p1, d1.select(w.net.data_idx["2"]), p2); t = T("level 2", t)
    p3 = ex.createPathSet(lv["3"].n_pairs); ex.join(uid["3"], p2, d1.select(w.net.data_idx["3"]), p3); t = T("level 3", t)
    ex.join(uid["4"], p3, p2, zero); t = T("level 4", t)
    del zero, p1, p2, p3, d1; ex.close(); t = T("teardown", t)
# resident step under stream choices
for label in ("own", "torch-default", "torch-side"):
    ex = api.JoinExec("method1", w.n_cases, w.n_ctrls, w.n_perms); ex.top_k = 10
    side = torch.cuda.Stream()
    if label == "torch-default": ex.set_stream(torch.cuda.current_stream().cuda_stream)
    if label == "torch-side": ex.set_stream(side.cuda_stream)
    ex.setValueTable(w.value_table); ex.setPermutedMasks(w.perm_masks)
    d1 = ex.createPathSet(w.gene_bits.shape[0]); d1.load_bits(w.gene_bits)
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        zero = ex.createPathSet(0); p1 = ex.createPathSet(lv["1a"].n_pairs)
        ex.join(uid["1a"], ex.createPathSet(lv["1a"].n_uids), d1.select(w.net.data_idx["1a"]), p1)
        p2 = ex.createPathSet(lv["2"].n_pairs); ex.join(uid["2"], p1, d1.select(w.net.data_idx["2"]), p2)
        p3 = ex.createPathSet(lv["3"].n_pairs); ex.join(uid["3"], p2, d1.select(w.net.data_idx["3"]), p3)
        ex.join(uid["4"], p3, p2, zero)
        del zero, p1, p2, p3
        torch.cuda.synchronize(); print(label, "resident step", rep, round((time.perf_counter() - t0) * 1e3, 2), "ms")
    del d1; ex.close()
