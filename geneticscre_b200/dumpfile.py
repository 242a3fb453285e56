"""Writer for the text dump that the reference's replay harness reads (test/harness.cpp:36-107, parsers in
test/test.cpp:23-117; format summarised in SURVEY.md App. B).

One record per line, ``"<label> <payload>"``; the label is ignored by the reader.  Order: path_length, num_cases,
num_ctrls; uids + signs lines for levels 1a, 1b, 2, 3, 4, 5; four row-index lines; data1, data2, perms (rows separated by
spaces, columns by commas); table (doubles).  The reference ships no writer for this format.
"""
from __future__ import annotations

import numpy as np

from . import synth


def _uids_line(level) -> str:
    loc = level.location.astype(np.int64)
    loc[loc == 0xFFFFFFFF] = -1
    return " ".join(f"{s}:{t}:{c}:{l}" for s, t, c, l in zip(level.src, level.trg, level.count, loc))


def _rows(mat, fmt) -> str:
    return " ".join(",".join(fmt(v) for v in row) for row in mat)


def write_dump(path: str, w: "synth.Workload", path_length: int = 5) -> None:
    net = w.net
    if "5" not in net.levels:
        raise ValueError("the harness reads all six levels: build the network with max_path_length=5")
    n = w.n_patients
    perm_bits = np.unpackbits(w.perm_masks.view(np.uint8), axis=1, bitorder="little")[:, :n].astype(bool)
    is_case = np.zeros(n, dtype=bool)
    is_case[: w.n_cases] = True
    perms = (perm_bits == is_case[None, :]).astype(np.int32)  # CaseORControl: 1 = label kept
    with open(path, "w") as f:
        f.write(f"path_length {path_length}\n")
        f.write(f"num_cases {w.n_cases}\n")
        f.write(f"num_ctrls {w.n_ctrls}\n")
        for name in ("1a", "1b", "2", "3", "4", "5"):
            lv = net.levels[name]
            f.write(f"uids{name} {_uids_line(lv)}\n")
            f.write(f"signs{name} {' '.join(str(int(s)) for s in lv.signs)}\n")
        for name in ("1a", "1b", "2", "3"):
            f.write(f"data_idx{name} {' '.join(str(int(i)) for i in net.data_idx[name])}\n")
        f.write(f"data1 {_rows(synth.unpack_bits(w.gene_bits, n), str)}\n")
        f.write(f"data2 {_rows(synth.unpack_bits(w.gene_bits2, n), str)}\n")
        f.write(f"perms {_rows(perms, str)}\n")
        f.write(f"table {_rows(w.value_table, lambda v: repr(float(v)))}\n")
