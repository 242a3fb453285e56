"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gcre_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gcre_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(engine):
    from geneticscre_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/gcre_b200.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), "ctypes binding and header disagree"


def test_no_cpu_fallback(engine):
    """Without a CUDA device the engine must refuse to run, not fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from geneticscre_b200 import _lib

    with pytest.raises(_lib.GcreError):
        engine.JoinExec("method1", 10, 10, 4)


def test_merge_topk_rule(engine):
    S = engine.Score
    merged = engine.merge_topk([[S(1.0, 5, 1, 1, 1), S(2.0, 3, 2, 1, 1)], [S(float("-inf"), -1, -1, 0, 0), S(2.0, 0, 9, 1, 1)]], 2)
    assert [(s.score, s.src, s.trg) for s in merged] == [(2.0, 3, 2), (2.0, 0, 9)]
    merged = engine.merge_topk([[S(1.0, 5, 1, 1, 1)]], 3)
    assert [(s.score, s.src) for s in merged] == [(float("-inf"), -1), (1.0, 5)]


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (checked textually over the package and headers)."""
    bad = []
    for base in ("geneticscre_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "gcre_oracle" in txt or "pyoracle" in txt:
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_host_pack_matches_numpy_packing():
    """gcre_host_pack_i32 (host threads, no GPU): bit c of row r set iff data[r][c] != 0, LSB-first words (src/gcre_paths.h:65-67)."""
    import ctypes as C

    import numpy as np

    from geneticscre_b200 import _lib, synth

    lib = _lib.load()
    rng = np.random.default_rng(7)
    for rows, cols, threads in ((1, 1, 1), (3, 63, 2), (5, 64, 1), (7, 65, 3), (40, 200, 4), (33, 1000, 0), (2, 0, 1)):
        data = (rng.random((rows, cols)) < 0.2).astype(np.int32) * rng.integers(-3, 4, size=(rows, cols), dtype=np.int32)
        w = (cols + 63) // 64
        out = np.full((rows, max(w, 1)), 0xAAAAAAAAAAAAAAAA, dtype=np.uint64)
        _lib.check(lib.gcre_host_pack_i32(data.ctypes.data_as(C.POINTER(C.c_int32)), rows, cols, out.ctypes.data_as(C.POINTER(C.c_uint64)), threads))
        if cols:
            want = synth.pack_bits((data != 0).astype(np.int32))
            assert np.array_equal(out[:, :w], want[:, :w]), (rows, cols)


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """The ctypes structures in geneticscre_b200/_lib.py must have the size and field offsets of the structs include/gcre_b200.h
    declares (a probe compiled with gcc prints them): a field added on one side only would shift every later field silently."""
    import ctypes as C
    import os
    import shutil
    import subprocess

    from geneticscre_b200 import _lib

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    pairs = {"gcre_score": _lib.ScoreC, "gcre_uid_ref": _lib.UidRefC, "gcre_exec_info": _lib.ExecInfoC, "gcre_join_opts": _lib.JoinOptsC,
             "gcre_decorated": _lib.DecoratedC}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "gcre_b200.h"', "int main(void) {"]
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, what, val = line.split()
        cls = pairs[cname]
        want = C.sizeof(cls) if what == "size" else getattr(cls, what).offset
        assert int(val) == want, f"{cname}.{what}: header {val}, ctypes {want}"
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in pairs.values())
