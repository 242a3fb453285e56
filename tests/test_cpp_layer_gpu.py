"""The header-only C++ class layer (include/gcre/*.h over the C ABI), compiled here with g++ -std=c++11 like the R package
would, run on the GPU: known-answer case, kept rows, row access, exception types.  Needs no reference checkout."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_class_layer_known_answers(tmp_path, engine):
    exe = str(tmp_path / "class_layer_kat")
    lib_dir = os.path.join(ROOT, "geneticscre_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-pthread", "-I", os.path.join(ROOT, "include", "gcre"),
                           os.path.join(ROOT, "tests", "cpp", "class_layer_kat.cpp"), "-L", lib_dir, "-lgcre_b200",
                           f"-Wl,-rpath,{lib_dir}", "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=dict(os.environ, GCRE_TIMER="1"))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "class layer: OK" in out.stdout
    assert "[gcre timer] level 2" in out.stdout
