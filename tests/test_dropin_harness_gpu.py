"""Drop-in check: the reference's own replay harness (test/harness.cpp, UNMODIFIED) compiled against include/gcre/ and
linked to the CUDA engine must print the same level-4 result block as the same harness built on the reference's join.

Both binaries are produced by oracle/build_ref.sh in the container that has the reference checkout and travel to the GPU
box under oracle/_ref/ (git-ignored); the test is skipped when they are absent.
"""
import os
import re
import subprocess

import pytest

from geneticscre_b200 import dumpfile, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
OUR_BIN = os.path.join(ROOT, "oracle", "_ref", "b200_harness")


def result_block(out):
    m = re.search(r"results : (\d+) \|(.*)\n\s+perms :(.*)\n", out)
    assert m, out[-2000:]
    entries = re.findall(r"(-?[\d.]+|-inf)\[(-?\d+):(-?\d+)\]", m.group(2))
    return int(m.group(1)), entries, m.group(3).split()


@pytest.mark.parametrize("n_gpus", [1, 2])
@pytest.mark.parametrize("method", ["method1", "method2"])
def test_reference_harness_runs_on_the_engine(tmp_path, engine, method, n_gpus):
    """n_gpus = 2: the C++ class layer drives two GPUs from one process (GCRE_GPUS): replicated path sets, kept levels on
    both, the score-only level-4/5 joins sharded by upstream row and merged on the host."""
    if not (os.path.exists(REF_BIN) and os.path.exists(OUR_BIN)):
        pytest.skip("oracle/_ref harness binaries not built (needs the reference checkout)")
    import torch

    if torch.cuda.device_count() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs")
    w = synth.make_workload(70, 90, 90, 260, 30, seed=17, max_path_length=5, real_table=True, max_freq=0.1, zero_frac=0.2)
    dump = str(tmp_path / "dump.txt")
    dumpfile.write_dump(dump, w)
    args = ["-f", dump, "-p", "30", "-m", method, "-k", "8", "-t", "0"]
    ref = subprocess.run([REF_BIN] + args, capture_output=True, text=True, timeout=300)
    ours = subprocess.run([OUR_BIN] + args, capture_output=True, text=True, timeout=300, env=dict(os.environ, GCRE_GPUS=str(n_gpus)))
    assert ref.returncode == 0, ref.stderr[-2000:]
    assert ours.returncode == 0, ours.stdout[-2000:] + ours.stderr[-2000:]
    n_r, e_r, p_r = result_block(ref.stdout)
    n_o, e_o, p_o = result_block(ours.stdout)
    assert n_r == n_o and p_r == p_o, (p_r, p_o)
    assert sorted(s for s, _, _ in e_r) == sorted(s for s, _, _ in e_o)
    kth = min(float(s) for s, _, _ in e_o)
    assert sorted(e for e in e_r if float(e[0]) > kth) == sorted(e for e in e_o if float(e[0]) > kth)
    assert ours.stdout.rstrip().endswith("done")
