"""SASS opcode histogram of the join kernels in the in-tree library (static instruction counts) -> profiles/r2_sass_opcodes.txt

    python profiles/sass_opcodes.py
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "geneticscre_b200", "libgcre_b200.so")
PICK = [
    ("method 1, score-only, running maxima", "join_sparse_kernelILi1ELb0EtLb0ELb0E"),
    ("method 2, score-only, thresholded", "join_sparse_kernelILi2ELb0EtLb0ELb1E"),
    ("method 1, KEEP, running maxima", "join_sparse_kernelILi1ELb1EtLb0ELb0E"),
    ("method 2, KEEP, thresholded", "join_sparse_kernelILi2ELb1EtLb0ELb1E"),
    ("method 2, score-only, running maxima", "join_sparse_kernelILi2ELb0EtLb0ELb0E"),
    ("dense method 1", "join_dense_kernelILi1ELb0E"),
    ("split-carrier method 1 NW=4", "join_sparse_sc_kernelILi1ELb0EtLi4ELb0E"),
    ("split-carrier method 2 NW=4", "join_sparse_sc_kernelILi2ELb0EtLi4ELb0E"),
    ("split-carrier method 1 NW=4, masks in shared memory (bulk async copy)", "join_sparse_sc_kernelILi1ELb0EtLi4ELb1E"),
    ("split-carrier method 2 NW=4, masks in shared memory (bulk async copy)", "join_sparse_sc_kernelILi2ELb0EtLi4ELb1E"),
    ("split-carrier method 2 NW=4, KEEP (emits the rows' counts)", "join_sparse_sc_kernelILi2ELb1EtLi4ELb0E"),
]
TENSOR = re.compile(r"^(UTMALDG|UTMASTG|UTC[A-Z]*MMA|LDTM|STTM|UTCBAR|HMMA|IMMA|QGMMA|UBLKCP|SYNCS)")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fns, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            fns[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            fns[cur][m.group(1)] += 1
    special = collections.Counter()
    for f, c in fns.items():
        for op, n in c.items():
            if TENSOR.match(op):
                special[op] += n
    out = ["# SASS opcode histogram of the join kernels in geneticscre_b200/libgcre_b200.so (cuobjdump -sass, sm_100a), static instruction counts.",
           "# No tensor-core or TMEM opcodes anywhere in the library: this path is bitwise integer counting + table gathers (north star).",
           "# The one sm_90+ mechanism in use is the 1-D bulk async copy (cp.async.bulk on an mbarrier: UBLKCP.S.G + SYNCS.*) that stages",
           "# the patient-major masks in shared memory in the PTS form of the split-carrier kernels.",
           "# The register / spill / shared-memory table of the same build is profiles/ptxas_sm100a.log.",
           f"# tensor / TMA / mbarrier opcodes found in the whole library: {dict(special) if special else 'none'}; kernels in the library: {len(fns)}", ""]
    for label, key in PICK:
        hit = [f for f in fns if key in f]
        if not hit:
            out.append(f"{label}: not in this build ({key})\n")
            continue
        c = fns[hit[0]]
        out.append(f"{label}  ({hit[0][:70]}...): {sum(c.values())} SASS instructions")
        out.append("  " + "  ".join(f"{op}:{n}" for op, n in c.most_common(22)) + "\n")
    open(os.path.join(ROOT, "profiles", "r2_sass_opcodes.txt"), "w").write("\n".join(out))
    print("\n".join(out[:8]))


if __name__ == "__main__":
    main()
