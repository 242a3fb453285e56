"""Per-source-line instruction / stall-sample attribution for one kernel of an ncu report (dev + evidence tool).

    python profiles/hotlines.py <report.ncu-rep> <kernel-substring> [launch-index] [top-N]

ncu's CSV source page is SASS-only, so SASS rows are mapped to CUDA lines through `nvdisasm -g` of the in-tree library
(same build as the one profiled), matched by instruction order.
"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "geneticscre_b200", "libgcre_b200.so")], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    out, cur_fn, cur_line, active = collections.OrderedDict(), None, None, False
    for line in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            cur_fn = m.group(1)
            out[cur_fn] = []
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur_fn:
            out[cur_fn].append((int(m.group(1), 16), m.group(2).strip(), cur_line))
    return out


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{ksub}", "--launch-skip", str(launch), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    kname = rows[0][1]
    hdr = rows[1]
    ie, ws, so = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
    data = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":  # next launch section
            break
        if len(r) > ie:
            data.append(r)
    fns = sass_lines(ksub)
    # pick the disassembled function whose demangled-ish name matches template args of the profiled kernel
    # mangled template-argument string of the profiled instance: (int)N -> LiNE, (bool)B -> LbBE, unsigned short -> t, unsigned int -> j
    targs = re.search(r"<(.*?)>\(", kname)
    key = "I"
    for a in (targs.group(1).split(", ") if targs else []):
        mi, mb = re.match(r"\(int\)(\d+)", a), re.match(r"\(bool\)(\d)", a)
        key += f"Li{mi.group(1)}E" if mi else f"Lb{mb.group(1)}E" if mb else {"unsigned short": "t", "unsigned int": "j"}.get(a, "")
    key += "E"
    fn = [f for f in fns if ksub in f and key in f][0]
    ins = fns[fn]
    assert len(ins) == len(data), (len(ins), len(data))
    per_line = collections.defaultdict(lambda: [0, 0])
    tot_i = tot_s = 0
    for (off, text, loc), r in zip(ins, data):
        i, s = int(r[ie]), int(r[ws])
        per_line[loc][0] += i
        per_line[loc][1] += s
        tot_i += i
        tot_s += s
    src_cache = {}

    def src(loc):
        if not loc:
            return ""
        f, l = loc
        if f not in src_cache:
            p = os.path.join(ROOT, "geneticscre_b200", "csrc", f)
            src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
        return src_cache[f][l - 1].strip()[:100] if l - 1 < len(src_cache[f]) else ""

    print(f"# {kname}  launch {launch}: {tot_i} warp instructions, {tot_s} stall samples, {len(ins)} SASS instructions")
    print("#  instr%   stall%   file:line   source")
    for loc, (i, s) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:topn]:
        print(f"{i / tot_i:7.2%}  {s / max(tot_s, 1):7.2%}   {loc[0] if loc else '?'}:{loc[1] if loc else 0:<4d}  {src(loc)}")


if __name__ == "__main__":
    main()
