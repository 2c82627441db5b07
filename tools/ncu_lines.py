"""Developer aid: executed warp instructions and stall samples of an ncu capture per source
function / source line (needs -lineinfo and `ncu --import-source on`).

usage: python tools/ncu_lines.py <file.ncu-rep> [--lines N] [--src parpe_b200/csrc/bdf_core.cuh]

Reads `ncu -i <rep> --page source --csv --print-source cuda,sass`: every CUDA source line is
followed by the SASS instructions attributed to it.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    nlines = 40
    if "--lines" in sys.argv:
        nlines = int(sys.argv[sys.argv.index("--lines") + 1])
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file = None
    cur_line = None
    hdr = None
    per_line = collections.defaultdict(lambda: [0, 0, 0, 0.0])  # inst, samples, static, thread_inst
    src_text = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = os.path.basename(r[1])
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_inst = hdr.index("Instructions Executed")
            i_samp = hdr.index("# Samples")
            i_thr = hdr.index("Thread Instructions Executed")
            continue
        if hdr is None:
            continue
        if r[0] != "":
            cur_line = (cur_file, int(r[0]))
            src_text[cur_line] = r[1]
            continue
        if len(r) <= i_inst or r[2] in ("...", ""):
            continue
        try:
            inst = int(r[i_inst])
            samp = int(r[i_samp])
            thr = int(r[i_thr])
        except ValueError:
            continue
        e = per_line[cur_line]
        e[0] += inst
        e[1] += samp
        e[2] += 1
        e[3] += thr
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    funcs = {}
    for fn in ("bdf_core.cuh", "kernels.cuh"):
        lst = []
        for i, l in enumerate(open(os.path.join(root, "parpe_b200/csrc", fn)).read().split("\n"), 1):
            m = re.match(r"\s*(?:template <[^>]*>\s*)?(?:PB_HD(?:_NOINLINE)?|__device__|__global__)\s+(?:static\s+)?(?:__forceinline__\s+)?[\w:<>\s\*&]+?\s(\w+)\(", l)
            if m:
                lst.append((i, m.group(1)))
        funcs[fn] = lst

    def fn_of(f, ln):
        if f not in funcs:
            return f
        name = f
        for i, n in funcs[f]:
            if i <= ln:
                name = n
            else:
                break
        return name

    agg = collections.defaultdict(lambda: [0, 0, 0, 0.0])
    for (f, ln), e in per_line.items():
        a = agg[fn_of(f, ln)]
        for k in range(4):
            a[k] += e[k]
    tot_i = sum(a[0] for a in agg.values())
    tot_s = sum(a[1] for a in agg.values())
    print("total executed warp instructions %d, samples %d" % (tot_i, tot_s))
    print("%-28s %12s %6s %8s %6s %7s %6s" % ("function", "inst", "%", "samples", "%", "static", "thr/i"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print("%-28s %12d %5.1f%% %8d %5.1f%% %7d %6.1f" % (k, a[0], 100.0 * a[0] / max(tot_i, 1), a[1],
                                                            100.0 * a[1] / max(tot_s, 1), a[2], a[3] / max(a[0], 1)))
    print()
    print("hottest lines by samples")
    for (f, ln), e in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:nlines]:
        print("%-14s %5d %11d %5.1f%% %7d %5.1f%% %4d  %s" % (f, ln, e[0], 100.0 * e[0] / max(tot_i, 1), e[1],
                                                             100.0 * e[1] / max(tot_s, 1), e[2], src_text[(f, ln)].strip()[:90]))


if __name__ == "__main__":
    main()
