#!/usr/bin/env python
"""Attribute an ncu source-page capture to CUDA source lines.

    ncu -i prof.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all libdismember_gpu.so ; nvdisasm -g -c capi.sm_100a.cubin > all.sass
    python tools/ncu_by_line.py sass.csv all.sass <mangled kernel name substring> [top N]

The ncu CSV lists the kernel's SASS in address order with stall samples and executed counts;
nvdisasm -g interleaves `//## File "...", line N` markers with the same instructions.  The two
are zipped instruction by instruction and summed per (file, line)."""
import collections
import csv
import re
import sys


def sass_lines(path, kernel):
    cur_line = ("?", 0)
    inside = False
    out = []
    for ln in open(path, errors="replace"):
        s = ln.strip()
        if s.startswith(".text."):
            inside = kernel in s
            continue
        if not inside:
            continue
        if s.startswith(".section") or s.startswith(".text"):
            inside = False
            continue
        m = re.match(r'//## File "(.*)", line (\d+)', s)
        if m:
            cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if s.startswith("/*") and "*/" in s:                    # /*0000*/  INSTR ;
            body = s.split("*/", 1)[1].strip()
            if body and not body.startswith("/*"):
                out.append((cur_line, body))
    return out


def main():
    csv_path, sass_path, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 40
    rows = list(csv.reader(open(csv_path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    sass = sass_lines(sass_path, kernel)
    if len(sass) != len(data):
        print(f"warning: {len(data)} profiled instructions vs {len(sass)} disassembled", file=sys.stderr)
    n = min(len(sass), len(data))
    samp = collections.Counter()
    execd = collections.Counter()
    stalls = collections.defaultdict(collections.Counter)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for i in range(n):
        key = sass[i][0]
        samp[key] += int(data[i][ix["# Samples"]])
        execd[key] += int(data[i][ix["Instructions Executed"]])
        for c in stall_cols:
            v = int(data[i][ix[c]])
            if v:
                stalls[key][c[6:]] += v
    tot_s, tot_e = sum(samp.values()), sum(execd.values())
    print(f"total samples {tot_s}, warp instructions {tot_e}")
    print(f"{'file:line':28s} {'samples':>8s} {'%':>6s} {'executed':>11s} {'%':>6s}  top stalls")
    order = execd.most_common(top) if "--by-executed" in sys.argv else samp.most_common(top)
    for key, _ in order:
        s = samp[key]
        st = ", ".join(f"{k}={v}" for k, v in stalls[key].most_common(3))
        print(f"{key[0] + ':' + str(key[1]):28s} {s:8d} {100 * s / tot_s:6.2f} {execd[key]:11d} {100 * execd[key] / tot_e:6.2f}  {st}")


if __name__ == "__main__":
    main()
