#!/usr/bin/env python
"""profiles/sass_summary.txt: per kernel of libdismember_gpu.so, the count of the SASS mnemonics that prove which hardware
path it uses (tcgen05 MMA / TMEM loads / TMA bulk + tensor copies / FP pipes).  `python tools/sass_summary.py > profiles/sass_summary.txt`"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dismember_b200", "libdismember_gpu.so")
MN = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCATOMSWS", "SYNCS", "FFMA", "DFMA", "HFMA2", "MUFU", "LDG", "LDGSTS", "ATOM", "RED", "BAR"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts, name, total = collections.OrderedDict(), None, collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        total[name] += 1
        for k in MN:
            if op == k or op.startswith(k + ".") or (k == "UTMALDG" and op.startswith("UTMALDG")):
                counts[name][k] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} | mnemonic counts per kernel (sm_100a); UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld,")
print("# UTMALDG = cp.async.bulk.tensor (TMA tensor map, incl. tile::gather4), UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops")
print(f"{'kernel':70s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in MN))
for n, c in counts.items():
    print(f"{n[:70]:70s} {total[n]:6d} " + " ".join(f"{c[k]:7d}" for k in MN))
