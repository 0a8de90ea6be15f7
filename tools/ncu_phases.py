#!/usr/bin/env python
"""Per-phase instruction / stall-sample totals of the fast beam kernel from an ncu source-page CSV.
usage: ncu_phases.py sass.csv all.sass beam_fast.cuh(as profiled) users_per_launch"""
import collections, csv, sys
sys.path.insert(0, __file__.rsplit("/", 1)[0])
import ncu_by_line as nb

csv_path, sass_path, src_path, users = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
rows = list(csv.reader(open(csv_path))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
sass = nb.sass_lines(sass_path, "beam_search_fast_kernel")
src = open(src_path).read().split("\n")
def find(s):
    for i, l in enumerate(src):
        if s in l: return i + 1
    raise KeyError(s)
marks = [("strict_rows", "template <int RPT>"), ("strict_batch", "__device__ __noinline__ void strict_score_batch"),
         ("select_fn", "__device__ __forceinline__ bool elect_one"), ("gather_fn", "template <int NIT>"),
         ("kernel_setup", "beam_search_fast_kernel(const BeamParams"), ("user_prologue", "// ---- next user (dynamic scheduler)"),
         ("cut", "// ---- which candidates stay"), ("cut_band", "if (n_keep != beam) {"), ("expand", "int nc;"), ("eps", "// ---- eps of this level"),
         ("gather", "// (A) gather rows"), ("mma_issue1", "// (B) S = X . K^T"), ("softmax", "// (C) Mask + SoftMax"),
         ("mma_issue2", "// (D) Hacc += P . H"), ("epilogue", "// (E) epilogue"), ("final", "// ---- K3: topk"), ("end", "// ---- per-level maxima")]
marks = [(n, find(s)) for n, s in marks]
def phase(f, l):
    if f != "beam_fast.cuh": return "inl:" + f
    p = "header"
    for n, ln in marks:
        if l >= ln: p = n
    return p
ex = collections.Counter(); sm = collections.Counter()
for i in range(min(len(sass), len(data))):
    (f, l), _ = sass[i]
    ph = phase(f, l)
    ex[ph] += int(data[i][ix["Instructions Executed"]]); sm[ph] += int(data[i][ix["# Samples"]])
te, ts = sum(ex.values()), sum(sm.values())
print(f"profiled {len(data)} vs disassembled {len(sass)} instructions; warp instr/user {te / users:.0f}")
for k, v in ex.most_common():
    print(f"{k:30s} instr/user {v / users:9.0f} {100 * v / te:5.1f}%   samples {100 * sm[k] / ts:5.1f}%")
