#!/usr/bin/env python3
"""bench.py -- users/sec of full TDM beam-search retrieval (beam 200, topk 10, depth ceil(log2 N)).

Default workload = BASELINE.json configs[1]: TDM synthetic 1M items, dim 64, beam 200, batch 1024,
1xB200.  One "step" = one pass of the hot path over one batch of 1024 synthetic users.

    python bench.py --gpus 1 --steps 64 --warmup 8
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...         # CPU restatement of the reference on the host cores

value  : whole-job users/s with queries and result buffers resident in HBM (dmg_tdm_retrieve_dev)
e2e    : same metric through the host-buffer C-ABI call (dmg_tdm_retrieve): pinned H2D of the
         B x T item ids and D2H of the topk ids/logits inside the timed region
timed region: the K steps are repeated R times (R sized so that every region lasts --target-sec, default 0.6 s; R is in
         config.repeats) with query batches from a pool of 64 distinct ones; ms_per_step = region / (K x R).
in flight: --inflight N (default 8) host threads, started BEFORE the timed region, each with its own handle (the engine +
         dmg_clone handles sharing its tables, the GPU form of the reference's per-thread model clones), take the steps
         round-robin: the kernels of different batches fill each other's gaps.
roofline: dominant kernel = wave_score_kernel (level-synchronous tcgen05 scorer, one launch per tree level; --arith strict:
         beam_search_kernel); algorithmic bytes per user = rows_scored*E*4 + T*E*4 + topk*8, rows_scored =
         256 + 400*(L-8); measured on a separate serial pass (one batch in flight) with CUDA events around every launch
         of the kernel (dmg_set_profiling / dmg_kernel_time): achieved = step bytes / summed launch durations of the step.
         roofline.traffic = ncu dram bytes per launch for THIS catalogue from profiles/traffic.json, else null.
structured: the same step on the structured ("trained-like") table of SURVEY 8d, where the users' beams diverge; with the
         distinct candidate rows one batch touches (dmg_wave_probe).
catalogues: --catalogues 10M,100M (default) repeats the measurement on the larger catalogues of the north star after the
         headline (100 M: 68.7 GB table + as much for its bf16 hi|lo copy; checked against the strict kernel).
Other paths: tools/bench_paths.py; sharded tables: tools/shard_check.py.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--items", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--beam", type=int, default=200)
    ap.add_argument("--topk", type=int, default=10)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--seq-len", type=int, default=10)
    ap.add_argument("--cpu-sample-sec", type=float, default=12.0, help="target CPU work for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify-strict", type=int, default=0,
                    help="re-run the first N users of the last batch with the strict fp32 kernel and compare ids / logit bits "
                         "(size-independent parity check for catalogues whose table is too large to ship to the CPU oracle)")
    ap.add_argument("--tau", type=float, default=None, help="certification band as a fraction of the worst-case bound")
    ap.add_argument("--target-sec", type=float, default=0.6, help="length of every timed region: the K steps are repeated R times")
    ap.add_argument("--catalogues", default="10M,100M", help="extra catalogues measured after the headline ('' = none)")
    ap.add_argument("--no-structured", action="store_true", help="skip the second workload on the structured ('trained-like') table")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1 only: skip the table-sharded legs (BASELINE configs 4 and 5, sharded TDM)")
    ap.add_argument("--shard-items", type=int, default=10_000_000, help="catalogue of the sharded TDM / JTM legs")
    ap.add_argument("--shard-dr-items", type=int, default=100_000_000, help="catalogue of the sharded Deep Retrieval leg")
    ap.add_argument("--inflight", type=int, default=0,
                    help="host threads / handles driving the GPU, one batch each in flight (1 = strictly serial steps; 0 = min(8, cores / ranks))")
    ap.add_argument("--arith", default="fast", choices=["fast", "strict"],
                    help="scorer arithmetic: tensor-core with certified cuts (same ids/logits) or strict fp32 SIMT")
    a = ap.parse_args()
    if a.inflight <= 0:
        a.inflight = 8
    return a


def algorithmic_bytes_per_user(L, E, T, topk, beam):
    s = beam.bit_length() - 1
    first = 2 * (1 << s)                       # children of the full start level
    rows = first + 2 * beam * max(L - s - 1, 0) if L > s else 0
    return rows, rows * E * 4 + T * E * 4 + topk * 8


def workload_name(args, items):
    return (f"TDM synthetic {items} items, dim={args.dim}, beam={args.beam}, batch={args.batch}, "
            f"topk={args.topk}, T={args.seq_len}")


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc, self.first = device, [], None, 0

    def start(self):
        if os.environ.get("DMG_BENCH_NO_CLOCKS"):                 # diagnostic only: is the sampler itself perturbing the run?
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def wait_first(self, timeout):
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        self.first = max(0, len(self.rows) - 1)                   # keep the sample that straddles the start of the timed region

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows[self.first:]:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_queries(args, step, rank, world):
    from dismember_b200 import synth
    return synth.queries(args.batch, args.seq_len, args.items, seed=4 + step * world + rank)


def cpu_run(args, tree_file, params, rows, seqs, threads):
    """Time the oracle (CPU restatement) on `seqs`; returns (users/s, seconds, outputs)."""
    from oracle import oracle as orc
    tree = orc.Tree.from_treefile(tree_file)
    model = orc.TdmModel(params, rows, args.dim, args.seq_len)
    t0 = time.perf_counter()
    out = model.retrieve_batch(tree, seqs, args.beam, args.topk, n_threads=threads)
    dt = time.perf_counter() - t0
    return len(seqs) / dt, dt, out


def cpu_tuned_run(args, tree_file, params, rows, seqs, threads, faithful=None, budget_s=10.0):
    """SURVEY 8(d)'s second CPU form (oracle/oracle_tuned.c): re-associated arithmetic (node-side W1x.x precomputed, W1a.Watt collapsed,
    polynomial exp, AVX-512/AVX2 clones), so NOT bit-equal to the reference -- a reported baseline, never the checker.  `agreement` says
    how close it stays to the faithful port on the same users."""
    from oracle import oracle as orc
    tree = orc.Tree.from_treefile(tree_file)
    model = orc.TdmModel(params, rows, args.dim, args.seq_len)
    t0 = time.perf_counter()
    tuned = orc.TunedTdm(tree, model, n_threads=threads)
    prep = time.perf_counter() - t0
    tuned.retrieve_batch(seqs[: 2 * threads], args.beam, args.topk, n_threads=threads)
    reps, dt, out = 0, 0.0, None
    while dt < budget_s and reps < 64:
        t0 = time.perf_counter()
        out = tuned.retrieve_batch(seqs, args.beam, args.topk, n_threads=threads)
        dt += time.perf_counter() - t0
        reps += 1
    res = {"value": len(seqs) * reps / dt, "unit": "users/s", "cores": threads, "kind": "port-tuned",
           "sample": f"{len(seqs)} users x {reps} passes, {dt:.1f} s; node-side precomputation {prep:.1f} s (+{rows * args.dim * 4 / 2**20:.0f} MiB), "
                     f"not in the rate"}
    if faithful is not None:
        same = out[0] == faithful[0]
        res["agreement"] = {"users_with_identical_topk": float(same.all(1).mean()),
                            "max_abs_logit_diff_on_same_ids": float(np.abs(out[1][same] - faithful[1][same]).max()) if same.any() else None}
    return res


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port: the JVM reference cannot run here)
    on all host cores, same metric/config, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dismember_b200 import synth
    from oracle import oracle as orc
    orc.build()
    tf = synth.tdm_tree(args.items, seed=1)
    L = tf.max_level
    rows = (1 << (L + 1)) - 1
    rng = np.random.Generator(np.random.PCG64(2))
    n_par = rows * args.dim + 3 * args.dim * args.dim + 2 * args.dim + 1
    params = (rng.standard_normal(n_par, dtype=np.float32) * np.float32(0.05))
    params[rows * args.dim + 3 * args.dim * args.dim: rows * args.dim + 3 * args.dim * args.dim + args.dim] = 0  # b1
    params[-1] = 0
    threads = os.cpu_count() or 1
    tree = orc.Tree.from_treefile(tf)
    model = orc.TdmModel(params, rows, args.dim, args.seq_len)
    # size the per-step sample so that K+W steps stay within a couple of minutes
    probe = make_queries(args, 0, 0, 1)[: 2 * threads]
    t0 = time.perf_counter()
    model.retrieve_batch(tree, probe, args.beam, args.topk, n_threads=threads)
    per_user = (time.perf_counter() - t0) / len(probe)
    budget = 90.0 / max(args.steps + args.warmup, 1)
    sample = int(max(threads, min(args.batch, budget / per_user)))
    qs = [make_queries(args, i, 0, 1)[:sample] for i in range(args.warmup + args.steps)]
    for w in range(args.warmup):
        model.retrieve_batch(tree, qs[w], args.beam, args.topk, n_threads=threads)
    t0 = time.perf_counter()
    for s in range(args.steps):
        model.retrieve_batch(tree, qs[args.warmup + s], args.beam, args.topk, n_threads=threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    rows_u, bytes_u = algorithmic_bytes_per_user(L, args.dim, args.seq_len, args.topk, args.beam)
    line = {
        "impl": "reference", "metric": "users/sec beam-search retrieval (beam=200, topk=10, depth=ceil(log2 N))",
        "value": value, "unit": "users/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, args.items), "levels": L, "rows_scored_per_user": rows_u,
                   "cpu_arm_users_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "users/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} users/step x {args.steps} steps, oracle/ C restatement (the Scala+MKL "
                                   f"reference cannot run: no JVM in the image)"},
        "e2e": {"value": value, "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:        # the tuned CPU form next to it (extra key; `value` above stays the faithful port, the arm the driver's ratio uses)
        line["cpu_baseline_tuned"] = cpu_tuned_run(args, tf, params, rows, qs[-1], threads, budget_s=8.0)
    except Exception as exc:  # noqa: BLE001 -- a baseline extra must never take the reference arm down
        line["cpu_baseline_tuned"] = {"unavailable": repr(exc)}
    print(json.dumps(line), flush=True)


class Workers:
    """NF persistent host threads, one per engine handle.  They are started at set-up and wait on a barrier, so no thread
    creation falls into a timed region; run() releases them for the steps lo..hi-1 (round-robin) and returns when all are done."""

    def __init__(self, n):
        self.n, self.job, self.errs = n, None, []
        self.go, self.done = threading.Barrier(n + 1), threading.Barrier(n + 1)
        self.threads = [threading.Thread(target=self._loop, args=(k,), daemon=True) for k in range(n)]
        for t in self.threads:
            t.start()

    def _loop(self, k):
        while True:
            self.go.wait()
            job = self.job
            if job is None:
                return
            step, lo, hi = job
            try:
                for i in range(lo + k, hi, self.n):
                    step(i, k)
            except Exception as ex:                                # noqa: BLE001
                self.errs.append(ex)
            self.done.wait()

    def run(self, step, lo, hi, before_release=None):
        self.job = (step, lo, hi)
        if before_release:
            before_release()
        self.go.wait()
        self.done.wait()
        if self.errs:
            raise self.errs[0]

    def stop(self):
        self.job = None
        self.go.wait()


def measure(args, env, items, structured=False, do_cpu=False, target_s=None, verify_strict=0, do_e2e=True, do_roofline=True):
    """One catalogue: build index + table, then value (device-resident), roofline (serial pass, dominant kernel timed alone),
    e2e (host buffers).  Returns a dict of results."""
    import torch
    from dismember_b200 import Engine, synth
    dev, local, world, rank, barrier, max_over_ranks = env["dev"], env["local"], env["world"], env["rank"], env["barrier"], env["max_over_ranks"]
    B, T, E, K, W = args.batch, args.seq_len, args.dim, args.steps, args.warmup
    target_s = args.target_sec if target_s is None else target_s
    tf = synth.tdm_tree(items, seed=1)
    L = tf.max_level
    rows = (1 << (L + 1)) - 1
    eng = Engine(local)
    eng.load_tree_tdm(L, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    params = None
    if structured:
        params = synth.din_params(rows, E, seed=2, structured=True)
        eng.load_din_weights(params, rows, E, T)
    else:
        eng.init_din_weights(np.float32, rows, E, T, seed=2)      # replicas: same table on every rank
    eng.set_arithmetic(args.arith)
    # waiting host threads spin while this rank's share of the cores covers them (lowest latency) and sleep otherwise (8 ranks x 8
    # threads on a 32-core host): dmg_set_sync_mode; the clones inherit it
    sync_mode = os.environ.get("DMG_BENCH_SYNC") or ("sleep" if env["world"] * args.inflight >= (os.cpu_count() or 1) else "spin")
    eng.set_sync_mode(sync_mode)
    if args.tau is not None:
        eng.set_fast_tolerance(args.tau)
    NF = max(1, args.inflight)
    engs = [eng] + [eng.clone() for _ in range(NF - 1)]           # one handle per host thread over ONE copy of the tables
    streams = [torch.cuda.Stream(dev) for _ in engs]              # non-default: the engines launch on them
    stream = streams[0]
    torch.cuda.set_stream(stream)
    for e, st in zip(engs, streams):
        e.set_stream(st.cuda_stream)
    P = 64                                                        # distinct query batches, cycled (every batch walks its own part of the table)
    host_q = [synth.queries(B, T, items, seed=4 + s * world + rank) for s in range(P)]
    dev_q = [torch.from_numpy(q).to(dev) for q in host_q]
    d_out = [(torch.empty((B, args.topk), dtype=torch.int32, device=dev), torch.empty((B, args.topk), dtype=torch.float32, device=dev),
              torch.empty((B,), dtype=torch.int32, device=dev)) for _ in engs]
    workers = Workers(NF)

    def step_dev(i, k=0, serial=False):
        o = d_out[k]
        fn = engs[k].tdm_retrieve_dev_sync if NF > 1 and not serial else engs[k].tdm_retrieve_dev
        fn(B, dev_q[i % P].data_ptr(), args.beam, args.topk, True, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr())

    e2e_last = [None] * NF

    def step_host(i, k=0):
        e2e_last[k] = engs[k].tdm_retrieve(host_q[i % P], args.beam, args.topk)

    def timed_dev(n_steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()

        def start():
            e0.record(stream)
            for st in streams[1:]:
                st.wait_event(e0)                                 # nothing of the timed steps starts before e0
        workers.run(step_dev, 0, n_steps, before_release=start)
        for st in streams[1:]:
            ek = torch.cuda.Event()
            ek.record(st)
            stream.wait_event(ek)                                 # e1 follows the last kernel of every stream
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---- warm-up, then size the timed region: the K steps are repeated R times (fresh batches from the pool) -----------
    workers.run(step_dev, 0, max(W, 3 * NF))                      # every handle past its first calls (scratch allocated, its step captured)
    barrier()
    probe_ms = timed_dev(max(K, 2 * NF))
    R = max(1, int(np.ceil(target_s * 1e3 / max(probe_ms * K / max(K, 2 * NF), 1e-3))))
    l0 = sum(e.launch_count for e in engs)
    dev_ms = timed_dev(K * R)
    launches = sum(e.launch_count for e in engs) - l0
    last_k = (K * R - 1) % NF
    last_items = d_out[last_k][0].cpu().numpy().copy()
    last_q = (K * R - 1) % P

    out = {"items": items, "levels": L, "rows": rows, "R": R, "NF": NF, "sync_mode": sync_mode, "value": world * B * K * R / (dev_ms * 1e-3),
           "ms_per_step": dev_ms / (K * R), "launches": launches / max(K * R, 1), "timed_region_s": dev_ms * 1e-3}

    # ---- roofline pass: ONE batch in flight on a handle of its own policy, the dominant kernel timed alone ---------------
    if do_roofline:
        os.environ["DMG_WAVE_SCORE_CTAS"] = "2"                   # a lone batch: two scorer CTAs per SM (the library's policy for a handle without clones)
        eng.set_profiling(True)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(3):
            step_dev(i, 0, serial=True)
        eng.kernel_time()
        barrier()
        s0.record(stream)
        for i in range(K):
            step_dev(i, 0, serial=True)
        s1.record(stream)
        barrier()
        out["serial_ms_per_step"] = max_over_ranks(s0.elapsed_time(s1)) / K
        kern_ms, kern_n = eng.kernel_time()
        eng.set_profiling(False)
        os.environ.pop("DMG_WAVE_SCORE_CTAS", None)
        out["kern_ms_per_step"], out["kern_launches_per_step"] = kern_ms / K, kern_n / K
    if args.arith == "fast":
        per = [e.fast_stats() for e in engs]
        out["fast_stats"] = {k: (max(p[k] for p in per) if k == "max_err_over_bound" else sum(p[k] for p in per)) for k in per[0]}

    # ---- e2e: host buffers through the C ABI (pinned H2D + D2H inside every call) ----------------------------------------
    if do_e2e:
        workers.run(step_host, 0, max(W, 3 * NF))                 # past every handle's first calls (its host-buffer step is captured once)
        barrier()
        t0 = time.perf_counter()
        workers.run(step_host, 0, K * R)
        torch.cuda.synchronize(dev)
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        out["e2e_value"], out["e2e_ms_per_step"], out["e2e_region_s"] = world * B * K * R / e2e_s, e2e_s / (K * R) * 1e3, e2e_s
        assert (e2e_last[last_k][0] == last_items).all(), "host-buffer and device-buffer paths disagree"

    if verify_strict > 0 and args.arith == "fast":
        nv = min(verify_strict, B)
        fi, fl, fc = eng.tdm_retrieve(host_q[last_q][:nv], args.beam, args.topk)
        eng.set_arithmetic("strict")
        si, sl, sc = eng.tdm_retrieve(host_q[last_q][:nv], args.beam, args.topk)
        eng.set_arithmetic("fast")
        out["strict_check"] = {"users_checked": nv, "ids_identical": bool((fi == si).all() and (fc == sc).all()),
                               "logits_bit_identical": bool((fl.view(np.uint32) == sl.view(np.uint32)).all())}

    if structured and args.arith == "fast" and rank == 0:
        # how much of the node table one batch really touches: distinct candidate rows per level (dmg_wave_probe)
        s_lvl = args.beam.bit_length() - 1
        uniq, tot = 0, 0
        for lvl in range(s_lvl + 1, L + 1):
            codes, _, counts, _ = eng.wave_probe(host_q[0], args.beam, lvl)
            m = np.arange(codes.shape[1])[None, :] < counts[:, None]
            uniq += len(np.unique(codes[m]))
            tot += int(counts.sum())
        out["unique_rows_per_batch"], out["rows_per_batch"] = uniq, tot
        out["unique_row_bytes_per_batch"] = uniq * E * 4

    # ---- cpu_baseline (rank 0, N=1 only): oracle on a bounded sample + parity check ---------------------------------------
    if do_cpu and rank == 0 and world == 1:
        from oracle import oracle as orc
        orc.build()
        if params is None:
            params = eng.download_din_weights()
        threads = os.cpu_count() or 1
        probe_v, _, _ = cpu_run(args, tf, params, rows, host_q[0][: 2 * threads], threads)
        n = int(max(2 * threads, min(B * P, args.cpu_sample_sec * probe_v)))
        nb = (n + B - 1) // B
        sample_q = np.concatenate([host_q[j] for j in range(nb)])[:n]
        v, dt, ref = cpu_run(args, tf, params, rows, sample_q, threads)
        gpu_i, gpu_l, _ = eng.tdm_retrieve(sample_q, args.beam, args.topk)
        out["cpu"] = {"value": v, "unit": "users/s", "cores": threads, "kind": "port",
                      "sample": f"{n} users from {nb} of the timed batches, {dt:.1f} s, oracle/ C restatement on {threads} "
                                f"host threads (the Scala+MKL reference cannot run here: no JVM)"}
        out["parity"] = {"users_checked": n, "ids_identical": bool((ref[0] == gpu_i).all()),
                         "logits_bit_identical": bool((ref[1].view(np.uint32) == gpu_l.view(np.uint32)).all())}
        out["cpu_tuned"] = cpu_tuned_run(args, tf, params, rows, sample_q, threads, ref)
    workers.stop()
    for e in reversed(engs):
        e.close()
    del dev_q, d_out
    torch.cuda.empty_cache()
    return out


def sharded_legs(args, env):
    """N > 1: the tables SHARDED over the ranks (csrc/shard.cu, csrc/dr.cu; NCCL inside libdismember_gpu.so) at the sizes BASELINE.json
    states -- configs[3] JTM item -> node weights on a 10 M item catalogue, configs[4] Deep Retrieval D=3 K=1000 on 100 M items -- plus
    TDM retrieval on the sharded 10 M table.  Every leg is checked bit for bit against an UNSHARDED engine on this rank's GPU (the
    tables are counter-based, so both hold the same values).  Returns one dict per leg (rank 0's view + whole-job rates)."""
    import torch
    import torch.distributed as dist
    from dismember_b200 import Engine, shard, synth
    world, rank, local, dev = env["world"], env["rank"], env["local"], env["dev"]
    T, E, B = args.seq_len, args.dim, args.batch
    out = {}
    ctl = dist.new_group(backend="gloo")                           # the 128-byte NCCL id and python objects travel here

    def sync():
        dist.barrier(group=ctl)
        torch.cuda.synchronize(dev)

    def timed(fn, reps):
        fn()                                                        # warm-up (buffers, NCCL channels)
        sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        sync()
        return env["max_over_ranks"](time.perf_counter() - t0) / reps

    eng = None
    try:
        # ---- TDM retrieval + JTM weights on the sharded node table ---------------------------------------------------------
        n_items = args.shard_items
        tf = synth.tdm_tree(n_items, seed=1)
        rows = (1 << (tf.max_level + 1)) - 1
        eng = shard.make_sharded_engine(local, group=ctl)
        eng.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
        eng.shard_init_din_weights(rows, E, T, seed=2)
        seqs = synth.queries(B, T, n_items, seed=100 + rank)
        res = {}

        def run_tdm():
            res["r"] = eng.shard_tdm_retrieve(seqs, args.beam, args.topk)
        dt = timed(run_tdm, 3)
        local_rows, global_rows, exchanged = eng.shard_info()
        full = Engine(local)
        full.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
        full.init_din_weights(np.float32, rows, E, T, seed=2)
        full.set_arithmetic("strict")
        n_chk = min(128, B)
        oi, ol, oc = full.tdm_retrieve(seqs[:n_chk], args.beam, args.topk)
        gi, gl, gc = res["r"]
        ok = bool((gi[:n_chk] == oi).all() and (gc[:n_chk] == oc).all() and (gl[:n_chk].view(np.uint32) == ol.view(np.uint32)).all())
        out["tdm_sharded"] = {"workload": f"TDM retrieval, node table of {n_items} items sharded by code range over {world} GPUs, "
                                          f"{B} users per rank, beam={args.beam}", "value": world * B / dt, "unit": "users/s",
                              "ms_per_step": dt * 1e3, "table_rows_global": global_rows, "table_rows_local": local_rows,
                              "rows_scored_for_other_ranks_per_step": exchanged // 4,
                              "parity_vs_unsharded_strict_engine": {"users_checked": n_chk, "ids_and_logit_bits_identical": ok}}
        # BASELINE configs[3]: JTM tree re-learning (item -> node weights of one level step) on the sharded table
        rng = np.random.Generator(np.random.PCG64(50 + rank))
        n_it, gap, old_level = 100_000, 4, 8
        n_samples = rng.integers(1, 9, n_it)
        off = np.zeros(n_it + 1, np.int64)
        off[1:] = np.cumsum(n_samples)
        sseq = synth.queries(int(off[-1]), T, n_items, seed=70 + rank)
        par = rng.integers((1 << old_level) - 1, (2 << old_level) - 1, n_it).astype(np.int32)

        def run_jtm():
            res["w"] = eng.shard_jtm_item_weights(off, sseq, par, old_level, old_level + gap, hierarchical=True, min_level=0)
        dtj = timed(run_jtm, 1)
        n_c = 4000
        want = full.jtm_item_weights(off[:n_c + 1], sseq[:off[n_c]], par[:n_c], old_level, old_level + gap, hierarchical=True, min_level=0)
        full.close()
        got = res["w"][:n_c]
        scorer_rows = int(off[-1]) * ((2 << gap) - 2)
        out["config4_jtm"] = {"workload": f"JTM item->node weights (TreeLearning.aggregateWeights, one level step, gap {gap}) on a {n_items} item "
                                          f"catalogue, node table sharded over {world} GPUs, {n_it} items ({int(off[-1])} samples) per rank",
                              "value": world * n_it / dtj, "unit": "items/s", "scorer_rows_per_s": world * scorer_rows / dtj, "seconds": dtj,
                              "parity_vs_unsharded_engine": {"items_checked": n_c, "weights_bit_identical":
                                                             bool((got.view(np.uint32) == want.view(np.uint32)).all())}}
        eng.close()
        eng = None
        torch.cuda.empty_cache()
        # ---- BASELINE configs[4]: Deep Retrieval D=3 K=1000, item tables sharded by item range --------------------------------
        n_dr, K, D, Ed, Bd, beam = args.shard_dr_items, 1000, 3, 32, 256, args.beam
        eng = shard.make_sharded_engine(local, group=ctl)
        eng.dr_init_synthetic(n_dr, K, D, T, Ed, J=2, seed=8)
        dseq = np.random.Generator(np.random.PCG64(90 + rank)).integers(-1, n_dr, (Bd, T)).astype(np.int32)

        def run_dr():
            res["d"] = eng.shard_dr_retrieve(dseq, beam, args.topk)
        dtd = timed(run_dr, 2)
        free, _ = torch.cuda.mem_get_info(dev)
        need = 3.0 * n_dr * Ed * 8 + 13.0 * K ** D + 4e9
        chk = {"skipped": f"the unsharded copy needs {need / 1e9:.0f} GB, {free / 1e9:.0f} GB free"}
        if need < free:
            full = Engine(local)
            full.dr_init_synthetic(n_dr, K, D, T, Ed, J=2, seed=8)
            ri, rs, rc = full.dr_retrieve(dseq, beam, args.topk)
            full.close()
            si, ss, sc = res["d"]
            chk = {"users_checked": Bd, "ids_identical": bool((si == ri).all() and (sc == rc).all()),
                   "scores_bit_identical": bool((ss.view(np.uint64) == rs.view(np.uint64)).all()), "results_per_user": float(rc.mean())}
        out["config5_deep_retrieval"] = {"workload": f"Deep Retrieval D={D} K={K}, {n_dr} items, E={Ed} (fp64), beam={beam}, item tables sharded by "
                                                     f"item range over {world} GPUs, {Bd} users per rank",
                                         "value": world * Bd / dtd, "unit": "users/s", "ms_per_step": dtd * 1e3,
                                         "item_table_gb_global": 3.0 * n_dr * Ed * 8 / 1e9, "parity_vs_unsharded_engine": chk}
    except Exception as ex:                                        # noqa: BLE001
        out["error"] = f"{type(ex).__name__}: {str(ex)[:300]}"
    finally:
        if eng is not None:
            try:
                eng.close()
            except Exception:                                      # noqa: BLE001
                pass
    out["data_plane"] = ("NCCL inside libdismember_gpu.so (dlopen libnccl.so.2): ncclSend/ncclRecv rounds per tree level, integer all-reduce of the "
                         "history tiles; torch.distributed only carries the NCCL id and the timing barrier")
    return out


def main():
    args = parse()
    # stdout carries the ONE JSON line and nothing else: libraries that write to fd 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(json_fd, "w")
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's log lines go to fd 1, which now is stderr (above): the init lines ("... rank r nranks N ... Init COMPLETE") of torch's
        # communicator and of the library's own (sharded legs) stay visible without touching the JSON line
        os.environ["NCCL_DEBUG"] = os.environ.get("DMG_NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, T, E, K, W = args.batch, args.seq_len, args.dim, args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    env = {"dev": dev, "local": local, "world": world, "rank": rank, "barrier": barrier, "max_over_ranks": max_over_ranks}
    import gc
    gc.collect()
    gc.disable()                                                  # no collector pause inside a timed region
    sys.setswitchinterval(1e-4)                                   # worker threads hand the GIL over within 0.1 ms
    sampler = ClockSampler(local)
    sampler.start()                                               # before the warm-up: nvidia-smi's start-up (NVML init on every GPU of
    sampler.wait_first(3.0)                                       # the box, once per rank) must not fall into a timed region
    sampler.mark()
    m = measure(args, env, args.items, do_cpu=not args.no_cpu_baseline, verify_strict=args.verify_strict)
    clocks = sampler.stop()
    gc.enable()

    L, rows = m["levels"], m["rows"]
    rows_u, bytes_u = algorithmic_bytes_per_user(L, E, T, args.topk, args.beam)
    peak, peak_src = hbm_peak()
    value, e2e_value = m["value"], m["e2e_value"]
    kern_ms = m.get("kern_ms_per_step", 0.0)
    achieved = B * bytes_u / (kern_ms * 1e-3) / 1e9 if kern_ms else None
    traffic = None
    fast_path = args.arith == "fast" and E in (16, 32, 64) and T <= 15   # E = 16 / 32 run the E = 64 tensor-core path on a zero-padded copy of the model
    kernel_name = "wave_score_kernel" if fast_path else "beam_search_kernel<float,%d>" % E
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(f"{kernel_name}@{args.items}")      # ncu dram bytes per launch for THIS catalogue, else null
        except Exception:
            traffic = None

    # ---- second workload: the structured ("trained-like") table, where the users' beams diverge -----------------------------
    structured = None
    if not args.no_structured and args.arith == "fast":
        ms = measure(args, env, args.items, structured=True, target_s=min(args.target_sec, 0.4), do_e2e=False, do_roofline=True,
                     verify_strict=min(256, B))
        structured = {"workload": workload_name(args, args.items) + ", structured table (rank-8 component: sibling scores differ by far more than rounding)",
                      "value": ms["value"], "unit": "users/s", "ms_per_step": ms["ms_per_step"], "R": ms["R"],
                      "frac_in_flight": ms["value"] / world * bytes_u / 1e9 / peak,
                      "kernel_frac": (B * bytes_u / (ms["kern_ms_per_step"] * 1e-3) / 1e9 / peak) if ms.get("kern_ms_per_step") else None,
                      "unique_rows_per_batch": ms.get("unique_rows_per_batch"), "rows_per_batch": ms.get("rows_per_batch"),
                      "unique_row_bytes_per_batch": ms.get("unique_row_bytes_per_batch"),
                      "fast_stats": ms.get("fast_stats"), "parity_fast_vs_strict_kernel": ms.get("strict_check")}

    # ---- the larger catalogues of the north star (same step, same timing; 100 M needs ~141 GB of HBM) -------------------------
    catalogues = {}
    for name in [c for c in args.catalogues.split(",") if c]:
        n_items = int(float(name.upper().replace("M", "e6").replace("K", "e3")))
        if n_items == args.items:
            continue
        try:
            import psutil
            if n_items >= 50_000_000 and psutil.virtual_memory().available / max(world, 1) < 20e9:
                catalogues[name] = {"skipped": "not enough host memory per rank to build the tree"}
                continue
            Lc = int(np.ceil(np.log2(n_items)))
            need = ((1 << (Lc + 1)) - 1) * E * 4 * 2 + 4e9
            free, _ = torch.cuda.mem_get_info(dev)
            if need > free:
                catalogues[name] = {"skipped": f"needs {need / 1e9:.0f} GB of HBM (table + bf16 hi|lo copy), {free / 1e9:.0f} GB free"}
                continue
            mc = measure(args, env, n_items, target_s=min(args.target_sec, 0.4), do_roofline=True,
                         verify_strict=min(1024, B) if n_items >= 50_000_000 else 0, do_cpu=False)
            _, bu = algorithmic_bytes_per_user(mc["levels"], E, T, args.topk, args.beam)
            catalogues[name] = {"items": n_items, "levels": mc["levels"], "value": mc["value"], "e2e": mc["e2e_value"], "unit": "users/s",
                                "ms_per_step": mc["ms_per_step"], "R": mc["R"], "node_table_gb": mc["rows"] * E * 4 / 1e9,
                                "frac_in_flight": mc["value"] / world * bu / 1e9 / peak,
                                "kernel_frac": (B * bu / (mc["kern_ms_per_step"] * 1e-3) / 1e9 / peak) if mc.get("kern_ms_per_step") else None,
                                "parity": mc.get("strict_check"), "fast_stats": mc.get("fast_stats")}
        except Exception as ex:                                    # noqa: BLE001
            catalogues[name] = {"error": str(ex)[:300]}

    sharded = None
    if world > 1 and not args.no_sharded:
        sharded = sharded_legs(args, env)

    if rank == 0:
        NF = m["NF"]
        line = {
            "metric": "users/sec beam-search retrieval (beam=200, topk=10, depth=ceil(log2 N))",
            "value": value, "unit": "users/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args, args.items), "levels": L, "rows_scored_per_user": rows_u,
                       "algorithmic_bytes_per_user": bytes_u, "node_table_gb": rows * E * 4 / 1e9,
                       "parallelism": f"replicas x{world}, users sharded, no collective",
                       "repeats": m["R"], "timed_region_s": m["timed_region_s"],
                       "timed_region": f"the {K} steps are repeated R={m['R']} times inside every timed region (device and e2e) with "
                                       f"query batches drawn from a pool of 64 distinct ones; ms_per_step = region / (steps x R)",
                       "batches_in_flight": NF, "host_wait": m.get("sync_mode"),
                       "in_flight": (f"{NF} host threads started before the timed region, one handle each (dmg_clone: one copy of the tables), "
                                     "take the steps round-robin; the roofline object is from a serial pass (one batch in flight)") if NF > 1 else "serial steps",
                       "serial_ms_per_step": m.get("serial_ms_per_step"),
                       "l2": f"inputs larger than L2: {rows * E * 4 / 1e9:.2f} GB node table + its bf16 hi|lo copy, 64 distinct query batches, no flush",
                       "arithmetic": ("level-synchronous tcgen05 bf16x3 tensor-core scorer (TMA tile::gather4 from a bf16 hi|lo copy of the "
                                      "table, 256 B per row like the fp32 rows) + certified cuts, strict fp32 re-score of near-cut candidates "
                                      "and of the topk (ids and logits bit-identical to the CPU oracle)")
                       if fast_path else "strict fp32 (sequential-k fma chains, bit-identical to the CPU oracle)",
                       "fast_stats": m.get("fast_stats")},
            "e2e": {"value": e2e_value, "unit": "users/s", "h2d_bytes_per_step": B * T * 4,
                    "d2h_bytes_per_step": B * args.topk * 8 + B * 4, "ms_per_step": m["e2e_ms_per_step"], "timed_region_s": m["e2e_region_s"]},
            "gpu_launches": int(round(m["launches"] * K)),
            "gpu_launches_per_step": m["launches"],
            "roofline_in_flight": {"achieved": value / world * bytes_u / 1e9, "unit": "GB/s", "frac": value / world * bytes_u / 1e9 / peak,
                                   "note": "whole step with the batches in flight (value x algorithmic bytes per user), not a kernel timed alone"},
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms_per_step": kern_ms, "kernel_launches_per_step": m.get("kern_launches_per_step"),
                         "kernel_ms_avg": kern_ms / max(m.get("kern_launches_per_step") or 1, 1),
                         "kernel_share_of_step": kern_ms / m["serial_ms_per_step"] if m.get("serial_ms_per_step") else None,
                         "measured_on": "serial pass (one batch in flight), CUDA events around every launch of the kernel (one per tree level); "
                                        "achieved = the step's algorithmic bytes / the summed launch durations of the step"},
            "cpu_baseline": m.get("cpu"),
            "cpu_baseline_tuned": m.get("cpu_tuned"),
            "parity": m.get("parity"),
            "parity_fast_vs_strict_kernel": m.get("strict_check"),
            "structured": structured,
            "catalogues": catalogues,
            "sharded": sharded,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
