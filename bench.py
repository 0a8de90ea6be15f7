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
in flight: --inflight N (default 4) host threads, each with its own handle (the engine + dmg_clone handles sharing its
         tables, the GPU form of the reference's per-thread model clones), take the steps round-robin, so the tail of one
         batch's persistent kernel overlaps the head of the next; both value and e2e time EXACTLY K steps this way.
         The roofline object is measured on a separate serial pass (one batch in flight, kernel timed alone).
roofline: dominant kernel = beam_search_fast_kernel (tcgen05 scorer + certified cuts; --arith strict: beam_search_kernel);
         algorithmic bytes per user = rows_scored*E*4 + T*E*4 + topk*8, rows_scored = 256 + 400*(L-8);
         kernel time = CUDA events around its launches on the engine's stream (dmg_set_profiling / dmg_kernel_time)
Other catalogues: --items 10000000 / 100000000 (--verify-strict N checks N users against the strict kernel when the table is
too large to ship to the CPU oracle).  Other paths: tools/bench_paths.py; sharded tables: tools/shard_check.py.
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
    ap.add_argument("--inflight", type=int, default=4,
                    help="host threads / handles driving the GPU, one batch each in flight (1 = strictly serial steps)")
    ap.add_argument("--arith", default="fast", choices=["fast", "strict"],
                    help="scorer arithmetic: tensor-core with certified cuts (same ids/logits) or strict fp32 SIMT")
    return ap.parse_args()


def algorithmic_bytes_per_user(L, E, T, topk, beam):
    s = beam.bit_length() - 1
    first = 2 * (1 << s)                       # children of the full start level
    rows = first + 2 * beam * max(L - s - 1, 0) if L > s else 0
    return rows, rows * E * 4 + T * E * 4 + topk * 8


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
        "config": {"workload": f"TDM synthetic {args.items} items, dim={args.dim}, beam={args.beam}, batch={args.batch} "
                               f"(CPU arm: {sample} users per step)", "levels": L, "rows_scored_per_user": rows_u},
        "cpu_baseline": {"value": value, "unit": "users/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} users/step x {args.steps} steps, oracle/ C restatement (the Scala+MKL "
                                   f"reference cannot run: no JVM in the image)"},
        "e2e": {"value": value, "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
    from dismember_b200 import Engine, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = os.environ.get("DMG_NCCL_DEBUG", "WARN")   # NCCL's version banner goes to stdout: keep the JSON line alone
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, T, E, K, W = args.batch, args.seq_len, args.dim, args.steps, args.warmup

    # ---- setup (untimed): index, node table, queries -------------------------------------
    tf = synth.tdm_tree(args.items, seed=1)
    L = tf.max_level
    rows = (1 << (L + 1)) - 1
    eng = Engine(local)
    eng.load_tree_tdm(L, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    eng.init_din_weights(np.float32, rows, E, T, seed=2)          # replicas: same table on every rank
    eng.set_arithmetic(args.arith)
    if args.tau is not None:
        eng.set_fast_tolerance(args.tau)
    NF = max(1, min(args.inflight, K))
    engs = [eng] + [eng.clone() for _ in range(NF - 1)]           # one handle per host thread over ONE copy of the tables
    streams = [torch.cuda.Stream(dev) for _ in engs]              # non-default: the engines launch on them
    stream = streams[0]
    torch.cuda.set_stream(stream)
    for e, st in zip(engs, streams):
        e.set_stream(st.cuda_stream)
    host_q = [make_queries(args, s, rank, world) for s in range(W + K)]
    dev_q = [torch.from_numpy(q).to(dev) for q in host_q]
    d_out = [(torch.empty((B, args.topk), dtype=torch.int32, device=dev), torch.empty((B, args.topk), dtype=torch.float32, device=dev),
              torch.empty((B,), dtype=torch.int32, device=dev)) for _ in engs]
    last_k = (K - 1) % NF                                         # the handle that runs the last timed step

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

    def step_dev(i, k=0, serial=False):
        o = d_out[k]
        # several handles: the synchronous device-buffer call (its host thread has nothing else to do; the strict redo kernel is
        # launched only for the batches that need it); one handle: the asynchronous call, steps queued back to back
        fn = engs[k].tdm_retrieve_dev_sync if NF > 1 and not serial else engs[k].tdm_retrieve_dev
        fn(B, dev_q[i].data_ptr(), args.beam, args.topk, True, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr())

    e2e_last = [None] * NF

    def step_host(i, k=0):
        e2e_last[k] = engs[k].tdm_retrieve(host_q[i], args.beam, args.topk)

    def run_steps(step, lo, hi):
        """steps lo..hi-1, round-robin over the NF handles, one host thread per handle"""
        if NF == 1:
            for i in range(lo, hi):
                step(i, 0)
            return
        errs = []

        def work(k):
            try:
                for i in range(lo + k, hi, NF):
                    step(i, k)
            except Exception as ex:                                # noqa: BLE001
                errs.append(ex)
        th = [threading.Thread(target=work, args=(k,)) for k in range(NF)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]

    # ---- value: device-resident, CUDA events on the launching streams --------------------
    import gc
    gc.collect()
    gc.disable()                                                  # no collector pause inside a 45 ms timed region
    sys.setswitchinterval(1e-4)                                   # worker threads hand the GIL over within 0.1 ms
    sampler = ClockSampler(local)
    sampler.start()                                               # before the warm-up: nvidia-smi's start-up (NVML init on every GPU of
    sampler.wait_first(3.0)                                       # the box, once per rank) must not fall into the timed region
    run_steps(step_dev, 0, W)
    barrier()
    sampler.mark()                                                # clocks are reported from the samples taken after this point
    l0 = sum(e.launch_count for e in engs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for st in streams[1:]:
        st.wait_event(e0)                                         # nothing of the timed steps starts before e0
    run_steps(step_dev, W, W + K)
    for st in streams[1:]:
        ek = torch.cuda.Event()
        ek.record(st)
        stream.wait_event(ek)                                     # e1 follows the last kernel of every stream
    e1.record(stream)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = sum(e.launch_count for e in engs) - l0
    last_items = d_out[last_k][0].cpu().numpy().copy()
    last_logits = d_out[last_k][1].cpu().numpy().copy()

    # ---- roofline pass: the same K steps, ONE batch in flight, the kernel timed alone ------
    eng.set_profiling(True)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s0.record(stream)
    for i in range(W, W + K):
        step_dev(i, 0, serial=True)
    s1.record(stream)
    barrier()
    serial_ms = max_over_ranks(s0.elapsed_time(s1))
    kern_ms, kern_n = eng.kernel_time()
    eng.set_profiling(False)
    fast_stats = None
    if args.arith == "fast":
        per = [e.fast_stats() for e in engs]
        fast_stats = {k: (max(p[k] for p in per) if k == "max_err_over_bound" else sum(p[k] for p in per)) for k in per[0]}

    # ---- e2e: host buffers through the C ABI (H2D + D2H inside) ---------------------------
    run_steps(step_host, 0, W)
    barrier()
    t0 = time.perf_counter()
    run_steps(step_host, W, W + K)
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop()
    gc.enable()
    e2e_out = e2e_last[last_k]
    assert (e2e_out[0] == last_items).all(), "host-buffer and device-buffer paths disagree"

    value = world * B * K / (dev_ms * 1e-3)
    e2e_value = world * B * K / e2e_s
    rows_u, bytes_u = algorithmic_bytes_per_user(L, E, T, args.topk, args.beam)
    peak, peak_src = hbm_peak()
    kern_avg_ms = kern_ms / max(kern_n, 1)
    achieved = B * bytes_u / (kern_avg_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("beam_search_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None

    strict_check = None
    if args.verify_strict > 0 and args.arith == "fast":
        nv = min(args.verify_strict, B)
        fi, fl, fc = eng.tdm_retrieve(host_q[W + K - 1][:nv], args.beam, args.topk)
        eng.set_arithmetic("strict")
        si, sl, sc = eng.tdm_retrieve(host_q[W + K - 1][:nv], args.beam, args.topk)
        eng.set_arithmetic("fast")
        strict_check = {"users_checked": nv, "ids_identical": bool((fi == si).all() and (fc == sc).all()),
                        "logits_bit_identical": bool((fl.view(np.uint32) == sl.view(np.uint32)).all())}

    # ---- cpu_baseline (rank 0, N=1 only): oracle on a bounded sample + parity check --------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        orc.build()
        params = eng.download_din_weights()
        threads = os.cpu_count() or 1
        probe_v, probe_dt, _ = cpu_run(args, tf, params, rows, host_q[W + K - 1][: 2 * threads], threads)
        n = int(max(2 * threads, min(B * K, args.cpu_sample_sec * probe_v)))
        nb = (n + B - 1) // B                                      # whole timed batches, newest first
        sample_q = np.concatenate([host_q[W + K - 1 - j] for j in range(nb)])[:n]
        v, dt, out = cpu_run(args, tf, params, rows, sample_q, threads)
        gpu_i, gpu_l, _ = eng.tdm_retrieve(sample_q, args.beam, args.topk)
        cpu = {"value": v, "unit": "users/s", "cores": threads, "kind": "port",
               "sample": f"{n} users from the last {nb} timed batches, {dt:.1f} s, oracle/ C restatement on {threads} "
                         f"host threads (the Scala+MKL reference cannot run here: no JVM)"}
        parity = {"users_checked": n, "ids_identical": bool((out[0] == gpu_i).all()),
                  "logits_bit_identical": bool((out[1].view(np.uint32) == gpu_l.view(np.uint32)).all())}

    if rank == 0:
        line = {
            "metric": "users/sec beam-search retrieval (beam=200, topk=10, depth=ceil(log2 N))",
            "value": value, "unit": "users/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"TDM synthetic {args.items} items, dim={E}, beam={args.beam}, batch={B}, "
                                   f"topk={args.topk}, T={T}", "levels": L, "rows_scored_per_user": rows_u,
                       "algorithmic_bytes_per_user": bytes_u, "node_table_gb": rows * E * 4 / 1e9,
                       "parallelism": f"replicas x{world}, users sharded, no collective",
                       "batches_in_flight": NF,
                       "in_flight": (f"{NF} host threads, one handle each (dmg_clone: one copy of the tables), take the steps round-robin; "
                                     "the roofline object is from a serial pass over the same steps") if NF > 1 else "serial steps",
                       "serial_ms_per_step": serial_ms / K,
                       "l2": f"inputs larger than L2: {rows * E * 4 / 1e9:.2f} GB node table, fresh queries every step, no flush",
                       "arithmetic": ("tcgen05 bf16x3 tensor-core scorer + certified cuts, strict fp32 re-score of "
                                      "near-cut candidates and of the topk (ids and logits bit-identical to the CPU oracle)")
                       if args.arith == "fast" else "strict fp32 (sequential-k fma chains, bit-identical to the CPU oracle)",
                       "fast_stats": fast_stats},
            "e2e": {"value": e2e_value, "unit": "users/s", "h2d_bytes_per_step": B * T * 4,
                    "d2h_bytes_per_step": B * args.topk * 8 + B * 4, "ms_per_step": e2e_s / K * 1e3},
            "gpu_launches": int(launches),
            "roofline_in_flight": {"achieved": value / world * bytes_u / 1e9, "unit": "GB/s", "frac": value / world * bytes_u / 1e9 / peak,
                                   "note": "whole step with the batches in flight (value x algorithmic bytes per user), not a kernel timed alone"},
            "roofline": {"bound": "hbm", "kernel": ("beam_search_fast_kernel" if args.arith == "fast" else "beam_search_kernel<float,%d>" % E), "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms_avg": kern_avg_ms, "kernel_launches_timed": int(kern_n),
                         "kernel_share_of_step": kern_ms / serial_ms if world == 1 else None,
                         "measured_on": "serial pass (one batch in flight), CUDA events around each launch of the kernel"},
            "cpu_baseline": cpu,
            "parity": parity,
            "parity_fast_vs_strict_kernel": strict_check,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    for e in reversed(engs):
        e.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
