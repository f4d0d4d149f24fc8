#!/usr/bin/env python
"""bench.py — Snekmer vectorize hot path on B200 (BASELINE.json config C2).

Workload (config.workload): synthetic UniRef-like proteins (log-normal lengths,
mean ~350, UniProt background + 0.1 % X), MIQS 10-letter alphabet, k = 3
(dense 1,000-k-mer basis), N = 1,000,000 sequences PER GPU (weak scaling: every
rank vectorises its own shard, no data-path collective).

One step = the whole vectorize rule body (kmerize.smk:67-129) over the batch:
pass 1 basis accumulation + finalisation (first-occurrence order), pass 2 dense
per-sequence counts written as int32 [N, K] to HBM.

  value      sequences/s, inputs already resident in HBM (CUDA events, max over ranks)
  e2e        same metric through snekmer_b200.pipeline.vectorize_host: pinned HOST
             residues/offsets in, HOST count matrix out, copies inside the timed region
  roofline   count_dense kernel: algorithmic bytes (R + 8(N+1) + 4NK) / its CUDA-event time
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the oracle port (numpy restatement of the reference) on the host cores,
             on a bounded sample of the same workload

`--impl reference` times only the CPU arm (rank 0), same JSON shape.
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
sys.path.insert(0, ROOT)

BG = dict(A=.122, L=.105, G=.084, R=.074, V=.071, D=.060, E=.057, P=.053, T=.050, S=.047, I=.047, F=.034, Q=.034,
          K=.025, M=.024, N=.022, Y=.022, H=.021, W=.014, C=.009)


def synth_proteins(n, seed, mean_len=350.0, sigma=0.6, x_frac=0.001):
    """L = clip(round(lognormal(mu, 0.6)), 30, 5000), iid background residues + 0.1 % X."""
    rng = np.random.default_rng(seed)
    mu = np.log(mean_len) - sigma * sigma / 2
    lens = np.clip(np.rint(rng.lognormal(mu, sigma, size=n)), 30, 5000).astype(np.int64)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    letters = np.frombuffer(("".join(BG.keys()) + "X").encode(), dtype=np.uint8)
    p = np.array(list(BG.values()), dtype=np.float64)
    p = np.append(p / p.sum() * (1 - x_frac), x_frac)
    cdf = np.cumsum(p)
    total = int(offsets[-1])
    res = np.empty(total, dtype=np.uint8)
    step = 1 << 26
    for i in range(0, total, step):
        u = rng.random(min(step, total - i), dtype=np.float32)
        res[i:i + len(u)] = letters[np.minimum(np.searchsorted(cdf, u), len(letters) - 1)]
    return res, offsets


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def cpu_arm(res, offsets, alphabet, k, sample_seqs):
    from oracle import cpu_baseline

    n = min(sample_seqs, len(offsets) - 1)
    off = offsets[:n + 1]
    r = cpu_baseline.vectorize_sample(res[:off[-1]], off, alphabet, k)
    return {"value": r["nseq"] / r["seconds"], "unit": "sequences/s", "cores": r["cores"], "kind": "port",
            "sample": f"first {n} sequences of the workload, two-pass vectorize (basis + dense counts), "
                      f"numpy oracle port, one process per shard; {r['seconds']:.2f} s", "K": r["K"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nseq", type=int, default=1_000_000, help="sequences per GPU")
    ap.add_argument("--cpu-sample", type=int, default=0, help="sequences in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    alphabet, k = "miqs", 3
    workload = f"C2: synthetic {args.nseq} proteins/GPU (lognormal len, mean~350, 0.1% X), miqs k=3, dense int32 counts"
    config = {"workload": workload, "alphabet": alphabet, "k": k, "nseq_per_gpu": args.nseq, "basis": "first-occurrence, K=1000",
              "l2": "inputs (0.35 GB) and output (4 GB) larger than L2", "parallelism": f"sequence-sharded x{world}, no collective"}

    if args.impl == "reference":
        if rank != 0:
            return
        ncores = os.cpu_count() or 1
        sample = args.cpu_sample or 20000 * min(ncores, 64)
        res, offsets = synth_proteins(sample, 2)
        vals = []
        for i in range(args.warmup + args.steps):
            r = cpu_arm(res, offsets, alphabet, k, sample)
            if i >= args.warmup:
                vals.append(r)
        v = float(np.mean([x["value"] for x in vals]))
        cb = dict(vals[-1], value=v)
        print(json.dumps({"impl": "reference", "metric": "sequences/sec vectorize", "value": v, "unit": "sequences/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * sample / v, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
                          "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist

    from snekmer_b200 import engine as E
    from snekmer_b200 import pipeline as P

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    res_np, offsets = synth_proteins(args.nseq, 2 + 1000 * rank)
    nres = int(offsets[-1])
    h_res = torch.empty(nres, dtype=torch.uint8, pin_memory=True)
    h_res.numpy()[:] = res_np
    batch = E.SequenceBatch.from_packed(res_np, offsets, dev)
    tab = E.alphabet_tables(alphabet, dev)
    S = tab.nsym ** k
    K = S
    out = torch.empty((batch.n, K), dtype=torch.int32, device=dev)
    count, first = E.basis_tables(S, dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    kern_ms = []

    def step(timed):
        count.zero_(); first.fill_(-1)
        E.basis_accumulate(batch, alphabet, k, count, first, 0)
        basis = E.basis_finalize(alphabet, k, count, first, 0)
        assert basis.K == K
        if timed:
            ev[2].record()
        E.count_dense(batch, alphabet, k, basis, out=out)
        if timed:
            ev[3].record()
        return basis

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        basis = step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev[0].record()
    for _ in range(args.steps):
        step(True)
        ev[3].synchronize()
        kern_ms.append(ev[2].elapsed_time(ev[3]))
    ev[1].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[1])
    clocks = sampler.stop()
    launches_per_step = 2 + 1 + 2 + 7 + 1      # memsets(2) accumulate(1) keys/emit(2) cub radix sort(~7) count(1)

    # ---- end to end: host buffers in, host matrix out -------------------------------
    e2e = None
    if not args.no_e2e:
        h_out = torch.empty((batch.n, K), dtype=torch.int32, pin_memory=True)
        for _ in range(2):
            P.vectorize_host(h_res, offsets, alphabet, k, out=h_out, device=dev)
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        n_e2e = max(3, min(args.steps, 5))
        t0.record()
        for _ in range(n_e2e):
            P.vectorize_host(h_res, offsets, alphabet, k, out=h_out, device=dev)
        t1.record()
        barrier()
        e2e_ms = t0.elapsed_time(t1) / n_e2e
        assert int(h_out[:1000].sum()) == int(out[:1000].sum().item())
    # ---- reduce over ranks ------------------------------------------------------------
    t = torch.tensor([total_ms, e2e_ms if not args.no_e2e else 0.0, float(np.mean(kern_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kern_ms_avg = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_per_step = total_ms / args.steps
    value = world * args.nseq / (ms_per_step * 1e-3)
    peak, peak_kind = peaks()
    alg_bytes = nres + 8 * (batch.n + 1) + 4 * batch.n * K
    achieved = alg_bytes / (kern_ms_avg * 1e-3) / 1e9
    line = {
        "metric": "sequences/sec vectorize", "value": value, "unit": "sequences/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
        "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "hbm", "kernel": "count_dense_kernel", "achieved": achieved, "peak": peak,
                     "peak_source": f"MEASURED_PEAKS.json ({peak_kind})", "unit": "GB/s", "frac": achieved / peak,
                     "algorithmic_bytes": alg_bytes, "kernel_ms": kern_ms_avg, "traffic": None},
    }
    if not args.no_e2e:
        line["e2e"] = {"value": world * args.nseq / (e2e_ms * 1e-3), "unit": "sequences/s", "ms_per_step": e2e_ms,
                       "h2d_bytes_per_step": nres + 8 * (batch.n + 1), "d2h_bytes_per_step": 4 * batch.n * K + 8,
                       "api": "snekmer_b200.pipeline.vectorize_host (pinned host residues/offsets -> pinned host int32 counts)"}
    if not args.no_cpu and world == 1:
        ncores = os.cpu_count() or 1
        sample = args.cpu_sample or 20000 * min(ncores, 64)
        line["cpu_baseline"] = cpu_arm(res_np, offsets, alphabet, k, sample)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
