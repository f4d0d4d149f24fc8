#!/usr/bin/env python
"""bench.py — the Snekmer hot path on B200.

Default workload = BASELINE.json config C2 (the configuration the headline metric is quoted on):
synthetic UniRef-like proteins (log-normal lengths, mean ~350, UniProt background + 0.1 % X),
MIQS 10-letter alphabet, k = 3 (dense 1,000-k-mer basis), N = 1,000,000 sequences PER GPU (weak
scaling: every rank vectorises its own shard, no data-path collective).  One step = the whole
vectorize rule body (kmerize.smk:67-129) over the batch: pass 1 basis accumulation +
finalisation (first-occurrence order), pass 2 dense per-sequence counts as int32 [N, K] in HBM.

  value      sequences/s, inputs already resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the public API with pinned HOST buffers in and HOST results
             out, copies inside the timed region
  roofline   dominant kernel: algorithmic bytes (or int8 ops) / its CUDA-event time against
             MEASURED_PEAKS.json; `traffic` = DRAM bytes per launch from the committed
             `ncu --set full` capture (profiles/traffic.json)
  cpu_baseline  the oracle port (numpy restatement of the reference) on the host cores, on a
             bounded sample of the same workload

--workload learn | apply | apply_sparse | sweep run the C3 / C4 / C5 shaped paths (sparse sort-based learn,
tcgen05 dense scoring, SpMM scoring, alphabet / k sweep) with the same JSON contract; they are not the driver's line.
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


def tensor_peak_int8():
    """Dense int8 tensor peak in TOP/s: 2 x the measured cuBLAS bf16 burst figure (the int8 pipe runs at twice the
    bf16 rate on sm_100a: 4.5 vs 2.25 P nominal); fallback 2 x 1590."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return 2.0 * float(json.load(f)["bf16_tflops"]), "2 x measured bf16 burst (MEASURED_PEAKS.json)"
    except Exception:
        return 2.0 * 1590.0, "2 x fallback bf16"


SYN6 = {"AGILMV": "A", "FWY": "F", "NQSTC": "N", "DE": "D", "KRH": "K", "P": "P"}       # 6-letter alphabet for C3/C4
SYN6_ORACLE = {"syn6": [(k, v) for k, v in SYN6.items()]}


def zipf_annotations(n, n_ann, frac_unannotated, seed):
    """annotation id per sequence ~ Zipf(1.1) over n_ann ids, -1 for the unannotated share."""
    rng = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, n_ann + 1) ** 1.1
    ids = rng.choice(n_ann, size=n, p=w / w.sum()).astype(np.int32)
    ids[rng.random(n) < frac_unannotated] = -1
    return ids


def cpu_arm(res, offsets, alphabet, k, sample_seqs):
    from oracle import cpu_baseline

    n = min(sample_seqs, len(offsets) - 1)
    off = offsets[:n + 1]
    r = cpu_baseline.vectorize_sample(res[:off[-1]], off, alphabet, k)
    return {"value": r["nseq"] / r["seconds"], "unit": "sequences/s", "cores": r["cores"], "kind": "port",
            "sample": f"first {n} sequences of the workload, two-pass vectorize (basis + dense counts), "
                      f"numpy oracle port, one process per shard; {r['seconds']:.2f} s", "K": r["K"]}


# =====================================================================================================
# workloads: each returns (line, cpu_fn) where cpu_fn(sample) -> cpu_baseline dict
# =====================================================================================================
class Ctx:
    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def timed(self, step, steps, warmup):
        """warm-up, then EXACTLY `steps` steps between barrier+sync pairs; returns total ms (this rank) and clocks."""
        torch = self.torch
        sampler = ClockSampler(self.local_rank)
        sampler.start()                      # nvidia-smi needs ~100 ms per sample: it runs from the warm-up on
        for _ in range(max(warmup, 3)):
            step(False)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(True)
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        # a timed region of a few ms is over before the first sample: keep the same work running (untimed) until the
        # sampler has seen the clocks under this load
        t_end = time.time() + 2.0
        while len(sampler.lines) < 5 and time.time() < t_end:
            step(False)
            torch.cuda.synchronize()
        return ms, sampler.stop()

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def run_vectorize(ctx, args):
    torch = ctx.torch
    from snekmer_b200 import engine as E
    from snekmer_b200 import pipeline as P

    alphabet, k = "miqs", 3
    world, dev = ctx.world, ctx.dev
    workload = f"C2: synthetic {args.nseq} proteins/GPU (lognormal len, mean~350, 0.1% X), miqs k=3, dense int32 counts"
    config = {"workload": workload, "alphabet": alphabet, "k": k, "nseq_per_gpu": args.nseq, "basis": "first-occurrence, K=1000",
              "l2": "inputs (0.35 GB) and output (4 GB) larger than L2", "parallelism": f"sequence-sharded x{world}, no collective"}
    res_np, offsets = synth_proteins(args.nseq, 2 + 1000 * ctx.rank)
    nres = int(offsets[-1])
    h_res = torch.empty(nres, dtype=torch.uint8, pin_memory=True)
    h_res.numpy()[:] = res_np
    batch = E.SequenceBatch.from_packed(res_np, offsets, dev)
    tab = E.alphabet_tables(alphabet, dev)
    S = tab.nsym ** k
    K = S
    out = torch.empty((batch.n, K), dtype=torch.int32, device=dev)
    count, first = E.basis_tables(S, dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kern_ms = []

    def step(timed):
        count.zero_(); first.fill_(-1)
        E.basis_accumulate(batch, alphabet, k, count, first, 0)
        basis = E.basis_finalize(alphabet, k, count, first, 0)
        assert basis.K == K
        if timed:
            ev[0].record()
        E.count_dense(batch, alphabet, k, basis, out=out)
        if timed:
            ev[1].record()
            ev[1].synchronize()
            kern_ms.append(ev[0].elapsed_time(ev[1]))

    total_ms, clocks = ctx.timed(step, args.steps, args.warmup)
    launches_per_step = 2 + 1 + 2 + 7 + 1      # fills(2) basis_kernel(1) keys/emit(2) cub radix sort(~7) count_dense_kernel(1)
    e2e_ms = 0.0
    if not args.no_e2e:
        # transport dtype: uint16 is lossless for sequences shorter than 65,536 residues (checked by the library)
        h_out = torch.empty((batch.n, K), dtype=torch.uint16, pin_memory=True)
        for _ in range(2):
            P.vectorize_host(h_res, offsets, alphabet, k, out=h_out, dtype=torch.uint16, device=dev)
        ctx.barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        n_e2e = max(3, min(args.steps, 5))
        t0.record()
        for _ in range(n_e2e):
            P.vectorize_host(h_res, offsets, alphabet, k, out=h_out, dtype=torch.uint16, device=dev)
        t1.record()
        ctx.barrier()
        e2e_ms = t0.elapsed_time(t1) / n_e2e
        assert int(h_out[:1000].numpy().astype(np.int64).sum()) == int(out[:1000].sum().item())
    total_ms, e2e_ms, kern_ms_avg = ctx.max_over_ranks([total_ms, e2e_ms, float(np.mean(kern_ms))])
    ms_per_step = total_ms / args.steps
    peak, peak_kind = peaks()
    alg_bytes = nres + 8 * (batch.n + 1) + 4 * batch.n * K
    achieved = alg_bytes / (kern_ms_avg * 1e-3) / 1e9
    line = {
        "metric": "sequences/sec vectorize", "value": world * args.nseq / (ms_per_step * 1e-3), "unit": "sequences/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config, "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "hbm", "kernel": "count_dense_kernel", "achieved": achieved, "peak": peak,
                     "peak_source": f"MEASURED_PEAKS.json ({peak_kind})", "unit": "GB/s", "frac": achieved / peak,
                     "algorithmic_bytes": alg_bytes, "kernel_ms": kern_ms_avg, "traffic": TRAFFIC.get("count_dense_kernel")},
    }
    if not args.no_e2e:
        line["e2e"] = {"value": world * args.nseq / (e2e_ms * 1e-3), "unit": "sequences/s", "ms_per_step": e2e_ms,
                       "h2d_bytes_per_step": nres + 8 * (batch.n + 1), "d2h_bytes_per_step": 2 * batch.n * K + 8,
                       "api": "snekmer_b200.pipeline.vectorize_host (pinned host residues/offsets -> pinned host uint16 counts [N, K]; "
                              "lossless: the library refuses uint16 when a sequence has more than 65,535 residues)"}
    cpu = lambda sample: cpu_arm(res_np, offsets, alphabet, k, sample)
    return line, cpu


def _learn_inputs(ctx, args, n_ann):
    from snekmer_b200 import alphabet as A

    A.register_alphabet("syn6", SYN6)
    res_np, offsets = synth_proteins(args.nseq, 3 + 1000 * ctx.rank)
    ann = zipf_annotations(args.nseq, n_ann, 0.30, 30 + ctx.rank)
    return res_np, offsets, ann


def run_learn(ctx, args):
    """C3 shape: 6-letter alphabet, k = 8 (S = 1,679,616), 20k annotations (Zipf 1.1, 30 % unannotated).
    Step = Totals table (basis accumulate) + sort-based sparse learn + (N > 1) exchange of the COO lists."""
    torch = ctx.torch
    from snekmer_b200 import engine as E

    alphabet, k, n_ann = "syn6", 8, 20000
    res_np, offsets, ann = _learn_inputs(ctx, args, n_ann)
    nres = int(offsets[-1])
    batch = E.SequenceBatch.from_packed(res_np, offsets, ctx.dev)
    d_ann = torch.from_numpy(ann).to(ctx.dev)
    S = 6 ** 8
    state = {}

    def step(timed):
        keys, vals, count = E.learn_sparse_with_totals(batch, alphabet, k, d_ann, n_ann)     # matrix + Totals row over ALL sequences
        if ctx.world > 1:
            keys, vals, _ = E.exchange_coo_by_annotation(keys, vals, S, n_ann)
            ctx.dist.all_reduce(count)
        state["nnz"] = keys.numel()
        state["totals"] = count

    total_ms, clocks = ctx.timed(step, args.steps, args.warmup)
    # end to end: pinned host residues + annotation ids in, COO list out
    h_res = torch.empty(nres, dtype=torch.uint8, pin_memory=True); h_res.numpy()[:] = res_np
    h_ann = torch.from_numpy(ann).pin_memory()
    e2e_ms = 0.0
    if not args.no_e2e:
        cap = E.learn_sparse(batch, alphabet, k, d_ann, n_ann)[0].numel() + 1024      # local entries; pinned result buffers (a pageable .cpu() ran at ~2 GB/s)
        h_k = torch.empty(cap, dtype=torch.int64, pin_memory=True)
        h_v = torch.empty(cap, dtype=torch.int64, pin_memory=True)

        def e2e_once():
            b = E.SequenceBatch.from_packed(h_res.numpy(), offsets, ctx.dev, pinned=True)
            kk, vv = E.learn_sparse(b, alphabet, k, h_ann.to(ctx.dev, non_blocking=True), n_ann)
            m = kk.numel()
            h_k[:m].copy_(kk, non_blocking=True)
            h_v[:m].copy_(vv, non_blocking=True)
            ctx.torch.cuda.synchronize()
            return h_k[:m], h_v[:m]
        e2e_once()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            kk, vv = e2e_once()
        ctx.torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) / 2 * 1e3
    total_ms, e2e_ms = ctx.max_over_ranks([total_ms, e2e_ms])
    ms = total_ms / args.steps
    nnz = int(state["nnz"])
    alg_bytes = nres + 8 * (batch.n + 1) + 4 * batch.n + 16 * nnz + 8 * S
    peak, peak_kind = peaks()
    line = {"metric": "sequences/sec learn", "value": ctx.world * args.nseq / (ms * 1e-3), "unit": "sequences/s", "n_gpus": ctx.world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": f"C3 shape: {args.nseq} proteins/GPU, 6-letter alphabet k=8 (S=1,679,616), 20k annotations Zipf(1.1), 30% unannotated; sparse COO matrix",
                       "nnz_rank0": nnz, "l2": "inputs 0.44 GB, keys 3.5 GB larger than L2",
                       "parallelism": f"sequence-sharded x{ctx.world}" + (", all_to_all of COO runs by annotation range + local merge" if ctx.world > 1 else "")},
            "clocks": clocks, "gpu_launches": 110 * args.steps,      # ~8 annotation slices x (fill, keys, histogram, 4 sort passes, RLE x2, append x2) + gather, Totals, column sums (profiles/r2f_launches_learn.csv)
            "roofline": {"bound": "hbm", "kernel": "learn step (Totals table + gather by annotation + 32-bit keys + radix sort + run-length encode per annotation slice)", "achieved": alg_bytes / (ms * 1e-3) / 1e9,
                         "peak": peak, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})", "unit": "GB/s",
                         "frac": alg_bytes / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": alg_bytes, "kernel_ms": ms,
                         "traffic": None, "note": "whole step; the sorts move ~4 x 8 B per annotated residue, far above the algorithmic bytes"}}
    if not args.no_e2e:
        line["e2e"] = {"value": ctx.world * args.nseq / (e2e_ms * 1e-3), "unit": "sequences/s", "ms_per_step": e2e_ms,
                       "h2d_bytes_per_step": nres + 8 * (batch.n + 1) + 4 * batch.n, "d2h_bytes_per_step": 16 * nnz,
                       "api": "engine.SequenceBatch.from_packed(pinned host) + engine.learn_sparse -> pinned host COO (wall clock)"}

    def cpu(sample):
        from oracle import cpu_baseline
        n = min(sample, len(offsets) - 1)
        r = cpu_baseline.learn_sample(res_np[:offsets[n]], offsets[:n + 1], ann[:n], "syn6", k, extra=SYN6_ORACLE)
        return {"value": r["nseq"] / r["seconds"], "unit": "sequences/s", "cores": r["cores"], "kind": "port",
                "sample": f"first {n} sequences, numpy (annotation,k-mer) unique-count per shard + merge; {r['seconds']:.2f} s"}
    return line, cpu


def run_apply(ctx, args):
    """Dense-basis apply (C2 basis: miqs k=3, K = 1000) against a 50k-annotation matrix: counts + norms +
    tcgen05 int8 scoring GEMM with fused top-2.  Queries shard over GPUs, the matrix is replicated."""
    torch = ctx.torch
    from snekmer_b200 import engine as E

    alphabet, k, n_ann = "miqs", 3, args.n_ann
    S = K = 1000
    tr_res, tr_off = synth_proteins(200_000, 77)                      # training set (same on every rank)
    tr_ann = zipf_annotations(200_000, n_ann, 0.0, 78)
    tb = E.SequenceBatch.from_packed(tr_res, tr_off, ctx.dev)
    M, _ = E.learn_dense(tb, alphabet, k, None, torch.from_numpy(tr_ann), n_ann)
    M = M[:n_ann].contiguous()
    prep = E.prepare_annotations(M)
    assert prep is not None
    res_np, offsets = synth_proteins(args.nseq, 4 + 1000 * ctx.rank)
    nres = int(offsets[-1])
    batch = E.SequenceBatch.from_packed(res_np, offsets, ctx.dev)
    Q = torch.empty((batch.n, K), dtype=torch.int32, device=ctx.dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kern_ms, out = [], {}

    def step(timed):
        E.count_dense(batch, alphabet, k, None, out=Q)
        qn2 = E.row_norm2(Q)
        if timed:
            ev[0].record()
        r = E.apply_tc(Q, prep, qn2)
        if timed:
            ev[1].record(); ev[1].synchronize()
            kern_ms.append(ev[0].elapsed_time(ev[1]))
        assert r is not None
        out["r"] = r

    total_ms, clocks = ctx.timed(step, args.steps, args.warmup)
    h_res = torch.empty(nres, dtype=torch.uint8, pin_memory=True); h_res.numpy()[:] = res_np
    e2e_ms = 0.0
    if not args.no_e2e:
        def e2e_once():
            b = E.SequenceBatch.from_packed(h_res.numpy(), offsets, ctx.dev, pinned=True)
            q = E.count_dense(b, alphabet, k, None, out=Q)
            r = E.apply_tc(q, prep, E.row_norm2(q))
            return r.top1.cpu(), r.score1.cpu(), r.score2.cpu()
        e2e_once(); ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            e2e_once()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) / 3 * 1e3
    total_ms, e2e_ms, k_ms = ctx.max_over_ranks([total_ms, e2e_ms, float(np.mean(kern_ms))])
    ms = total_ms / args.steps
    ops = 2.0 * batch.n * prep.issued_macs_per_query()              # int8 MACs x 2 actually issued (per tile: width x planes x padded K)
    tiles = prep.tiles()
    peak, src = tensor_peak_int8()
    line = {"metric": "sequences/sec apply", "value": ctx.world * args.nseq / (ms * 1e-3), "unit": "sequences/s", "n_gpus": ctx.world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8 x int8 -> int32 (exact), float64 scaling", "data": "synthetic",
            "config": {"workload": f"dense apply: {args.nseq} queries/GPU (mean len 350) vs {n_ann} annotations, miqs k=3 (K=1000), "
                                   f"{len(tiles)} annotation tiles, {float(tiles[:, 2].mean()):.2f} base-256 digit planes per tile (max {prep.n_planes}); "
                                   f"counts + norms + tcgen05 GEMM + top-2",
                       "l2": "query counts 4 GB larger than L2", "parallelism": f"query-sharded x{ctx.world}, matrix replicated, no collective"},
            "clocks": clocks, "gpu_launches": 4 * args.steps,
            "roofline": {"bound": "tensor", "kernel": "apply_tc_kernel", "achieved": ops / (k_ms * 1e-3) / 1e12, "peak": peak,
                         "peak_source": src, "unit": "TOP/s", "frac": ops / (k_ms * 1e-3) / 1e12 / peak, "kernel_ms": k_ms,
                         "algorithmic_ops": ops, "traffic": TRAFFIC.get("apply_tc_kernel")}}
    if not args.no_e2e:
        line["e2e"] = {"value": ctx.world * args.nseq / (e2e_ms * 1e-3), "unit": "sequences/s", "ms_per_step": e2e_ms,
                       "h2d_bytes_per_step": nres + 8 * (batch.n + 1), "d2h_bytes_per_step": 20 * batch.n,
                       "api": "SequenceBatch.from_packed(pinned host) + count_dense + apply_tc -> host top-1/score/runner-up (wall clock)"}

    def cpu(sample):
        from oracle import cpu_baseline
        n = min(sample // 16, len(offsets) - 1)
        basis = np.arange(S, dtype=np.uint64)
        r = cpu_baseline.apply_dense_sample(res_np[:offsets[n]], offsets[:n + 1], alphabet, k, basis, M.cpu().numpy())
        return {"value": r["nseq"] / r["seconds"], "unit": "sequences/s", "cores": r["cores"], "kind": "port",
                "sample": f"first {n} queries, numpy counts + float64 BLAS cosine + argpartition top-2 per shard; {r['seconds']:.2f} s"}
    return line, cpu


def run_apply_sparse(ctx, args):
    """C4 shape: 6-letter alphabet k = 8 basis (S = 1,679,616), 50k annotations; queries as CSR over codes, exact
    integer SpMM + top-2 (one CTA per query, all annotations in shared-memory accumulators)."""
    torch = ctx.torch
    from snekmer_b200 import engine as E

    alphabet, k, n_ann, S = "syn6", 8, args.n_ann, 6 ** 8
    from snekmer_b200 import alphabet as A
    A.register_alphabet("syn6", SYN6)
    tr_res, tr_off = synth_proteins(args.ntrain, 79)
    tr_ann = zipf_annotations(args.ntrain, n_ann, 0.0, 80)
    tb = E.SequenceBatch.from_packed(tr_res, tr_off, ctx.dev)
    keys, vals = E.learn_sparse(tb, alphabet, k, torch.from_numpy(tr_ann), n_ann)
    tile = E.SPARSE_MAX_ANN
    cscs = [E.csc_build(keys, vals, S, min(tile, n_ann - a0), a0) for a0 in range(0, n_ann, tile)]
    res_np, offsets = synth_proteins(args.nseq, 5 + 1000 * ctx.rank)
    nres = int(offsets[-1])
    batch = E.SequenceBatch.from_packed(res_np, offsets, ctx.dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    out, kern_ms = {}, []

    def score(rowptr, cols, cvals):
        idxs, scs = [], []
        for c in cscs:
            r = E.apply_sparse(rowptr, cols, cvals, c, batch.max_len)
            if len(cscs) == 1:
                return r
            i = torch.stack([r.top1.to(torch.int64), r.top2.to(torch.int64)])
            idxs.append(torch.where(i >= 0, i + c.ann_lo, i)); scs.append(torch.stack([r.score1, r.score2]))
        return E.merge_top2(torch.stack(idxs), torch.stack(scs))

    def step(timed):
        rowptr, cols, cvals = E.count_csr(batch, alphabet, k, None)
        if timed:
            ev[0].record()
        out["r"] = score(rowptr, cols, cvals)
        if timed:
            ev[1].record(); ev[1].synchronize()
            kern_ms.append(ev[0].elapsed_time(ev[1]))
        out["csr"] = (rowptr, cols, cvals)

    total_ms, clocks = ctx.timed(step, args.steps, args.warmup)
    rowptr, cols, cvals = out["csr"]
    nnzq = int(cols.numel())
    # useful multiply-accumulates: every query entry meets every entry of its k-mer's column
    macs = 0
    for c in cscs:
        ci = cols.to(torch.int64)
        macs += int((c.colptr[ci + 1] - c.colptr[ci]).sum().item())
    h_res = torch.empty(nres, dtype=torch.uint8, pin_memory=True); h_res.numpy()[:] = res_np
    e2e_ms = 0.0
    if not args.no_e2e:
        def e2e_once():
            b = E.SequenceBatch.from_packed(h_res.numpy(), offsets, ctx.dev, pinned=True)
            r = score(*E.count_csr(b, alphabet, k, None))
            return r.top1.cpu(), r.score1.cpu(), r.score2.cpu()
        e2e_once(); ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            e2e_once()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) / 3 * 1e3
    total_ms, e2e_ms, k_ms = ctx.max_over_ranks([total_ms, e2e_ms, float(np.mean(kern_ms))])
    ms = total_ms / args.steps
    peak, peak_kind = peaks()
    entry_bytes = 4 if all(c.packed is not None for c in cscs) else 8      # packed CSC: annotation << 16 | value in one word
    alg_bytes = entry_bytes * macs + 8 * nnzq + 40 * batch.n          # gathered CSC entries, query entries, row pointers + results
    line = {"metric": "sequences/sec apply (sparse)", "value": ctx.world * args.nseq / (ms * 1e-3), "unit": "sequences/s", "n_gpus": ctx.world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "uint32 exact dots, float64 scaling", "data": "synthetic",
            "config": {"workload": f"C4 shape: {args.nseq} queries/GPU vs {n_ann} annotations learned from {args.ntrain} proteins, 6-letter k=8 "
                                   f"(S=1,679,616), nnz(M)={int(keys.numel())}, nnz(Q)={nnzq}; CSR counts + exact integer SpMM + top-2",
                       "l2": "CSC 0.54 GB larger than L2", "parallelism": f"query-sharded x{ctx.world}, matrix replicated in {len(cscs)} annotation slice(s)"},
            "clocks": clocks, "gpu_launches": (6 + len(cscs) + (len(cscs) > 1)) * args.steps,
            "roofline": {"bound": "hbm", "kernel": "apply_sparse_kernel", "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": peak,
                         "peak_source": f"MEASURED_PEAKS.json ({peak_kind})", "unit": "GB/s", "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes": alg_bytes, "kernel_ms": k_ms, "useful_macs": macs, "gmacs_per_s": macs / (k_ms * 1e-3) / 1e9,
                         "traffic": TRAFFIC.get("apply_sparse_kernel"),
                         "note": f"bytes = the CSC entries the gather formulation touches ({entry_bytes} B per multiply-accumulate, served by L2 and HBM); the kernel is latency / shared-atomic bound, see DESIGN.md"}}
    if not args.no_e2e:
        line["e2e"] = {"value": ctx.world * args.nseq / (e2e_ms * 1e-3), "unit": "sequences/s", "ms_per_step": e2e_ms,
                       "h2d_bytes_per_step": nres + 8 * (batch.n + 1), "d2h_bytes_per_step": 20 * batch.n,
                       "api": "SequenceBatch.from_packed(pinned host) + count_csr + apply_sparse -> host top-1/score/runner-up (wall clock)"}
    return line, None


SWEEP_POINTS = [("hydro", 8), ("hydro", 14), ("solvacc", 8), ("solvacc", 14), ("standard", 4), ("standard", 8), ("standard", 12),
                ("miqs", 3), ("miqs", 6), ("miqs", 10), ("miqs", 14), (None, 2), (None, 5), (None, 8), (None, 11), (None, 14)]


def run_sweep(ctx, args):
    """C5: alphabet / k sweep (2..20 letters, k = 2..14).  One step = every point once: basis + per-sequence
    counts through engine.vectorize (table kernels up to 2^27 codes, sort-based wide path beyond)."""
    torch = ctx.torch
    from snekmer_b200 import engine as E

    res_np, offsets = synth_proteins(args.nseq, 5 + 1000 * ctx.rank)
    nres = int(offsets[-1])
    batch = E.SequenceBatch.from_packed(res_np, offsets, ctx.dev)
    points = SWEEP_POINTS
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in points]
    per_ms = [[] for _ in points]
    info = [None] * len(points)

    def step(timed):
        for i, (a, k) in enumerate(points):
            if timed:
                ev[i][0].record()
            v = E.vectorize(batch, a, k)
            if timed:
                ev[i][1].record()
            info[i] = (v.path, v.K, int(v.vals.numel()) if v.vals is not None else None)
            del v
        if timed:
            torch.cuda.synchronize()
            for i in range(len(points)):
                per_ms[i].append(ev[i][0].elapsed_time(ev[i][1]))

    total_ms, clocks = ctx.timed(step, args.steps, args.warmup)
    (total_ms,) = ctx.max_over_ranks([total_ms])
    ms = total_ms / args.steps
    pts = []
    alg_total = 0
    for i, (a, k) in enumerate(points):
        tab = E.alphabet_tables(a, ctx.dev)
        path, K, nnz = info[i]
        m = float(np.mean(per_ms[i]))
        # B_vec (dense: R + 8(N+1) + 4NK; sparse: R + 16(N+1) + 12 nnz) + B_basis (R + 8(N+1) + 24 K)
        vec_b = nres + 8 * (batch.n + 1) + (4 * batch.n * K if path == "dense" else 8 * (batch.n + 1) + (12 if path == "csr" else 16) * nnz)
        alg = vec_b + nres + 8 * (batch.n + 1) + 24 * K
        alg_total += alg
        pts.append({"alphabet": str(a), "nsym": tab.nsym, "k": k, "log2_space": round(k * float(np.log2(tab.nsym)), 1), "path": path, "K": K,
                    "nnz": nnz, "ms": m, "seq_per_s": ctx.world * args.nseq / (m * 1e-3), "hbm_frac": alg / (m * 1e-3) / 1e9 / peaks()[0]})
    peak, peak_kind = peaks()
    line = {"metric": "sequences/sec vectorize (alphabet/k sweep)", "value": ctx.world * args.nseq * len(points) / (ms * 1e-3),
            "unit": "sequences/s", "n_gpus": ctx.world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "uint32/uint64 codes, int32 counts", "data": "synthetic",
            "config": {"workload": f"C5: {len(points)} (alphabet, k) points x {args.nseq} proteins/GPU; value = point-vectorisations of a sequence per second",
                       "points": pts, "l2": "inputs 0.35 GB x N/1e6, keys 8-16 B per residue: larger than L2",
                       "parallelism": f"sequence-sharded x{ctx.world}, no collective"},
            "clocks": clocks, "gpu_launches": sum(12 if p["path"] == "dense" else 25 for p in pts) * args.steps,
            "roofline": {"bound": "hbm", "kernel": "whole sweep step (table kernels + radix / segmented sorts)", "achieved": alg_total / (ms * 1e-3) / 1e9,
                         "peak": peak, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})", "unit": "GB/s",
                         "frac": alg_total / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": alg_total, "kernel_ms": ms, "traffic": None,
                         "note": "per-point fractions in config.points; the sort-based points move 10-20 x their algorithmic bytes"}}

    def cpu(sample):
        from oracle import cpu_baseline
        n = min(max(sample // 8, 1000), len(offsets) - 1)
        out = []
        for a, k in [("miqs", 6), (None, 8), (None, 14)]:
            r = cpu_baseline.vectorize_sparse_sample(res_np[:offsets[n]], offsets[:n + 1], a, k)
            out.append({"alphabet": str(a), "k": k, "seq_per_s": r["nseq"] / r["seconds"], "K": r["K"]})
        v = len(out) / sum(1.0 / o["seq_per_s"] for o in out)
        return {"value": v, "unit": "sequences/s", "cores": r["cores"], "kind": "port",
                "sample": f"first {n} sequences, 3 of the sweep points (harmonic mean), numpy oracle port, one process per shard", "points": out}
    return line, cpu


TRAFFIC = {}     # kernel -> dram bytes per launch from the committed `ncu --set full` capture (profiles/), else absent


def _load_traffic():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            TRAFFIC.update({k: v for k, v in json.load(f).items() if not k.startswith("_")})
    except Exception:
        pass


def reference_arm(args):
    """The CPU arm alone (rank 0): the oracle port of the reference's numpy path on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    alphabet, k = "miqs", 3
    workload = f"C2: synthetic {args.nseq} proteins/GPU (lognormal len, mean~350, 0.1% X), miqs k=3, dense int32 counts"
    config = {"workload": workload, "alphabet": alphabet, "k": k, "nseq_per_gpu": args.nseq, "basis": "first-occurrence, K=1000",
              "l2": "n/a (CPU)", "parallelism": f"{min(ncores, 64)} host processes, one per shard"}
    sample = args.cpu_sample or 20000 * min(ncores, 64)
    res, offsets = synth_proteins(sample, 2)
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_arm(res, offsets, alphabet, k, sample)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([x["value"] for x in vals]))
    cb = dict(vals[-1], value=v)
    _emit({"impl": "reference", "metric": "sequences/sec vectorize", "value": v, "unit": "sequences/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * sample / v, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
           "cpu_baseline": cb,
           "e2e": {"value": v, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


_REAL_STDOUT = None


def _capture_stdout():
    """Everything any library writes to fd 1 (NCCL's version banner, torchrun notices) goes to stderr; the ONE JSON line
    is written to the real stdout by _emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    _capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="vectorize", choices=["vectorize", "learn", "apply", "apply_sparse", "sweep"],
                    help="vectorize = the headline (BASELINE.json config C2); the others are the C3 / C4 shaped paths")
    ap.add_argument("--nseq", type=int, default=0, help="sequences per GPU (default per workload)")
    ap.add_argument("--n-ann", type=int, default=50000)
    ap.add_argument("--ntrain", type=int, default=400_000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="sequences in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if not args.nseq:
        args.nseq = {"vectorize": 1_000_000, "learn": 1_250_000, "apply": 1_000_000, "apply_sparse": 200_000, "sweep": 200_000}[args.workload]
    if args.impl == "reference":
        reference_arm(args)
        return
    _load_traffic()
    ctx = Ctx(args)
    line, cpu = {"vectorize": run_vectorize, "learn": run_learn, "apply": run_apply, "apply_sparse": run_apply_sparse, "sweep": run_sweep}[args.workload](ctx, args)
    if ctx.rank == 0:
        if not args.no_cpu and ctx.world == 1 and cpu is not None:
            ncores = os.cpu_count() or 1
            line["cpu_baseline"] = cpu(args.cpu_sample or 20000 * min(ncores, 64))
        _emit(line)
    ctx.finish()


if __name__ == "__main__":
    main()
