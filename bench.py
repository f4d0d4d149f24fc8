#!/usr/bin/env python
"""bench.py — the Snekmer hot path on B200.

The line's headline workload is BASELINE.json config C2 (the configuration the metric is quoted on for one GPU):
synthetic UniRef-like proteins (log-normal lengths, mean ~350, UniProt background + 0.1 % X), MIQS 10-letter alphabet,
k = 3 (dense 1,000-k-mer basis), N = 1,000,000 sequences PER GPU (weak scaling: every rank vectorises its own shard).
One step = the whole vectorize rule body (kmerize.smk:67-129) over the batch: basis in first-occurrence order
(order-only walk with device-side early exit: min_filter = 0 needs no occurrence counts) + dense per-sequence counts
as int32 [N, K] in HBM.

  value      sequences/s, inputs already resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the public API with pinned HOST buffers in and HOST results out, copies inside the
             timed region (lossless uint8 + escape-list transport; the other transports are listed beside it)
  roofline   dominant kernel: algorithmic bytes (or int8 ops) / its CUDA-event time against MEASURED_PEAKS.json;
             `traffic` = DRAM bytes per launch from the committed `ncu --set full` capture (profiles/traffic.json)
  cpu_baseline  the oracle port (numpy restatement of the reference) on the host cores, bounded sample

BASELINE.json's metric is "vectorize+apply at 1/2/4/8 GPUs", and configs C3 / C4 are the multi-GPU ones, so the SAME
line carries, under "workloads", one sub-record per further path — each with the same contract (value, ms_per_step,
roofline, e2e, gpu_launches) plus the collective it runs at N > 1 and an in-run parity check:

  apply         dense-basis scoring: counts + norms + tcgen05 int8 GEMM with fused top-2; queries sharded, matrix replicated
  learn         C3 shape (6-letter alphabet, k = 8, 20k annotations): sparse sort-based learn; at N > 1 the all_to_all of
                COO runs by balanced annotation range + merge tree + all_reduce(Totals) (learn.smk:467-494)
  apply_sparse  C4 shape: CSR counts + exact integer SpMM + top-2, query-sharded with the matrix replicated AND, at N > 1,
                annotation-sharded (all_gather of the query CSR, all_gather of per-shard top-2, 2-way merge; apply.smk:278-342)
  c1            BASELINE configs[0]: the reference's own bundled learn/apply case (7,069 proteins, alphabet 2, k = 8), N = 1 only

--workload X runs one of them alone (also: sweep = C5).  `--impl reference` times only the CPU arm (rank 0), same JSON
shape and the same `config` dict.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import traceback

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
        """index: a GPU index or a comma-separated list — ONE nvidia-smi process samples all of them (the profiling recipe's
        clocks line); one process per rank polling the driver every 100 ms perturbed 1-ms steps at 8 ranks."""
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


def _json(path):
    try:
        with open(os.path.join(ROOT, path)) as f:
            return json.load(f)
    except Exception:
        return {}


def peaks():
    v = _json("MEASURED_PEAKS.json").get("hbm_gbs")
    return (float(v), "MEASURED_PEAKS.json (measured)") if v else (6650.0, "fallback (B200_PROFILING.md)")


def tensor_peak_int8():
    """Dense int8 tensor peak in TOP/s: measured with cuBLASLt IMMA (scripts/measure_peaks.py -> profiles/peaks_extra.json);
    else 2 x the measured cuBLAS bf16 burst figure (the int8 pipe runs at twice the bf16 rate on sm_100a)."""
    v = _json("profiles/peaks_extra.json").get("int8_tops")
    if v:
        return float(v), "profiles/peaks_extra.json: cuBLASLt int8 GEMM 8192^3, CUDA events (scripts/measure_peaks.py)"
    b = _json("MEASURED_PEAKS.json").get("bf16_tflops")
    return (2.0 * float(b), "2 x measured bf16 burst (MEASURED_PEAKS.json)") if b else (2.0 * 1590.0, "2 x fallback bf16")


def fma_peak_fp32():
    v = _json("profiles/peaks_extra.json").get("fp32_fma_tflops")
    if v:
        return float(v), "profiles/peaks_extra.json: skm_bench_fma_f32, CUDA events (scripts/measure_peaks.py)"
    return 74.5, "nominal 148 SMs x 128 lanes x 2 x 1.965 GHz"


SYN6 = {"AGILMV": "A", "FWY": "F", "NQSTC": "N", "DE": "D", "KRH": "K", "P": "P"}       # 6-letter alphabet for C3/C4
SYN6_ORACLE = {"syn6": [(k, v) for k, v in SYN6.items()]}


def zipf_annotations(n, n_ann, frac_unannotated, seed):
    """annotation id per sequence ~ Zipf(1.1) over n_ann ids, -1 for the unannotated share."""
    rng = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, n_ann + 1) ** 1.1
    ids = rng.choice(n_ann, size=n, p=w / w.sum()).astype(np.int32)
    ids[rng.random(n) < frac_unannotated] = -1
    return ids


def c2_config(nseq):
    """The headline workload's config — the SAME dict in both arms."""
    return {"workload": f"C2: synthetic {nseq} proteins/GPU (lognormal len, mean~350, 0.1% X), miqs k=3, dense int32 counts",
            "alphabet": "miqs", "k": 3, "nseq_per_gpu": nseq, "basis": "first-occurrence, K=1000",
            "l2": "inputs (0.35 GB) and output (4 GB) per GPU are larger than L2",
            "parallelism": "sequence-sharded, one shard per GPU (reference arm: one shard per host process), no collective"}


def cpu_arm(res, offsets, alphabet, k, sample_seqs):
    from oracle import cpu_baseline

    n = min(sample_seqs, len(offsets) - 1)
    off = offsets[:n + 1]
    r = cpu_baseline.vectorize_sample(res[:off[-1]], off, alphabet, k)
    return {"value": r["nseq"] / r["seconds"], "unit": "sequences/s", "cores": r["cores"], "kind": "port",
            "sample": f"first {n} sequences of the workload, two-pass vectorize (basis + dense counts), "
                      f"numpy oracle port, one process per shard; {r['seconds']:.2f} s", "K": r["K"]}


# =====================================================================================================
# harness
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
        self.cpus, self.all_cpus = None, os.sched_getaffinity(0)
        if not os.environ.get("SKM_NO_CPU_BIND"):
            from snekmer_b200 import dist as D
            self.cpus = D.bind_to_local_cpus(self.local_rank)     # pinned buffers NUMA-local to this rank's GPU
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def run_cpu(self, cpu, sample):
        """The CPU baseline leg uses ALL host cores (its worker processes inherit the affinity), not only the GPU-local ones."""
        os.sched_setaffinity(0, self.all_cpus)
        try:
            return cpu(sample)
        finally:
            if self.cpus:
                os.sched_setaffinity(0, self.cpus)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def all_ok(self, ok):
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def timed(self, step, steps, warmup, clocks=True):
        """warm-up, then EXACTLY `steps` steps between barrier+sync pairs; returns total ms (this rank) and clocks."""
        torch = self.torch
        # rank 0 samples the GPUs of ALL ranks (median clock over all samples, union of the throttle reasons)
        if os.environ.get("SKM_NO_CLOCKS"):            # experiment: is the sampler what perturbs 1-ms steps at 8 ranks?
            clocks = False
        vis = [x.strip() for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip()]
        ids = [vis[i] if i < len(vis) else str(i) for i in range(self.world)]        # nvidia-smi wants PHYSICAL indices / UUIDs
        sampler = ClockSampler(",".join(ids)) if clocks and self.rank == 0 else None
        if sampler:
            sampler.start()                  # nvidia-smi needs ~100 ms per sample: it runs from the warm-up on
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
        if not clocks:
            return ms, None
        # a timed region of a few ms is over before the first sample: keep the same work running (untimed) until the
        # sampler has seen the clocks under this load — at N > 1 on EVERY rank for a fixed time, so that the one sampler
        # (rank 0) sees all GPUs under the load of the step
        t_end = time.time() + (2.0 if self.world == 1 else 0.7)
        while time.time() < t_end and (self.world > 1 or len(sampler.lines) < 5):
            step(False)
            torch.cuda.synchronize()
        return ms, (sampler.stop() if sampler else None)

    def count_launches(self, step, fallback):
        """Kernels launched by ONE step, counted from a CUPTI trace of an extra untimed step (torch.profiler sees every
        kernel of the process, libskm_b200's included).  Returns (count, {kernel name: launches}) or (fallback, None)."""
        torch = self.torch
        try:
            from torch.profiler import ProfilerActivity, profile

            torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                step(False)
                torch.cuda.synchronize()
            names = {}
            for e in prof.events():
                if str(getattr(e, "device_type", "")).endswith("CUDA"):
                    n = e.name
                    if n.lower().startswith(("memcpy", "memset")):
                        continue
                    short = n.split("<")[0].split("(")[0].replace("void ", "").strip()
                    names[short] = names.get(short, 0) + 1
            total = sum(names.values())
            if total > 0:
                return total, names
        except Exception:          # noqa: BLE001 — e.g. running under ncu: CUPTI is taken
            pass
        return fallback, None

    def free(self):
        import gc
        gc.collect()
        self.torch.cuda.empty_cache()

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def _e2e_time(ctx, fn, reps):
    """Wall clock per call of fn() (host buffers in, host results out; fn synchronises), max over ranks."""
    fn()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    ctx.barrier()
    return ctx.max_over_ranks([ms])[0]


# =====================================================================================================
# C2: vectorize (the headline)
# =====================================================================================================
def run_vectorize(ctx, args, steps, warmup):
    torch = ctx.torch
    from snekmer_b200 import engine as E
    from snekmer_b200 import pipeline as P

    alphabet, k = "miqs", 3
    world, dev = ctx.world, ctx.dev
    res_np, offsets = synth_proteins(args.nseq, 2 + 1000 * ctx.rank)
    nres = int(offsets[-1])
    h_res = torch.empty(nres, dtype=torch.uint8, pin_memory=True)
    h_res.numpy()[:] = res_np
    batch = E.SequenceBatch.from_packed(res_np, offsets, dev)
    tab = E.alphabet_tables(alphabet, dev)
    S = K = tab.nsym ** k
    out = torch.empty((batch.n, K), dtype=torch.int32, device=dev)
    first = torch.empty(S, dtype=torch.int64, device=dev)
    state = torch.zeros(4, dtype=torch.int32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kern_ms = []

    # the launch-bound front of the step (table reset, 4 walk chunks, key build, sort, emit) is ONE CUDA-graph launch
    plan = E.OrderOnlyPlan(batch, alphabet, k, capture=not os.environ.get("SKM_NO_GRAPH"))

    def step(timed):
        # pass 1: basis order (kmerize.smk:89-104), pass 2: counts (kmerize.smk:112-120); K is read back while pass 2 runs
        basis, counts = E.vectorize_order_only(batch, alphabet, k, out=out, count_events=ev if timed else None, plan=plan)
        assert basis.K == K and counts.data_ptr() == out.data_ptr()
        if timed:
            ev[1].synchronize()
            kern_ms.append(ev[0].elapsed_time(ev[1]))

    total_ms, clocks = ctx.timed(step, steps, warmup)
    launches, names = ctx.count_launches(step, 14)
    first.fill_(-1); state.zero_()
    E.basis_first_progressive(batch, alphabet, k, first, state, 0)
    saturated_at = int(first.max().item())
    e2e, e2e_ms = {}, 0.0
    if not args.no_e2e:
        reps = max(3, min(steps, 5))
        h2d = nres + 8 * (batch.n + 1)
        chk = int(out[:1000].sum().item())
        for tr, width, dt in (("uint8", K, torch.uint8), ("bits", (K + 7) // 8, torch.uint8), ("uint16", K, torch.uint16)):
            h_out = torch.empty((batch.n, width), dtype=dt, pin_memory=True)
            box = {}

            def once():
                box["r"] = P.vectorize_host(h_res, offsets, alphabet, k, out=h_out, transport=tr, device=dev)
            ms = _e2e_time(ctx, once, reps)
            r = box["r"]
            d2h = h_out.numel() * h_out.element_size() + 8 + (12 * int(r.escapes[0].size) if r.escapes else 0)
            if tr == "bits":
                assert int(r.presence()[:1000].sum()) == int((out[:1000] > 0).sum().item())
            else:
                esc = int((r.escapes[2][r.escapes[0] < 1000] - 255).sum()) if r.escapes else 0
                assert int(r.data[:1000].numpy().astype(np.int64).sum()) + esc == chk
            e2e[tr] = {"value": world * args.nseq / (ms * 1e-3), "ms_per_step": ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "pcie_gbs_per_rank": (h2d + d2h) / (ms * 1e-3) / 1e9}
            del h_out, r, box
        e2e_ms = e2e["uint8"]["ms_per_step"]
    total_ms, kern_ms_avg = ctx.max_over_ranks([total_ms, float(np.mean(kern_ms))])
    ms_per_step = total_ms / steps
    peak, peak_src = peaks()
    alg_bytes = nres + 8 * (batch.n + 1) + 4 * batch.n * K
    achieved = alg_bytes / (kern_ms_avg * 1e-3) / 1e9
    step_bytes = alg_bytes + nres + 8 * (batch.n + 1) + 24 * S          # B_vec + B_basis
    line = {
        "metric": "sequences/sec vectorize", "value": world * args.nseq / (ms_per_step * 1e-3), "unit": "sequences/s",
        "n_gpus": world, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": c2_config(args.nseq), "clocks": clocks, "gpu_launches": launches * steps, "launches_per_step": names,
        "cuda_graph": ("basis walk + finalisation replayed as one graph launch, count pass launched eagerly behind it" if plan.graph is not None
                       else "off (eager launches)"),
        "roofline": {"bound": "hbm", "kernel": "count_dense_warp_kernel", "achieved": achieved, "peak": peak,
                     "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "algorithmic_bytes": alg_bytes, "kernel_ms": kern_ms_avg,
                     "traffic": TRAFFIC.get("count_dense_warp_kernel", TRAFFIC.get("count_dense_kernel")),
                     "whole_step": {"algorithmic_bytes": step_bytes, "ms": ms_per_step, "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                                    "note": "B_vec + B_basis of SURVEY 8(d) over the whole step; the basis walk stops on the device once all "
                                            f"{S} codes have a first position (last one at residue {saturated_at} of {nres})"}},
    }
    if not args.no_e2e:
        m = e2e["uint8"]
        line["e2e"] = {"value": m["value"], "unit": "sequences/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": m["h2d_bytes_per_step"],
                       "d2h_bytes_per_step": m["d2h_bytes_per_step"], "pcie_gbs_per_rank": m["pcie_gbs_per_rank"],
                       "cpu_affinity_rank0": (f"{len(ctx.cpus)} CPUs local to the GPU (NVML): {ctx.cpus[0]}..{ctx.cpus[-1]}" if ctx.cpus else "not bound"),
                       "api": "snekmer_b200.pipeline.vectorize_host(transport='uint8'): pinned host residues/offsets -> pinned host uint8 "
                              "counts [N, K] + escape list (row, col, count) of the entries >= 255; lossless",
                       "transports": {"uint16 (2 B per count, lossless below 65,536 residues per sequence)": e2e["uint16"],
                                      "bits (the reference's own vectorize payload: bit-packed presence matrix, kmerize.smk:112-120)": e2e["bits"]}}
    cpu = lambda sample: cpu_arm(res_np, offsets, alphabet, k, sample)
    return line, cpu


# =====================================================================================================
# dense apply (tcgen05)
# =====================================================================================================
def run_apply(ctx, args, steps, warmup):
    """Dense-basis apply (C2 basis: miqs k=3, K = 1000) against a 50k-annotation matrix: counts + norms +
    tcgen05 int8 scoring GEMM with fused top-2.  Queries shard over GPUs, the matrix is replicated."""
    torch = ctx.torch
    from snekmer_b200 import engine as E
    from snekmer_b200 import pipeline as P

    alphabet, k, n_ann = "miqs", 3, args.n_ann
    nseq = args.nseq_apply
    S = K = 1000
    tr_res, tr_off = synth_proteins(200_000, 77)                      # training set (same on every rank)
    tr_ann = zipf_annotations(200_000, n_ann, 0.0, 78)
    tb = E.SequenceBatch.from_packed(tr_res, tr_off, ctx.dev)
    M, _ = E.learn_dense(tb, alphabet, k, None, torch.from_numpy(tr_ann), n_ann)
    M = M[:n_ann].contiguous()
    prep = E.prepare_annotations(M)
    assert prep is not None
    res_np, offsets = synth_proteins(nseq, 4 + 1000 * ctx.rank)
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

    total_ms, _ = ctx.timed(step, steps, warmup, clocks=False)
    launches, names = ctx.count_launches(step, 8)
    # parity inside the run: a random sample of queries re-scored by the exact CUDA-core path must agree on the
    # predictions and to 2 ulp on the scores
    r = out["r"]
    sel = torch.from_numpy(np.random.default_rng(1).choice(batch.n, size=min(2000, batch.n), replace=False)).to(ctx.dev)
    ex = E.apply_dense(Q[sel].contiguous(), M, tensor_cores=False)
    ok = bool(torch.equal(ex.top1, r.top1[sel]) and torch.equal(ex.top2, r.top2[sel]) and
              torch.allclose(ex.score1, r.score1[sel], rtol=1e-14, atol=0) and torch.allclose(ex.score2, r.score2[sel], rtol=1e-14, atol=0))
    ok = ctx.all_ok(ok)
    e2e_ms = 0.0
    if not args.no_e2e:
        h_res = torch.empty(nres, dtype=torch.uint8, pin_memory=True); h_res.numpy()[:] = res_np
        bufs = (torch.empty(batch.n, dtype=torch.int32, pin_memory=True), torch.empty(batch.n, dtype=torch.int32, pin_memory=True),
                torch.empty(batch.n, dtype=torch.float64, pin_memory=True), torch.empty(batch.n, dtype=torch.float64, pin_memory=True))
        box = {}

        def once():
            box["h"] = P.apply_host(h_res, offsets, alphabet, k, prep, out=bufs, device=ctx.dev)
        e2e_ms = _e2e_time(ctx, once, 3)
        assert np.array_equal(box["h"].top1, r.top1.cpu().numpy())
    total_ms, k_ms = ctx.max_over_ranks([total_ms, float(np.mean(kern_ms))])
    ms = total_ms / steps
    ops = 2.0 * batch.n * prep.issued_macs_per_query()              # int8 MACs x 2 actually issued (per tile: width x planes x padded K)
    useful = 2.0 * batch.n * n_ann * K
    tiles = prep.tiles()
    peak, src = tensor_peak_int8()
    rec = {"metric": "sequences/sec apply", "value": ctx.world * nseq / (ms * 1e-3), "unit": "sequences/s", "n_gpus": ctx.world,
           "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
           "dtype": "int8 x int8 -> int32 (exact), float64 scaling", "data": "synthetic",
           "config": {"workload": f"dense apply: {nseq} queries/GPU (mean len 350) vs {n_ann} annotations, miqs k=3 (K=1000), "
                                  f"{len(tiles)} annotation tiles, {float(tiles[:, 2].mean()):.2f} base-256 digit planes per tile (max {prep.n_planes}); "
                                  f"counts + norms + tcgen05 GEMM + top-2",
                      "l2": "query counts 4 GB larger than L2", "parallelism": f"query-sharded x{ctx.world}, matrix replicated"},
           "collective": None, "comm_ms": 0.0,
           "parity_check": ("ok" if ok else "FAILED") + f": {int(sel.numel())} sampled queries per rank re-scored by the exact CUDA-core path (skm_apply_dense)",
           "gpu_launches": launches * steps, "launches_per_step": names,
           "roofline": {"bound": "tensor", "kernel": "apply_tc_kernel", "achieved": ops / (k_ms * 1e-3) / 1e12, "peak": peak,
                        "peak_source": src, "unit": "TOP/s", "frac": ops / (k_ms * 1e-3) / 1e12 / peak, "kernel_ms": k_ms,
                        "algorithmic_ops": ops, "useful_ops": useful, "useful_tops": useful / (k_ms * 1e-3) / 1e12,
                        "traffic": TRAFFIC.get("apply_tc_kernel")}}
    if not args.no_e2e:
        rec["e2e"] = {"value": ctx.world * nseq / (e2e_ms * 1e-3), "unit": "sequences/s", "ms_per_step": e2e_ms,
                      "h2d_bytes_per_step": nres + 8 * (batch.n + 1), "d2h_bytes_per_step": 24 * batch.n,
                      "api": "snekmer_b200.pipeline.apply_host: pinned host residues -> (counts stay in HBM) -> tcgen05 scoring -> "
                             "pinned host top-1 / top-2 ids and scores (24 B per query)"}

    def cpu(sample):
        from oracle import cpu_baseline
        n = min(sample // 16, len(offsets) - 1)
        basis = np.arange(S, dtype=np.uint64)
        r_ = cpu_baseline.apply_dense_sample(res_np[:offsets[n]], offsets[:n + 1], alphabet, k, basis, M.cpu().numpy())
        return {"value": r_["nseq"] / r_["seconds"], "unit": "sequences/s", "cores": r_["cores"], "kind": "port",
                "sample": f"first {n} queries, numpy counts + float64 BLAS cosine + argpartition top-2 per shard; {r_['seconds']:.2f} s"}
    return rec, cpu


# =====================================================================================================
# C3: sparse learn (+ the all_to_all exchange at N > 1)
# =====================================================================================================
def run_learn(ctx, args, steps, warmup):
    """C3 shape: 6-letter alphabet, k = 8 (S = 1,679,616), 20k annotations (Zipf 1.1, 30 % unannotated).
    Step = sort-based sparse learn + Totals row + (N > 1) exchange of the COO lists and all_reduce of the Totals."""
    torch, dist = ctx.torch, ctx.dist
    from snekmer_b200 import alphabet as A
    from snekmer_b200 import engine as E

    A.register_alphabet("syn6", SYN6)
    alphabet, k, n_ann = "syn6", 8, 20000
    nseq = args.nseq_learn
    res_np, offsets = synth_proteins(nseq, 3 + 1000 * ctx.rank)
    ann = zipf_annotations(nseq, n_ann, 0.30, 30 + ctx.rank)
    nres = int(offsets[-1])
    batch = E.SequenceBatch.from_packed(res_np, offsets, ctx.dev)
    d_ann = torch.from_numpy(ann).to(ctx.dev)
    S = 6 ** 8
    state, comm_ms = {}, []
    cev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    from snekmer_b200 import dist as D
    phases = {"bounds": [], "all_to_all": [], "merge": [], "all_reduce_totals": []}
    count_bits = 64 - (n_ann * S - 1).bit_length()
    pev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]

    lev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    local_ev_ms, wall_ms, warm_ev_ms = [], [], []
    if os.environ.get("SKM_PEER_WARM") and ctx.world > 1:           # experiment: peer mappings opened, exchange through NCCL
        D.peer_buffers().ensure(1 << 27)

    def step(timed):
        t_w = time.perf_counter()
        lev[0].record()
        keys, vals, count = E.learn_sparse_with_totals(batch, alphabet, k, d_ann, n_ann)     # matrix + Totals row over ALL sequences
        lev[1].record()
        if not timed:
            lev[1].synchronize()
            warm_ev_ms.append(round(lev[0].elapsed_time(lev[1]), 3))
        state["local_nnz"] = keys.numel()
        rng_ = (0, n_ann)
        if ctx.world > 1:
            # E.exchange_coo_by_annotation, phase by phase (CUDA events on this stream; the NCCL work is ordered with it)
            if timed:
                cev[0].record(); pev[0].record()
            bounds = D.balanced_annotation_bounds(keys, S, n_ann)
            if timed:
                pev[1].record()
            kb = [a * S for a in bounds]
            done = False
            if E.peer_exchange_enabled(keys):
                # fused pack + all_to_all: one kernel stores the packed entries into the owners' buffers over NVLink
                pushed = D.push_coo_by_key_range(keys, vals, kb, count_bits)
                if pushed is not None:
                    ptr, runs, flag = pushed
                    if timed:
                        pev[2].record()
                    k3, v3, dn = E.coo_merge_runs_packed(ptr, runs, count_bits, ctx.dev)
                    m, bad = torch.cat([dn, flag.to(torch.int64)]).tolist()
                    if not bad:
                        keys, vals, done = k3[:m], v3[:m], True
                    del k3, v3
            if not done:
                k2, v2, runs = D.alltoall_coo_by_key_range(keys, vals, kb, return_runs=True)
                if timed:
                    pev[2].record()
                keys, vals = E.coo_merge_runs(k2, v2, runs)
                del k2, v2
            if timed:
                pev[3].record()
            rng_ = (bounds[ctx.rank], bounds[ctx.rank + 1])
            dist.all_reduce(count)
            if timed:
                pev[4].record()
                cev[1].record(); cev[1].synchronize()
                comm_ms.append(cev[0].elapsed_time(cev[1]))
                for i, name in enumerate(phases):
                    phases[name].append(pev[i].elapsed_time(pev[i + 1]))
        if timed:
            lev[1].synchronize()
            local_ev_ms.append(lev[0].elapsed_time(lev[1]))
            wall_ms.append((time.perf_counter() - t_w) * 1e3)
        state.update(keys=keys, vals=vals, totals=count, range=rng_)

    total_ms, _ = ctx.timed(step, steps, max(warmup, 3) + 2, clocks=False)    # the caching allocator needs two steps to hold two generations of the COO buffers
    launches, names = ctx.count_launches(step, 110)
    # ---- parity inside the run -----------------------------------------------------------------------
    keys, vals, (a_lo, a_hi) = state["keys"], state["vals"], state["range"]
    if ctx.world > 1:
        # rows of three annotations (a frequent, a mid, the rarest id) recomputed by ONE rank from the sequences of all
        # ranks must equal what the exchange left on the owning rank; the all-reduced Totals must add up
        check_ids = [50, 1000, n_ann - 1]
        sel = np.flatnonzero(np.isin(ann, check_ids))
        mine = ([res_np[offsets[i]:offsets[i + 1]].tobytes() for i in sel], ann[sel].tolist())
        parts = [None] * ctx.world
        dist.all_gather_object(parts, mine)
        seqs = [s for p in parts for s in p[0]]
        ids = np.array([a for p in parts for a in p[1]], dtype=np.int32)
        offs = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in seqs], out=offs[1:])
        sb = E.SequenceBatch.from_packed(np.frombuffer(b"".join(seqs), dtype=np.uint8), offs, ctx.dev)
        rk, rv = E.learn_sparse(sb, alphabet, k, torch.from_numpy(ids), n_ann)
        ok = True
        for a in check_ids:
            if a_lo <= a < a_hi:
                b = torch.tensor([a * S, (a + 1) * S], dtype=torch.int64, device=ctx.dev)
                i0, i1 = torch.searchsorted(keys, b).tolist()
                j0, j1 = torch.searchsorted(rk, b).tolist()
                ok = ok and (i1 - i0) == (j1 - j0) and bool(torch.equal(keys[i0:i1], rk[j0:j1]) and torch.equal(vals[i0:i1], rv[j0:j1]))
        # Totals: the all-reduced table sums to the valid windows of all ranks
        local_tot = E.kmer_totals(batch, alphabet, k).sum().to(torch.float64).reshape(1)
        dist.all_reduce(local_tot)
        ok = ok and int(local_tot.item()) == int(state["totals"].sum().item())
        ok = ctx.all_ok(ok and keys.numel() > 0 and bool((keys[1:] > keys[:-1]).all()))
        parity = (("ok" if ok else "FAILED") + f": rows of annotations {check_ids} after the exchange == one-rank recomputation from "
                  f"{len(seqs)} gathered sequences; keys strictly sorted; all-reduced Totals == sum of the ranks' window counts")
    else:
        n_chk = min(20000, batch.n)
        sub = E.SequenceBatch.from_packed(res_np[:offsets[n_chk]], offsets[:n_chk + 1], ctx.dev)
        k1, v1 = E.learn_sparse(sub, alphabet, k, d_ann[:n_chk], n_ann)
        k2, v2 = E.learn_sparse(sub, alphabet, k, d_ann[:n_chk], n_ann, method="global")
        ok = bool(torch.equal(k1, k2) and torch.equal(v1, v2))
        parity = ("ok" if ok else "FAILED") + f": single rank — grouped learn == one global 64-bit sort on the first {n_chk} sequences"
    # end to end: pinned host residues + annotation ids in, COO list out
    e2e_ms = 0.0
    nnz = int(state["keys"].numel())
    if not args.no_e2e:
        from snekmer_b200 import pipeline as P
        h_res = torch.empty(nres, dtype=torch.uint8, pin_memory=True); h_res.numpy()[:] = res_np
        h_ann = torch.from_numpy(ann).pin_memory()
        cap = int(state["local_nnz"]) + 1024
        h_w = torch.empty(cap, dtype=torch.int64, pin_memory=True)
        box = {}

        def once():
            box["r"] = P.learn_host(h_res, offsets, h_ann, alphabet, k, n_ann, out=h_w, device=ctx.dev)
        e2e_ms = _e2e_time(ctx, once, 2)
    total_ms, c_ms = ctx.max_over_ranks([total_ms, float(np.mean(comm_ms)) if comm_ms else 0.0])
    ms = total_ms / steps
    local_nnz = int(state["local_nnz"])
    alg_bytes = nres + 8 * (batch.n + 1) + 4 * batch.n + 16 * local_nnz + 8 * S
    peak, peak_src = peaks()
    rec = {"metric": "sequences/sec learn", "value": ctx.world * nseq / (ms * 1e-3), "unit": "sequences/s", "n_gpus": ctx.world,
           "steps": steps, "warmup": max(warmup, 3) + 2, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
           "dtype": "int64", "data": "synthetic",
           "config": {"workload": f"C3 shape: {nseq} proteins/GPU, 6-letter alphabet k=8 (S=1,679,616), 20k annotations Zipf(1.1), 30% unannotated; sparse COO matrix",
                      "nnz_local": local_nnz, "nnz_after_exchange_rank0": nnz, "l2": "inputs 0.44 GB, keys 3.5 GB larger than L2",
                      "parallelism": f"sequence-sharded x{ctx.world}"},
           "collective": (("all_reduce(per-annotation histogram) + skm_coo_pack_push (own kernel: packed 8-byte entries stored into the owners' "
                           "receive buffers over NVLink peer memory, fenced by all_reduce) + merge tree + all_reduce(Totals)"
                           if E.peer_exchange_enabled(state["keys"]) else
                           "all_reduce(per-annotation histogram) + all_to_all_single(COO keys, values by balanced annotation range) + merge tree + all_reduce(Totals)")
                          if ctx.world > 1 else None),
           "comm_ms": c_ms, "local_ms": ms - c_ms,
           "comm_phases_ms_this_rank": {n_: float(np.mean(v)) for n_, v in phases.items() if v} or None,
           "local_learn_ms_events_this_rank": [round(x, 3) for x in local_ev_ms] or None,
           "step_wall_ms_this_rank": [round(x, 3) for x in wall_ms] or None,
           "untimed_local_learn_ms_events_this_rank": warm_ev_ms[:8] or None,
           "learn_method": os.environ.get("SKM_LEARN_METHOD", "hybrid"),
           "parity_check": parity, "gpu_launches": launches * steps, "launches_per_step": names,
           "roofline": {"bound": "hbm", "kernel": "learn step (gather by annotation + per-slice 32-bit keys + sort + run-length encode + Totals)",
                        "achieved": alg_bytes / ((ms - c_ms) * 1e-3) / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                        "frac": alg_bytes / ((ms - c_ms) * 1e-3) / 1e9 / peak, "algorithmic_bytes": alg_bytes, "kernel_ms": ms - c_ms,
                        "traffic": None, "note": "local part of the step; B_learn of SURVEY 8(d)"}}
    if not args.no_e2e:
        rec["e2e"] = {"value": ctx.world * nseq / (e2e_ms * 1e-3), "unit": "sequences/s", "ms_per_step": e2e_ms,
                      "h2d_bytes_per_step": nres + 8 * (batch.n + 1) + 4 * batch.n,
                      "d2h_bytes_per_step": (8 if box["r"].packed is not None else 16) * box["r"].nnz,
                      "api": "snekmer_b200.pipeline.learn_host: pinned host residues + annotation ids -> pinned host COO in the packed "
                             "exchange format (key << count_bits | count, 8 B per entry; HostCOO.keys() / vals() unpack); wall clock"}

    def cpu(sample):
        from oracle import cpu_baseline
        n = min(sample, len(offsets) - 1)
        r = cpu_baseline.learn_sample(res_np[:offsets[n]], offsets[:n + 1], ann[:n], "syn6", k, extra=SYN6_ORACLE)
        return {"value": r["nseq"] / r["seconds"], "unit": "sequences/s", "cores": r["cores"], "kind": "port",
                "sample": f"first {n} sequences, numpy (annotation,k-mer) unique-count per shard + merge; {r['seconds']:.2f} s"}
    return rec, cpu


# =====================================================================================================
# C4: sparse apply (SpMM), replicated and annotation-sharded
# =====================================================================================================
def run_apply_sparse(ctx, args, steps, warmup):
    """C4 shape: 6-letter alphabet k = 8 basis (S = 1,679,616), 50k annotations; queries as CSR over codes, exact
    integer SpMM + top-2 (one CTA per query, all annotations in shared-memory accumulators)."""
    torch, dist = ctx.torch, ctx.dist
    from snekmer_b200 import alphabet as A
    from snekmer_b200 import dist as D
    from snekmer_b200 import engine as E

    alphabet, k, n_ann, S = "syn6", 8, args.n_ann, 6 ** 8
    nseq = args.nseq_sparse
    A.register_alphabet("syn6", SYN6)
    tr_res, tr_off = synth_proteins(args.ntrain, 79)
    tr_ann = zipf_annotations(args.ntrain, n_ann, 0.0, 80)
    tb = E.SequenceBatch.from_packed(tr_res, tr_off, ctx.dev)
    keys, vals = E.learn_sparse(tb, alphabet, k, torch.from_numpy(tr_ann), n_ann)
    nnz_m = int(keys.numel())
    del tb
    tile = E.SPARSE_MAX_ANN
    cscs = [E.csc_build(keys, vals, S, min(tile, n_ann - a0), a0) for a0 in range(0, n_ann, tile)]
    res_np, offsets = synth_proteins(nseq, 5 + 1000 * ctx.rank)
    nres = int(offsets[-1])
    batch = E.SequenceBatch.from_packed(res_np, offsets, ctx.dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    out, kern_ms = {}, []

    def score(rowptr, cols, cvals, parts, max_len):
        """Top-2 over the annotation slices `parts`, annotation ids GLOBAL."""
        idxs, scs = [], []
        for c in parts:
            r = E.apply_sparse(rowptr, cols, cvals, c, max_len)
            if len(parts) == 1 and c.ann_lo == 0:
                return r
            i = torch.stack([r.top1.to(torch.int64), r.top2.to(torch.int64)])
            idxs.append(torch.where(i >= 0, i + c.ann_lo, i)); scs.append(torch.stack([r.score1, r.score2]))
        return E.merge_top2(torch.stack(idxs), torch.stack(scs))

    def step(timed):
        rowptr, cols, cvals = E.count_csr(batch, alphabet, k, None)
        if timed:
            ev[0].record()
        out["r"] = score(rowptr, cols, cvals, cscs, batch.max_len)
        if timed:
            ev[1].record(); ev[1].synchronize()
            kern_ms.append(ev[0].elapsed_time(ev[1]))
        out["csr"] = (rowptr, cols, cvals)

    total_ms, _ = ctx.timed(step, steps, warmup, clocks=False)
    launches, names = ctx.count_launches(step, 8)
    rowptr, cols, cvals = out["csr"]
    nnzq = int(cols.numel())
    macs = 0                    # useful multiply-accumulates: every query entry meets every entry of its k-mer's column
    for c in cscs:
        ci = cols.to(torch.int64)
        macs += int((c.colptr[ci + 1] - c.colptr[ci]).sum().item())
    # ---- annotation-sharded layout (N > 1): every rank scores ALL queries against its slice of the annotations ----
    sharded = None
    parity = "n/a (single rank): the tests compare this path with the dense float64 oracle"
    ok = True
    if ctx.world > 1:
        # annotation ranges balanced by what a slice costs in the SpMM: every entry weighted by how often its k-mer occurs
        # (entry counts alone put the 613 largest families — half of the entries, most of the work — on one rank)
        col_tot = torch.zeros(S, dtype=torch.int64, device=ctx.dev)
        E.check(E.lib().skm_coo_colsum(keys.data_ptr(), vals.data_ptr(), keys.numel(), S, col_tot.data_ptr(), torch.cuda.current_stream().cuda_stream))
        bounds = D.balanced_annotation_bounds(keys, S, n_ann, weights=col_tot[keys % S])
        del col_tot
        a_lo, a_hi = bounds[ctx.rank], bounds[ctx.rank + 1]
        mine = [E.csc_build(keys, vals, S, min(tile, a_hi - a0), a0) for a0 in range(a_lo, a_hi, tile)]
        max_len_all = int(ctx.max_over_ranks([float(batch.max_len)])[0])
        cm, res_sh = [], {}
        ph = {"gather_csr": [], "score": [], "gather_top2_merge": []}
        sev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        nq_all = ctx.world * batch.n

        def step_sharded(timed):
            rp, cc, vv = E.count_csr(batch, alphabet, k, None)
            if timed:
                sev[0].record()
            lens = (rp[1:] - rp[:-1]).contiguous()
            all_lens = [torch.empty_like(lens) for _ in range(ctx.world)]
            dist.all_gather(all_lens, lens)                                    # same number of queries on every rank
            cc_all, vv_all = D.allgather_tables(cc, vv)                        # the query CSR of every rank
            rp_all = torch.zeros(nq_all + 1, dtype=torch.int64, device=ctx.dev)
            torch.cumsum(torch.cat(all_lens), 0, out=rp_all[1:])
            if timed:
                sev[1].record()
            if mine:
                r = score(rp_all, cc_all, vv_all, mine, max_len_all)
                t1, t2, s1, s2 = r.top1, r.top2, r.score1, r.score2
            else:                                                              # a rank without annotations contributes no candidate
                t1 = torch.full((nq_all,), -1, dtype=torch.int32, device=ctx.dev); t2 = t1.clone()
                s1 = torch.full((nq_all,), float("-inf"), dtype=torch.float64, device=ctx.dev); s2 = s1.clone()
            if timed:
                sev[2].record()
            idx, sc = D.allgather_top2(t1, t2, s1, s2, 0)                      # per-shard top-2 of all queries (ids already global)
            m = E.merge_top2(idx, sc)
            if timed:
                sev[3].record(); sev[3].synchronize()
                cm.append(sev[0].elapsed_time(sev[1]) + sev[2].elapsed_time(sev[3]))
                ph["gather_csr"].append(sev[0].elapsed_time(sev[1])); ph["score"].append(sev[1].elapsed_time(sev[2]))
                ph["gather_top2_merge"].append(sev[2].elapsed_time(sev[3]))
            res_sh["r"] = m

        sh_ms, _ = ctx.timed(step_sharded, steps, warmup, clocks=False)
        m, r = res_sh["r"], out["r"]
        q0, q1 = ctx.rank * batch.n, (ctx.rank + 1) * batch.n
        ok = bool(torch.equal(m.top1[q0:q1], r.top1) and torch.equal(m.score1[q0:q1], r.score1) and
                  torch.equal(m.score2[q0:q1], r.score2) and torch.equal(m.top2[q0:q1], r.top2))
        ok = ctx.all_ok(ok)
        mine_ph = [float(np.mean(ph[n_])) for n_ in ("gather_csr", "score", "gather_top2_merge")]
        t_ph = torch.tensor(mine_ph, dtype=torch.float64, device=ctx.dev)
        all_ph = [torch.empty_like(t_ph) for _ in range(ctx.world)]
        dist.all_gather(all_ph, t_ph)
        sh_ms, sh_comm = ctx.max_over_ranks([sh_ms, float(np.mean(cm))])
        sharded = {"ms_per_step": sh_ms / steps, "comm_ms": sh_comm, "value": ctx.world * nseq / (sh_ms / steps * 1e-3),
                   "annotation_ranges": bounds,
                   "phases_ms_per_rank [gather_csr, score, gather_top2+merge]": [[round(x, 2) for x in t.tolist()] for t in all_ph],
                   "note": "a rank that finishes its slice early waits in the top-2 all_gather: comm_ms includes that skew",
                   "collective": "all_gather(query CSR: row lengths, codes, counts) + all_gather(per-shard top-2 ids, scores) + 2-way merge (skm_top2_merge)"}
        parity = ("ok" if ok else "FAILED") + ": annotation-sharded top-2 (ids and float64 scores) of this rank's queries == the replicated-matrix result, bit for bit"
    e2e_ms = 0.0
    if not args.no_e2e:
        h_res = torch.empty(nres, dtype=torch.uint8, pin_memory=True); h_res.numpy()[:] = res_np

        def once():
            b = E.SequenceBatch.from_packed(h_res.numpy(), offsets, ctx.dev, pinned=True)
            r = score(*E.count_csr(b, alphabet, k, None), cscs, batch.max_len)
            return r.top1.cpu(), r.top2.cpu(), r.score1.cpu(), r.score2.cpu()
        e2e_ms = _e2e_time(ctx, once, 3)
    total_ms, k_ms = ctx.max_over_ranks([total_ms, float(np.mean(kern_ms))])
    ms = total_ms / steps
    hbm, hbm_src = peaks()
    fma, fma_src = fma_peak_fp32()
    entry_bytes = 4 if all(c.packed is not None for c in cscs) else 8      # packed CSC: annotation << 16 | value in one word
    gathered = entry_bytes * macs + 8 * nnzq + 40 * batch.n                # what the gather formulation touches (L2 + HBM)
    compulsory = 12 * nnzq + 12 * nnz_m + 16 * batch.n
    t_bytes, t_flops = compulsory / (hbm * 1e9), 2.0 * macs / (fma * 1e12)
    t_roof = max(t_bytes, t_flops)
    rec = {"metric": "sequences/sec apply (sparse)", "value": ctx.world * nseq / (ms * 1e-3), "unit": "sequences/s", "n_gpus": ctx.world,
           "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
           "dtype": "uint32 exact dots, float64 scaling", "data": "synthetic",
           "config": {"workload": f"C4 shape: {nseq} queries/GPU vs {n_ann} annotations learned from {args.ntrain} proteins, 6-letter k=8 "
                                  f"(S=1,679,616), nnz(M)={nnz_m}, nnz(Q)={nnzq}; CSR counts + exact integer SpMM + top-2",
                      "l2": "CSC 0.54 GB larger than L2", "parallelism": f"query-sharded x{ctx.world}, matrix replicated in {len(cscs)} annotation slice(s)"},
           "collective": sharded["collective"] if sharded else None, "comm_ms": sharded["comm_ms"] if sharded else 0.0,
           "annotation_sharded": sharded, "parity_check": parity, "gpu_launches": launches * steps, "launches_per_step": names,
           "roofline": {"bound": "hbm" if t_bytes >= t_flops else "fp32-fma", "kernel": "apply_sparse_kernel", "kernel_ms": k_ms,
                        "achieved": 2.0 * macs / (k_ms * 1e-3) / 1e12, "peak": fma, "unit": "TFLOP/s", "peak_source": fma_src,
                        "frac": t_roof / (k_ms * 1e-3),
                        "definition": "SURVEY 8(d): roofline time = max(useful flops / fp32-FMA peak, compulsory bytes / HBM); frac = roofline time / kernel time",
                        "useful_flops": 2.0 * macs, "flops_term_ms": t_flops * 1e3, "compulsory_bytes": compulsory, "bytes_term_ms": t_bytes * 1e3,
                        "hbm_peak": hbm, "hbm_peak_source": hbm_src,
                        "gathered_bytes": gathered, "gathered_gbs": gathered / (k_ms * 1e-3) / 1e9, "gathered_frac_of_hbm": gathered / (k_ms * 1e-3) / 1e9 / hbm,
                        "traffic": TRAFFIC.get("apply_sparse_kernel"),
                        "note": f"gathered_* counts every CSC word the gather formulation touches ({entry_bytes} B per multiply-accumulate, mostly L2 hits)"}}
    if not args.no_e2e:
        rec["e2e"] = {"value": ctx.world * nseq / (e2e_ms * 1e-3), "unit": "sequences/s", "ms_per_step": e2e_ms,
                      "h2d_bytes_per_step": nres + 8 * (batch.n + 1), "d2h_bytes_per_step": 24 * batch.n,
                      "api": "SequenceBatch.from_packed(pinned host) + count_csr + apply_sparse -> host top-2 ids / scores (wall clock)"}
    return rec, None


# =====================================================================================================
# C1: the reference's own bundled learn / apply case
# =====================================================================================================
C1_DIR = os.path.join(ROOT, "tests", "golden", "c1")
C1_FILES = ["UP000322080_2603819", "UP000322981_424902"]


def _c1_inputs():
    import gzip

    files = []
    for nb in C1_FILES:
        ids, seqs, cur = [], [], None
        with gzip.open(os.path.join(C1_DIR, nb + ".fasta.gz"), "rt") as f:
            for line in f:
                line = line.rstrip("\r\n")
                if line.startswith(">"):
                    if cur is not None:
                        seqs.append("".join(cur))
                    ids.append(line[1:].split(None, 1)[0])
                    cur = []
                elif cur is not None:
                    cur.append(line.strip())
        if cur is not None:
            seqs.append("".join(cur))
        files.append((ids, seqs))
    ann = {}
    with open(os.path.join(C1_DIR, "c1.ann")) as f:
        next(f)
        for line in f:
            a, b = line.rstrip("\n").split("\t")
            ann[a] = b
    return files, ann


def run_c1(ctx, args, steps, warmup):
    """BASELINE configs[0]: .test/config_learnapp.yaml (alphabet 2, k 8) on the bundled proteomes — vectorize both
    files, learn the per-annotation count matrices, merge, score every sequence against the merged matrix (eval_apply).
    value = device path on resident batches; e2e = the in-memory rule cores from Python strings to host results."""
    torch = ctx.torch
    from snekmer_b200 import engine as E
    from snekmer_b200 import rules as R

    a, k = 2, 8
    files, ann = _c1_inputs()
    nseq = sum(len(f[0]) for f in files)
    names = sorted(set(ann.values()))
    aidx = {x: i for i, x in enumerate(names)}
    batches, ann_ids = [], []
    for ids, seqs in files:
        batches.append(E.SequenceBatch.from_strings(seqs))
        ann_ids.append(torch.from_numpy(np.array([aidx.get(ann.get(i.split("|")[1], ""), -1) for i in ids], dtype=np.int32)))
    nres = sum(b.nres for b in batches)
    out = {}

    def step(timed):
        Ms, bases, Cs = [], [], []
        for b, ai in zip(batches, ann_ids):
            basis = E.build_basis(b, a, k, 0, counts=False)
            C = E.count_dense(b, a, k, basis)
            M, tot = E.learn_dense(b, a, k, basis, ai, len(names))
            Ms.append(M); bases.append(basis); Cs.append(C)
        # merge: both files hold the saturated 3^8 space; file 2's columns are re-indexed onto file 1's order
        idx = bases[1].col_of_code[bases[0].codes].to(torch.int64)                 # column in file 2 of file 1's k-mer j
        Mm = (Ms[0][:len(names)] + E.gather_columns(Ms[1][:len(names)].contiguous(), idx)).contiguous()
        out["r"] = [E.apply_dense(Cs[0], Mm), E.apply_dense(E.gather_columns(Cs[1], idx), Mm)]

    total_ms, _ = ctx.timed(step, steps, warmup, clocks=False)
    launches, lnames = ctx.count_launches(step, 60)
    ms = total_ms / steps

    def chain():
        res = []
        for ids, seqs in files:
            v = R.vectorize_records(ids, seqs, a, k)
            res.append((v, R.learn_counts(v.ids, v.seqs, v.kmerlist, ann)))
        tables = [R.CountsTable(["Totals"] + lr.annotations, lr.kmerlist, np.concatenate([[lr.total_seqs], lr.seq_count]),
                                np.concatenate([[lr.totals.sum()], lr.M.sum(axis=1)]), np.concatenate([lr.totals[None], lr.M])) for _, lr in res]
        merged = R.merge_tables(tables)
        return [R.cosine_top2(v.seqs, v.kmerlist, merged).top1.cpu() for v, _ in res]
    e2e_ms = 0.0
    if not args.no_e2e:
        e2e_ms = _e2e_time(ctx, chain, 2)
    rec = {"metric": "sequences/sec vectorize+learn+apply (C1)", "value": nseq / (ms * 1e-3), "unit": "sequences/s", "n_gpus": 1,
           "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "replicas only",
           "dtype": "int32 counts, int64 matrix, float64 cosine",
           "data": "the reference's bundled proteomes (tests/golden/c1, 7,069 proteins) + the survey's seeded annotations",
           "config": {"workload": f"C1: .test/config_learnapp.yaml (alphabet 2, k 8, K = 6,561) on {nseq} bundled proteins ({nres} residues), 40 annotations: "
                                  "vectorize x2 + learn x2 + merge + cosine top-2 of every sequence"},
           "collective": None, "comm_ms": 0.0, "parity_check": "tests/test_gpu_c1.py compares this chain with the files the unmodified reference wrote",
           "gpu_launches": launches * steps, "launches_per_step": lnames, "roofline": None}
    if not args.no_e2e:
        rec["e2e"] = {"value": nseq / (e2e_ms * 1e-3), "unit": "sequences/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": 2 * nres, "d2h_bytes_per_step": 8 * nseq,
                      "api": "rules.vectorize_records + learn_counts + merge_tables + cosine_top2 (Python strings in, host arrays out; dominated by host string handling)"}

    def cpu(sample):
        from oracle import skm_oracle as O
        t0 = time.perf_counter()
        lut, syms = O.build_lut(a)
        per = []
        for ids, seqs in files:
            res, offs = O.pack(seqs)
            si, pos, code, valid = O.window_codes(res, offs, lut, len(syms), k)
            basis, _ = O.basis_codes(si, pos, code, valid, 0)
            C = O.count_matrix(si, code, valid, len(seqs), basis)
            anns, M, nsq, totals, _ = O.learn_matrix(ids, C, ann)
            per.append((basis, C, anns, M))
        order = {int(c): j for j, c in enumerate(per[0][0])}
        idx = np.array([order[int(c)] for c in per[1][0]])
        rows = per[0][2] + [x for x in per[1][2] if x not in set(per[0][2])]
        Mm = np.zeros((len(rows), len(per[0][0])), dtype=np.int64)
        for (basis, C, anns, M), ci in ((per[0], np.arange(len(per[0][0]))), (per[1], idx)):
            for r_, an in enumerate(anns):
                Mm[rows.index(an), ci] += M[r_]
        for (basis, C, anns, M), ci in ((per[0], np.arange(len(per[0][0]))), (per[1], idx)):
            Q = np.zeros((C.shape[0], Mm.shape[1]), dtype=np.int64)
            Q[:, ci] = C
            O.top2(O.cosine_scores(Q, Mm))
        dt = time.perf_counter() - t0
        return {"value": nseq / dt, "unit": "sequences/s", "cores": 1, "kind": "port",
                "sample": f"the whole C1 case (7,069 proteins), numpy oracle port, one process; {dt:.2f} s "
                          "(the reference's own rule bodies need 156 s on one core for it, SURVEY 3.5)"}
    return rec, cpu


# =====================================================================================================
# C5: alphabet / k sweep
# =====================================================================================================
SWEEP_POINTS = [("hydro", 8), ("hydro", 14), ("solvacc", 8), ("solvacc", 14), ("standard", 4), ("standard", 8), ("standard", 12),
                ("miqs", 3), ("miqs", 6), ("miqs", 10), ("miqs", 14), (None, 2), (None, 5), (None, 8), (None, 11), (None, 14)]


def run_sweep(ctx, args, steps, warmup):
    """C5: alphabet / k sweep (2..20 letters, k = 2..14).  One step = every point once: basis + per-sequence
    counts through engine.vectorize (table kernels up to 2^27 codes, sort-based wide path beyond)."""
    torch = ctx.torch
    from snekmer_b200 import engine as E

    nseq = args.nseq_sparse
    res_np, offsets = synth_proteins(nseq, 5 + 1000 * ctx.rank)
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

    total_ms, clocks = ctx.timed(step, steps, warmup)
    launches, names = ctx.count_launches(step, 300)
    (total_ms,) = ctx.max_over_ranks([total_ms])
    ms = total_ms / steps
    pts = []
    alg_total = 0
    peak, peak_src = peaks()
    for i, (a, k) in enumerate(points):
        tab = E.alphabet_tables(a, ctx.dev)
        path, K, nnz = info[i]
        m = float(np.mean(per_ms[i]))
        # B_vec (dense: R + 8(N+1) + 4NK; sparse: R + 16(N+1) + 12 nnz) + B_basis (R + 8(N+1) + 24 K)
        vec_b = nres + 8 * (batch.n + 1) + (4 * batch.n * K if path == "dense" else 8 * (batch.n + 1) + (12 if path == "csr" else 16) * nnz)
        alg = vec_b + nres + 8 * (batch.n + 1) + 24 * K
        alg_total += alg
        pts.append({"alphabet": str(a), "nsym": tab.nsym, "k": k, "log2_space": round(k * float(np.log2(tab.nsym)), 1), "path": path, "K": K,
                    "nnz": nnz, "ms": m, "seq_per_s": ctx.world * nseq / (m * 1e-3), "hbm_frac": alg / (m * 1e-3) / 1e9 / peak})
    line = {"metric": "sequences/sec vectorize (alphabet/k sweep)", "value": ctx.world * nseq * len(points) / (ms * 1e-3),
            "unit": "sequences/s", "n_gpus": ctx.world, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "uint32/uint64 codes, int32 counts", "data": "synthetic",
            "config": {"workload": f"C5: {len(points)} (alphabet, k) points x {nseq} proteins/GPU; value = point-vectorisations of a sequence per second",
                       "points": pts, "l2": "inputs 0.35 GB x N/1e6, keys 8-16 B per residue: larger than L2",
                       "parallelism": f"sequence-sharded x{ctx.world}, no collective"},
            "clocks": clocks, "gpu_launches": launches * steps, "launches_per_step": names,
            "roofline": {"bound": "hbm", "kernel": "whole sweep step (table kernels + radix / segmented sorts)", "achieved": alg_total / (ms * 1e-3) / 1e9,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s",
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
    TRAFFIC.update({k: v for k, v in _json("profiles/traffic.json").items() if not k.startswith("_")})


# =====================================================================================================
# reference arm: the CPU port alone
# =====================================================================================================
def reference_arm(args):
    """The CPU arm alone (rank 0): the oracle port of the reference's numpy path on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    alphabet, k = "miqs", 3
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
           "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": c2_config(args.nseq),
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


RUNNERS = {"vectorize": run_vectorize, "apply": run_apply, "learn": run_learn, "apply_sparse": run_apply_sparse, "c1": run_c1, "sweep": run_sweep}


def main():
    _capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all"] + sorted(RUNNERS),
                    help="all (default) = the headline C2 vectorize line + the apply / learn / apply_sparse / c1 sub-records under "
                         "'workloads'; a name runs that workload alone")
    ap.add_argument("--nseq", type=int, default=1_000_000, help="C2: sequences per GPU")
    ap.add_argument("--nseq-apply", type=int, default=1_000_000, help="dense apply: queries per GPU")
    ap.add_argument("--nseq-learn", type=int, default=1_250_000, help="C3: proteins per GPU")
    ap.add_argument("--nseq-sparse", type=int, default=200_000, help="C4 / C5: sequences per GPU")
    ap.add_argument("--n-ann", type=int, default=50000)
    ap.add_argument("--ntrain", type=int, default=400_000)
    ap.add_argument("--sub-steps", type=int, default=0, help="steps of the sub-workloads (0 = min(steps, 5), at least 3)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="sequences in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return
    _load_traffic()
    ctx = Ctx(args)
    ncores = os.cpu_count() or 1
    cpu_sample = args.cpu_sample or 20000 * min(ncores, 64)
    want_cpu = not args.no_cpu and ctx.world == 1 and ctx.rank == 0
    sub_steps = args.sub_steps or max(3, min(args.steps, 5))
    if args.workload != "all":
        line, cpu = RUNNERS[args.workload](ctx, args, args.steps, args.warmup)
        if want_cpu and cpu is not None:
            line["cpu_baseline"] = ctx.run_cpu(cpu, cpu_sample)
        if ctx.rank == 0 or ctx.world == 1:          # (RANK=1 WORLD_SIZE=1 reruns another rank's shard on one GPU)
            _emit(line)
        ctx.finish()
        return
    line, cpu = run_vectorize(ctx, args, args.steps, args.warmup)
    if want_cpu:
        line["cpu_baseline"] = ctx.run_cpu(cpu, cpu_sample)
    del cpu
    ctx.free()
    line["workloads"] = {}
    subs = ["apply", "learn", "apply_sparse"] + (["c1"] if ctx.world == 1 else [])
    failed = []
    for name in subs:
        t0 = time.time()
        try:
            rec, cpu = RUNNERS[name](ctx, args, sub_steps, 3)
            if want_cpu and cpu is not None:
                rec["cpu_baseline"] = ctx.run_cpu(cpu, cpu_sample if name != "learn" else cpu_sample // 2)
            rec["wall_s"] = round(time.time() - t0, 1)
            del cpu
        except Exception as e:          # noqa: BLE001 — the headline line must still be printed; the failure is in the record
            rec = {"error": f"{type(e).__name__}: {e}", "traceback": traceback.format_exc()[-1500:]}
            failed.append(name)
        line["workloads"][name] = rec
        ctx.free()
    line["gpu_launches"] += sum(int(r.get("gpu_launches", 0)) for r in line["workloads"].values())
    line["parity_check"] = ("ok" if not failed and all("FAILED" not in str(r.get("parity_check", "")) for r in line["workloads"].values())
                            else f"FAILED: {failed}")
    if ctx.rank == 0:
        _emit(line)
    ctx.finish()


if __name__ == "__main__":
    main()
