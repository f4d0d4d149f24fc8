#!/usr/bin/env python
"""Extra peaks for the scoring rooflines, measured the way MEASURED_PEAKS.json measures bf16 (CUDA events, best of N
after warm-up; SURVEY 8(d) asks for them): dense int8 tensor throughput (cuBLASLt IMMA through torch._int_mm,
8192^3, 2*N^3 ops) and fp32 FMA throughput (libskm_b200's skm_bench_fma_f32: 148 x 16 CTAs x 256 threads x 16
independent chains).  Writes profiles/peaks_extra.json.

    python scripts/measure_peaks.py [out.json]
"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snekmer_b200._native import check, lib  # noqa: E402


def best_ms(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "peaks_extra.json")
    dev = torch.device("cuda", 0)
    res = {"gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__}
    n = 8192
    a = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
    b = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
    try:
        ms = best_ms(lambda: torch._int_mm(a, b))
        res["int8_tops"] = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
        res["int8_how"] = f"torch._int_mm (cuBLASLt int8 x int8 -> int32) {n}^3, 2*N^3 ops, best of 10, CUDA events: {ms:.3f} ms"
    except Exception as e:          # noqa: BLE001
        res["int8_tops"] = None
        res["int8_how"] = f"torch._int_mm failed: {e!r}"
    x = torch.randn(n, n, dtype=torch.bfloat16, device=dev)
    y = torch.randn(n, n, dtype=torch.bfloat16, device=dev)
    ms = best_ms(lambda: torch.matmul(x, y))
    res["bf16_tflops_here"] = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    blocks = sms * 16
    d_out = torch.zeros(blocks * 256, dtype=torch.float32, device=dev)
    flops = C.c_double(0.0)
    iters = 1 << 16

    def fma():
        check(lib().skm_bench_fma_f32(iters, blocks, d_out.data_ptr(), C.byref(flops), torch.cuda.current_stream().cuda_stream))
    ms = best_ms(fma)
    res["fp32_fma_tflops"] = flops.value / (ms * 1e-3) / 1e12
    res["fp32_fma_how"] = (f"skm_bench_fma_f32: {blocks} CTAs x 256 threads x 16 independent chains x {iters} FMAs, 2 flops each, "
                           f"best of 10, CUDA events: {ms:.3f} ms; nominal {sms} SMs x 128 lanes x 2 x 1.965 GHz = {sms * 128 * 2 * 1.965e9 / 1e12:.1f} TFLOP/s")
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
