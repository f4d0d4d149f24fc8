"""Write-only and copy bandwidth of this GPU (CUDA events), to put count_dense's 92 %-write stream in context."""
import torch
dev = torch.device("cuda", 0)
n = 1 << 30                                   # 4 GiB of int32
a = torch.empty(n, dtype=torch.int32, device=dev)
b = torch.empty(n, dtype=torch.int32, device=dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.zero_());        print(f"zero_ (memset)      4 GiB: {ms:.3f} ms  {4 * 2**30 / ms / 1e6:.0f} GB/s written")
ms = t(lambda: a.fill_(7));       print(f"fill_ (kernel)      4 GiB: {ms:.3f} ms  {4 * 2**30 / ms / 1e6:.0f} GB/s written")
ms = t(lambda: b.copy_(a));       print(f"copy_ 4 GiB -> 4 GiB     : {ms:.3f} ms  {8 * 2**30 / ms / 1e6:.0f} GB/s read+written")
ms = t(lambda: a.sum());          print(f"sum (read only)     4 GiB: {ms:.3f} ms  {4 * 2**30 / ms / 1e6:.0f} GB/s read")
