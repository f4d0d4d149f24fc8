"""Debug: per-tile timeline of CTA pair 0 of apply_tc_kernel (library built with SKM_EXTRA_NVCC=-DSKM_TC_PROF)."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from snekmer_b200 import engine as E, _native
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
nq, ann = 148 * 128, 50000
Q = torch.randint(0, 4, (nq, K), generator=g, device=dev, dtype=torch.int32)
M = torch.randint(0, 201, (ann, K), generator=g, device=dev, dtype=torch.int64)
prep = E.prepare_annotations(M)
qn2 = E.row_norm2(Q)
for _ in range(3):
    E.apply_tc(Q, prep, qn2)
torch.cuda.synchronize()
lib = _native.lib()
fn = ctypes.CDLL(_native.LIB_PATH if hasattr(_native, "LIB_PATH") else os.path.join(os.path.dirname(_native.__file__), "lib", "libskm_b200.so")).skm_debug_tc_timeline
buf = np.zeros((6, 512), dtype=np.int64)
fn(buf.ctypes.data_as(ctypes.c_void_p))
t0 = buf[0, 0]
n = 196
names = ["mma_start", "mma_issued", "epiA_start", "epiA_end", "epiB_start", "epiB_end"]
print("tile " + " ".join(f"{x:>11s}" for x in names) + "   period  epiA  epiB")
for t in list(range(0, 24)) + list(range(24, n, 8)):
    row = buf[:, t] - t0
    per = buf[0, t] - buf[0, t - 1] if t else 0
    print(f"{t:4d} " + " ".join(f"{int(x):11d}" for x in row) + f" {int(per):8d} {int(buf[3,t]-buf[2,t]):5d} {int(buf[5,t]-buf[4,t]):5d}")
print("total", int(buf[3, n - 1] - t0), "per tile", (buf[3, n - 1] - t0) / n)
