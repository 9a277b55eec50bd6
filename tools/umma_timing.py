import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from timewarp_b200 import _lib
lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
print("n_mma   N ts noswz | issue_cycles total_cycles per_mma")
for N in (256, 240, 160, 128, 96, 80, 64, 32):
    for ts in (0, 1):
        for nosw in (0, 1):
            for rep in range(2):
                _lib.check(lib.tw_debug_umma_timing(512, N, ts, nosw, out.data_ptr(), None), "timing")
                torch.cuda.synchronize()
            a, b = out.tolist()
            print(f"512 {N:4d} {ts}   {nosw}   | {a:8d} {b:8d} {b/512:7.1f}")
