import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from timewarp_b200 import _lib
lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
print("n_mma commit_every N ts wait_each | issue_cycles total_cycles per_mma")
for N in (128, 256, 64):
    for ts in (0, 1):
        for ce, we in ((384, 0), (24, 0), (8, 0), (4, 0), (1, 0), (24, 1), (8, 1)):
            for rep in range(2):
                _lib.check(lib.tw_debug_umma_timing(384, ce, N, ts, we, out.data_ptr(), None), "timing")
                torch.cuda.synchronize()
            a, b = out.tolist()
            print(f"384 {ce:4d} {N:4d} {ts} {we} | {a:8d} {b:8d} {b/384:7.1f}")
