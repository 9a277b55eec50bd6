"""Cycles per block of TS MMAs with different bookkeeping around the block (tcgen05.commit, fences, elect.sync, mbarrier polls).
Usage: python tools/umma_overhead.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from timewarp_b200 import _lib
_lib.load()
dll = C.CDLL(os.path.join(os.path.dirname(_lib.__file__), "libtimewarp_b200.so"))
fn = dll.tw_debug_umma_overhead
fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
fn.restype = C.c_int
out = torch.zeros(2, dtype=torch.int64, device="cuda")
def t(per_block, N, flags, n_blocks=256):
    for rep in range(2):
        assert fn(n_blocks, per_block, N, flags, out.data_ptr(), None) == 0
        torch.cuda.synchronize()
    return out.tolist()[1] / n_blocks
names = {0: "one elected thread, no bookkeeping", 1: "+ commit per block", 17: "+ 2 commits per block", 2: "+ fence::after per block", 8: "+ try_wait per block",
         4: "elect.sync + __syncwarp per block", 5: "elect + commit", 7: "elect + commit + fence", 15: "elect + commit + fence + try_wait", 31: "elect + 2 commits + fence + try_wait"}
for per_block, N in ((0, 80), (4, 80), (8, 80), (15, 80), (8, 160)):
    hw = per_block * (N / 2 + 0.9)
    print(f"--- {per_block} TS MMAs of N = {N} per block (tensor time {hw:.0f} cycles)")
    for fl, nm in names.items():
        print(f"   flags {fl:2d} {nm:42s} {t(per_block, N, fl):8.1f} cycles per block")
