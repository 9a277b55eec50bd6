"""Issue rate of M128 x N x K16 tcgen05 MMAs as a function of WHERE the shared-memory operands live (byte offsets of the A
units and of the B tile inside the dynamic shared-memory window), with real or zero operand data.
Usage: python tools/umma_timing2.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from timewarp_b200 import _lib
_lib.load()
dll = C.CDLL(os.path.join(os.path.dirname(_lib.__file__), "libtimewarp_b200.so"))
fn = dll.tw_debug_umma_timing3
fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
fn.restype = C.c_int
out = torch.zeros(2, dtype=torch.int64, device="cuda")
def t(N, ts, fill, units, a_off, b_off):
    for rep in range(2):
        assert fn(512, N, ts, out.data_ptr(), fill, units, a_off * 1024, b_off * 1024, None) == 0, (N, ts, a_off, b_off)
        torch.cuda.synchronize()
    return out.tolist()[1] / 512
print("   N |  SS(1 A unit)  SS(2 A units)  TS     [A at 0 KB, B at 64 KB, real data]")
for N in (256, 240, 160, 128, 96, 80, 64, 32):
    print(f"{N:4d} | {t(N, 0, 1, 1, 0, 64):8.1f} {t(N, 0, 1, 2, 0, 64):12.1f} {t(N, 1, 1, 1, 0, 64):10.1f}")
print("placement scan, SS N=80: A offset (KB) x B offset (KB)")
offs = (0, 32, 64, 128, 160)
print("      " + " ".join(f"{b:6d}" for b in offs))
for a in offs:
    print(f"{a:5d} " + " ".join((f"{t(80, 0, 1, 1, a, b):6.1f}" if abs(a - b) >= 32 else "     -") for b in offs))
