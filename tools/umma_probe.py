"""Run the tcgen05 probe over a grid of operand placements/layouts and print the error of each."""
import itertools
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from timewarp_b200 import _lib

AM = {0: "A smem K-major SW128", 1: "A smem K-major noswz", 2: "A smem MN-major SW128", 3: "A TMEM"}
BM = {0: "B K-major SW128", 1: "B K-major noswz", 2: "B MN-major SW128"}


def run(N, K, a_mode, b_mode, d_col=0, a_col=256, seed=0):
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(128, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    out = torch.full((128, N), float("nan"), device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(lib.tw_debug_umma_probe(A.data_ptr(), B.data_ptr(), out.data_ptr(), N, K, a_mode, b_mode, d_col, a_col, status.data_ptr(), None), "probe")
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ B.bfloat16().float().T
    err = (out - ref).abs().max().item()
    return err, int(status.item()), ref.abs().max().item()


if __name__ == "__main__":
    cases = []
    for a, b in itertools.product(range(4), range(3)):
        cases.append((128, 64, a, b, 0))
    cases += [(256, 128, 0, 0, 0), (256, 128, 3, 0, 0), (128, 128, 0, 0, 128), (80, 80, 0, 1, 0), (80, 80, 1, 1, 0), (80, 80, 0, 1, 8), (80, 80, 0, 1, 65),
              (72, 80, 0, 1, 0), (64, 256, 3, 0, 64), (128, 256, 0, 0, 384), (16, 16, 0, 0, 0), (48, 48, 1, 1, 0)]
    for N, K, a, b, dc in cases:
        try:
            err, st, mx = run(N, K, a, b, dc)
            print(f"N={N:3d} K={K:3d} {AM[a]:22s} {BM[b]:18s} d_col={dc:3d}: max|err|={err:.3e} (ref max {mx:.1f}) status={st} {'OK' if err < 1e-2 and st == 0 else 'FAIL'}")
        except Exception as e:
            print(f"N={N} K={K} a={a} b={b} d_col={dc}: EXC {e}")
            break
