"""is a single-pass `kind::tf32` tensor-core path (2 units of tensor time per product instead of the 3 of
the bf16 hi/lo split) inside the parity budget?  CPU emulation on the bench configuration (FULL model, 2olx-65, proposal and
init-scale weights): every GEMM input of the oracle is rounded the way the tensor core would see it, the fp32 accumulation is
kept, and log p / the MH exponent are compared with the oracle's fp64 run.

    python tools/tf32_numerics.py            # ~2 min on 8 cores; prints a table and writes profiles/r02_tf32_numerics.json

modes: fp32 (no rounding), tf32 (both operands rounded to 10 explicit mantissa bits, round-to-nearest), tf32w (weights pre-rounded,
activations TRUNCATED as the hardware does without a cvt.rna), bf16 (both operands 7 bits), bf16x3 (a w ~ hi hi + lo hi + hi lo).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import flow_oracle as fo  # noqa: E402
from timewarp_b200.peptides import tetrapeptide_2olx  # noqa: E402

F = torch.nn.functional
_linear = F.linear


def rna(x, drop_bits):  # round fp32 to nearest keeping 23 - drop_bits mantissa bits
    i = x.contiguous().view(torch.int32)
    half = 1 << (drop_bits - 1)
    return ((i + half) & ~((1 << drop_bits) - 1)).view(torch.float32)


def trunc(x, drop_bits):
    return (x.contiguous().view(torch.int32) & ~((1 << drop_bits) - 1)).view(torch.float32)


def make_linear(mode):
    def lin(x, w, b=None):
        if x.dtype != torch.float32 or mode == "fp32":
            return _linear(x, w, b)
        if mode == "tf32":
            y = _linear(rna(x, 13), rna(w, 13))
        elif mode == "tf32w":
            y = _linear(trunc(x, 13), rna(w, 13))
        elif mode == "bf16":
            y = _linear(rna(x, 16), rna(w, 16))
        elif mode == "bf16x3":
            xh, wh = rna(x, 16), rna(w, 16)
            xl, wl = rna(x - xh, 16), rna(w - wh, 16)
            y = _linear(xh, wh) + _linear(xl, wh) + _linear(xh, wl)
        else:
            raise ValueError(mode)
        return y if b is None else y + b

    return lin


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    pep = tetrapeptide_2olx()
    o = fo.OracleConfig()
    B, V = 16, pep.num_atoms
    out = {}
    for weights in ("proposal", "init"):
        sd = bench.bench_state_dict(o, weights)
        sd64 = fo.to_dtype(sd, torch.float64)
        g = torch.Generator().manual_seed(0)
        x = torch.tensor(pep.coords_nm, dtype=torch.float32)[None] + 0.005 * torch.randn(B, V, 3, generator=g)
        at = torch.tensor(pep.atom_types)[None].repeat(B, 1)
        mask = torch.zeros(B, V, dtype=torch.bool)
        xv = torch.randn(B, V, 3, generator=g)
        zc = torch.randn(1, B, V, 3, generator=g) * torch.exp(sd["coords_prior_log_scale"])
        zv = torch.randn(1, B, V, 3, generator=g) * torch.exp(sd["velocs_prior_log_scale"])
        with torch.no_grad():
            yc, yv, pxy64 = fo.conditional_sample_with_logp(sd64, o, at, x.double(), xv.double(), mask, 1, zc.double(), zv.double(), distance_mode="direct")
            pyx64 = fo.log_likelihood(sd64, o, at, yc[0], yv[0], x.double(), xv.double(), mask, distance_mode="direct")
            y32, yv32 = yc[0].float(), yv[0].float()
            res = {}
            for mode in ("fp32", "tf32", "tf32w", "bf16", "bf16x3"):
                F.linear = make_linear(mode)
                try:
                    # same proposal y for every mode (the fp64 one, rounded): isolates the density error
                    pyx = fo.log_likelihood(sd, o, at, y32, yv32, x, xv, mask, distance_mode="direct")
                    _, _, pxy = fo.conditional_sample_with_logp(sd, o, at, x, xv, mask, 1, zc, zv, distance_mode="direct")
                finally:
                    F.linear = _linear
                e_yx = (pyx.double() - pyx64)
                e_xy = (pxy[0].double() - pxy64[0])
                dexp = (e_xy - e_yx).abs()  # error of the MH exponent (energies excluded)
                res[mode] = {"rel_err_logp_max": float((e_yx.abs() / pyx64.abs()).max()), "abs_err_logp_nats_max": float(e_yx.abs().max()),
                             "abs_err_logp_nats_median": float(e_yx.abs().median()), "exponent_err_nats_median": float(dexp.median()),
                             "exponent_err_nats_max": float(dexp.max()),
                             "expected_flip_fraction": float((1 - torch.exp(-dexp)).mean() * 0.4)}  # P(u between p and p') ~ p |d|, p ~ 0.4
            out[weights] = res
            print(f"weights = {weights}: |log p| ~ {float(pyx64.abs().mean()):.0f} nats")
            print(f"{'mode':8s} {'rel err max':>12s} {'|err| max':>10s} {'|err| med':>10s} {'exp err med':>12s} {'exp err max':>12s} {'flips/decision':>15s}")
            for m, r in res.items():
                print(f"{m:8s} {r['rel_err_logp_max']:12.2e} {r['abs_err_logp_nats_max']:10.2e} {r['abs_err_logp_nats_median']:10.2e} "
                      f"{r['exponent_err_nats_median']:12.2e} {r['exponent_err_nats_max']:12.2e} {r['expected_flip_fraction']:15.2e}")
    with open(os.path.join(ROOT, "profiles", "r02_tf32_numerics.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
