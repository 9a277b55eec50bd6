"""Diagnostic: per-tensor gradient error of the CUDA backward vs the oracle's fp64 autograd (and fp32 for scale)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import flow_oracle as fo
from tests.common import FULL_O, GOLDEN, build_model, EMPTY_ADJ, EMPTY_EBI
name = sys.argv[1] if len(sys.argv) > 1 else "grads_full_ad22"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
d = np.load(os.path.join(GOLDEN, name + ".npz"))
g = {k: torch.from_numpy(d[k]) for k in d.files if d[k].dtype.kind != "U"}
m, sd = build_model(FULL_O, prec, int(g["weight_seed"]))
m.train()
kw = dict(atom_types=g["atom_types"].cuda(), x_coords=g["x_coords"].cuda(), x_velocs=g["x_velocs"].cuda(), y_coords=g["y_coords"].cuda(),
          y_velocs=g["y_velocs"].cuda(), adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=g["masked_elements"].cuda())
loss = m(**kw); loss.backward(); torch.cuda.synchronize()
ours = {k: p.grad.cpu().double() for k, p in m.named_parameters()}
args = (g["atom_types"], g["x_coords"], g["x_velocs"], g["y_coords"], g["y_velocs"], g["masked_elements"])
l32, g32 = fo.nll_loss_and_grads(sd, FULL_O, *args)
sd64 = fo.to_dtype(sd, torch.float64)
l64, g64 = fo.nll_loss_and_grads(sd64, FULL_O, args[0], *[a.double() for a in args[1:5]], args[5], distance_mode="direct")
print("loss ours", float(loss), "fp32", float(l32), "fp64", float(l64))
rows = []
for k in g64:
    n = float(g64[k].norm())
    rows.append((float((ours[k] - g64[k]).norm()) / max(n, 1e-12), float((g32[k].double() - g64[k]).norm()) / max(n, 1e-12), n, k))
rows.sort(reverse=True)
print("worst 12 (ours vs fp64, ref-fp32 vs fp64, norm, name)")
for r in rows[:12]:
    print("%.2e %.2e %.3e %s" % r)
import statistics
print("median ours", statistics.median(r[0] for r in rows), "median fp32", statistics.median(r[1] for r in rows))
