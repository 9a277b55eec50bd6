"""One small `local`-attention training step (sampling pass + density pass in one graph, both backward entry points) for
`compute-sanitizer --tool memcheck`: B = 3 (ragged), V = 22."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timewarp_b200 as tw
from timewarp_b200.peptides import alanine_dipeptide
from timewarp_b200.synthetic import synth_state_dict

pep = alanine_dipeptide()
cfg = tw.kernel_transformer_nvp_config("bf16x3")
enc = cfg.encoder_layer_config
enc.attention_type, enc.lengthscales, enc.normalise_kernel_values, enc.max_radius, enc.num_heads = "local", None, None, 0.45, 6
m = tw.custom_transformer_nvp_constructor(cfg)
m.load_state_dict(synth_state_dict(m, 0))
m = m.cuda().train()
B, V = 3, pep.num_atoms
x = (torch.tensor(pep.coords_nm, dtype=torch.float32)[None] + 0.01 * torch.randn(B, V, 3)).cuda()
xv = torch.randn(B, V, 3).cuda()
at = torch.tensor(pep.atom_types)[None].repeat(B, 1).cuda()
mask = torch.zeros(B, V, dtype=torch.bool).cuda()
mask[1, 17:] = True
mask[2, 5:] = True
kw = dict(adj_list=torch.zeros(0, 2, dtype=torch.long).cuda(), edge_batch_idx=torch.zeros(0, dtype=torch.long).cuda(), masked_elements=mask)
yc, yv, lp = m.conditional_sample_with_logp(atom_types=at, x_coords=x, x_velocs=xv, num_samples=1, **kw)
ll = m.log_likelihood(atom_types=at, x_coords=yc[0], x_velocs=yv[0], y_coords=x, y_velocs=xv, **kw)
(lp[0] - ll + (yc[0] ** 2).sum((-1, -2))).mean().backward()
torch.cuda.synchronize()
print("memcheck local-attention train run done", float(lp.sum()), sum(int(p.grad is not None) for p in m.parameters()))
