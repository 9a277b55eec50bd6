"""One flow pass (log_likelihood, B chains of 2olx) for ncu: warm-up pass, then the profiled pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timewarp_b200 as tw
from oracle import flow_oracle as fo
from timewarp_b200.peptides import tetrapeptide_2olx

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
pep = tetrapeptide_2olx()
m = tw.custom_transformer_nvp_constructor(tw.kernel_transformer_nvp_config(prec))
m.load_state_dict(fo.synth_state_dict(fo.OracleConfig(), 0))
m = m.cuda().eval()
g = torch.Generator().manual_seed(0)
x = (torch.tensor(pep.coords_nm, dtype=torch.float32)[None] + 0.01 * torch.randn(B, 65, 3, generator=g)).cuda()
y = x + 0.02 * torch.randn(B, 65, 3, generator=g).cuda()
xv, yv = torch.randn(B, 65, 3, generator=g).cuda(), torch.randn(B, 65, 3, generator=g).cuda()
at = torch.tensor(pep.atom_types)[None].repeat(B, 1).cuda()
mask = torch.zeros(B, 65, dtype=torch.bool).cuda()
e = torch.zeros(0, 2, dtype=torch.long).cuda()
inference = os.environ.get("TW_PROFILE_TAPED", "0") != "1"  # default: the inference path (fused kernels); 1 = taped training forward
for _ in range(passes):
  with torch.set_grad_enabled(not inference):
    ll = m.log_likelihood(atom_types=at, x_coords=x, x_velocs=xv, y_coords=y, y_velocs=yv, adj_list=e, edge_batch_idx=e[:, 0], masked_elements=mask)
  torch.cuda.synchronize()
print(ll[:4].tolist())
