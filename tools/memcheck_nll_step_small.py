"""Two NLL training steps for `compute-sanitizer --tool memcheck`: kernel attention, 40 ragged samples of up to 22 atoms (880 tokens =
7 token tiles: the CTAs of the fused FFN-backward kernel cross token tiles), hand-written backward, optim.FlatAdam (parameters re-homed
into the flat buffer on the first step, weight images re-packed from the new storage on the second)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timewarp_b200 as tw
from timewarp_b200.optim import FlatAdam
from timewarp_b200.synthetic import synth_state_dict

torch.manual_seed(0)
m = tw.custom_transformer_nvp_constructor(tw.kernel_transformer_nvp_config(sys.argv[1] if len(sys.argv) > 1 else "bf16x3"))
m.load_state_dict(synth_state_dict(m, 0))
m = m.cuda().train()
opt = FlatAdam(m, lr=1e-4, weight_decay=1e-3)
B, V = 40, 22
lengths = torch.randint(9, V + 1, (B,))
lengths[0] = V
mask = torch.arange(V)[None, :] >= lengths[:, None]
keep = (~mask)[:, :, None]
x = 0.3 * torch.randn(B, V, 3) * keep
y = (x + 0.02 * torch.randn(B, V, 3)) * keep
kw = dict(atom_types=(torch.randint(0, 5, (B, V)) * (~mask)).cuda(), x_coords=x.cuda(), x_velocs=(torch.randn(B, V, 3) * keep).cuda(),
          y_coords=y.cuda(), y_velocs=(torch.randn(B, V, 3) * keep).cuda(), adj_list=torch.zeros(0, 2, dtype=torch.long).cuda(),
          edge_batch_idx=torch.zeros(0, dtype=torch.long).cuda(), masked_elements=mask.cuda())
losses = []
for _ in range(2):
    opt.zero_grad(set_to_none=True)
    loss = m(**kw)
    loss.backward()
    opt.step()
    losses.append(float(loss))
torch.cuda.synchronize()
print("memcheck nll steps done", losses)
