import sys; sys.path.insert(0, '/root/repo')
import torch, copy
import timewarp_b200 as tw
from oracle import flow_oracle as fo
from timewarp_b200.peptides import alanine_dipeptide
dev = torch.device('cuda')
pep = alanine_dipeptide(); V = pep.num_atoms; batch = 64
g = torch.Generator().manual_seed(0)
x = torch.tensor(pep.coords_nm, dtype=torch.float32)[None] + 0.01 * torch.randn(batch, V, 3, generator=g)
y = x + 0.02 * torch.randn(batch, V, 3, generator=g)
kw = dict(atom_types=torch.tensor(pep.atom_types)[None].repeat(batch, 1).to(dev), x_coords=x.to(dev),
          x_velocs=torch.randn(batch, V, 3, generator=g).to(dev), y_coords=y.to(dev), y_velocs=torch.randn(batch, V, 3, generator=g).to(dev),
          adj_list=torch.zeros(0, 2, dtype=torch.long, device=dev), edge_batch_idx=torch.zeros(0, dtype=torch.long, device=dev),
          masked_elements=torch.zeros(batch, V, dtype=torch.bool, device=dev))
for name, okw, stn in (("foreach", dict(capturable=True), False), ("foreach-none", dict(capturable=True), True), ("fused", dict(fused=True, capturable=True), True), ("plain", dict(), True)):
    model = tw.custom_transformer_nvp_constructor(tw.kernel_transformer_nvp_config("bf16x3"))
    model.load_state_dict(fo.synth_state_dict(fo.OracleConfig(), 0))
    model = model.to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, **okw)
    losses = []
    for it in range(6):
        opt.zero_grad(set_to_none=stn)
        loss = model(**kw); loss.backward(); opt.step()
        losses.append(round(float(loss.detach()), 4))
    print(name, losses)
