"""MH steps of B chains of 2olx (the bench workload) for ncu: `warm` untimed steps (weights are packed in the
first), then `steps` more.  tools/summarize_launches.py --skip-pack reports the shares of the step kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timewarp_b200 as tw
from timewarp_b200.synthetic import synth_state_dict
from timewarp_b200.energy import PeptidePotentialEnergy
from timewarp_b200.forcefield import amber_like_system
from timewarp_b200.peptides import tetrapeptide_2olx
from timewarp_b200.sampling import MHChains

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
pep = tetrapeptide_2olx()
m = tw.custom_transformer_nvp_constructor(tw.kernel_transformer_nvp_config(prec))
m.load_state_dict(synth_state_dict(m, 0))
m = m.cuda().eval()
g = torch.Generator().manual_seed(0)
x = (torch.tensor(pep.coords_nm, dtype=torch.float32)[None] + 0.005 * torch.randn(B, pep.num_atoms, 3, generator=g)).cuda()
at = torch.tensor(pep.atom_types)[None].repeat(B, 1).cuda()
mask = torch.zeros(B, pep.num_atoms, dtype=torch.bool).cuda()
chains = MHChains(m, PeptidePotentialEnergy(amber_like_system(pep)), at, mask, x)
for _ in range(steps):
    chains.step()
torch.cuda.synchronize()
print(chains.acceptance_rate().mean().item())
