"""Small invocations of the CUDA-core kernels with hand-rolled shared-memory protocols (energy + forces, Langevin steps, local
attention, fp32 flow pass, MH decision) for `compute-sanitizer --tool racecheck`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timewarp_b200 as tw
from timewarp_b200 import md, sampling
from timewarp_b200.energy import PeptidePotentialEnergy
from timewarp_b200.forcefield import amber_like_system
from timewarp_b200.peptides import alanine_dipeptide
from timewarp_b200.synthetic import synth_state_dict

pep = alanine_dipeptide()
sysd = amber_like_system(pep)
energy = PeptidePotentialEnergy(sysd)
B, V = 3, pep.num_atoms
x = (torch.tensor(pep.coords_nm, dtype=torch.float32)[None] + 0.003 * torch.randn(B, V, 3)).cuda()
e, f = energy.energy_and_forces(x)
sim = md.Simulation(sysd, md.LangevinIntegrator(310.0, 0.3, 0.0005))
x2, v2 = sim.step(x, torch.zeros_like(x), 3)
at = torch.tensor(pep.atom_types)[None].repeat(B, 1).cuda()
mask = torch.zeros(B, V, dtype=torch.bool).cuda()
for att, extra in (("kernel", dict(lengthscales=[0.3, 1.0], normalise_kernel_values=True)), ("local", dict(max_radius=0.4))):
    enc = tw.CustomAttentionEncoderLayerConfig(d_model=16, dim_feedforward=32, dropout=0.0, num_heads=2, attention_type=att, **extra)
    cfg = tw.CustomAttentionTransformerNVPConfig(atom_embedding_dim=8, latent_mlp_hidden_dims=[24], num_coupling_layers=2,
                                                 num_transformer_layers=1, encoder_layer_config=enc, precision="fp32")
    m = tw.custom_transformer_nvp_constructor(cfg)
    m.load_state_dict(synth_state_dict(m, 0))
    m = m.cuda().eval()
    with torch.no_grad():
        yc, yv, lp = m.conditional_sample_with_logp(atom_types=at, x_coords=x, x_velocs=torch.randn_like(x), adj_list=torch.zeros(0, 2, dtype=torch.long).cuda(),
                                                    edge_batch_idx=torch.zeros(0, dtype=torch.long).cuda(), masked_elements=mask, num_samples=1)
chains = sampling.MHChains(m, energy, at, mask, x)
chains.step()
torch.cuda.synchronize()
print("racecheck run done", float(e.sum()), float(lp.sum()))
