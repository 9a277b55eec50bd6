"""CUDA-event time per launch of each kernel class (tw_prof_enable: 1 FFN, 2 attention, 3 in/out MLPs, 4 energy) inside eagerly
launched MH steps of the bench workload (1024 chains x 65 atoms).  Usage: python tools/time_classes.py [chains] [steps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timewarp_b200 as tw
from timewarp_b200 import _lib
from timewarp_b200.synthetic import synth_state_dict
from timewarp_b200.energy import PeptidePotentialEnergy
from timewarp_b200.forcefield import amber_like_system
from timewarp_b200.peptides import tetrapeptide_2olx
from timewarp_b200.sampling import MHChains

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pep = tetrapeptide_2olx()
m = tw.custom_transformer_nvp_constructor(tw.kernel_transformer_nvp_config("bf16x3"))
m.load_state_dict(synth_state_dict(m, 0))
m = m.cuda().eval()
g = torch.Generator().manual_seed(0)
x = (torch.tensor(pep.coords_nm, dtype=torch.float32)[None] + 0.005 * torch.randn(B, pep.num_atoms, 3, generator=g)).cuda()
at = torch.tensor(pep.atom_types)[None].repeat(B, 1).cuda()
mask = torch.zeros(B, pep.num_atoms, dtype=torch.bool).cuda()
chains = MHChains(m, PeptidePotentialEnergy(amber_like_system(pep)), at, mask, x)
for _ in range(3):
    chains.step()
torch.cuda.synchronize()
lib = _lib.load()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    chains.step()
e1.record()
torch.cuda.synchronize()
print(f"step {e0.elapsed_time(e1) / steps:.3f} ms")
for cls, name in ((1, "ffn"), (2, "attention"), (3, "in/out mlp"), (4, "energy")):
    lib.tw_prof_enable(cls)
    for _ in range(steps):
        chains.step()
    ms, n = C.c_double(0), C.c_longlong(0)
    _lib.check(lib.tw_prof_collect(C.byref(ms), C.byref(n)), "collect")
    lib.tw_prof_enable(0)
    print(f"{name:12s} {n.value / steps:6.1f} scopes/step  {ms.value / max(n.value, 1) * 1e3:8.1f} us/scope  {ms.value / steps:7.3f} ms/step")
