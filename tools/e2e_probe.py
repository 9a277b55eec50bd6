"""Where the end-to-end overhead of bench.py comes from: the MH step with host buffers in / out, replayed as a CUDA graph or launched
eagerly, and the pieces of the host round trip on their own (1024 chains x 65 atoms, bf16x3)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import timewarp_b200 as tw
from timewarp_b200.energy import PeptidePotentialEnergy
from timewarp_b200.peptides import tetrapeptide_2olx
from timewarp_b200.sampling import MHChains

dev = torch.device("cuda", 0)
pep = tetrapeptide_2olx()
model = tw.custom_transformer_nvp_constructor(tw.kernel_transformer_nvp_config("bf16x3"))
model.load_state_dict(bench.bench_state_dict(model, "synthetic"))
model = model.to(dev).eval()
sysd, _ = bench.bench_system(pep)
energy = PeptidePotentialEnergy(sysd)
x0, at, mask = bench.synthetic_chains(pep, 1024, seed=1000)
torch.manual_seed(0)
chains = MHChains(model, energy, at.to(dev), mask.to(dev), x0.to(dev))
for _ in range(3):
    chains.step()
chains.capture_graph(warmup=1)
chains.step()
hx = x0.clone().pin_memory(); hat, hmask = at.clone().pin_memory(), mask.clone().pin_memory()
hy = torch.empty_like(hx).pin_memory(); hacc = torch.empty(1024, dtype=torch.bool).pin_memory()


def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n


def e2e(step):
    def f():
        chains.x.copy_(hx, non_blocking=True); chains.atom_types.copy_(hat, non_blocking=True); chains.mask.copy_(hmask, non_blocking=True)
        chains.e_pot_x.copy_((energy(chains.x) / chains.kbT).squeeze(-1))
        acc = step()
        hy.copy_(chains.x, non_blocking=True); hacc.copy_(acc, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        hx.copy_(hy)
    return f


def sync_after(step):
    def f():
        step(); torch.cuda.current_stream().synchronize()
    return f


print("graph, back to back        : %.3f ms device, %.3f ms wall" % timed(chains.step))
print("eager, back to back        : %.3f ms device, %.3f ms wall" % timed(chains._step_impl))
print("graph + sync per step      : %.3f ms device, %.3f ms wall" % timed(sync_after(chains.step)))
print("eager + sync per step      : %.3f ms device, %.3f ms wall" % timed(sync_after(chains._step_impl)))
print("e2e (host in/out), graph   : %.3f ms device, %.3f ms wall" % timed(e2e(chains.step)))
print("e2e (host in/out), eager   : %.3f ms device, %.3f ms wall" % timed(e2e(chains._step_impl)))
