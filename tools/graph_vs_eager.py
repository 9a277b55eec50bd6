"""MH step of the bench workload launched eagerly vs replayed as one CUDA graph (MHChains.capture_graph), with the SM clock
and power draw sampled during each timed region.  Usage: python tools/graph_vs_eager.py [chains] [steps]"""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import timewarp_b200 as tw
from timewarp_b200.energy import PeptidePotentialEnergy
from timewarp_b200.forcefield import amber99sbildn_obc2
from timewarp_b200.peptides import tetrapeptide_2olx
from timewarp_b200.sampling import MHChains

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
pep = tetrapeptide_2olx()
m = tw.custom_transformer_nvp_constructor(tw.kernel_transformer_nvp_config("bf16x3"))
m.load_state_dict(bench.bench_state_dict(m, "proposal"))
m = m.cuda().eval()
g = torch.Generator().manual_seed(0)
x = (torch.tensor(pep.coords_nm, dtype=torch.float32)[None] + 0.005 * torch.randn(B, pep.num_atoms, 3, generator=g)).cuda()
at = torch.tensor(pep.atom_types)[None].repeat(B, 1).cuda()
mask = torch.zeros(B, pep.num_atoms, dtype=torch.bool).cuda()


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows, self.stop = [], False

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.rows.append((float(out[0]), float(out[1])))
            except Exception:
                pass
            time.sleep(0.05)


def timed(fn, label):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s = Sampler()
    s.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    host_ms = (time.perf_counter() - t0) * 1e3 / steps  # time the host needs to ENQUEUE a step
    torch.cuda.synchronize()
    s.stop = True
    s.join()
    clk = sorted(r[0] for r in s.rows) or [0.0]
    pw = sorted(r[1] for r in s.rows) or [0.0]
    print(f"{label:28s} {e0.elapsed_time(e1) / steps:8.3f} ms per step   host enqueue {host_ms:7.3f} ms   SM clock median {clk[len(clk) // 2]:6.0f} MHz   "
          f"power median {pw[len(pw) // 2]:6.0f} W   ({len(s.rows)} samples)")


chains = MHChains(m, PeptidePotentialEnergy(amber99sbildn_obc2(pep)), at, mask, x)
timed(chains.step, "eager launches")
chains.capture_graph()
timed(chains.step, "one CUDA graph per step")
chains.release_graph()
timed(chains.step, "eager again")
