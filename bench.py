#!/usr/bin/env python
"""Benchmark of the Timewarp MH hot path on B200 (BASELINE.json: "MH proposals/sec").

One STEP = one Metropolis-Hastings iteration of `--chains` independent chains of the synthetic
4-residue peptide (2olx, 65 atoms): reverse flow pass (proposal + log p_xy), potential + kinetic
energies, forward flow pass (log p_yx), accept/reject  -- utils/evaluation_utils.py:589-689 with
num_proposal_steps == 1 applied to every chain (BASELINE.json configs[2]).  Proposals/s = chains * steps / time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by the driver with torch.distributed.run (one rank per GPU, NCCL); chains are
sharded (weak scaling: --chains per GPU), no data-path collective, one all-gather of acceptance
statistics after the timed region.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "mh_proposals_per_sec"
UNIT = "proposals/s"


def f_atom(V):  # algorithmic FLOPs per atom per flow pass
    return 71_663_616 + 73_728 * V


def ffn_flops_per_token(D=128, F=2048):
    return 2 * (D * F + F * D)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu_index, [], threading.Event()
        # NVML in-process (the library behind nvidia-smi): a query costs microseconds.  Spawning nvidia-smi five times a second
        # stalls work submission for milliseconds each time, which a region with one host round trip per step (e2e) pays in full.
        self.source, self._nvml, self._h = "nvidia-smi", None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml, self._h, self.source = pynvml, pynvml.nvmlDeviceGetHandleByIndex(gpu_index), "nvml"
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n, h = self._nvml, self._h
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        try:
            bits = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            bits = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        try:
            power = n.nvmlDeviceGetPowerUsage(h) / 1000.0
        except Exception:
            power = float("nan")
        flag = lambda m: "Active" if bits & m else "Not Active"  # noqa: E731
        # nvml.h: SwPowerCap 0x4, HwSlowdown 0x8, SwThermalSlowdown 0x20, HwThermalSlowdown 0x40
        return [str(self.gpu), str(sm), str(mx), f"{power:.1f}", hex(bits), flag(0x8), flag(0x40), flag(0x20), flag(0x4)]

    def run(self):
        while not self.stop_flag.is_set():
            if self._nvml is not None:
                try:
                    self.rows.append(self._sample_nvml())
                    self.stop_flag.wait(0.2)
                    continue
                except Exception:
                    self._nvml, self.source = None, "nvidia-smi"
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "source": self.source}


def synthetic_chains(pep, n_chains, seed):
    """Chain starts: 2olx MD frame jittered with N(0, 0.005^2) nm."""
    g = torch.Generator().manual_seed(seed)
    x = torch.tensor(pep.coords_nm, dtype=torch.float32)[None] + 0.005 * torch.randn(n_chains, pep.num_atoms, 3, generator=g)
    at = torch.tensor(pep.atom_types)[None].repeat(n_chains, 1)
    mask = torch.zeros(n_chains, pep.num_atoms, dtype=torch.bool)
    return x, at, mask


def workload_config(args, energy_name):
    """`config` of the JSON line: describes the WORKLOAD only, so that both arms (ours / reference) print the same dict."""
    V = 65
    return {"workload": f"mh_2olx65_chains{args.chains}_per_gpu", "atoms": V, "chains_per_gpu": args.chains,
            "model": f"kernel_transformer_nvp (35.97M params, synthetic weights: {args.weights})",
            "energy": energy_name,
            "l2": "working set per step (activations+workspace) >> 126 MB L2; no explicit flush",
            "algorithmic_tflops_per_step": args.chains * 2 * V * f_atom(V) / 1e12}


def bench_system(pep):
    """The potential-energy system of the benchmark peptide: the ff99SB-ILDN + OBC2 table pinned to the reference's golden
    energies when the product has it (forcefield.amber99sbildn_obc2), else the synthetic Amber-like parameters."""
    from timewarp_b200 import forcefield as ff

    if hasattr(ff, "amber99sbildn_obc2") and getattr(ff, "AMBER99SBILDN_PINNED", False):
        return ff.amber99sbildn_obc2(pep), "ff99SB-ILDN + GB-OBC2 (pinned to the reference's golden 2olx energies)"
    return ff.amber_like_system(pep), "synthetic Amber-like + GB-OBC2"


def bench_state_dict(src, mode):
    """Synthetic weights of the full architecture.  `src` is the product model (GPU arm: timewarp_b200.synthetic) or an
    oracle config (CPU-baseline / reference arm: oracle.flow_oracle) -- the two generators give identical tensors
    (tests/test_abi_cpu.py::test_synthetic_weights_match_oracle_generator).  mode "init": random-init-scale parameters (proposals of a random
    flow are never accepted).  mode "proposal" (default): the same tensors with the last layer of every out_mlp scaled by
    1e-5 (shifts far below the proposal width) and the prior scales set to (5e-4 nm, 1): a near-identity flow whose
    proposals are local moves, so the accept
    branch of the MH rule is exercised.  Identical arithmetic and shapes either way."""
    if isinstance(src, torch.nn.Module):
        from timewarp_b200.synthetic import synth_state_dict

        sd = synth_state_dict(src, 0)
        n_hidden = len(src.flow.chain[0].scale_transformer.out_mlp.linears()) - 1
    else:  # CPU legs only: the oracle is the thing being timed there
        from oracle import flow_oracle as fo

        sd = fo.synth_state_dict(src, 0)
        n_hidden = len(src.latent_mlp_hidden_dims)
    if mode == "proposal":
        last = 2 * n_hidden
        for k in list(sd):
            if f".out_mlp._layers.{last}." in k:
                sd[k] = sd[k] * 1e-5
        sd["coords_prior_log_scale"] = torch.tensor(float(np.log(5e-4)))
        sd["velocs_prior_log_scale"] = torch.tensor(0.0)
    return sd


# --------------------------------------------------------------------------------------------
def cpu_reference_iteration(flow, sd, o, sysd, kbT, x, at, mask, gen):
    """One MH iteration on the CPU, evaluation_utils.py:589-668 with one proposal per chain.  `flow` is the UNMODIFIED
    reference model (oracle/_ref, byte-compiled from /root/reference by oracle/build_ref.py) driven through its public
    `conditional_sample_with_logp` / `log_likelihood` (the default torch generator supplies its latents), or None: the
    oracle port of the same arithmetic (oracle/flow_oracle.py).  Energies: the fp64 numpy port (OpenMM is not installable
    offline)."""
    from oracle import energy_oracle as eo
    from oracle import flow_oracle as fo

    n = x.shape[0]
    xv = torch.randn(x.shape, generator=gen)
    with torch.no_grad():
        if flow is not None:
            kw = dict(atom_types=at, adj_list=torch.zeros(0, 2, dtype=torch.long), edge_batch_idx=torch.zeros(0, dtype=torch.long),
                      masked_elements=mask)
            yc, yv, p_xy = flow.conditional_sample_with_logp(x_coords=x, x_velocs=xv, num_samples=1, **kw)
            p_yx = flow.log_likelihood(x_coords=yc[0], x_velocs=yv[0], y_coords=x, y_velocs=xv, **kw)
        else:
            zc = torch.randn((1,) + tuple(x.shape), generator=gen) * torch.exp(sd["coords_prior_log_scale"])
            zv = torch.randn((1,) + tuple(x.shape), generator=gen) * torch.exp(sd["velocs_prior_log_scale"])
            yc, yv, p_xy = fo.conditional_sample_with_logp(sd, o, at, x, xv, mask, 1, zc, zv)
            p_yx = fo.log_likelihood(sd, o, at, yc[0], yv[0], x, xv, mask)
    e_x = torch.from_numpy(eo.potential_energy(sysd, x.numpy().astype(np.float64))).float() / kbT
    e_y = torch.from_numpy(eo.potential_energy(sysd, yc[0].numpy().astype(np.float64))).float() / kbT
    e_kin = 0.5 * (yv[0] ** 2).sum((-1, -2)) - 0.5 * (xv**2).sum((-1, -2))
    ex = (e_y - e_x) + e_kin + p_xy[0] - p_yx
    acc = torch.rand(n, generator=gen) < torch.clamp(torch.exp(-ex), max=1.0)
    return torch.where(acc[:, None, None], yc[0], x), acc


def time_cpu_reference(sample_chains, iters, warmup, seed=0, weights="proposal"):
    """Times the CPU path on all host threads.  Returns a dict: value (proposals/s), cores, ms per iteration, kind, sample."""
    from oracle import flow_oracle as fo
    from oracle import ref_flow
    from timewarp_b200.forcefield import MOLAR_GAS_CONSTANT_R
    from timewarp_b200.peptides import tetrapeptide_2olx

    torch.set_num_threads(os.cpu_count() or 1)
    pep = tetrapeptide_2olx()
    o = fo.OracleConfig()
    sd = bench_state_dict(o, weights)
    flow, kind = None, "port"
    if ref_flow.available():
        try:
            flow = ref_flow.full_model(sd)
            kind = "reference(flow)+port(energy)"
        except Exception as e:  # pragma: no cover
            sys.stderr.write(f"bench: oracle/_ref not usable ({e}); timing the oracle port\n")
    sysd, energy_name = bench_system(pep)
    sysd = sysd.as_float32()
    kbT = 310.0 * MOLAR_GAS_CONSTANT_R
    x, at, mask = synthetic_chains(pep, sample_chains, seed)
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    acc_n = 0
    for _ in range(warmup):
        x, _ = cpu_reference_iteration(flow, sd, o, sysd, kbT, x, at, mask, gen)
    times = []
    for _ in range(iters):
        t0 = time.perf_counter()
        x, acc = cpu_reference_iteration(flow, sd, o, sysd, kbT, x, at, mask, gen)
        times.append(time.perf_counter() - t0)
        acc_n += int(acc.sum())
    flow_name = ("unmodified reference modules (oracle/_ref, byte-compiled from /root/reference): ConditionalFlowDensityModel."
                 "conditional_sample_with_logp + .log_likelihood, torch-CPU fp32") if flow is not None else "oracle/flow_oracle.py (torch-CPU fp32)"
    return {"value": sample_chains * iters / sum(times), "cores": torch.get_num_threads(), "ms": 1e3 * sum(times) / iters, "kind": kind,
            "acceptance_rate": acc_n / max(sample_chains * iters, 1), "energy": energy_name,
            "sample": f"{sample_chains} chains x {iters} MH iterations (+{warmup} warm-up) of the same 2olx-65 workload: flow 2 passes through "
                      f"{flow_name} + 2 energies each through oracle/energy_oracle.py (numpy fp64, {energy_name})"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S = args.cpu_sample
    r = time_cpu_reference(S, args.steps, args.warmup, weights=args.weights)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, r["energy"]),
        "kernel_config": {"chains_timed_per_step": S,
                          "note": "the reference's CPU path for this workload: proposals/s of one CPU process is insensitive to the number "
                                  "of chains per call (SURVEY.md section 6: 20-22 /s at 64...256 chains on 8 cores), so each step is a "
                                  f"bounded sample of {S} chains; OpenMM is not installable offline, so the energies come from the fp64 "
                                  "numpy port"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "acceptance_rate_mean": r["acceptance_rate"],
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def time_nll_training(dev, precision, batch=256, steps=5, warmup=3, use_graph=True):
    """BASELINE.json configs[1]: alanine-dipeptide (22 atoms) NLL training, batch 256 on one GPU: forward (taped) +
    hand-written backward + Adam step.  Returns atoms/s (B * V / step time, CUDA events)."""
    import timewarp_b200 as tw
    from timewarp_b200.peptides import alanine_dipeptide
    from timewarp_b200.synthetic import synth_state_dict

    pep = alanine_dipeptide()
    V = pep.num_atoms
    model = tw.custom_transformer_nvp_constructor(tw.kernel_transformer_nvp_config(precision))
    model.load_state_dict(synth_state_dict(model, 0))
    model = model.to(dev).train()
    from timewarp_b200.optim import FlatAdam

    opt = FlatAdam(model, lr=1e-4)  # torch.optim.Adam's update rule as one launch over the flat gradient buffer
    g = torch.Generator().manual_seed(0)
    x = torch.tensor(pep.coords_nm, dtype=torch.float32)[None] + 0.01 * torch.randn(batch, V, 3, generator=g)
    y = x + 0.02 * torch.randn(batch, V, 3, generator=g)
    kw = dict(atom_types=torch.tensor(pep.atom_types)[None].repeat(batch, 1).to(dev), x_coords=x.to(dev),
              x_velocs=torch.randn(batch, V, 3, generator=g).to(dev), y_coords=y.to(dev), y_velocs=torch.randn(batch, V, 3, generator=g).to(dev),
              adj_list=torch.zeros(0, 2, dtype=torch.long, device=dev), edge_batch_idx=torch.zeros(0, dtype=torch.long, device=dev),
              masked_elements=torch.zeros(batch, V, dtype=torch.bool, device=dev))

    def step():
        opt.zero_grad(set_to_none=True)  # the backward hands out views of one flat, freshly zeroed gradient buffer
        loss = model(**kw)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    # The step is launch-bound at this size (5632 tokens, ~1500 kernel launches): replay it as ONE CUDA graph
    # (forward with tape + hand-written backward + Adam).  Falls back to eager launches if capture fails.
    launch = "eager"
    graph_step = None
    if use_graph:
        try:
            static_loss = None
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_loss = step()
            graph_step = lambda: (g.replay(), static_loss)[1]  # noqa: E731
            for _ in range(2):
                graph_step()
            torch.cuda.synchronize()
            launch = "one CUDA graph per training step"
        except Exception as e:  # pragma: no cover
            graph_step = None
            launch = f"eager (graph capture failed: {str(e)[:80]})"
            torch.cuda.synchronize()
    run = graph_step or step
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        loss = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"metric": "nll_train_atoms_per_sec", "value": batch * V / (ms / 1e3), "unit": "atoms/s", "ms_per_step": ms,
            "config": {"workload": f"nll_train_ad22_batch{batch}", "atoms": V, "batch": batch, "optimizer": "Adam (optim.FlatAdam, one launch)", "precision": precision,
                       "launch": launch},
            "final_loss": float(loss.detach())}


def time_reference_s_mode(model, energy, pep, dev, S, num_samples=96):
    """`sample_with_model` exactly as evaluate.py --mh drives it (utils/evaluation_utils.py:468-745): one chain, S parallel
    proposals from the current state per iteration, the chain advances to the first accepted one (one 4-byte host read per
    iteration).  Throughput = proposals evaluated per second (iterations x S / wall time incl. the host bookkeeping)."""
    from timewarp_b200 import sampling

    class _Batch:
        atom_coords = torch.tensor(pep.coords_nm, dtype=torch.float32)[None]
        atom_velocs = torch.zeros(1, pep.num_atoms, 3)
        atom_types = torch.tensor(pep.atom_types)[None]
        masked_elements = torch.zeros(1, pep.num_atoms, dtype=torch.bool)
        adj_list = torch.tensor(pep.bonds)
        edge_batch_idx = torch.zeros(len(pep.bonds), dtype=torch.long)

    calls = [0]
    inner = model.conditional_sample_with_logp

    def counted(*a, **kw):
        calls[0] += 1
        return inner(*a, **kw)

    model.conditional_sample_with_logp = counted
    try:
        masses = torch.tensor(pep.masses, dtype=torch.float32)
        kw = dict(accept=True, random_velocs=True, resample_velocs=True, num_proposal_steps=S)
        sampling.sample_with_model(_Batch(), model, dev, energy, masses, 4, **kw)  # warm-up (workspace for S samples)
        torch.cuda.synchronize()
        calls[0] = 0
        t0 = time.perf_counter()
        coords, _, accepted, stats = sampling.sample_with_model(_Batch(), model, dev, energy, masses, num_samples, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    finally:
        del model.conditional_sample_with_logp
    return {"metric": "mh_proposals_per_sec", "value": calls[0] * S / dt, "unit": "proposals/s", "iterations": calls[0],
            "proposals_per_iteration": S, "chain_states": int(len(stats)), "accepted": int(accepted), "ms_per_iteration": dt / calls[0] * 1e3,
            "note": "sample_with_model, 1 chain x S proposals per iteration (shared conditioning), host bookkeeping and per-iteration sync included"}


def measure_nll_dp(dev, rank, world, prec, B, steps, warmup):
    """Data-parallel NLL training (BASELINE.json configs[3]: dipeptide set, batch 2048 over 8 GPUs = 256 per GPU, ONE
    gradient all-reduce per step -- train_deepspeed.py:99-120,186-188).  Synthetic 2AA-like ragged batches: atom counts
    uniform in [17, 51], padded to the batch maximum with `masked_elements`.  Every rank calls this; returns the result
    dict (atoms/s = un-padded atoms of all ranks / max-over-ranks device time; `allreduce_ms` = device time of the flat
    gradient all-reduce inside the step, CUDA events)."""
    import timewarp_b200 as tw
    from timewarp_b200 import distributed as twd
    from timewarp_b200.synthetic import synth_state_dict

    model = tw.custom_transformer_nvp_constructor(tw.kernel_transformer_nvp_config(prec))
    model.load_state_dict(synth_state_dict(model, 0))
    model = model.to(dev).train()
    twd.broadcast_parameters(model.parameters())
    from timewarp_b200.optim import FlatAdam

    opt = FlatAdam(model, lr=1e-4)
    trainer = twd.DataParallelTrainer(model, opt)
    g = torch.Generator().manual_seed(100 + rank)
    lengths = torch.randint(17, 52, (B,), generator=g)
    V = int(lengths.max())
    mask = torch.arange(V)[None, :] >= lengths[:, None]
    keep = (~mask)[:, :, None]
    x = 0.3 * torch.randn(B, V, 3, generator=g) * keep
    y = (x + 0.02 * torch.randn(B, V, 3, generator=g)) * keep
    batch = dict(atom_types=(torch.randint(0, 5, (B, V), generator=g) * (~mask)).to(dev), x_coords=x.to(dev),
                 x_velocs=(torch.randn(B, V, 3, generator=g) * keep).to(dev), y_coords=y.to(dev),
                 y_velocs=(torch.randn(B, V, 3, generator=g) * keep).to(dev), adj_list=torch.zeros(0, 2, dtype=torch.long, device=dev),
                 edge_batch_idx=torch.zeros(0, dtype=torch.long, device=dev), masked_elements=mask.to(dev))
    atoms = torch.tensor([float(lengths.sum())], device=dev, dtype=torch.float64)
    for _ in range(warmup):
        trainer.step(batch)
    trainer.time_collective = True
    trainer.collective_ms()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = trainer.step(batch)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), trainer.collective_ms() * steps], device=dev, dtype=torch.float64)
    # replicas must stay bit-identical: compare a parameter checksum across ranks
    chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(atoms)
        lo, hi = chk.clone(), chk.clone()
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
        torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
        in_sync = bool((lo == hi).item())
    else:
        in_sync = True
    ms, ar_ms = float(t[0].item()) / steps, float(t[1].item()) / steps
    nbytes = int(trainer.collective_bytes)
    res = {"metric": "nll_train_atoms_per_sec", "value": float(atoms.item()) / (ms / 1e3), "unit": "atoms/s", "n_gpus": world,
           "steps": steps, "warmup": warmup, "step_ms": ms, "ms_per_step": ms, "allreduce_ms": ar_ms, "bytes": nbytes,
           "allreduce_busbw_gbs": (nbytes * 2 * (world - 1) / world / (ar_ms / 1e3) / 1e9) if (world > 1 and ar_ms > 0) else None,
           "allreduce_share_of_step": ar_ms / ms if ms > 0 else None, "higher_is_better": True, "scaling": "weak", "dtype": prec,
           "config": {"workload": f"nll_train_2aa_like_batch{B}_per_gpu", "batch_per_gpu": B, "global_batch": B * world, "padded_atoms": V,
                      "atoms_per_step_all_ranks": float(atoms.item()), "optimizer": "Adam (optim.FlatAdam, one launch)", "precision": prec,
                      "collective": "one NCCL all-reduce over the flat fp32 gradient buffer per step, in place"},
           "final_loss": float(loss.detach()), "replicas_in_sync": in_sync}
    del trainer, opt, model
    torch.cuda.empty_cache()
    return res


def run_nll(args):
    """`--workload nll`: the data-parallel NLL training workload alone.  Prints ONE JSON line."""
    from timewarp_b200 import distributed as twd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    twd.init_from_env("nccl", dev)
    res = measure_nll_dp(dev, rank, world, args.precision, args.batch, args.steps, args.warmup)
    if rank == 0:
        res.update({"vs_baseline": None, "data": "synthetic"})
        print(json.dumps(res), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def run_ours(args):
    import ctypes as C

    import timewarp_b200 as tw
    from timewarp_b200 import _lib
    from timewarp_b200.energy import PeptidePotentialEnergy
    from timewarp_b200.peptides import tetrapeptide_2olx
    from timewarp_b200.sampling import MHChains

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    lib = _lib.load()
    pep = tetrapeptide_2olx()
    V = pep.num_atoms
    model = tw.custom_transformer_nvp_constructor(tw.kernel_transformer_nvp_config(args.precision))
    model.load_state_dict(bench_state_dict(model, args.weights))
    model = model.to(dev).eval()
    sysd, energy_name = bench_system(pep)
    energy = PeptidePotentialEnergy(sysd)
    x0, at, mask = synthetic_chains(pep, args.chains, seed=1000 + rank)
    torch.manual_seed(args.seed + rank)  # per-rank generator
    chains = MHChains(model, energy, at.to(dev), mask.to(dev), x0.to(dev))

    for _ in range(args.warmup):
        chains.step()
    if args.graph:
        chains.capture_graph(warmup=1)  # one MH iteration = ONE CUDA graph replay (no per-kernel host launches)
        chains.step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- device-resident timed region: exactly K steps -------------------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        chains.step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)

    # ---- the same iteration launched eagerly, with CUDA events around every fused-FFN launch (dominant kernel):
    # kernel count per step and the FFN launch time for the roofline.  Outside the timed region: graph replays have no
    # host-side launches to bracket.
    n_prof = 3
    lib.tw_prof_enable(1)
    launches0 = lib.tw_debug_launch_count()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(n_prof):
        chains._step_impl()
    pe1.record()
    torch.cuda.synchronize()
    ms_prof = pe0.elapsed_time(pe1)
    launches_per_step = (lib.tw_debug_launch_count() - launches0) // n_prof
    launches = launches_per_step * args.steps
    ffn_ms, ffn_scopes = C.c_double(0), C.c_longlong(0)
    _lib.check(lib.tw_prof_collect(C.byref(ffn_ms), C.byref(ffn_scopes)), "tw_prof_collect")
    # ... and the same for the attention layer (the "kernel-attention tile" of the north star), one more eager step
    lib.tw_prof_enable(2)
    chains._step_impl()
    attn_ms, attn_scopes = C.c_double(0), C.c_longlong(0)
    _lib.check(lib.tw_prof_collect(C.byref(attn_ms), C.byref(attn_scopes)), "tw_prof_collect")
    lib.tw_prof_enable(0)

    # ---- end-to-end region: host (pinned) state in, host state + decisions out, every step --
    hx = x0.clone().pin_memory()
    hat, hmask = at.clone().pin_memory(), mask.clone().pin_memory()
    hy = torch.empty_like(hx).pin_memory()
    hacc = torch.empty(args.chains, dtype=torch.bool).pin_memory()
    h2d = hx.numel() * 4 + hat.numel() * 8 + hmask.numel()
    d2h = hy.numel() * 4 + hacc.numel()

    def e2e_step():
        chains.x.copy_(hx, non_blocking=True)
        chains.atom_types.copy_(hat, non_blocking=True)
        chains.mask.copy_(hmask, non_blocking=True)
        chains.e_pot_x.copy_((energy(chains.x) / chains.kbT).squeeze(-1))  # host-fed state: its energy is part of the call
        acc = chains.step()
        hy.copy_(chains.x, non_blocking=True)
        hacc.copy_(acc, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads the result
        hx.copy_(hy)

    e2e_step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=3)

    t = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
    acc_rate = chains.acceptance_rate()
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
        gathered = [torch.empty_like(acc_rate) for _ in range(world)]
        dist.all_gather(gathered, acc_rate)  # the path's only collective: acceptance statistics
        acc_rate = torch.cat(gathered)
    ms_total, ms_e2e = t.tolist()
    total_chains = args.chains * world
    # ---- secondary: the workload that really communicates (BASELINE.json configs[3]), on every rank, after the MH numbers
    nll_dp = None
    if not args.no_nll:
        nll_dp = {}
        for prec in ("bf16", "bf16x3"):  # configs[3] says bf16; bf16x3 is the parity-grade default of this repo
            try:
                nll_dp[prec] = measure_nll_dp(dev, rank, world, prec, args.batch, 5, 3)
            except Exception as e:  # the MH line must not depend on the training path
                nll_dp[prec] = {"metric": "nll_train_atoms_per_sec", "error": str(e)[:200]}
    value = total_chains * args.steps / (ms_total / 1e3)
    e2e_value = total_chains * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        peaks = load_peaks()
        M = args.chains * V
        n_ffn = max(int(ffn_scopes.value), 1)
        ffn_ms_avg = ffn_ms.value / n_ffn
        ffn_flops = 2 * M * ffn_flops_per_token()  # one timed scope = FFN of BOTH conditioner networks over all tokens
        achieved = ffn_flops / (ffn_ms_avg / 1e3) / 1e12 if ffn_ms_avg > 0 else 0.0
        issued_factor = 3 if args.precision == "bf16x3" else 1
        # attention layer (both networks per timed scope).  Algorithmic = the reference's arithmetic per token: value projection
        # 196 608 + mixing 1 536 V + out projection 196 608.  Issued = what the fused kernel really feeds
        # the tensor pipe: per (sample, head) MMA1 128 x 128 x VP and MMA2 128 x 128 x 128, every row of the 128-row tile.
        n_attn = max(int(attn_scopes.value), 1)
        attn_ms_avg = attn_ms.value / n_attn
        attn_alg = 2 * M * (393216 + 1536 * V)
        VP = (V + 15) // 16 * 16
        H = int(model._cfg.num_heads)
        # issued by the feature-major kernels (csrc/attn_fm3.cu for VP <= 80, csrc/attn_fm.cu above): per (group of G samples, head)
        # the projection as 8 K-steps of M128 x N x K16 with N = G VP tokens, per (sample, head) the mixing as VP/16 K-steps of
        # M128 x VP x K16; x3 for the split
        G = max(1, (80 if VP <= 80 else 160) // VP)
        groups = (args.chains + G - 1) // G
        attn_issued = 2 * issued_factor * H * (groups * 8 * (2 * 128 * G * VP * 16) + args.chains * (VP // 16) * (2 * 128 * VP * 16))
        attn_roofline = {"bound": "tensor (in-kernel trace: 2.7 k cycles per (sample, head) for 1.88 k cycles of tensor work; the rest is issue and hand-over latency around N = 80 MMAs, DESIGN.md section 5)",
                         "kernel": "fused attention layer, feature-major deep pipeline k_attn_fm3 (W_o W_v projection of every head, per-sample mixing, residual + LayerNorm; W_c multicast over CTA pairs), both conditioner nets",
                         "achieved": attn_alg / (attn_ms_avg / 1e3) / 1e12 if attn_ms_avg > 0 else 0.0, "peak": peaks["tf_sustained"],
                         "unit": "TFLOP/s", "issued_tflops": attn_issued / (attn_ms_avg / 1e3) / 1e12 if attn_ms_avg > 0 else 0.0,
                         "avg_launch_ms": attn_ms_avg, "launches_timed": n_attn}
        attn_roofline["frac"] = attn_roofline["achieved"] / peaks["tf_sustained"]
        attn_roofline["issued_frac"] = attn_roofline["issued_tflops"] / peaks["tf_sustained"]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ffn_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(args.precision)
            except Exception:
                traffic = None
        step_flops = args.chains * 2 * V * f_atom(V)
        # the CPU baseline is a reported side figure: rank 0 at N = 1 only (the reference arm carries it at every N)
        run_cpu = not args.no_cpu_baseline and world == 1
        cpu = time_cpu_reference(args.cpu_sample, 8, 2, weights=args.weights) if run_cpu else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "bf16x3": "bf16x3(f32-accumulate)", "bf16": "bf16"}[args.precision], "data": "synthetic",
            "config": workload_config(args, energy_name),
            "kernel_config": {"precision": args.precision, "energy_kernel": "on-GPU fp64",
                              "launch": "one CUDA graph replay per MH step" if args.graph else "eager (host launches hidden behind the kernels)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor", "kernel": "fused FFN (linear1+ReLU+linear2+residual), both conditioner nets",
                         "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sustained"],
                         "peak_source": peaks["source"] + " (sustained bf16 cuBLAS)", "issued_mma_factor": issued_factor, "issued_frac": issued_factor * achieved / peaks["tf_sustained"],
                         "avg_launch_ms": ffn_ms_avg, "launches_timed": n_ffn, "share_of_step": ffn_ms.value / ms_prof,
                         "timed_in": f"{n_prof} eagerly launched steps after the timed region (CUDA events around each launch)", "traffic": traffic},
            "roofline_attention": attn_roofline,
            "whole_step_algorithmic_tflops": step_flops * args.steps / (ms_total / 1e3) / 1e12,
            "acceptance_rate_mean": float(acc_rate.mean().item()),
        }
        if world == 1 and not args.no_nll:
            try:  # the reference's own driver shape: ONE chain, S = chains proposals per iteration
                line["reference_s_mode"] = time_reference_s_mode(model, energy, pep, dev, args.chains)
            except Exception as e:
                line["reference_s_mode"] = {"error": str(e)[:200]}
        if nll_dp is not None:
            line["secondary"] = dict(nll_dp["bf16"])
            line["secondary"]["bf16x3"] = nll_dp["bf16x3"]
        if world == 1 and not args.no_nll:
            try:  # BASELINE.json configs[1]: AD-22 batch 256 on one GPU, step replayed as one CUDA graph
                line["secondary_ad22"] = time_nll_training(dev, args.precision)
            except Exception as e:
                line["secondary_ad22"] = {"metric": "nll_train_atoms_per_sec", "error": str(e)[:200]}
        if cpu is not None:
            line["cpu_baseline"] = {"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "kind": cpu["kind"], "sample": cpu["sample"]}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=1024, help="chains per GPU (BASELINE configs[2]: 1024)")
    ap.add_argument("--precision", default=os.environ.get("TW_PRECISION", "bf16x3"), choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--cpu-sample", type=int, default=64, help="chains in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--weights", default="proposal", choices=["proposal", "init"], help="synthetic weight set (see bench_state_dict)")
    ap.add_argument("--graph", action=argparse.BooleanOptionalAction, default=True,
                    help="replay one CUDA graph per MH step instead of launching every kernel from the host: same GPU time (25.60 vs 25.75 "
                    "ms per step, tools/graph_vs_eager.py), but enqueueing a step eagerly costs the host 20.8 ms -- with one process per GPU "
                    "on a shared host the replay keeps the step device-bound (--no-graph: eager launches)")
    ap.add_argument("--workload", default="mh", choices=["mh", "nll"], help="mh: the headline MH benchmark; nll: data-parallel NLL training")
    ap.add_argument("--batch", type=int, default=256, help="--workload nll: samples per GPU")
    ap.add_argument("--no-nll", action="store_true", help="skip the secondary NLL-training throughput measurement")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "nll":
        run_nll(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
