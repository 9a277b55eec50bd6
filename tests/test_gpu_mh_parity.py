"""MH drivers against the full-loop CPU oracle (oracle/mh_oracle.py) on a replayed RNG stream.

* `sample_with_model` (utils/evaluation_utils.py:468-745) chain-for-chain: S in {1, 10}, adaptive on / off.
* The bench configuration (FULL model, bf16x3, 2olx-65, proposal weights): >= 20 lock-step `MHChains.step()` iterations of
  256 chains; the decision flip set is reported (gpurun_out/mh_flip_set.json) and bounded.
* `explore` (exploration.py:229-250) step-for-step.

"MH acceptance bit-exact given fixed RNG" (north star) can only hold where |u - p_acc| exceeds the numerical error of the
exponent; these tests measure that error and assert that every decision outside that band is identical."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import flow_oracle as fo
from oracle import mh_oracle as mo
from tests.common import FULL_O, TINY_O, CudaReplayDraws, build_model, proposal_weights
from timewarp_b200 import sampling
from timewarp_b200.chirality import compute_chirality_sign, find_chirality_centers
from timewarp_b200.energy import PeptidePotentialEnergy
from timewarp_b200.forcefield import amber99sbildn_obc2, amber_like_system
from timewarp_b200.peptides import alanine_dipeptide, tetrapeptide_2olx

pytestmark = pytest.mark.gpu

EPS_P = 5e-3  # a decision may differ from the oracle's only where |u - p_acc| < EPS_P


class _Batch:
    def __init__(self, pep):
        self.atom_coords = torch.tensor(pep.coords_nm, dtype=torch.float32)[None]
        self.atom_velocs = torch.zeros_like(self.atom_coords)
        self.atom_types = torch.tensor(pep.atom_types)[None]
        self.masked_elements = torch.zeros(1, pep.num_atoms, dtype=torch.bool)
        self.adj_list = torch.tensor(pep.bonds)
        self.edge_batch_idx = torch.zeros(len(pep.bonds), dtype=torch.long)


def _model(o, precision, weights):
    m, sd = build_model(o, precision, 0)
    if weights == "proposal":
        sd = proposal_weights(sd, len(o.latent_mlp_hidden_dims))
        m.load_state_dict(sd)
    return m, sd


@pytest.mark.parametrize("S,adaptive,rotate", [(1, False, False), (10, False, False), (10, True, False), (10, False, True)])
@pytest.mark.parametrize("chirality", [False, True])
def test_sample_with_model_matches_oracle_chain(S, adaptive, chirality, rotate):
    pep = alanine_dipeptide()
    m, sd = _model(TINY_O, "fp32", "proposal")
    sysd = amber_like_system(pep)
    energy = PeptidePotentialEnergy(sysd)
    batch = _Batch(pep)
    masses = torch.tensor(pep.masses, dtype=torch.float32)
    centers = ref_signs = None
    if chirality:
        centers = find_chirality_centers(batch.adj_list, batch.atom_types)
        ref_signs = compute_chirality_sign(batch.atom_coords.cuda(), centers.cuda())
    kw = dict(accept=True, random_velocs=True, resample_velocs=True, num_proposal_steps=S, adaptive_parallelism=adaptive,
              acceptance_rate_smoothing_factor=0.3 if adaptive else 0.01, rotate=rotate)
    n = 30
    for seed in (5, 6, 7, 8):  # a seed whose first decisions all lie outside the error band (fp32 path: |u - p_acc| > 1e-4)
        torch.manual_seed(seed)
        np.random.seed(seed)  # (rotate=True draws its rotations from numpy's global generator, evaluation_utils.py:604-605)
        coords, velocs, accepted, stats = sampling.sample_with_model(batch, m, torch.device("cuda"), energy, masses, n,
                                                                     reference_signs=ref_signs, chirality_centers=centers, **kw)
        torch.manual_seed(seed)
        np.random.seed(seed)
        trace = []
        o_coords, o_velocs, o_acc, o_st = mo.sample_with_model(
            sd, TINY_O, sysd.as_float32(), energy.kbT, batch.atom_types, batch.atom_coords, batch.atom_velocs, batch.masked_elements, masses,
            n, CudaReplayDraws(), reference_signs=None if ref_signs is None else ref_signs.cpu(), chirality_centers=centers, trace=trace, **kw)
        # states up to the first decision inside the error band are comparable; afterwards the two chains may legitimately differ
        k_ok = 0
        for t in trace:
            if (np.abs(t["u"] - t["p_acc"]) < 1e-4).any():
                break
            k_ok += t["first_acc_idx"] + 1
        if k_ok >= min(10, len(o_st["acceptance"])):
            break
    assert k_ok >= min(10, len(o_st["acceptance"])), "every seed tried puts a decision inside the error band too early"
    if k_ok == len(o_st["acceptance"]):
        assert accepted == o_acc and len(coords) == len(o_coords)
    np.testing.assert_allclose(coords[: k_ok + 1], o_coords[: k_ok + 1], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(velocs[: k_ok + 1], o_velocs[: k_ok + 1], rtol=1e-4, atol=2e-5)
    np.testing.assert_array_equal(stats.acceptance_indicator[:k_ok], o_st["acceptance_indicator"][:k_ok])
    for f, name in (("p_xy", "p_xy"), ("p_yx", "p_yx")):
        np.testing.assert_allclose(getattr(stats, f)[:k_ok], o_st[name][:k_ok], rtol=1e-4, atol=1e-4)
    for f in ("exponent", "energies_pot", "energies_kin", "energies_pot_delta", "energies_kin_delta"):
        np.testing.assert_allclose(getattr(stats, f)[:k_ok], o_st[f][:k_ok], rtol=1e-4, atol=2e-2)
    np.testing.assert_allclose(stats.acceptance[:k_ok], o_st["acceptance"][:k_ok], rtol=0, atol=2e-2)
    if chirality is False and S == 10:
        assert o_st["acceptance_indicator"].any(), "the test configuration must exercise the accept branch"


def test_mh_chains_bench_configuration_flip_set():
    """>= 20 lock-step iterations of the bench configuration against the oracle, re-synchronised on the GPU state every
    iteration.  Writes the measured flip set / exponent-error distribution to gpurun_out/mh_flip_set.json."""
    steps = int(os.environ.get("TW_MH_PARITY_STEPS", "20"))
    B = int(os.environ.get("TW_MH_PARITY_CHAINS", "256"))
    pep = tetrapeptide_2olx()
    V = pep.num_atoms
    m, sd = _model(FULL_O, "bf16x3", "proposal")
    sysd = amber99sbildn_obc2(pep)  # the bench's energy: ff99SB-ILDN + OBC2 pinned to the reference's OpenMM fixtures
    s32 = sysd.as_float32()
    energy = PeptidePotentialEnergy(sysd)
    kbT = energy.kbT
    g = torch.Generator().manual_seed(1000)
    x0 = torch.tensor(pep.coords_nm, dtype=torch.float32)[None] + 0.005 * torch.randn(B, V, 3, generator=g)
    at = torch.tensor(pep.atom_types)[None].repeat(B, 1)
    mask = torch.zeros(B, V, dtype=torch.bool)
    torch.manual_seed(0)
    chains = sampling.MHChains(m, energy, at.cuda(), mask.cuda(), x0.cuda())
    torch.set_num_threads(os.cpu_count() or 1)
    c = lambda t: t.detach().cpu()  # noqa: E731
    err_exp, err_pxy, err_pyx, err_e, gap, flips_gap, dec_gpu, dec_ref = [], [], [], [], [], [], [], []
    for it in range(steps):
        x_before, e_before = c(chains.x).clone(), c(chains.e_pot_x).clone()
        torch.manual_seed(100 + it)
        acc = c(chains.step())
        torch.manual_seed(100 + it)
        nc, nv, o_acc, u, rec = mo.mh_step(sd, FULL_O, s32, kbT, at, x_before, mask, CudaReplayDraws(), e_pot_x=e_before)
        last = {k: c(v) for k, v in chains.last.items()}
        err_exp.append((last["exponent"] - rec.exponent).numpy())
        err_pxy.append((last["p_xy"] - rec.p_xy).numpy())
        err_pyx.append((last["p_yx"] - rec.p_yx).numpy())
        err_e.append((last["e_pot_y"] - rec.e_pot_y).numpy())
        d = (u - rec.p_acc).abs().numpy()
        gap.append(d)
        flip = (acc != o_acc).numpy()
        flips_gap.append(d[flip])
        dec_gpu.append(acc.numpy()), dec_ref.append(o_acc.numpy())
        # the accepted states are the oracle's proposals (to fp32 round-off of the sampling pass)
        both = (acc & o_acc).numpy()
        np.testing.assert_allclose(c(chains.x).numpy()[both], rec.y_coords.numpy()[both], rtol=1e-4, atol=2e-5)
        same_rej = (~acc & ~o_acc).numpy()
        np.testing.assert_array_equal(c(chains.x).numpy()[same_rej], x_before.numpy()[same_rej])
    cat = lambda a: np.concatenate(a)  # noqa: E731
    err_exp, gap, flips_gap = cat(err_exp), cat(gap), cat(flips_gap)
    dec_gpu, dec_ref = cat(dec_gpu), cat(dec_ref)
    q = lambda a: {"median": float(np.median(np.abs(a))), "p99": float(np.quantile(np.abs(a), 0.99)), "max": float(np.abs(a).max())}  # noqa: E731
    report = {
        "config": f"FULL kernel_transformer_nvp, bf16x3, proposal weights, 2olx-65, {B} chains x {steps} lock-step iterations",
        "decisions": int(dec_gpu.size), "accept_rate_gpu": float(dec_gpu.mean()), "accept_rate_oracle": float(dec_ref.mean()),
        "flips": int((dec_gpu != dec_ref).sum()), "flip_gaps_abs_u_minus_p": sorted(float(v) for v in flips_gap),
        "eps_p": EPS_P, "decisions_inside_eps_band": int((gap < EPS_P).sum()),
        "exponent_error_nats": q(err_exp), "p_xy_error_nats": q(cat(err_pxy)), "p_yx_error_nats": q(cat(err_pyx)),
        "e_pot_y_error_kT": q(cat(err_e)),
    }
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/mh_flip_set.json", "w") as f:
        json.dump(report, f, indent=1)
    print("MH flip set:", json.dumps(report))
    assert 0.05 < dec_ref.mean() < 0.95, "the configuration must exercise both branches of the MH rule"
    assert (flips_gap < EPS_P).all(), f"a decision outside the +-{EPS_P} band differs from the oracle"
    assert report["flips"] <= max(3, 0.02 * dec_gpu.size)
    assert report["exponent_error_nats"]["max"] < 2e-2 and report["exponent_error_nats"]["median"] < 2e-3


@pytest.mark.parametrize("threshold", [300.0, 0.0])
def test_explore_matches_oracle_steps(threshold):
    """exploration.py:229-250 step for step (FULL model, bf16x3, 2olx): proposals, energies and the keep / reject decision
    against the oracle on the replayed stream; decisions may differ only where |dE - threshold| is inside the energy error."""
    pep = tetrapeptide_2olx()
    V, P, steps = pep.num_atoms, 48, 4
    m, sd = _model(FULL_O, "bf16x3", "proposal")
    sysd = amber99sbildn_obc2(pep)
    s32 = sysd.as_float32()
    energy = PeptidePotentialEnergy(sysd)
    x = torch.tensor(pep.coords_nm, dtype=torch.float32)[None]
    at = torch.tensor(pep.atom_types)[None]
    mask = torch.zeros(1, V, dtype=torch.bool)
    centers = find_chirality_centers(torch.tensor(pep.bonds), at)
    ref_signs = compute_chirality_sign(x.cuda(), centers.cuda())
    g = torch.Generator().manual_seed(2)
    v0 = torch.randn(1, V, 3, generator=g)
    torch.manual_seed(9)
    pos, en, n_acc = sampling.explore(m, energy, at.cuda(), mask.cuda(), x.cuda(), v0.cuda(), num_steps=steps, num_chains=P, threshold=threshold,
                                      chirality_centers=centers, reference_signs=ref_signs)
    pos, en = pos.cpu().reshape(steps, P, V, 3), en.cpu().reshape(steps, P)
    torch.manual_seed(9)
    draws = CudaReplayDraws()
    y, yv = x.repeat(P, 1, 1), v0.repeat(P, 1, 1)
    e = torch.from_numpy(np.asarray(mo.eo.potential_energy(s32, y.numpy().astype(np.float64)))).float()
    n_flip, total_acc = 0, 0
    for s in range(steps):
        y_o, e_o, ok, y_new, e_new, yv = mo.explore_step(sd, FULL_O, s32, at.repeat(P, 1), y, yv, e, mask.repeat(P, 1), draws, threshold=threshold,
                                                         chirality_centers=centers, reference_signs=ref_signs.cpu())
        moved = (pos[s] != y).flatten(1).any(1)  # the product kept the proposal
        flip = moved != ok
        assert ((e_new - e - threshold).abs()[flip] < 5e-2).all()  # kJ/mol: only ties inside the energy error may differ
        agree = ~flip
        np.testing.assert_allclose(pos[s][agree & ok].numpy(), y_new[agree & ok].numpy(), rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(en[s][agree].numpy(), e_o[agree].numpy(), rtol=1e-5, atol=5e-2)
        n_flip += int(flip.sum())
        total_acc += int(moved.sum())
        y, e = pos[s], en[s]  # re-synchronise on the product's state
    assert n_flip <= 2
    if threshold == 0.0:
        assert 0 < total_acc < steps * P
    assert int(n_acc.sum()) == total_acc
