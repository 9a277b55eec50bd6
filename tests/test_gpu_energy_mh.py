"""Energy kernel vs the fp64 numpy oracle, MH / exploration / chirality kernels vs torch restatements
of the reference lines, and the sampling drivers end to end."""
import math

import numpy as np
import pytest
import torch

from oracle import energy_oracle as eo
from oracle import flow_oracle as fo
from tests.common import EMPTY_ADJ, EMPTY_EBI, FULL_O, TINY_O, build_model, load_golden
from timewarp_b200.chirality import check_symmetry_change, compute_chirality_sign, find_chirality_centers
from timewarp_b200.energy import OpenmmPotentialEnergyTorch, PeptidePotentialEnergy
from timewarp_b200.forcefield import amber_like_system
from timewarp_b200.peptides import alanine_dipeptide, tetrapeptide_2olx
from timewarp_b200 import sampling

pytestmark = pytest.mark.gpu


def _confs(pep, B, seed, noise=0.01):
    g = np.random.default_rng(seed)
    return (pep.coords_nm[None] + noise * g.standard_normal((B, pep.num_atoms, 3))).astype(np.float32)


@pytest.mark.parametrize("pep_fn,gb", [(alanine_dipeptide, "obc2"), (tetrapeptide_2olx, "obc2"), (tetrapeptide_2olx, "obc1")])
def test_energy_matches_oracle(pep_fn, gb):
    pep = pep_fn()
    sysd = amber_like_system(pep, gb=gb)
    energy = PeptidePotentialEnergy(sysd)
    x = _confs(pep, 33, 0)
    e, terms = energy(torch.from_numpy(x).cuda(), return_terms=True)
    assert e.shape == (33, 1) and e.dtype == torch.float32
    ref_terms = eo.energy_terms(sysd.as_float32(), x.astype(np.float64))
    # test_md.py:35-47 checks potential energies at atol 1e-3 kJ/mol; fp32 output rounding of |E|~1e3 is ~6e-5
    np.testing.assert_allclose(terms.cpu().numpy(), ref_terms, rtol=2e-6, atol=1e-3)
    np.testing.assert_allclose(e.cpu().numpy()[:, 0], ref_terms.sum(-1), rtol=2e-6, atol=1e-3)
    assert len(set(np.round(ref_terms[0], 6))) >= 3  # >= 3 distinct force components (test_md.py:50-83)


def test_energy_cutoff_and_options():
    pep = tetrapeptide_2olx()
    sysd = amber_like_system(pep)
    sysd.cutoff = 0.9  # make the cutoff bite inside the molecule
    sysd.reaction_field_eps = 78.3
    x = _confs(pep, 5, 1)
    e = PeptidePotentialEnergy(sysd)(torch.from_numpy(x).cuda())
    np.testing.assert_allclose(e.cpu().numpy()[:, 0], eo.potential_energy(sysd.as_float32(), x), rtol=2e-6, atol=1e-3)
    sysd.use_gb = False
    e = PeptidePotentialEnergy(sysd)(torch.from_numpy(x).cuda())
    np.testing.assert_allclose(e.cpu().numpy()[:, 0], eo.potential_energy(sysd.as_float32(), x), rtol=2e-6, atol=1e-3)


def test_energy_properties_and_nonfinite():
    pep = tetrapeptide_2olx()
    energy = OpenmmPotentialEnergyTorch(amber_like_system(pep), None, platform_name="CUDA")
    assert abs(energy.kbT - 2.577483411627504) < 1e-12 and energy.num_particles == 65
    x = torch.from_numpy(_confs(pep, 4, 2)).cuda()
    e = energy(x)
    # rigid motion invariance
    th = 0.9
    R = torch.tensor([[math.cos(th), -math.sin(th), 0], [math.sin(th), math.cos(th), 0], [0, 0, 1.0]], device="cuda")
    e2 = energy(x @ R.T + torch.tensor([0.5, -0.2, 1.0], device="cuda"))
    torch.testing.assert_close(e, e2, rtol=1e-5, atol=5e-3)
    # leading dims are flattened like the reference (openmm_bridge.py:292)
    assert energy(x.reshape(2, 2, 65, 3)).shape == (4, 1)
    # overlapping atoms / NaN coordinates give non-finite energies, never a crash (a12)
    bad = x.clone()
    bad[0, 40] = bad[0, 0]  # two non-bonded atoms on top of each other: LJ/Coulomb singularity
    bad[1, 3, 0] = float("nan")
    eb = energy(bad)
    assert not torch.isfinite(eb[0]).all() and not torch.isfinite(eb[1]).all() and torch.isfinite(eb[2:]).all()
    assert energy(x[:0]).shape == (0, 1)


def test_chirality_kernels_match_reference_golden():
    g = load_golden("chirality_2olx")
    signs = compute_chirality_sign(g["coords"].cuda(), g["centers"].cuda())
    assert torch.equal(signs.cpu(), g["signs"])
    changed = check_symmetry_change(g["coords"].cuda(), g["centers"].cuda(), g["ref_signs"].cuda())
    assert torch.equal(changed.cpu(), g["changed"])
    # mirror => every centre flips (reference tests/test_chirality.py:91-99)
    mirrored = g["coords"] * torch.tensor([1.0, 1.0, -1.0])
    assert torch.equal(compute_chirality_sign(mirrored.cuda(), g["centers"].cuda()).cpu(), -g["signs"])


def test_mh_accept_kernel():
    n = 4096
    g = torch.Generator().manual_seed(0)
    epx, epy, ekx, eky, pxy, pyx = (torch.randn(n, generator=g) * s for s in (30, 30, 5, 5, 200, 200))
    u = torch.rand(n, generator=g)
    epy[5] = float("nan"); epy[6] = float("inf"); epy[7] = -float("inf")
    ex_ref = ((epy - epx) + (eky - ekx) + pxy) - pyx  # evaluation_utils.py:644-663
    p_ref = torch.min(torch.tensor(1.0), torch.exp(-ex_ref))
    acc_ref = u < p_ref
    ex, pa, acc, first = sampling.mh_accept(*(t.cuda() for t in (epx, epy, ekx, eky, pxy, pyx, u)), want_first=True)
    torch.testing.assert_close(ex.cpu(), ex_ref, rtol=0, atol=0, equal_nan=True)
    torch.testing.assert_close(pa.cpu(), p_ref, rtol=2e-6, atol=0, equal_nan=True)
    # decisions are identical except where |u - p_acc| is inside exp() round-off
    diff = acc.view(torch.bool).cpu() != acc_ref
    assert (diff & ((u - p_ref).abs() > 1e-6)).sum() == 0
    assert not acc[5] and not acc[6] and acc[7]  # NaN/+inf reject, -inf accepts (a12)
    assert int(first.item()) == int(acc.view(torch.bool).cpu().nonzero()[0])
    # state update in place only where accepted
    V = 7
    x, xv, y, yv = (torch.randn(n, V, 3, generator=g).cuda() for _ in range(4))
    x0, xv0 = x.clone(), xv.clone()
    _, _, acc2, _ = sampling.mh_accept(*(t.cuda() for t in (epx, epy, ekx, eky, pxy, pyx, u)), x_coords=x, x_velocs=xv, y_coords=y, y_velocs=yv)
    a = acc2.view(torch.bool)[:, None, None]
    assert torch.equal(x, torch.where(a, y, x0)) and torch.equal(xv, torch.where(a, yv, xv0))
    # nothing accepted -> -1
    _, _, _, first = sampling.mh_accept(epx.cuda(), (epx + 1e6).cuda(), ekx.cuda(), ekx.cuda(), pxy.cuda(), pxy.cuda(), u.cuda(), want_first=True)
    assert int(first.item()) == -1


def test_kinetic_energy():
    v = torch.randn(9, 22, 3)
    m = torch.rand(22) * 10 + 1
    kbT = 2.5774834
    a = sampling.compute_kinetic_energy(v.cuda(), None, random_velocs=True)
    torch.testing.assert_close(a.cpu(), 0.5 * (v**2.0).sum(-1).sum(-1), rtol=1e-5, atol=1e-5)
    b = sampling.compute_kinetic_energy(v.cuda(), m.cuda(), random_velocs=False, kbT=kbT)
    torch.testing.assert_close(b.cpu(), 0.5 * (m * (v**2.0).sum(-1)).sum(-1) / kbT, rtol=1e-5, atol=1e-5)


class _Batch:
    def __init__(self, pep, dev):
        self.atom_coords = torch.tensor(pep.coords_nm, dtype=torch.float32)[None]
        self.atom_velocs = torch.zeros_like(self.atom_coords)
        self.atom_types = torch.tensor(pep.atom_types)[None]
        self.masked_elements = torch.zeros(1, pep.num_atoms, dtype=torch.bool)
        self.adj_list = torch.tensor(pep.bonds)
        self.edge_batch_idx = torch.zeros(len(pep.bonds), dtype=torch.long)


@pytest.mark.parametrize("random_velocs", [True, False])
@pytest.mark.parametrize("accept", [True, False])
@pytest.mark.parametrize("S", [1, 10])
def test_sample_with_model_shapes(random_velocs, accept, S):
    """Reference tests/test_evaluation_utils.py:112-138: len(sampled_coords) == len(chain_stats.acceptance) + 1."""
    if not accept and S > 1:
        pytest.skip("reference raises for accept=False with several proposals")
    pep = alanine_dipeptide()
    m, _ = build_model(TINY_O, "fp32", 0)
    energy = PeptidePotentialEnergy(amber_like_system(pep))
    batch = _Batch(pep, "cuda")
    centers = find_chirality_centers(batch.adj_list, batch.atom_types)
    ref_signs = compute_chirality_sign(batch.atom_coords.cuda(), centers.cuda())
    torch.manual_seed(0)
    coords, velocs, accepted, stats = sampling.sample_with_model(
        batch, m, torch.device("cuda"), energy, torch.tensor(pep.masses, dtype=torch.float32), num_samples=12, accept=accept,
        random_velocs=random_velocs, resample_velocs=random_velocs, num_proposal_steps=S, reference_signs=ref_signs, chirality_centers=centers)
    assert len(coords) == len(stats.acceptance) + 1 == len(velocs)
    assert coords.shape[1:] == (22, 3) and len(stats) >= 12
    assert 0 <= accepted <= len(stats)
    for f in ("p_xy", "p_yx", "exponent", "energies_pot", "energies_kin", "energies_pot_delta", "energies_kin_delta", "acceptance_indicator"):
        assert len(getattr(stats, f)) == len(stats)


def test_mh_chains_against_oracle_step():
    """One lock-step MH iteration over B chains == the reference formulae evaluated with the CPU oracles."""
    pep = alanine_dipeptide()
    m, sd = build_model(TINY_O, "fp32", 0)
    sysd = amber_like_system(pep)
    energy = PeptidePotentialEnergy(sysd)
    B = 16
    x0 = torch.from_numpy(_confs(pep, B, 3)).cuda()
    at = torch.tensor(pep.atom_types)[None].repeat(B, 1).cuda()
    mask = torch.zeros(B, 22, dtype=torch.bool, device="cuda")
    torch.manual_seed(11)
    chains = sampling.MHChains(m, energy, at, mask, x0)
    x_before, e_before = chains.x.clone(), chains.e_pot_x.clone()
    torch.manual_seed(12)
    acc = chains.step()
    # replay the RNG stream: velocities, latents (coords, velocs), uniforms
    torch.manual_seed(12)
    xv = torch.randn_like(x_before)
    zc = torch.empty(1, B, 22, 3, device="cuda").normal_() * torch.exp(m.coords_prior_log_scale.detach())
    zv = torch.empty(1, B, 22, 3, device="cuda").normal_() * torch.exp(m.velocs_prior_log_scale.detach())
    u = torch.rand(B, device="cuda")
    c = lambda t: t.detach().cpu()  # noqa: E731
    yc, yv, p_xy = fo.conditional_sample_with_logp(sd, TINY_O, c(at), c(x_before), c(xv), c(mask), 1, c(zc), c(zv), distance_mode="direct")
    p_yx = fo.log_likelihood(sd, TINY_O, c(at), yc[0], yv[0], c(x_before), c(xv), c(mask), distance_mode="direct")
    kbT = energy.kbT
    e_y = torch.from_numpy(eo.potential_energy(sysd.as_float32(), yc[0].numpy().astype(np.float64))).float() / kbT
    e_kin = 0.5 * (yv[0] ** 2).sum((-1, -2)) - 0.5 * (c(xv) ** 2).sum((-1, -2))
    ex = (e_y - c(e_before)) + e_kin + p_xy[0] - p_yx
    torch.testing.assert_close(c(chains.last["exponent"]), ex, rtol=1e-4, atol=2e-2)
    p_acc = torch.clamp(torch.exp(-ex), max=1.0)
    safe = (c(u) - p_acc).abs() > 1e-3
    assert torch.equal(c(acc)[safe], (c(u) < p_acc)[safe])
    torch.testing.assert_close(c(chains.x), torch.where(c(acc)[:, None, None], yc[0], c(x_before)), rtol=1e-4, atol=1e-5)


def test_explore_rule():
    pep = tetrapeptide_2olx()
    m, _ = build_model(TINY_O, "fp32", 0)
    energy = PeptidePotentialEnergy(amber_like_system(pep))
    x = torch.tensor(pep.coords_nm, dtype=torch.float32)[None].cuda()
    at = torch.tensor(pep.atom_types)[None].cuda()
    mask = torch.zeros(1, 65, dtype=torch.bool, device="cuda")
    centers = find_chirality_centers(torch.tensor(pep.bonds), at.cpu())
    ref_signs = compute_chirality_sign(x, centers.cuda())
    torch.manual_seed(0)
    pos, en, n_acc = sampling.explore(m, energy, at, mask, x, torch.randn_like(x), num_steps=3, num_chains=8, threshold=300.0,
                                      chirality_centers=centers, reference_signs=ref_signs)
    assert pos.shape == (24, 65, 3) and en.shape == (24,) and n_acc.shape == (8,)
    # energies never rise by more than the threshold between consecutive kept states (exploration.py:243-246)
    e = en.reshape(3, 8)
    e0 = energy(x).squeeze()
    assert ((e[0] - e0) <= 300.0 + 1e-3).all() and ((e[1:] - e[:-1]) <= 300.0 + 1e-3).all()
    # stored energies are the energies of the stored positions
    torch.testing.assert_close(energy(pos).squeeze(-1), en, rtol=1e-5, atol=2e-2)


@pytest.mark.parametrize("random_velocs", [True, False])
def test_sample_on_batches_against_oracle(random_velocs):
    """utils/evaluation_utils.py:190-353 on three B=1 dataset pairs: shapes of the 11-tuple, and every log-density /
    acceptance probability against the CPU oracles on the replayed RNG stream."""
    pep = alanine_dipeptide()
    m, sd = build_model(TINY_O, "fp32", 0)
    sysd = amber_like_system(pep)
    energy = PeptidePotentialEnergy(sysd)
    masses = torch.tensor(pep.masses, dtype=torch.float32)
    g = torch.Generator().manual_seed(5)
    batches = []
    for _ in range(3):
        b = _Batch(pep, "cuda")
        b.atom_coords = b.atom_coords + 0.01 * torch.randn(1, 22, 3, generator=g)
        b.atom_velocs = torch.randn(1, 22, 3, generator=g)
        b.atom_coord_targets = b.atom_coords + 0.02 * torch.randn(1, 22, 3, generator=g)
        b.atom_veloc_targets = torch.randn(1, 22, 3, generator=g)
        batches.append(b)
    torch.manual_seed(21)
    out = sampling.sample_on_batches(batches, m, torch.device("cuda"), energy, False, masses, random_velocs=random_velocs)
    y_c, y_v, t_c, t_v, c_c, c_v, ll_rev, ll_fwd, ll_rev_tr, ll_fwd_tr, acc = out
    for a in (y_c, y_v, t_c, t_v, c_c, c_v):
        assert a.shape == (3, 22, 3)
    for a in (ll_rev, ll_fwd, ll_rev_tr, ll_fwd_tr, acc):
        assert a.shape == (3, 1)
    # replay
    torch.manual_seed(21)
    kbT = energy.kbT
    sgn = 1.0 if random_velocs else -1.0
    for i, b in enumerate(batches):
        x = b.atom_coords
        if random_velocs:
            xv = torch.randn(1, 22, 3, device="cuda").cpu()
            tv = torch.randn(1, 22, 3, device="cuda").cpu()
        else:
            xv, tv = b.atom_velocs, b.atom_veloc_targets
        zc = (torch.empty(1, 1, 22, 3, device="cuda").normal_() * torch.exp(m.coords_prior_log_scale.detach())).cpu()
        zv = (torch.empty(1, 1, 22, 3, device="cuda").normal_() * torch.exp(m.velocs_prior_log_scale.detach())).cpu()
        at, mask = b.atom_types, b.masked_elements
        yc, yv, p_xy = fo.conditional_sample_with_logp(sd, TINY_O, at, x, xv, mask, 1, zc, zv, distance_mode="direct")
        yc, yv = yc[0], yv[0]
        np.testing.assert_allclose(y_c[i], yc[0].numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(c_v[i], xv[0].numpy(), rtol=0, atol=0)
        p_yx = fo.log_likelihood(sd, TINY_O, at, yc, sgn * yv, x, sgn * xv, mask, distance_mode="direct")
        p_xy_tr = fo.log_likelihood(sd, TINY_O, at, x, xv, b.atom_coord_targets, tv, mask, distance_mode="direct")
        p_yx_tr = fo.log_likelihood(sd, TINY_O, at, b.atom_coord_targets, sgn * tv, x, sgn * xv, mask, distance_mode="direct")
        for got, ref in ((ll_fwd[i], p_xy[0]), (ll_rev[i], p_yx), (ll_fwd_tr[i], p_xy_tr), (ll_rev_tr[i], p_yx_tr)):
            np.testing.assert_allclose(got, ref.numpy(), rtol=1e-4, atol=1e-4)
        s32 = sysd.as_float32()
        e_pot = (eo.potential_energy(s32, yc.numpy().astype(np.float64)) - eo.potential_energy(s32, x.numpy().astype(np.float64))) / kbT
        ke = lambda v: (0.5 * (v**2).sum((-1, -2)) if random_velocs else 0.5 * (masses * (v**2).sum(-1)).sum(-1) / kbT)  # noqa: E731
        ex = torch.from_numpy(e_pot).float() + ke(yv) - ke(xv) + p_xy[0] - p_yx
        p_ref = torch.clamp(torch.exp(-ex), max=1.0).numpy()
        np.testing.assert_allclose(acc[i], p_ref, rtol=5e-2, atol=1e-6)
    # data_augmentation=True (evaluation_utils.py:227-229): one random rigid motion per batch, drawn like the reference draws it
    # (scipy rotation from numpy's generator, one CPU torch.randn(1, 3)); equal to running on hand-transformed batches
    from timewarp_b200.dataloader import DenseMolDynBatch
    from timewarp_b200.equivariance import transform_batch

    dense = [DenseMolDynBatch(names=["ad"], atom_types=b.atom_types, adj_list=b.adj_list, edge_batch_idx=b.edge_batch_idx,
                              atom_coords=b.atom_coords, atom_velocs=b.atom_velocs, atom_forces=torch.zeros_like(b.atom_coords),
                              atom_coord_targets=b.atom_coord_targets, atom_veloc_targets=b.atom_veloc_targets,
                              atom_force_targets=torch.zeros_like(b.atom_coords), masked_elements=b.masked_elements) for b in batches]
    np.random.seed(3)
    torch.manual_seed(21)
    aug = sampling.sample_on_batches(dense, m, torch.device("cuda"), energy, True, masses, random_velocs=random_velocs)
    np.random.seed(3)
    torch.manual_seed(21)
    moved = [transform_batch(b) for b in dense]  # (CPU generator only: the CUDA stream of the model is untouched)
    ref = sampling.sample_on_batches(moved, m, torch.device("cuda"), energy, False, masses, random_velocs=random_velocs)
    for a, b in zip(aug, ref):
        np.testing.assert_array_equal(a, b)
    assert not np.allclose(aug[4], out[4])  # the conditioning states really moved
    with pytest.raises(AssertionError):  # like the reference: only dense batches can be augmented
        sampling.sample_on_batches(batches, m, torch.device("cuda"), energy, True, masses)


def test_mh_chains_graph_replay_matches_eager():
    """One MH iteration replayed as a CUDA graph == the same iteration launched eagerly (same RNG stream, same state)."""
    pep = alanine_dipeptide()
    m, _ = build_model(TINY_O, "fp32", 0)
    energy = PeptidePotentialEnergy(amber_like_system(pep))
    B = 8
    x0 = torch.from_numpy(_confs(pep, B, 5)).cuda()
    at = torch.tensor(pep.atom_types)[None].repeat(B, 1).cuda()
    mask = torch.zeros(B, 22, dtype=torch.bool, device="cuda")
    results = []
    for use_graph in (False, True):
        torch.manual_seed(3)
        chains = sampling.MHChains(m, energy, at, mask, x0, accept=False)  # accept everything: the state moves every step
        if use_graph:
            chains.capture_graph(warmup=2)
        else:
            chains.step(), chains.step()  # a capture pass records kernels without running them: no state / RNG change
        torch.manual_seed(4)
        for _ in range(3):
            acc = chains.step()
        torch.cuda.synchronize()
        results.append((chains.x.clone(), chains.last["exponent"].clone(), acc.clone(), chains.n_steps))
    assert results[0][3] == results[1][3] == 5
    torch.testing.assert_close(results[0][0], results[1][0], rtol=0, atol=0)
    torch.testing.assert_close(results[0][1], results[1][1], rtol=0, atol=0, equal_nan=True)
    assert torch.equal(results[0][2], results[1][2])


@pytest.mark.parametrize("pep_fn,gb", [(alanine_dipeptide, "obc2"), (tetrapeptide_2olx, "obc2"), (tetrapeptide_2olx, "obc1")])
def test_forces_match_finite_differences(pep_fn, gb):
    """Analytic forces of the CUDA kernel == -dU/dx by central differences of the fp64 oracle (every term incl. the
    chain rule through the OBC Born radii); autograd through the energy module returns the same gradient."""
    pep = pep_fn()
    sysd = amber_like_system(pep, gb=gb)
    energy = PeptidePotentialEnergy(sysd)
    x = _confs(pep, 2, 11, noise=0.01)
    xt = torch.from_numpy(x).cuda()
    e, f = energy.energy_and_forces(xt)
    s32 = sysd.as_float32()
    h = 1e-5
    rng = np.random.default_rng(0)
    picks = [(b, int(i), int(k)) for b in range(2) for i, k in zip(rng.integers(0, pep.num_atoms, 12), rng.integers(0, 3, 12))]
    fmax = float(f.abs().max())
    for b, i, k in picks:
        xp, xm = x[b:b + 1].astype(np.float64).copy(), x[b:b + 1].astype(np.float64).copy()
        xp[0, i, k] += h
        xm[0, i, k] -= h
        num = -(eo.potential_energy(s32, xp)[0] - eo.potential_energy(s32, xm)[0]) / (2 * h)
        assert abs(float(f[b, i, k]) - num) < 2e-5 * fmax + 1e-2, (b, i, k, float(f[b, i, k]), num)
    # autograd: d(sum w_b U_b)/dx = -w_b F_b
    xg = xt.clone().requires_grad_(True)
    w = torch.tensor([[0.5], [-2.0]], device="cuda")
    (energy(xg) * w).sum().backward()
    torch.testing.assert_close(xg.grad, -f * w[:, :, None], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(energy(xt), e, rtol=0, atol=0)  # the no-grad path gives the same energies
    # translation invariance: forces sum to zero
    assert float(f.sum(1).abs().max()) < 1e-3 * fmax


# ------------------------------------------------------------------------------------------
# OpenMM integrator steps inside the chain: tw_langevin_steps vs oracle/md_oracle.py (itself pinned to
# the reference's trajectory fixtures, tests/test_md_oracle.py)
def _md_setup(pep, friction=0.3, middle=False, temperature=310.0):
    from timewarp_b200 import md

    sysd = amber_like_system(pep)
    cls = md.LangevinMiddleIntegrator if middle else md.LangevinIntegrator
    return sysd, md.Simulation(sysd, cls(temperature, friction, 0.0005), seed=7)


@pytest.mark.parametrize("middle", [False, True])
def test_langevin_kernel_matches_oracle(middle):
    from oracle import md_oracle as mo

    pep = alanine_dipeptide()
    sysd, sim = _md_setup(pep, middle=middle)
    s32 = sysd.as_float32()
    m32 = np.asarray(sysd.masses, dtype=np.float32).astype(np.float64)
    rng = np.random.default_rng(3)
    B, steps = 3, 3
    x = _confs(pep, B, 4, noise=0.005)
    v = (rng.standard_normal(x.shape) * np.sqrt(sim.kbT / m32)[:, None]).astype(np.float32)
    noise = rng.standard_normal((steps, B, pep.num_atoms, 3)).astype(np.float32)
    xg, vg = sim.step(torch.from_numpy(x).cuda(), torch.from_numpy(v).cuda(), steps, noise=torch.from_numpy(noise).cuda())
    assert xg.shape == x.shape and xg.dtype == torch.float32
    xo, vo = mo.integrate(lambda c: eo.potential_energy(s32, c), x, v, noise, m32, 0.0005, 0.3, sim.kbT, middle=middle)
    np.testing.assert_allclose(xg.cpu().numpy(), xo, rtol=0, atol=5e-7)
    np.testing.assert_allclose(vg.cpu().numpy(), vo, rtol=0, atol=5e-5)
    assert np.abs(xo - x).max() > 1e-3  # the steps moved the atoms well beyond the tolerance
    # zero steps / empty batch are no-ops; inputs are not modified
    x_t = torch.from_numpy(x).cuda()
    x0, v0 = sim.step(x_t, torch.from_numpy(v).cuda(), 0)
    assert torch.equal(x0, x_t) and torch.equal(x_t.cpu(), torch.from_numpy(x))
    xe, _ = sim.step(x_t[:0], torch.from_numpy(v).cuda()[:0], 2)
    assert xe.shape == (0, pep.num_atoms, 3)


def test_langevin_kernel_conserves_energy_without_friction():
    """friction = 0 turns both rules into leapfrog / velocity Verlet: the total energy (kinetic part at the half-kicked
    velocities, as OpenMM reports it) has no drift over 2000 steps."""
    from oracle import md_oracle as mo

    pep = tetrapeptide_2olx()
    for middle in (False, True):
        sysd, sim = _md_setup(pep, friction=0.0, middle=middle)
        energy = PeptidePotentialEnergy(sysd)
        m = np.asarray(sysd.masses, dtype=np.float64)
        x = torch.from_numpy(_confs(pep, 4, 5, noise=0.002)).cuda()
        v = sim.velocities_to_temperature(x)

        def total(x, v):
            u, f = energy.energy_and_forces(x)
            vv, ff = v.cpu().numpy().astype(np.float64), f.cpu().numpy().astype(np.float64)
            # both rules kick first: the stored velocity is half a kick short of the velocity that belongs to x
            ke = mo.leapfrog_kinetic_energy(vv, ff, m, 0.0005)
            return u.cpu().numpy()[:, 0].astype(np.float64) + ke, ke

        e0, ke0 = total(x, v)
        x1, v1 = sim.step(x, v, 2000)
        e1, _ = total(x1, v1)
        assert np.isfinite(e1).all()
        assert np.abs(e1 - e0).max() < 0.02 * ke0.mean(), (middle, e0, e1, ke0)
        assert (x1 - x).abs().max() > 0.01  # 1 ps of dynamics


def test_langevin_kernel_thermostat_and_rng_stream():
    """In-kernel Philox noise: reproducible per (seed, offset), advancing between calls, and the friction + noise pair
    equilibrates the kinetic energy to kT/2 per degree of freedom."""
    pep = alanine_dipeptide()
    sysd, sim = _md_setup(pep, friction=20.0)
    x = torch.from_numpy(_confs(pep, 256, 6, noise=0.002)).cuda()
    v = torch.zeros_like(x)
    xa, va = sim.step(x, v, 50)
    xb, vb = sim.step(x, v, 50)  # the stream advanced: different noise
    assert not torch.equal(va, vb)
    _, sim2 = _md_setup(pep, friction=20.0)
    xc, vc = sim2.step(x, v, 50)  # same seed, same offset: identical
    assert torch.equal(xa, xc) and torch.equal(va, vc)
    x1, v1 = sim.step(x, v, 3000)  # 1.5 ps, gamma t = 30
    m = sim.masses(x.device)
    ke_per_dof = (0.5 * m[None, :, None] * v1 * v1).mean().item()
    assert abs(ke_per_dof / (0.5 * sim.kbT) - 1.0) < 0.03, ke_per_dof / (0.5 * sim.kbT)


def test_sample_with_model_with_integrator_steps():
    """openmm_on_current / openmm_on_proposal (utils/evaluation_utils.py:594-602,623-626): integrator steps inside the MH loop."""
    from timewarp_b200 import md

    pep = alanine_dipeptide()
    sysd = amber_like_system(pep)
    energy = PeptidePotentialEnergy(sysd)
    sim = md.Simulation(sysd, md.get_simulation_environment_integrator("T1-peptides"))
    assert isinstance(sim.integrator, md.LangevinIntegrator) and not isinstance(sim.integrator, md.LangevinMiddleIntegrator)
    assert isinstance(md.get_simulation_environment_integrator("T1B-peptides"), md.LangevinMiddleIntegrator)
    m, _ = build_model(TINY_O, "fp32", 0)
    batch = _Batch(pep, "cuda")
    masses = torch.as_tensor(pep.masses, dtype=torch.float32)
    # (accept=False with the proposal-side steps: a rejected proposal would leave no trace of them in the chain)
    for kw in (dict(openmm_on_current=True, accept=True), dict(openmm_on_proposal=True, accept=False)):
        torch.manual_seed(0)
        coords, velocs, accepted, stats = sampling.sample_with_model(
            batch, m, torch.device("cuda"), energy, masses, num_samples=6, random_velocs=True, resample_velocs=True,
            num_openmm_steps=5, sim=sim, num_proposal_steps=1, **kw)
        assert coords.shape == (7, pep.num_atoms, 3) and np.isfinite(coords).all() and len(stats) == 6
        torch.manual_seed(0)
        plain = sampling.sample_with_model(batch, m, torch.device("cuda"), energy, masses, num_samples=6, accept=kw["accept"],
                                           random_velocs=True, resample_velocs=True, num_proposal_steps=1)[0]
        assert not np.allclose(plain, coords)
    with pytest.raises(ValueError):
        md.openmm_step(sim, torch.zeros(1, pep.num_atoms, 3, device="cuda"))


def test_kinetic_energy_matches_openmm_fixture():
    """compute_kinetic_energy with masses (utils/evaluation_utils.py:432-436) on the velocities of the reference's OpenMM
    trajectory fixture reproduces the kinetic energies OpenMM recorded (simulation/tests/test_md.py:35-47 data)."""
    import os
    from tests.common import GOLDEN

    g = np.load(os.path.join(GOLDEN, "langevin_2olx_pairs.npz"))
    pep = tetrapeptide_2olx()
    m = pep.masses
    v = g["ke_velocities"].astype(np.float64) + 0.5 * float(g["timestep_ps"]) * g["ke_forces"].astype(np.float64) / m[:, None]
    kbT = 2.577483411627504
    ke = sampling.compute_kinetic_energy(torch.from_numpy(v.astype(np.float32)).cuda(), torch.from_numpy(m.astype(np.float32)).cuda(),
                                         random_velocs=False, kbT=kbT)
    np.testing.assert_allclose(ke.cpu().numpy() * kbT, g["ke_openmm"], rtol=3e-6, atol=0)


def test_sample_trajectory_writes_and_resumes(tmp_path):
    """sample_trajectory.py:217-281: chunk files `{protein}_trajectory_model_{i}.npz` (positions[::10], time); a second call
    with a larger budget resumes after the chunks on disk, starting from the last saved position."""
    pep = alanine_dipeptide()
    m, _ = build_model(TINY_O, "fp32", 0)
    energy = PeptidePotentialEnergy(amber_like_system(pep))
    masses = torch.tensor(pep.masses, dtype=torch.float32)
    out = str(tmp_path / "chain")
    batch = _Batch(pep, "cuda")
    assert sampling.sample_trajectory(batch, m, torch.device("cuda"), energy, masses, out, "ad", num_samples=40, saving_interval=20,
                                      mh=False, conserve_chirality=True) == 0
    import os
    assert sorted(os.listdir(out)) == ["ad_trajectory_model_0.npz", "ad_trajectory_model_1.npz"]
    z1 = np.load(os.path.join(out, "ad_trajectory_model_1.npz"))
    assert z1["positions"].shape == (3, 22, 3) and float(z1["time"]) > 0  # 21 states thinned by 10
    batch2 = _Batch(pep, "cuda")
    sampling.sample_trajectory(batch2, m, torch.device("cuda"), energy, masses, out, "ad", num_samples=60, saving_interval=20, mh=False)
    assert sorted(os.listdir(out))[-1] == "ad_trajectory_model_2.npz" and len(os.listdir(out)) == 3
    z2 = np.load(os.path.join(out, "ad_trajectory_model_2.npz"))
    np.testing.assert_array_equal(z2["positions"][0], z1["positions"][-1])  # resumed from the last saved position
    with pytest.raises(AssertionError):
        sampling.sample_trajectory(batch, m, torch.device("cuda"), energy, masses, out, "ad", num_samples=5, saving_interval=20)


def test_energy_kernel_reference_golden_energies_with_openmm():
    """The reference's golden potential energies (simulation/tests/test_md.py:35-47, atol 1e-3 kJ/mol) through the CUDA kernel;
    needs OpenMM + its Amber XML files to build the System (skipped otherwise; the CPU twin is in tests/test_forcefield_cpu.py)."""
    pytest.importorskip("openmm")
    import io
    import os
    from openmm import app, unit
    from tests.common import GOLDEN
    from timewarp_b200.forcefield import system_description_from_openmm

    g = np.load(os.path.join(GOLDEN, "langevin_2olx_pairs.npz"))
    pdb = app.PDBFile(io.StringIO(str(g["state0_pdb"])))
    ff = app.ForceField("amber99sbildn.xml", "amber99_obc.xml")
    system = ff.createSystem(pdb.topology, nonbondedMethod=app.CutoffNonPeriodic, nonbondedCutoff=2.0 * unit.nanometer, constraints=None)
    energy = PeptidePotentialEnergy(system_description_from_openmm(system))
    e = energy(torch.from_numpy(g["pot_positions"]).cuda()).cpu().numpy()[:, 0]
    np.testing.assert_allclose(e, g["pot_openmm"], rtol=0, atol=1e-3)


def test_mh_chains_with_integrator_steps():
    """MHChains with `openmm_on_current` / `openmm_on_proposal`: every chain takes its integrator steps in one launch; the carried
    potential energy follows the moved state; rejected chains still move (on_current) / do not move (on_proposal)."""
    from timewarp_b200 import md

    pep = alanine_dipeptide()
    sysd = amber_like_system(pep)
    energy = PeptidePotentialEnergy(sysd)
    m, _ = build_model(TINY_O, "fp32", 0)  # random weights: (almost) every proposal is rejected
    B = 16
    x0 = torch.from_numpy(_confs(pep, B, 8, noise=0.002)).cuda()
    at = torch.tensor(pep.atom_types)[None].repeat(B, 1).cuda()
    mask = torch.zeros(B, pep.num_atoms, dtype=torch.bool).cuda()
    masses = torch.tensor(pep.masses, dtype=torch.float32)
    sim = md.Simulation(sysd, md.get_simulation_environment_integrator("T1-peptides"))
    torch.manual_seed(2)
    cur = sampling.MHChains(m, energy, at, mask, x0, masses=masses, sim=sim, num_openmm_steps=20, openmm_on_current=True)
    acc = cur.step()
    moved = (cur.x - x0).abs().amax((1, 2))
    assert torch.all(moved > 1e-4) and torch.isfinite(cur.x).all()
    rej = ~acc
    assert rej.any()
    torch.testing.assert_close(cur.e_pot_x[rej], (energy(cur.x) / energy.kbT).squeeze(-1)[rej], rtol=1e-6, atol=1e-4)
    torch.manual_seed(2)
    prop = sampling.MHChains(m, energy, at, mask, x0, masses=masses, sim=sim, num_openmm_steps=20, openmm_on_proposal=True)
    acc2 = prop.step()
    assert torch.equal(prop.x[~acc2], x0[~acc2])  # a rejected proposal leaves no trace of its integrator steps
    with pytest.raises(AssertionError):
        sampling.MHChains(m, energy, at, mask, x0, sim=sim, num_openmm_steps=5, openmm_on_current=True)  # masses are required


def test_energy_kernel_reference_golden_energies_and_forces():
    """The reference's own energy check (simulation/tests/test_md.py:35-47: 40 OpenMM frames of 2olx, preset "T1-peptides")
    through the CUDA kernel with the pinned ff99SB-ILDN / OBC2 table: energies at atol 0.02 kJ/mol (the reference's 1e-3 holds
    between two runs of its single-precision platform; against fp64 accumulation the floor is that platform's rounding,
    measured max 0.011), forces at the reference's rtol 0.05 with atol 0.2 kJ/mol/nm (rms 0.03 of 930)."""
    import os

    from tests.common import GOLDEN
    from timewarp_b200.forcefield import amber99sbildn_obc2

    g = np.load(os.path.join(GOLDEN, "energy_2olx_openmm.npz"))
    pep = tetrapeptide_2olx()
    c_term = [i for i, (n, r) in enumerate(zip(pep.atom_names, pep.residue_index)) if n == "C" and r == max(pep.residue_index)][0]
    for tag, sysd in (("cpu", amber99sbildn_obc2(pep)), ("wide", amber99sbildn_obc2(pep, improper_choice={c_term: 0}))):
        energy = PeptidePotentialEnergy(sysd)
        x = torch.from_numpy(g[f"{tag}_positions"]).cuda()
        e, f = energy.energy_and_forces(x)
        e, f = e.cpu().numpy()[:, 0].astype(np.float64), f.cpu().numpy().astype(np.float64)
        ref_e, ref_f = g[f"{tag}_potential"], g[f"{tag}_forces"].astype(np.float64)
        print(tag, "dE mean %.4f std %.4f max %.4f; force rms residual %.4f" % ((e - ref_e).mean(), (e - ref_e).std(), np.abs(e - ref_e).max(),
                                                                                  np.sqrt(((f - ref_f) ** 2).mean())))
        np.testing.assert_allclose(e, ref_e, rtol=0, atol=0.02)
        np.testing.assert_allclose(f, ref_f, rtol=0.05, atol=0.2)
        assert np.sqrt(((f - ref_f) ** 2).mean()) < 0.1
        np.testing.assert_allclose(e, eo.potential_energy(sysd.as_float32(), g[f"{tag}_positions"].astype(np.float64)), rtol=0, atol=1e-3)


def test_openmm_provider_energy_modules(tmp_path):
    """OpenMMProvider (utils/openmm/openmm_provider.py:110-143): the module it builds from `<protein>-traj-state0.pdb` gives the
    energies of the pinned 2olx system, is cached, and feeds losses.compute_potential_energy."""
    from timewarp_b200.energy import OpenMMProvider
    from timewarp_b200.forcefield import amber99sbildn_obc2
    from timewarp_b200.losses import compute_potential_energy
    from timewarp_b200.peptides import tetrapeptide_2olx

    pep = tetrapeptide_2olx()
    with open(tmp_path / "2olx-traj-state0.pdb", "w") as f:
        for i, (n, rn, ri, xyz) in enumerate(zip(pep.atom_names, pep.residue_names, pep.residue_index, pep.coords_nm * 10.0)):
            name = (" " + n) if len(n) < 4 else n
            f.write("ATOM  %5d %-4s %3s A%4d    %8.3f%8.3f%8.3f  1.00  0.00\n" % (i + 1, name, rn, ri, xyz[0], xyz[1], xyz[2]))
        f.write("ENDMDL\n")
    prov = OpenMMProvider(str(tmp_path), parameters="T1-peptides", device="cuda", cache_size=2)
    mod = prov.get_potential_energy_module("2olx")
    assert prov.get_potential_energy_module("2olx") is mod
    x = torch.from_numpy(_confs(pep, 4, 9)).cuda()
    ref = PeptidePotentialEnergy(amber99sbildn_obc2(pep), temperature=310.0).cuda()(x)
    torch.testing.assert_close(mod(x), ref, atol=2e-3, rtol=0)  # (PDB coordinates are rounded to 1e-3 A: only the topology is used)
    mask = torch.zeros(4, pep.num_atoms, dtype=torch.bool, device="cuda")
    u = compute_potential_energy(x, ["2olx"] * 4, mask, prov)
    torch.testing.assert_close(u, ref.squeeze(-1) / prov.kbT, atol=1e-3, rtol=0)
