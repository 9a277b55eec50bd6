"""Host logic of the energy-based losses (timewarp_b200/losses.py) and of the OpenMM-shaped integrator objects
(timewarp_b200/md.py) that needs no GPU: energy bookkeeping (losses.py:24-149), presets (simulation/md.py:13-125), and the
no-CPU-fallback rule."""
import numpy as np
import pytest
import torch

from timewarp_b200 import _lib, losses, md
from timewarp_b200.forcefield import amber_like_system
from timewarp_b200.peptides import alanine_dipeptide


class _Quadratic:
    """A differentiable stand-in for the energy module: U(x) = k * sum x^2 per conformation, [B, 1]."""

    def __init__(self, k, kbT=2.5):
        self.k, self.kbT = k, kbT

    def __call__(self, coords):
        return self.k * (coords**2).sum((-1, -2))[:, None]


def test_kinetic_energy_modes():
    v, m = torch.randn(3, 5, 3), torch.rand(3, 5) + 0.5
    torch.testing.assert_close(losses.compute_kinetic_energy(v, None, None, random_velocs=True), 0.5 * (v**2).sum((-1, -2)))
    torch.testing.assert_close(losses.compute_kinetic_energy(v, m, 2.5), 0.5 * (m * (v**2).sum(-1)).sum(-1) / 2.5)


def test_potential_energy_segments_masking_and_gradient():
    prov = losses.EnergyProvider({"a": _Quadratic(1.0), "b": _Quadratic(3.0)}, {"a": torch.ones(4), "b": torch.ones(2)})
    assert prov.kbT == 2.5
    x = torch.randn(5, 4, 3, requires_grad=True)
    mask = torch.zeros(5, 4, dtype=torch.bool)
    mask[3:, 2:] = True  # protein "b" has two atoms: the padding must not reach its energy
    names = ["a", "a", "a", "b", "b"]
    seg = losses.compute_potential_energy(x, names, mask, prov, segments=[0, 3, 5])
    loop = losses.compute_potential_energy(x, names, mask, prov)
    torch.testing.assert_close(seg, loop)
    want = torch.cat([(x[:3] ** 2).sum((-1, -2)), 3.0 * (x[3:, :2] ** 2).sum((-1, -2))]) / 2.5
    torch.testing.assert_close(seg, want)
    seg.sum().backward()
    assert torch.all(x.grad[3:, 2:] == 0) and torch.all(x.grad[:3] != 0)
    e, (pot, kin) = losses.compute_energy(x.detach(), torch.ones(5, 4, 3), names, mask, prov, random_velocs=False, segments=[0, 3, 5])
    torch.testing.assert_close(kin, torch.tensor([6.0, 6.0, 6.0, 3.0, 3.0]) / 2.5)  # masses padded with zeros (losses.py:129-138)
    torch.testing.assert_close(e, pot + kin)
    with pytest.raises(AssertionError):
        losses.EnergyProvider({"a": _Quadratic(1.0, kbT=2.5), "b": _Quadratic(1.0, kbT=2.6)})
    with pytest.raises(ValueError):
        losses.AcceptanceLoss(prov, high_energy_threshold=300.0)


def test_md_presets_and_integrators():
    old = md.get_parameters_from_preset("T1-peptides")  # simulation/md.py:31-37,75-82
    assert old["integrator"] == "LangevinIntegrator" and old["forcefield"] == "amber99-implicit"
    assert (old["temperature"], old["friction"], old["timestep"]) == (310.0, 0.3, 0.0005)
    new = md.get_parameters_from_preset("T1B-peptides")
    assert new["integrator"] == "LangevinMiddleIntegrator" and new["forcefield"] == "amber14-implicit"
    assert md.get_parameters_from_preset({"x": 1}) == {"x": 1}
    with pytest.raises(ValueError, match="Invalid preset name"):
        md.get_parameters_from_preset("charmm36")
    it = md.get_simulation_environment_integrator("alanine-dipeptide")
    assert type(it) is md.LangevinIntegrator and it.kind == _lib.TW_INTEGRATOR_LANGEVIN
    assert (it.getTemperature(), it.getFriction(), it.getStepSize()) == (310.0, 0.3, 0.0005)
    assert md.get_simulation_environment_integrator("amber14-implicit").kind == _lib.TW_INTEGRATOR_LANGEVIN_MIDDLE

    class Q:  # openmm Quantity-like arguments are unwrapped
        def __init__(self, v):
            self._value = v

    assert md.LangevinIntegrator(Q(300.0), Q(1.0), Q(0.002)).getStepSize() == 0.002


def test_simulation_has_no_cpu_fallback():
    pep = alanine_dipeptide()
    sim = md.Simulation(amber_like_system(pep), md.LangevinIntegrator(310.0, 0.3, 0.0005))
    assert abs(sim.kbT - 2.577483411627504) < 1e-12
    x = torch.tensor(pep.coords_nm, dtype=torch.float32)[None]
    with pytest.raises(_lib.TimewarpB200Error, match="CUDA only"):
        sim.step(x, torch.zeros_like(x), 1)
    with pytest.raises(ValueError):
        md.openmm_step(sim, x)
    v = sim.velocities_to_temperature(x)  # host-side draw: N(0, kT/m) per component
    assert v.shape == x.shape and np.isfinite(v.numpy()).all()


def test_loss_wrapper_and_dispatch():
    """LossWrapper / get_loss / wrap_or_replace_loss / unwrap_loss_wrapper (losses.py:215-303) and the NLL glue (losses.py:321-356)
    with a stand-in model that records what it is called with."""
    calls = []

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(2))

        def forward(self, **kw):
            calls.append(kw)
            return kw["y_velocs"].sum() * 0 + self.w.sum()

    class Batch:
        atom_coords, atom_coord_targets = torch.zeros(2, 3, 3), torch.ones(2, 3, 3)
        atom_velocs, atom_veloc_targets = torch.full((2, 3, 3), 2.0), torch.full((2, 3, 3), 3.0)
        atom_types, masked_elements = torch.zeros(2, 3, dtype=torch.long), torch.zeros(2, 3, dtype=torch.bool)
        adj_list, edge_batch_idx = torch.zeros(0, 2, dtype=torch.long), torch.zeros(0, dtype=torch.long)

    m = Model()
    w = losses.LossWrapper(m, losses.NegativeLogLikelihoodLoss(random_velocs=False))
    w(Batch, device="cpu")
    assert torch.equal(calls[-1]["x_velocs"], Batch.atom_velocs) and torch.equal(calls[-1]["y_velocs"], Batch.atom_veloc_targets)
    assert torch.equal(calls[-1]["y_coords"], Batch.atom_coord_targets)
    torch.manual_seed(0)
    losses.LossWrapper(m, losses.NegativeLogLikelihoodLoss())(Batch)
    torch.manual_seed(0)
    a, b = torch.randn(2, 3, 3), torch.randn(2, 3, 3)  # conditioning velocities are drawn first (losses.py:333-334)
    assert torch.equal(calls[-1]["x_velocs"], a) and torch.equal(calls[-1]["y_velocs"], b)
    # (un)wrapping
    w2 = losses.wrap_or_replace_loss(losses.LossWrapper(w, None), losses.NegativeLogLikelihoodLoss())
    assert w2.module is m and losses.unwrap_loss_wrapper(losses.LossWrapper(w2)) is m
    with pytest.raises(AssertionError):
        losses.LossWrapper(m)(Batch)
    with pytest.raises(TypeError):
        losses.get_loss(object(), m, Batch)
    # state dicts of a wrapper ("module." prefix, DeepSpeed checkpoints) and of the bare module both load
    sd = {"w": torch.tensor([1.0, 2.0])}
    w.load_state_dict(sd)
    assert torch.equal(m.w.data, sd["w"])
    w.load_state_dict({"module.w": torch.tensor([3.0, 4.0])})
    assert torch.equal(m.w.data, torch.tensor([3.0, 4.0]))
    assert set(w.state_dict()) == {"module.w"}
