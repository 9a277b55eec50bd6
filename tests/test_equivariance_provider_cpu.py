"""Host logic of the batch augmentation (equivariance/equivariance_transforms.py:15-175) and of the multi-protein energy
provider (utils/openmm/openmm_provider.py:20-175): no GPU needed."""
import os

import numpy as np
import pytest
import torch

from timewarp_b200.dataloader import DenseMolDynBatch
from timewarp_b200 import equivariance as eq

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FIELDS = ("atom_coords", "atom_velocs", "atom_forces", "atom_coord_targets", "atom_veloc_targets", "atom_force_targets")


def _batch(d):
    B, V = d["in_atom_coords"].shape[:2]
    return DenseMolDynBatch(names=["a", "b", "c"], atom_types=torch.tensor(d["in_atom_types"]),
                            adj_list=torch.tensor([[0, 1], [1, 2], [7, 8]]), edge_batch_idx=torch.tensor([0, 0, 1]),
                            masked_elements=torch.zeros(B, V, dtype=torch.bool), **{k: torch.tensor(d["in_" + k]) for k in FIELDS})


def test_transform_batch_matches_reference_golden():
    """Same seeds, same draws (scipy's rotation from numpy's global generator, ONE torch.randn(1, 3)): bit-equal to the
    reference's transform_batch."""
    d = np.load(os.path.join(GOLD, "equivariance_batch.npz"))
    np.random.seed(5)
    torch.manual_seed(6)
    out = eq.transform_batch(_batch(d))
    for k in FIELDS:
        np.testing.assert_array_equal(getattr(out, k).numpy(), d["out_" + k], err_msg=k)
    np.testing.assert_array_equal(out.atom_types.numpy(), d["out_atom_types"])
    np.testing.assert_array_equal(out.adj_list.numpy(), d["out_adj_list"])
    assert out.names == ["a", "b", "c"]


def test_rigid_motion_properties():
    d = np.load(os.path.join(GOLD, "equivariance_batch.npz"))
    b = _batch(d)
    out = eq.transform_batch(b)
    dist = lambda x: torch.cdist(x, x)  # noqa: E731
    torch.testing.assert_close(dist(out.atom_coords), dist(b.atom_coords), atol=1e-5, rtol=1e-5)  # coordinates: isometry
    torch.testing.assert_close(out.atom_velocs.norm(dim=-1), b.atom_velocs.norm(dim=-1), atol=1e-5, rtol=1e-5)  # vectors: rotation only
    # the displacement x -> target rotates like a vector (the translation cancels)
    torch.testing.assert_close((out.atom_coord_targets - out.atom_coords).norm(dim=-1), (b.atom_coord_targets - b.atom_coords).norm(dim=-1),
                               atol=1e-5, rtol=1e-5)


def test_composition_and_permutation():
    t = eq.Translation(torch.tensor([[1.0, 2.0, 3.0]]))
    r = eq.Rotation(torch.tensor([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]]))
    x = torch.tensor([[1.0, 0.0, 0.0]])
    both = t + r  # translation first, rotation second
    torch.testing.assert_close(both.transform_coord(x), torch.tensor([[-2.0, 2.0, 3.0]]))
    torch.testing.assert_close(both.transform_veloc(x), torch.tensor([[0.0, 1.0, 0.0]]))  # vectors are not translated
    assert len((both + t).transforms) == 3 and len((both + both).transforms) == 4
    p = eq.Permutation(torch.tensor([2, 0, 1]))
    feats = torch.tensor([10, 11, 12])
    assert p.transform_pointwise_feature(feats).tolist() == [11, 12, 10]  # point i moves to slot permutation[i]
    assert p.transform_adjacency_list(torch.tensor([[0, 1]])).tolist() == [[2, 0]]
    with pytest.raises(NotImplementedError):
        p.transform_coord(torch.zeros(2, 3, 3))


def test_provider_cache_is_fifo(tmp_path, monkeypatch):
    """Cache semantics of openmm_provider.py:57-75,110-143 with the module construction stubbed out (no GPU here)."""
    from timewarp_b200 import energy as en

    built = []

    class Fake:
        def __init__(self, system, integrator, platform_name="CUDA"):
            built.append(system)

        def to(self, device):
            return self

    monkeypatch.setattr(en, "OpenmmPotentialEnergyTorch", Fake)
    monkeypatch.setattr(en.OpenMMProvider, "get_system", lambda self, protein: protein)
    prov = en.OpenMMProvider(str(tmp_path), parameters="T1-peptides", device="cpu", cache_size=2)
    a = prov.get_potential_energy_module("A")
    assert prov.get_potential_energy_module("A") is a and built == ["A"]
    prov.get_potential_energy_module("B")
    prov.get_potential_energy_module("C")  # drops A (first in)
    assert list(prov._potential_energy_cache) == ["B", "C"]
    assert prov.get_potential_energy_module("A") is not a and built == ["A", "B", "C", "A"]
    assert list(prov._potential_energy_cache) == ["C", "A"]
    prov.clear_cache_to_size(1)
    assert list(prov._potential_energy_cache) == ["A"]
    none = en.OpenMMProvider(str(tmp_path), device="cpu", cache_size=0)
    none.get_potential_energy_module("A")
    assert none._potential_energy_cache == {}
    assert abs(prov.kbT - 8.31446261815324e-3 * 310.0) < 1e-12


def test_provider_finds_pdb_and_builds_the_pinned_system(tmp_path):
    """get_system / get_masses walk `pdb_dirs` for <protein>-traj-state0.pdb (openmm_provider.py:87-108) and build the
    ff99SB-ILDN + OBC2 description of the 2olx tetrapeptide (the system pinned to the reference's golden energies)."""
    from timewarp_b200.energy import OpenMMProvider
    from timewarp_b200.peptides import tetrapeptide_2olx

    pep = tetrapeptide_2olx()
    sub = tmp_path / "nested" / "dir"
    sub.mkdir(parents=True)
    with open(sub / "2olx-traj-state0.pdb", "w") as f:
        for i, (n, rn, ri, xyz) in enumerate(zip(pep.atom_names, pep.residue_names, pep.residue_index, pep.coords_nm * 10.0)):
            name = (" " + n) if len(n) < 4 else n
            f.write("ATOM  %5d %-4s %3s A%4d    %8.3f%8.3f%8.3f  1.00  0.00\n" % (i + 1, name, rn, ri, xyz[0], xyz[1], xyz[2]))
        f.write("ENDMDL\n")
    prov = OpenMMProvider([str(tmp_path)], parameters="T1-peptides", device="cpu")
    sysd = prov.get_system("2olx")
    assert sysd.getNumParticles() == pep.num_atoms
    np.testing.assert_allclose(prov.get_masses("2olx").numpy(), pep.masses, rtol=1e-6)
    with pytest.raises(ValueError):
        prov.get_system("nope")
