"""Batch / wire formats (timewarp_b200/dataloader.py) against golden outputs of the reference's own collate and trajectory
pairing (tests/golden/make_golden.py::dataloader_case)."""
import os

import numpy as np
import pytest
import torch

from timewarp_b200 import dataloader as dl
from timewarp_b200.peptides import alanine_dipeptide

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("atom_types", "adj_list", "atom_coords", "atom_velocs", "atom_forces", "atom_coord_targets", "atom_veloc_targets", "atom_force_targets")


def _golden():
    d = np.load(os.path.join(GOLDEN, "dataloader_collate.npz"))
    return {k: d[k] for k in d.files}


def test_dense_collate_matches_reference():
    g = _golden()
    pts = [dl.MolDynDatapoint(name=f"mol{i}", **{k: torch.from_numpy(g[f"in{i}_{k}"]) for k in FIELDS}) for i in range(3)]
    assert [p.num_atoms for p in pts] == [22, 15, 9]
    batch = dl.moldyn_dense_collate_fn(pts)
    assert batch.names == ["mol0", "mol1", "mol2"]
    for k in ("atom_types", "adj_list", "edge_batch_idx", "atom_coords", "atom_velocs", "atom_forces", "atom_coord_targets",
              "atom_veloc_targets", "atom_force_targets", "masked_elements"):
        got = getattr(batch, k)
        assert got.dtype == torch.from_numpy(g[f"batch_{k}"]).dtype, k
        assert torch.equal(got, torch.from_numpy(g[f"batch_{k}"])), k
    assert torch.equal(dl.lengths_to_mask(torch.tensor([3, 1, 4])), torch.from_numpy(g["lengths_mask"]))
    # padding is zeros and flagged; the keyword view feeds the model interface
    assert batch.masked_elements[1, 15:].all() and not batch.masked_elements[1, :15].any()
    assert float(batch.atom_coords[2, 9:].abs().max()) == 0.0
    kw = batch.model_kwargs()
    assert set(kw) == {"atom_types", "x_coords", "x_velocs", "y_coords", "y_velocs", "adj_list", "edge_batch_idx", "masked_elements"}
    half = batch.tofp16()
    assert half.atom_coords.dtype == torch.float16 and half.atom_types.dtype == torch.int64


def test_trajectory_pairing_matches_reference():
    g = _golden()
    ad = alanine_dipeptide()
    path = os.path.join(GOLDEN, "synthetic_ad-traj-arrays.npz")
    for sw, eq in ((1, False), (10, False), (100, True), (1000, False)):
        info = dl.load_pdb_trace_data("synthetic_ad", ad, path, step_width=sw, equal_data_spacing=eq)
        key = f"pairs_sw{sw}_eq{int(eq)}"
        np.testing.assert_array_equal(np.stack(info.coord_features), g[key + "_coord_features"])
        np.testing.assert_array_equal(np.stack(info.veloc_targets), g[key + "_veloc_targets"])
        np.testing.assert_array_equal(info.node_types, g[key + "_node_types"])
        assert {tuple(sorted(b)) for b in info.adj_list.tolist()} == {tuple(sorted(b)) for b in g[key + "_adj_list"].tolist()}
    pts = dl.datapoints_from_trajectory(info)
    assert len(pts) == len(info.coord_features) and pts[0].atom_coords.dtype == torch.float32 and pts[0].atom_types.dtype == torch.int64
    batch = dl.moldyn_dense_collate_fn(pts[:2])
    assert batch.atom_coords.shape == (2, 22, 3) and not batch.masked_elements.any()


def test_pdb_topology_reader(tmp_path):
    ad = alanine_dipeptide()
    lines = ["REMARK   1 CREATED WITH OPENMM 7.4.1"]
    for i, (n, r, ri, xyz) in enumerate(zip(ad.atom_names, ad.residue_names, ad.residue_index, ad.coords_nm * 10.0)):
        el = next(c for c in n if c.isalpha())
        lines.append("ATOM  %5d %-4s %3s A%4d    %8.3f%8.3f%8.3f  1.00  0.00          %2s" % (i + 1, n if len(n) == 4 else " " + n, r, ri, *xyz, el))
    lines += ["TER", "END"]
    p = tmp_path / "ad.pdb"
    p.write_text("\n".join(lines) + "\n")
    topo = dl.read_pdb_topology(str(p))
    assert topo.atom_names == ad.atom_names and topo.residue_names == ad.residue_names
    np.testing.assert_allclose(topo.coords_nm, ad.coords_nm, atol=1e-4)
    assert np.array_equal(topo.bonds, ad.bonds) and np.array_equal(topo.atom_types, ad.atom_types)


def test_raw_moldyn_dataset_directory(tmp_path, capsys):
    """RawMolDynDataset (datasets/iterable_datasets.py:21-129) over a directory with one complete trajectory, one without its
    npz and one corrupt one: names, graceful skipping, datapoints identical to the direct loader, collate-ready."""
    import shutil
    from timewarp_b200 import datasets as ds

    ad = alanine_dipeptide()
    lines = []
    for i, (n, r, ri, xyz) in enumerate(zip(ad.atom_names, ad.residue_names, ad.residue_index, ad.coords_nm * 10.0)):
        el = next(c for c in n if c.isalpha())
        lines.append("ATOM  %5d %-4s %3s A%4d    %8.3f%8.3f%8.3f  1.00  0.00          %2s" % (i + 1, n if len(n) == 4 else " " + n, r, ri, *xyz, el))
    pdb = "\n".join(lines + ["TER", "END"]) + "\n"
    for name in ("good", "lonely", "broken"):
        (tmp_path / f"{name}-traj-state0.pdb").write_text(pdb)
    src = os.path.join(GOLDEN, "synthetic_ad-traj-arrays.npz")
    shutil.copy(src, tmp_path / "good-traj-arrays.npz")
    d = dict(np.load(src))
    d["positions"] = d["positions"].copy()
    d["positions"][1] += 1000.0  # a jump no integrator step makes
    np.savez(tmp_path / "broken-traj-arrays.npz", **d)
    (tmp_path / "notes.txt").write_text("ignored")
    data = ds.RawMolDynDataset(data_dir=str(tmp_path), step_width=1)
    assert ds.get_pdb_names(tmp_path) == ["broken", "good", "lonely"] and data.pdb_names == ("broken", "good", "lonely")
    points = list(data.make_iterator(data.pdb_names))
    out = capsys.readouterr().out
    assert "W: lonely data not fully present." in out and "W: broken trajectory has" in out
    want = dl.datapoints_from_trajectory(dl.load_pdb_trace_data("good", ad, src, step_width=1))
    assert len(points) == len(want) > 0 and all(p.name == "good" for p in points)
    for a, b in zip(points, want):
        assert torch.equal(a.atom_coords, b.atom_coords) and torch.equal(a.atom_coord_targets, b.atom_coord_targets)
        assert torch.equal(a.atom_types, b.atom_types) and torch.equal(a.adj_list, b.adj_list)
    batch = dl.moldyn_dense_collate_fn(points[:3])
    assert batch.atom_coords.shape == (3, 22, 3)
    with pytest.raises(RuntimeError, match="unexpected exception"):
        (tmp_path / "bad-traj-state0.pdb").write_text(pdb)
        (tmp_path / "bad-traj-arrays.npz").write_text("not an npz")
        list(ds.RawMolDynDataset(str(tmp_path), 1).make_iterator(["bad"]))
