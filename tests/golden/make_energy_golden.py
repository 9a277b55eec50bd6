"""Copies the reference's OpenMM energy / force fixtures for 2olx into ONE small golden file (authoring container only:
reads /root/reference, which does not exist on the GPU box).

    python tests/golden/make_energy_golden.py

* `cpu_*`    : all 40 frames of simulation/testdata/implicit-2olx-traj-cpu-arrays.npz -- the arrays the reference's own test
               simulation/tests/test_md.py:35-47 compares against OpenMM (preset "T1-peptides": amber99sbildn + OBC2,
               CutoffNonPeriodic 2 nm).
* `wide_*`   : every 7th frame of testdata/output/2olx-traj-arrays.npz -- a longer trajectory that leaves the Asn chi1 trans
               rotamer (the torsion series of timewarp_b200/amber99.py::ILDN are exercised over their whole range) and whose
               C-terminal carboxylate improper was written with the two oxygens in the other order.
Positions / forces float32 as stored by the reference, energies float64 (kJ/mol, column 0 = potential)."""
import os

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

cpu = np.load(os.path.join(REF, "simulation/testdata/implicit-2olx-traj-cpu-arrays.npz"))
wide = np.load(os.path.join(REF, "testdata/output/2olx-traj-arrays.npz"))
sel = np.arange(0, wide["positions"].shape[0], 7)
np.savez_compressed(
    os.path.join(HERE, "energy_2olx_openmm.npz"),
    cpu_positions=cpu["positions"], cpu_forces=cpu["forces"], cpu_potential=cpu["energies"][:, 0],
    wide_positions=wide["positions"][sel], wide_forces=wide["forces"][sel], wide_potential=wide["energies"][sel, 0], wide_frames=sel,
)
print("wrote", os.path.join(HERE, "energy_2olx_openmm.npz"))
