"""`RawMolDynDataset` (datasets/iterable_datasets.py:21-129): the directory-of-trajectories dataset the reference's drivers read
their initial states from (sample_trajectory.py:209-216, exploration.py:202-208, sample.py:63): every `<name>-traj-state0.pdb` +
`<name>-traj-arrays.npz` pair in `data_dir` yields conditioning / target `MolDynDatapoint`s `step_width` integrator steps apart.
(The LMDB-backed training datasets need the `lmdb` package, which this environment does not have.)"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from functools import cached_property
from typing import Iterator, Optional, Sequence

from .dataloader import CoordDeltaTooBig, MolDynDatapoint, TrajectoryInformation, datapoints_from_trajectory, load_pdb_trace_data

STATE0_SUFFIX = "-traj-state0.pdb"


def get_pdb_names(data_dir) -> list:
    """datasets/iterable_datasets.py:21-28."""
    return sorted(f[: -len(STATE0_SUFFIX)] for f in os.listdir(data_dir) if f.endswith(STATE0_SUFFIX))


@dataclass(frozen=True)
class RawMolDynDataset:
    data_dir: str
    step_width: int
    equal_data_spacing: bool = field(default=False)

    @cached_property
    def pdb_names(self) -> Sequence[str]:
        names = get_pdb_names(self.data_dir)
        print(f"I: Found {len(names)} trace files in {self.data_dir}.")
        return tuple(names)

    def pdb_file_name(self, pdb_name: str) -> str:
        return f"{self.data_dir}/{pdb_name}{STATE0_SUFFIX}"

    def npz_file_name(self, pdb_name: str) -> str:
        return f"{self.data_dir}/{pdb_name}-traj-arrays.npz"

    def _gracefully_load_pdb_trace_data(self, pdb_name: str) -> Optional[TrajectoryInformation]:
        """A trajectory, or None (with a warning) when its files are missing or it fails the step-distance sanity check;
        anything else is an error (iterable_datasets.py:58-80)."""
        try:
            return load_pdb_trace_data(pdb_name, self.pdb_file_name(pdb_name), self.npz_file_name(pdb_name), step_width=self.step_width,
                                       equal_data_spacing=self.equal_data_spacing)
        except FileNotFoundError:
            print(f"W: {pdb_name} data not fully present.")
        except CoordDeltaTooBig as e:
            print(f"W: {e}.")
        except Exception as e:
            raise RuntimeError(
                f"Got unexpected exception while trying to load {self.pdb_file_name(pdb_name)} / {self.npz_file_name(pdb_name)}") from e
        return None

    def make_iterator(self, pdb_names: Sequence[str]) -> Iterator[MolDynDatapoint]:
        """Datapoints of the named trajectories, in order (iterable_datasets.py:82-129)."""
        for pdb_name in pdb_names:
            info = self._gracefully_load_pdb_trace_data(pdb_name)
            if info is None:
                continue
            yield from datapoints_from_trajectory(info)
