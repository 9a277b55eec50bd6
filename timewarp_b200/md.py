"""OpenMM-shaped integrator objects and `openmm_step` on the GPU.

The reference's MH driver can interleave OpenMM MD steps with the flow proposals (`openmm_on_current` / `openmm_on_proposal`,
utils/evaluation_utils.py:559-565,594-602,623-626) through `openmm_step(sim, coords, velocs, num_steps, integrator)`
(utils/evaluation_utils.py:439-464), which drives an `openmm.app.Simulation` built by simulation/md.py:100-125,190-216.
Here `Simulation` holds a `SystemDescription` + an integrator description and `step` runs the `tw_langevin_steps` CUDA kernel:
all `num_steps` force evaluations and updates of a whole batch of conformations in ONE launch, positions / velocities /
forces resident in shared memory -- no host round trip per step, no per-sample context.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _lib
from .energy import PeptidePotentialEnergy
from .forcefield import MOLAR_GAS_CONSTANT_R, SystemDescription


class LangevinIntegrator:
    """mm.LangevinIntegrator(temperature [K], friction [1/ps], timestep [ps]) -- simulation/md.py:113-114."""

    kind = _lib.TW_INTEGRATOR_LANGEVIN

    def __init__(self, temperature: float, friction: float, timestep: float):
        self._temperature, self._friction, self._timestep = (float(getattr(q, "_value", q)) for q in (temperature, friction, timestep))

    def getTemperature(self) -> float:
        return self._temperature

    def getFriction(self) -> float:
        return self._friction

    def getStepSize(self) -> float:
        return self._timestep


class LangevinMiddleIntegrator(LangevinIntegrator):
    """mm.LangevinMiddleIntegrator -- simulation/md.py:115-123."""

    kind = _lib.TW_INTEGRATOR_LANGEVIN_MIDDLE


def get_parameters_from_preset(preset_or_dataset_name):
    """simulation/md.py:13-97: dataset / preset name -> simulation parameters (temperature K, friction 1/ps, timestep ps)."""
    if isinstance(preset_or_dataset_name, dict):
        return preset_or_dataset_name
    old = {"T1-peptides", "HP-1400", "HP-4000", "alanine-dipeptide", "amber99-implicit-old"}
    name = {"T1B-peptides": "amber14-implicit"}.get(preset_or_dataset_name, preset_or_dataset_name)
    if name in old:
        return {"forcefield": "amber99-implicit", "temperature": 310.0, "friction": 0.3, "timestep": 0.0005, "integrator": "LangevinIntegrator"}
    if name in ("amber99-implicit", "amber14-implicit", "amber14-explicit"):
        return {"forcefield": name, "temperature": 310.0, "friction": 0.3, "timestep": 0.0005, "waterbox_pad": 1.0,
                "integrator": "LangevinMiddleIntegrator"}
    raise ValueError("Invalid preset name '%s'" % name)


def get_simulation_environment_integrator(parameters):
    """simulation/md.py:100-125."""
    p = get_parameters_from_preset(parameters)
    cls = {"LangevinIntegrator": LangevinIntegrator, "LangevinMiddleIntegrator": LangevinMiddleIntegrator}[p["integrator"]]
    return cls(p["temperature"], p["friction"], p["timestep"])


class Simulation:
    """The part of `openmm.app.Simulation` the MH driver touches (`context` state in, `step(n)`, state out), batched.
    `system` is a SystemDescription (force-field arrays + masses), `integrator` one of the classes above."""

    def __init__(self, system: SystemDescription, integrator: LangevinIntegrator, seed: int = 0):
        self.system, self.integrator = system, integrator
        self._energy = PeptidePotentialEnergy(system, temperature=integrator.getTemperature())
        self._masses = {}
        self._seed, self._offset = int(seed), 0

    @property
    def kbT(self) -> float:
        return self.integrator.getTemperature() * MOLAR_GAS_CONSTANT_R

    def masses(self, device) -> Tensor:
        if device not in self._masses:
            self._masses[device] = torch.as_tensor(np.asarray(self.system.masses), dtype=torch.float32).to(device).contiguous()
        return self._masses[device]

    def velocities_to_temperature(self, coords: Tensor) -> Tensor:
        """context.setVelocitiesToTemperature: v ~ N(0, kT/m) per component (drawn from torch's generator of the device)."""
        std = torch.sqrt(self.kbT / self.masses(coords.device))[:, None]
        return torch.randn_like(coords, dtype=torch.float32) * std

    def step(self, coords: Tensor, velocs: Tensor, num_steps: int, noise: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """`num_steps` integrator steps of every conformation in coords [..., N, 3] (nm) / velocs (nm/ps); returns new tensors.
        noise: optional standard normals [num_steps, B, N, 3]; default = in-kernel Philox stream (seed of the Simulation,
        advancing with every call)."""
        if coords.device.type != "cuda":
            raise _lib.TimewarpB200Error(f"coords are on {coords.device}: the integrator kernel runs on CUDA only (no CPU fallback)")
        N = self.system.n_atoms
        assert coords.shape[-2:] == (N, 3) and velocs.shape == coords.shape
        dev = coords.device
        en = self._energy
        if en._struct is None or en._struct_device != dev:
            en._build(dev)
        x = coords.detach().reshape(-1, N, 3).to(torch.float32).contiguous().clone()
        v = velocs.detach().reshape(-1, N, 3).to(torch.float32).contiguous().clone()
        B = x.shape[0]
        if noise is not None:
            noise = noise.to(device=dev, dtype=torch.float32).contiguous()
            assert noise.shape == (num_steps, B, N, 3), f"noise must be [num_steps, B, N, 3], got {tuple(noise.shape)}"
        it = self.integrator
        _lib.check(
            _lib.load().tw_langevin_steps(C.byref(en._struct), _lib.ptr(x), _lib.ptr(v), _lib.ptr(self.masses(dev)), B, int(num_steps),
                                          it.kind, it.getStepSize(), it.getFriction(), self.kbT, _lib.ptr(noise), self._seed,
                                          self._offset, torch.cuda.current_stream(dev).cuda_stream),
            "tw_langevin_steps",
        )
        if noise is None:
            # a thread draws ceil(3N / 128) normals per step, one 32-bit Philox output each (Box-Muller pairs)
            self._offset += int(num_steps) * (-(-3 * N // 128)) + 2
        return x.reshape(coords.shape).to(coords.dtype), v.reshape(coords.shape).to(coords.dtype)


def openmm_step(sim: Simulation, coords: Tensor, velocs: Optional[Tensor] = None, num_steps: int = 1, integrator=None) -> Tuple[Tensor, Tensor]:
    """utils/evaluation_utils.py:439-464 (same signature): set positions (+ velocities, or velocities to the integrator's
    temperature), `sim.step(num_steps)`, read positions and velocities back.  The reference handles one conformation
    (`coords.squeeze(0)`); here every leading row is an independent conformation."""
    if velocs is None:
        if integrator is None:
            raise ValueError("either `velocs` or `integrator` needs to be specified")
        velocs = sim.velocities_to_temperature(coords)
    return sim.step(coords, velocs, num_steps)
