"""Potential-energy module with the call surface of the reference's OpenmmPotentialEnergyTorch
(utils/openmm/openmm_bridge.py:252-307): `energy(coords[..., N, 3]) -> [B, 1]` in kJ/mol on the
input's device/dtype, `.kbT`, `.num_particles`, `.get_integrator()`.  The arithmetic runs in the
`tw_peptide_energy` CUDA kernel -- no OpenMM, no host round trip, no per-sample Python loop."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from . import _lib
from .forcefield import MOLAR_GAS_CONSTANT_R, SystemDescription


class _Integrator:
    """Stand-in for the openmm integrator object the reference passes around (only the temperature is used)."""

    def __init__(self, temperature_kelvin: float):
        self._t = float(temperature_kelvin)

    def getTemperature(self) -> float:
        return self._t


class _EnergyFn(torch.autograd.Function):
    """U(x) with the gradient bgflow attaches to the OpenMM energy (openmm_bridge.py:56-60): dU/dx = -force, forces from
    the same kernel launch (analytic, fp64 accumulation)."""

    @staticmethod
    def forward(ctx, module, coords):
        energy, forces = module._evaluate(coords, want_forces=True)
        ctx.save_for_backward(forces)
        ctx.in_shape, ctx.in_dtype = coords.shape, coords.dtype
        return energy

    @staticmethod
    def backward(ctx, grad_out):
        (forces,) = ctx.saved_tensors  # [B,N,3] float32
        g = -forces * grad_out.reshape(-1, 1, 1).to(forces.dtype)
        return None, g.reshape(ctx.in_shape).to(ctx.in_dtype)


class PeptidePotentialEnergy(nn.Module):
    def __init__(self, system: SystemDescription, temperature: Optional[float] = None):
        super().__init__()
        self.system = system
        self.num_particles = system.n_atoms
        self._integrator = _Integrator(system.temperature if temperature is None else temperature)
        self._dev_arrays = {}
        self._struct: Optional[_lib.EnergySystem] = None
        self._struct_device = None

    def get_integrator(self):
        return self._integrator

    @property
    def kbT(self) -> float:
        """T * R in kJ/mol (openmm_bridge.py:299-307); 2.577483411627504 at 310 K."""
        return self._integrator.getTemperature() * MOLAR_GAS_CONSTANT_R

    def _build(self, device):
        s = self.system
        d = {}

        def put(name, arr, dtype):
            t = torch.as_tensor(np.ascontiguousarray(arr), dtype=dtype).to(device).contiguous()
            d[name] = t
            return t.data_ptr() if t.numel() else None

        es = _lib.EnergySystem()
        es.n_atoms = s.n_atoms
        es.n_bonds = len(s.bond_idx)
        es.bond_idx = put("bond_idx", s.bond_idx, torch.int32)
        es.bond_param = put("bond_param", s.bond_param, torch.float32)
        es.n_angles = len(s.angle_idx)
        es.angle_idx = put("angle_idx", s.angle_idx, torch.int32)
        es.angle_param = put("angle_param", s.angle_param, torch.float32)
        es.n_torsions = len(s.torsion_idx)
        es.torsion_idx = put("torsion_idx", s.torsion_idx, torch.int32)
        es.torsion_param = put("torsion_param", s.torsion_param, torch.float32)
        es.charge = put("charge", s.charge, torch.float32)
        es.sigma = put("sigma", s.sigma, torch.float32)
        es.epsilon = put("epsilon", s.epsilon, torch.float32)
        es.excluded = put("excluded", s.excluded, torch.uint8)
        es.n_exceptions = len(s.exception_idx)
        es.exception_idx = put("exception_idx", s.exception_idx, torch.int32)
        es.exception_param = put("exception_param", s.exception_param, torch.float32)
        es.cutoff, es.reaction_field_eps, es.one_4pi_eps0 = s.cutoff, s.reaction_field_eps, s.one_4pi_eps0
        es.use_gb = 1 if s.use_gb else 0
        es.gb_radius = put("gb_radius", s.gb_radius, torch.float32)
        es.gb_scale = put("gb_scale", s.gb_scale, torch.float32)
        es.gb_alpha, es.gb_beta, es.gb_gamma, es.gb_offset = s.gb_alpha, s.gb_beta, s.gb_gamma, s.gb_offset
        es.solute_dielectric, es.solvent_dielectric = s.solute_dielectric, s.solvent_dielectric
        es.surface_area_energy = s.surface_area_energy
        self._dev_arrays, self._struct, self._struct_device = d, es, device

    def _evaluate(self, coords: Tensor, want_forces: bool = False, want_terms: bool = False):
        assert coords.size(-1) == 3, f"last dimension is expected to be of size 3 but it is {coords.size(-1)}"
        assert (
            coords.size(-2) == self.num_particles
        ), f"size {coords.size()} does not align with expected number of particles {self.num_particles}"
        if coords.device.type != "cuda":
            raise _lib.TimewarpB200Error(f"coords are on {coords.device}: the energy kernel runs on CUDA only (no CPU fallback)")
        dev = coords.device
        if self._struct is None or self._struct_device != dev:
            self._build(dev)
        x = coords.detach().reshape(-1, self.num_particles, 3).to(torch.float32).contiguous()
        B = x.shape[0]
        out = torch.empty(B, dtype=torch.float32, device=dev)
        forces = torch.empty(B, self.num_particles, 3, dtype=torch.float32, device=dev) if want_forces else None
        terms = torch.empty(B, 5, dtype=torch.float32, device=dev) if want_terms else None
        _lib.check(
            _lib.load().tw_peptide_energy(C.byref(self._struct), _lib.ptr(x), B, _lib.ptr(out), _lib.ptr(forces), _lib.ptr(terms),
                                          torch.cuda.current_stream(dev).cuda_stream),
            "tw_peptide_energy",
        )
        energy = out.to(coords.dtype).reshape(-1, 1)  # [B,1], input dtype/device (openmm_bridge.py:228)
        if want_terms:
            return energy, forces, terms
        return energy, forces

    def forward(self, coords: Tensor, return_terms: bool = False) -> Tensor:
        """U(coords[..., N, 3]) -> [B, 1] kJ/mol.  Differentiable w.r.t. coords (gradient = -force) like the reference's
        bgflow-bridged energy; without grad the force computation is skipped."""
        if return_terms:
            energy, _, terms = self._evaluate(coords, want_terms=True)
            return energy, terms
        if coords.requires_grad and torch.is_grad_enabled():
            return _EnergyFn.apply(self, coords)
        return self._evaluate(coords)[0]

    def energy_and_forces(self, coords: Tensor):
        """(U [B,1] kJ/mol, F [B,N,3] kJ/mol/nm = -dU/dx) from one kernel launch (OpenMMBridge.evaluate, openmm_bridge.py:170-249)."""
        return self._evaluate(coords, want_forces=True)


class OpenmmPotentialEnergyTorch(PeptidePotentialEnergy):
    """Constructor-compatible alias: OpenmmPotentialEnergyTorch(system, integrator, platform_name, ...)
    (openmm_bridge.py:259-279) where `system` is a SystemDescription and `integrator` anything with
    getTemperature() in kelvin.  platform_* arguments are accepted and ignored (the platform is the GPU)."""

    def __init__(self, system, integrator=None, platform_name: str = "CUDA", platform_properties=None, bridge_kwargs=None):
        t = None
        if integrator is not None:
            t = integrator.getTemperature()
            t = float(getattr(t, "_value", t))
        super().__init__(system, temperature=t)


class OpenMMProvider:
    """Energy modules of several proteins behind one object -- utils/openmm/openmm_provider.py:20-175, same constructor
    arguments, methods and FIFO cache behaviour (`cache_size` modules; the oldest entry is dropped first; `cache_size <= 0`
    disables caching).  `get_system(protein)` finds `<protein>-traj-state0.pdb` under `pdb_dirs` like the reference and builds
    the ff99SB-ILDN + OBC2 `SystemDescription` from the residue table of `timewarp_b200/amber99.py` (the reference calls OpenMM's
    `ForceField.createSystem`; residues outside the table raise).  `losses.compute_energy` and the energy-based losses take
    this object wherever they take the reference's."""

    def __init__(self, pdb_dirs, parameters: str = "T1B-peptides", device="cuda", cache_size: int = 8):
        import os

        if isinstance(pdb_dirs, (str, os.PathLike)):
            pdb_dirs = [pdb_dirs]
        self.pdb_dirs = list(pdb_dirs)
        self.parameters = parameters
        self.device = device if isinstance(device, torch.device) else torch.device(device)
        self.cache_size = cache_size
        self._potential_energy_cache = {}
        self._masses = {}

    def clear_cache(self):
        self._potential_energy_cache.clear()

    def clear_cache_to_size(self, size=None):
        size = self.cache_size if size is None else size
        if size <= 0:
            self.clear_cache()
            return self
        while len(self._potential_energy_cache) > size:
            oldest = next(iter(self._potential_energy_cache))  # dicts keep insertion order: first in, first out
            del self._potential_energy_cache[oldest]
        return self

    def get_integrator(self):
        from .md import get_simulation_environment_integrator

        return get_simulation_environment_integrator(self.parameters)

    @property
    def kbT(self) -> float:
        return 8.31446261815324e-3 * float(self.get_integrator().getTemperature())  # kJ/mol

    def _peptide(self, protein: str):
        import os

        from .dataloader import read_pdb_topology

        for pdb_dir in self.pdb_dirs:
            for dirpath, _, _ in os.walk(str(pdb_dir)):
                candidate = os.path.join(dirpath, f"{protein}-traj-state0.pdb")
                if os.path.isfile(candidate):
                    return read_pdb_topology(candidate, name=protein)
        raise ValueError(f"could not find PDB file for {protein} in any of the provided paths {self.pdb_dirs}")

    def get_system(self, protein: str):
        from .forcefield import amber99sbildn_obc2

        return amber99sbildn_obc2(self._peptide(protein))

    def get_potential_energy_module(self, protein: str):
        if protein in self._potential_energy_cache:
            return self._potential_energy_cache[protein]
        self.clear_cache_to_size(self.cache_size - 1)  # room for one more
        module = OpenmmPotentialEnergyTorch(self.get_system(protein), self.get_integrator(), platform_name="CUDA").to(self.device)
        if self.cache_size > 0:
            self._potential_energy_cache[protein] = module
        return module

    def get_masses(self, protein: str) -> Tensor:
        if protein not in self._masses:
            self._masses[protein] = torch.tensor([float(m) for m in self._peptide(protein).masses], device=self.device)  # (float32 like the reference)
        return self._masses[protein]
