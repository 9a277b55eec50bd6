"""CPU restatements of the reference's algorithms for the hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and the CPU-baseline / reference legs of bench.py may import this package; the product
(timewarp_b200/) never does and has no CPU fallback.

  flow_oracle.py    the kernel-attention RealNVP flow (torch, any dtype): pinned to golden vectors generated from the unmodified
                    reference (tests/golden/*.npz, tests/golden/make_golden.py), values and autograd gradients
  energy_oracle.py  implicit-solvent Amber potential energy (numpy fp64): PARITY UNPINNED against the reference (OpenMM and its
                    parameter files are absent); pinned by closed-form cases; the pinning test against the reference's golden
                    energies is committed and runs wherever OpenMM exists (tests/test_forcefield_cpu.py)
  md_oracle.py      OpenMM integrator steps / kinetic energy (numpy fp64): LangevinIntegrator pinned by the consecutive steps
                    recorded in the reference's own trajectory fixtures; LangevinMiddleIntegrator parity unpinned
"""
