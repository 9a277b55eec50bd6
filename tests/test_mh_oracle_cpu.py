"""Host-side checks of the MH-loop oracle (oracle/mh_oracle.py, restating utils/evaluation_utils.py:468-745): the
reference's own shape invariant (tests/test_evaluation_utils.py:112-138), the first-accept truncation, the adaptive
number of proposals, and that every stored state is the previous state or the accepted proposal."""
import numpy as np
import pytest
import torch

from oracle import flow_oracle as fo
from oracle import mh_oracle as mo
from tests.common import TINY_O
from timewarp_b200.forcefield import MOLAR_GAS_CONSTANT_R, amber_like_system
from timewarp_b200.peptides import alanine_dipeptide
from timewarp_b200.sampling import compute_num_proposal_steps


def _setup():
    pep = alanine_dipeptide()
    sd = fo.synth_state_dict(TINY_O, 0)
    # local proposals so that some are accepted: small shifts, narrow coordinate prior
    for k in list(sd):
        if ".out_mlp._layers.2." in k:
            sd[k] = sd[k] * 1e-5
    sd["coords_prior_log_scale"] = torch.tensor(float(np.log(5e-4)))
    sd["velocs_prior_log_scale"] = torch.tensor(0.0)
    sysd = amber_like_system(pep).as_float32()
    kbT = 310.0 * MOLAR_GAS_CONSTANT_R
    at = torch.tensor(pep.atom_types)[None]
    x = torch.tensor(pep.coords_nm, dtype=torch.float32)[None]
    mask = torch.zeros(1, pep.num_atoms, dtype=torch.bool)
    return pep, sd, sysd, kbT, at, x, mask


@pytest.mark.parametrize("random_velocs", [True, False])
@pytest.mark.parametrize("S,adaptive", [(1, False), (6, False), (6, True)])
def test_chain_invariants(random_velocs, S, adaptive):
    pep, sd, sysd, kbT, at, x, mask = _setup()
    masses = torch.tensor(pep.masses, dtype=torch.float32)
    trace = []
    coords, velocs, accepted, st = mo.sample_with_model(
        sd, TINY_O, sysd, kbT, at, x, torch.zeros_like(x), mask, masses, 14, mo.TorchDraws(torch.Generator().manual_seed(3)),
        accept=True, random_velocs=random_velocs, resample_velocs=random_velocs, num_proposal_steps=S, adaptive_parallelism=adaptive,
        acceptance_rate_smoothing_factor=0.5 if adaptive else 0.01, trace=trace)
    n = len(st["acceptance"])
    assert len(coords) == n + 1 == len(velocs) and n >= 14  # the reference's invariant
    for k, v in st.items():
        assert len(v) == n, k
    assert 0 <= accepted <= len(trace)
    # every stored state repeats its predecessor unless its indicator says "accepted"
    ind = st["acceptance_indicator"]
    for i in range(n):
        if not ind[i]:
            np.testing.assert_array_equal(coords[i + 1], coords[i])
    pos = 0
    for t in trace:
        k = t["first_acc_idx"] + 1
        assert not ind[pos : pos + k - 1].any()  # everything before the first accepted proposal was rejected
        if t["u"] is not None:
            np.testing.assert_array_equal(ind[pos : pos + k], (t["u"] < t["p_acc"])[:k])
        pos += k
    assert pos == n
    if adaptive:
        assert trace[0]["S"] == 1 or trace[0]["S"] == compute_num_proposal_steps(1e-3, max_num_proposal_steps=S)
        assert all(1 <= t["S"] <= S for t in trace)
        if random_velocs:
            assert len({t["S"] for t in trace}) > 1  # the moving average of the acceptance changed the number of proposals
    else:
        assert all(t["S"] == S for t in trace)
    np.testing.assert_allclose(st["exponent"], st["energies_pot_delta"] + st["energies_kin_delta"] + st["p_xy"] - st["p_yx"], rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(st["acceptance"], np.minimum(1.0, np.exp(-st["exponent"].astype(np.float64))), rtol=1e-4, atol=1e-7)


def test_accept_everything_requires_one_proposal():
    pep, sd, sysd, kbT, at, x, mask = _setup()
    masses = torch.tensor(pep.masses, dtype=torch.float32)
    draws = mo.TorchDraws(torch.Generator().manual_seed(0))
    coords, _, accepted, st = mo.sample_with_model(sd, TINY_O, sysd, kbT, at, x, torch.zeros_like(x), mask, masses, 5, draws, accept=False,
                                                   random_velocs=True, resample_velocs=True, num_proposal_steps=1)
    assert accepted == 5 and len(coords) == 6 and st["acceptance_indicator"].all()
    with pytest.raises(ValueError):
        mo.sample_with_model(sd, TINY_O, sysd, kbT, at, x, torch.zeros_like(x), mask, masses, 5, draws, accept=False, num_proposal_steps=3)


def test_num_proposal_steps_matches_product_host_code():
    for p in (0.0, 1e-3, 0.05, 0.3, 0.9, 0.999, 1.0):
        for mx in (1, 10, 100):
            assert mo.compute_num_proposal_steps(p, max_num_proposal_steps=mx) == compute_num_proposal_steps(p, max_num_proposal_steps=mx)


def test_mh_step_and_explore_step_shapes():
    pep, sd, sysd, kbT, at, x, mask = _setup()
    B = 5
    g = torch.Generator().manual_seed(1)
    xb = x.repeat(B, 1, 1) + 0.005 * torch.randn(B, pep.num_atoms, 3, generator=g)
    draws = mo.TorchDraws(g)
    nc, nv, acc, u, rec = mo.mh_step(sd, TINY_O, sysd, kbT, at.repeat(B, 1), xb, mask.repeat(B, 1), draws)
    assert nc.shape == xb.shape and acc.shape == (B,) and rec.exponent.shape == (B,)
    np.testing.assert_array_equal(nc[~acc].numpy(), xb[~acc].numpy())
    np.testing.assert_array_equal(nc[acc].numpy(), rec.y_coords[acc].numpy())
    e0 = mo.potential_energy_kT(sysd, xb, 1.0)
    y, e, ok, y_new, e_new, v_next = mo.explore_step(sd, TINY_O, sysd, at.repeat(B, 1), xb, torch.randn(B, pep.num_atoms, 3, generator=g), e0,
                                                     mask.repeat(B, 1), draws, threshold=0.0)
    assert ((e - e0) <= 1e-6).all() and v_next.shape == xb.shape
    np.testing.assert_array_equal(y[~ok].numpy(), xb[~ok].numpy())
