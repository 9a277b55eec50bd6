"""Pin the oracle's autograd (oracle.flow_oracle.nll_loss_and_grads) against gradients produced by
`loss.backward()` on the unmodified reference model (tests/golden/make_golden.py::grad_case)."""
import os

import numpy as np
import pytest
import torch

from oracle import flow_oracle as fo

FULL = fo.OracleConfig()


def load_grad_case(golden_dir, name):
    d = np.load(os.path.join(golden_dir, name + ".npz"))
    g = {k: (torch.from_numpy(d[k]) if d[k].dtype.kind != "U" else d[k]) for k in d.files}
    return g


@pytest.mark.parametrize("name", ["grads_full_ad22", "grads_full_ad22_ragged", "grads_full_ad22_learnable", "grads_full_ad22_chebyshev"])
def test_oracle_autograd_matches_reference(golden_dir, name):
    torch.set_num_threads(max(torch.get_num_threads(), 4))
    g = load_grad_case(golden_dir, name)
    FULL = fo.OracleConfig()
    if name.endswith("learnable"):
        FULL = fo.OracleConfig(attention_type="learnable_kernel")
    if name.endswith("chebyshev"):
        FULL = fo.OracleConfig(attention_type="chebyshev_kernel", cheb_order=12, force_asymptotic_zero=False)
    sd = fo.synth_state_dict(FULL, int(g["weight_seed"]))
    loss, grads = fo.nll_loss_and_grads(sd, FULL, g["atom_types"], g["x_coords"], g["x_velocs"], g["y_coords"], g["y_velocs"],
                                        g["masked_elements"])
    torch.testing.assert_close(loss, g["loss"], rtol=2e-6, atol=2e-6)
    names = [str(n) for n in g["grad_names"]]
    assert set(names) == set(grads)  # the oracle differentiates exactly the reference's Parameters
    norms = torch.tensor([float(grads[n].double().norm()) for n in names], dtype=torch.float64)
    torch.testing.assert_close(norms, g["grad_norms"].double(), rtol=2e-4, atol=1e-7)
    for k in g:
        if k.startswith("grad::"):
            ref = g[k]
            got = grads[k[6:]]
            got = got[:8] if got.numel() > 20000 else got
            err = (got - ref).norm() / ref.norm().clamp_min(1e-12)
            assert err < 2e-4, (k, float(err))
