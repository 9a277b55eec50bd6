"""Golden vectors for the batch augmentation, made by the UNMODIFIED reference's `transform_batch`
(equivariance/equivariance_transforms.py:153-175) with seeded generators.  Authoring container only:

    python tests/golden/make_equivariance_golden.py   ->   tests/golden/equivariance_batch.npz
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
LINK_DIR = "/tmp/tw_ref_pkg"
os.makedirs(LINK_DIR, exist_ok=True)
if not os.path.exists(os.path.join(LINK_DIR, "timewarp")):
    os.symlink(REF, os.path.join(LINK_DIR, "timewarp"))
sys.path.insert(0, LINK_DIR)
for n in ("pymol2", "mdtraj", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(n, types.ModuleType(n))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from timewarp.dataloader import DenseMolDynBatch  # noqa: E402
from timewarp.equivariance.equivariance_transforms import transform_batch  # noqa: E402

g = torch.Generator().manual_seed(11)
B, V = 3, 7
f = lambda: torch.randn(B, V, 3, generator=g)  # noqa: E731
fields = dict(atom_coords=f(), atom_velocs=f(), atom_forces=f(), atom_coord_targets=f(), atom_veloc_targets=f(), atom_force_targets=f())
batch = DenseMolDynBatch(names=["a", "b", "c"], atom_types=torch.randint(0, 5, (B, V), generator=g),
                         adj_list=torch.tensor([[0, 1], [1, 2], [7, 8]]), edge_batch_idx=torch.tensor([0, 0, 1]),
                         masked_elements=torch.zeros(B, V, dtype=torch.bool), **fields)
np.random.seed(5)
torch.manual_seed(6)
out = transform_batch(batch)
np.savez(os.path.join(HERE, "equivariance_batch.npz"), **{"in_" + k: v.numpy() for k, v in fields.items()},
         in_atom_types=batch.atom_types.numpy(), **{"out_" + k: getattr(out, k).numpy() for k in fields},
         out_atom_types=out.atom_types.numpy(), out_adj_list=out.adj_list.numpy())
print("ok")
