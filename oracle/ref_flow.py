"""TEST INFRASTRUCTURE -- loader for the byte-compiled, UNMODIFIED reference flow model in oracle/_ref/reference_flow.zip (built by
oracle/build_ref.py).  Import recipe = SURVEY.md Appendix C: stub modules for the plotting / trajectory packages the
reference imports at module top but never uses on this path, stdlib `profile` imported before the reference's own
profile.py could shadow it, and a `__hash__` for one dataclass that Python >= 3.11 rejects as a mutable default."""
from __future__ import annotations

import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(REF_DIR, "reference_flow.zip")
STAMP = os.path.join(REF_DIR, "reference_flow.python_version")


def available() -> bool:
    return os.path.exists(ARCHIVE) and os.path.exists(STAMP) and open(STAMP).read().strip() == sys.version.split()[0]


def load():
    """Returns (custom_transformer_nvp_constructor, CustomAttentionTransformerNVPConfig, CustomAttentionEncoderLayerConfig)."""
    if not available():
        raise ImportError("oracle/_ref is not built for this interpreter (python -m oracle.build_ref in the authoring container)")
    import cProfile  # noqa: F401
    import profile  # noqa: F401

    if ARCHIVE not in sys.path:
        sys.path.insert(0, ARCHIVE)  # zipimport: `timewarp.*` and the top-level `utilities` package the reference imports
    for n in ("pymol2", "mdtraj", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(n, types.ModuleType(n))
    import timewarp.modules.model_wrappers.flow as F

    F.ConditionalFlowDensityConfig.__hash__ = lambda s: id(s)
    from timewarp.model_configs import CustomAttentionTransformerNVPConfig
    from timewarp.model_constructor import custom_transformer_nvp_constructor
    from timewarp.modules.layers.custom_attention_encoder import CustomAttentionEncoderLayerConfig

    return custom_transformer_nvp_constructor, CustomAttentionTransformerNVPConfig, CustomAttentionEncoderLayerConfig


def full_model(state_dict):
    """The flagship configuration (configs/kernel_transformer_nvp.yaml:19-30) built by the reference's constructor and
    loaded with `state_dict` (strict)."""
    ctor, NVPConfig, EncConfig = load()
    enc = EncConfig(d_model=128, dim_feedforward=2048, dropout=0.0, num_heads=6, attention_type="kernel",
                    lengthscales=[0.1, 0.2, 0.5, 0.7, 1.0, 1.2], normalise_kernel_values=True)
    model = ctor(NVPConfig(atom_embedding_dim=32, latent_mlp_hidden_dims=[256], num_coupling_layers=8, num_transformer_layers=3,
                           encoder_layer_config=enc))
    model.load_state_dict(state_dict, strict=True)
    return model.eval()
