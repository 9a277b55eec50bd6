"""CPU-only checks: the C-ABI library loads and exports every symbol include/timewarp_b200.h declares,
host logic (config validation, parameter table, state_dict contract, seeded init) -- no compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import timewarp_b200 as tw
from timewarp_b200 import _lib
from timewarp_b200.build import build_library
from oracle import flow_oracle as fo
from tests.common import FULL_O, TINY_O, GOLDEN, model_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build_library()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "timewarp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(tw_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.exported_symbols()) == declared
    assert lib.tw_abi_version() == 1


def test_struct_layout_matches_c(lib):
    # sizes computed by the C compiler for the same declarations
    import subprocess, tempfile, textwrap
    src = textwrap.dedent("""
        #include <stdio.h>
        #include "timewarp_b200.h"
        int main(){ printf("%zu %zu\\n", sizeof(tw_flow_config), sizeof(tw_energy_system)); return 0; }
    """)
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        a, b = map(int, subprocess.check_output([os.path.join(d, "s")]).split())
    assert a == C.sizeof(_lib.FlowConfig) and b == C.sizeof(_lib.EnergySystem)


def test_param_table_and_state_dict_keys(lib):
    for o in (TINY_O, FULL_O):
        m = tw.custom_transformer_nvp_constructor(model_config(o, "fp32"))
        sd = m.state_dict()
        assert list(sd.keys()) == list(fo.state_dict_shapes(o).keys())
        for k, shape in fo.state_dict_shapes(o).items():
            assert tuple(sd[k].shape) == shape, k
        assert lib.tw_flow_num_params(C.byref(m._cfg)) == len(m._ordered_params())
    assert sum(p.numel() for p in m.parameters()) == 35_971_282


def test_seeded_init_matches_reference():
    ref = np.load(os.path.join(GOLDEN, "tiny_init_seed0.npz"))
    torch.manual_seed(0)
    m = tw.custom_transformer_nvp_constructor(model_config(TINY_O, "fp32"))
    sd = m.state_dict()
    assert sorted(sd.keys()) == sorted(ref.files)
    for k in ref.files:
        assert np.array_equal(sd[k].numpy(), ref[k]), k


def test_seeded_init_matches_reference_local():
    """`local` attention module tree (qkv_proj before output_proj, local_self_attention.py:33-41)."""
    from tests.common import TINY_LOC
    ref = np.load(os.path.join(GOLDEN, "tiny_local_init_seed0.npz"))
    torch.manual_seed(0)
    sd = tw.custom_transformer_nvp_constructor(model_config(TINY_LOC, "fp32")).state_dict()
    assert sorted(sd.keys()) == sorted(ref.files)
    for k in ref.files:
        assert np.array_equal(sd[k].numpy(), ref[k]), k


def test_reference_checkpoint_loads(tmp_path):
    """utilities/model_utils.py:32-63: a checkpoint written by the reference's own save_model (tests/golden/ckpt, made by
    make_golden.py::checkpoint_case) loads verbatim, from the file or from a run directory; DeepSpeed-style "module" key too."""
    from oracle import flow_oracle as fo
    from timewarp_b200 import checkpoint as ck

    def ctor(data):
        assert "module" in data or data["step"] == 123  # extra keys of save_model(**kwargs) are handed to the constructor
        return tw.custom_transformer_nvp_constructor(model_config(TINY_O, "fp32"))

    want = fo.synth_state_dict(TINY_O, 0)
    for path in (os.path.join(GOLDEN, "ckpt"), os.path.join(GOLDEN, "ckpt", "run0", "best_model.pt")):
        m = ck.load_model(path, ctor, weights_only=True)
        sd = m.state_dict()
        assert set(sd) == set(want)
        for k in want:
            assert torch.equal(sd[k], want[k]), k
    ck.save_model(tmp_path / "best_model.pt", m, step=123)  # round trip through our writer
    assert set(ck.load_model_state_dict(tmp_path)) == set(want)
    # optimizer state survives exactly (reference tests/test_training_utils.py:23-67)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    for q in m.parameters():
        q.grad = torch.full_like(q, 0.01)
    opt.step()
    os.makedirs(tmp_path / "opt")
    ck.save_model(tmp_path / "opt" / "best_model.pt", m, optimizer=opt)
    data = ck.load_checkpoint_in_subdir(tmp_path / "opt")
    opt2 = torch.optim.Adam(ctor({"module": None}).parameters(), lr=1e-3)
    opt2.load_state_dict(data["optimizer_state_dict"])
    for a, b in zip(opt.state_dict()["state"].values(), opt2.state_dict()["state"].values()):
        assert torch.equal(a["exp_avg"], b["exp_avg"]) and torch.equal(a["exp_avg_sq"], b["exp_avg_sq"])
    for k, v in m.state_dict().items():
        assert torch.equal(data["model_state_dict"][k], v)
    torch.save({"module": m.state_dict()}, tmp_path / "ds.pt")
    m2 = ck.load_model(tmp_path / "ds.pt", ctor)
    assert torch.equal(m2.state_dict()["flow.atom_embedder.weight"], m.state_dict()["flow.atom_embedder.weight"])
    os.makedirs(tmp_path / "a"), os.makedirs(tmp_path / "b")
    for d in "ab":
        torch.save({}, tmp_path / d / "x.pt")
    with pytest.raises(AssertionError):
        ck.load_checkpoint_in_subdir(tmp_path, "x.pt")  # two candidates: unique_item fails like the reference


def test_synthetic_weights_match_oracle_generator():
    """bench.py loads the product model from timewarp_b200.synthetic and its CPU baseline from the oracle's generator: the two
    must give identical parameters (every attention variant, two seeds)."""
    from oracle import flow_oracle as fo
    from tests.common import TINY_C, TINY_L, TINY_LOC
    from timewarp_b200.synthetic import synth_state_dict

    for o in (TINY_O, TINY_L, TINY_C, TINY_LOC):
        m = tw.custom_transformer_nvp_constructor(model_config(o, "fp32"))
        for seed in (0, 3):
            a, b = synth_state_dict(m, seed), fo.synth_state_dict(o, seed)
            assert set(a) == set(b)
            for k in a:
                assert torch.equal(a[k], b[k]), (o.attention_type, seed, k)
            m.load_state_dict(a, strict=True)


def test_config_validation(lib):
    bad = model_config(TINY_O, "fp32")
    bad.num_coupling_layers = 3
    with pytest.raises(AssertionError, match="even number of coupling layers"):
        tw.custom_transformer_nvp_constructor(bad)
    bad = model_config(TINY_O, "fp32")
    bad.position_layer_index_mod_2 = 2
    with pytest.raises(AssertionError):
        tw.custom_transformer_nvp_constructor(bad)
    bad = model_config(TINY_O, "fp32")
    bad.encoder_layer_config.attention_type = "global"
    with pytest.raises(RuntimeError, match="Unknown attention type"):  # custom_attention_encoder.py:211-212
        tw.custom_transformer_nvp_constructor(bad)
    # tensor-core precisions reject layer sizes they do not cover (status TW_ERR_UNSUPPORTED)
    m = tw.custom_transformer_nvp_constructor(model_config(TINY_O, "bf16x3"))
    n = C.c_size_t()
    assert lib.tw_flow_workspace_bytes(C.byref(m._cfg), 4, 4, 22, C.byref(n)) == 4
    assert b"tensor-core" in lib.tw_last_error()
    # NULL pointers are reported, not dereferenced
    m = tw.custom_transformer_nvp_constructor(model_config(TINY_O, "fp32"))
    assert lib.tw_flow_log_likelihood(C.byref(m._cfg), None, None, None, None, None, None, None, 1, 3, 1, None, None, None, None, None, 0, None) == 1
    assert lib.tw_attn_scores(None, None, None, 1, 3, 2, None, None) == 1
    assert lib.tw_attn_scores(None, None, None, 0, 3, 2, None, None) == 0  # empty batch is a no-op


def test_no_cpu_fallback():
    m = tw.custom_transformer_nvp_constructor(model_config(TINY_O, "fp32"))
    B, V = 2, 5
    kw = dict(atom_types=torch.zeros(B, V, dtype=torch.long), x_coords=torch.zeros(B, V, 3), x_velocs=torch.zeros(B, V, 3),
              adj_list=torch.zeros(0, 2, dtype=torch.long), edge_batch_idx=torch.zeros(0, dtype=torch.long),
              masked_elements=torch.zeros(B, V, dtype=torch.bool))
    with pytest.raises(_lib.TimewarpB200Error, match="no CPU fallback"):
        m.log_likelihood(y_coords=torch.zeros(B, V, 3), y_velocs=torch.zeros(B, V, 3), **kw)
    with pytest.raises(_lib.TimewarpB200Error, match="no CPU fallback"):
        m.conditional_sample(num_samples=1, **kw)
    with pytest.raises(NotImplementedError):
        m.flow.chain[0].scale_transformer.in_mlp(torch.zeros(1, 17))
    from timewarp_b200.energy import PeptidePotentialEnergy
    from timewarp_b200.forcefield import amber_like_system
    from timewarp_b200.peptides import alanine_dipeptide
    e = PeptidePotentialEnergy(amber_like_system(alanine_dipeptide()))
    assert abs(e.kbT - 2.577483411627504) < 1e-12  # utils/evaluation_utils_o2.py:17
    with pytest.raises(_lib.TimewarpB200Error, match="no CPU fallback"):
        e(torch.zeros(1, 22, 3))
    with pytest.raises(AssertionError):
        e(torch.zeros(1, 21, 3))


def test_flat_adam_host_side():
    """optim.FlatAdam without the CUDA backward: the reference-shaped constructor (utilities/training_utils.py:356-368), a loud
    error instead of a silent torch fallback, and its state dict."""
    from types import SimpleNamespace
    from timewarp_b200.optim import FlatAdam, get_optimizer
    m = tw.custom_transformer_nvp_constructor(model_config(TINY_O, "bf16x3"))
    opt = get_optimizer(m, SimpleNamespace(optimizer="Adam", learning_rate=3e-4, weight_decay=1e-2, warmup_steps=0))
    assert isinstance(opt, FlatAdam) and isinstance(opt, torch.optim.Optimizer)
    g = opt.param_groups[0]
    assert (g["lr"], g["weight_decay"], g["betas"], g["eps"]) == (3e-4, 1e-2, (0.9, 0.999), 1e-8)
    assert len(g["params"]) == sum(1 for p in m.parameters() if p.requires_grad)
    with pytest.raises(_lib.TimewarpB200Error, match="no gradient to apply"):
        opt.step()
    for q in m.parameters():
        q.grad = torch.zeros_like(q)
    with pytest.raises(_lib.TimewarpB200Error, match="no CPU fallback"):
        opt.step()
    sd = opt.state_dict()
    assert sd["flat"] and sd["exp_avg"] is None and sd["hyper"]["lr"] == 3e-4
    opt.load_state_dict({**sd, "hyper": {**sd["hyper"], "lr": 1e-5}})
    assert opt.param_groups[0]["lr"] == 1e-5
    with pytest.raises(ValueError):
        opt.load_state_dict(torch.optim.Adam(m.parameters()).state_dict())
    opt.zero_grad(set_to_none=True)
    # checkpoints move between the two flat layouts parameter by parameter
    src = torch.arange(12.0)
    dst = torch.zeros(16)
    FlatAdam._remap(src, [0, 4, 8], dst, [8, 0, 12], [3, 4, 2])
    assert dst.tolist() == [4.0, 5.0, 6.0, 7.0, 0, 0, 0, 0, 0.0, 1.0, 2.0, 0, 8.0, 9.0, 0, 0]
    with pytest.raises(ValueError, match="different parameters"):
        opt.load_state_dict({**sd, "exp_avg": torch.zeros(4), "exp_avg_sq": torch.zeros(4), "offsets": [0], "numels": [4], "step": 1.0})


def test_chirality_centers_host():
    from timewarp_b200.chirality import find_chirality_centers
    g = np.load(os.path.join(GOLDEN, "chirality_2olx.npz"))
    c = find_chirality_centers(torch.from_numpy(g["bonds"]), torch.from_numpy(g["atom_types"]))
    assert c.tolist() == g["centers"].tolist()


def test_num_proposal_steps_and_chainstats(tmp_path):
    from timewarp_b200.sampling import ChainStats, compute_num_proposal_steps
    # utils/evaluation_utils.py:32-64
    assert compute_num_proposal_steps(1e-3, max_num_proposal_steps=100) == 100
    assert compute_num_proposal_steps(0.5) == 4  # ceil(log(0.1)/log(0.5))
    assert compute_num_proposal_steps(1.0) == 1
    assert compute_num_proposal_steps(0.0, max_num_proposal_steps=7) == 7
    a = np.arange(10.0)
    s = ChainStats(*(a.copy() for _ in range(9)))
    assert len(s) == 10 and len(s.thin(3)) == 4 and len(s[2:5]) == 3
    s.save(tmp_path / "c.pkl")
    assert np.array_equal(ChainStats.load(tmp_path / "c.pkl").p_xy, a)
