"""Generate golden vectors by running the UNMODIFIED reference (imported from /root/reference).

Run in the authoring container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  Inputs are synthetic; weights come from
`oracle.flow_oracle.synth_state_dict` (a pure function of key/shape/seed) loaded into the
reference model with `load_state_dict(strict=True)`, so nothing but small arrays is stored.

Import recipe = SURVEY.md Appendix C (stub modules for absent plotting deps, one dataclass
__hash__ shim for Python >= 3.11; no reference code is modified or copied).
"""
import os
import sys
import types
import profile, cProfile  # noqa: F401,E401  (import stdlib `profile` before reference/profile.py can shadow it)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REF = "/root/reference"
LINK_DIR = "/tmp/tw_ref_pkg"
os.makedirs(LINK_DIR, exist_ok=True)
if not os.path.exists(os.path.join(LINK_DIR, "timewarp")):
    os.symlink(REF, os.path.join(LINK_DIR, "timewarp"))
sys.path.insert(0, LINK_DIR)
sys.path.append(REF)
for n in ("pymol2", "mdtraj", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(n, types.ModuleType(n))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import timewarp.modules.model_wrappers.flow as F  # noqa: E402

F.ConditionalFlowDensityConfig.__hash__ = lambda s: id(s)
from timewarp.model_constructor import custom_transformer_nvp_constructor  # noqa: E402
from timewarp.model_configs import CustomAttentionTransformerNVPConfig  # noqa: E402
from timewarp.modules.layers.custom_attention_encoder import CustomAttentionEncoderLayerConfig  # noqa: E402
from timewarp.modules.layers.kernel_attention import compute_kernel_attention_scores  # noqa: E402
from timewarp.utils import chirality as ref_chirality  # noqa: E402

from oracle.flow_oracle import OracleConfig, synth_state_dict  # noqa: E402
from timewarp_b200.peptides import alanine_dipeptide, tetrapeptide_2olx  # noqa: E402


def ref_model(cfg: OracleConfig, seed: int):
    enc = CustomAttentionEncoderLayerConfig(
        d_model=cfg.d_model,
        dim_feedforward=cfg.dim_feedforward,
        dropout=0.0,
        num_heads=cfg.num_heads if cfg.attention_type == "local" else len(cfg.lengthscales),
        attention_type=cfg.attention_type,
        lengthscales=None if cfg.attention_type == "local" else list(cfg.lengthscales),
        normalise_kernel_values=None if cfg.attention_type == "local" else True,
        max_radius=cfg.max_radius if cfg.attention_type == "local" else None,
        cheb_order=cfg.cheb_order if cfg.attention_type == "chebyshev_kernel" else None,
        force_asymptotic_zero=cfg.force_asymptotic_zero if cfg.attention_type == "chebyshev_kernel" else None,
    )
    mc = CustomAttentionTransformerNVPConfig(
        atom_embedding_dim=cfg.atom_embedding_dim,
        latent_mlp_hidden_dims=list(cfg.latent_mlp_hidden_dims),
        num_coupling_layers=cfg.num_coupling_layers,
        num_transformer_layers=cfg.num_transformer_layers,
        encoder_layer_config=enc,
        position_layer_index_mod_2=cfg.position_layer_index_mod_2,
    )
    model = custom_transformer_nvp_constructor(mc)
    sd = synth_state_dict(cfg, seed)
    if cfg.attention_type == "chebyshev_kernel":
        # the reference builds `cheb_coeffs` as an EXPANDED tensor (stride 0 over heads, kernel_attention.py:328-330): the
        # heads share storage and load_state_dict cannot copy into it.  Give every module a real [H, order] tensor instead.
        named = dict(model.named_parameters())
        for k in [k for k in sd if k.endswith("cheb_coeffs")]:
            named[k].data = sd[k].clone()
        missing = model.load_state_dict({k: v for k, v in sd.items() if not k.endswith("cheb_coeffs")}, strict=False)
        assert all(k.endswith("cheb_coeffs") for k in missing.missing_keys) and not missing.unexpected_keys
    else:
        model.load_state_dict(sd, strict=True)
    model.eval()
    return model, sd


def synth_batch(peptide, B, seed, lengths=None):
    """x = pdb + N(0, 0.01^2), y = x + N(0, 0.02^2) nm, velocities N(0,1); ragged -> zero padded."""
    g = torch.Generator().manual_seed(seed)
    V = peptide.num_atoms
    base = torch.tensor(peptide.coords_nm, dtype=torch.float32)
    x = base[None] + 0.01 * torch.randn(B, V, 3, generator=g)
    y = x + 0.02 * torch.randn(B, V, 3, generator=g)
    xv = torch.randn(B, V, 3, generator=g)
    yv = torch.randn(B, V, 3, generator=g)
    at = torch.tensor(peptide.atom_types)[None].repeat(B, 1)
    mask = torch.zeros(B, V, dtype=torch.bool)
    if lengths is not None:
        for b, n in enumerate(lengths):
            mask[b, n:] = True
        keep = (~mask)[:, :, None]
        x, y, xv, yv = x * keep, y * keep, xv * keep, yv * keep
        at = at * (~mask)
    return at, x, xv, y, yv, mask


EMPTY_ADJ = torch.zeros(0, 2, dtype=torch.long)
EMPTY_EBI = torch.zeros(0, dtype=torch.long)


def run_case(name, cfg, peptide, B, seed, lengths=None, sample_S=3, wseed=0, trace_layer0=True):
    model, sd = ref_model(cfg, wseed)
    at, x, xv, y, yv, mask = synth_batch(peptide, B, seed, lengths)
    out = dict(atom_types=at.numpy(), x_coords=x.numpy(), x_velocs=xv.numpy(), y_coords=y.numpy(), y_velocs=yv.numpy(),
               masked_elements=mask.numpy(), weight_seed=np.int64(wseed))
    kw = dict(atom_types=at, x_coords=x, x_velocs=xv, adj_list=EMPTY_ADJ, edge_batch_idx=EMPTY_EBI, masked_elements=mask)
    with torch.no_grad():
        out["log_likelihood"] = model.log_likelihood(y_coords=y, y_velocs=yv, **kw).numpy()
        out["loss"] = model(y_coords=y, y_velocs=yv, **kw).numpy()
        # fp64 truth from the same reference code
        m64 = ref_model(cfg, wseed)[0].double()  # (a deepcopy would keep the Chebyshev basis lambdas bound to the fp32 module)
        kw64 = dict(atom_types=at, x_coords=x.double(), x_velocs=xv.double(), adj_list=EMPTY_ADJ,
                    edge_batch_idx=EMPTY_EBI, masked_elements=mask)
        out["log_likelihood_f64"] = m64.log_likelihood(y_coords=y.double(), y_velocs=yv.double(), **kw64).numpy()
        # attention scores straight from the reference function
        com = (x * (~mask)[:, :, None]).sum(1, keepdim=True) / (~mask).sum(1)[:, None, None]
        ls = torch.tensor(cfg.lengthscales, dtype=torch.float32)
        if cfg.attention_type == "learnable_kernel":  # density direction: the first attention layer executed (cache quirk)
            ls = torch.exp(sd["flow.chain.0.scale_transformer.encoder_layers.0.self_attn.attention.log_lengthscales"])
        if cfg.attention_type == "local":
            out["scores"] = np.zeros((0,), np.float32)  # dot-product attention: no position-only score tensor
        elif cfg.attention_type == "chebyshev_kernel":  # scores of the first attention layer (every layer has its own)
            from timewarp.modules.layers.kernel_attention import chebyshev_basis_function
            cc = sd["flow.chain.0.scale_transformer.encoder_layers.0.self_attn.attention.cheb_coeffs"]
            out["scores"] = compute_kernel_attention_scores(
                query=x - com, key=x - com, masked_elements=mask, lengthscales=ls,
                basis_function=lambda s_: chebyshev_basis_function(s_, cfg.cheb_order, cc, cfg.force_asymptotic_zero)).numpy()
        else:
          out["scores"] = compute_kernel_attention_scores(query=x - com, key=x - com, masked_elements=mask, lengthscales=ls).numpy()

        # sampling, S=1 over the whole batch (exploration.py shape) -- same RNG consumption as the reference
        torch.manual_seed(1234 + seed)
        yc1, yv1, lp1 = model.conditional_sample_with_logp(num_samples=1, **kw)
        torch.manual_seed(1234 + seed)
        xc_c = x - com
        zc1 = torch.distributions.Normal(torch.zeros_like(xc_c), torch.exp(model.coords_prior_log_scale)).rsample((1,))
        zv1 = torch.distributions.Normal(torch.zeros_like(xv), torch.exp(model.velocs_prior_log_scale)).rsample((1,))
        out.update(s1_z_coords=zc1.numpy(), s1_z_velocs=zv1.numpy(), s1_y_coords=yc1.numpy(), s1_y_velocs=yv1.numpy(), s1_logp=lp1.numpy())

        # sampling, S proposals from one state (sample_with_model shape, B == 1)
        kw1 = dict(atom_types=at[:1], x_coords=x[:1], x_velocs=xv[:1], adj_list=EMPTY_ADJ, edge_batch_idx=EMPTY_EBI,
                   masked_elements=mask[:1])
        torch.manual_seed(4321 + seed)
        ycS, yvS, lpS = model.conditional_sample_with_logp(num_samples=sample_S, **kw1)
        torch.manual_seed(4321 + seed)
        zcS = torch.distributions.Normal(torch.zeros_like(x[:1]), torch.exp(model.coords_prior_log_scale)).rsample((sample_S,))
        zvS = torch.distributions.Normal(torch.zeros_like(xv[:1]), torch.exp(model.velocs_prior_log_scale)).rsample((sample_S,))
        out.update(sS_z_coords=zcS.numpy(), sS_z_velocs=zvS.numpy(), sS_y_coords=ycS.numpy(), sS_y_velocs=yvS.numpy(), sS_logp=lpS.numpy())
        # MH reverse move density (evaluation_utils.py:648-657) for those proposals
        S = sample_S
        p_yx = model.log_likelihood(
            atom_types=at[:1].repeat(S, 1), y_coords=x[:1].repeat(S, 1, 1), y_velocs=xv[:1].repeat(S, 1, 1),
            x_coords=ycS.squeeze(1), x_velocs=yvS.squeeze(1), adj_list=EMPTY_ADJ, edge_batch_idx=EMPTY_EBI,
            masked_elements=mask[:1].repeat(S, 1))
        out["sS_p_yx"] = p_yx.numpy()

        if trace_layer0:
            layer0 = model.flow.chain[0]
            feats = model.flow.atom_embedder(at)
            cache = model.cache.empty_like()
            scale, shift = layer0._get_scale_and_shift(
                atom_types=at, z_coords=y - x, z_velocs=yv, x_features=feats, x_coords=x - com, x_velocs=xv,
                adj_list=EMPTY_ADJ, edge_batch_idx=EMPTY_EBI, masked_elements=mask, logger=None, cache=cache)
            out["layer0_scale"] = scale.numpy()
            out["layer0_shift"] = shift.numpy()
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(name, "ll", out["log_likelihood"], "f64", out["log_likelihood_f64"], "loss", out["loss"])
    return sd


GRAD_FULL_KEYS = [
    "flow.atom_embedder.weight", "coords_prior_log_scale", "velocs_prior_log_scale",
    "flow.chain.0.scale_transformer.in_mlp._layers.0.weight", "flow.chain.0.scale_transformer.in_mlp._layers.2.bias",
    "flow.chain.3.shift_transformer.encoder_layers.1.self_attn.values_proj.weight",
    "flow.chain.3.shift_transformer.encoder_layers.1.self_attn.attention._out_projection.weight",
    "flow.chain.5.scale_transformer.encoder_layers.2.linear1.bias", "flow.chain.5.scale_transformer.encoder_layers.2.linear2.bias",
    "flow.chain.5.scale_transformer.encoder_layers.0.norm1.weight", "flow.chain.5.scale_transformer.encoder_layers.0.norm2.bias",
    "flow.chain.7.shift_transformer.out_mlp._layers.2.weight", "flow.chain.7.shift_transformer.out_mlp._layers.0.bias",
    "flow.chain.6.shift_transformer.encoder_layers.0.linear2.weight",
]


def grad_case(name, cfg, peptide, B, seed, lengths=None, wseed=0, extra_keys=()):
    """NLL loss (density_model_base.py:27-42) and its autograd gradients from the unmodified reference:
    L2 norm of every parameter gradient + a few full tensors."""
    model, sd = ref_model(cfg, wseed)
    model.train()
    at, x, xv, y, yv, mask = synth_batch(peptide, B, seed, lengths)
    loss = model(atom_types=at, x_coords=x, x_velocs=xv, y_coords=y, y_velocs=yv, adj_list=EMPTY_ADJ, edge_batch_idx=EMPTY_EBI,
                 masked_elements=mask)
    loss.backward()
    out = dict(atom_types=at.numpy(), x_coords=x.numpy(), x_velocs=xv.numpy(), y_coords=y.numpy(), y_velocs=yv.numpy(),
               masked_elements=mask.numpy(), weight_seed=np.int64(wseed), loss=loss.detach().numpy())
    names, norms = [], []
    for k, p in model.named_parameters():
        names.append(k)
        norms.append(0.0 if p.grad is None else float(p.grad.double().norm()))
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms)
    named = dict(model.named_parameters())
    for k in [k for k in GRAD_FULL_KEYS if k in named] + list(extra_keys):  # (`local` modules have other attention parameter names)
        g = named[k].grad
        out["grad::" + k] = (g[:8] if g.numel() > 20000 else g).numpy()  # large matrices: first 8 rows
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, "loss", float(loss), "total grad norm", float(np.sqrt((np.array(norms) ** 2).sum())))


TINY = OracleConfig(atom_embedding_dim=8, latent_mlp_hidden_dims=[24], num_coupling_layers=4, num_transformer_layers=2,
                    d_model=16, dim_feedforward=32, lengthscales=[0.3, 1.0])
FULL = OracleConfig()


def chirality_case():
    """Known answers for utils/chirality.py on the 2olx topology (hand-derived bonds)."""
    pep = tetrapeptide_2olx()
    adj = torch.tensor(pep.bonds)
    at = torch.tensor(pep.atom_types)[None]
    centers = ref_chirality.find_chirality_centers(adj, at)
    g = torch.Generator().manual_seed(5)
    coords = torch.tensor(pep.coords_nm, dtype=torch.float32)[None] + 0.01 * torch.randn(6, pep.num_atoms, 3, generator=g)
    coords[1] = coords[1] * torch.tensor([1.0, 1.0, -1.0])  # mirror -> every centre flips
    # sample 3: swap two substituents of one centre only
    c0 = centers[0]
    tmp = coords[3, c0[1]].clone(); coords[3, c0[1]] = coords[3, c0[2]]; coords[3, c0[2]] = tmp
    ref_signs = ref_chirality.compute_chirality_sign(coords[:1], centers)
    signs = ref_chirality.compute_chirality_sign(coords, centers)
    changed = ref_chirality.check_symmetry_change(coords, centers, ref_signs)
    np.savez_compressed(os.path.join(HERE, "chirality_2olx.npz"), bonds=pep.bonds, atom_types=at.numpy(), centers=centers.numpy(),
                        coords=coords.numpy(), ref_signs=ref_signs.numpy(), signs=signs.numpy(), changed=changed.numpy())
    print("chirality centers", centers.tolist(), "changed", changed.tolist())


def init_case():
    """Reference default initialisation under a fixed torch seed (construction-order contract)."""
    enc = CustomAttentionEncoderLayerConfig(d_model=TINY.d_model, dim_feedforward=TINY.dim_feedforward, dropout=0.0,
                                            num_heads=len(TINY.lengthscales), attention_type="kernel",
                                            lengthscales=list(TINY.lengthscales), normalise_kernel_values=True)
    mc = CustomAttentionTransformerNVPConfig(atom_embedding_dim=TINY.atom_embedding_dim, latent_mlp_hidden_dims=list(TINY.latent_mlp_hidden_dims),
                                             num_coupling_layers=TINY.num_coupling_layers, num_transformer_layers=TINY.num_transformer_layers,
                                             encoder_layer_config=enc)
    torch.manual_seed(0)
    model = custom_transformer_nvp_constructor(mc)
    sd = {k: v.numpy() for k, v in model.state_dict().items()}
    np.savez_compressed(os.path.join(HERE, "tiny_init_seed0.npz"), **sd)
    print("init case:", len(sd), "tensors")


FULL_LEARNABLE = OracleConfig(attention_type="learnable_kernel")
TINY_LEARNABLE = OracleConfig(atom_embedding_dim=8, latent_mlp_hidden_dims=[24], num_coupling_layers=4, num_transformer_layers=2,
                              d_model=16, dim_feedforward=32, lengthscales=[0.3, 1.0], attention_type="learnable_kernel")


TINY_CHEB = OracleConfig(atom_embedding_dim=8, latent_mlp_hidden_dims=[24], num_coupling_layers=4, num_transformer_layers=2,
                         d_model=16, dim_feedforward=32, lengthscales=[0.3, 1.0], attention_type="chebyshev_kernel", cheb_order=6,
                         force_asymptotic_zero=True)
FULL_CHEB = OracleConfig(attention_type="chebyshev_kernel", cheb_order=12, force_asymptotic_zero=False)


def chebyshev_cases():
    """`chebyshev_kernel` attention: per-layer Chebyshev-rational basis functions, no score sharing."""
    ad = alanine_dipeptide()
    run_case("tiny_ad_chebyshev", TINY_CHEB, ad, B=3, seed=23, lengths=[22, 15, 9], sample_S=3, trace_layer0=False)
    run_case("full_ad22_chebyshev", FULL_CHEB, ad, B=3, seed=24, sample_S=2, trace_layer0=False)


TINY_LOCAL = OracleConfig(atom_embedding_dim=8, latent_mlp_hidden_dims=[24], num_coupling_layers=4, num_transformer_layers=2,
                          d_model=16, dim_feedforward=32, lengthscales=[], attention_type="local", max_radius=0.3, num_heads=3)
FULL_LOCAL = OracleConfig(lengthscales=[], attention_type="local", max_radius=0.45, num_heads=6)


def local_cases():
    """`local` attention: dot-product attention over the atoms within max_radius."""
    ad = alanine_dipeptide()
    # construction-order contract of the `local` module tree: reference default initialisation under a fixed torch seed
    enc = CustomAttentionEncoderLayerConfig(d_model=TINY_LOCAL.d_model, dim_feedforward=TINY_LOCAL.dim_feedforward, dropout=0.0,
                                            num_heads=TINY_LOCAL.num_heads, attention_type="local", max_radius=TINY_LOCAL.max_radius)
    mc = CustomAttentionTransformerNVPConfig(atom_embedding_dim=TINY_LOCAL.atom_embedding_dim,
                                             latent_mlp_hidden_dims=list(TINY_LOCAL.latent_mlp_hidden_dims),
                                             num_coupling_layers=TINY_LOCAL.num_coupling_layers,
                                             num_transformer_layers=TINY_LOCAL.num_transformer_layers, encoder_layer_config=enc)
    torch.manual_seed(0)
    sd = {k: v.numpy() for k, v in custom_transformer_nvp_constructor(mc).state_dict().items()}
    np.savez_compressed(os.path.join(HERE, "tiny_local_init_seed0.npz"), **sd)
    run_case("tiny_ad_local", TINY_LOCAL, ad, B=3, seed=25, lengths=[22, 15, 9], sample_S=3, trace_layer0=False)
    run_case("full_ad22_local", FULL_LOCAL, ad, B=3, seed=26, sample_S=2, trace_layer0=False)


def learnable_cases():
    """`learnable_kernel` attention: per-layer log_lengthscales, of which the reference uses only
    the first executed layer's (cache key quirk) -- layer 0 for log_likelihood, the last coupling layer when sampling."""
    ad = alanine_dipeptide()
    run_case("tiny_ad_learnable", TINY_LEARNABLE, ad, B=3, seed=21, lengths=[22, 15, 9], sample_S=3, trace_layer0=False)
    run_case("full_ad22_learnable", FULL_LEARNABLE, ad, B=3, seed=22, sample_S=2, trace_layer0=False)


def dataloader_case():
    """Batch / wire formats: the reference's own dense collate on ragged synthetic datapoints, and
    its conditioning/target pairing (`load_pdb_trace_data`) on a small synthetic log-spaced trajectory file.  mdtraj is
    not installed: `md.load` is replaced by a stand-in topology built from the alanine-dipeptide template (the pairing
    logic under test does not depend on it)."""
    import timewarp.dataloader as ref_dl

    ad = alanine_dipeptide()
    g = torch.Generator().manual_seed(31)
    pts = []
    for i, n in enumerate([22, 15, 9]):
        f = lambda: torch.randn(n, 3, generator=g)  # noqa: E731
        bonds = torch.tensor([b for b in ad.bonds.tolist() if max(b) < n], dtype=torch.int64)
        pts.append(ref_dl.MolDynDatapoint(name=f"mol{i}", atom_types=torch.tensor(ad.atom_types[:n]), adj_list=bonds, atom_coords=f(),
                                          atom_velocs=f(), atom_forces=f(), atom_coord_targets=f(), atom_veloc_targets=f(),
                                          atom_force_targets=f()))
    batch = ref_dl.moldyn_dense_collate_fn(pts)
    out = {f"in{i}_{k}": getattr(p, k).numpy() for i, p in enumerate(pts) for k in
           ("atom_types", "adj_list", "atom_coords", "atom_velocs", "atom_forces", "atom_coord_targets", "atom_veloc_targets", "atom_force_targets")}
    out.update({f"batch_{k}": getattr(batch, k).numpy() for k in
                ("atom_types", "adj_list", "edge_batch_idx", "atom_coords", "atom_velocs", "atom_forces", "atom_coord_targets",
                 "atom_veloc_targets", "atom_force_targets", "masked_elements")})
    out["lengths_mask"] = ref_dl.lengths_to_mask(torch.tensor([3, 1, 4])).numpy()

    # synthetic trajectory: log-spaced steps like simulation output (1, 10, 100, 1000 after every 10000-step block)
    rng = np.random.default_rng(7)
    steps = np.array(sorted({base + off for base in range(0, 60000, 10000) for off in (0, 1, 10, 100, 1000)}), dtype=np.int64)
    T, V = len(steps), ad.num_atoms
    traj = dict(step=steps, positions=rng.standard_normal((T, V, 3)).astype(np.float32),
                velocities=rng.standard_normal((T, V, 3)).astype(np.float32), forces=rng.standard_normal((T, V, 3)).astype(np.float32))
    traj_path = os.path.join(HERE, "synthetic_ad-traj-arrays.npz")
    np.savez_compressed(traj_path, **traj)

    class _A:
        def __init__(self, i, sym):
            self.index, self.element = i, types.SimpleNamespace(symbol=sym)

    atoms = [_A(i, e) for i, e in enumerate(ad.elements)]
    topo = types.SimpleNamespace(atoms=atoms, bonds=[types.SimpleNamespace(atom1=atoms[a], atom2=atoms[b]) for a, b in ad.bonds.tolist()])
    ref_dl.md.load = lambda path: types.SimpleNamespace(topology=topo)
    for sw, eq in ((1, False), (10, False), (100, True), (1000, False)):
        info = ref_dl.load_pdb_trace_data("synthetic_ad", "unused.pdb", traj_path, step_width=sw, equal_data_spacing=eq)
        out[f"pairs_sw{sw}_eq{int(eq)}_coord_features"] = np.stack(info.coord_features) if info.coord_features else np.zeros((0, V, 3), np.float32)
        out[f"pairs_sw{sw}_eq{int(eq)}_veloc_targets"] = np.stack(info.veloc_targets) if info.veloc_targets else np.zeros((0, V, 3), np.float32)
        out[f"pairs_sw{sw}_eq{int(eq)}_node_types"] = info.node_types
        out[f"pairs_sw{sw}_eq{int(eq)}_adj_list"] = info.adj_list
        print("pairs", sw, eq, len(info.coord_features))
    np.savez_compressed(os.path.join(HERE, "dataloader_collate.npz"), **out)



def checkpoint_case():
    """A checkpoint written by the reference's own `save_model` (utilities/model_utils.py:12-29) for the model of the
    `tiny_ad` case -> tests/golden/ckpt/run0/best_model.pt."""
    from utilities.model_utils import save_model

    model = ref_model(TINY, 0)[0]
    d = os.path.join(HERE, "ckpt", "run0")
    os.makedirs(d, exist_ok=True)
    save_model(os.path.join(d, "best_model.pt"), model, step=123)
    print("checkpoint case written")


def md_case():
    """Consecutive integrator steps and kinetic energies recorded in the reference's OpenMM trajectory fixtures (preset
    "T1-peptides": LangevinIntegrator 310 K, 0.3 / ps, 0.5 fs, simulation/md.py:75-82) -> tests/golden/langevin_2olx_pairs.npz.
    Pins oracle/md_oracle.py (update rule, noise scale, leapfrog kinetic energy, OpenMM element masses)."""
    pdb = open(os.path.join(REF, "simulation/testdata/implicit-2olx-traj-cpu-state0.pdb")).read().splitlines()
    elements = np.array([l[76:78].strip() for l in pdb if l.startswith(("ATOM", "HETATM"))])
    out = {"elements": elements, "timestep_ps": 0.0005, "friction_per_ps": 0.3, "temperature_K": 310.0}
    x0, v0, f0, x1, v1 = [], [], [], [], []
    for rel in ("simulation/testdata/implicit-2olx-traj-cpu-arrays.npz", "simulation/testdata/implicit-2olx-traj-arrays.npz",
                "testdata/output/2olx-traj-arrays.npz"):
        d = np.load(os.path.join(REF, rel))
        st = d["step"]
        first = [i for i in range(len(st) - 1) if st[i + 1] == st[i] + 1][:10]  # 10 steps per fixture keep the file small
        for i in first:
            if True:
                x0.append(d["positions"][i]), v0.append(d["velocities"][i]), f0.append(d["forces"][i])
                x1.append(d["positions"][i + 1]), v1.append(d["velocities"][i + 1])
    out.update(x0=np.stack(x0), v0=np.stack(v0), f0=np.stack(f0), x1=np.stack(x1), v1=np.stack(v1))
    d = np.load(os.path.join(REF, "simulation/testdata/implicit-2olx-traj-cpu-arrays.npz"))  # checked by simulation/tests/test_md.py:35-47
    out.update(ke_velocities=d["velocities"], ke_forces=d["forces"], ke_openmm=d["energies"][:, 1])
    # the reference's golden potential energies (simulation/tests/test_md.py:35-47, atol 1e-3 kJ/mol) + the topology they
    # belong to: reproducible wherever OpenMM 7.7 + amber99sbildn/amber99_obc exist (tests/test_forcefield_cpu.py, skipped here)
    out.update(pot_positions=d["positions"], pot_openmm=d["energies"][:, 0], state0_pdb=np.array("\n".join(pdb)))
    np.savez_compressed(os.path.join(HERE, "langevin_2olx_pairs.npz"), **out)
    print("md case:", len(x0), "consecutive steps,", len(d["step"]), "kinetic energies")


def learnable_grad_case():
    """learnable_kernel training: the gradient reaches the log_lengthscales of the first attention layer the pass executes
    (flow.chain.0.scale_transformer.encoder_layers.0) and no other (cache quirk: their .grad stays None -> norm 0)."""
    ad = alanine_dipeptide()
    grad_case("grads_full_ad22_learnable", FULL_LEARNABLE, ad, B=3, seed=31, lengths=[22, 17, 12],
              extra_keys=["flow.chain.0.scale_transformer.encoder_layers.0.self_attn.attention.log_lengthscales"])


def local_grad_case():
    """`local` attention training (local_self_attention.py:46-119 under autograd): qkv_proj / output_proj gradients."""
    ad = alanine_dipeptide()
    grad_case("grads_full_ad22_local", FULL_LOCAL, ad, B=3, seed=33, lengths=[22, 17, 12],
              extra_keys=["flow.chain.3.shift_transformer.encoder_layers.1.self_attn.qkv_proj.weight",
                          "flow.chain.3.shift_transformer.encoder_layers.1.self_attn.output_proj.weight"])


def chebyshev_grad_case():
    """chebyshev_kernel training: every attention layer's cheb_coeffs receives its own gradient."""
    ad = alanine_dipeptide()
    grad_case("grads_full_ad22_chebyshev", FULL_CHEB, ad, B=3, seed=32, lengths=[22, 17, 12],
              extra_keys=["flow.chain.0.scale_transformer.encoder_layers.0.self_attn.attention.cheb_coeffs",
                          "flow.chain.5.shift_transformer.encoder_layers.2.self_attn.attention.cheb_coeffs"])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "chebyshev_grad":
        chebyshev_grad_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "local_grad":
        local_grad_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "learnable_grad":
        learnable_grad_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "checkpoint":
        checkpoint_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "md":
        md_case()
        sys.exit(0)
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "learnable":
        learnable_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "local":
        local_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "chebyshev":
        chebyshev_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "dataloader":
        dataloader_case()
        sys.exit(0)
    init_case()
    ad, olx = alanine_dipeptide(), tetrapeptide_2olx()
    run_case("tiny_ad_ragged", TINY, ad, B=3, seed=11, lengths=[22, 15, 9], sample_S=4)
    run_case("tiny_ad", TINY, ad, B=4, seed=12, sample_S=4)
    run_case("full_ad22", FULL, ad, B=4, seed=0, sample_S=3)
    run_case("full_ad22_ragged", FULL, ad, B=3, seed=1, lengths=[22, 17, 12], sample_S=2)
    run_case("full_2olx65", FULL, olx, B=2, seed=2, sample_S=2)
    chirality_case()
    grad_case("grads_full_ad22", FULL, ad, B=4, seed=3)
    grad_case("grads_full_ad22_ragged", FULL, ad, B=3, seed=4, lengths=[22, 17, 12])
    learnable_cases()
    learnable_grad_case()
    chebyshev_cases()
    chebyshev_grad_case()
    local_cases()
    local_grad_case()
    md_case()
    checkpoint_case()
    dataloader_case()

