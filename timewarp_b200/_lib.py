"""ctypes binding of libtimewarp_b200.so (the C ABI declared in include/timewarp_b200.h).

The library is the product: if it is missing or fails to load, every compute entry point raises.
There is no CPU or pure-PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtimewarp_b200.so")

TW_OK = 0
TW_ATTENTION_KERNEL, TW_ATTENTION_LOCAL, TW_ATTENTION_CHEBYSHEV = 0, 1, 2
TW_INTEGRATOR_LANGEVIN, TW_INTEGRATOR_LANGEVIN_MIDDLE = 0, 1
TW_MAX_MLP_HIDDEN = 4
TW_MAX_HEADS = 16
PRECISION = {"fp32": 0, "bf16x3": 1, "bf16": 2}
TW_FLOW_DISPLACEMENT_TARGET = 1


class FlowConfig(C.Structure):
    _fields_ = [
        ("atom_embedding_dim", C.c_int32),
        ("num_mlp_hidden", C.c_int32),
        ("mlp_hidden_dims", C.c_int32 * TW_MAX_MLP_HIDDEN),
        ("num_coupling_layers", C.c_int32),
        ("num_transformer_layers", C.c_int32),
        ("d_model", C.c_int32),
        ("dim_feedforward", C.c_int32),
        ("num_heads", C.c_int32),
        ("position_layer_index_mod_2", C.c_int32),
        ("num_atom_types", C.c_int32),
        ("layer_norm_eps", C.c_float),
        ("precision", C.c_int32),
        ("attention_type", C.c_int32),  # TW_ATTENTION_*: 0 kernel / learnable_kernel, 1 local, 2 chebyshev_kernel
        ("cheb_order", C.c_int32),
        ("force_asymptotic_zero", C.c_int32),
        ("max_radius", C.c_float),  # local: neighbourhood radius in nm
    ]


class EnergySystem(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int32),
        ("n_bonds", C.c_int32), ("bond_idx", C.c_void_p), ("bond_param", C.c_void_p),
        ("n_angles", C.c_int32), ("angle_idx", C.c_void_p), ("angle_param", C.c_void_p),
        ("n_torsions", C.c_int32), ("torsion_idx", C.c_void_p), ("torsion_param", C.c_void_p),
        ("charge", C.c_void_p), ("sigma", C.c_void_p), ("epsilon", C.c_void_p), ("excluded", C.c_void_p),
        ("n_exceptions", C.c_int32), ("exception_idx", C.c_void_p), ("exception_param", C.c_void_p),
        ("cutoff", C.c_double), ("reaction_field_eps", C.c_double), ("one_4pi_eps0", C.c_double),
        ("use_gb", C.c_int32), ("gb_radius", C.c_void_p), ("gb_scale", C.c_void_p),
        ("gb_alpha", C.c_double), ("gb_beta", C.c_double), ("gb_gamma", C.c_double), ("gb_offset", C.c_double),
        ("solute_dielectric", C.c_double), ("solvent_dielectric", C.c_double), ("surface_area_energy", C.c_double),
    ]  # fmt: skip


_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_SIGNATURES = {
    "tw_abi_version": (C.c_int, []),
    "tw_last_error": (C.c_char_p, []),
    "tw_debug_launch_count": (C.c_longlong, []),
    "tw_prof_enable": (C.c_int, [C.c_int]),
    "tw_prof_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "tw_flow_num_params": (C.c_int, [C.POINTER(FlowConfig)]),
    "tw_flow_workspace_bytes": (C.c_int, [C.POINTER(FlowConfig), _I64, _I64, _I64, C.POINTER(C.c_size_t)]),
    "tw_attn_scores": (C.c_int, [_P, _P, _P, _I64, _I64, _I32, _P, _P]),
    "tw_flow_packed_bytes": (C.c_int, [C.POINTER(FlowConfig), C.POINTER(C.c_size_t)]),
    "tw_flow_pack_weights": (C.c_int, [C.POINTER(FlowConfig), _P, _P, C.c_size_t, _P]),
    "tw_flow_scale_shift": (C.c_int, [C.POINTER(FlowConfig), _P, _I32, _P, _P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, _P, C.c_size_t, _P]),
    "tw_flow_log_likelihood": (C.c_int, [C.POINTER(FlowConfig), _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I32, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "tw_flow_sample": (C.c_int, [C.POINTER(FlowConfig), _P, _P, _P, _P, _P, _I64, _I64, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "tw_peptide_energy": (C.c_int, [C.POINTER(EnergySystem), _P, _I64, _P, _P, _P, _P]),
    "tw_langevin_steps": (C.c_int, [C.POINTER(EnergySystem), _P, _P, _P, _I64, _I32, _I32, C.c_double, C.c_double, C.c_double, _P,
                                    C.c_uint64, C.c_uint64, _P]),
    "tw_chirality": (C.c_int, [_P, _P, _P, _I64, _I64, _I32, _P, _P, _P]),
    "tw_kinetic_energy": (C.c_int, [_P, _P, C.c_float, _I64, _I64, _P, _P]),
    "tw_mh_accept": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "tw_debug_umma_probe": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "tw_debug_umma_timing": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "tw_debug_set_ffn_trace": (C.c_int, [_P]),
    "tw_debug_set_trace": (C.c_int, [C.c_int, _P]),
    "tw_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _P, _P]),
    "tw_threshold_accept": (C.c_int, [_P, _P, _P, _P, C.c_float, _I64, _I64, _P, _P]),
    "tw_flow_train_bytes": (C.c_int, [C.POINTER(FlowConfig), _I64, _I64, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "tw_flow_log_likelihood_train": (C.c_int, [C.POINTER(FlowConfig), _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I32, _P, _P, _P, C.c_size_t, _P]),
    "tw_flow_log_likelihood_backward": (C.c_int, [C.POINTER(FlowConfig), _P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, C.c_size_t, _P, C.c_size_t, _P]),
    "tw_flow_log_likelihood_backward_inputs": (C.c_int, [C.POINTER(FlowConfig), _P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, C.c_size_t, _P,
                                                         C.c_size_t, _P, _P, _P, _P, _P]),
    "tw_flow_sample_train": (C.c_int, [C.POINTER(FlowConfig), _P, _P, _P, _P, _P, _I64, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "tw_flow_sample_backward": (C.c_int, [C.POINTER(FlowConfig), _P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, _P, _P, C.c_size_t, _P, C.c_size_t,
                                          _P, _P, _P]),
    "tw_debug_gemm": (C.c_int, [C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, C.c_size_t, _P]),
}  # fmt: skip

_lib: Optional[C.CDLL] = None


class TimewarpB200Error(RuntimeError):
    pass


def exported_symbols():
    return sorted(_SIGNATURES)


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TimewarpB200Error(
            f"{LIB_PATH} not found: build it with `python -m timewarp_b200.build` (or __graft_entry__.build()). "
            "timewarp_b200 has no CPU / PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != TW_OK:
        msg = load().tw_last_error()
        raise TimewarpB200Error(f"{what} failed with status {status}: {msg.decode() if msg else ''}")


def ptr(t) -> Optional[int]:
    """Device pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()
