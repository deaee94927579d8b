"""ctypes binding of libmdm_b200.so (the C-ABI declared in include/mdm_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MDM_B200_LIB: developer knob to A/B a differently-built copy of the same library (tools/op_bench.py)
LIB_PATH = os.environ.get("MDM_B200_LIB") or os.path.join(_HERE, "csrc", "libmdm_b200.so")

MDM_F32, MDM_BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_SILU, ACT_EXPFEAT = 0, 1, 2, 3

_ERR = {1: "invalid argument", 2: "CUDA error", 3: "unsupported configuration"}


class MdmError(RuntimeError):
    pass


class GemmEpi(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p), ("rowscale", C.c_void_p), ("rowmask", C.c_void_p),
        ("resid", C.c_void_p), ("ld_resid", C.c_int), ("resid_mod", C.c_int),
        ("alpha", C.c_float), ("beta", C.c_float), ("act", C.c_int),
        ("out_f32", C.c_void_p), ("ld_f32", C.c_int),
        ("out_bf16", C.c_void_p), ("ld_bf16", C.c_int), ("bf16_pre_resid", C.c_int), ("pair_tiles", C.c_int),
        ("tile_k", C.c_void_p), ("mn_major", C.c_int),
    ]


class RowOp(C.Structure):
    _fields_ = [
        ("inp", C.c_void_p), ("in_dt", C.c_int),
        ("ln1_w", C.c_void_p), ("ln1_b", C.c_void_p), ("l2norm", C.c_int),
        ("out1_f32", C.c_void_p), ("out1_a", C.c_void_p),
        ("ln2_w", C.c_void_p), ("ln2_b", C.c_void_p),
        ("film", C.c_void_p), ("rows_per_seq", C.c_int), ("silu", C.c_int),
        ("out2_f32", C.c_void_p), ("out2_a", C.c_void_p), ("out0_a", C.c_void_p),
    ]


class Bgemm(C.Structure):
    _fields_ = [("A", C.c_void_p), ("a_dt", C.c_int), ("a_z1", C.c_long), ("a_z2", C.c_long), ("a_rs", C.c_long), ("a_cs", C.c_long),
                ("B", C.c_void_p), ("b_dt", C.c_int), ("b_z1", C.c_long), ("b_z2", C.c_long), ("b_rs", C.c_long), ("b_cs", C.c_long),
                ("C", C.c_void_p), ("c_dt", C.c_int), ("c_z1", C.c_long), ("c_z2", C.c_long), ("c_rs", C.c_long), ("c_cs", C.c_long),
                ("Z1", C.c_int), ("Z2", C.c_int), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("alpha", C.c_float), ("accumulate", C.c_int),
                ("m_limit", C.c_void_p), ("k_limit", C.c_void_p), ("limit_shift", C.c_int), ("tensor_cores", C.c_int)]


EP_MAX_RANKS = 8


class EpPeers(C.Structure):
    _fields_ = [("xp", C.c_void_p * EP_MAX_RANKS), ("rowscale", C.c_void_p * EP_MAX_RANKS),
                ("yp", C.c_void_p * EP_MAX_RANKS), ("cnt", C.c_void_p * EP_MAX_RANKS),
                ("flags", C.c_void_p * EP_MAX_RANKS)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_long, C.c_float

_SIGS = {
    "mdm_gemm_bf16": [_P, _I, _L, _P, _I, _L, _I, _I, _I, _P, _I, _P, C.POINTER(GemmEpi), _I, _P],
    "mdm_gemm_f32": [_P, _I, _L, _P, _I, _L, _I, _I, _I, _P, _I, _P, C.POINTER(GemmEpi), _P],
    "mdm_rowop": [C.POINTER(RowOp), _L, _I, _I, _P],
    "mdm_gemm_rowop": [C.POINTER(RowOp), _L, _I, _P, _I, _L, _I, C.POINTER(GemmEpi), _P],
    "mdm_gemm_ln": [_P, _I, _L, _P, _I, _L, _I, _I, _I, C.POINTER(GemmEpi), C.POINTER(RowOp), _P],
    "mdm_gemm_gate": [_P, _I, _L, _P, _I, _L, _I, _I, _I, C.POINTER(GemmEpi), _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "mdm_fastattn": [_P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P],
    "mdm_fastattn_ordered": [_P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "mdm_lincross_ctx": [_P, _P, _I, _P, _I, _I, _I, _I, _P, _P],
    "mdm_lincross_apply": [_P, _I, _P, _I, _I, _I, _I, _P, _P],
    "mdm_lincross_apply_ex": [_P, _I, _P, _P, _I, _I, _I, _I, _P, _P],
    "mdm_lincross_apply_style": [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "mdm_transpose_cast_bf16": [_P, _L, _I, _I, _P, _P],
    "mdm_softmax_cross": [_P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P],
    "mdm_moe_gate": [_P, _L, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "mdm_moe_gate_forced": [_P, _L, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "mdm_moe_scan": [_P, _P, _P, _L, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "mdm_moe_permute": [_P, _L, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P],
    "mdm_moe_combine_film": [_P, _I, _P, _L, _I, _I, _P, _P, _P, _I, _P, _P],
    "mdm_softmax_topk": [_P, _L, _I, _I, _P, _P, _P, _P],
    "mdm_timestep_embedding": [_P, _I, _I, _P, _I, _P],
    "mdm_gated_mix": [_P, _P, _L, _P, _I, _P],
    "mdm_pad_cast": [_P, _L, _I, _P, _I, _I, _P],
    "mdm_cfg_update": [_P, _P, _P, _P, _P, _P, _I, _F, _I, _I, _L, _P, _P, _P],
    "mdm_q_sample": [_P, _P, _P, _P, _I, _I, _L, _P, _P],
    "mdm_p_mean_variance": [_P, _P, _P, _P, _P, _I, _I, _I, _L, _P, _P, _P, _P],
    "mdm_recover_from_ric": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "mdm_masked_mse": [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P],
    "mdm_rowop_bwd": [C.POINTER(RowOp), _L, _I, _I, _P, _P, _P, _P, C.POINTER(C.c_int), C.POINTER(C.c_int), _P],
    "mdm_rowop_bwd2": [C.POINTER(RowOp), _L, _I, _I, _P, _P, _I, _P, _P, _P, _I, C.POINTER(C.c_int), C.POINTER(C.c_int), _P],
    "mdm_gelu_fwd": [_P, _L, _P, _P],
    "mdm_gelu_bwd": [_P, _P, _L, _P, _P],
    "mdm_rowscale_bf16": [_P, _P, _L, _I, _P, _P],
    "mdm_seg_colsum_bf16": [_P, _I, _P, _P, _I, _P, _P],
    "mdm_transpose_split_bf16": [_P, _L, _I, _I, _I, _P, _P],
    "mdm_sum_partials": [_P, _I, _L, _I, _P, _P],
    "mdm_colsum_bf16": [_P, _L, _I, _I, _P, _P],
    "mdm_ddim_update": [_P, _P, _P, _P, _P, _P, _P, _I, _F, _F, _I, _I, _L, _P, _P, _P],
    "mdm_ep_counts": [_P, _P, _L, _I, _I, _I, _I, _I, C.POINTER(EpPeers), _P, _P, _P, _P],
    "mdm_ep_scan": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "mdm_ep_dispatch": [_P, _L, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, C.POINTER(EpPeers), _I, _P, _P],
    "mdm_ep_combine_film": [C.POINTER(EpPeers), _I, _P, _L, _I, _I, _I, _P, _P, _P, _I, _P, _P],
    "mdm_ep_barrier": [C.POINTER(EpPeers), _I, _I, _P, _P, _P],
    "mdm_ipc_get_handle": [_P, _P, C.POINTER(C.c_long)],
    "mdm_ipc_open_handle": [_P, C.POINTER(C.c_void_p)],
    "mdm_ipc_close_handle": [_P],
    "mdm_bgemm": [C.POINTER(Bgemm), _P],
    "mdm_sizeof_bgemm": [],
    "mdm_fa_prep": [_P, _I, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P],
    "mdm_fa_feat": [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    "mdm_fa_out_bwd": [_P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _P, _P, C.POINTER(C.c_int), _P],
    "mdm_fa_out_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P],
    "mdm_fa_feat_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P],
    "mdm_fa_prep_bwd": [_P, _I, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, C.POINTER(C.c_int), _P],
    "mdm_head_softmax": [_P, _I, _I, _I, _I, _I, _P, _P],
    "mdm_head_softmax_bwd": [_P, _P, _I, _I, _I, _I, _P, _I, _P],
    "mdm_key_softmax": [_P, _P, _I, _I, _I, _I, _P],
    "mdm_key_softmax_bwd": [_P, _P, _I, _I, _I, _I, _P],
    "mdm_col_softmax": [_P, _I, _P, _I, _I, _I, _P, _P],
    "mdm_col_softmax_bwd": [_P, _P, _I, _I, _I, _P, _I, _P],
    "mdm_moe_combine_sum": [_P, _I, _P, _P, _L, _I, _I, _P, _P],
    "mdm_moe_combine_bwd": [_P, _I, _P, _P, _L, _I, _I, _P, _P, _P, _P],
    "mdm_moe_gate_bwd_logits": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P, _P],
    "mdm_moe_unpermute_bwd": [_P, _I, _P, _P, _P, _L, _I, _I, _I, _I, _P, _P],
    "mdm_moe_wgrad_tables": [_P, _P, _L, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "mdm_act_fwd": [_P, _I, _L, _I, _P, _P],
    "mdm_act_bwd": [_P, _P, _I, _L, _I, _P, _P],
    "mdm_axpby": [_P, _I, _F, _P, _I, _F, _L, _P, _I, _P],
    "mdm_gated_mix_bwd": [_P, _P, _P, _L, _P, _P, _P],
    "mdm_masked_mse_grad": [_P, _P, _P, _I, _I, _I, _F, _P, _P],
    "mdm_colsum": [_P, _I, _L, _I, _L, _I, _P, _P],
    "mdm_colsum_prod": [_P, _P, _P, _L, _I, _I, _P, _P],
    "mdm_transpose_split": [_P, _I, _L, _I, _L, _I, _I, _P, _P],
    "mdm_seg_colsum": [_P, _I, _I, _P, _P, _I, _I, _P, _P],
    "mdm_grad_clip_coef": [_P, _L, _F, _P, _I, _P, _P],
    "mdm_adam_step": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _I, _P, _P, _P],
    "mdm_num_sms": [],
    "mdm_set_pdl": [_I],
    "mdm_sizeof_gemm_epi": [],
    "mdm_sizeof_rowop": [],
    "mdm_sizeof_ep_peers": [],
}

_lib = None


def load():
    """Load the library once; raise (never fall back) if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MdmError(
            "libmdm_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `python -m motiondiffusion_moe_b200.build`. There is no CPU fallback."
            % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: loud by design
        fn.argtypes = args
        fn.restype = C.c_int
    lib.mdm_version.restype = C.c_char_p
    lib.mdm_version.argtypes = []
    for fn, st in ((lib.mdm_sizeof_gemm_epi, GemmEpi), (lib.mdm_sizeof_rowop, RowOp), (lib.mdm_sizeof_ep_peers, EpPeers),
                   (lib.mdm_sizeof_bgemm, Bgemm)):
        if fn() != C.sizeof(st):     # a stale .so or a drifted binding: refuse to pass short structs to the kernels
            raise MdmError("struct layout mismatch between _lib.py and %s: %s is %d bytes here, %d in the library"
                           % (LIB_PATH, st.__name__, C.sizeof(st), fn()))
    _lib = lib
    return lib


LAUNCHES = [0]   # kernels launched through this binding (every C-ABI compute call launches exactly one)


def check(status, what):
    LAUNCHES[0] += 1
    if status != 0:
        raise MdmError("%s failed: %s (status %d)" % (what, _ERR.get(status, "unknown"), status))


def exported_symbols():
    return list(_SIGS.keys()) + ["mdm_version"]
