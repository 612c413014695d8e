"""ctypes binding of liblgcn_b200.so (declared in include/lgcn_b200.h).

The shared library is the product; this module only loads it and declares argument types.  There is
no fallback: if the library is missing or a call is rejected, a RuntimeError is raised.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int32, c_int64, c_size_t, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblgcn_b200.so")

MAX_Z = 8
MAX_TOPK = 128


class SpmmPlan(Structure):
    _fields_ = [
        ("seg_len", c_int32), ("n_long", c_int32), ("n_segs", c_int32), ("n_items", c_int32),
        ("d_max", c_int32), ("pad", c_int32),
        ("items", c_void_p), ("seginfo", c_void_p), ("counters", c_void_p), ("partials", c_void_p), ("acc", c_void_p),
        ("hinted_indices", c_void_p),
    ]


class SpmmPeers(Structure):
    _fields_ = [("n_peers", c_int32), ("multicast", c_int32), ("y", c_void_p * 7), ("p", c_void_p * 7)]


class BprFeat(Structure):
    _fields_ = [("n_parts", c_int32), ("part", c_int32), ("scalars_dev", c_void_p), ("records_local", c_void_p),
                ("records_peer", c_void_p * 8)]


class AdamScalars(Structure):
    _fields_ = [
        ("step_size", c_float), ("bc2_sqrt", c_float), ("beta1", c_float), ("beta2", c_float),
        ("w1", c_float), ("w2", c_float), ("eps", c_float), ("pad0", c_float),
        ("step", c_int32), ("pad1", c_int32), ("lr_d", c_double), ("beta1_d", c_double), ("beta2_d", c_double),
    ]


_P = c_void_p
_SIGNATURES = {
    "lgcn_abi_version": (ctypes.c_int, []),
    "lgcn_last_error": (c_char_p, []),
    "lgcn_device_info": (ctypes.c_int, [POINTER(c_int32)]),
    "lgcn_enable_peer_access": (ctypes.c_int, [c_int32]),
    "lgcn_debug_poke": (ctypes.c_int, [_P, c_float, c_int32, POINTER(c_int32), _P]),
    "lgcn_csr_build_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "lgcn_csr_build": (ctypes.c_int, [_P, _P, c_int64, c_int32, c_int32, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "lgcn_degree_accumulate": (ctypes.c_int, [_P, _P, c_int64, c_int32, c_int32, _P, _P, _P]),
    "lgcn_degree_finalize": (ctypes.c_int, [_P, c_int32, _P, _P, _P]),
    "lgcn_csr_rows_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "lgcn_csr_rows_emit": (ctypes.c_int, [_P, _P, c_int64, c_int32, c_int32, c_int32, c_int32, c_int64, _P, _P, _P, c_size_t, _P]),
    "lgcn_csr_rows_finish": (ctypes.c_int, [c_int64, c_int64, c_int32, c_int32, c_int32, c_int32, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "lgcn_rank_barrier": (ctypes.c_int, [_P, POINTER(c_void_p), c_int32, c_int32, _P, _P, c_int32, _P]),
    "lgcn_copy_words": (ctypes.c_int, [_P, _P, c_int64, c_int32, _P]),
    "lgcn_coo_to_csr": (ctypes.c_int, [_P, _P, c_int64, c_int32, _P, _P, _P]),
    "lgcn_spmm_plan_count": (ctypes.c_int, [_P, c_int32, c_int32, _P, _P]),
    "lgcn_spmm_hint_indices": (ctypes.c_int, [_P, c_int64, _P, c_int32, _P, _P]),
    "lgcn_spmm_plan_count_slab": (ctypes.c_int, [_P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, _P, _P]),
    "lgcn_spmm_plan_fill_slab": (ctypes.c_int, [_P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, _P, _P, _P, c_size_t, _P]),
    "lgcn_spmm_plan_workspace_bytes": (c_size_t, [c_int32]),
    "lgcn_spmm_plan_fill": (ctypes.c_int, [_P, c_int32, c_int32, c_int32, _P, _P, _P, c_size_t, _P]),
    "lgcn_spmm_f32": (ctypes.c_int, [_P, _P, _P, c_int32, c_int32, _P, _P, c_float, c_float,
                                     POINTER(c_void_p), c_int32, POINTER(SpmmPlan), _P, _P, POINTER(SpmmPeers), _P]),
    "lgcn_spmm_adam_f32": (ctypes.c_int, [_P, _P, _P, c_int32, c_int32, _P, _P, c_float, c_float,
                                          POINTER(c_void_p), c_int32, _P, _P, _P, _P, POINTER(SpmmPlan), _P, _P, POINTER(SpmmPeers), c_int32, _P]),
    "lgcn_debug_gather_rows": (ctypes.c_int, [_P, _P, c_int64, c_int32, c_int32, c_int32, _P, _P]),
    "lgcn_debug_spmm_variant": (ctypes.c_int, [ctypes.c_int]),
    "lgcn_adam_init": (ctypes.c_int, [_P, c_double, c_double, c_double, c_double, c_int32, _P]),
    "lgcn_adam_tick": (ctypes.c_int, [_P, _P]),
    "lgcn_step_begin": (ctypes.c_int, [_P, _P, c_int32, _P, _P, c_int64, _P]),
    "lgcn_adam_f32": (ctypes.c_int, [_P, _P, _P, _P, c_int64, _P, _P]),
    "lgcn_bpr_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "lgcn_bpr_fwd_bwd": (ctypes.c_int, [_P, _P, _P, _P, c_int32, _P, c_int32, c_int32, c_int32,
                                        c_float, c_float, c_float, c_float, _P, _P, c_int32, c_int32,
                                        c_int32, _P, c_size_t, _P, _P, _P]),
    "lgcn_bpr_feat_partial": (ctypes.c_int, [_P, _P, _P, _P, c_int32, _P, c_int32, c_int32, c_int32, POINTER(BprFeat), _P, c_size_t, _P]),
    "lgcn_bpr_feat_finish": (ctypes.c_int, [_P, _P, _P, _P, c_int32, _P, c_int32, c_int32, c_int32,
                                            c_float, c_float, c_float, c_float, _P, _P, c_int32, POINTER(BprFeat),
                                            _P, c_size_t, _P, _P, _P]),
    "lgcn_bpr_clear_rows": (ctypes.c_int, [_P, _P, _P, _P, c_int32, _P, c_int32, c_int32, _P]),
    "lgcn_batch_advance": (ctypes.c_int, [_P, c_int32, _P]),
    "lgcn_batch_masks": (ctypes.c_int, [_P, _P, _P, c_int32, _P, c_int32, c_int32, _P, _P, _P, _P, c_int32, _P]),
    "lgcn_batch_masks_rows": (ctypes.c_int, [_P, _P, _P, c_int32, _P, c_int32, c_int32, c_int32, _P, _P]),
    "lgcn_popgate_param_count": (c_int32, [c_int32, c_int32, c_int32]),
    "lgcn_popgate_fuse": (ctypes.c_int, [_P, c_int32, c_int32, c_int32, _P, _P, c_int32, c_int32, c_float, _P, _P, _P]),
    "lgcn_popgate_bpr_workspace_bytes": (c_size_t, [c_int32]),
    "lgcn_popgate_bpr_fwd_bwd": (ctypes.c_int, [_P, _P, _P, _P, c_int32, _P, c_int32, c_int32, c_int32, _P, _P, c_int32, c_int32,
                                                c_float, c_float, c_float, _P, _P, _P, _P, c_size_t, _P]),
    "lgcn_score_topk_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "lgcn_score_topk": (ctypes.c_int, [_P, _P, _P, c_int32, c_int32, c_int32, _P, _P, c_int32, c_int32,
                                       _P, _P, _P, c_size_t, _P]),
    "lgcn_score_topk_tc_supported": (ctypes.c_int, [c_int32, c_int32]),
    "lgcn_score_topk_tc_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "lgcn_score_topk_tc": (ctypes.c_int, [_P, _P, _P, c_int32, c_int32, c_int32, _P, _P, c_int32, c_int32,
                                          _P, _P, _P, _P, _P, c_size_t, _P]),
    "lgcn_score_topk_tc_debug_layout": (ctypes.c_int, [c_int32, c_int32, _P]),
    "lgcn_score_topk_tc_host_position": (c_int32, [c_int32, c_int32]),
    "lgcn_score_topk_tc_host_item": (c_int32, [c_int32, c_int32]),
    "lgcn_score_topk_tc_position_space": (c_int32, [c_int32]),
    "lgcn_score_topk_tc_item_positions": (ctypes.c_int, [_P, c_int64, c_int32, _P, _P]),
    "lgcn_score_dense": (ctypes.c_int, [_P, _P, _P, c_int32, c_int32, c_int32, _P, _P]),
    "lgcn_score_dense_tc_supported": (ctypes.c_int, [c_int32]),
    "lgcn_score_dense_tc_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "lgcn_score_dense_tc": (ctypes.c_int, [_P, _P, _P, c_int32, c_int32, c_int32, _P, _P, c_size_t, _P]),
    "lgcn_rank_metrics_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "lgcn_rank_metrics": (ctypes.c_int, [_P, c_int32, c_int32, _P, _P, _P, c_int32, _P, _P, c_size_t, _P]),
    "lgcn_sample_bpr": (ctypes.c_int, [_P, _P, c_int32, c_int32, c_int64, ctypes.c_uint64, ctypes.c_uint64, _P, _P, _P, _P, _P]),
    "lgcn_sampler_seed": (None, [c_uint32]),
    "lgcn_sampler_get_state": (None, [_P]),
    "lgcn_sampler_set_state": (ctypes.c_int, [_P]),
    "lgcn_sample_negative": (c_int64, [c_int32, c_int32, c_int64, _P, _P, c_int32, _P]),
    "lgcn_sample_negative_by_user": (c_int64, [_P, c_int64, c_int32, c_int32, _P, _P, c_int32, _P]),
    "lgcn_randint": (c_int32, [c_int32]),
    "lgcn_parse_interactions": (c_int64, [ctypes.c_char_p, _P, _P, c_int64, _P, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load the shared library (once). Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C <package>/csrc`). There is no CPU or PyTorch fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.lgcn_abi_version() != 1:
        raise RuntimeError(f"liblgcn_b200 ABI version {lib.lgcn_abi_version()} != 1")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().lgcn_last_error()
        raise RuntimeError(f"liblgcn_b200 {what}: {msg.decode() if msg else 'error'} (rc={rc})")
