"""ctypes binding of libmillion_b200.so (the C ABI declared in include/million_b200.h).

There is no fallback: if the library is missing the import of any op raises, loudly.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MILLION_B200_LIB: development override (A/B runs of two builds on the same box); the default is the in-tree build
LIB_PATH = os.environ.get("MILLION_B200_LIB") or os.path.join(HERE, "libmillion_b200.so")

MILLION_F16, MILLION_BF16, MILLION_F32 = 0, 1, 2
MILLION_OK, MILLION_ERR_INVALID, MILLION_ERR_UNSUPPORTED, MILLION_ERR_CUDA = 0, 1, 2, 3
V_ROWMAJOR, V_TRANSPOSED, V_PAGED = 0, 1, 2
IMPL_AUTO, IMPL_GENERIC, IMPL_FAST, IMPL_GRID = 0, 1, 2, 3
ATTN_PARTIAL_ONLY = 1
ATTN_FUSED_SPLITKV = 2
ATTN_PDL = 4
ABI_VERSION = 8

c_i32, c_i64, c_u32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_void_p


class MillionError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libmillion_b200 status {status}: {msg}")
        self.status = status


class AttnParams(ctypes.Structure):
    """Mirror of `million_attn_params` (include/million_b200.h) — field order is the ABI."""
    _fields_ = [
        ("struct_size", c_u32), ("io_dtype", c_i32), ("impl", c_i32), ("flags", c_i32),
        ("bs", c_i32), ("nh", c_i32), ("nh_k", c_i32), ("d", c_i32), ("M", c_i32), ("C", c_i32),
        ("nk", c_i32), ("r", c_i32),
        ("q", c_vp), ("k_codes", c_vp), ("k_head_stride", c_i64),
        ("v_layout", c_i32), ("page_size", c_i32), ("v_codes", c_vp), ("v_head_stride", c_i64), ("v_ld", c_i64),
        ("v_page_ids", c_vp), ("n_pages", c_i32), ("res_len", c_i32),
        ("k_cent", c_vp), ("v_cent", c_vp), ("k_res", c_vp), ("v_res", c_vp),
        ("out", c_vp), ("workspace", c_vp), ("workspace_bytes", c_i64), ("n_splits", c_i32),
        ("partial", c_vp), ("prepared_codebook", c_vp),
        ("k_out", c_i32), ("v_out", c_i32),
        ("k_out_idx", c_vp), ("k_out_val", c_vp), ("k_out_head_stride", c_i64),
        ("v_out_idx", c_vp), ("v_out_val", c_vp), ("v_out_head_stride", c_i64),
        ("p2p_state", c_vp),
        ("k_new", c_vp), ("v_new", c_vp), ("r_dev", c_vp),
        ("code_bytes", c_i32), ("reserved0", c_i32),
    ]


# symbol -> (restype, argtypes); tests/test_abi.py checks every symbol of the header is here and exported
SIGNATURES = {
    "million_abi_version": (ctypes.c_int, []),
    "million_last_error": (ctypes.c_char_p, []),
    "million_device_info": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)] * 3),
    "million_pq_encoder_prepared_bytes": (c_i64, [ctypes.c_int] * 3),
    "million_pq_encoder_prepare": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "million_pq_encoder_grid_prepared_bytes": (c_i64, [ctypes.c_int] * 3),
    "million_pq_encoder_grid_prepare": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "million_pq_encoder_auto_prepared_bytes": (c_i64, [ctypes.c_int] * 3),
    "million_pq_encoder_auto_prepare": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "million_pq_encode": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, c_vp, c_vp, c_vp, ctypes.c_int, c_i64, c_i64, c_i64, c_i64,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]),
    "million_pq_encode_paged": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_int, c_i64,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]),
    "million_pq_decode": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, c_i64, c_i64, c_vp, c_vp, ctypes.c_int, c_i64,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]),
    "million_pq_outlier_split": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]),
    "million_pq_outlier_apply": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, c_vp, c_vp, ctypes.c_int, c_i64, c_i64,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]),
    "million_pq_codebook_prepared_bytes": (c_i64, [ctypes.c_int] * 3),
    "million_pq_codebook_prepare": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "million_pq_decode_attn_workspace_bytes": (c_i64, [ctypes.c_int] * 5),
    "million_pq_decode_attn_default_splits": (ctypes.c_int, [ctypes.c_int] * 3),
    "million_pq_decode_attn": (ctypes.c_int, [ctypes.POINTER(AttnParams), c_vp]),
    "million_lse_merge": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, ctypes.c_int, c_vp, ctypes.c_int, c_vp]),
    "million_splitkv_symmetric_bytes": (c_i64, [ctypes.c_int, c_i64, ctypes.c_int]),
    "million_splitkv_state_bytes": (c_i64, []),
    "million_splitkv_state_init": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_i64, c_vp]),
    "million_splitkv_set_timeout": (ctypes.c_int, [c_vp, c_i64, c_vp]),
    "million_splitkv_push_merge": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_i64, ctypes.c_int, c_vp, ctypes.c_int, c_vp, ctypes.c_int, c_vp]),
    "million_window_append": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, c_vp]),
    "million_counter_add": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_vp]),
    "million_rope_qk": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp] + [ctypes.c_int] * 5 + [c_vp]),
    "million_window_shift": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, c_vp]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension is the product and there is no fallback. "
                "Build it with `python -m million_b200._build` (or `__graft_entry__.build()`).")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.million_abi_version() != ABI_VERSION:
            raise ImportError(f"libmillion_b200 ABI {handle.million_abi_version()} != expected {ABI_VERSION}; rebuild")
        _lib = handle
    return _lib


def check(status):
    if status != MILLION_OK:
        raise MillionError(status, lib().million_last_error().decode(errors="replace"))
