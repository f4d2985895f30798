"""ctypes binding of ``libuvlt_sm100.so`` (C ABI declared in ``include/uvlt.h``).

The library is the product: there is no CPU or PyTorch fallback.  ``load()`` raises if the shared object is
missing, and every wrapper raises ``RuntimeError`` with ``uvlt_last_error()`` when a call returns non-zero.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# UVLT_LIB: another build of the same C ABI (A/B timing of two library builds on one box, tools/r2_gpu_ab.sh)
LIB_PATH = os.environ.get("UVLT_LIB") or os.path.join(_HERE, "libuvlt_sm100.so")

_lib = None
_lock = threading.Lock()

c_void_p, c_int, c_int32, c_int64, c_float, c_longlong = C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_float, C.c_longlong


class UvltConfig(C.Structure):
    """Mirror of ``struct uvlt_config`` (include/uvlt.h)."""

    _fields_ = [
        ("embed_dim", c_int32), ("num_heads", c_int32), ("depth", c_int32), ("mlp_hidden", c_int32),
        ("template_size", c_int32), ("search_size", c_int32), ("text_len", c_int32), ("fusion_start", c_int32),
        ("head_channels", c_int32), ("vocab_size", c_int32), ("max_position", c_int32), ("max_batch", c_int32),
        ("softmax_one", c_int32), ("offset_sigmoid", c_int32), ("txt_token_mean", c_int32),
        ("num_cont_layers", c_int32), ("cont_layers", c_int32 * 32),
    ]


class UvltOutputs(C.Structure):
    """Mirror of ``struct uvlt_outputs`` (include/uvlt.h)."""

    _fields_ = [
        ("tokens", c_void_p), ("cls_score", c_void_p), ("bbox_map", c_void_p), ("pred_boxes", c_void_p),
        ("cont_score", c_void_p), ("cont_prob", c_void_p), ("logits", c_void_p), ("prompts", c_void_p),
        ("batch", c_int32), ("n_tokens", c_int32), ("embed_dim", c_int32), ("feat_size", c_int32),
        ("cont_cols", c_int32), ("reserved", c_int32),
    ]


WANT_LOGITS = 1
SKIP_TEXT = 2
TEXT_CACHED = 4
FRAME_SLOT1 = 8
NO_SYNC = 16

_P = c_void_p
# name -> (restype, argtypes); must list every symbol include/uvlt.h declares (tests/test_cabi_symbols.py checks)
SIGNATURES = {
    "uvlt_abi_version": (c_int, []),
    "uvlt_last_error": (C.c_char_p, []),
    "uvlt_create": (c_int, [C.POINTER(UvltConfig), C.POINTER(c_void_p)]),
    "uvlt_destroy": (None, [c_void_p]),
    "uvlt_set_weight": (c_int, [c_void_p, C.c_char_p, c_void_p, C.POINTER(c_int64), c_int32]),
    "uvlt_finalize_weights": (c_int, [c_void_p]),
    "uvlt_set_option": (c_int, [c_void_p, C.c_char_p, c_int32]),
    "uvlt_forward_test": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int32, c_int32, C.POINTER(UvltOutputs), _P]),
    "uvlt_forward_train": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int32, c_int32, C.POINTER(UvltOutputs), _P]),
    "uvlt_backbone": (c_int, [_P, _P, _P, _P, _P, _P, c_int32, c_int32, C.POINTER(UvltOutputs), _P]),
    "uvlt_forward_prompt": (c_int, [_P, _P, _P, _P, _P, _P, c_int32, _P, _P]),
    "uvlt_head": (c_int, [_P, _P, _P, _P, c_int32, C.POINTER(UvltOutputs), _P]),
    "uvlt_track_decode": (c_int, [_P, _P, c_int32, _P, _P, _P, _P]),
    "uvlt_track_frame_host": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int32, c_int32, c_int32, _P, _P, _P, _P]),
    "uvlt_track_frame_image_host": (c_int, [_P, _P, c_int32, c_int32, _P, C.c_double, _P, _P, _P, _P, _P, _P, c_int32,
                                            c_int32, c_int32, _P, _P, _P, _P]),
    "uvlt_stream_sync": (c_int, [_P]),
    "uvlt_step_wait": (c_int, [_P, c_int32]),
    "uvlt_op_crop_resize": (c_int, [_P, c_int32, c_int32, _P, C.c_double, c_int32, _P, _P, c_int32, _P]),
    "uvlt_op_box_update": (c_int, [_P, _P, c_int32, c_int32, c_int32, _P, c_int32, _P]),
    "uvlt_op_anno2mask": (c_int, [_P, c_int32, _P, c_int32, _P]),
    "uvlt_op_normalize_u8": (c_int, [_P, _P, c_int32, c_int32, _P]),
    "uvlt_op_grounding_resize": (c_int, [_P, c_int32, c_int32, c_int32, _P, c_int32, _P]),
    "uvlt_text_encode": (c_int, [_P, _P, _P, _P, c_int32, _P]),
    "uvlt_upload_frames": (c_int, [_P, _P, c_int64, c_int64, c_int64, _P]),
    "uvlt_upload_frames_slot": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int32, _P]),
    "uvlt_upload_frames_2d": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P]),
    "uvlt_last_launch_count": (c_int, [c_void_p]),
    "uvlt_op_gemm": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "uvlt_gemm_plan": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, C.POINTER(c_int32)]),
    "uvlt_runtime_switches": (c_int, [C.POINTER(c_int32)]),
    "uvlt_op_gemm_splitk": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, C.POINTER(c_int), _P]),
    "uvlt_op_gemm_grouped": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_longlong, c_longlong,
                                     c_int, _P]),
    "uvlt_op_attention": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, c_int, _P]),
    "uvlt_op_layernorm": (c_int, [_P, c_longlong, c_int, c_int, _P, _P, c_int, c_int, _P, _P, _P, c_float, c_int,
                                  c_int, _P]),
    "uvlt_op_patch_im2col": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, c_longlong, c_int, _P]),
    "uvlt_op_im2col3x3": (c_int, [_P, c_int, c_longlong, c_longlong, c_longlong, c_int, c_int, c_int, c_int, _P, _P]),
    "uvlt_op_bert_embed": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_longlong, c_int, _P, c_int, c_int, c_int, c_int, _P]),
    "uvlt_op_build_bias": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
}


def load():
    """dlopen the in-tree library and attach prototypes.  Raises ``OSError`` when it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise OSError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C uvltrack_b200/csrc`).  There is no fallback path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing -> loud failure
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int, what: str = "uvlt call") -> None:
    if rc != 0:
        msg = load().uvlt_last_error()
        raise RuntimeError(f"{what} failed (rc={rc}): {msg.decode() if msg else 'unknown error'}")


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def current_stream():
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)
