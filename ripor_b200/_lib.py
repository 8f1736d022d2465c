"""ctypes binding of libriporb200.so (the C ABI declared in include/riporb200.h).

The product path has no CPU fallback: if the library is missing this module raises at import of the
first symbol, and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libriporb200.so")

c_i32, c_i64, c_f32, c_f64, c_void_p, c_char_p = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_void_p, C.c_char_p
P = C.POINTER


class TrieInfo(C.Structure):
    _fields_ = [("n_docs", c_i64), ("n_unique", c_i64), ("n_nodes", c_i64), ("n_children", c_i64),
                ("bytes", c_i64), ("L", c_i32), ("V", c_i32), ("code_bytes", c_i32), ("on_device", c_i32)]


class EngineConfig(C.Structure):
    _fields_ = [("d_model", c_i32), ("num_heads", c_i32), ("d_kv", c_i32), ("d_ff", c_i32),
                ("num_layers", c_i32), ("num_decoder_layers", c_i32), ("vocab_size", c_i32),
                ("num_buckets", c_i32), ("max_distance", c_i32), ("layer_norm_eps", c_f32),
                ("decoder_vocab_size", c_i32), ("docid_len", c_i32), ("shared_output_input_embeds", c_i32),
                ("scaleup_output_hidden", c_i32), ("max_batch", c_i32), ("max_beams", c_i32),
                ("max_src_len", c_i32), ("precision", c_i32), ("device", c_i32)]


PRECISIONS = {"fp32": 0, "tf32x3": 1, "bf16x3": 2, "tf32": 3, "bf16": 4, "fp16x3": 5}

# name -> (restype, argtypes); every symbol include/riporb200.h declares.
SIGNATURES = {
    "rb200_version": (c_char_p, []),
    "rb200_last_error": (c_char_p, []),
    "rb200_trie_build": (C.c_int, [c_void_p, C.c_int, c_i64, C.c_int, C.c_int, C.c_int, P(c_void_p)]),
    "rb200_trie_free": (C.c_int, [c_void_p]),
    "rb200_trie_get_info": (C.c_int, [c_void_p, P(TrieInfo)]),
    "rb200_trie_level_counts": (C.c_int, [c_void_p, c_void_p]),
    "rb200_trie_save": (C.c_int, [c_void_p, c_char_p]),
    "rb200_trie_load": (C.c_int, [c_char_p, P(c_void_p)]),
    "rb200_trie_save_tagged": (C.c_int, [c_void_p, c_char_p, C.c_uint64]),
    "rb200_trie_load_tagged": (C.c_int, [c_char_p, P(C.c_uint64), P(c_void_p)]),
    "rb200_trie_leaf_expand": (C.c_int, [c_void_p, c_void_p, c_i64, C.c_int, c_void_p, c_void_p, c_void_p]),
    "rb200_docid_json_open": (C.c_int, [c_char_p, C.c_int, P(c_void_p)]),
    "rb200_docid_json_free": (C.c_int, [c_void_p]),
    "rb200_docid_json_info": (C.c_int, [c_void_p, P(c_i64), P(c_i32), P(c_i32), P(c_i64)]),
    "rb200_docid_json_codes": (C.c_int, [c_void_p, P(c_void_p)]),
    "rb200_docid_json_keys": (C.c_int, [c_void_p, P(c_void_p), P(c_void_p)]),
    "rb200_trie_build_from_table": (C.c_int, [c_void_p, C.c_int, C.c_int, P(c_void_p)]),
    "rb200_unpack_codes": (C.c_int, [c_void_p, c_i64, c_i64, C.c_int, C.c_int, c_void_p]),
    "rb200_trie_mask_host": (C.c_int, [c_void_p, c_void_p, c_i64, C.c_int, c_void_p]),
    "rb200_trie_leaf_docs": (C.c_int, [c_void_p, c_i64, P(c_void_p), P(c_i64)]),
    "rb200_trie_find_leaf": (C.c_int, [c_void_p, c_void_p, P(c_i64)]),
    "rb200_trie_upload": (C.c_int, [c_void_p, C.c_int]),
    "rb200_trie_mask_device": (C.c_int, [c_void_p, c_void_p, c_i64, C.c_int, c_void_p, c_void_p]),
    "rb200_beam_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P(c_void_p)]),
    "rb200_beam_free": (C.c_int, [c_void_p]),
    "rb200_beam_reset": (C.c_int, [c_void_p, c_void_p, C.c_int, c_void_p]),
    "rb200_beam_reset_beams": (C.c_int, [c_void_p, c_void_p, C.c_int, C.c_int, c_void_p]),
    "rb200_beam_forced_tail": (C.c_int, [c_void_p, c_void_p, C.c_int, c_void_p, C.c_int, c_void_p]),
    "rb200_beam_step": (C.c_int, [c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, c_void_p, c_void_p, C.c_int,
                                  c_void_p]),
    "rb200_beam_finalize": (C.c_int, [c_void_p, c_void_p, C.c_int, c_f64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rb200_beam_view": (C.c_int, [c_void_p, C.c_int, P(c_void_p)]),
    "rb200_beam_current_step": (C.c_int, [c_void_p]),
    "rb200_engine_create": (C.c_int, [P(EngineConfig), P(c_void_p)]),
    "rb200_engine_free": (C.c_int, [c_void_p]),
    "rb200_engine_resize": (C.c_int, [c_void_p, C.c_int, C.c_int, C.c_int]),
    "rb200_engine_next_input": (C.c_int, [c_void_p, P(c_void_p)]),
    "rb200_engine_forward": (C.c_int, [c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, c_void_p, C.c_int,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
    "rb200_engine_last_freeze_histogram": (C.c_int, [c_void_p, c_void_p, C.c_int, P(c_i64)]),
    "rb200_engine_set_weight": (C.c_int, [c_void_p, c_char_p, c_void_p, c_i64, c_void_p]),
    "rb200_engine_finalize_weights": (C.c_int, [c_void_p, c_void_p]),
    "rb200_engine_workspace_bytes": (c_i64, [c_void_p]),
    "rb200_engine_search": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rb200_engine_search_host": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rb200_engine_encode": (C.c_int, [c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, c_void_p]),
    "rb200_engine_encoder_states": (C.c_int, [c_void_p, P(c_void_p)]),
    "rb200_engine_decode_step": (C.c_int, [c_void_p, c_void_p, C.c_int, c_void_p, c_void_p]),
    "rb200_engine_beam": (C.c_int, [c_void_p, P(c_void_p)]),
    "rb200_engine_last_launch_count": (c_i64, [c_void_p]),
    "rb200_engine_last_tail_step": (C.c_int, [c_void_p]),
    "rb200_engine_set_profiling": (C.c_int, [c_void_p, C.c_int]),
    "rb200_engine_get_profile": (C.c_int, [c_void_p, P(c_f64), P(c_f64), P(c_i64)]),
    "rb200_engine_get_profile_bytes": (C.c_int, [c_void_p, P(c_f64)]),
    "rb200_relative_position_bucket": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "rb200_gemm_bench": (C.c_int, [C.c_int, c_i64, c_i64, c_i64, C.c_int, C.c_int, C.c_int, P(c_f64), c_void_p]),
    "rb200_gemm": (C.c_int, [C.c_int, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, C.c_int, C.c_int, c_void_p]),
}

_lib = None


class RB200Error(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the C-ABI library; fails loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RB200Error(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU or PyTorch fallback for this path)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError here = header and library out of sync
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(status: int) -> None:
    if status != 0:
        msg = lib().rb200_last_error().decode("utf-8", "replace")
        if status == -1:
            raise ValueError(msg)
        raise RB200Error(f"riporb200 error {status}: {msg}")


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)
