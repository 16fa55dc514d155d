"""ctypes binding of include/mimi_b200.h. The product path fails loudly if the extension is missing."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

c_void_p = C.c_void_p

LAYER_FIELDS = (
    "input_layernorm_weight", "input_layernorm_bias", "q_proj_weight", "k_proj_weight", "v_proj_weight",
    "o_proj_weight", "self_attn_layer_scale", "post_attention_layernorm_weight",
    "post_attention_layernorm_bias", "fc1_weight", "fc2_weight", "mlp_layer_scale",
)


class LayerWeights(C.Structure):
    _fields_ = [(n, c_void_p) for n in LAYER_FIELDS]


class Weights(C.Structure):
    _fields_ = [
        ("conv_weight", c_void_p * 14),
        ("conv_bias", c_void_p * 14),
        ("layer", LayerWeights * 8),
        ("downsample_weight", c_void_p),
        ("semantic_input_proj_weight", c_void_p),
        ("acoustic_input_proj_weight", c_void_p),
        ("embed_sum", c_void_p * 32),
        ("cluster_usage", c_void_p * 32),
        ("rope_inv_freq", c_void_p),
    ]


class DecoderWeights(C.Structure):
    _fields_ = [
        ("semantic_output_proj_weight", c_void_p),
        ("acoustic_output_proj_weight", c_void_p),
        ("upsample_weight", c_void_p),
        ("layer", LayerWeights * 8),
        ("conv_in_weight", c_void_p),
        ("conv_in_bias", c_void_p),
        ("up_weight", c_void_p * 4),
        ("up_bias", c_void_p * 4),
        ("res_a_weight", c_void_p * 4),
        ("res_a_bias", c_void_p * 4),
        ("res_b_weight", c_void_p * 4),
        ("res_b_bias", c_void_p * 4),
        ("conv_out_weight", c_void_p),
        ("conv_out_bias", c_void_p),
    ]


PHASE_BEGIN, PHASE_FRONT, PHASE_FINISH = 1, 2, 3      # MIMI_B200_PHASE_* (include/mimi_b200.h)

# every symbol include/mimi_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mimi_b200_abi_version": (C.c_int, []),
    "mimi_b200_create": (C.c_int, [C.POINTER(c_void_p), C.c_int]),
    "mimi_b200_destroy": (None, [c_void_p]),
    "mimi_b200_last_error": (C.c_char_p, [c_void_p]),
    "mimi_b200_load_weights": (C.c_int, [c_void_p, C.POINTER(Weights)]),
    "mimi_b200_encoded_frames": (C.c_int64, [C.c_int64]),
    "mimi_b200_workspace_bytes": (C.c_int, [c_void_p, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_size_t)]),
    "mimi_b200_encode": (C.c_int, [c_void_p, c_void_p, C.c_int, C.c_int64, c_void_p, C.c_int, c_void_p,
                                   c_void_p, c_void_p, C.c_size_t, c_void_p]),
    "mimi_b200_encode_phase": (C.c_int, [c_void_p, C.c_int, C.c_int, C.c_int, c_void_p, C.c_int, C.c_int64, c_void_p, C.c_int,
                                         c_void_p, c_void_p, c_void_p, C.c_size_t, c_void_p]),
    "mimi_b200_host_pack": (C.c_int, [c_void_p, C.c_int64, c_void_p, c_void_p, c_void_p, C.c_int, C.c_int]),
    "mimi_b200_debug_tap": (C.c_int, [c_void_p, C.c_int, c_void_p, C.c_size_t, C.POINTER(C.c_int64),
                                      C.POINTER(C.c_int), c_void_p]),
    "mimi_b200_debug_set": (C.c_int, [c_void_p, C.c_int, C.c_int]),
    "mimi_b200_profile_read": (C.c_int, [c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "mimi_b200_debug_tc_gemm": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                          c_void_p, c_void_p]),
    "mimi_b200_debug_shift_probe": (C.c_int, [c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, c_void_p, c_void_p]),
    "mimi_b200_resample_out_len": (C.c_int64, [C.c_int64, C.c_int, C.c_int]),
    "mimi_b200_resample": (C.c_int, [c_void_p, c_void_p, C.c_int64, c_void_p, C.c_int, C.c_int, C.c_int,
                                     c_void_p, C.c_int64, c_void_p]),
    "mimi_b200_utf8_bytes_per_frame": (C.c_int64, [C.c_int, C.c_uint32, C.c_int]),
    "mimi_b200_codes_to_utf8": (C.c_int, [c_void_p, c_void_p, C.c_int, C.c_int, C.c_int64, c_void_p,
                                          C.c_uint32, C.c_int, c_void_p, C.c_int64, c_void_p, c_void_p]),
    "mimi_b200_load_decoder_weights": (C.c_int, [c_void_p, C.POINTER(DecoderWeights)]),
    "mimi_b200_decode_workspace_bytes": (C.c_int, [c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_size_t)]),
    "mimi_b200_decode": (C.c_int, [c_void_p, c_void_p, C.c_int, C.c_int, C.c_int64, c_void_p, c_void_p, C.c_size_t, c_void_p]),
    "mimi_b200_codes_pack_u16": (C.c_int, [c_void_p, c_void_p, C.c_int64, c_void_p, c_void_p]),
    "mimi_b200_range_overflow": (C.c_int, [c_void_p, C.c_int]),
    "mimi_b200_launch_count": (C.c_int64, [c_void_p]),
}

_lib = None


class MimiB200Error(RuntimeError):
    pass


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen libmimi_b200.so (building it in-tree first if it is missing or stale and nvcc exists)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    alt = os.environ.get("MIMI_B200_LIB")        # A/B of two builds on one box (tools/ab.sh): an explicit library file
    if path is None and alt:
        path = alt
    p = path or _build.LIB_PATH
    if path is None:
        try:
            _build.build()
        except Exception as e:  # stale-but-present library is still usable; a missing one is fatal
            if not os.path.exists(p):
                raise MimiB200Error(f"libmimi_b200.so is missing and could not be built: {e}") from e
            import warnings
            warnings.warn(f"libmimi_b200.so is older than its sources and could not be rebuilt ({e}); loading the existing "
                          "library", RuntimeWarning, stacklevel=2)
    if not os.path.exists(p):
        raise MimiB200Error(f"{p} not found: the CUDA extension is required (there is no CPU fallback)")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)           # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.mimi_b200_abi_version() != 1:
        raise MimiB200Error("libmimi_b200.so ABI version mismatch")
    if path is None or path == alt:
        _lib = lib
    return lib


def check(lib: C.CDLL, handle, rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.mimi_b200_last_error(handle)
        raise MimiB200Error(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")
