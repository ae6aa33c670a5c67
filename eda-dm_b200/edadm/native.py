"""ctypes binding of libedadm.so (C ABI declared in include/edadm.h).

The library is built in-tree by eda-dm_b200/build.py; there is no CPU fallback -- if the shared
object is missing or a call fails the caller gets an exception.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libedadm.so")


class EdadmError(RuntimeError):
    pass


P = c_void_p
_SIGNATURES = {
    "edadm_last_error": (c_char_p, []),
    "edadm_abi_version": (c_int, []),
    "edadm_reduce_slots": (c_int, []),
    "edadm_uaq_fwd": (c_int, [P, P, P, P, P, c_int64, c_int64, c_int64, c_int, P, P, c_float, c_uint64, c_uint64, P]),
    "edadm_uaq_bwd": (c_int, [P, P, P, P, c_int64, c_int64, c_int64, c_int, P, P, c_float, c_uint64, c_uint64, P, P, c_int, P, P]),
    "edadm_adaround_fwd": (c_int, [P, P, P, P, c_int64, c_int64, c_int64, c_int, c_int, P, P, P]),
    "edadm_adaround_bwd": (c_int, [P, P, P, P, P, c_int64, c_int64, c_int64, c_int, P, c_int, P]),
    "edadm_adaround_init_alpha": (c_int, [P, P, c_int64, c_int64, c_int64, P, P]),
    "edadm_round_reg": (c_int, [P, c_int64, c_float, c_float, P, P, c_int, P, P]),
    "edadm_lp_loss_fwd": (c_int, [P, P, c_int64, c_float, c_float, P, P, P]),
    "edadm_lp_loss_bwd": (c_int, [P, P, c_int64, c_float, c_float, P, P, P]),
    "edadm_fused_adam": (c_int, [P, c_int, P, P, P, P, P, c_double, c_double, c_float, c_int, P]),
    "edadm_mse_search_scores": (c_int, [P, c_int64, c_int64, P, P, c_int, c_int, c_float, P, P]),
    "edadm_act_quant_nhwc": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, c_int, c_int, P, P, c_int, c_float, c_int64, P]),
    "edadm_act_quant_rows": (c_int, [P, P, P, c_int64, c_int, c_int, P, P, c_int, c_int, P, P, c_int, c_float, c_int, c_int64, P]),
    "edadm_gn_fold": (c_int, [P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, c_float, P, P, P]),
    "edadm_norm_act_quant_nhwc": (c_int, [P, P, P, c_int, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, c_int, c_int, P, P, c_int, P]),
    "edadm_im2col_u8": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "edadm_conv_rowsum": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "edadm_pack_weight": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P]),
    "edadm_qgemm_i8": (c_int, [P, c_int, c_int, c_int, c_int, c_int, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, P]),
    "edadm_qgemm_i8_codes": (c_int, [P, c_int64, c_int, P, c_int, c_int, c_int, P, P, P, P, P, P, P, c_int, P, P, c_int, P, c_int, P, P]),
    "edadm_split_bf16": (c_int, [P, c_int64, c_int64, P, P, c_int64, P, P, c_int64, P]),
    "edadm_gemm_bf16x3": (c_int, [P, P, P, P, c_int64, c_int, c_int64, c_int64, P, P, c_int, P]),
    "edadm_split_bf16_batched": (c_int, [P, c_int64, c_int64, c_int64, P, P, c_int64, P, P, c_int64, P]),
    "edadm_gn_fold_cat": (c_int, [P, c_int, P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, c_float, P, P, P]),
    "edadm_act_quant_nhwc_slice": (c_int, [P, P, P, c_int, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P]),
    "edadm_qgemm_i8_split": (c_int, [P, c_int, c_int, c_int, c_int, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P,
                                     P, P, P, P, c_int, P]),
    "edadm_qgemm_i8_rows_post": (c_int, [P, c_int64, c_int, P, c_int, c_int, c_int, P, P, P, P, P, P, P, c_int, P, P]),
    "edadm_split_filter_bf16": (c_int, [P, c_int, c_int, c_int, c_int, P, P, c_int64, P, P, c_int64, P]),
    "edadm_split_nhwc_bf16": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "edadm_conv_bf16x3": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P, c_int, c_int, c_int, c_int, c_int64, P, P, P]),
    "edadm_split_shift_bf16": (c_int, [P, P, P, c_int64, c_int, c_int, c_int, P]),
    "edadm_conv_wgrad_bf16x3": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, P]),
    "edadm_gemm_bf16x3_grouped": (c_int, [P, P, P, P, c_int64, c_int64, c_int, c_int64, c_int64, P, P, c_int, P]),
    "edadm_norm_act_pool2": (c_int, [P, P, P, c_int, P, c_int, c_int, c_int, c_int, P]),
    "edadm_upsample2x_codes": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P]),
    "edadm_layernorm_quant_rows": (c_int, [P, P, P, c_float, P, P, c_int64, c_int, c_int, P, P, c_int, P]),
    "edadm_layernorm_quant_rows_multi": (c_int, [P, P, P, c_float, c_int, P, P, P, P, P, c_int64, c_int, c_int, P]),
    "edadm_geglu_quant_rows": (c_int, [P, P, P, c_int64, c_int, c_int, P, P, c_int, P]),
    "edadm_pack_weight_w4": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P]),
    "edadm_qgemm_w4a8": (c_int, [P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, c_int, c_int, c_int, P]),
    "edadm_conv3x3_small_n": (c_int, [P, P, P, P, P, c_int, P, c_int, c_int, c_int, c_int, c_int, P]),
    "edadm_qattn_fwd": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, c_int,
                                c_float, P, c_int64, c_int64, c_int64, c_int64, P]),
}

_lib = None

# kernels each entry point launches (for bench.py's gpu_launches accounting)
KERNELS_PER_CALL = {"uaq_fwd": 1, "uaq_bwd": 1, "adaround_fwd": 1, "adaround_bwd": 1, "adaround_init_alpha": 1,
                    "round_reg": 2, "lp_loss_fwd": 2, "lp_loss_bwd": 1, "act_quant_nhwc": 1, "gn_fold": 1, "norm_act_quant_nhwc": 1, "conv3x3_small_n": 1, "layernorm_quant_rows": 1, "layernorm_quant_rows_multi": 1, "norm_act_pool2": 1, "upsample2x_codes": 1, "geglu_quant_rows": 1, "act_quant_rows": 1,
                    "im2col_u8": 1, "conv_rowsum": 1, "pack_weight": 1, "pack_weight_w4": 1, "qgemm_i8": 1, "qgemm_i8_codes": 1, "qgemm_w4a8": 1, "qattn_fwd": 1}
launch_counter = {"kernels": 0, "calls": {}}


def exported_symbols():
    """Names every build of the library must export (checked by the CPU test-suite)."""
    return sorted(_SIGNATURES)


def load_library(path: str = None):
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise EdadmError(
            f"{path} not found: build it with `python eda-dm_b200/build.py` "
            "(or __graft_entry__.build()); there is no CPU fallback for the quantized path")
    handle = ctypes.CDLL(path)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(handle, name, None)
        if fn is None:
            raise EdadmError(f"{path} does not export {name}")
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = handle
    return handle


class _Lib:
    """Call-through proxy that turns non-zero return codes into EdadmError."""

    def __getattr__(self, name):
        handle = load_library()
        fn = getattr(handle, "edadm_" + name)

        if fn.restype is not c_int or name in ("abi_version", "reduce_slots"):
            return fn

        n_kernels = KERNELS_PER_CALL.get(name, 1)

        def call(*args):
            launch_counter["kernels"] += n_kernels
            launch_counter["calls"][name] = launch_counter["calls"].get(name, 0) + 1
            rc = fn(*args)
            if rc != 0:
                msg = handle.edadm_last_error()
                raise EdadmError(f"edadm_{name} failed ({rc}): {msg.decode() if msg else ''}")

        call.__name__ = name
        object.__setattr__(self, name, call)
        return call


lib = _Lib()
