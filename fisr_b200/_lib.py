"""ctypes binding of ``libfisr_b200.so`` (C ABI in ``include/fisr_b200.h``).

There is no CPU path: if the library is missing, or no sm_100 GPU is present when a context is
created, the import / constructor raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfisr_b200.so")

FISR_OK = 0
FISR_E_OVERFLOW = -5   # non-finite gradient (loss scale too high)
PREC_F16X3 = 0      # fp16 (hi, lo) split operands, fp32-class result (default)
PREC_F16 = 1        # single fp16 operands, fast mode
PREC_F16F8 = 2      # fp16 main term + fp8 cross terms (2 MMA units per K slice), inference only

_lib: Optional[C.CDLL] = None


class FisrError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FisrError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(fisr_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    sig = {
        "fisr_create": (i, [i, C.POINTER(vp)]),
        "fisr_destroy": (None, [vp]),
        "fisr_last_error": (C.c_char_p, [vp]),
        "fisr_set_precision": (i, [vp, i]),
        "fisr_get_precision": (i, [vp]),
        "fisr_num_params": (i, []),
        "fisr_param_name": (C.c_char_p, [i]),
        "fisr_param_shape": (i, [i, C.POINTER(i)]),
        "fisr_set_param": (i, [vp, C.c_char_p, vp, sz]),
        "fisr_get_param": (i, [vp, C.c_char_p, vp, sz]),
        "fisr_forward": (i, [vp, vp, i, i, i, vp, vp, vp, vp]),
        "fisr_forward_host": (i, [vp, vp, i, i, i, vp, vp, vp]),
        "fisr_window_device": (i, [vp, vp, vp, vp, i, i, i, i, i, i, vp, vp]),
        "fisr_units_device": (i, [vp, vp, vp, vp, i, i, i, i, i, C.POINTER(i), i, i, vp, vp]),
        "fisr_window_device_f32": (i, [vp, vp, vp, vp, i, i, i, i, vp, vp]),
        "fisr_window_host": (i, [vp, vp, vp, vp, i, i, i, i, vp]),
        "fisr_window_host_f32": (i, [vp, vp, vp, vp, i, i, i, i, vp]),
        "fisr_window_submit": (i, [vp, i, vp, vp, vp, i, i, i, i, vp]),
        "fisr_window_wait": (i, [vp, i]),
        "fisr_warp_device": (i, [vp, vp, vp, f, vp, i, i, f, vp]),
        "fisr_warp_host": (i, [vp, vp, vp, f, vp, i, i, f]),
        "fisr_warp_batch_device": (i, [vp, vp, vp, vp, i, f, vp, i, i, f, vp]),
        "fisr_groups2ovlp": (i, [vp, vp, i, i, i, vp, vp]),
        "fisr_temporal_loss": (i, [vp, vp, vp, vp, vp, i, i, i, C.POINTER(f), C.POINTER(f), vp]),
        "fisr_train_forward": (i, [vp, vp, vp, vp, vp, vp, vp, i, i, i, C.POINTER(f), C.POINTER(f), vp]),
        "fisr_adam_step": (i, [vp, C.POINTER(vp), i, f, f, f, f]),
        "fisr_adam_steps": (C.c_longlong, [vp]),
        "fisr_adam_reset": (i, [vp, C.c_longlong]),
        "fisr_adam_set_steps": (i, [vp, C.c_longlong]),
        "fisr_get_adam_slot": (i, [vp, C.c_char_p, i, vp, sz]),
        "fisr_set_adam_slot": (i, [vp, C.c_char_p, i, vp, sz]),
        "fisr_conv3x3": (i, [vp, vp, vp, vp, vp, i, i, i, i, i, i, i, vp, vp]),
        "fisr_train_backward": (i, [vp, vp, vp, vp, vp, vp, vp, i, i, i, C.POINTER(f), C.POINTER(f), vp]),
        "fisr_get_grad": (i, [vp, C.c_char_p, vp, sz]),
        "fisr_adam_apply": (i, [vp, f, f, f, f]),
        "fisr_train_step": (i, [vp, vp, vp, vp, vp, vp, vp, i, i, i, C.POINTER(f), f, C.POINTER(f), vp]),
        "fisr_set_loss_scale": (i, [vp, f]),
        "fisr_get_loss_scale": (f, [vp, i, i, i]),
        "fisr_set_wgrad_exact": (i, [vp, i]),
        "fisr_dgrad3x3": (i, [vp, vp, vp, vp, vp, i, i, i, i, i, i, vp, vp]),
        "fisr_profile_train": (i, [vp, i, i, i, i, i, C.POINTER(f), C.POINTER(C.c_double), C.c_char_p, i]),
        "fisr_wgrad3x3": (i, [vp, vp, vp, i, i, i, i, i, f, vp, vp]),
        "fisr_debug_conv_output": (i, [vp, C.c_char_p, vp, sz]),
        "fisr_profile_ops": (i, [vp, i, i, i, i, i, C.POINTER(f), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                 C.POINTER(i), C.c_char_p, i]),
        "fisr_ipc_alloc": (i, [vp, sz, C.POINTER(vp), C.c_char_p]),
        "fisr_ipc_open": (i, [vp, C.c_char_p, C.POINTER(vp)]),
        "fisr_ipc_close": (i, [vp, vp]),
        "fisr_ipc_free": (i, [vp, vp]),
        "fisr_copy2d_async": (i, [vp, vp, sz, vp, sz, sz, sz, vp]),
        "fisr_pwc_create": (i, [i, C.POINTER(vp)]),
        "fisr_pwc_destroy": (None, [vp]),
        "fisr_pwc_last_error": (C.c_char_p, [vp]),
        "fisr_pwc_num_params": (i, []),
        "fisr_pwc_param_name": (C.c_char_p, [i]),
        "fisr_pwc_param_shape": (i, [i, C.POINTER(i)]),
        "fisr_pwc_set_param": (i, [vp, C.c_char_p, vp, sz]),
        "fisr_pwc_forward": (i, [vp, vp, vp, i, i, i, vp, vp]),
        "fisr_pwc_prepare_pair": (i, [vp, vp, vp, i, vp, i, i, i, vp, vp, vp]),
        "fisr_pwc_finish_flow": (i, [vp, vp, i, i, i, i, i, i, i, vp, i, vp, i, C.c_double, vp, vp]),
        "fisr_pwc_debug_flow": (i, [vp, i, vp, sz]),
        "fisr_pwc_launch_count": (C.c_longlong, [vp]),
        "fisr_launch_count": (C.c_longlong, [vp]),
        "fisr_plan_info": (i, [vp, i, i, i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i), C.POINTER(sz)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)       # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EXPORTS = [
    "fisr_create", "fisr_destroy", "fisr_last_error", "fisr_set_precision", "fisr_get_precision", "fisr_num_params",
    "fisr_param_name", "fisr_param_shape", "fisr_set_param", "fisr_get_param", "fisr_forward", "fisr_forward_host",
    "fisr_window_device", "fisr_units_device", "fisr_window_device_f32", "fisr_window_host", "fisr_window_host_f32", "fisr_window_submit", "fisr_window_wait", "fisr_warp_device", "fisr_warp_host", "fisr_warp_batch_device",
    "fisr_groups2ovlp", "fisr_temporal_loss", "fisr_train_forward", "fisr_adam_step", "fisr_adam_steps", "fisr_adam_reset", "fisr_adam_set_steps", "fisr_get_adam_slot", "fisr_set_adam_slot",
    "fisr_train_backward", "fisr_get_grad", "fisr_adam_apply", "fisr_train_step", "fisr_set_loss_scale", "fisr_get_loss_scale", "fisr_set_wgrad_exact",
    "fisr_dgrad3x3", "fisr_profile_train", "fisr_conv3x3", "fisr_wgrad3x3", "fisr_debug_conv_output", "fisr_profile_ops", "fisr_launch_count", "fisr_plan_info",
    "fisr_pwc_create", "fisr_pwc_destroy", "fisr_pwc_last_error", "fisr_pwc_num_params", "fisr_pwc_param_name", "fisr_pwc_param_shape",
    "fisr_pwc_set_param", "fisr_pwc_forward", "fisr_pwc_prepare_pair", "fisr_pwc_finish_flow", "fisr_pwc_debug_flow", "fisr_pwc_launch_count",
    "fisr_ipc_alloc", "fisr_ipc_open", "fisr_ipc_close", "fisr_ipc_free", "fisr_copy2d_async",
]
