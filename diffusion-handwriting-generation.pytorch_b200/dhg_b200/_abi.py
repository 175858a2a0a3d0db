"""ctypes binding of include/dhg_b200.h.  There is no fallback: if the library
is missing or cannot be loaded, importing the compute entry points raises."""
import ctypes
import os

from .build import LIB_PATH

_lib = None

c_i32, c_i64, c_u64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64
c_vp, c_cp = ctypes.c_void_p, ctypes.c_char_p


class DhgConfig(ctypes.Structure):
    _fields_ = [("num_layers", c_i32), ("channels", c_i32)]


class DebugEpilogue(ctypes.Structure):
    """dhg_debug_epilogue of include/dhg_b200.h (test hook)."""
    _fields_ = [
        ("bias", c_vp), ("rowbias", c_vp), ("rowbias_cols", c_i32), ("res_pre", c_vp), ("res_pre_pitch", c_i32), ("ln", c_i32),
        ("gamma", c_vp), ("beta", c_vp), ("film_bstride", c_i32), ("res_post", c_vp), ("res_post_pitch", c_i32),
        ("res_post_up", c_i32), ("res_post_period_lo", c_i32), ("out_raw", c_vp), ("out_raw_pitch", c_i32),
        ("out_act", c_vp), ("out_act_pitch", c_i32), ("period", c_i32), ("pad_first", c_i32), ("nvalid", c_i32),
        ("dot_w", c_vp), ("dot_out", c_vp), ("dot_act", c_i32), ("split_io", c_i32),
        ("dual_a2", c_vp), ("dual_lda2", c_i32), ("dual_K2", c_i32), ("dual_w2", c_vp), ("dual_w1_rows", c_i32), ("w_row_off", c_i32),
    ]


class DebugAttn(ctypes.Structure):
    """dhg_debug_attn of include/dhg_b200.h (test hook)."""
    _fields_ = [
        ("q", c_vp), ("k", c_vp), ("v", c_vp), ("o", c_vp),
        ("q_pitch", c_i32), ("k_pitch", c_i32), ("v_pitch", c_i32), ("o_pitch", c_i32),
        ("q_period", c_i32), ("q_pad", c_i32), ("k_period", c_i32), ("k_pad", c_i32),
        ("B", c_i32), ("H", c_i32), ("D", c_i32), ("Tq", c_i32), ("Tk", c_i32),
        ("q_rows", c_i32), ("k_rows", c_i32), ("text", c_vp),
    ]


# name -> (restype, argtypes); every symbol include/dhg_b200.h declares
SIGNATURES = {
    "dhg_last_error": (c_cp, []),
    "dhg_abi_version": (c_i32, []),
    "dhg_create": (c_i32, [c_i32, ctypes.POINTER(DhgConfig), ctypes.POINTER(c_vp)]),
    "dhg_destroy": (c_i32, [c_vp]),
    "dhg_num_weights": (c_i32, [c_vp]),
    "dhg_weight_name": (c_cp, [c_vp, c_i32]),
    "dhg_weight_ndim": (c_i32, [c_vp, c_i32]),
    "dhg_weight_dim": (c_i64, [c_vp, c_i32, c_i32]),
    "dhg_load_weight": (c_i32, [c_vp, c_cp, c_vp, ctypes.POINTER(c_i64), c_i32]),
    "dhg_set_schedule": (c_i32, [c_vp, c_vp, c_vp]),
    "dhg_finalize": (c_i32, [c_vp]),
    "dhg_plan": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32]),
    "dhg_denoise": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "dhg_sample": (c_i32, [c_vp, c_i32, c_vp, c_vp, c_u64, c_vp, c_vp, c_i32, c_vp, c_vp]),
    "dhg_sample_host": (c_i32, [c_vp, c_i32, c_vp, c_vp, c_u64, c_vp, c_vp, c_i32, c_vp]),
    "dhg_check_errors": (c_i32, [c_vp, c_vp]),
    "dhg_posterior_step": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "dhg_last_launch_count": (c_i64, [c_vp]),
    "dhg_plan_bytes": (c_i64, [c_vp]),
    "dhg_set_option": (c_i32, [c_vp, c_cp, c_i32]),
    "dhg_debug_read": (c_i64, [c_vp, c_cp, c_vp, c_i64]),
    "dhg_debug_tc_gemm": (c_i32, [c_i32, c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "dhg_debug_time_text": (c_i32, [c_vp, c_i32, c_i32, ctypes.POINTER(ctypes.c_float)]),
    "dhg_debug_attention": (c_i32, [c_i32, ctypes.POINTER(DebugAttn), c_i32, c_i32, ctypes.POINTER(ctypes.c_float), c_vp]),
    "dhg_debug_tc_gemm_ex": (c_i32, [c_i32, c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_i32, ctypes.POINTER(DebugEpilogue),
                                     c_i32, ctypes.POINTER(ctypes.c_float), c_vp]),
    "dhg_style_last_error": (c_cp, []),
    "dhg_style_create": (c_i32, [c_i32, ctypes.POINTER(c_vp)]),
    "dhg_style_destroy": (c_i32, [c_vp]),
    "dhg_style_load_weight": (c_i32, [c_vp, c_cp, c_vp, ctypes.POINTER(c_i64), c_i32]),
    "dhg_style_finalize": (c_i32, [c_vp]),
    "dhg_style_extract": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "dhg_train_last_error": (c_cp, []),
    "dhg_train_scratch_doubles": (c_i32, []),
    "dhg_train_perturb": (c_i32, [c_i32, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp]),
    "dhg_train_loss": (c_i32, [c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "dhg_train_sqnorm": (c_i32, [c_i32, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "dhg_train_adam_step": (c_i32, [c_i32, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                    ctypes.c_double, ctypes.c_double, c_vp, ctypes.c_double, c_i32, c_vp]),
    # forward + backward pass of the denoiser (csrc/train_step.cu)
    "dhg_trainer_last_error": (c_cp, []),
    "dhg_trainer_param_count": (c_i64, [c_i32, c_i32]),
    "dhg_trainer_param_info": (c_i32, [c_i32, c_i32, c_i32, c_vp, c_i32, c_vp, c_vp]),
    "dhg_trainer_create": (c_i32, [c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "dhg_trainer_destroy": (c_i32, [c_vp]),
    "dhg_trainer_workspace_bytes": (c_i64, [c_vp]),
    "dhg_trainer_last_launches": (c_i64, [c_vp]),
    "dhg_trainer_set_option": (c_i32, [c_cp, c_i32]),
    "dhg_trainer_forward": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "dhg_trainer_backward": (c_i32, [c_vp, c_vp, c_vp, c_vp]),
}


class DhgError(RuntimeError):
    pass


def lib():
    """Load (once) and return the C-ABI library.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DhgError(
                f"{LIB_PATH} is not built. Run `python __graft_entry__.py build` "
                "(needs nvcc). There is no CPU or PyTorch fallback for this path."
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype, fn.argtypes = res, args
        _lib = handle
        # experiment switches for the kernels, e.g. DHG_OPTS="tap_shift=0,w_resident=0"
        for kv in filter(None, os.environ.get("DHG_OPTS", "").split(",")):
            k, v = kv.split("=")
            if handle.dhg_set_option(None, k.strip().encode(), int(v)) != 0:
                raise DhgError(handle.dhg_last_error().decode())
    return _lib


def check(rc):
    if rc != 0:
        raise DhgError(lib().dhg_last_error().decode("utf-8", "replace"))
