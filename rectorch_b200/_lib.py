"""ctypes binding of libb200vae.so (include/b200vae.h).

The shared library is the only compute path of this package: if it is missing or
cannot be loaded every compute call raises -- there is no CPU / PyTorch fallback.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_uint64,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200vae.so")
MAX_LAYERS = 8

EINVAL, ECUDA, ESTATE, ECAPACITY = -1, -2, -3, -4


class B200VaeError(RuntimeError):
    """Raised for any non-zero return code of the C ABI."""

    def __init__(self, code, msg):
        super().__init__("b200vae error %d: %s" % (code, msg))
        self.code = code


class Config(Structure):
    _fields_ = [("device", c_int32), ("is_vae", c_int32), ("n_enc", c_int32), ("n_dec", c_int32),
                ("enc_dims", c_int32 * (MAX_LAYERS + 1)), ("dec_dims", c_int32 * (MAX_LAYERS + 1)),
                ("max_batch", c_int32), ("max_batch_nnz", c_int64), ("use_tensor_cores", c_int32),
                ("cond_dim", c_int32)]


_lib = None

_SIGS = {
    "b200vae_last_error": (c_char_p, []),
    "b200vae_version": (c_int, []),
    "b200vae_ctx_create": (c_int, [POINTER(c_void_p), POINTER(Config)]),
    "b200vae_ctx_destroy": (c_int, [c_void_p]),
    "b200vae_bind_params": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                    POINTER(c_int64), POINTER(c_int64)]),
    "b200vae_sync_weights": (c_int, [c_void_p, c_void_p]),
    "b200vae_bind_shadow": (c_int, [c_void_p, c_void_p, c_int64]),
    "b200vae_defer_wait_event": (c_int, [c_void_p, c_void_p]),
    "b200vae_dp_pack": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "b200vae_dp_unpack": (c_int, [c_void_p, c_int32, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                  c_void_p, c_int64, c_void_p]),
    "b200vae_set_w1_sharding": (c_int, [c_void_p, c_void_p, c_int32, c_int32]),
    "b200vae_w1_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "b200vae_check_error_flag": (c_int, [c_void_p]),
    "b200vae_wait_wd_ready": (c_int, [c_void_p, c_void_p]),
    "b200vae_bind_csr": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64]),
    "b200vae_dense_to_csr": (c_int, [c_void_p, c_int, c_void_p, c_int32, c_void_p]),
    "b200vae_expand_batch": (c_int, [c_void_p, c_int, c_void_p, c_int32, c_void_p, c_void_p]),
    "b200vae_forward_backward": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int, c_float, c_float,
                                         c_float, c_uint64, c_uint64, c_int64, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p]),
    "b200vae_enc0_grad": (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_float, c_uint64, c_uint64, c_int64,
                                  c_void_p]),
    "b200vae_adam_step": (c_int, [c_void_p, c_float, c_float, c_float, c_float, c_float, c_float,
                                  c_int64, c_void_p]),
    "b200vae_adam_step_range": (c_int, [c_void_p, c_float, c_float, c_float, c_float, c_float, c_float,
                                        c_int64, c_int64, c_int64, c_int, c_void_p]),
    "b200vae_adam_step_split": (c_int, [c_void_p, c_float, c_float, c_float, c_float, c_float, c_float,
                                        c_int64, c_void_p, c_int32, c_int, c_void_p]),
    "b200vae_train_step": (c_int, [c_void_p, c_void_p, c_int32, c_int, c_float, c_float, c_float,
                                   c_uint64, c_int64, c_void_p, c_void_p, c_float, c_float, c_float,
                                   c_float, c_float, c_void_p, c_void_p]),
    "b200vae_train_step_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_float,
                                        c_float, c_float, c_uint64, c_int64, c_float, c_float,
                                        c_void_p, c_void_p]),
    "b200vae_predict": (c_int, [c_void_p, c_void_p, c_int32, c_int, c_int, c_float, c_uint64,
                                c_uint64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200vae_decode": (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "b200vae_topk_metrics": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, POINTER(c_int32),
                                     POINTER(c_int32), c_int32, c_void_p, c_void_p, c_void_p]),
    "b200vae_topk_metrics_csr": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p,
                                         c_void_p]),
    "b200vae_expand_rows_raw": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p,
                                        c_void_p]),
    "b200vae_dense_to_csr_raw": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_int64, c_void_p]),
    "b200vae_multinomial_nll_rows": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "b200vae_kl_rows": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "b200vae_gemm_f16": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int,
                                  c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p]),
    "b200vae_dec_fwd_lse": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                    c_void_p, c_void_p]),
    "b200vae_probe_launch": (c_int, [c_int, c_int, c_int, c_int, c_void_p]),
    "b200vae_launch_count": (c_int64, [c_void_p, c_int]),
    "b200vae_set_timing": (c_int, [c_void_p, c_int]),
    "b200vae_set_deterministic": (c_int, [c_void_p, c_int]),
    "b200vae_kernel_ms": (c_float, [c_void_p, c_int]),
    "b200vae_timing_report": (c_int, [c_void_p, c_char_p, c_int]),
    "b200vae_build_cond_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "b200vae_ease_gram": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "b200vae_ease_solve": (c_int, [c_void_p, c_int32, ctypes.c_double, c_void_p, c_void_p]),
    "b200vae_ease_scores": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "b200vae_csv_open": (c_int, [POINTER(c_void_p), c_char_p, ctypes.c_char, c_int]),
    "b200vae_csv_info": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int32), POINTER(c_int64),
                                 POINTER(c_int64), POINTER(c_int64)]),
    "b200vae_csv_value_column": (c_char_p, [c_void_p]),
    "b200vae_csv_to_csr": (c_int, [c_void_p, c_int64, c_int64, c_int32, c_int, c_void_p, c_void_p, c_void_p,
                                   POINTER(c_int64)]),
    "b200vae_csv_close": (c_int, [c_void_p]),
}

EXPORTS = tuple(sorted(_SIGS))


def lib():
    """Load (once) and return the ctypes handle; raises if the engine is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200VaeError(ECUDA, "libb200vae.so is not built (run `python -m rectorch_b200.build`); "
                                      "rectorch_b200 has no CPU fallback")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().b200vae_last_error()
        raise B200VaeError(rc, msg.decode("utf-8", "replace") if msg else "unknown error")
    return rc


def ptr(t):
    """Device/host pointer of a torch tensor (or None) as c_void_p."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """The caller's current torch stream on `device` (default: the current device)."""
    import torch
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)
