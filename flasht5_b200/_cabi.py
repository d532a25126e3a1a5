"""ctypes binding of libb200t5.so (the C ABI declared in include/b200t5.h).

This is the only place the shared library is loaded.  There is no fallback of any kind: if the
library has not been built, or the device is not an sm_100 part, the ops raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200T5_LIB") or os.path.join(_HERE, "libb200t5.so")   # env override: developer builds only

F16, BF16, F32 = 0, 1, 2
_DTYPE_CODE = {torch.float16: F16, torch.bfloat16: BF16, torch.float32: F32}

I64x4 = C.c_int64 * 4


ATTN_DETERMINISTIC = 1      # b200t5_attn_params.flags (include/b200t5.h)
ATTN_DBIAS_F32 = 2
ATTN_DBIAS_ACCUMULATE = 4


class AttnParams(C.Structure):
    """Mirror of `b200t5_attn_params` (include/b200t5.h)."""
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("D", C.c_int32),
        ("dtype", C.c_int32), ("causal", C.c_int32), ("bias_B", C.c_int32), ("bias_H", C.c_int32),
        ("sm_scale", C.c_float), ("device", C.c_int32), ("flags", C.c_int32),
        ("stream", C.c_void_p),
        ("q", C.c_void_p), ("q_strides", I64x4),
        ("k", C.c_void_p), ("k_strides", I64x4),
        ("v", C.c_void_p), ("v_strides", I64x4),
        ("bias", C.c_void_p), ("bias_strides", I64x4),
        ("o", C.c_void_p), ("o_strides", I64x4),
        ("lse", C.c_void_p),
        ("dout", C.c_void_p), ("do_strides", I64x4),
        ("dq", C.c_void_p), ("dq_strides", I64x4),
        ("dk", C.c_void_p), ("dk_strides", I64x4),
        ("dv", C.c_void_p), ("dv_strides", I64x4),
        ("dbias", C.c_void_p), ("dbias_strides", I64x4),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class RpeParams(C.Structure):
    """Mirror of `b200t5_rpe_params` (include/b200t5.h)."""
    _fields_ = [
        ("table", C.c_void_p), ("table_stride_b", C.c_int64), ("table_stride_h", C.c_int64),
        ("table_dtype", C.c_int32), ("num_buckets", C.c_int32),
        ("lut", C.c_void_p), ("lut_zero", C.c_int32), ("lut_len", C.c_int32),
        ("const_lo", C.c_int32), ("const_hi", C.c_int32),
        ("band", C.c_void_p), ("dtable", C.c_void_p),
    ]


class AdamwTensor(C.Structure):
    """Mirror of `b200t5_adamw_tensor` (include/b200t5.h)."""
    _fields_ = [
        ("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("comp", C.c_void_p),
        ("numel", C.c_int64), ("first_chunk", C.c_int32), ("sqrt_numel", C.c_float), ("ss_base", C.c_float),
        ("ss_floor", C.c_float), ("neg_lr_wd", C.c_float), ("reserved", C.c_int32),
    ]


# every symbol include/b200t5.h declares: name -> (restype, argtypes)
_i64, _i32, _f, _vp, _sz = C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_size_t
SYMBOLS = {
    "b200t5_attn_fwd": (_i32, [C.POINTER(AttnParams)]),
    "b200t5_attn_fwd_workspace_bytes": (_sz, [C.POINTER(AttnParams)]),
    "b200t5_attn_bwd_workspace_bytes": (_sz, [C.POINTER(AttnParams)]),
    "b200t5_attn_bwd": (_i32, [C.POINTER(AttnParams)]),
    "b200t5_rmsnorm_fwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _f, _i32, _i32, _i32, _vp]),
    "b200t5_rmsnorm_bwd_workspace_bytes": (_sz, [_i64]),
    "b200t5_rmsnorm_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i64, _i64, _i64, _i64, _i64, _i32, _i32,
                                  _i32, _vp]),
    "b200t5_ce_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _i64, _i64, _f, _f, _f, _i64, _i32, _i32, _vp]),
    "b200t5_ce_bwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _f, _f, _f, _i64, _i32, _i32,
                             _vp]),
    "b200t5_t5_bias_fwd": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200t5_t5_bias_bwd": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200t5_rpe_band_len": (_i32, [_i32, _i32]),
    "b200t5_rpe_band": (_i32, [C.POINTER(RpeParams), _i32, _i32, _i32, _vp]),
    "b200t5_attn_rpe_fwd": (_i32, [C.POINTER(AttnParams), C.POINTER(RpeParams)]),
    "b200t5_attn_rpe_bwd_workspace_bytes": (_sz, [C.POINTER(AttnParams), C.POINTER(RpeParams)]),
    "b200t5_attn_rpe_bwd": (_i32, [C.POINTER(AttnParams), C.POINTER(RpeParams)]),
    "b200t5_adamw_chunk_elems": (_i32, []),
    "b200t5_adamw_workspace_bytes": (_sz, [_i32, _i32]),
    "b200t5_adamw_scale_step": (_i32, [_vp, _i32, _vp, _i32, _vp, _sz, _i32, _i32, _i32, _f, _f, _f, _i32, _i32, _vp]),
    "b200t5_abi_version": (_i32, []),
    "b200t5_last_error": (C.c_char_p, []),
    "b200t5_launch_count": (C.c_uint64, []),
    "b200t5_device_supported": (_i32, [_i32]),
    "b200t5_profile_enable": (_i32, [_i32]),
    "b200t5_profile_collect": (_i32, [C.POINTER(C.c_int), C.POINTER(C.c_float), _i32]),
}
ABI_VERSION = 3

_lib = None
_lock = threading.Lock()


def load():
    """Load (once) and return the shared library; raises if it is missing or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m flasht5_b200.build` "
                "(nvcc, sm_100a).  flasht5_b200 has no fallback path.")
        # torch has already loaded libcudart; RTLD_GLOBAL is not needed (cudart is linked statically).
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.b200t5_abi_version() != ABI_VERSION:
            raise ImportError(f"libb200t5.so ABI {lib.b200t5_abi_version()} != binding ABI {ABI_VERSION}")
        _lib = lib
    return _lib


def last_error() -> str:
    return load().b200t5_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def launch_count() -> int:
    return int(load().b200t5_launch_count())


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPE_CODE[dt]
    except KeyError:
        raise TypeError(f"unsupported dtype {dt}") from None


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("flasht5_b200 ops run on sm_100 CUDA devices only (got a %s tensor); "
                               "there is no CPU fallback" % t.device.type)


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def strides4(t: torch.Tensor):
    return I64x4(*t.stride())


def profile_enable(on: bool):
    load().b200t5_profile_enable(1 if on else 0)


def profile_collect(cap: int = 4096):
    """[(kernel_id, ms), ...] for the main attention kernels launched since the last collect
    (synchronise the stream first)."""
    ids = (C.c_int * cap)()
    ms = (C.c_float * cap)()
    n = load().b200t5_profile_collect(ids, ms, cap)
    return [(int(ids[i]), float(ms[i])) for i in range(n)]
