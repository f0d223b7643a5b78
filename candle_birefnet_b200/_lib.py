"""ctypes binding of libbirefnet_b200.so (the C ABI in include/birefnet_b200.h).

The library is built in-tree by `candle_birefnet_b200.build`; a missing library is an ImportError-grade failure,
never a silent fallback -- there is no CPU implementation of the hot path in this package.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

HERE = Path(__file__).resolve().parent
import os
LIB_PATH = Path(os.environ["BRN_LIB_PATH"]) if os.environ.get("BRN_LIB_PATH") else HERE / "libbirefnet_b200.so"   # A/B builds
HEADER = HERE.parent / "include" / "birefnet_b200.h"

OK = 0
PREC_FP32, PREC_BF16, PREC_FP16 = 0, 1, 2
DEFORM_CPU_FALLBACK, DEFORM_DEFORMABLE = 0, 1
F32, BF16, F16 = 0, 1, 2


class BrnConfig(C.Structure):
    _fields_ = [("embed_dim", C.c_int32), ("depths", C.c_int32 * 4), ("num_heads", C.c_int32 * 4),
                ("window_size", C.c_int32), ("mlp_ratio", C.c_int32), ("patch_size", C.c_int32),
                ("precision", C.c_int32), ("deform_mode", C.c_int32), ("micro_batch", C.c_int32)]


class BrnError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"[brn_status {status}] {msg}")
        self.status = status


def declared_symbols() -> list[str]:
    """Every function the public header declares (used by the symbol-export test)."""
    return re.findall(r"BRN_API\s+[\w\s\*]+?\b(brn_\w+)\s*\(", HEADER.read_text())


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: run `python -m candle_birefnet_b200.build` (nvcc, sm_100a). "
                          "There is no fallback implementation.")
    L = C.CDLL(str(LIB_PATH))
    fp, vp, i32, i64 = C.POINTER(C.c_float), C.c_void_p, C.c_int32, C.c_int64
    L.brn_last_error.restype = C.c_char_p
    L.brn_version.restype = C.c_char_p
    L.brn_config_swin_l.argtypes = [C.POINTER(BrnConfig)]
    L.brn_config_swin_l.restype = None
    L.brn_config_swin_b.argtypes = [C.POINTER(BrnConfig)]
    L.brn_config_swin_b.restype = None
    for name in ("brn_config_swin_t", "brn_config_swin_s"):
        getattr(L, name).argtypes = [C.POINTER(BrnConfig)]
        getattr(L, name).restype = None
    L.brn_model_create.argtypes = [C.POINTER(BrnConfig), C.c_int, C.POINTER(vp)]
    L.brn_model_destroy.argtypes = [vp]
    L.brn_model_destroy.restype = None
    L.brn_model_set_tensor.argtypes = [vp, C.c_char_p, vp, C.c_int, C.POINTER(i64), C.c_int]
    L.brn_model_num_tensors.argtypes = [vp]
    L.brn_model_tensor_info.argtypes = [vp, i32, C.POINTER(C.c_char_p), C.POINTER(i64), C.POINTER(i32)]
    L.brn_model_load_safetensors.argtypes = [vp, C.c_char_p, C.POINTER(i32)]
    L.brn_model_finalize.argtypes = [vp]
    L.brn_model_set_precision.argtypes = [vp, C.c_int]
    L.brn_model_set_deform_mode.argtypes = [vp, C.c_int]
    L.brn_model_set_cuda_graph.argtypes = [vp, C.c_int]
    for name in ("brn_forward_logits", "brn_forward"):
        getattr(L, name).argtypes = [vp, vp, i32, i32, i32, C.c_int, vp, C.c_int, vp]
    L.brn_backbone_forward.argtypes = [vp, vp, i32, i32, i32, C.c_int, C.POINTER(vp), C.c_int, vp]
    L.brn_features_forward.argtypes = [vp, vp, i32, i32, i32, C.c_int, C.POINTER(vp), C.c_int, vp]
    L.brn_decoder_forward.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, C.c_int, vp, vp]
    L.brn_window_attention.argtypes = [C.c_int, C.c_int, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    L.brn_deform_conv2d.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]
    L.brn_deformable_conv2d.argtypes = [C.c_int, C.c_int, C.c_int, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32,
                                        i32, vp]
    L.brn_linear.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    L.brn_preprocess_rgb8.argtypes = [C.c_int, vp, i32, i32, i32, i32, i32, vp]
    L.brn_postprocess_mask.argtypes = [C.c_int, vp, i32, i32, i32, i32, i32, vp]
    L.brn_infer_rgb8.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    L.brn_ln_linear.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    L.brn_swin_mlp.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp]
    L.brn_conv2d.argtypes = [C.c_int, C.c_int, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    L.brn_bench_op.argtypes = [C.c_int, C.c_int, C.c_int] + [i32] * 10 + [fp]
    L.brn_sharded_create.argtypes = [C.POINTER(BrnConfig), C.POINTER(i32), i32, C.POINTER(vp)]
    L.brn_sharded_destroy.argtypes = [vp]
    L.brn_sharded_destroy.restype = None
    L.brn_sharded_num_devices.argtypes = [vp]
    L.brn_sharded_num_devices.restype = i32
    L.brn_sharded_set_tensor.argtypes = [vp, C.c_char_p, vp, C.c_int, C.POINTER(i64), C.c_int]
    L.brn_sharded_load_safetensors.argtypes = [vp, C.c_char_p, C.POINTER(i32)]
    L.brn_sharded_finalize.argtypes = [vp]
    for name in ("brn_sharded_forward_logits", "brn_sharded_forward"):
        getattr(L, name).argtypes = [vp, vp, i32, i32, i32, vp]
    L.brn_host_alloc.argtypes = [C.c_size_t]
    L.brn_host_alloc.restype = vp
    L.brn_host_free.argtypes = [vp]
    L.brn_host_free.restype = None
    L.brn_launch_count.argtypes = [vp]
    L.brn_launch_count.restype = i64
    L.brn_launch_count_reset.argtypes = [vp]
    L.brn_launch_count_reset.restype = None
    L.brn_profile_enable.argtypes = [vp, C.c_int]
    L.brn_profile_enable.restype = None
    L.brn_profile_get.argtypes = [vp, C.POINTER(C.POINTER(C.c_char_p)), C.POINTER(fp), C.POINTER(C.POINTER(C.c_double))]
    L.brn_profile_get.restype = i32
    L.brn_kernel_class_times.argtypes = [vp, fp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i32), i32]
    L.brn_kernel_class_times.restype = i32
    _lib = L
    return L


def check(status: int) -> None:
    if status != OK:
        raise BrnError(status, lib().brn_last_error().decode())
