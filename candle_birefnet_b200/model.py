"""Host-side mirror of the reference crate's public API for the hot path.

    reference (Rust, src/lib.rs:6-14)                     here
    ---------------------------------------------------   -----------------------------------------------
    BiRefNetConfig::swin_l()          birefnet.rs:64-66   BiRefNetConfig.swin_l()
    BiRefNet::new(config, vb)         birefnet.rs:389     BiRefNet.new(config, vb)   (vb: key -> ndarray, or a
                                                          .safetensors path -- VarBuilder::from_tensors,
                                                          examples/infer_image.rs:35-40)
    model.forward_logits(&x)          birefnet.rs:412     model.forward_logits(x)    x: float32 [B,3,H,W]
    model.forward(&x) / Module        birefnet.rs:466     model.forward(x) / model(x)
    model.backbone.forward(&x)        swin.rs:768         model.backbone_forward(x) -> 4 NCHW maps
    DeformableConv2d::new / forward   deform_conv.rs:29   ops.DeformableConv2d(in, out, k, stride, padding, vb)(x)

Inputs/outputs are host numpy arrays (copied inside the C call) or torch CUDA tensors (zero-copy device pointers).
All arithmetic happens in libbirefnet_b200.so; this file only marshals pointers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Mapping, Tuple, Union

import numpy as np

from . import _lib
from ._lib import BrnConfig, BrnError, check, lib


@dataclass
class SwinConfig:
    """SwinConfig (src/swin.rs:13-23); head_dim 32, window 12 (swin_b / swin_l) or 7 (swin_t / swin_s)."""
    embed_dim: int = 192
    depths: Tuple[int, int, int, int] = (2, 2, 18, 2)
    num_heads: Tuple[int, int, int, int] = (6, 12, 24, 48)
    window_size: int = 12
    mlp_ratio: int = 4
    patch_size: int = 4

    @staticmethod
    def swin_l() -> "SwinConfig":
        return SwinConfig()

    @staticmethod
    def swin_b() -> "SwinConfig":
        """SwinConfig::swin_b (src/swin.rs:54-66)."""
        return SwinConfig(embed_dim=128, num_heads=(4, 8, 16, 32))

    @staticmethod
    def swin_t() -> "SwinConfig":
        """SwinConfig::swin_t (src/swin.rs:27-38): window 7."""
        return SwinConfig(embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), window_size=7)

    @staticmethod
    def swin_s() -> "SwinConfig":
        """SwinConfig::swin_s (src/swin.rs:41-52): window 7."""
        return SwinConfig(embed_dim=96, depths=(2, 2, 18, 2), num_heads=(3, 6, 12, 24), window_size=7)


@dataclass
class BiRefNetConfig:
    """BiRefNetConfig (src/birefnet.rs:13-30) plus the two runtime knobs of this implementation."""
    size: Tuple[int, int] = (1024, 1024)
    backbone: str = "swin_v1_l"
    swin: SwinConfig = field(default_factory=SwinConfig.swin_l)
    mul_scl_ipt: bool = True
    ms_supervision: bool = True
    dec_ipt: bool = True
    use_aspp_deformable: bool = True
    precision: str = "fp16"            # tcgen05 path with "fp16" | "bf16" operands, or "fp32" (SIMT FMA path)
    deform_mode: str = "deformable"    # "deformable" (Metal path semantics) | "cpu_fallback" (candle CPU semantics)
    micro_batch: int = 0

    @staticmethod
    def swin_l() -> "BiRefNetConfig":
        return BiRefNetConfig()


_PREC = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "fp16": _lib.PREC_FP16}
_DEF = {"cpu_fallback": _lib.DEFORM_CPU_FALLBACK, "deformable": _lib.DEFORM_DEFORMABLE}


def _is_torch_cuda(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


class BiRefNet:
    def __init__(self, handle: C.c_void_p, config: BiRefNetConfig, device: int):
        self._h = handle
        self.config = config
        self.device = device

    # ---- construction -------------------------------------------------------------------------------------
    @staticmethod
    def new(config: BiRefNetConfig, vb: Union[Mapping[str, np.ndarray], str], device: int = 0) -> "BiRefNet":
        L = lib()
        c = BrnConfig()
        c.embed_dim = config.swin.embed_dim
        for i in range(4):
            c.depths[i] = config.swin.depths[i]
            c.num_heads[i] = config.swin.num_heads[i]
        c.window_size, c.mlp_ratio, c.patch_size = config.swin.window_size, config.swin.mlp_ratio, config.swin.patch_size
        c.precision, c.deform_mode, c.micro_batch = _PREC[config.precision], _DEF[config.deform_mode], config.micro_batch
        h = C.c_void_p()
        check(L.brn_model_create(C.byref(c), device, C.byref(h)))
        m = BiRefNet(h, config, device)
        if isinstance(vb, str):
            # the library's own safetensors reader (brn_model_load_safetensors): no Python-side parsing
            try:
                n = C.c_int32(0)
                check(L.brn_model_load_safetensors(h, vb.encode(), C.byref(n)))
                check(L.brn_model_finalize(h))
            except Exception:
                L.brn_model_destroy(h)
                m._h = None
                raise
            return m
        try:
            for key in m.tensor_keys():
                if key not in vb:
                    raise BrnError(3, f"cannot find tensor {key}")   # candle: `vb.get` fails inside BiRefNet::new
                a = np.ascontiguousarray(vb[key])
                if a.dtype == np.float32:
                    dt = _lib.F32
                elif a.dtype == np.float16:
                    dt = _lib.F16
                else:
                    a = a.astype(np.float32)
                    dt = _lib.F32
                shape = (C.c_int64 * a.ndim)(*a.shape)
                check(L.brn_model_set_tensor(h, key.encode(), a.ctypes.data_as(C.c_void_p), dt, shape, a.ndim))
            check(L.brn_model_finalize(h))
        except Exception:
            L.brn_model_destroy(h)
            m._h = None
            raise
        return m

    @staticmethod
    def new_synthetic(config: BiRefNetConfig, seed: int = 0, weight_set: str = "B", offset_sigma: float = 2.0,
                      device: int = 0) -> "BiRefNet":
        """BiRefNet::new on seeded random-init weights (no network for the HF checkpoint): the schema comes from the
        handle itself, the values from `synth.synthetic_weights`."""
        from .synth import synthetic_weights
        L = lib()
        probe = BiRefNet._create(config, device)
        try:
            sc = probe.tensor_schema()
        finally:
            probe.close()
        return BiRefNet.new(config, synthetic_weights(sc, seed, weight_set, offset_sigma), device)

    @staticmethod
    def _create(config: BiRefNetConfig, device: int) -> "BiRefNet":
        L = lib()
        c = BrnConfig()
        c.embed_dim = config.swin.embed_dim
        for i in range(4):
            c.depths[i] = config.swin.depths[i]
            c.num_heads[i] = config.swin.num_heads[i]
        c.window_size, c.mlp_ratio, c.patch_size = config.swin.window_size, config.swin.mlp_ratio, config.swin.patch_size
        c.precision, c.deform_mode, c.micro_batch = _PREC[config.precision], _DEF[config.deform_mode], config.micro_batch
        h = C.c_void_p()
        check(L.brn_model_create(C.byref(c), device, C.byref(h)))
        return BiRefNet(h, config, device)

    def tensor_keys(self) -> List[str]:
        L = lib()
        out = []
        key = C.c_char_p()
        shape = (C.c_int64 * 4)()
        rank = C.c_int32()
        for i in range(L.brn_model_num_tensors(self._h)):
            check(L.brn_model_tensor_info(self._h, i, C.byref(key), shape, C.byref(rank)))
            out.append(key.value.decode())
        return out

    def tensor_schema(self) -> Dict[str, Tuple[int, ...]]:
        L = lib()
        out = {}
        key = C.c_char_p()
        shape = (C.c_int64 * 4)()
        rank = C.c_int32()
        for i in range(L.brn_model_num_tensors(self._h)):
            check(L.brn_model_tensor_info(self._h, i, C.byref(key), shape, C.byref(rank)))
            out[key.value.decode()] = tuple(shape[d] for d in range(rank.value))
        return out

    def set_precision(self, precision: str) -> None:
        check(lib().brn_model_set_precision(self._h, _PREC[precision]))
        self.config.precision = precision

    def set_deform_mode(self, mode: str) -> None:
        check(lib().brn_model_set_deform_mode(self._h, _DEF[mode]))
        self.config.deform_mode = mode

    def set_cuda_graph(self, on: bool) -> None:
        """Replay the forward as a CUDA graph from the second call with the same buffers/shape (default on)."""
        check(lib().brn_model_set_cuda_graph(self._h, 1 if on else 0))

    def close(self) -> None:
        if self._h:
            lib().brn_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- forward -------------------------------------------------------------------------------------------
    def _run(self, fn, x, out=None, stream=None):
        if _is_torch_cuda(x):
            import torch
            assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4 and x.shape[1] == 3
            B, _, H, W = x.shape
            if x.device.index != self.device:
                raise BrnError(1, f"input lives on cuda:{x.device.index}, the handle on cuda:{self.device}")
            if out is None:
                out = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
            elif not (_is_torch_cuda(out) and out.device == x.device and out.dtype == torch.float32 and
                      out.is_contiguous() and tuple(out.shape) == (B, 1, H, W)):
                raise BrnError(5, f"`out` must be a contiguous float32 CUDA tensor [{B},1,{H},{W}] on {x.device}")
            s = stream if stream is not None else torch.cuda.current_stream(x.device).cuda_stream
            if not s:
                s = 1   # cudaStreamLegacy: torch's default stream, explicitly (NULL would mean the handle's own stream)
            check(fn(self._h, C.c_void_p(x.data_ptr()), B, H, W, 1, C.c_void_p(out.data_ptr()), 1, C.c_void_p(s)))
            return out
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 4 or x.shape[1] != 3:
            raise BrnError(5, f"expected [B,3,H,W], got {x.shape}")
        B, _, H, W = x.shape
        if out is None:
            out = np.empty((B, 1, H, W), dtype=np.float32)
        elif not (isinstance(out, np.ndarray) and out.dtype == np.float32 and out.flags.c_contiguous and
                  out.shape == (B, 1, H, W)):
            raise BrnError(5, f"`out` must be a C-contiguous float32 array [{B},1,{H},{W}]")
        check(fn(self._h, x.ctypes.data_as(C.c_void_p), B, H, W, 0, out.ctypes.data_as(C.c_void_p), 0, None))
        return out

    def forward_logits(self, x, out=None, stream=None):
        """BiRefNet::forward_logits (src/birefnet.rs:412-461)."""
        return self._run(lib().brn_forward_logits, x, out, stream)

    def forward(self, x, out=None, stream=None):
        """BiRefNet::forward (src/birefnet.rs:466-469): sigmoid(logits)."""
        return self._run(lib().brn_forward, x, out, stream)

    __call__ = forward

    def _feature_call(self, fn, x, chans, stream=None):
        """Shared marshalling of brn_backbone_forward / brn_features_forward: host numpy or torch CUDA tensors."""
        if _is_torch_cuda(x):
            import torch
            assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4 and x.shape[1] == 3
            if x.device.index != self.device:
                raise BrnError(1, f"input lives on cuda:{x.device.index}, the handle on cuda:{self.device}")
            B, _, H, W = x.shape
            outs = [torch.empty((B, chans[i], H // (4 << i), W // (4 << i)), dtype=torch.float32, device=x.device)
                    for i in range(4)]
            ptrs = (C.c_void_p * 4)(*[C.c_void_p(o.data_ptr()) for o in outs])
            s = stream if stream is not None else torch.cuda.current_stream(x.device).cuda_stream
            check(fn(self._h, C.c_void_p(x.data_ptr()), B, H, W, 1, ptrs, 1, C.c_void_p(s or 1)))
            return outs
        x = np.ascontiguousarray(x, dtype=np.float32)
        B, _, H, W = x.shape
        outs = [np.empty((B, chans[i], H // (4 << i), W // (4 << i)), dtype=np.float32) for i in range(4)]
        ptrs = (C.c_void_p * 4)(*[o.ctypes.data_as(C.c_void_p) for o in outs])
        check(fn(self._h, x.ctypes.data_as(C.c_void_p), B, H, W, 0, ptrs, 0, None))
        return outs

    def backbone_forward(self, x, stream=None):
        """SwinTransformer::forward (src/swin.rs:768-797) -> [x1,x2,x3,x4] NCHW float32."""
        E = self.config.swin.embed_dim
        return self._feature_call(lib().brn_backbone_forward, x, [E << i for i in range(4)], stream)

    def features_forward(self, x, stream=None):
        """First half of forward_logits (src/birefnet.rs:412-454) -> [x1,x2,x3,x4_cxt] NCHW float32: the multi-scale
        features the squeeze module and decoder consume (what examples/bench_inference.rs times as `backbone`)."""
        E = self.config.swin.embed_dim
        return self._feature_call(lib().brn_features_forward, x, [2 * (E << i) for i in range(3)] + [30 * E], stream)

    def decoder_forward(self, x, x1, x2, x3, x4, out=None, stream=None):
        """SqueezeModule + BiRefNetDecoder::forward (src/birefnet.rs:86-94, 278-376) on given features."""
        if _is_torch_cuda(x):
            import torch
            ts = (x, x1, x2, x3, x4)
            assert all(_is_torch_cuda(t) and t.dtype == torch.float32 and t.is_contiguous() for t in ts)
            B, _, H, W = x.shape
            if out is None:
                out = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
            s = stream if stream is not None else torch.cuda.current_stream(x.device).cuda_stream
            check(lib().brn_decoder_forward(self._h, *[C.c_void_p(t.data_ptr()) for t in ts], B, H, W, 1,
                                            C.c_void_p(out.data_ptr()), C.c_void_p(s or 1)))
            return out
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (x, x1, x2, x3, x4)]
        B, _, H, W = arrs[0].shape
        out = np.empty((B, 1, H, W), dtype=np.float32)
        check(lib().brn_decoder_forward(self._h, *[a.ctypes.data_as(C.c_void_p) for a in arrs], B, H, W, 0,
                                        out.ctypes.data_as(C.c_void_p), None))
        return out

    def infer_rgb8(self, rgb: np.ndarray, size: Tuple[int, int] = (1024, 1024)) -> np.ndarray:
        """examples/infer_image.rs:44-105 on the device: uint8 RGB [B,h,w,3] -> uint8 masks [B,h,w] (resize to `size`
        with Triangle, ImageNet normalise, forward_logits, sigmoid, u8, Lanczos3 resize back)."""
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        if rgb.ndim == 3:
            rgb = rgb[None]
        B, h, w, c = rgb.shape
        if c != 3:
            raise BrnError(5, f"expected [B,h,w,3] uint8, got {rgb.shape}")
        out = np.empty((B, h, w), dtype=np.uint8)
        check(lib().brn_infer_rgb8(self._h, rgb.ctypes.data_as(C.c_void_p), B, h, w, size[0], size[1],
                                   out.ctypes.data_as(C.c_void_p)))
        return out

    # ---- introspection -------------------------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(lib().brn_launch_count(self._h))

    def reset_launch_count(self) -> None:
        lib().brn_launch_count_reset(self._h)

    KERNEL_CLASSES = ("gemm_tcgen05", "attn_tcgen05", "deform_tcgen05", "gemm_simt", "attn_simt", "layernorm", "glue")

    def profile(self, on=True) -> None:
        """on = 1/True: per-stage CUDA events; on = 2: also events around every kernel launch (per-class sums)."""
        lib().brn_profile_enable(self._h, int(on))

    def kernel_class_times(self):
        """{class: dict(ms, flops, bytes, launches)} of the last forward run with profile(2)."""
        n = len(self.KERNEL_CLASSES)
        ms = (C.c_float * n)()
        fl = (C.c_double * n)()
        by = (C.c_double * n)()
        cnt = (C.c_int32 * n)()
        lib().brn_kernel_class_times(self._h, ms, fl, by, cnt, n)
        return {k: dict(ms=float(ms[i]), flops=float(fl[i]), bytes=float(by[i]), launches=int(cnt[i]))
                for i, k in enumerate(self.KERNEL_CLASSES)}

    def profile_get(self) -> List[Tuple[str, float]]:
        names = C.POINTER(C.c_char_p)()
        ms = C.POINTER(C.c_float)()
        fl = C.POINTER(C.c_double)()
        n = lib().brn_profile_get(self._h, C.byref(names), C.byref(ms), C.byref(fl))
        return [(names[i].decode(), float(ms[i])) for i in range(n)]
