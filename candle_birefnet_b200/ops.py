"""Operator-level entry points (host numpy in / out) -- the reference's native boundaries, used by parity tests."""
from __future__ import annotations

import ctypes as C
from ctypes import c_float as C_float

import numpy as np

from . import _lib
from ._lib import check, lib

_PREC = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "fp16": _lib.PREC_FP16}


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def linear(a, w, bias=None, residual=None, act: int = 0, precision: str = "bf16", device: int = 0) -> np.ndarray:
    """candle_nn::linear (+ activation / residual epilogue): act 0 none, 1 relu, 2 exact-erf gelu."""
    a, w, bias, residual = _f32(a), _f32(w), _f32(bias), _f32(residual)
    M, K = a.shape
    N = w.shape[0]
    assert w.shape == (N, K)
    out = np.empty((M, N), dtype=np.float32)
    check(lib().brn_linear(device, _PREC[precision], _p(a), _p(w), _p(bias), _p(residual), M, N, K, act, _p(out)))
    return out


def ln_linear(x, gamma, beta, w, bias=None, act: int = 0, precision: str = "fp16", device: int = 0) -> np.ndarray:
    """LayerNorm folded into the consuming linear layer (norm1 -> qkv, norm2 -> fc1; src/swin.rs:355,217,407,104)."""
    x, gamma, beta, w, bias = _f32(x), _f32(gamma), _f32(beta), _f32(w), _f32(bias)
    M, K = x.shape
    N = w.shape[0]
    assert w.shape == (N, K) and gamma.shape == (K,) and beta.shape == (K,)
    out = np.empty((M, N), dtype=np.float32)
    check(lib().brn_ln_linear(device, _PREC[precision], _p(x), _p(gamma), _p(beta), _p(w), _p(bias), M, N, K, act, _p(out)))
    return out


def swin_mlp(x, gamma, beta, w1, b1, w2, b2, precision: str = "fp16", fused: int = -1, device: int = 0,
             with_stats: bool = False):
    """x + fc2(gelu_erf(fc1(LayerNorm(x)))): `Mlp::forward` inside the Swin block (src/swin.rs:103-107,407).

    fused=1 demands the single-kernel path (C in {128, 192}, hidden = 4C), 0 the two-GEMM path, -1 the model's own
    choice.  with_stats: also return the [M, 2] (mean, rstd) the epilogue emits for the next block's folded norm1."""
    x, gamma, beta, w1, b1, w2, b2 = (_f32(t) for t in (x, gamma, beta, w1, b1, w2, b2))
    M, Cc = x.shape
    hidden = w1.shape[0]
    assert w1.shape == (hidden, Cc) and w2.shape == (Cc, hidden)
    out = np.empty((M, Cc), dtype=np.float32)
    st = np.empty((M, 2), dtype=np.float32) if with_stats else None
    check(lib().brn_swin_mlp(device, _PREC[precision], _p(x), _p(gamma), _p(beta), _p(w1), _p(b1), _p(w2), _p(b2), M, Cc,
                             hidden, fused, _p(out), _p(st)))
    return (out, st) if with_stats else out


def conv2d(x, weight, bias=None, act: int = 0, precision: str = "bf16", device: int = 0) -> np.ndarray:
    """candle_nn::conv2d, stride 1, padding k//2, NCHW."""
    x, weight, bias = _f32(x), _f32(weight), _f32(bias)
    B, Cin, H, W = x.shape
    O, _, k, _ = weight.shape
    out = np.empty((B, O, H, W), dtype=np.float32)
    check(lib().brn_conv2d(device, _PREC[precision], _p(x), _p(weight), _p(bias), B, Cin, H, W, O, k, act, _p(out)))
    return out


def deform_conv2d(x, offset, mask, weight, bias=None, stride: int = 1, padding: int = None, precision: str = "bf16",
                  device: int = 0) -> np.ndarray:
    """DeformableConv2d / DeformConvASPP Metal math (src/deform_conv.rs:102-215, src/aspp.rs:58-165) ==
    torchvision.ops.deform_conv2d(x, offset, weight, bias, stride, padding, mask=mask)."""
    x, offset, mask, weight, bias = _f32(x), _f32(offset), _f32(mask), _f32(weight), _f32(bias)
    B, Cin, H, W = x.shape
    O, _, k, _ = weight.shape
    if padding is None:
        padding = k // 2
    Ho, Wo = (H + 2 * padding - k) // stride + 1, (W + 2 * padding - k) // stride + 1
    assert offset.shape == (B, 2 * k * k, Ho, Wo) and mask.shape == (B, k * k, Ho, Wo)
    out = np.empty((B, O, Ho, Wo), dtype=np.float32)
    check(lib().brn_deform_conv2d(device, _PREC[precision], _p(x), _p(offset), _p(mask), _p(weight), _p(bias), B, Cin,
                                  H, W, O, k, stride, padding, _p(out)))
    return out


class DeformableConv2d:
    """DeformableConv2d (src/deform_conv.rs:16-99; crate root export src/lib.rs:13): `new(in, out, k, stride, padding, vb)`
    with vb keys offset_conv.{weight,bias}, modulator_conv.{weight,bias}, regular_conv.{weight,bias}; `forward(x)`."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, stride: int, padding: int, vb,
                 precision: str = "fp16", deform_mode: str = "deformable", device: int = 0):
        k = kernel_size
        shapes = {"offset_conv.weight": (2 * k * k, in_channels, k, k), "offset_conv.bias": (2 * k * k,),
                  "modulator_conv.weight": (k * k, in_channels, k, k), "modulator_conv.bias": (k * k,),
                  "regular_conv.weight": (out_channels, in_channels, k, k), "regular_conv.bias": (out_channels,)}
        self.w = {}
        for key, shp in shapes.items():
            if key not in vb:
                raise _lib.BrnError(3, f"cannot find tensor {key}")          # candle: vb.get fails inside new()
            a = _f32(vb[key])
            if a.shape != shp:
                raise _lib.BrnError(5, f"shape mismatch for {key}: {a.shape} != {shp}")
            self.w[key] = a
        self.cin, self.cout, self.k, self.stride, self.padding = in_channels, out_channels, k, stride, padding
        self.precision, self.deform_mode, self.device = precision, deform_mode, device

    def forward(self, x) -> np.ndarray:
        x = _f32(x)
        B, Cin, H, W = x.shape
        if Cin != self.cin:
            raise _lib.BrnError(5, f"expected {self.cin} input channels, got {Cin}")
        k, s, p = self.k, self.stride, self.padding
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        out = np.empty((B, self.cout, Ho, Wo), dtype=np.float32)
        w = self.w
        mode = {"cpu_fallback": _lib.DEFORM_CPU_FALLBACK, "deformable": _lib.DEFORM_DEFORMABLE}[self.deform_mode]
        check(lib().brn_deformable_conv2d(self.device, _PREC[self.precision], mode, _p(x), B, Cin, H, W,
                                          _p(w["offset_conv.weight"]), _p(w["offset_conv.bias"]),
                                          _p(w["modulator_conv.weight"]), _p(w["modulator_conv.bias"]),
                                          _p(w["regular_conv.weight"]), _p(w["regular_conv.bias"]), self.cout, k, s, p,
                                          _p(out)))
        return out

    __call__ = forward


def window_attention(qkv, bias, hp: int, wp: int, shift: int, precision: str = "bf16", device: int = 0) -> np.ndarray:
    """WindowAttention::forward_standard (src/swin.rs:266-311) on window-ordered qkv [n_windows,144,3*heads*32]."""
    qkv, bias = _f32(qkv), _f32(bias)
    nwin, n, c3 = qkv.shape
    heads = bias.shape[0]
    ws = {144: 12, 49: 7}[n]
    assert c3 == 3 * heads * 32 and bias.shape == (heads, n, n)
    out = np.empty((nwin, n, heads * 32), dtype=np.float32)
    check(lib().brn_window_attention(device, _PREC[precision], _p(qkv), _p(bias), nwin, heads, ws, hp, wp, shift, _p(out)))
    return out


def preprocess_rgb8(rgb: np.ndarray, H: int = 1024, W: int = 1024, device: int = 0) -> np.ndarray:
    """examples/infer_image.rs:44-67: Triangle resize_exact + ImageNet normalise.  rgb uint8 [h,w,3] or [B,h,w,3]."""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    if rgb.ndim == 3:
        rgb = rgb[None]
    B, h, w, c = rgb.shape
    assert c == 3
    out = np.empty((B, 3, H, W), dtype=np.float32)
    check(lib().brn_preprocess_rgb8(device, _p(rgb), B, h, w, H, W, _p(out)))
    return out


def postprocess_mask(logits: np.ndarray, orig_h: int, orig_w: int, device: int = 0) -> np.ndarray:
    """examples/infer_image.rs:85-105: sigmoid -> u8 -> Lanczos3 resize.  logits float32 [B,1,H,W] or [B,H,W]."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    if logits.ndim == 4:
        logits = logits[:, 0]
    B, H, W = logits.shape
    out = np.empty((B, orig_h, orig_w), dtype=np.uint8)
    check(lib().brn_postprocess_mask(device, _p(np.ascontiguousarray(logits)), B, H, W, orig_h, orig_w, _p(out)))
    return out


def bench_op(kind: str, B: int, H: int, W: int, Cin: int, N: int = 0, k: int = 1, act: int = 0, with_res: bool = False,
             out_f32: bool = False, iters: int = 20, precision: str = "fp16", device: int = 0) -> float:
    """Mean device ms per launch of one kernel on synthetic device-resident data (kind: gemm | attn | deform | mlp;
    mlp: M = B*H*W rows of width Cin, with_res = the fused kernel, else fc1 + fc2)."""
    ms = C_float()
    kid = {"gemm": 0, "attn": 1, "deform": 2, "mlp": 3}[kind]
    check(lib().brn_bench_op(device, _PREC[precision], kid, B, H, W, Cin, N, k, act, int(with_res), int(out_f32), iters,
                             C.byref(ms)))
    return float(ms.value)
