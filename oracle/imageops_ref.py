"""CPU oracle for the steps either side of the hot path -- TEST INFRASTRUCTURE ONLY.

Restates what `examples/infer_image.rs` does around `forward_logits`:

  * `:50`     `img.resize_exact(1024, 1024, FilterType::Triangle)`   -> `resize(..., "triangle")`
  * `:54-67`  ImageNet mean/std normalisation into `[1,3,1024,1024]`  -> `normalize_imagenet`
  * `:85-97`  `sigmoid` -> `(v * 255.0).clamp(0, 255) as u8`          -> `mask_to_u8`
  * `:100-105` `imageops::resize(&mask, orig_w, orig_h, Lanczos3)`    -> `resize(..., "lanczos3")`

The resampling code lives in a third-party crate that is NOT under /root/reference: `image` 0.25.9
(`Cargo.lock:1102-1105`).  This file restates the published algorithm of its `imageops::sample` module
(`resize` = `vertical_sample` into an f32 image, then `horizontal_sample` with clamp + round-half-away to the pixel
type; per output sample the source window is `[floor(c - s), ceil(c + s))` around `c = (o + 0.5) * ratio` with support
`s = filter.support * max(ratio, 1)`, weights `kernel((i - (c - 0.5)) / max(ratio, 1))` normalised to sum 1, all in
f32, accumulation `t += p * w` in tap order).  PARITY UNPINNED: the crate source is not in this image and the
reference holds no resized fixture; `tests/test_oracle.py` pins this restatement against PIL's BILINEAR / LANCZOS
(same filters, fixed-point coefficients) to <= 1 LSB.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32

IMAGENET_MEAN = (0.485, 0.456, 0.406)     # examples/infer_image.rs:54
IMAGENET_STD = (0.229, 0.224, 0.225)      # examples/infer_image.rs:55


def _triangle(x: np.float32) -> np.float32:
    """image::imageops::sample::triangle_kernel."""
    a = F(abs(x))
    return F(1.0) - a if a < F(1.0) else F(0.0)


def _sinc(t: np.float32) -> np.float32:
    a = F(t * F(math.pi))
    return F(1.0) if t == F(0.0) else F(F(np.sin(a)) / a)


def _lanczos3(x: np.float32) -> np.float32:
    """image::imageops::sample::lanczos3_kernel."""
    t = F(3.0)
    return F(_sinc(x) * _sinc(F(x / t))) if abs(x) < t else F(0.0)


FILTERS = {"triangle": (_triangle, F(1.0)), "lanczos3": (_lanczos3, F(3.0))}


def sample_weights(n_in: int, n_out: int, filt: str):
    """Per output index: (left, [normalised f32 weights]) -- the window/weight loop shared by vertical_sample and
    horizontal_sample."""
    kernel, support = FILTERS[filt]
    ratio = F(F(n_in) / F(n_out))
    sratio = ratio if ratio >= F(1.0) else F(1.0)
    src_support = F(support * sratio)
    out = []
    for o in range(n_out):
        c = F((F(o) + F(0.5)) * ratio)
        left = int(math.floor(F(c - src_support)))
        left = min(max(left, 0), n_in - 1)
        right = int(math.ceil(F(c + src_support)))
        right = min(max(right, left + 1), n_in)
        c = F(c - F(0.5))
        ws = [kernel(F(F(F(i) - c) / sratio)) for i in range(left, right)]
        s = F(0.0)
        for w in ws:
            s = F(s + w)
        out.append((left, [F(w / s) for w in ws]))
    return out


def _apply(img: np.ndarray, axis: int, n_out: int, filt: str) -> np.ndarray:
    """One separable pass in f32: out[o] = sum_i img[left + i] * w[i], `t += p * w` in tap order (no fused
    multiply-add: Rust evaluates the product and the sum as two rounded f32 operations)."""
    img = np.moveaxis(img.astype(np.float32), axis, 0)
    res = np.empty((n_out,) + img.shape[1:], np.float32)
    for o, (left, ws) in enumerate(sample_weights(img.shape[0], n_out, filt)):
        t = np.zeros(img.shape[1:], np.float32)
        for i, w in enumerate(ws):
            t = (t + (img[left + i] * w).astype(np.float32)).astype(np.float32)
        res[o] = t
    return np.moveaxis(res, 0, axis)


def resize(img_u8: np.ndarray, new_h: int, new_w: int, filt: str) -> np.ndarray:
    """image::imageops::resize on an 8-bit image [H,W] or [H,W,C]: vertical pass (f32 result), horizontal pass,
    clamp to [0,255], round half away from zero.  Same-size requests are a copy, like the crate."""
    assert img_u8.dtype == np.uint8
    if img_u8.shape[0] == new_h and img_u8.shape[1] == new_w:
        return img_u8.copy()
    tmp = _apply(img_u8, 0, new_h, filt)
    out = _apply(tmp, 1, new_w, filt)
    out = np.clip(out, F(0.0), F(255.0))
    fl = np.floor(out)
    return (fl + ((out - fl) >= F(0.5))).astype(np.uint8)   # f32::round (half away from zero) on non-negative values


def normalize_imagenet(rgb_u8: np.ndarray) -> np.ndarray:
    """examples/infer_image.rs:54-67: (p / 255 - mean) / std per channel -> [1,3,H,W] f32."""
    assert rgb_u8.dtype == np.uint8 and rgb_u8.ndim == 3 and rgb_u8.shape[2] == 3
    x = rgb_u8.astype(np.float32) / F(255.0)
    out = np.empty((1, 3) + rgb_u8.shape[:2], np.float32)
    for c in range(3):
        out[0, c] = ((x[..., c] - F(IMAGENET_MEAN[c])) / F(IMAGENET_STD[c])).astype(np.float32)
    return out


def preprocess(rgb_u8: np.ndarray, size: int = 1024) -> np.ndarray:
    """examples/infer_image.rs:44-69."""
    return normalize_imagenet(resize(rgb_u8, size, size, "triangle"))


def mask_to_u8(prob: np.ndarray) -> np.ndarray:
    """examples/infer_image.rs:93-97: `(v * 255.0).clamp(0.0, 255.0) as u8` (truncation)."""
    v = np.clip(prob.astype(np.float32) * F(255.0), F(0.0), F(255.0))
    return v.astype(np.uint8)


def postprocess(logits: np.ndarray, orig_h: int, orig_w: int) -> np.ndarray:
    """examples/infer_image.rs:85-105 on one [H,W] logit map: sigmoid -> u8 -> Lanczos3 resize to the original size."""
    x = logits.astype(np.float32)
    prob = (F(1.0) / (F(1.0) + np.exp(-x))).astype(np.float32)
    return resize(mask_to_u8(prob), orig_h, orig_w, "lanczos3")
