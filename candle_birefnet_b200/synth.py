"""Seeded synthetic weights / inputs for benchmarks and smoke runs (there is no network for the HF checkpoint).

Every tensor of the schema gets non-trivial values from its own numpy PCG64 stream keyed by
(seed, crc32(name)), so the values do not depend on generation order and are identical on every machine:
Linear/Conv W ~ N(0, 1/fan_in), biases U(+-1/sqrt(fan_in)), LN/BN gamma 1+0.1N, beta 0.1N, BN mean 0.1N,
var U(0.5,1.5), relative-position table 0.5N.  weight-set "A" zeroes `*.offset_conv.*` / `*.modulator_conv.*`
(offsets 0, modulator 1: deformable conv == the reference's CPU fallback); "B" draws them so that the
offsets are ~N(0, offset_sigma) pixels.
"""
from __future__ import annotations

import zlib
from typing import Dict, Mapping, Tuple

import numpy as np


def synthetic_weights(schema: Mapping[str, Tuple[int, ...]], seed: int = 0, weight_set: str = "A",
                      offset_sigma: float = 2.0) -> Dict[str, np.ndarray]:
    assert weight_set in ("A", "B")
    out: Dict[str, np.ndarray] = {}
    for name, shape in schema.items():
        shape = tuple(int(s) for s in shape)
        rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))
        is_off = ".offset_conv." in name or ".modulator_conv." in name
        if name.endswith(".running_mean"):
            v = 0.1 * rng.standard_normal(shape)
        elif name.endswith(".running_var"):
            v = rng.uniform(0.5, 1.5, shape)
        elif name.endswith("relative_position_bias_table"):
            v = 0.5 * rng.standard_normal(shape)
        elif name.endswith(".weight") and len(shape) == 1:          # LayerNorm / BatchNorm gamma
            v = 1.0 + 0.1 * rng.standard_normal(shape)
        elif name.endswith(".weight"):
            fan_in = int(np.prod(shape[1:]))
            v = rng.standard_normal(shape) * np.sqrt(1.0 / fan_in)
        elif name.endswith(".bias"):
            wshape = tuple(schema[name[:-5] + ".weight"])
            if len(wshape) == 1:                                     # LayerNorm / BatchNorm beta
                v = 0.1 * rng.standard_normal(shape)
            else:
                v = rng.uniform(-1.0, 1.0, shape) / (1.0 if is_off else np.sqrt(float(np.prod(wshape[1:]))))
        else:
            raise ValueError(f"unrecognised tensor name {name}")
        if is_off:
            if weight_set == "A":
                v = np.zeros(shape)
            elif ".offset_conv." in name:
                v = v * offset_sigma if name.endswith(".weight") else v * 0.5 * offset_sigma
        out[name] = np.ascontiguousarray(v, dtype=np.float32)
    return out


def synthetic_input(b: int, h: int, w: int, seed: int = 1234) -> np.ndarray:
    """x ~ N(0,1), the distribution the reference's benches use (examples/bench_inference.rs:30)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.standard_normal((b, 3, h, w)).astype(np.float32)
