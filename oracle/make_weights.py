"""Seeded synthetic weights with the reference's safetensors key schema -- TEST INFRASTRUCTURE ONLY.

Key names and shapes follow what `BiRefNet::new` asks its VarBuilder for
(src/birefnet.rs:389-409, :170-273; src/swin.rs:98-99,130-141,333-338,486-487,677-680,754;
src/decoder.rs:44-45,65,104-114; src/aspp.rs:39-45,247-290) == the keys of HF
`ZhengPeng7/BiRefNet/model.safetensors` (687 tensors / 220,202,578 params for swin_l).

Every tensor is non-trivial (candle's VarMap init would leave the relative-position table,
LN affine and BN statistics at identity).  Each tensor is drawn from its own numpy PCG64 stream keyed by
(seed, crc32(name)), so values do not depend on generation order and are reproducible across machines.

weight-set "A": `*.offset_conv.*` and `*.modulator_conv.*` are zero, so offsets == 0 and modulator == 2*sigmoid(0) == 1:
               the reference's CPU fallback (plain conv) and true deformable conv coincide.
weight-set "B": random offset/modulator convs, scaled so offsets are ~N(0, offset_sigma) pixels.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np

from .birefnet_ref import Config


def schema(cfg: Config) -> "OrderedDict[str, Tuple[Tuple[int, ...], str]]":
    """name -> (shape, kind) with kind in {w, b, ln_g, ln_b, bn_m, bn_v, bn_g, bn_b, table}."""
    s: "OrderedDict[str, Tuple[Tuple[int, ...], str]]" = OrderedDict()

    def lin(p, o, i, bias=True):
        s[p + ".weight"] = ((o, i), "w")
        if bias:
            s[p + ".bias"] = ((o,), "b")

    def ln(p, c):
        s[p + ".weight"] = ((c,), "ln_g")
        s[p + ".bias"] = ((c,), "ln_b")

    def conv(p, o, i, k, bias=True):
        s[p + ".weight"] = ((o, i, k, k), "w")
        if bias:
            s[p + ".bias"] = ((o,), "b")

    def bn(p, c):
        s[p + ".running_mean"] = ((c,), "bn_m")
        s[p + ".running_var"] = ((c,), "bn_v")
        s[p + ".weight"] = ((c,), "bn_g")
        s[p + ".bias"] = ((c,), "bn_b")

    E = cfg.embed_dim
    conv("bb.patch_embed.proj", E, 3, cfg.patch_size)
    ln("bb.patch_embed.norm", E)
    nstage = len(cfg.depths)
    for i in range(nstage):
        C = E << i
        for j in range(cfg.depths[i]):
            p = f"bb.layers.{i}.blocks.{j}"
            ln(p + ".norm1", C)
            lin(p + ".attn.qkv", 3 * C, C)
            lin(p + ".attn.proj", C, C)
            s[p + ".attn.relative_position_bias_table"] = (((2 * cfg.window_size - 1) ** 2, cfg.num_heads[i]), "table")
            ln(p + ".norm2", C)
            lin(p + ".mlp.fc1", cfg.mlp_ratio * C, C)
            lin(p + ".mlp.fc2", C, cfg.mlp_ratio * C)
        if i < nstage - 1:
            ln(f"bb.layers.{i}.downsample.norm", 4 * C)
            lin(f"bb.layers.{i}.downsample.reduction", 2 * C, 4 * C, bias=False)
        ln(f"bb.norm{i}", C)

    def dec_blk(p, cin, cout):
        conv(p + ".conv_in", 64, cin, 3)
        bn(p + ".bn_in", 64)
        a = p + ".dec_att"
        branches = [(a + ".aspp1", 1)] + [(f"{a}.aspp_deforms.{i}", k) for i, k in enumerate((1, 3, 7))]
        for bp, k in branches:
            conv(bp + ".atrous_conv.offset_conv", 2 * k * k, 64, k)
            conv(bp + ".atrous_conv.modulator_conv", k * k, 64, k)
            conv(bp + ".atrous_conv.regular_conv", 256, 64, k, bias=False)
            bn(bp + ".bn", 256)
        conv(a + ".global_avg_pool.1", 256, 64, 1, bias=False)
        bn(a + ".global_avg_pool.2", 256)
        conv(a + ".conv1", 64, 1280, 1, bias=False)
        bn(a + ".bn1", 64)
        conv(p + ".conv_out", cout, 64, 3)
        bn(p + ".bn_out", cout)

    d = cfg.decoder_channels()
    lat = d["lat"]
    dec_blk("squeeze_module.0", cfg.x4_channels(), lat[3])
    for n in range(5):
        conv(f"decoder.ipt_blk{n + 1}.conv1", 64, d["ipt_in"][n], 3)
        conv(f"decoder.ipt_blk{n + 1}.conv_out", d["ipt_out"][n], 64, 3)
    for n, (ci, co) in zip((4, 3, 2, 1), zip(d["dec_in"], d["dec_out"])):
        dec_blk(f"decoder.decoder_block{n}", ci, co)
    for n, c in zip((4, 3, 2), (lat[2], lat[1], lat[0])):
        conv(f"decoder.lateral_block{n}.conv", c, c, 1)
    for n, c in zip((4, 3, 2), d["dec_out"][:3]):
        conv(f"decoder.gdt_convs_{n}.0", 16, c, 3)
        bn(f"decoder.gdt_convs_{n}.1", 16)
        conv(f"decoder.gdt_convs_attn_{n}.0", 1, 16, 1)
        conv(f"decoder.gdt_convs_pred_{n}.0", 1, 16, 1)       # loaded, unused in forward (birefnet.rs:230-232)
        conv(f"decoder.conv_ms_spvn_{n}", 1, c, 1)            # loaded, unused in forward (birefnet.rs:241-243)
    conv("decoder.conv_out1.0", 1, d["final"], 1)
    return s


def make_weights(cfg: Config, seed: int = 0, weight_set: str = "A", offset_sigma: float = 2.0) -> Dict[str, np.ndarray]:
    """float32 numpy tensors for every key of `schema(cfg)`; values from the package's seeded generator
    (candle_birefnet_b200/synth.py) so tests, bench and smoke runs share one definition of "random-init"."""
    from candle_birefnet_b200.synth import synthetic_weights
    return synthetic_weights({k: v[0] for k, v in schema(cfg).items()}, seed, weight_set, offset_sigma)


def as_torch(weights: Dict[str, np.ndarray], dtype=None):
    import torch
    return {k: (torch.from_numpy(v) if dtype is None else torch.from_numpy(v).to(dtype)) for k, v in weights.items()}


def make_input(b: int, h: int, w: int, seed: int = 1234) -> np.ndarray:
    from candle_birefnet_b200.synth import synthetic_input
    return synthetic_input(b, h, w, seed)


def count_params(cfg: Config) -> Tuple[int, int]:
    sc = schema(cfg)
    return len(sc), int(sum(int(np.prod(s)) for s, _ in sc.values()))


if __name__ == "__main__":
    print(count_params(Config.swin_l()), count_params(Config.mini()))
