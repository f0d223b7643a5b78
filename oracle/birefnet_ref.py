"""CPU oracle for BiRefNet (Swin backbone) `forward_logits` -- TEST INFRASTRUCTURE ONLY.

This file is a PyTorch-CPU restatement of the reference crate's forward pass
(`/root/reference/src/{swin,birefnet,decoder,aspp,deform_conv}.rs`), written
op-for-op so every function can be checked against the Rust it follows (the
file:line of the Rust is cited in each docstring).  It is the *checker* for the
CUDA path: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it.  The product path
(`candle_birefnet_b200`) never does.

PARITY UNPINNED: the reference holds no golden vectors, known-answer tests or
fixtures for this path, and its arithmetic lives in an
un-vendored candle fork (candle-core/candle-nn 0.9.2 @ imperatormk/candle
674fa161, Cargo.lock:230-232) that cannot be built here (no Rust toolchain).
The oracle is therefore pinned only against independent formulations of the
same published algorithms (torch `F.*` ops, `torchvision.ops.deform_conv2d`,
HF `transformers` Swin, closed-form index maps) in `tests/test_oracle.py`.

Two deformable-conv modes exist because the reference's CPU path does not
deform:
  * ``cpu_fallback`` -- what candle-CPU really computes: `regular_conv(x)`
    (`src/aspp.rs:183-185`, `src/deform_conv.rs:95-98`).
  * ``deformable``   -- what the Metal path / PyTorch BiRefNet compute:
    torchvision DCNv2 semantics (`src/aspp.rs:58-165`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
W = Dict[str, Tensor]

WINDOW = 12          # src/swin.rs:74
SHIFT = WINDOW // 2  # src/swin.rs:548
LN_EPS = 1e-5        # src/swin.rs:333
BN_EPS = 1e-5        # src/decoder.rs:105
MASK_VALUE = -100.0  # src/swin.rs:651


@dataclass
class Config:
    """Mirror of SwinConfig (src/swin.rs:13-23) + BiRefNetConfig (src/birefnet.rs:13-30).

    head_dim 32 variants: window 12 (swin_b, swin_l) and window 7 (swin_t, swin_s).  `swin_l()` is the one
    the reference ever builds (src/birefnet.rs:390-391); `mini()` is a reduced
    width/depth variant of the same architecture used to keep CPU tests fast.
    """
    embed_dim: int = 192
    depths: Tuple[int, ...] = (2, 2, 18, 2)
    num_heads: Tuple[int, ...] = (6, 12, 24, 48)
    window_size: int = WINDOW
    mlp_ratio: int = 4
    patch_size: int = 4
    name: str = "swin_l"

    @staticmethod
    def swin_l() -> "Config":
        return Config()

    # src/swin.rs:54-66
    @staticmethod
    def swin_b() -> "Config":
        return Config(embed_dim=128, num_heads=(4, 8, 16, 32), name="swin_b")

    # src/swin.rs:27-38 and :41-52 (window 7: 49-token windows, shift 3)
    @staticmethod
    def swin_t() -> "Config":
        return Config(embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), window_size=7, name="swin_t")

    @staticmethod
    def swin_s() -> "Config":
        return Config(embed_dim=96, depths=(2, 2, 18, 2), num_heads=(3, 6, 12, 24), window_size=7, name="swin_s")

    @staticmethod
    def mini7() -> "Config":
        return Config(embed_dim=64, depths=(2, 2, 2, 2), num_heads=(2, 4, 8, 16), window_size=7, name="mini7")

    @staticmethod
    def mini() -> "Config":
        return Config(embed_dim=64, depths=(2, 2, 2, 2), num_heads=(2, 4, 8, 16), name="mini")

    # src/swin.rs:83-87
    def stage_channels(self) -> List[int]:
        return [self.embed_dim << i for i in range(len(self.depths))]

    # src/birefnet.rs:50-53 (mul_scl_ipt doubles)
    def lateral_channels(self) -> List[int]:
        return [2 * c for c in self.stage_channels()]

    # src/birefnet.rs:56-61
    def x4_channels(self) -> int:
        lat = self.lateral_channels()
        return lat[3] + lat[0] + lat[1] + lat[2]

    # src/birefnet.rs:180,202-207
    def decoder_channels(self):
        lat = self.lateral_channels()
        ipt_out = [48, 96, 192, 384, 384]
        # ipt_blk inputs are the image2patches channel counts 3*g*g; for swin_l they
        # coincide with the lat_ch expressions of src/birefnet.rs:189-193.
        ipt_in = [3, 48, 192, 768, 3072]
        dec_out = [lat[2], lat[1], lat[0], lat[0] // 2]
        dec_in = [lat[3] + ipt_out[4], dec_out[0] + ipt_out[3], dec_out[1] + ipt_out[2], dec_out[2] + ipt_out[1]]
        return dict(lat=lat, ipt_in=ipt_in, ipt_out=ipt_out, dec_in=dec_in, dec_out=dec_out,
                    final=dec_out[3] + ipt_out[0])


# ----------------------------------------------------------------------------------------------
# candle op vocabulary
# ----------------------------------------------------------------------------------------------

def linear(x: Tensor, w: W, p: str, bias: bool = True) -> Tensor:
    """candle_nn::linear / linear_no_bias: x @ W^T (+ b)."""
    return F.linear(x, w[p + ".weight"], w[p + ".bias"] if bias else None)


def layer_norm(x: Tensor, w: W, p: str) -> Tensor:
    """candle_nn::layer_norm(dim, 1e-5): biased variance, eps inside sqrt."""
    return F.layer_norm(x, (x.shape[-1],), w[p + ".weight"], w[p + ".bias"], LN_EPS)


def conv2d(x: Tensor, w: W, p: str, stride: int = 1, padding: int = 0, bias: bool = True) -> Tensor:
    """candle_nn::conv2d / conv2d_no_bias: cross-correlation, zero padding."""
    return F.conv2d(x, w[p + ".weight"], w[p + ".bias"] if bias else None, stride=stride, padding=padding)


def batch_norm_eval(x: Tensor, w: W, p: str) -> Tensor:
    """candle_nn::batch_norm(C,1e-5).forward_t(x,false): (x-mu)/sqrt(var+eps)*g+b."""
    return F.batch_norm(x, w[p + ".running_mean"], w[p + ".running_var"], w[p + ".weight"], w[p + ".bias"],
                        training=False, eps=BN_EPS)


def upsample_bilinear2d(x: Tensor, h: int, wd: int) -> Tensor:
    """Tensor::upsample_bilinear2d(h, w, align_corners=true) (src/birefnet.rs, 16 sites)."""
    return F.interpolate(x, size=(h, wd), mode="bilinear", align_corners=True)


# ----------------------------------------------------------------------------------------------
# swin.rs
# ----------------------------------------------------------------------------------------------

def build_relative_position_index(ws: int) -> Tensor:
    """src/swin.rs:166-210: index[(i,j),(k,l)] = (i-k+ws-1)*(2ws-1) + (j-l+ws-1)."""
    n = ws * ws
    idx = torch.zeros((n, n), dtype=torch.long)
    for i in range(ws):
        for j in range(ws):
            for k in range(ws):
                for l in range(ws):
                    idx[i * ws + j, k * ws + l] = (i - k + ws - 1) * (2 * ws - 1) + (j - l + ws - 1)
    return idx


_REL_INDEX_CACHE: Dict[int, Tensor] = {}


def cached_bias(table: Tensor, ws: int) -> Tensor:
    """src/swin.rs:148-152: table[index].reshape(N,N,heads).permute(2,0,1) -> [heads,N,N]."""
    if ws not in _REL_INDEX_CACHE:
        _REL_INDEX_CACHE[ws] = build_relative_position_index(ws)
    n = ws * ws
    bias = table.index_select(0, _REL_INDEX_CACHE[ws].flatten())
    return bias.reshape(n, n, -1).permute(2, 0, 1).contiguous()


def roll_2d(x: Tensor, shift_h: int, shift_w: int) -> Tensor:
    """src/swin.rs:412-444: roll via narrow + cat on dims 1 and 2 of [B,H,W,C]."""
    _, h, wd, _ = x.shape
    shift_h = ((shift_h % h) + h) % h
    shift_w = ((shift_w % wd) + wd) % wd
    if shift_h == 0 and shift_w == 0:
        return x
    if shift_h > 0:
        x = torch.cat([x.narrow(1, h - shift_h, shift_h), x.narrow(1, 0, h - shift_h)], 1)
    if shift_w > 0:
        x = torch.cat([x.narrow(2, wd - shift_w, shift_w), x.narrow(2, 0, wd - shift_w)], 2)
    return x


def window_partition(x: Tensor, ws: int) -> Tensor:
    """src/swin.rs:446-459."""
    b, h, wd, c = x.shape
    x = x.reshape(b, h // ws, ws, wd // ws, ws, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(b * (h // ws) * (wd // ws), ws * ws, c)


def window_reverse(windows: Tensor, ws: int, h: int, wd: int) -> Tensor:
    """src/swin.rs:461-475."""
    c = windows.shape[-1]
    b = windows.shape[0] // ((h // ws) * (wd // ws))
    x = windows.reshape(b, h // ws, wd // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(b, h, wd, c)


def create_attention_mask(hp: int, wp: int, ws: int, shift: int, dtype) -> Tensor:
    """src/swin.rs:603-655: 3x3 region ids, partition, -100 where ids differ. -> [nW,N,N]."""
    img = torch.zeros((hp, wp), dtype=torch.float32)
    h_slices = [(0, hp - ws), (hp - ws, hp - shift), (hp - shift, hp)]
    w_slices = [(0, wp - ws), (wp - ws, wp - shift), (wp - shift, wp)]
    cnt = 0
    for hs, he in h_slices:
        for wsl, we in w_slices:
            img[hs:he, wsl:we] = cnt
            cnt += 1
    mask = img.reshape(1, hp // ws, ws, wp // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws)
    diff = mask.unsqueeze(1) - mask.unsqueeze(2)
    return torch.where(diff != 0, torch.full_like(diff, MASK_VALUE), torch.zeros_like(diff)).to(dtype)


def window_attention(xw: Tensor, w: W, p: str, heads: int, ws: int, mask: Tensor | None) -> Tensor:
    """WindowAttention::forward + forward_standard (src/swin.rs:212-311)."""
    b_, n, c = xw.shape
    hd = c // heads
    scale = float(hd) ** -0.5                                           # :134
    qkv = linear(xw, w, p + ".qkv").reshape(b_, n, 3, heads, hd).permute(2, 0, 3, 1, 4)  # :217-219
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * scale                                                        # :278
    attn = q @ k.transpose(-2, -1)                                       # :281
    attn = attn + cached_bias(w[p + ".relative_position_bias_table"], ws).unsqueeze(0)  # :284-285
    if mask is not None:                                                 # :288-297
        nw = mask.shape[0]
        attn = attn.reshape(b_ // nw, nw, heads, n, n) + mask.unsqueeze(0).unsqueeze(2)
        attn = attn.reshape(b_, heads, n, n)
    attn = torch.softmax(attn, dim=-1)                                   # :300
    x = (attn @ v).transpose(1, 2).reshape(b_, n, c)                     # :303-307
    return linear(x, w, p + ".proj")                                     # :310


def mlp(x: Tensor, w: W, p: str) -> Tensor:
    """Mlp::forward (src/swin.rs:103-107): fc1 -> exact-erf GELU -> fc2."""
    return linear(F.gelu(linear(x, w, p + ".fc1")), w, p + ".fc2")


def swin_block(x: Tensor, h: int, wd: int, w: W, p: str, heads: int, ws: int, shift: int,
               attn_mask: Tensor) -> Tensor:
    """SwinTransformerBlock::forward (src/swin.rs:350-410)."""
    b, l, c = x.shape
    assert l == h * wd                                                   # :352
    shortcut = x
    x = layer_norm(x, w, p + ".norm1").reshape(b, h, wd, c)              # :355-356
    pad_r = (ws - wd % ws) % ws                                          # :359-360
    pad_b = (ws - h % ws) % ws
    if pad_r > 0 or pad_b > 0:                                           # :362-366 zeros AFTER norm1
        x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b))
    hp, wp = x.shape[1], x.shape[2]
    if shift > 0:                                                        # :371-377
        x = roll_2d(x, -shift, -shift)
    xw = window_partition(x, ws)                                         # :380
    mask = attn_mask if shift > 0 else None                              # :383
    aw = window_attention(xw, w, p + ".attn", heads, ws, mask)           # :384
    x = window_reverse(aw, ws, hp, wp)                                   # :387
    if shift > 0:                                                        # :390-394
        x = roll_2d(x, shift, shift)
    if pad_r > 0 or pad_b > 0:                                           # :397-401
        x = x[:, :h, :wd, :]
    x = x.reshape(b, h * wd, c)
    x = shortcut + x                                                     # :406
    return x + mlp(layer_norm(x, w, p + ".norm2"), w, p + ".mlp")        # :407


def patch_merging(x: Tensor, h: int, wd: int, w: W, p: str) -> Tensor:
    """PatchMerging::forward (src/swin.rs:491-527): order (0,0),(1,0),(0,1),(1,1)."""
    b, _, c = x.shape
    x = x.reshape(b, h, wd, c)
    if h % 2 == 1 or wd % 2 == 1:                                        # :496-503
        x = F.pad(x, (0, 0, 0, wd % 2, 0, h % 2))
        h, wd = h + h % 2, wd + wd % 2
    x = x.reshape(b, h // 2, 2, wd // 2, 2, c)
    x0, x1, x2, x3 = x[:, :, 0, :, 0], x[:, :, 1, :, 0], x[:, :, 0, :, 1], x[:, :, 1, :, 1]  # :510-516
    x = torch.cat([x0, x1, x2, x3], -1).reshape(b, (h // 2) * (wd // 2), 4 * c)
    return linear(layer_norm(x, w, p + ".norm"), w, p + ".reduction", bias=False)  # :525-526


def basic_layer(x: Tensor, h: int, wd: int, w: W, p: str, depth: int, heads: int, ws: int, downsample: bool):
    """BasicLayer::forward (src/swin.rs:578-601)."""
    hp = ((h + ws - 1) // ws) * ws
    wp = ((wd + ws - 1) // ws) * ws
    attn_mask = create_attention_mask(hp, wp, ws, ws // 2, x.dtype)      # :584
    for i in range(depth):
        shift = 0 if i % 2 == 0 else ws // 2                             # :552
        x = swin_block(x, h, wd, w, f"{p}.blocks.{i}", heads, ws, shift, attn_mask)
    x_out = x
    if downsample:
        return x_out, h, wd, patch_merging(x, h, wd, w, p + ".downsample"), (h + 1) // 2, (wd + 1) // 2
    return x_out, h, wd, x, h, wd


def patch_embed(x: Tensor, w: W, p: str, patch: int) -> Tensor:
    """PatchEmbed::forward (src/swin.rs:692-714): conv k=s=patch, LN over C."""
    _, _, h, wd = x.shape
    if wd % patch != 0 or h % patch != 0:
        x = F.pad(x, (0, (patch - wd % patch) % patch, 0, (patch - h % patch) % patch))
    x = conv2d(x, w, p + ".proj", stride=patch)
    b, c, wh, ww = x.shape
    x = layer_norm(x.flatten(2).transpose(1, 2), w, p + ".norm")
    return x.transpose(1, 2).reshape(b, c, wh, ww)


def swin_forward(x: Tensor, w: W, cfg: Config, p: str = "bb") -> List[Tensor]:
    """SwinTransformer::forward (src/swin.rs:768-797) -> 4 NCHW feature maps."""
    x = patch_embed(x, w, p + ".patch_embed", cfg.patch_size)
    _, _, h, wd = x.shape
    x = x.flatten(2).transpose(1, 2)                                     # :774
    outs = []
    n = len(cfg.depths)
    for i in range(n):
        x_out, oh, ow, x, h, wd = basic_layer(x, h, wd, w, f"{p}.layers.{i}", cfg.depths[i], cfg.num_heads[i],
                                              cfg.window_size, i < n - 1)
        xn = layer_norm(x_out, w, f"{p}.norm{i}")                        # :784
        outs.append(xn.reshape(xn.shape[0], oh, ow, -1).permute(0, 3, 1, 2))  # :786-788
    return outs


# ----------------------------------------------------------------------------------------------
# aspp.rs / deform_conv.rs / decoder.rs
# ----------------------------------------------------------------------------------------------

def deform_conv_aspp(x: Tensor, w: W, p: str, k: int, mode: str) -> Tensor:
    """DeformConvASPP::forward (src/aspp.rs:168-187); Metal math at :58-165."""
    pad = k // 2                                                         # :258
    offset = conv2d(x, w, p + ".offset_conv", padding=pad)               # :171
    mask = 2.0 / (1.0 + torch.exp(-conv2d(x, w, p + ".modulator_conv", padding=pad)))  # :173-174
    if mode == "cpu_fallback":                                           # :183-185
        return conv2d(x, w, p + ".regular_conv", padding=pad, bias=False)
    assert mode == "deformable"
    import torchvision
    return torchvision.ops.deform_conv2d(x, offset, w[p + ".regular_conv.weight"], None, stride=1, padding=pad,
                                         dilation=1, mask=mask)


def deformable_conv2d(x: Tensor, w: W, p: str, k: int, stride: int, pad: int, mode: str) -> Tensor:
    """DeformableConv2d::forward (src/deform_conv.rs:82-99; Metal math :102-215): bias + stride variant."""
    offset = conv2d(x, w, p + ".offset_conv", stride=stride, padding=pad)
    mask = 2.0 / (1.0 + torch.exp(-conv2d(x, w, p + ".modulator_conv", stride=stride, padding=pad)))
    if mode == "cpu_fallback":
        return conv2d(x, w, p + ".regular_conv", stride=stride, padding=pad)
    import torchvision
    return torchvision.ops.deform_conv2d(x, offset, w[p + ".regular_conv.weight"], w[p + ".regular_conv.bias"],
                                         stride=stride, padding=pad, dilation=1, mask=mask)


def aspp_module_deformable(x: Tensor, w: W, p: str, k: int, mode: str) -> Tensor:
    """ASPPModuleDeformable::forward (src/aspp.rs:217-223)."""
    return F.relu(batch_norm_eval(deform_conv_aspp(x, w, p + ".atrous_conv", k, mode), w, p + ".bn"))


def aspp_deformable(x: Tensor, w: W, p: str, mode: str) -> Tensor:
    """ASPPDeformable::forward (src/aspp.rs:303-333)."""
    x1 = aspp_module_deformable(x, w, p + ".aspp1", 1, mode)
    outs = [aspp_module_deformable(x, w, f"{p}.aspp_deforms.{i}", k, mode) for i, k in enumerate((1, 3, 7))]
    _, _, h, wd = x.shape
    x5 = x.mean(dim=-2, keepdim=True).mean(dim=-1, keepdim=True)         # :314
    x5 = F.relu(batch_norm_eval(conv2d(x5, w, p + ".global_avg_pool.1", bias=False), w, p + ".global_avg_pool.2"))
    x5 = x5.expand(-1, -1, h, wd)                                        # :318 nearest from 1x1
    out = torch.cat([x1] + outs + [x5], 1)
    return F.relu(batch_norm_eval(conv2d(out, w, p + ".conv1", bias=False), w, p + ".bn1"))


def basic_dec_blk(x: Tensor, w: W, p: str, mode: str) -> Tensor:
    """BasicDecBlk::forward (src/decoder.rs:126-141)."""
    x = F.relu(batch_norm_eval(conv2d(x, w, p + ".conv_in", padding=1), w, p + ".bn_in"))
    x = aspp_deformable(x, w, p + ".dec_att", mode)
    return batch_norm_eval(conv2d(x, w, p + ".conv_out", padding=1), w, p + ".bn_out")


def simple_convs(x: Tensor, w: W, p: str) -> Tensor:
    """SimpleConvs::forward (src/decoder.rs:50-56): no activation in between."""
    return conv2d(conv2d(x, w, p + ".conv1", padding=1), w, p + ".conv_out", padding=1)


def gdt_convs(x: Tensor, w: W, p: str) -> Tensor:
    """GdtConvs::forward (src/birefnet.rs:111-118)."""
    return F.relu(batch_norm_eval(conv2d(x, w, p + ".0", padding=1), w, p + ".1"))


def image2patches(x: Tensor, th: int, tw: int) -> Tensor:
    """local fn image2patches (src/birefnet.rs:288-300)."""
    b, c, h, wd = x.shape
    gh, gw = h // th, wd // tw
    return x.reshape(b, c, gh, th, gw, tw).permute(0, 1, 2, 4, 3, 5).reshape(b, c * gh * gw, th, tw)


# ----------------------------------------------------------------------------------------------
# birefnet.rs
# ----------------------------------------------------------------------------------------------

def decoder_forward(x: Tensor, x1: Tensor, x2: Tensor, x3: Tensor, x4: Tensor, w: W, mode: str,
                    p: str = "decoder") -> Tensor:
    """BiRefNetDecoder::forward (src/birefnet.rs:278-376)."""
    _, _, h, wd = x.shape
    h3, w3 = x3.shape[2:]
    h2, w2 = x2.shape[2:]
    h1, w1 = x1.shape[2:]
    ipt5 = simple_convs(image2patches(x, h // 32, wd // 32), w, p + ".ipt_blk5")    # :304-305
    ipt4 = simple_convs(image2patches(x, h // 16, wd // 16), w, p + ".ipt_blk4")
    ipt3 = simple_convs(image2patches(x, h // 8, wd // 8), w, p + ".ipt_blk3")
    ipt2 = simple_convs(image2patches(x, h // 4, wd // 4), w, p + ".ipt_blk2")
    ipt1 = simple_convs(x, w, p + ".ipt_blk1")                                      # :320

    def gate(pk: Tensor, n: int) -> Tensor:                                          # :327-329
        g = gdt_convs(pk, w, f"{p}.gdt_convs_{n}")
        return pk * torch.sigmoid(conv2d(g, w, f"{p}.gdt_convs_attn_{n}.0"))

    p4 = gate(basic_dec_blk(torch.cat([x4, ipt5], 1), w, p + ".decoder_block4", mode), 4)
    p3_in = upsample_bilinear2d(p4, h3, w3) + conv2d(x3, w, p + ".lateral_block4.conv")          # :332-334
    p3 = gate(basic_dec_blk(torch.cat([p3_in, upsample_bilinear2d(ipt4, h3, w3)], 1), w,
                            p + ".decoder_block3", mode), 3)
    p2_in = upsample_bilinear2d(p3, h2, w2) + conv2d(x2, w, p + ".lateral_block3.conv")
    p2 = gate(basic_dec_blk(torch.cat([p2_in, upsample_bilinear2d(ipt3, h2, w2)], 1), w,
                            p + ".decoder_block2", mode), 2)
    p1_in = upsample_bilinear2d(p2, h1, w1) + conv2d(x1, w, p + ".lateral_block2.conv")
    p1 = basic_dec_blk(torch.cat([p1_in, upsample_bilinear2d(ipt2, h1, w1)], 1), w, p + ".decoder_block1", mode)
    final_in = torch.cat([upsample_bilinear2d(p1, h, wd), upsample_bilinear2d(ipt1, h, wd)], 1)  # :372-374
    return conv2d(final_in, w, p + ".conv_out1.0")                                   # :375


def features(x: Tensor, w: W, cfg: Config) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """First half of BiRefNet::forward_logits (src/birefnet.rs:412-454): x1..x3 and the cxt-concatenated x4."""
    _, _, h, wd = x.shape
    x1, x2, x3, x4 = swin_forward(x, w, cfg)
    xh = upsample_bilinear2d(x, h // 2, wd // 2)                         # :425
    f1, f2, f3, f4 = swin_forward(xh, w, cfg)                            # :426
    x1 = torch.cat([x1, upsample_bilinear2d(f1, *x1.shape[2:])], 1)      # :435-443
    x2 = torch.cat([x2, upsample_bilinear2d(f2, *x2.shape[2:])], 1)
    x3 = torch.cat([x3, upsample_bilinear2d(f3, *x3.shape[2:])], 1)
    x4 = torch.cat([x4, upsample_bilinear2d(f4, *x4.shape[2:])], 1)
    h4, w4 = x4.shape[2:]
    x4 = torch.cat([upsample_bilinear2d(x1, h4, w4), upsample_bilinear2d(x2, h4, w4),
                    upsample_bilinear2d(x3, h4, w4), x4], 1)             # :450-453
    return x1, x2, x3, x4


def forward_logits(x: Tensor, w: W, cfg: Config, mode: str = "cpu_fallback") -> Tensor:
    """BiRefNet::forward_logits (src/birefnet.rs:412-461)."""
    x1, x2, x3, x4 = features(x, w, cfg)
    x4 = basic_dec_blk(x4, w, "squeeze_module.0", mode)                  # :457 (SqueezeModule, :86-94)
    return decoder_forward(x, x1, x2, x3, x4, w, mode)                   # :460


def forward(x: Tensor, w: W, cfg: Config, mode: str = "cpu_fallback") -> Tensor:
    """BiRefNet::forward (src/birefnet.rs:466-469)."""
    return torch.sigmoid(forward_logits(x, w, cfg, mode))


def to_dtype(w: W, dtype) -> W:
    return {k: v.to(dtype) for k, v in w.items()}
