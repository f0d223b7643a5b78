"""GPU parity tests, operator level, through the C ABI (ctypes) against the oracle's statement of each op.

fp32 path (SIMT FMA): tolerance 1e-4 relative to the output scale.
bf16 path (tcgen05): the oracle is evaluated on bf16-rounded operands in fp64, so what remains is accumulation
order (fp32) and the bf16 rounding of stored activations; tolerances are written at each test.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import candle_birefnet_b200 as cb
from oracle import birefnet_ref as R

pytestmark = pytest.mark.gpu


def bf16r(a):
    return torch.from_numpy(np.asarray(a, dtype=np.float32)).bfloat16().float().numpy()


def r16(a, precision):
    """operands as the 16-bit tensor-core path sees them"""
    t = torch.from_numpy(np.asarray(a, dtype=np.float32))
    if precision == "bf16":
        return t.bfloat16().float().numpy()
    if precision == "fp16":
        return t.half().float().numpy()
    return t.numpy()


def relerr(got, exp):
    return float(np.abs(got - exp).max() / (np.abs(exp).max() + 1e-12))


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 192, 192), (1000, 576, 192), (257, 48, 96), (144, 3, 64),
                                    (5184, 2304, 768),
                                    # several work items per CTA with the N tile changing from item to item: bias staged
                                    # once per kernel (N <= 1152) / fetched one tile ahead (N > 1152)
                                    (20000, 768, 256), (20000, 1536, 128)])
@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
def test_linear(M, N, K, precision):
    rng = np.random.default_rng(M + N + K)
    a = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    got = cb.ops.linear(a, w, b, precision=precision)
    exp = r16(a, precision).astype(np.float64) @ r16(w, precision).astype(np.float64).T + b
    assert relerr(got, exp) < 2e-5


@pytest.mark.parametrize("act", [1, 2])
@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
def test_linear_epilogues(act, precision):
    rng = np.random.default_rng(act)
    M, N, K = 500, 256, 128
    a = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    res = rng.standard_normal((M, N)).astype(np.float32)
    got = cb.ops.linear(a, w, b, residual=res, act=act, precision=precision)
    aa, ww = r16(a, precision), r16(w, precision)
    z = torch.from_numpy(aa.astype(np.float64) @ ww.astype(np.float64).T + b)
    z = F.relu(z) if act == 1 else F.gelu(z)      # exact-erf GELU (src/swin.rs:105)
    exp = z.numpy() + res
    assert relerr(got, exp) < 2e-5


@pytest.mark.parametrize("C,O,k,H,W", [(64, 64, 3, 32, 32), (64, 256, 1, 16, 24), (64, 147, 7, 16, 16), (224, 64, 3, 64, 64),
                                        (48, 64, 3, 20, 12), (128, 16, 3, 8, 8),
                                        # 3x3 tap-group mode (N <= 80, 16x8-pixel tiles): ragged tiles, N = 80 / 16 / 1
                                        (96, 80, 3, 24, 40), (64, 16, 3, 40, 24), (64, 1, 3, 32, 16), (480, 64, 3, 17, 9)])
@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
def test_conv2d(C, O, k, H, W, precision):
    rng = np.random.default_rng(C + O + k)
    x = rng.standard_normal((2, C, H, W)).astype(np.float32)
    w = (rng.standard_normal((O, C, k, k)) / np.sqrt(C * k * k)).astype(np.float32)
    b = rng.standard_normal(O).astype(np.float32)
    got = cb.ops.conv2d(x, w, b, act=1, precision=precision)
    xx, ww = r16(x, precision), r16(w, precision)
    exp = F.relu(F.conv2d(torch.from_numpy(xx).double(), torch.from_numpy(ww).double(), torch.from_numpy(b).double(),
                          padding=k // 2)).numpy()
    assert relerr(got, exp) < 2e-5


@pytest.mark.parametrize("k,sigma", [(1, 0.0), (1, 2.0), (3, 0.5), (3, 8.0), (7, 2.0)])
@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
def test_deform_conv2d(k, sigma, precision):
    import torchvision
    rng = np.random.default_rng(k * 10 + int(sigma))
    B, C, H, W, O = 2, 64, 20, 28, 256
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = (rng.standard_normal((O, C, k, k)) / np.sqrt(C * k * k)).astype(np.float32)
    off = (rng.standard_normal((B, 2 * k * k, H, W)) * sigma).astype(np.float32)   # incl. far out-of-bounds samples
    msk = rng.uniform(0, 2, (B, k * k, H, W)).astype(np.float32)
    got = cb.ops.deform_conv2d(x, off, msk, w, precision=precision)
    xx, ww = r16(x, precision), r16(w, precision)
    exp = torchvision.ops.deform_conv2d(torch.from_numpy(xx).double(), torch.from_numpy(off).double(),
                                        torch.from_numpy(ww).double(), None, padding=k // 2,
                                        mask=torch.from_numpy(msk).double()).numpy()
    # 16-bit paths: the gathered, modulated samples are rounded once to bf16 / fp16 before the MMA
    assert relerr(got, exp) < {"bf16": 1e-2, "fp16": 2e-3, "fp32": 1e-4}[precision]


@pytest.mark.parametrize("nimg,hp,wp,heads,shift", [(1, 24, 36, 2, 0), (2, 24, 36, 2, 6), (1, 12, 12, 6, 6), (1, 72, 72, 24, 6),
                                                     (1, 264, 264, 6, 6)])
@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
def test_window_attention(nimg, hp, wp, heads, shift, precision):
    """softmax(scale q k^T + bias (+ -100 region mask)) v, the plain chain of examples/test_flash_bias.rs:30-36."""
    rng = np.random.default_rng(hp + heads + shift)
    nw = (hp // 12) * (wp // 12)
    C = heads * 32
    qkv = rng.standard_normal((nimg * nw, 144, 3 * C)).astype(np.float32)
    bias = (rng.standard_normal((heads, 144, 144)) * 0.5).astype(np.float32)
    got = cb.ops.window_attention(qkv, bias, hp, wp, shift, precision=precision)
    t = torch.from_numpy(r16(qkv, precision)).double()
    q, k, v = [t[..., i * C:(i + 1) * C].reshape(nimg * nw, 144, heads, 32).permute(0, 2, 1, 3) for i in range(3)]
    s = (q * 32 ** -0.5) @ k.transpose(-1, -2) + torch.from_numpy(bias).double()
    if shift:
        m = R.create_attention_mask(hp, wp, 12, 6, torch.float64)
        s = (s.reshape(nimg, nw, heads, 144, 144) + m[None, :, None]).reshape(nimg * nw, heads, 144, 144)
    exp = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(nimg * nw, 144, C).numpy()
    err = np.abs(got - exp).max()
    # 16-bit paths: q (scaled), bias, P and the output are rounded to bf16 / fp16; |out| <= ~1
    assert err < {"bf16": 3e-2, "fp16": 4e-3, "fp32": 2e-5}[precision], err


@pytest.mark.parametrize("M,N,K", [(300, 576, 192), (1000, 1152, 384), (777, 3072, 768), (200, 4608, 1536),
                                    # persistent CTAs walking several (M, N) tiles: column sums + bias staged once
                                    # (N = 576) / fetched one tile ahead (N = 1536)
                                    (40000, 576, 192), (20000, 1536, 384)])
@pytest.mark.parametrize("mean_over_std", [0.0, 3.0, 30.0])
@pytest.mark.parametrize("precision", ["fp16", "bf16"])
@pytest.mark.parametrize("act", [0, 2])
def test_ln_linear_fold(M, N, K, mean_over_std, precision, act):
    """LayerNorm folded into the consuming GEMM (norm1 -> qkv, norm2 -> fc1) against
    F.layer_norm -> linear in fp64.  The epilogue forms rstd * (acc - mean * colsum) + bias': with |row mean| >> row std
    acc and mean * colsum are both large and must cancel (fp32 accumulate, column sums of the ROUNDED weights), and the
    raw 16-bit copy of x carries a rounding error relative to |x|, not |x - mean| -- the tolerance scales with that."""
    if act == 2 and (K != 768 or mean_over_std == 30.0):
        pytest.skip("GELU variant: one shape")
    rng = np.random.default_rng(M + N + K + int(mean_over_std))
    x = rng.standard_normal((M, K))
    x = x * rng.uniform(0.5, 4.0, size=(M, 1)) + mean_over_std * rng.uniform(0.5, 1.5, size=(M, 1)) * np.sign(rng.standard_normal((M, 1)))
    x = x.astype(np.float32)
    gamma = (1.0 + 0.1 * rng.standard_normal(K)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(K)).astype(np.float32)
    w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    got = cb.ops.ln_linear(x, gamma, beta, w, b, act=act, precision=precision)
    xd = torch.from_numpy(x).double()
    ln = F.layer_norm(xd, (K,), torch.from_numpy(gamma).double(), torch.from_numpy(beta).double(), 1e-5)
    exp = ln @ torch.from_numpy(w).double().T + torch.from_numpy(b).double()
    if act == 2:
        exp = F.gelu(exp)
    exp = exp.numpy()
    eps = 2.0 ** -11 if precision == "fp16" else 2.0 ** -8
    # operand rounding (x relative to |x|/std, W) + output rounding, on O(1) outputs of K-term sums
    tol = eps * (4.0 + 2.0 * (1.0 + mean_over_std))
    err = np.abs(got - exp).max() / max(1.0, np.abs(exp).max())
    assert err < tol, (err, tol)


@pytest.mark.parametrize("M,C", [(1000, 192), (148 * 128 + 205, 192), (149 * 128, 192), (777, 128), (150 * 128 + 1, 128), (100, 192),
                                 (3 * 148 * 128 + 5, 192)])
@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_swin_mlp_fused(M, C, precision):
    """The fused MLP kernel of the early stages (mlp_tcgen05.cu) against x + fc2(gelu(fc1(LN(x)))) in fp64
    (src/swin.rs:103-107,407), against the two-GEMM path it replaces, and its emitted row statistics against the
    statistics of its own output.  Row counts cover partial tiles, one CTA per tile (M < 148 tiles) and 2-CTA clusters
    with an odd tile count (the second CTA of the last pair has no tile)."""
    rng = np.random.default_rng(M + C)
    hid = 4 * C
    x = (rng.standard_normal((M, C)) * rng.uniform(0.5, 3.0, size=(M, 1)) + rng.uniform(-2.0, 2.0, size=(M, 1))).astype(np.float32)
    gamma = (1.0 + 0.1 * rng.standard_normal(C)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(C)).astype(np.float32)
    w1 = (rng.standard_normal((hid, C)) / np.sqrt(C)).astype(np.float32)
    b1 = (0.5 * rng.standard_normal(hid)).astype(np.float32)
    w2 = (rng.standard_normal((C, hid)) / np.sqrt(hid)).astype(np.float32)
    b2 = (0.5 * rng.standard_normal(C)).astype(np.float32)
    got, st = cb.ops.swin_mlp(x, gamma, beta, w1, b1, w2, b2, precision=precision, fused=1, with_stats=True)
    two = cb.ops.swin_mlp(x, gamma, beta, w1, b1, w2, b2, precision=precision, fused=0)
    xd = torch.from_numpy(x).double()
    ln = F.layer_norm(xd, (C,), torch.from_numpy(gamma).double(), torch.from_numpy(beta).double(), 1e-5)
    h = F.gelu(ln @ torch.from_numpy(w1).double().T + torch.from_numpy(b1).double())
    exp = (xd + h @ torch.from_numpy(w2).double().T + torch.from_numpy(b2).double()).numpy()
    eps = 2.0 ** -11 if precision == "fp16" else 2.0 ** -8
    scale = max(1.0, np.abs(exp).max())
    err = np.abs(got - exp).max() / scale
    assert err < 12.0 * eps, (err, eps)
    # same operands, same roundings (hidden activations rounded to the operand type on both paths); fp32 accumulation
    # order over the hidden dimension differs (one 4C-long chain vs 128-wide chunks): a few fp32 ulps
    assert np.abs(got - two).max() / scale < 1e-5, np.abs(got - two).max()
    mean, rstd = got.astype(np.float64).mean(1), 1.0 / np.sqrt(got.astype(np.float64).var(1) + 1e-5)
    assert np.abs(st[:, 0] - mean).max() < 1e-4 * scale
    assert np.abs(st[:, 1] / rstd - 1.0).max() < 1e-3


@pytest.mark.parametrize("M,C,hid", [(500, 384, 1536), (300, 192, 384), (260, 96, 384)])
def test_swin_mlp_two_gemm_path(M, C, hid):
    """Shapes the fused kernel does not take (C = 384: the A tile + one chunk of weights exceed shared memory; mlp_ratio
    != 4; C = 96) run as folded fc1 + fc2 GEMMs behind the same entry point; fused=1 must refuse them loudly."""
    rng = np.random.default_rng(M + C + hid)
    x = rng.standard_normal((M, C)).astype(np.float32)
    gamma = (1.0 + 0.1 * rng.standard_normal(C)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(C)).astype(np.float32)
    w1 = (rng.standard_normal((hid, C)) / np.sqrt(C)).astype(np.float32)
    b1 = (0.5 * rng.standard_normal(hid)).astype(np.float32)
    w2 = (rng.standard_normal((C, hid)) / np.sqrt(hid)).astype(np.float32)
    b2 = (0.5 * rng.standard_normal(C)).astype(np.float32)
    got, st = cb.ops.swin_mlp(x, gamma, beta, w1, b1, w2, b2, precision="fp16", fused=-1, with_stats=True)
    xd = torch.from_numpy(x).double()
    ln = F.layer_norm(xd, (C,), torch.from_numpy(gamma).double(), torch.from_numpy(beta).double(), 1e-5)
    h = F.gelu(ln @ torch.from_numpy(w1).double().T + torch.from_numpy(b1).double())
    exp = (xd + h @ torch.from_numpy(w2).double().T + torch.from_numpy(b2).double()).numpy()
    scale = max(1.0, np.abs(exp).max())
    assert np.abs(got - exp).max() / scale < 12.0 * 2.0 ** -11
    assert np.abs(st[:, 0] - got.astype(np.float64).mean(1)).max() < 1e-4 * scale
    with pytest.raises(cb._lib.BrnError):
        cb.ops.swin_mlp(x, gamma, beta, w1, b1, w2, b2, precision="fp16", fused=1)


def test_preprocess_matches_image_crate_restatement():
    """examples/infer_image.rs:44-67 on the device: Triangle resize_exact + ImageNet normalise of the reference's own
    test photo, against oracle/imageops_ref.py (restatement of the `image` 0.25.9 sampling code).  Triangle weights are
    plain IEEE f32 arithmetic on both sides and the kernels do not contract multiply-adds: bit-exact."""
    from pathlib import Path
    from oracle import imageops_ref as IM
    rgb = np.load(Path(__file__).parent / "golden" / "cat_768_u8.npz")["rgb"]
    got = cb.ops.preprocess_rgb8(rgb, 1024, 1024)
    exp = IM.preprocess(rgb, 1024)
    assert got.shape == (1, 3, 1024, 1024)
    assert np.array_equal(got, exp), float(np.abs(got - exp).max())
    # down-scaling (support grows with the ratio), non-square, batch of 2, and the same-size copy path
    two = np.stack([rgb[:600, :700], rgb[100:700, 50:750]])
    got = cb.ops.preprocess_rgb8(two, 256, 320)
    for b in range(2):
        assert np.array_equal(got[b], IM.normalize_imagenet(IM.resize(two[b], 256, 320, "triangle"))[0])
    same = cb.ops.preprocess_rgb8(rgb, 768, 768)
    assert np.array_equal(same, IM.normalize_imagenet(rgb))


def test_postprocess_matches_image_crate_restatement():
    """examples/infer_image.rs:85-105: sigmoid -> u8 (truncation) -> Lanczos3 resize.  expf / sinf differ from numpy's in
    the last ulp, which can move a value across a truncation or rounding boundary: <= 1 level, on < 0.1 % of the pixels."""
    from oracle import imageops_ref as IM
    rng = np.random.default_rng(3)
    logits = (rng.standard_normal((2, 256, 320)) * 3).astype(np.float32)
    logits[0, :40] = 30.0       # saturated regions: Lanczos overshoot must clamp
    logits[0, 40:80] = -30.0
    for (oh, ow) in ((256, 320), (768, 768), (100, 517)):
        got = cb.ops.postprocess_mask(logits, oh, ow)
        for b in range(2):
            exp = IM.postprocess(logits[b], oh, ow)
            d = np.abs(got[b].astype(int) - exp.astype(int))
            assert d.max() <= 1 and (d > 0).mean() < 1e-3, (oh, ow, int(d.max()), float((d > 0).mean()))


def test_infer_rgb8_end_to_end(mini_cfg, mini_weights_B):
    """brn_infer_rgb8 == preprocess -> forward_logits -> postprocess chained through the separate entry points."""
    from pathlib import Path
    rgb = np.load(Path(__file__).parent / "golden" / "cat_768_u8.npz")["rgb"][::3, ::3].copy()      # 256 x 256
    two = np.stack([rgb, rgb[::-1].copy()])
    cfg = cb.BiRefNetConfig(swin=cb.SwinConfig(embed_dim=mini_cfg.embed_dim, depths=tuple(mini_cfg.depths),
                                               num_heads=tuple(mini_cfg.num_heads)), precision="fp16")
    m = cb.BiRefNet.new(cfg, mini_weights_B)
    try:
        got = m.infer_rgb8(two, size=(128, 160))
        x = cb.ops.preprocess_rgb8(two, 128, 160)
        exp = cb.ops.postprocess_mask(m.forward_logits(x), 256, 256)
        assert got.shape == (2, 256, 256) and got.dtype == np.uint8
        assert np.array_equal(got, exp)
    finally:
        m.close()


@pytest.mark.parametrize("k,stride,padding", [(3, 1, 1), (3, 2, 1), (3, 2, 0), (5, 1, 2), (1, 1, 0), (7, 3, 2)])
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_deform_conv2d_stride_padding_bias(k, stride, padding, precision):
    """The operator at the geometry the reference's module allows (src/deform_conv.rs:29-46: any kernel / stride /
    padding, bias), against torchvision.  C = 32 and non-"same" geometries take the SIMT kernel in every precision."""
    import torchvision
    rng = np.random.default_rng(k * 100 + stride * 10 + padding)
    B, C, H, W, O = 2, 32, 19, 23, 40
    Ho, Wo = (H + 2 * padding - k) // stride + 1, (W + 2 * padding - k) // stride + 1
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = (rng.standard_normal((O, C, k, k)) / np.sqrt(C * k * k)).astype(np.float32)
    b = rng.standard_normal(O).astype(np.float32)
    off = (rng.standard_normal((B, 2 * k * k, Ho, Wo)) * 1.5).astype(np.float32)
    msk = rng.uniform(0, 2, (B, k * k, Ho, Wo)).astype(np.float32)
    got = cb.ops.deform_conv2d(x, off, msk, w, b, stride=stride, padding=padding, precision=precision)
    xx = r16(x, precision)
    exp = torchvision.ops.deform_conv2d(torch.from_numpy(xx).double(), torch.from_numpy(off).double(),
                                        torch.from_numpy(w).double(), torch.from_numpy(b).double(), stride=stride,
                                        padding=padding, mask=torch.from_numpy(msk).double()).numpy()
    assert got.shape == exp.shape
    assert relerr(got, exp) < 1e-4


@pytest.mark.parametrize("cin,cout,k,stride,padding", [(64, 256, 3, 1, 1), (64, 256, 7, 1, 3), (16, 24, 3, 2, 1), (8, 8, 5, 1, 0)])
@pytest.mark.parametrize("mode", ["deformable", "cpu_fallback"])
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_deformable_conv2d_module(cin, cout, k, stride, padding, mode, precision):
    """DeformableConv2d::new + forward (src/deform_conv.rs:29-99) against the oracle's restatement, both device
    behaviours (Metal path / candle-CPU fallback).  The (64 -> 256, same padding) cases run the tcgen05 kernels."""
    rng = np.random.default_rng(cin + cout + k + stride)
    B, H, W = 2, 24, 32
    vb = {"offset_conv.weight": (rng.standard_normal((2 * k * k, cin, k, k)) * 0.5 / np.sqrt(cin * k * k)),
          "offset_conv.bias": rng.standard_normal(2 * k * k) * 0.5,
          "modulator_conv.weight": rng.standard_normal((k * k, cin, k, k)) / np.sqrt(cin * k * k),
          "modulator_conv.bias": rng.standard_normal(k * k) * 0.1,
          "regular_conv.weight": rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k),
          "regular_conv.bias": rng.standard_normal(cout)}
    vb = {n: v.astype(np.float32) for n, v in vb.items()}
    x = rng.standard_normal((B, cin, H, W)).astype(np.float32)
    m = cb.DeformableConv2d(cin, cout, k, stride, padding, vb, precision=precision, deform_mode=mode)
    got = m(x)
    wt = {"m." + n: torch.from_numpy(v).double() for n, v in vb.items()}
    exp = R.deformable_conv2d(torch.from_numpy(x).double(), wt, "m", k, stride, padding, mode).numpy()
    assert got.shape == exp.shape
    tc = precision == "fp16" and cin == 64 and stride == 1 and padding == k // 2
    # fp16 tensor-core path: offsets come from an fp16-operand conv (sampling positions move by ~1e-3 px) and the
    # sampled values are rounded to fp16; the SIMT path is fp32 arithmetic (on fp16-rounded x when precision = fp16)
    tol = 1e-2 if tc else (2e-3 if precision == "fp16" else 1e-4)
    assert relerr(got, exp) < tol, relerr(got, exp)
    with pytest.raises(cb.BrnError) as e:
        cb.DeformableConv2d(cin, cout, k, stride, padding, {n: v for n, v in vb.items() if n != "regular_conv.bias"})
    assert e.value.status == 3


@pytest.mark.parametrize("nimg,hp,wp,heads,shift", [(1, 14, 21, 3, 0), (2, 14, 21, 3, 3), (1, 7, 7, 6, 3), (1, 70, 70, 3, 3)])
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_window_attention_window7(nimg, hp, wp, heads, shift, precision):
    """The same plain chain for swin_t / swin_s windows (src/swin.rs:27-52: window 7, shift 3; 49 tokens per window),
    mask from the oracle's create_attention_mask (regions split at hp - 7 and hp - 3, src/swin.rs:608-629)."""
    rng = np.random.default_rng(hp + heads + shift)
    n, nw = 49, (hp // 7) * (wp // 7)
    C = heads * 32
    qkv = rng.standard_normal((nimg * nw, n, 3 * C)).astype(np.float32)
    bias = (rng.standard_normal((heads, n, n)) * 0.5).astype(np.float32)
    got = cb.ops.window_attention(qkv, bias, hp, wp, shift, precision=precision)
    t = torch.from_numpy(r16(qkv, precision)).double()
    q, k, v = [t[..., i * C:(i + 1) * C].reshape(nimg * nw, n, heads, 32).permute(0, 2, 1, 3) for i in range(3)]
    s = (q * 32 ** -0.5) @ k.transpose(-1, -2) + torch.from_numpy(bias).double()
    if shift:
        m = R.create_attention_mask(hp, wp, 7, 3, torch.float64)
        s = (s.reshape(nimg, nw, heads, n, n) + m[None, :, None]).reshape(nimg * nw, heads, n, n)
    exp = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(nimg * nw, n, C).numpy()
    err = np.abs(got - exp).max()
    # window 7 runs the SIMT kernel in every precision: fp32 arithmetic on (for fp16) fp16-rounded q, k, v and output
    assert err < {"fp16": 4e-3, "fp32": 2e-5}[precision], err
