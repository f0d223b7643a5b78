"""Structured diagnostics for the tcgen05 kernels (prints, never asserts): run on the GPU box."""
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
import candle_birefnet_b200 as cb
from oracle import birefnet_ref as R


def bf16r(a):
    return torch.from_numpy(np.asarray(a, dtype=np.float32)).bfloat16().float().numpy()


def describe(tag, got, exp):
    d = np.abs(got - exp)
    print(f"[{tag}] max|d|={d.max():.4g} mean|d|={d.mean():.4g} max|exp|={np.abs(exp).max():.4g} "
          f"nan={np.isnan(got).sum()} zeros={int((got == 0).sum())}/{got.size}")
    if d.max() > 1e-2 * (np.abs(exp).max() + 1e-9):
        bad = d > 1e-2 * np.abs(exp).max()
        rows, cols = np.where(bad.reshape(bad.shape[0], -1)) if bad.ndim >= 2 else (np.where(bad)[0], None)
        print(f"   bad fraction {bad.mean():.4f}; bad rows%8 hist {np.bincount(rows % 8, minlength=8)}; "
              f"first bad rows {np.unique(rows)[:10]}")
        if cols is not None:
            print(f"   bad cols%64 hist(16-bins) {np.bincount((cols % 64) // 4, minlength=16)}; first bad cols {np.unique(cols)[:10]}")


def main():
    rng = np.random.default_rng(0)
    for (M, N, K) in [(128, 64, 64), (128, 16, 16), (128, 64, 256), (256, 256, 64), (300, 192, 192), (1000, 576, 192)]:
        a = rng.standard_normal((M, K)).astype(np.float32)
        w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
        try:
            got = cb.ops.linear(a, w, None, precision="bf16")
            describe(f"linear {M}x{N}x{K}", got, bf16r(a).astype(np.float64) @ bf16r(w).astype(np.float64).T)
        except Exception as e:
            print("linear failed", (M, N, K), e)
    for (C, O, k, H, W) in [(64, 64, 1, 16, 8), (64, 64, 3, 16, 8), (64, 64, 3, 32, 32), (128, 32, 3, 8, 8)]:
        x = rng.standard_normal((1, C, H, W)).astype(np.float32)
        w = (rng.standard_normal((O, C, k, k)) / np.sqrt(C * k * k)).astype(np.float32)
        try:
            got = cb.ops.conv2d(x, w, None, precision="bf16")
            exp = torch.nn.functional.conv2d(torch.from_numpy(bf16r(x)).double(), torch.from_numpy(bf16r(w)).double(), padding=k // 2).numpy()
            describe(f"conv C{C} O{O} k{k} {H}x{W}", got.reshape(O, -1), exp.reshape(O, -1))
        except Exception as e:
            print("conv failed", e)
    for (hp, wp, heads, shift) in [(12, 12, 1, 0), (12, 12, 2, 0), (24, 24, 2, 6)]:
        nw = (hp // 12) * (wp // 12)
        C = heads * 32
        qkv = rng.standard_normal((nw, 144, 3 * C)).astype(np.float32)
        bias = (rng.standard_normal((heads, 144, 144)) * 0.5).astype(np.float32)
        t = torch.from_numpy(bf16r(qkv)).double()
        q, k, v = [t[..., i * C:(i + 1) * C].reshape(nw, 144, heads, 32).permute(0, 2, 1, 3) for i in range(3)]
        s = (q * 32 ** -0.5) @ k.transpose(-1, -2) + torch.from_numpy(bias).double()
        if shift:
            m = R.create_attention_mask(hp, wp, 12, 6, torch.float64)
            s = s + m[:, None]
        exp = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(nw, 144, C).numpy()
        for prec in ("fp32", "bf16"):
            try:
                got = cb.ops.window_attention(qkv, bias, hp, wp, shift, precision=prec)
                describe(f"attn {prec} hp{hp} heads{heads} shift{shift}", got.reshape(nw * 144, C), exp.reshape(nw * 144, C))
            except Exception as e:
                print("attn failed", prec, e)


if __name__ == "__main__":
    main()
