"""Per-kernel micro-benchmarks on the shapes of the Swin-L 1024^2 batch-16 step (device-resident synthetic data)."""
import sys
sys.path.insert(0, ".")
from candle_birefnet_b200 import ops

PEAK = 1388.4


def gemm(tag, M, N, K, k=1, act=0, res=False, f32=False, H=None, W=None, B=1, prec="fp16"):
    if H is None:
        B, H, W = 1, 1, M
    ms = ops.bench_op("gemm", B, H, W, K, N, k, act, int(res), f32, precision=prec)
    fl = 2.0 * B * H * W * N * K * k * k
    print(f"gemm {tag:28s} M={B*H*W:8d} N={N:5d} K={k*k}x{K:5d} act={act} res={int(res)}: {ms*1e3:9.1f} us  {fl/ms/1e9:7.1f} TF/s ({fl/ms/1e9/PEAK*100:4.1f}%)", flush=True)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "one":   # one <M> <N> <K> <act> <res> <f32>
        M, N, K, act, res, f32 = [int(v) for v in sys.argv[2:8]]
        gemm("one", M, N, K, act=act, res=bool(res), f32=bool(f32))
        return
    if which in ("all", "gemm"):
        gemm("s2 fc1 gelu", 65536, 3072, 768, act=2)
        gemm("s2 qkv", 82944, 2304, 768)
        gemm("s2 fc2 res", 65536, 768, 3072, res=True, f32=True)
        gemm("s2 proj res", 82944, 768, 768, res=True, f32=True)
        gemm("s0 fc1 gelu", 1048576, 768, 192, act=2)
        gemm("s0 qkv", 1115136, 576, 192)
        gemm("s0 fc2 res", 1048576, 192, 768, res=True, f32=True)
        gemm("s1 fc1 gelu", 262144, 1536, 384, act=2)
        gemm("s3 fc1 gelu", 16384, 6144, 1536, act=2)
        gemm("s3 fc2 res", 16384, 1536, 6144, res=True, f32=True)
        gemm("k1 offmod N=3", 0, 3, 64, k=1, act=3, f32=True, B=16, H=256, W=256)
        gemm("k3 offmod N=27", 0, 27, 64, k=3, act=3, f32=True, B=16, H=256, W=256)
        gemm("k7 offmod N=147", 0, 147, 64, k=7, act=3, f32=True, B=16, H=256, W=256)
        gemm("dec1 conv_in 480->64", 0, 64, 480, k=3, act=1, B=16, H=256, W=256)
        gemm("dec1 conv1 1024->64", 0, 64, 1024, k=1, act=1, B=16, H=256, W=256)
        gemm("dec1 conv_out 64->192", 0, 192, 64, k=3, f32=True, B=16, H=256, W=256)
        gemm("lat2 384->384 res", 0, 384, 384, k=1, res=True, B=16, H=256, W=256)
    if which == "s2":     # the four GEMMs of a stage-2 block + the stage-1 / stage-3 MLP pair
        gemm("s2 fc1 gelu", 65536, 3072, 768, act=2)
        gemm("s2 qkv", 82944, 2304, 768)
        gemm("s2 fc2 res", 65536, 768, 3072, res=True, f32=True)
        gemm("s2 proj res", 82944, 768, 768, res=True, f32=True)
        gemm("s1 fc1 gelu", 262144, 1536, 384, act=2)
        gemm("s1 fc2 res", 262144, 384, 1536, res=True, f32=True)
        gemm("s3 fc1 gelu", 16384, 6144, 1536, act=2)
        gemm("s3 fc2 res", 16384, 1536, 6144, res=True, f32=True)
    if which == "tg":
        gemm("dec1 conv_in 480->64", 0, 64, 480, k=3, act=1, B=16, H=256, W=256)
        gemm("dec2 conv_in 960->64", 0, 64, 960, k=3, act=1, B=16, H=128, W=128)
        gemm("dec3 conv_in 1920->64", 0, 64, 1920, k=3, act=1, B=16, H=64, W=64)
        gemm("sq conv_in 5760->64", 0, 64, 5760, k=3, act=1, B=16, H=32, W=32)
        gemm("ipt 48->64", 0, 64, 48, k=3, B=16, H=256, W=256)
        gemm("gdt 384->16", 0, 16, 384, k=3, act=1, B=16, H=128, W=128)
        gemm("k3 offmod N=27", 0, 27, 64, k=3, act=3, f32=True, B=16, H=256, W=256)
        gemm("conv_out 64->192", 0, 192, 64, k=3, f32=True, B=16, H=256, W=256)
    if which == "res":
        gemm("s0 proj res", 1115136, 192, 192, res=True, f32=True)
        gemm("s0 fc2 res", 1048576, 192, 768, res=True, f32=True)
        gemm("s1 proj res", 278784, 384, 384, res=True, f32=True)
        gemm("s1 fc2 res", 262144, 384, 1536, res=True, f32=True)
        gemm("s2 proj res", 82944, 768, 768, res=True, f32=True)
        gemm("s2 fc2 res", 65536, 768, 3072, res=True, f32=True)
    if which == "lat":
        gemm("lat2 res16", 0, 384, 384, k=1, res=2, B=16, H=256, W=256)
        gemm("lat2 nores", 0, 384, 384, k=1, B=16, H=256, W=256)
        gemm("lat2 res32 out32", 0, 384, 384, k=1, res=True, f32=True, B=16, H=256, W=256)
        gemm("lat2 linear res16", 1048576, 384, 384, res=2)
        gemm("lat2 linear nores", 1048576, 384, 384)
        gemm("lat2 N=192 nores", 0, 192, 384, k=1, B=16, H=256, W=256)
        gemm("lat3 768 res16", 0, 768, 768, k=1, res=2, B=16, H=128, W=128)
        gemm("lat3 768 nores", 0, 768, 768, k=1, B=16, H=128, W=128)
    if which == "small":
        for N in (3, 16, 32, 64):
            for act in (0, 3):
                for f32 in (True, False):
                    gemm(f"N={N} act={act} f32={int(f32)}", 0, N, 64, k=1, act=act, f32=f32, B=16, H=256, W=256)
        gemm("patch embed K=48", 1048576, 192, 48, f32=True)
    if which in ("all", "attn"):
        for (nw, heads, side) in ((7744, 6, 22), (1936, 12, 11), (576, 24, 6), (144, 48, 3)):
            for shift in (0, 6):
                ms = ops.bench_op("attn", nw, side, side, heads, 0, shift, precision="fp16")
                fl = 4.0 * 144 * 144 * 32 * nw * heads
                print(f"attn windows={nw:5d} heads={heads:2d} shift={shift}: {ms*1e3:9.1f} us  {fl/ms/1e9:7.1f} TF/s  {nw*heads/ms/1e3:8.1f} units/us", flush=True)
    if which == "attnm":   # the model's attention launches (token-order output, pad fix-up): full-resolution grids at batch 16
        for (nw, heads, side, pad) in ((7744, 6, 22, 8), (1936, 12, 11, 4), (576, 24, 6, 8), (144, 48, 3, 4), (64, 48, 2, 8)):
            for shift in (0, 6):
                ms0 = ops.bench_op("attn", nw, side, side, heads, 0, shift, precision="fp16")
                ms1 = ops.bench_op("attn", nw, side, side, heads, pad, shift, with_res=True, precision="fp16")
                print(f"attn windows={nw:5d} heads={heads:2d} shift={shift} pad={pad}: window-order {ms0*1e3:8.1f} us {nw*heads/ms0/1e3:6.1f} units/us | "
                      f"token-order+pads {ms1*1e3:8.1f} us {nw*heads/ms1/1e3:6.1f} units/us", flush=True)
        return
    if which == "mlp":     # the MLP half of a Swin block at the model's stage-0 / stage-1 row counts (batch 16, merged grids)
        for (M, Cc) in ((1310720, 192), (327680, 128)):
            for prec in ("fp16", "bf16"):
                ms0 = ops.bench_op("mlp", 1, 1, M, Cc, with_res=False, precision=prec)
                ms1 = ops.bench_op("mlp", 1, 1, M, Cc, with_res=True, precision=prec)
                fl = 16.0 * M * Cc * Cc
                gb = 12.0 * M * Cc
                print(f"mlp M={M} C={Cc} {prec}: fc1+fc2 {ms0*1e3:8.1f} us {fl/ms0/1e9:6.1f} TF/s | fused {ms1*1e3:8.1f} us "
                      f"{fl/ms1/1e9:6.1f} TF/s {gb/ms1/1e6:6.0f} GB/s compulsory", flush=True)
        return
    if which == "mlp2":    # fc1 (LayerNorm fold + GELU) and fc2 (fp32 residual + emit) as the model launches them, stage 2 / 3 rows
        for (M, Cc) in ((81920, 768), (20480, 1536), (327680, 384)):
            ms0 = ops.bench_op("mlp", 1, 1, M, Cc, with_res=False, precision="fp16")
            fl = 16.0 * M * Cc * Cc
            print(f"mlp2 M={M} C={Cc}: fc1+fc2 {ms0*1e3:8.1f} us {fl/ms0/1e9:6.1f} TF/s", flush=True)
        return
    if which == "attn1":   # attn1 <windows> <heads> <side> <shift>
        nw, heads, side, shift = [int(v) for v in sys.argv[2:6]]
        ms = ops.bench_op("attn", nw, side, side, heads, 0, shift, precision="fp16")
        fl = 4.0 * 144 * 144 * 32 * nw * heads
        print(f"attn windows={nw:5d} heads={heads:2d} shift={shift}: {ms*1e3:9.1f} us  {fl/ms/1e9:7.1f} TF/s  {nw*heads/ms/1e3:8.1f} units/us", flush=True)
        return
    if which == "deform1":   # deform1 <side> <k>
        side, k = int(sys.argv[2]), int(sys.argv[3])
        ms = ops.bench_op("deform", 16, side, side, 64, 256, k, act=1, precision="fp16")
        fl = 2.0 * 16 * side * side * 256 * k * k * 64
        print(f"deform {side}^2 k={k}: {ms*1e3:9.1f} us  {fl/ms/1e9:7.1f} TF/s", flush=True)
        return
    if which in ("all", "deform"):
        for (side, k) in ((256, 7), (256, 3), (256, 1), (128, 7), (64, 7)):
            ms = ops.bench_op("deform", 16, side, side, 64, 256, k, act=1, precision="fp16")
            fl = 2.0 * 16 * side * side * 256 * k * k * 64
            print(f"deform {side}^2 k={k}: {ms*1e3:9.1f} us  {fl/ms/1e9:7.1f} TF/s", flush=True)


if __name__ == "__main__":
    main()
