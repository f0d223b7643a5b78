"""Round-2 experiment: parity numbers of the benchmarked configuration, per precision policy (prints JSON lines)."""
import json, os, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import candle_birefnet_b200 as cb
from oracle import birefnet_ref as R
from oracle import imageops_ref as IM
from oracle.make_weights import as_torch, make_input, make_weights

sig = lambda z: 1.0 / (1.0 + np.exp(-z))
def iou(a, b):
    a, b = a > 0.5, b > 0.5
    return float(np.logical_and(a, b).sum() / max(1, np.logical_or(a, b).sum()))
def pc(cfg, p, m):
    return cb.BiRefNetConfig(swin=cb.SwinConfig(embed_dim=cfg.embed_dim, depths=tuple(cfg.depths), num_heads=tuple(cfg.num_heads)), precision=p, deform_mode=m)
out = open(ROOT / "gpurun_out" / "exp_parity.jsonl", "a")
def emit(**kw):
    s = json.dumps(kw); print(s, flush=True); out.write(s + "\n"); out.flush()

# 1. mini backbone + features errors
cfg = R.Config.mini()
wA = make_weights(cfg, seed=0, weight_set="A"); wB = make_weights(cfg, seed=0, weight_set="B", offset_sigma=2.0)
m = cb.BiRefNet.new(pc(cfg, "fp32", "deformable"), wA)
for hw in [(128, 160), (256, 256)]:
    x = make_input(2, hw[0], hw[1], seed=5)
    exp = R.swin_forward(torch.from_numpy(x), as_torch(wA), cfg)
    for p in ("fp32", "bf16", "fp16"):
        m.set_precision(p); got = m.backbone_forward(x)
        for i in range(4):
            e = exp[i].numpy()
            emit(what="backbone_mini", hw=hw, prec=p, i=i, rel=float(np.abs(got[i]-e).max()/np.abs(e).max()),
                 rms=float(np.sqrt(np.mean((got[i]-e)**2))/np.sqrt(np.mean(e**2))))
m.close()
m = cb.BiRefNet.new(pc(cfg, "fp32", "deformable"), wB)
for hw in [(32, 32), (96, 160), (64, 96), (224, 96)]:
    x = make_input(2, hw[0], hw[1], seed=17)
    exp = R.features(torch.from_numpy(x), as_torch(wB), cfg)
    for p in ("fp32", "bf16", "fp16"):
        m.set_precision(p); got = m.features_forward(x)
        for i in range(4):
            e = exp[i].numpy()
            emit(what="features_mini", hw=hw, prec=p, i=i, shape=list(got[i].shape), rel=float(np.abs(got[i]-e).max()/max(1.0, np.abs(e).max())))
m.close()

# 2. Swin-L 1024^2
cfg = R.Config.swin_l()
rgb = np.load(ROOT / "tests/golden/cat_768_u8.npz")["rgb"]
inputs = {"randn": make_input(1, 1024, 1024, seed=1234), "cat": IM.preprocess(rgb, 1024)}
for wset, mode in (("A", "cpu_fallback"), ("B", "deformable")):
    w = make_weights(cfg, seed=0, weight_set=wset, offset_sigma=2.0)
    m = cb.BiRefNet.new(pc(cfg, "fp16", mode), w)
    wt = as_torch(w)
    for kind, x in inputs.items():
        t0 = time.time()
        with torch.no_grad():
            exp = R.forward_logits(torch.from_numpy(x), wt, cfg, mode).numpy()
        t_or = time.time() - t0
        near = float((np.abs(exp) < 5e-3).mean())
        for p, dec in (("fp32", ""), ("fp16", ""), ("bf16", "fp16"), ("bf16", "bf16")):
            m.set_precision(p)
            if dec: os.environ["BRN_BF16_DECODER"] = dec
            else: os.environ.pop("BRN_BF16_DECODER", None)
            got = m.forward_logits(x)
            emit(what="swin_l_1024", wset=wset, mode=mode, input=kind, prec=p, dec=dec, oracle_s=round(t_or, 1),
                 max_dlogit=float(np.abs(got-exp).max()), mean_dlogit=float(np.abs(got-exp).mean()),
                 max_dsig=float(np.abs(sig(got)-sig(exp)).max()), iou=iou(sig(got), sig(exp)),
                 logit_std=float(exp.std()), frac_within_5e3=near)
        os.environ.pop("BRN_BF16_DECODER", None)
    m.close()
