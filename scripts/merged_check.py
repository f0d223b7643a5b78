"""Merged two-grid backbone pass == two separate passes, bit for bit (run with and without BRN_SPLIT_BACKBONE=1)."""
import os, sys, hashlib
sys.path.insert(0, ".")
import numpy as np
import candle_birefnet_b200 as cb
from candle_birefnet_b200.synth import synthetic_input
cfg = cb.BiRefNetConfig(swin=cb.SwinConfig.swin_l(), precision="fp16", deform_mode="deformable")
m = cb.BiRefNet.new_synthetic(cfg, seed=0, weight_set="B", offset_sigma=2.0, device=0)
x = synthetic_input(2, 512, 384, seed=3)
y = m.forward_logits(x)
print("split" if os.environ.get("BRN_SPLIT_BACKBONE") else "merged", hashlib.sha256(y.tobytes()).hexdigest(), float(y.mean()), float(np.abs(y).max()))
