"""Per-stage error report of the CUDA path vs the oracle on the mini config (prints, never asserts)."""
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
import candle_birefnet_b200 as cb
from oracle import birefnet_ref as R
from oracle.make_weights import as_torch, make_input, make_weights


def main():
    cfg = R.Config.mini()
    for wset in ("A", "B"):
        wnp = make_weights(cfg, seed=0, weight_set=wset)
        w = as_torch(wnp)
        pc = cb.BiRefNetConfig(swin=cb.SwinConfig(embed_dim=cfg.embed_dim, depths=tuple(cfg.depths), num_heads=tuple(cfg.num_heads)),
                               precision="fp32", deform_mode="deformable")
        m = cb.BiRefNet.new(pc, wnp)
        x = make_input(1, 128, 160, seed=5)
        xt = torch.from_numpy(x)
        feats = R.swin_forward(xt, w, cfg)
        x1, x2, x3, x4 = R.features(xt, w, cfg)
        for prec in ("fp32", "bf16"):
            m.set_precision(prec)
            try:
                got = m.backbone_forward(x)
                print(f"[{wset} {prec}] backbone max|d| per stage:", [float(np.abs(g - e.numpy()).max()) for g, e in zip(got, feats)],
                      "max|exp|", [float(e.abs().max()) for e in feats])
            except Exception as e:
                print(f"[{wset} {prec}] backbone failed:", e)
            for mode in ("cpu_fallback", "deformable"):
                m.set_deform_mode(mode)
                try:
                    sq = R.basic_dec_blk(x4, w, "squeeze_module.0", mode)
                    exp = R.decoder_forward(xt, x1, x2, x3, sq, w, mode).numpy()
                    got = m.decoder_forward(x, x1.numpy(), x2.numpy(), x3.numpy(), x4.numpy())
                    print(f"[{wset} {prec} {mode}] decoder max|dlogit| {np.abs(got - exp).max():.5g} (std {exp.std():.3g})")
                    got = m.forward_logits(x)
                    print(f"[{wset} {prec} {mode}] forward max|dlogit| {np.abs(got - exp).max():.5g}  launches {m.launch_count()}")
                    m.reset_launch_count()
                except Exception as e:
                    print(f"[{wset} {prec} {mode}] failed:", e)
        m.close()


if __name__ == "__main__":
    main()
