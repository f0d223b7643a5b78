"""Generates tests/golden/*.npz from the oracle (run here, in the build container; the fixtures travel to the GPU box).

    python -m tests.golden.make_golden
"""
from pathlib import Path

import numpy as np
import torch

from oracle import birefnet_ref as R
from oracle.make_weights import as_torch, make_input, make_weights

OUT = Path(__file__).parent


def main():
    torch.set_num_threads(8)
    cfg = R.Config.mini()
    w = as_torch(make_weights(cfg, seed=0, weight_set="B", offset_sigma=2.0))
    x = torch.from_numpy(make_input(1, 64, 96, seed=7))
    d = {}
    for mode in ("cpu_fallback", "deformable"):
        d["logits_" + mode] = R.forward_logits(x, w, cfg, mode).numpy()
    for i, ft in enumerate(R.swin_forward(x, w, cfg)):
        d[f"feat{i}"] = ft.numpy()
    np.savez_compressed(OUT / "mini_64x96.npz", **d)
    print({k: v.shape for k, v in d.items()})


if __name__ == "__main__":
    main()
