"""Generates tests/golden/cat_768_u8.npz from the reference's only image asset.

    python tests/golden/make_cat_fixture.py      (in the build container: /root/reference is read-only there and does
                                                  not exist on the GPU box, so the decoded pixels are committed)

The fixture is the decoded RGB8 raster of /root/reference/examples/assets/cat.png (768x768), the input the reference's
CLI feeds through `resize_exact(1024, 1024, Triangle)` + ImageNet normalisation (examples/infer_image.rs:44-67).
SURVEY.md section 8d names it as the realistic input of the mask-IoU check.  Stored as zlib-compressed uint8.
"""
from pathlib import Path

import numpy as np
from PIL import Image

SRC = Path("/root/reference/examples/assets/cat.png")
DST = Path(__file__).resolve().parent / "cat_768_u8.npz"

if __name__ == "__main__":
    im = Image.open(SRC)
    rgb = np.asarray(im.convert("RGB"), dtype=np.uint8)          # image::DynamicImage::to_rgb8
    assert rgb.shape == (768, 768, 3), rgb.shape
    np.savez_compressed(DST, rgb=rgb, mode=np.array(im.mode))
    print(DST, DST.stat().st_size, "bytes; source mode", im.mode)
