"""Image sharding across GPUs: one process per GPU, contiguous batch split, weights replicated,
no collective on the data path -- every image's forward is independent (BatchNorm is eval-mode everywhere,
src/decoder.rs:129,139; src/aspp.rs:220,316,330).  `torch.distributed` is used only as plumbing: rank/world
discovery, barriers for timing, and (optionally) collecting the masks on rank 0.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Tuple

import numpy as np


def shard_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous split of n images over `world` ranks; the first n % world ranks get one extra image."""
    base, extra = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def forward_sharded(forward: Callable[[np.ndarray], np.ndarray], x: np.ndarray, rank: int, world: int,
                    gather: bool = False, group=None) -> Optional[np.ndarray]:
    """Runs `forward` (e.g. BiRefNet.forward_logits) on this rank's contiguous slice of the global batch `x`.

    gather=False: returns the local masks [b_local,1,H,W] (the bench path: nothing crosses ranks).
    gather=True : rank 0 returns the full [B,1,H,W] in input order, other ranks return None; the exchange is a
                  host-side gather of the 1-channel masks (4 MiB per 1024^2 image), not part of the timed hot path.
    """
    lo, hi = shard_bounds(x.shape[0], world)[rank]
    local = forward(x[lo:hi]) if hi > lo else np.zeros((0, 1) + x.shape[2:], np.float32)
    if not gather:
        return local
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    parts = [None] * world if rank == 0 else None
    dist.gather_object(local, parts, dst=0, group=group)
    if rank != 0:
        return None
    return np.concatenate(parts, axis=0)


class ShardedBiRefNet:
    """One process, several GPUs: `brn_sharded_*` of the C ABI (one model handle + one host thread per GPU, contiguous
    image split, weights replicated, no collective).  Mirrors BiRefNet.new / forward_logits / forward on host arrays."""

    def __init__(self, config, vb, devices):
        import ctypes as C
        from . import _lib
        from .model import BiRefNet, _DEF, _PREC
        L = _lib.lib()
        c = _lib.BrnConfig()
        c.embed_dim = config.swin.embed_dim
        for i in range(4):
            c.depths[i] = config.swin.depths[i]
            c.num_heads[i] = config.swin.num_heads[i]
        c.window_size, c.mlp_ratio, c.patch_size = config.swin.window_size, config.swin.mlp_ratio, config.swin.patch_size
        c.precision, c.deform_mode, c.micro_batch = _PREC[config.precision], _DEF[config.deform_mode], config.micro_batch
        devs = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        _lib.check(L.brn_sharded_create(C.byref(c), devs, len(devices), C.byref(h)))
        self._h, self.config, self.devices = h, config, list(devices)
        try:
            if isinstance(vb, str):
                n = C.c_int32(0)
                _lib.check(L.brn_sharded_load_safetensors(h, vb.encode(), C.byref(n)))
            else:
                probe = BiRefNet._create(config, devices[0])
                try:
                    keys = probe.tensor_keys()
                finally:
                    probe.close()
                for key in keys:
                    if key not in vb:
                        raise _lib.BrnError(3, f"cannot find tensor {key}")
                    a = np.ascontiguousarray(vb[key], dtype=np.float32)
                    shape = (C.c_int64 * a.ndim)(*a.shape)
                    _lib.check(L.brn_sharded_set_tensor(h, key.encode(), a.ctypes.data_as(C.c_void_p), _lib.F32, shape, a.ndim))
            _lib.check(L.brn_sharded_finalize(h))
        except Exception:
            L.brn_sharded_destroy(h)
            self._h = None
            raise

    def _run(self, fn, x, out=None):
        import ctypes as C
        from . import _lib
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 4 or x.shape[1] != 3:
            raise _lib.BrnError(5, f"expected [B,3,H,W], got {x.shape}")
        B, _, H, W = x.shape
        if out is None:
            out = np.empty((B, 1, H, W), dtype=np.float32)
        _lib.check(fn(self._h, x.ctypes.data_as(C.c_void_p), B, H, W, out.ctypes.data_as(C.c_void_p)))
        return out

    def forward_logits(self, x, out=None):
        from . import _lib
        return self._run(_lib.lib().brn_sharded_forward_logits, x, out)

    def forward(self, x, out=None):
        from . import _lib
        return self._run(_lib.lib().brn_sharded_forward, x, out)

    __call__ = forward

    def close(self):
        from . import _lib
        if self._h:
            _lib.lib().brn_sharded_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
