"""Single-process image sharding through the C ABI (`brn_sharded_*`): one handle + one host thread
per GPU, contiguous image split, no collective.  The contract: sharded == unsharded, bit for bit, for every batch size
including ragged splits and batches smaller than the device list.  On a 1-GPU box the device list names cuda:0
several times -- the threads, the split and the per-handle workspaces are the same code path; with more GPUs visible
the same test also spans devices."""
import numpy as np
import pytest
import torch

import candle_birefnet_b200 as cb
from candle_birefnet_b200.shard import ShardedBiRefNet, shard_bounds
from oracle.make_weights import make_input

pytestmark = pytest.mark.gpu


def py_cfg(cfg, precision="fp16"):
    return cb.BiRefNetConfig(swin=cb.SwinConfig(embed_dim=cfg.embed_dim, depths=tuple(cfg.depths),
                                               num_heads=tuple(cfg.num_heads)), precision=precision, deform_mode="deformable")


@pytest.mark.parametrize("nshards", [2, 3])
def test_sharded_equals_single_handle(mini_cfg, mini_weights_B, nshards):
    ngpu = torch.cuda.device_count()
    devices = [i % ngpu for i in range(nshards)]
    single = cb.BiRefNet.new(py_cfg(mini_cfg), mini_weights_B)
    sharded = ShardedBiRefNet(py_cfg(mini_cfg), mini_weights_B, devices)
    try:
        for B in (1, 2, 5, 8):                      # B < shards, ragged and even splits
            x = make_input(B, 96, 128, seed=50 + B)
            want = single.forward_logits(x)
            got = sharded.forward_logits(x)
            assert np.array_equal(got, want), (B, devices)
            assert np.array_equal(sharded.forward(x), single.forward(x))
        # the split rule is the one of the process-per-GPU path
        assert [hi - lo for lo, hi in shard_bounds(5, nshards)] == ([3, 2] if nshards == 2 else [2, 2, 1])
    finally:
        single.close()
        sharded.close()
    prev = torch.cuda.current_device()
    assert prev == 0                                # the library leaves the caller's current device alone


def test_sharded_errors(mini_cfg, mini_weights_B):
    with pytest.raises(cb.BrnError) as e:
        ShardedBiRefNet(py_cfg(mini_cfg), mini_weights_B, [0, 99])
    assert e.value.status == 1                      # device index out of range
    w = dict(mini_weights_B)
    w.pop("bb.norm2.weight")
    with pytest.raises(cb.BrnError) as e:
        ShardedBiRefNet(py_cfg(mini_cfg), w, [0, 0])
    assert e.value.status == 3
