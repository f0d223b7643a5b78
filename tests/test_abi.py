"""CPU tests of the boundary: the C-ABI library loads, exports every symbol the header declares, and fails loudly
(no CPU fallback) when there is no GPU.  No compute calls here."""
import ctypes as C

import numpy as np
import pytest

import candle_birefnet_b200 as cb
from candle_birefnet_b200 import _lib
from oracle import birefnet_ref as R
from tests.conftest import has_gpu


def test_library_exports_every_declared_symbol():
    L = cb.lib()
    syms = _lib.declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/birefnet_b200.h but not exported"
    assert b"sm_100a" in L.brn_version()


def test_config_swin_l_matches_reference():
    c = _lib.BrnConfig()
    cb.lib().brn_config_swin_l(C.byref(c))
    # SwinConfig::swin_l (src/swin.rs:69-80)
    assert c.embed_dim == 192 and list(c.depths) == [2, 2, 18, 2] and list(c.num_heads) == [6, 12, 24, 48]
    assert c.window_size == 12 and c.mlp_ratio == 4 and c.patch_size == 4
    py = cb.BiRefNetConfig.swin_l()
    assert py.swin.embed_dim == 192 and py.size == (1024, 1024) and py.mul_scl_ipt


def test_config_swin_b_matches_reference():
    c = _lib.BrnConfig()
    cb.lib().brn_config_swin_b(C.byref(c))
    # SwinConfig::swin_b (src/swin.rs:54-66)
    assert c.embed_dim == 128 and list(c.depths) == [2, 2, 18, 2] and list(c.num_heads) == [4, 8, 16, 32]
    assert c.window_size == 12 and c.mlp_ratio == 4 and c.patch_size == 4
    py = cb.SwinConfig.swin_b()
    assert py.embed_dim == 128 and py.num_heads == (4, 8, 16, 32) and py.window_size == 12
    assert R.Config.swin_b().stage_channels() == [128, 256, 512, 1024]


def test_config_swin_t_s_match_reference():
    c = _lib.BrnConfig()
    cb.lib().brn_config_swin_t(C.byref(c))
    # SwinConfig::swin_t (src/swin.rs:27-38)
    assert c.embed_dim == 96 and list(c.depths) == [2, 2, 6, 2] and list(c.num_heads) == [3, 6, 12, 24] and c.window_size == 7
    cb.lib().brn_config_swin_s(C.byref(c))
    # SwinConfig::swin_s (src/swin.rs:41-52)
    assert c.embed_dim == 96 and list(c.depths) == [2, 2, 18, 2] and list(c.num_heads) == [3, 6, 12, 24] and c.window_size == 7
    assert cb.SwinConfig.swin_t().window_size == 7 and cb.SwinConfig.swin_s().depths == (2, 2, 18, 2)
    assert R.Config.swin_t().stage_channels() == [96, 192, 384, 768]


def test_product_path_has_no_oracle_or_cpu_fallback():
    import inspect
    import candle_birefnet_b200.model as m
    import candle_birefnet_b200.ops as o
    for mod in (m, o, _lib):
        src = inspect.getsource(mod)
        assert "oracle" not in src.replace("no CPU", "")
        assert "import torch.nn" not in src


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    with pytest.raises(cb.BrnError) as e:
        cb.BiRefNet.new(cb.BiRefNetConfig.swin_l(), {})
    assert e.value.status == 2 and "no CPU fallback" in str(e.value)
    with pytest.raises(cb.BrnError):
        cb.ops.linear(np.zeros((4, 8), np.float32), np.zeros((4, 8), np.float32))


def test_null_and_bad_arguments_return_status():
    L = cb.lib()
    assert L.brn_model_create(None, 0, None) != 0
    assert b"null" in L.brn_last_error()
    assert L.brn_model_num_tensors(None) == 0
    assert L.brn_forward_logits(None, None, 1, 32, 32, 0, None, 0, None) != 0
