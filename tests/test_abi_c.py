"""The C ABI used from C: tests/c/abi_smoke.c is compiled with gcc against include/birefnet_b200.h and linked with the
in-tree libbirefnet_b200.so (no Python, no ctypes in the call path).  CPU: it builds warning-free and the probe passes
(loud failure without a GPU).  GPU: create -> load_safetensors -> finalize -> forward_logits / forward from C gives
the same bits as the ctypes path.  Plus: every prototype of the header agrees with the ctypes declaration in
candle_birefnet_b200/_lib.py (argument count and kind), so a header / binding drift cannot pass silently."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import candle_birefnet_b200 as cb
from candle_birefnet_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "candle_birefnet_b200"


@pytest.fixture(scope="module")
def smoke_bin(tmp_path_factory):
    cb.lib()                                            # the library exists (built in-tree)
    out = tmp_path_factory.mktemp("abi") / "abi_smoke"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O1", "-I", str(ROOT / "include"), str(ROOT / "tests/c/abi_smoke.c"),
           "-o", str(out), "-L", str(PKG), "-l:libbirefnet_b200.so", f"-Wl,-rpath,{PKG}", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_c_program_builds_and_probe_passes(smoke_bin):
    r = subprocess.run([str(smoke_bin), "probe"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "probe ok" in r.stdout


@pytest.mark.gpu
def test_c_program_matches_ctypes_path(smoke_bin, mini_cfg, mini_weights_B, tmp_path):
    from safetensors.numpy import save_file
    from oracle.make_weights import make_input
    wf = tmp_path / "mini.safetensors"
    save_file({k: np.ascontiguousarray(v) for k, v in mini_weights_B.items()}, str(wf))
    x = make_input(2, 96, 128, seed=3)
    x.tofile(tmp_path / "in.f32")
    args = [str(smoke_bin), "run", str(wf), str(tmp_path / "in.f32"), "2", "96", "128", str(mini_cfg.embed_dim),
            *[str(h) for h in mini_cfg.num_heads], str(tmp_path / "out.f32")]
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(tmp_path / "out.f32", dtype=np.float32).reshape(2, 1, 96, 128)
    cfg = cb.BiRefNetConfig(swin=cb.SwinConfig(embed_dim=mini_cfg.embed_dim, depths=tuple(mini_cfg.depths),
                                               num_heads=tuple(mini_cfg.num_heads)), precision="fp16", deform_mode="deformable")
    m = cb.BiRefNet.new(cfg, mini_weights_B)
    want = m.forward_logits(x)
    m.close()
    assert np.array_equal(got, want)


# ---- header prototypes vs ctypes declarations ----------------------------------------------------------------
def _kind_of_c(decl: str) -> str:
    d = decl.strip()
    if "*" in d or "[" in d:
        return "ptr"
    t = d.rsplit(" ", 1)[0].replace("const ", "").strip() if " " in d else d
    return {"int": "i32", "int32_t": "i32", "int64_t": "i64", "size_t": "size", "float": "f32"}.get(t, "?" + t)


def _kind_of_ctypes(t) -> str:
    if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
        return "ptr"
    return {C.c_int: "i32", C.c_int32: "i32", C.c_int64: "i64", C.c_size_t: "size", C.c_float: "f32"}.get(t, "?" + str(t))


def test_ctypes_declarations_match_header():
    text = re.sub(r"/\*.*?\*/", "", _lib.HEADER.read_text(), flags=re.S)
    protos = re.findall(r"BRN_API\s+([\w\s\*]+?)\b(brn_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    assert len(protos) == len(_lib.declared_symbols())
    L = cb.lib()
    for ret, name, params in protos:
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [p for p in params.split(",")]
        fn = getattr(L, name)
        if fn.argtypes is None:
            assert not plist, f"{name}: header declares {len(plist)} parameters, _lib.py declares none"
            continue
        assert len(fn.argtypes) == len(plist), f"{name}: header has {len(plist)} parameters, ctypes {len(fn.argtypes)}"
        for i, (cdecl, ct) in enumerate(zip(plist, fn.argtypes)):
            assert _kind_of_c(cdecl) == _kind_of_ctypes(ct), f"{name} parameter {i}: header `{cdecl.strip()}` vs ctypes {ct}"
