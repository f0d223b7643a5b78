"""rust/ cannot be compiled here (no cargo); these checks keep the `-sys` crate's source in step with the header so the
binding a maintainer builds elsewhere is the ABI this repository tests: same functions, same parameter counts, same
struct layout and constants."""
import re
from pathlib import Path

from candle_birefnet_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
SYS = (ROOT / "rust" / "birefnet-b200-sys" / "src" / "lib.rs").read_text()
HDR = re.sub(r"/\*.*?\*/", "", _lib.HEADER.read_text(), flags=re.S)


def _split_params(s: str):
    s = " ".join(s.split())
    return [] if s in ("", "void") else s.split(",")


def test_extern_block_declares_exactly_the_header_functions():
    hdr = {name: len(_split_params(params))
           for _, name, params in re.findall(r"BRN_API\s+([\w\s\*]+?)\b(brn_\w+)\s*\(([^;]*?)\)\s*;", HDR, flags=re.S)}
    rust = {name: len([p for p in _split_params(params) if p.strip()])
            for name, params in re.findall(r"pub fn (brn_\w+)\s*\(([^;]*?)\)\s*(?:->[^;]+)?;", SYS, flags=re.S)}
    assert set(rust) == set(hdr), (sorted(set(hdr) - set(rust)), sorted(set(rust) - set(hdr)))
    for name, n in hdr.items():
        assert rust[name] == n, f"{name}: header has {n} parameters, rust/birefnet-b200-sys has {rust[name]}"


def test_config_struct_and_constants_match_header():
    fields_h = re.findall(r"int32_t\s+(\w+)(?:\[(\d)\])?;", re.search(r"typedef struct \{(.*?)\} brn_config;", HDR, flags=re.S).group(1))
    fields_r = re.findall(r"pub (\w+): (?:\[i32; (\d)\]|i32),", re.search(r"pub struct brn_config \{(.*?)\}", SYS, flags=re.S).group(1))
    assert fields_h == fields_r
    for enum_body in re.findall(r"typedef enum \{(.*?)\}", HDR, flags=re.S):
        for name, val in re.findall(r"(BRN_\w+)\s*=\s*(\d+)", enum_body):
            m = re.search(rf"pub const {name}: [\w:]+ = (\d+);", SYS)
            assert m and m.group(1) == val, name
    assert len(_lib.BrnConfig._fields_) == len(fields_h)


def test_wrapper_crate_keeps_the_reference_signatures():
    src = (ROOT / "rust" / "candle-birefnet-b200" / "src" / "lib.rs").read_text()
    for sig in ("pub fn swin_l() -> Self", "pub fn new(config: BiRefNetConfig, vb: VarBuilder) -> Result<Self>",
                "pub fn forward_logits(&self, x: &Tensor) -> Result<Tensor>", "pub fn forward(&self, x: &Tensor) -> Result<Tensor>",
                "impl Module for BiRefNet", "impl Module for DeformableConv2d"):
        assert sig in src, sig
    assert "pub fn new(in_channels: usize, out_channels: usize, kernel_size: usize, stride: usize, padding: usize" in src
