"""Builds libbirefnet_b200.so in-tree with nvcc for sm_100a (no torch dependency in the library)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libbirefnet_b200.so"
SOURCES = ["abi.cu", "model.cu", "kernels_simt.cu", "kernels_glue.cu", "gemm_tcgen05.cu", "attn_tcgen05.cu", "mlp_tcgen05.cu",
           "deform_tcgen05.cu", "ln_kernels.cu", "final_kernel.cu", "prepost.cu", "safetensors_loader.cpp", "sharded.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stamp() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*")) + [HERE.parent / "include" / "birefnet_b200.h", Path(__file__)]):
        if p.is_file():
            h.update(p.name.encode())
            h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_variant(name: str, defines: list[str]) -> Path:
    """A/B experiment build: libbirefnet_b200_<name>.so with extra -D flags (select it with BRN_LIB_PATH)."""
    out = HERE / f"libbirefnet_b200_{name}.so"
    bdir = HERE / "build" / name
    bdir.mkdir(parents=True, exist_ok=True)
    procs, objs = [], []
    for s in SOURCES:
        o = bdir / (s + ".o")
        procs.append(subprocess.Popen([_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", str(CSRC / s), "-o", str(o)]))
        objs.append(str(o))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc compilation failed")
    subprocess.run([_nvcc(), "-shared", "-o", str(out), *objs, "-lcudart"], check=True)
    return out


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp_file = HERE / "build" / "stamp"
    stamp = _stamp()
    if not force and LIB.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return LIB
    (HERE / "build").mkdir(exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        o = HERE / "build" / (s + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(CSRC / s), "-o", str(o)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(o))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {s} ---\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"--- {s} ---\n{out}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    cmd = [_nvcc(), "-shared", "-o", str(LIB), *objs, "-lcudart"]
    subprocess.run(cmd, check=True)
    stamp_file.write_text(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
