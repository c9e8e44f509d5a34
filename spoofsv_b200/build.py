"""In-tree nvcc build of libspoofsv_b200.so (sm_100a only; no torch in the link).

    python -m spoofsv_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libspoofsv_b200.so"
SOURCES = ["abi.cu", "backward.cu", "conv_f32.cu", "conv_tc.cu", "conv_tc2.cu", "conv_tc32.cu", "decode_ws.cu", "griffinlim.cu", "misc.cu", "train_small.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the CUDA library cannot be built")
    return exe


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "spoofsv_b200.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False, defines=(), variant: str | None = None) -> Path:
    """Compile the library.  `variant` (development aid for A/B kernel experiments): build with the extra `-D` flags in
    `defines` into build/variants/<variant>/libspoofsv_b200.so instead of the product library; load it by pointing
    SSV_B200_LIB at it (spoofsv_b200/_lib.py)."""
    lib = LIB
    out_dir = PKG / "build"
    if variant:
        out_dir = PKG / "build" / "variants" / variant
        lib = out_dir / "libspoofsv_b200.so"
        force = True
    elif not force and not _stale():
        return LIB
    objs = []
    out_dir.mkdir(parents=True, exist_ok=True)
    common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", *ARCH, *[f"-D{d}" for d in defines]]
    if verbose:
        common += ["-Xptxas", "-v"]
    procs = []
    for src in SOURCES:
        obj = out_dir / (src[:-3] + ".o")
        cmd = [_nvcc(), *common, "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} (exit {p.returncode})\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc compilation failed")
    cmd = [_nvcc(), "-shared", *ARCH, "-o", str(lib), *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc link failed")
    return lib


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    var = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--variant=")), None)
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, variant=var)
    print(path)
