"""In-tree build of the CUDA library (sm_100a only).  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libb200fa.so")
SOURCES = ["b200fa_api.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + ["../../include/b200fa.h"]
TUNING_LIB = os.path.join(OUT_DIR, "libb200fa_tuning.so")  # -DB200FA_TUNING: env knobs, timeline stamps (profiles/ tools only)
EXAMPLE_SRC = os.path.join(os.path.dirname(HERE), "examples", "kernel_test_dropin.cu")
EXAMPLE_BIN = os.path.join(os.path.dirname(HERE), "examples", "_build", "kernel_test_dropin")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, tuning: bool = False) -> str:
    """Compile csrc/ into _build/libb200fa.so (skipped when up to date).  Returns the library path.
    tuning=True builds _build/libb200fa_tuning.so instead (same sources, -DB200FA_TUNING; load it with B200FA_LIB=...)."""
    out = TUNING_LIB if tuning else LIB
    if not force and not _stale(out):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OUT_DIR, exist_ok=True)
    extra = os.environ.get("B200FA_NVCC_EXTRA", "").split() + (["-DB200FA_TUNING"] if tuning else [])
    cmd = [nvcc, *NVCC_FLAGS, *extra, *(["-Xptxas", "-v"] if verbose else []), "-o", out,
           *[os.path.join(CSRC, s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


def build_example(force: bool = False) -> str:
    """Compile examples/kernel_test_dropin.cu — the reference's kernel_test flow as a CUDA C++ caller of the C ABI — against
    include/b200fa.h and link it with libb200fa.so (rpath relative to the binary, so it runs from the tree on any box)."""
    lib = build()
    if not force and os.path.exists(EXAMPLE_BIN) and os.path.getmtime(EXAMPLE_BIN) >= max(os.path.getmtime(EXAMPLE_SRC), os.path.getmtime(lib)):
        return EXAMPLE_BIN
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(os.path.dirname(EXAMPLE_BIN), exist_ok=True)
    cmd = [nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-I", os.path.join(os.path.dirname(HERE), "include"),
           EXAMPLE_SRC, "-L", OUT_DIR, "-lb200fa", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../ggml-cuda-experiments_b200/_build", "-o", EXAMPLE_BIN]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed (example):\n" + res.stdout + res.stderr)
    return EXAMPLE_BIN


if __name__ == "__main__":
    print(build(force=True, verbose=True))
