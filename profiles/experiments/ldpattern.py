#!/usr/bin/env python
"""Builds ldpattern.cu as a shared library and runs its five load-shape variants (see the .cu header)."""
import ctypes as C
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "libldpattern.so")
if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, "ldpattern.cu")):
    subprocess.run(["nvcc", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared", "-o", so,
                    os.path.join(HERE, "ldpattern.cu")], check=True)
if torch.cuda.is_available():
    lib = C.CDLL(so)
    lib.ldpattern_run.restype = C.c_float
    lib.ldpattern_run.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    total = 2 << 30
    buf = torch.ones(total, dtype=torch.uint8, device="cuda"); sink = torch.zeros(1, dtype=torch.int32, device="cuda")
    names = ["LDG.128  8 rows x  64 B (half lines)", "LDG.128  4 rows x 128 B", "LDG.256  8 rows x 128 B", "LDG.256  4 rows x 256 B",
             "LDG.128  linear 512 B", "decode shape: K 8x64B + V 4x128B, 1 GiB apart", "same, CTA slabs at 512 KB power-of-2 bases"]
    for ctas in (296, 2048, 4144):
        print(f"ctas={ctas} (4 warps each, 2 resident per SM, 16 x 16 B per lane in flight)")
        for m, n in enumerate(names):
            v = lib.ldpattern_run(m, buf.data_ptr(), total, ctas, sink.data_ptr())
            print(f"  mode{m} {n:40s}: {v:7.0f} GB/s", flush=True)
            if v < 0:
                raise SystemExit("CUDA error, stopping")
    torch.cuda.synchronize()
