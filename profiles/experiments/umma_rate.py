#!/usr/bin/env python
"""Builds umma_rate.cu and runs it (see the .cu header)."""
import ctypes as C, os, subprocess
HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "libumma_rate.so")
src = os.path.join(HERE, "umma_rate.cu")
if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
    subprocess.run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared", "-o", so, src], check=True)
import torch
if torch.cuda.is_available():
    torch.zeros(1, device="cuda")
    C.CDLL(so).umma_rate_main()
