#!/usr/bin/env python
"""Diagnostics: per-item clock64 stamps of one CTA of the persistent prefill kernel (b200fa_debug_set).
  python profiles/prefill_items.py [--n 2048] [--cta 0] [--nocausal]"""
import argparse, ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from __graft_entry__ import load_package
P = load_package()
ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=2048); ap.add_argument("--cta", type=int, default=0)
ap.add_argument("--nocausal", action="store_true"); ap.add_argument("--heads", type=int, default=32)
a = ap.parse_args()
dev = torch.device("cuda", 0); D = 128; H = a.heads
q, k, v = [(torch.rand((1, H, a.n, D), device=dev) * 2 - 1).half() for _ in range(3)]
dump = torch.zeros(2 * 128 * 128 + 256, dtype=torch.float32, device=dev)
lib = P.lib(); lib.b200fa_debug_set.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int]
for rep in range(3):
    dump.zero_()
    lib.b200fa_debug_set(None, dump.data_ptr(), a.cta)
    P.flash_attn_ext(q, k, v, None, flags=0 if a.nocausal else P.FLAG_CAUSAL); torch.cuda.synchronize()
lib.b200fa_debug_set(None, None, 0)
st = dump.view(torch.int64)[:2 * 32 * 8].cpu().numpy().reshape(2, 32, 8)
t0 = min(st[t][0, 0] for t in range(2) if st[t][0, 0] > 0)
for t in range(2):
    print(f"tile {t}: item | work idx | halves | start | softmax loop | -> O read out | -> stored | cycles per half")
    for k in range(32):
        r = st[t][k]
        if r[0] == 0: break
        print(f"   {k:2d} | {r[4]:5d} | {r[5]:3d} | {r[0]-t0:8d} | {r[1]-r[0]:7d} | {r[2]-r[1]:6d} | {r[3]-r[2]:6d} | {(r[1]-r[0])/max(r[5],1):7.0f}")
    print(f"   total {st[t][:, 3].max() - t0} cycles")
