#!/usr/bin/env python
"""Diagnostics: clock64 stamps of the softmax loop of one CTA of the prefill kernel (b200fa_debug_set).
  python profiles/prefill_timeline.py [--n 2048] [--cta 0] [--nocausal]"""
import argparse, os, sys
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
lib = P.lib(); lib.b200fa_debug_set.argtypes = [__import__("ctypes").c_void_p] * 2 + [__import__("ctypes").c_int]
for rep in range(2):
    dump.zero_()
    lib.b200fa_debug_set(None, dump.data_ptr(), a.cta)
    P.flash_attn_ext(q, k, v, None, flags=0 if a.nocausal else P.FLAG_CAUSAL); torch.cuda.synchronize()
lib.b200fa_debug_set(None, None, 0)
full = dump.view(torch.int64)[:3 * 64 * 8].cpu().numpy().reshape(3, 64, 8)
st = full[:2]
mm = full[2][full[2][:, 0] > 0]
if len(mm):
    print("MMA warp: per iteration: wait V | wait P0 + issue PV0 | wait K | issue QK0 | wait P1 + issue PV1 | issue QK1 | period")
    for i, r in enumerate(mm[:20]):
        nxt = mm[i + 1, 0] if i + 1 < len(mm) else r[6]
        print(f"  it {i:2d} @ {r[0]-mm[0,0]:8d}: {r[1]-r[0]:6d} | {r[2]-r[1]:6d} | {r[3]-r[2]:6d} | {r[4]-r[3]:6d} | {r[5]-r[4]:6d} | {r[6]-r[5]:6d} | {nxt-r[0]:6d}")
for t in range(2):
    rows = st[t][st[t][:, 0] > 0]
    if len(rows) == 0: continue
    t0 = rows[0, 0]
    e = st[t][63]
    print(f"tile {t}: epilogue: last arrive -> pv_done seen {e[0] - rows[-1, 5]} | stage pass0 {e[1] - e[0]} | store pass0 {e[2] - e[1]} | stage pass1 {e[3] - e[2]} | store pass1 {e[4] - e[3]}")
    print(f"tile {t}: loop entered {st[t][63, 6] - t0} cycles before the first wait; epilogue finished {st[t][63, 7] - rows[-1, 5]} cycles after the last arrive")
    print(f"tile {t}: {len(rows)} iterations; per iteration: wait_S | tmem_ld | max(+resc) | exp+sum+pack | wait_pv+st+arrive | total   (cycles)")
    for i, r in enumerate(rows[:18]):
        nxt = rows[i + 1, 0] if i + 1 < len(rows) else r[5]
        print(f"  it {i:2d} @ {r[0]-t0:8d}: {r[1]-r[0]:6d} | {r[2]-r[1]:5d} | {r[3]-r[2]:5d} | {r[4]-r[3]:5d} | {r[5]-r[4]:5d} | {r[5]-r[0]:6d}")
