#!/usr/bin/env python
"""Diagnostics: per-CTA %globaltimer stamps of the stream decode kernel (b200fa_debug_timeline).
  python profiles/timeline.py --hq 32 --hk 32 --nkv 4096 [--q8] [--batch B]
Prints, relative to the earliest CTA start: start, first landed stage, end of streaming, fold done, end (ns; min/median/max over CTAs)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from __graft_entry__ import load_package  # noqa: E402

P = load_package()
ap = argparse.ArgumentParser()
ap.add_argument("--hq", type=int, default=32); ap.add_argument("--hk", type=int, default=32)
ap.add_argument("--nkv", type=int, default=4096); ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--q8", action="store_true"); ap.add_argument("--nomask", action="store_true")
ap.add_argument("--cold", action="store_true", help="flush L2 before the stamped call (code, Q and mask then also come from DRAM)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
D = 128
row_bytes = 136 if a.q8 else 256
kv_bytes = 2 * a.batch * a.hk * a.nkv * row_bytes
nsets = min(12, max(3, int(400e6 // kv_bytes) + 1))  # K/V sets rotate (> L2 in total): K/V always come from DRAM, code and Q stay warm
ks, vs = [], []
for _ in range(nsets):
    k = (torch.rand((a.batch, a.hk, a.nkv, D), device=dev) * 2 - 1).half(); v = (torch.rand((a.batch, a.hk, a.nkv, D), device=dev) * 2 - 1).half()
    if a.q8:
        k, v = P.quantize_q8_0(k), P.quantize_q8_0(v)
    ks.append(k); vs.append(v)
q = torch.rand((a.batch, a.hq, 1, D), device=dev) * 2 - 1
mask = None if a.nomask else torch.zeros((32, a.nkv), dtype=torch.float16, device=dev)
stamps = torch.zeros((160 + 256, 8), dtype=torch.int64, device=dev)  # 160 CTAs x 8, then 512 chunks x 4 of CTA 0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for it in range(3):
    if a.cold:
        flush.fill_(it)  # evict everything from L2 (code and Q too)
    else:
        for w in range(2 * nsets):  # warm code / Q / mask; K/V of the stamped call were last touched nsets calls ago
            P.flash_attn_ext(q, ks[w % nsets], vs[w % nsets], mask)
    stamps.zero_()
    torch.cuda.synchronize()
    P.lib().b200fa_debug_timeline(stamps.data_ptr())
    P.flash_attn_ext(q, ks[0], vs[0], mask)
    torch.cuda.synchronize()
    P.lib().b200fa_debug_timeline(None)
    s_all = stamps.cpu().numpy()
    s = s_all[:160]
    ch = s_all[160:].reshape(-1, 4)
    s = s[s[:, 0] > 0]
    t0 = s[:, 7].min() if (s[:, 7] > 0).all() else s[:, 0].min()  # earliest kernel entry (tuning builds stamp it)
    names = ["start", "first stage", "stream end", "fold done", "end"]
    print(f"run {it}: {P.last_dispatch()} ctas={len(s)}")
    for i, n in enumerate(names):
        col = s[:, i] - t0
        print(f"  {n:12s} min {col.min():7d}  med {int(np.median(col)):7d}  max {col.max():7d} ns")
    if it == 2:
        order = np.argsort(s[:, 2])
        print("  per CTA (sorted by end of streaming): cta | first  stream_end  fold_done  end")
        for o in list(order[:6]) + list(order[-24:]):
            r = s[o]
            print(f"    {o:4d} | {r[1]-t0:6d} {r[2]-t0:6d} {r[3]-t0:6d} {r[4]-t0:6d}")
        print("  the 10 CTAs that end last: cta | stream_end fold_done synced end")
        for o in np.argsort(s[:, 4])[-10:]:
            r = s[o]
            print(f"    {o:4d} | {r[2]-t0:6d} {r[3]-t0:6d} {r[5]-t0:6d} {r[4]-t0:6d}   producer exit {r[6]-t0:6d}  idx1 {r[1]-t0:6d}")

# per-chunk view of CTA 0 (ns): stage freed -> operations issued -> landed -> released by the first consumer warp
ch = ch[(ch[:, 2] > 0)]
e0 = s_all[0, 7] if s_all[0, 7] > 0 else s_all[0, 0]
print("CTA 0 entry -> ready for first chunk:", s_all[0, 6] - e0, "ns")
print("CTA 0 entry -> consumers start:", s_all[0, 0] - e0, "ns; first chunks (free, issued, landed, done) relative to entry:")
for r in ch[:14]:
    print("   ", [int(x - e0) if x > 0 else -1 for x in r])
if len(ch) > 16:
    c = ch[12:-2]
    have_p = c[:, 0] > 0
    print(f"CTA 0, {len(c)} chunks: issue cost med {int(np.median((c[:,1]-c[:,0])[have_p]))} ns; issued->landed med {int(np.median((c[:,2]-c[:,1])[have_p]))} ns; "
          f"landed->released med {int(np.median(c[:,3]-c[:,2]))} ns; chunk period med {int(np.median(np.diff(c[:,2])))} ns")
    nxt = c[1:, 0] - c[:-1, 3]
