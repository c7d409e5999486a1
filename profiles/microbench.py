#!/usr/bin/env python
"""Tuning tool: times one attention shape through the C ABI with CUDA-graph replays (rotating K/V sets > L2).
  python profiles/microbench.py --hq 32 --hk 32 --nkv 4096 [--batch 1] [--nq 1] [--q8] [--causal] [--steps 200]
Prints us/step, GB/s (K+V bytes) and TFLOP/s.  Env B200FA_SPLITS overrides the split count."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from __graft_entry__ import load_package  # noqa: E402

P = load_package()
dev = torch.device("cuda", 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hq", type=int, default=32); ap.add_argument("--hk", type=int, default=32)
    ap.add_argument("--nkv", type=int, default=4096); ap.add_argument("--nq", type=int, default=1)
    ap.add_argument("--batch", type=int, default=1); ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--q8", action="store_true"); ap.add_argument("--causal", action="store_true")
    ap.add_argument("--nomask", action="store_true"); ap.add_argument("--dense", action="store_true")
    ap.add_argument("--qf16", action="store_true"); ap.add_argument("--copy", action="store_true", help="also time a torch copy of the same bytes")
    a = ap.parse_args()
    D = 128
    row_bytes = 136 if a.q8 else 256
    kv_bytes = 2 * a.batch * a.hk * a.nkv * row_bytes
    nsets = max(2, int(400e6 // kv_bytes) + 1)
    nsets = min(nsets, 8)

    def mk(seed):
        g = torch.Generator(device=dev); g.manual_seed(seed)
        x = (torch.rand((a.batch, a.nkv, a.hk, D), generator=g, device=dev) * 2 - 1).to(torch.float16)
        if a.q8:
            return P.quantize_q8_0(x.permute(0, 2, 1, 3).contiguous())
        return x.permute(0, 2, 1, 3).contiguous() if a.dense else x.permute(0, 2, 1, 3)
    ks = [mk(10 + s) for s in range(nsets)]; vs = [mk(20 + s) for s in range(nsets)]
    q = (torch.rand((a.batch, a.hq, a.nq, D), device=dev) * 2 - 1)
    if a.qf16:
        q = q.half()
    mask = None
    if not a.nomask:
        rows = max(32, (a.nq + 31) // 32 * 32)
        mask = torch.zeros((rows, a.nkv), dtype=torch.float16, device=dev)
        if a.causal:
            mask = torch.full((rows, a.nkv), float("-inf"), dtype=torch.float16, device=dev).triu(1 + a.nkv - a.nq)
    flags = (P.FLAG_CAUSAL if a.causal else 0) | P.FLAG_WORKSPACE_ZEROED
    dst = torch.empty((a.batch, a.nq, a.hq, D), device=dev)
    qt = P.TYPE_F16 if a.qf16 else P.TYPE_F32
    ws = P.Workspace(P.workspace_size(qt, P.TYPE_Q8_0 if a.q8 else P.TYPE_F16, D, a.nq, a.hq, a.batch, a.nkv, a.hk, a.batch, flags))

    def step(i):
        P.flash_attn_ext(q, ks[i % nsets], vs[i % nsets], mask, dst=dst, flags=flags, workspace=ws)

    def timeit(fn, steps):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize()
        chunk = min(steps, 50)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(chunk):
                fn(i)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps // chunk):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (steps // chunk * chunk) * 1e3

    us = timeit(step, a.steps)
    flops = 4 * a.batch * a.hq * a.nq * a.nkv * D * (0.5 if a.causal and a.nq == a.nkv else 1.0)
    print(f"{P.last_dispatch()} launches={P.last_launch_count()} splits_env={os.environ.get('B200FA_SPLITS')} "
          f"us={us:.2f} GB/s={kv_bytes / us / 1e3:.0f} TFLOP/s={flops / us / 1e6:.1f} kv_MB={kv_bytes / 1e6:.1f} nsets={nsets}")
    if a.copy:
        src = [torch.empty(kv_bytes // 2, dtype=torch.uint8, device=dev) for _ in range(nsets)]
        dstc = torch.empty(kv_bytes // 2, dtype=torch.uint8, device=dev)
        usc = timeit(lambda i: dstc.copy_(src[i % nsets]), a.steps)
        print(f"torch copy of kv_bytes/2 (read+write = kv_bytes): us={usc:.2f} GB/s={kv_bytes / usc / 1e3:.0f}")
        red = [torch.empty(kv_bytes // 4, dtype=torch.float32, device=dev) for _ in range(nsets)]
        usr = timeit(lambda i: torch.sum(red[i % nsets]), a.steps)
        print(f"torch sum over kv_bytes (read only): us={usr:.2f} GB/s={kv_bytes / usr / 1e3:.0f}")


if __name__ == "__main__":
    main()
