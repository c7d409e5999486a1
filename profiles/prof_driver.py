#!/usr/bin/env python
"""Small driver for ncu captures: runs one named workload a few times through the C ABI on plain streams
(no CUDA graphs, no torch.distributed) so every kernel shows up as its own launch.

  python profiles/prof_driver.py c2|c3|c3mask|c4|c5|chunk [iters]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from __graft_entry__ import load_package  # noqa: E402

P = load_package()
dev = torch.device("cuda", 0)
D = 128


def rnd(shape, seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    return (torch.rand(shape, generator=g, device=dev) * 2 - 1).to(torch.float16)


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    if wl == "c2":
        H, n_kv = 32, 4096
        ks = [rnd((1, n_kv, H, D), 10 + s).permute(0, 2, 1, 3) for s in range(4)]
        vs = [rnd((1, n_kv, H, D), 20 + s).permute(0, 2, 1, 3) for s in range(4)]
        q = (torch.rand((1, 1, H, D), device=dev) * 2 - 1).permute(0, 2, 1, 3)
        mask = torch.zeros((32, n_kv), dtype=torch.float16, device=dev)
        dst = torch.empty((1, 1, H, D), device=dev)
        ws = P.Workspace(P.workspace_size(0, 1, D, 1, H, 1, n_kv, H, 1))
        step = lambda i: P.flash_attn_ext(q, ks[i % 4], vs[i % 4], mask, dst=dst, workspace=ws)  # noqa: E731
    elif wl in ("c3", "c3mask"):
        n, H = 2048, 32
        q, k, v = rnd((1, H, n, D), 1), rnd((1, H, n, D), 2), rnd((1, H, n, D), 3)
        mask = torch.full((n, n), float("-inf"), dtype=torch.float16, device=dev).triu(1)
        dst = torch.empty((1, n, H, D), device=dev)
        ws = P.Workspace(P.workspace_size(1, 1, D, n, H, 1, n, H, 1))
        fl = P.FLAG_CAUSAL if wl == "c3" else 0
        step = lambda i: P.flash_attn_ext(q, k, v, mask, dst=dst, flags=fl, workspace=ws)  # noqa: E731
    elif wl == "c4":
        Hq, Hk, B, n_kv = 32, 8, 64, 8192
        k, v = rnd((B, Hk, n_kv, D), 4), rnd((B, Hk, n_kv, D), 5)
        q = torch.rand((B, Hq, 1, D), device=dev) * 2 - 1
        mask = torch.zeros((32, n_kv), dtype=torch.float16, device=dev)
        dst = torch.empty((B, 1, Hq, D), device=dev)
        ws = P.Workspace(P.workspace_size(0, 1, D, 1, Hq, B, n_kv, Hk, B))
        step = lambda i: P.flash_attn_ext(q, k, v, mask, dst=dst, workspace=ws)  # noqa: E731
    elif wl == "c5":
        Hq, Hk, n_kv = 32, 8, 131072
        kq = P.quantize_q8_0(rnd((1, Hk, n_kv, D), 6)); vq = P.quantize_q8_0(rnd((1, Hk, n_kv, D), 7))
        q = torch.rand((1, Hq, 1, D), device=dev) * 2 - 1
        dst = torch.empty((1, 1, Hq, D), device=dev)
        ws = P.Workspace(P.workspace_size(0, 8, D, 1, Hq, 1, n_kv, Hk, 1))
        step = lambda i: P.flash_attn_ext(q, kq, vq, None, dst=dst, workspace=ws)  # noqa: E731
    elif wl == "chunk":  # chunked prefill: 256 new queries against a 32K f16 cache (split-KV prefill + combine)
        n_q, n_kv, H = 256, 32768, 32
        q, k, v = rnd((1, H, n_q, D), 1), rnd((1, H, n_kv, D), 2), rnd((1, H, n_kv, D), 3)
        dst = torch.empty((1, n_q, H, D), device=dev)
        ws = P.Workspace(P.workspace_size(1, 1, D, n_q, H, 1, n_kv, H, 1))
        step = lambda i: P.flash_attn_ext(q, k, v, None, dst=dst, flags=P.FLAG_WORKSPACE_ZEROED, workspace=ws)  # noqa: E731
    else:
        raise SystemExit("unknown workload " + wl)
    for i in range(iters):
        step(i)
    torch.cuda.synchronize()
    print(wl, "ok", P.last_dispatch(), P.last_launch_count())


if __name__ == "__main__":
    main()
