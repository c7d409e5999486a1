import sys
sys.path.insert(0, "/root/repo")
import torch
from __graft_entry__ import load_package
P = load_package()
dev = torch.device("cuda", 0)
D, Hq, Hk = 128, 32, 8
n_local = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
q8 = len(sys.argv) > 2 and sys.argv[2] == "q8"
def mk(seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    x = (torch.rand((1, Hk, n_local, D), generator=g, device=dev) * 2 - 1).half()
    return P.quantize_q8_0(x) if q8 else x
nsets = 3
ks = [mk(1 + s) for s in range(nsets)]; vs = [mk(10 + s) for s in range(nsets)]
q = torch.rand((1, Hq, 1, D), device=dev) * 2 - 1
ws = P.Workspace(P.workspace_size(0, 8 if q8 else 1, D, 1, Hq, 1, n_local, Hk, 1))
xch = P.PeerExchange.local(1, Hq, D)[0]
part = torch.empty((Hq, D + 2), device=dev); dst = torch.empty((Hq, D), device=dev)
def timeit(fn, steps=200, chunk=20):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(chunk): fn(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps // chunk): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3
fl = P.FLAG_WORKSPACE_ZEROED
a = timeit(lambda i: P.flash_attn_partial(q, ks[i % nsets], vs[i % nsets], n_kv_total=n_local, workspace=ws, out=part, flags=fl))
def three(i):
    P.flash_attn_partial_scatter(q, ks[i % nsets], vs[i % nsets], xch, n_kv_total=n_local, workspace=ws, flags=fl)
    P.merge_partials_wait(xch, dst=dst)
b = timeit(three)
c = timeit(lambda i: P.flash_attn_seqpar(q, ks[i % nsets], vs[i % nsets], xch, n_kv_total=n_local, workspace=ws, dst=dst, flags=fl))
print(f"n_local={n_local} q8={q8}: partial only {a:.1f} us | partial+scatter+merge (3 launches) {b:.1f} us | fused one kernel {c:.1f} us")
xch.close()
