import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import oracle
from common import make_mask, synth_qkv
from gpu_common import pkg, to_dev
P = pkg()
cases = [(256,256,4,4,"none"), (128,256,2,2,"none"), (384,384,4,1,"none"), (200,257,4,2,"noise"), (512,512,8,2,"causal")]
for (n_q, n_kv, H, Hk, mk) in cases:
    Q, K, V = synth_qkv(128, n_q, n_kv, H, Hk)
    mask = make_mask(mk, n_q, n_kv)
    ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), oracle.view_of(mask) if mask is not None else None, 1/np.sqrt(128), round_q_f16=True)
    q,k,v = to_dev(Q), to_dev(K), to_dev(V); m = to_dev(mask) if mask is not None else None
    nbad = 0
    for rep in range(200):
        out = P.flash_attn_ext(q, k, v, m)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        err = np.abs(got - ref)
        bad = err > 2e-3 + 1e-2*np.abs(ref)
        if bad.any():
            nbad += 1
            if nbad <= 3:
                idx = np.nonzero(bad)
                print("  rep", rep, "max", err.max(), "rows", np.unique(idx[1])[:12], "n_rows", len(np.unique(idx[1])), "heads", np.unique(idx[2]), "dims", len(np.unique(idx[3])))
    print(n_q, n_kv, H, Hk, mk, "bad reps:", nbad, "/200")
