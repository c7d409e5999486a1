import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import oracle
from common import make_mask, synth_qkv
from gpu_common import pkg, to_dev
P = pkg()
cases = [(200,256,"noise"), (200,257,"zeros"), (192,257,"noise"), (196,257,"noise"), (200,258,"noise"),(200,264,"noise"), (200,257,"noise"), (72, 257, "noise"), (200,385,"noise")]
for (n_q, n_kv, mk) in cases:
    H, Hk = 4, 2
    Q, K, V = synth_qkv(128, n_q, n_kv, H, Hk)
    mask = make_mask(mk, n_q, n_kv)
    ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), oracle.view_of(mask) if mask is not None else None, 1/np.sqrt(128), round_q_f16=True)
    q,k,v = to_dev(Q), to_dev(K), to_dev(V); m = to_dev(mask) if mask is not None else None
    out = P.flash_attn_ext(q, k, v, m); torch.cuda.synchronize()
    out = P.flash_attn_ext(q, k, v, m); torch.cuda.synchronize()
    got = out.cpu().numpy()
    err = np.abs(got - ref)
    bad = err > 2e-3 + 1e-2*np.abs(ref)
    idx = np.nonzero(bad)
    print(n_q, n_kv, mk, P.last_dispatch(), "max", err.max(), "rows", np.unique(idx[1])[:12], "n_rows", len(np.unique(idx[1])))
    # padded mask rows variant
    if mask is not None:
        mm = np.zeros(((n_q+31)//32*32, n_kv), np.float16); mm[:n_q] = mask
        out = P.flash_attn_ext(q, k, v, to_dev(mm)); torch.cuda.synchronize()
        err = np.abs(out.cpu().numpy() - ref); bad = err > 2e-3 + 1e-2*np.abs(ref)
        print("   padded mask: max", err.max(), "n_rows", len(np.unique(np.nonzero(bad)[1])))
