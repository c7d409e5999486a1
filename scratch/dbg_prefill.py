import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import oracle
from common import make_mask, synth_qkv
from gpu_common import pkg, to_dev
P = pkg()
for (n_q, n_kv) in [(200, 129), (130, 70), (200, 257), (256, 129), (128, 129)]:
    for mk in ["none", "noise"]:
        H, Hk = 4, 2
        Q, K, V = synth_qkv(128, n_q, n_kv, H, Hk)
        mask = make_mask(mk, n_q, n_kv)
        ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), oracle.view_of(mask) if mask is not None else None, 1/np.sqrt(128), round_q_f16=True)
        out = P.flash_attn_ext(to_dev(Q), to_dev(K), to_dev(V), to_dev(mask) if mask is not None else None)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        err = np.abs(got - ref)  # [b][q][head][D]
        bad = err > 2e-3 + 1e-2*np.abs(ref)
        print(n_q, n_kv, mk, P.last_dispatch(), "max", err.max(), "bad rows:", np.unique(np.nonzero(bad)[1])[:20], "heads", np.unique(np.nonzero(bad)[2]), "dims", np.unique(np.nonzero(bad)[3])[:10])
