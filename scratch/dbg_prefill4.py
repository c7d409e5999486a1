import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import oracle
from common import make_mask, synth_qkv
from gpu_common import pkg, to_dev
P = pkg()
n_q, n_kv, H, Hk = 200, 257, 1, 1
Q, K, V = synth_qkv(128, n_q, n_kv, H, Hk)
noise = make_mask("noise", n_q, n_kv)
def run(mask, tag):
    ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), oracle.view_of(mask), 1/np.sqrt(128), round_q_f16=True)
    out = P.flash_attn_ext(to_dev(Q), to_dev(K), to_dev(V), to_dev(mask)); torch.cuda.synchronize()
    err = np.abs(out.cpu().numpy() - ref); bad = err > 2e-3 + 1e-2*np.abs(ref)
    print(tag, "max", err.max(), "bad rows", np.unique(np.nonzero(bad)[1]))
    return out.cpu().numpy(), ref
m = noise.copy(); m[:] = noise[0:1]; run(m, "column-only mask")
m = noise.copy(); m[:, 128:] = 0; run(m, "noise only in tile 0")
m = noise.copy(); m[:, :128] = 0; m[:, 256:] = 0; run(m, "noise only in tile 1")
m = noise.copy(); m[:, :256] = 0; run(m, "noise only in tile 2 (1 col)")
m = np.zeros_like(noise); m[:, 256] = -np.inf; run(m, "last col -inf")
m = np.zeros_like(noise); m[192:, 5] = 3.0; run(m, "single bump col 5 rows>=192")
m = np.zeros_like(noise); m[:, 256] = 3.0; got, ref = run(m, "last col +3")
print(got[0,190:200,0,:3]); print(ref[0,190:200,0,:3])
