import sys, os, ctypes
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import oracle
from common import make_mask, synth_qkv
from gpu_common import pkg, to_dev
P = pkg()
lib = P.lib(); lib.b200fa_debug_set.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int]
word = torch.zeros(1, dtype=torch.int64).pin_memory()
# mapped host pointer
cudart = ctypes.CDLL("libcudart.so")
dptr = ctypes.c_void_p()
cudart.cudaHostGetDevicePointer(ctypes.byref(dptr), ctypes.c_void_p(word.data_ptr()), 0)
lib.b200fa_debug_set(dptr, None, 0)
args = [int(x) for x in sys.argv[1:6]]
n_q, n_kv, H, Hk, B = args
causal = int(sys.argv[6])
Q, K, V = synth_qkv(128, n_q, n_kv, H, Hk, n_batch=B)
q, k, v = to_dev(Q), to_dev(K), to_dev(V)
try:
    out = P.flash_attn_ext(q, k, v, None, flags=P.FLAG_CAUSAL if causal else 0)
    torch.cuda.synchronize()
    print("ok", P.last_dispatch())
    mask = make_mask("causal", n_q, n_kv) if causal else None
    ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), oracle.view_of(mask) if mask is not None else None, 1/np.sqrt(128), round_q_f16=True)
    err = np.abs(out.cpu().numpy() - ref)
    print("max err", err.max())
except Exception as e:
    print("FAILED", str(e)[:200])
    w = word.item() & 0xFFFFFFFFFFFFFFFF
    print("timeout word = %#x  code=%d cta=%d parity=%d" % (w, (w >> 32) & 0xFFFF, (w >> 8) & 0xFFFFFF, w & 0xFF))
