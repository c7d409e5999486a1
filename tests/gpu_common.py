"""Helpers for the -m gpu parity tests: numpy inputs -> product call through the C ABI -> numpy."""
import numpy as np

import oracle
from __graft_entry__ import load_package
from common import assert_close


def pkg():
    return load_package()


def to_dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def kv_cache_view(t):
    """Dense [b][h][n][D] torch tensor -> same logical view over memory laid out [b][n][h][D]
    (the ggml KV-cache view the reference's test_llama passes, flash-matrix.cu:203-204)."""
    return t.permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3)


def run_both(Q, K, V, mask, scale=None, flags=0, q_f16=False, dst_f16=False, cache_view=False, q8=False,
             mask_pad=None, atol=None, rtol=None, what="", drop_mask_for_product=False):
    """Q f32 [b][h][q][D], K/V f16 [bk][hk][kv][D], mask f16 [q][kv] or None.  Returns (got, ref)."""
    import torch
    P = pkg()
    D = Q.shape[-1]
    scale = scale if scale is not None else 1.0 / np.sqrt(D)
    Qn = Q.astype(np.float16) if q_f16 else Q
    if q8:
        Kq = oracle.quantize_q8_0(K.astype(np.float32)); Vq = oracle.quantize_q8_0(V.astype(np.float32))
        kview, vview = oracle.view_of(Kq, oracle.TYPE_Q8_0), oracle.view_of(Vq, oracle.TYPE_Q8_0)
        k, v = to_dev(Kq), to_dev(Vq)
    else:
        kview, vview = oracle.view_of(K), oracle.view_of(V)
        k, v = to_dev(K), to_dev(V)
    ref = oracle.flash_attn_ext(oracle.view_of(Qn), kview, vview, oracle.view_of(mask) if mask is not None else None,
                                scale, round_q_f16=True)
    q = to_dev(Qn)
    if cache_view:
        q, k, v = kv_cache_view(q), kv_cache_view(k), kv_cache_view(v)
    m = None
    if mask is not None and not drop_mask_for_product:
        mm = mask
        if mask_pad:
            rows = (mask.shape[0] + mask_pad - 1) // mask_pad * mask_pad
            mm = np.zeros((rows, mask.shape[1]), np.float16); mm[:mask.shape[0]] = mask
        m = to_dev(mm)
    out = P.flash_attn_ext(q, k, v, m, scale=scale, flags=flags, dst_dtype=torch.float16 if dst_f16 else torch.float32)
    torch.cuda.synchronize()
    got = out.float().cpu().numpy()
    kw = {}
    if atol is not None: kw["atol"] = atol
    if rtol is not None: kw["rtol"] = rtol
    assert_close(got, ref, what or P.last_dispatch(), **kw)
    return got, ref
