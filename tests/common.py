"""Shared helpers for the parity tests: data recipes, tolerance, golden loaders."""
import json
import os

import numpy as np

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

# north_star tolerance for f16 output; f32 output is held to the same bound (it is far inside it)
ATOL, RTOL = 2e-3, 1e-2


def make_mask(kind, n_q, n_kv):
    """f16 [n_q][n_kv] additive mask, rows = queries (flash-llama.h:151,194). None for 'none'."""
    if kind == "none":
        return None
    m = np.zeros((n_q, n_kv), np.float16)
    if kind == "tail56":
        m[:, n_kv - 56:] = -np.inf
    elif kind == "causal":  # bottom-right aligned: query i sees kv <= i + (n_kv - n_q)
        for i in range(n_q):
            m[i, max(0, i + (n_kv - n_q) + 1):] = -np.inf
    elif kind == "noise":
        m[:] = oracle.uniform_pm1(4, (n_q, n_kv)).astype(np.float16)
    elif kind != "zeros":
        raise ValueError(kind)
    return m


def pad_mask_rows(mask, multiple=32):
    """ggml pads mask rows to 32 (flash-matrix.cu:203 PADD(batch,32)); padding rows are zeros."""
    if mask is None:
        return None
    n_q, n_kv = mask.shape
    rows = (n_q + multiple - 1) // multiple * multiple
    out = np.zeros((rows, n_kv), np.float16)
    out[:n_q] = mask
    return out


def synth_qkv(D, n_q, n_kv, n_head, n_head_kv, n_batch=1, seeds=(1, 2, 3)):
    """Reference recipe (utils.h:57-61, kernel_test.h:45-47) with fixed seeds.
    Dense per-head layout: Q f32 [b][head][n_q][D], K/V f16 [b][kv_head][n_kv][D]."""
    Q = oracle.uniform_pm1(seeds[0], (n_batch, n_head, n_q, D))
    K = oracle.uniform_pm1(seeds[1], (n_batch, n_head_kv, n_kv, D)).astype(np.float16)
    V = oracle.uniform_pm1(seeds[2], (n_batch, n_head_kv, n_kv, D)).astype(np.float16)
    return Q, K, V


def assert_close(got, ref, what="", atol=ATOL, rtol=RTOL):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert np.isfinite(got).all(), f"{what}: non-finite values in result"
    err = np.abs(got - ref)
    bound = atol + rtol * np.abs(ref)
    worst = float((err / bound).max()) if err.size else 0.0
    assert worst <= 1.0, (f"{what}: max_abs={err.max():.3e} max_rel={(err / np.maximum(np.abs(ref), 1e-6)).max():.3e} "
                          f"worst err/bound={worst:.2f} (atol={atol}, rtol={rtol})")
    return float(err.max()) if err.size else 0.0


def load_kat():
    return json.load(open(os.path.join(GOLDEN, "kat_flash_attn_f32.json")))


def load_ref_host_cases():
    return np.load(os.path.join(GOLDEN, "ref_host_cases.npz"))
