"""Diagnostic driver for the tcgen05 prefill kernel (not a pytest file): runs one small case with the
kernel's dump hooks on and reports which stage (S = QK^T, P/O = PV, final) first deviates.
Usage: python tests/prefill_debug.py [n_q n_kv n_head causal]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402

import oracle  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402


def main():
    n_q, n_kv, H, causal = (int(x) for x in (sys.argv[1:5] + ["128", "128", "1", "0"][len(sys.argv) - 1:]))
    P = load_package()
    lib = P.lib()
    D = 128
    Q = oracle.uniform_pm1(1, (1, H, n_q, D)); K = oracle.uniform_pm1(2, (1, H, n_kv, D)).astype(np.float16)
    V = oracle.uniform_pm1(3, (1, H, n_kv, D)).astype(np.float16)
    Qh = Q.astype(np.float16)
    word = torch.zeros(1, dtype=torch.int64).pin_memory()
    dump = torch.zeros(2 * 128 * 128 + 256, dtype=torch.float32, device="cuda")
    n_qt = (n_q + 127) // 128
    dump_cta = (n_qt - 1) * H  # CTA order is heavy-first: query tile 0 / head 0 is at (n_qt-1)*H
    lib.b200fa_debug_set(C.c_void_p(word.data_ptr()), C.c_void_p(dump.data_ptr()), dump_cta)
    q, k, v = (torch.from_numpy(x).cuda() for x in (Qh, K, V))
    mask = None
    if causal:
        mask = np.zeros((n_q, n_kv), np.float16)
        for i in range(n_q):
            mask[i, i + (n_kv - n_q) + 1:] = -np.inf
    try:
        out = P.flash_attn_ext(q, k, v, None, flags=P.FLAG_CAUSAL if causal else 0)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print("LAUNCH/SYNC FAILED:", e, "timeout word = %#x" % (word.item() & 0xFFFFFFFFFFFFFFFF))
        return 1
    print("dispatch:", P.last_dispatch(), "launches:", P.last_launch_count(), "timeout word = %#x" % (word.item() & 0xFFFFFFFFFFFFFFFF))
    d = dump.cpu().numpy()
    S = d[:128 * 128].reshape(128, 128); O = d[128 * 128:2 * 128 * 128].reshape(128, 128)
    l = d[2 * 128 * 128:2 * 128 * 128 + 128]; m = d[2 * 128 * 128 + 128:]
    rows = min(128, n_q); cols = min(128, n_kv)
    Sref = Qh[0, 0, :rows].astype(np.float64) @ K[0, 0, :cols].astype(np.float64).T
    eS = np.abs(S[:rows, :cols] - Sref)
    print(f"S  (QK^T tile 0): max err {eS.max():.3e}   |S|max {np.abs(Sref).max():.3f}")
    if eS.max() > 1e-2:
        bad = np.argwhere(eS > 1e-2)
        print("   first bad (row, col):", bad[:8].tolist(), " n_bad:", len(bad))
        print("   S[0,:8]   ", S[0, :8]); print("   Sref[0,:8]", Sref[0, :8])
        print("   S[1,:8]   ", S[1, :8]); print("   Sref[1,:8]", Sref[1, :8])
        # is it a permutation / transposition?
        print("   corr with Sref^T:", np.abs(S[:rows, :cols] - Sref.T[:rows, :cols]).max() if rows == cols else "n/a")
    ref = oracle.flash_attn_ext(oracle.view_of(Qh), oracle.view_of(K), oracle.view_of(V),
                                oracle.view_of(mask) if mask is not None else None, 1 / np.sqrt(D))
    got = out.cpu().numpy()
    e = np.abs(got - ref)
    print(f"final: max err {e.max():.3e}  finite={np.isfinite(got).all()}  l[0..3]={l[:4]}  m[0..3]={m[:4]}")
    if e.max() > 2e-3:
        # unnormalised O / l of the dumped CTA vs reference rows
        On = O / np.maximum(l[:, None], 1e-30)
        eo = np.abs(On[:rows] - ref[0, :rows, 0, :])
        print(f"   dumped-CTA O/l vs ref: max err {eo.max():.3e}; worst rows {np.argsort(-eo.max(1))[:6].tolist()}")
        print("   got[0,0,0,:8]", got[0, 0, 0, :8]); print("   ref[0,0,0,:8]", ref[0, 0, 0, :8])
        idx = np.unravel_index(np.argmax(e), e.shape)
        print("   worst idx", idx, got[idx], ref[idx])
    return 0 if e.max() <= 2e-3 else 2


if __name__ == "__main__":
    sys.exit(main())
