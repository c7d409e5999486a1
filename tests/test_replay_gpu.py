"""-m gpu: capture replay (the reference's fixture test, flash-matrix.cu:66-73 / test_llama) through tensor-dump files.
The real llama.cpp captures are absent from the reference, so a capture set is synthesised in the layouts test_llama reads
(q [head][n_q][D] f32, k [head][kv][D] f16, v TRANSPOSED [head][D][kv] f16, mask f16, qkv [n_q][head][D] f32), the expected
`qkv` coming from the reference's own host attention (oracle/_ref) or the C oracle."""
import ctypes as C

import numpy as np
import pytest

import oracle
from common import make_mask, synth_qkv
from gpu_common import pkg

pytestmark = pytest.mark.gpu


def _expected(Q, K, VT, mask, scale):
    H, n_q, D = Q.shape
    Hk, n_kv, _ = K.shape
    try:
        r = oracle.ref_host()
        out = np.zeros((n_q, H, D), np.float32)
        scores = np.zeros((H, n_q, n_kv), np.float32)
        rc = r.ref_host_attention_llama(Q.ctypes.data, K.ctypes.data, VT.ctypes.data, mask.ctypes.data, out.ctypes.data,
                                        scores.ctypes.data, D, n_q, n_kv, H, Hk, C.c_float(scale), 4)
        assert rc == 0
        return out
    except OSError:
        V = np.ascontiguousarray(VT.transpose(0, 2, 1))
        ref = oracle.flash_attn_ext(oracle.view_of(Q[None]), oracle.view_of(K[None]), oracle.view_of(V[None]), oracle.view_of(mask), scale)
        return ref.reshape(n_q, H, D)


@pytest.mark.parametrize("n_q,n_kv,H,Hk,kind", [(1, 256, 32, 32, "tail56"), (1, 256, 32, 8, "zeros"), (40, 320, 8, 8, "causal"),
                                                 (130, 256, 4, 2, "noise")])
def test_replay_capture(tmp_path, n_q, n_kv, H, Hk, kind):
    P = pkg()
    D = 128
    Q, K, V = synth_qkv(D, n_q, n_kv, H, Hk)
    Q, K, V = Q[0], K[0], V[0]
    VT = np.ascontiguousarray(V.transpose(0, 2, 1))
    mask = make_mask(kind, n_q, n_kv)
    scale = 1.0 / np.sqrt(D)
    exp = _expected(Q, K, VT, mask, scale)
    tag = str(n_kv)
    paths = P.capture_paths(str(tmp_path), tag)
    for part, arr in zip(P.tensor_io.CAPTURE_PARTS, (Q, K, VT, mask, exp)):
        P.write_tensor(paths[part], f"fa-{part}", arr)
    r = P.replay_capture(str(tmp_path), tag)
    assert r["out"].shape == (n_q, H, D)
    bound = 2e-3 + 1e-2 * np.abs(r["ref"])    # north_star tolerance: max-abs 2e-3, rel 1e-2
    assert (np.abs(r["out"] - r["ref"]) <= bound).all(), (r["max_abs"], r["dispatch"])


@pytest.mark.parametrize("n_q,n_kv,H,Hk", [(1, 256, 32, 8), (5, 320, 8, 8)])
def test_replay_capture_in_the_cache_view_layout(tmp_path, n_q, n_kv, H, Hk):
    """The reference's live GPU arm reads the same files as [n_kv][head][D] for both K and V (flash-matrix.cu:141,149): the layout is
    derived from the shapes stored in the file headers, and an explicit layout that does not fit the files is an error."""
    P = pkg()
    D = 128
    Q, K, V = synth_qkv(D, n_q, n_kv, H, Hk)
    Q, K, V = Q[0], K[0], V[0]
    mask = make_mask("noise", n_q, n_kv)
    exp = _expected(Q, K, np.ascontiguousarray(V.transpose(0, 2, 1)), mask, 1.0 / np.sqrt(D))
    paths = P.capture_paths(str(tmp_path), "cv")
    Kc, Vc = np.ascontiguousarray(K.transpose(1, 0, 2)), np.ascontiguousarray(V.transpose(1, 0, 2))   # [n_kv][head][D]
    for part, arr in zip(P.tensor_io.CAPTURE_PARTS, (Q, Kc, Vc, mask, exp)):
        P.write_tensor(paths[part], f"fa-{part}", arr)
    r = P.replay_capture(str(tmp_path), "cv")
    assert r["kv_layout"] == "cache_view"
    assert (np.abs(r["out"] - r["ref"]) <= 2e-3 + 1e-2 * np.abs(r["ref"])).all(), r["max_abs"]
    r2 = P.replay_capture(str(tmp_path), "cv", kv_layout="cache_view")
    assert np.array_equal(r2["out"], r["out"])


def test_replay_capture_rejects_shapes_that_fit_no_layout(tmp_path):
    P = pkg()
    D, n_q, n_kv, H = 128, 1, 256, 4
    Q, K, V = synth_qkv(D, n_q, n_kv, H, H)
    paths = P.capture_paths(str(tmp_path), "bad")
    bad_k = np.zeros((H, n_kv // 2, D), np.float16)  # half the keys the mask announces
    for part, arr in zip(P.tensor_io.CAPTURE_PARTS, (Q[0], bad_k, bad_k, make_mask("zeros", n_q, n_kv), np.zeros((n_q, H, D), np.float32))):
        P.write_tensor(paths[part], f"fa-{part}", arr)
    with pytest.raises(P.B200FAError):
        P.replay_capture(str(tmp_path), "bad")
