"""-m gpu: seeded random shapes through the C ABI against the CPU oracle — every kernel family, every mask kind, f16/q8_0
K/V, f32/f16 Q and dst, GQA and batch broadcast, cache-view strides.  Deterministic (fixed seeds)."""
import numpy as np
import pytest

import oracle
from common import make_mask, synth_qkv
from gpu_common import pkg, run_both

pytestmark = pytest.mark.gpu


def _case(seed):
    r = np.random.RandomState(seed)
    D = int(r.choice([128, 128, 128, 64]))
    gqa = int(r.choice([1, 1, 2, 4, 8]))
    Hk = int(r.choice([1, 2, 3, 4, 8]))
    H = Hk * gqa
    B = int(r.choice([1, 1, 2, 3]))
    family = r.choice(["decode", "decode", "burst", "prefill", "prefill"])
    if family == "decode":
        n_q = 1
        n_kv = int(r.choice([1, 7, 64, 65, 127, 300, 1000, 2049, 4100]))
    elif family == "burst":
        n_q = int(r.choice([2, 3, 5, 9, 17, 33]))
        n_kv = int(r.choice([n_q, 64, 200, 513, 1025]))
    else:
        n_q = int(r.choice([64, 65, 128, 130, 255, 256, 300, 384, 520]))
        n_kv = int(r.choice([n_q, n_q + 1, n_q + 77, 2 * n_q, max(64, n_q - 30), 640]))
    mask_kind = str(r.choice(["none", "zeros", "noise", "causal"]))
    if mask_kind == "causal" and n_kv < n_q:
        mask_kind = "noise"
    q8 = bool(r.rand() < 0.3)
    if q8 and (n_kv * (D // 32 * 34)) % 16 and family != "prefill":
        n_kv += 1  # keep some q8_0 cases on the stream kernel (even row counts), others fall back by construction
    return dict(D=D, n_q=n_q, n_kv=n_kv, H=H, Hk=Hk, B=B, mask_kind=mask_kind, q8=q8, q_f16=bool(r.rand() < 0.4),
                dst_f16=bool(r.rand() < 0.3), cache_view=bool(r.rand() < 0.4) and not q8, flag=bool(r.rand() < 0.5), seed=seed)


@pytest.mark.parametrize("seed", list(range(64)))
def test_random_shape(seed):
    c = _case(seed)
    Q, K, V = synth_qkv(c["D"], c["n_q"], c["n_kv"], c["H"], c["Hk"], n_batch=c["B"], seeds=(seed + 1, seed + 2, seed + 3))
    mask = make_mask(c["mask_kind"], c["n_q"], c["n_kv"])
    flags = pkg().FLAG_CAUSAL if (c["mask_kind"] == "causal" and c["flag"]) else 0
    run_both(Q, K, V, mask, flags=flags, q_f16=c["q_f16"], dst_f16=c["dst_f16"], cache_view=c["cache_view"], q8=c["q8"],
             mask_pad=32 if c["flag"] else None, what=str(c))


def _case2(seed):
    """Shapes around the dispatch boundaries added late in round 1: bursts on virtual KV heads, small n_q on the tile kernel,
    split-KV prefill (few items, long KV), q8_0 on the prefill path, padded head sizes everywhere."""
    r = np.random.RandomState(1000 + seed)
    D = int(r.choice([128, 128, 64, 80, 96, 40]))
    gqa = int(r.choice([1, 2, 4, 8]))
    Hk = int(r.choice([1, 2, 3]))
    H = Hk * gqa
    B = int(r.choice([1, 1, 2]))
    family = r.choice(["vheads", "small_q", "chunk", "chunk"])
    if family == "vheads":
        n_q = int(r.choice([2, 3, 5, 8, 11, 16]))
        n_kv = int(r.choice([n_q + 3, 200, 1000, 2500]))
    elif family == "small_q":
        n_q = int(r.choice([17, 20, 33, 48, 63, 64, 65]))
        n_kv = int(r.choice([n_q, 130, 700, 1500]))
    else:
        n_q = int(r.choice([17, 64, 100, 128, 200, 256]))
        n_kv = int(r.choice([2048, 2100, 3000, 4200]))
    mask_kind = str(r.choice(["none", "zeros", "noise", "causal", "causal"]))
    if mask_kind == "causal" and n_kv < n_q:
        mask_kind = "noise"
    q8 = bool(r.rand() < 0.35) and D in (64, 128)
    return dict(D=D, n_q=n_q, n_kv=n_kv, H=H, Hk=Hk, B=B, mask_kind=mask_kind, q8=q8, q_f16=bool(r.rand() < 0.4),
                dst_f16=bool(r.rand() < 0.3), cache_view=bool(r.rand() < 0.4) and not q8, flag=bool(r.rand() < 0.5), seed=seed)


@pytest.mark.parametrize("seed", list(range(72)))
def test_random_shape_dispatch_boundaries(seed):
    c = _case2(seed)
    Q, K, V = synth_qkv(c["D"], c["n_q"], c["n_kv"], c["H"], c["Hk"], n_batch=c["B"], seeds=(seed + 11, seed + 12, seed + 13))
    mask = make_mask(c["mask_kind"], c["n_q"], c["n_kv"])
    flags = pkg().FLAG_CAUSAL if (c["mask_kind"] == "causal" and c["flag"]) else 0
    run_both(Q, K, V, mask, flags=flags, q_f16=c["q_f16"], dst_f16=c["dst_f16"], cache_view=c["cache_view"], q8=c["q8"],
             mask_pad=32 if c["flag"] else None, what=str(c))


def _case3(seed):
    """Round 2: GQA-packed tiles — every power-of-two group, query counts from one position to several packed tile pairs, KV ranges
    long enough to be split into segments, padded head sizes, q8_0 caches, masks with and without the flag."""
    r = np.random.RandomState(2000 + seed)
    D = int(r.choice([128, 128, 128, 64, 80, 112]))
    gqa = int(r.choice([2, 4, 8, 16, 32]))
    Hk = int(r.choice([1, 2, 3])) if gqa < 32 else 1
    H = Hk * gqa
    B = int(r.choice([1, 1, 2]))
    n_q = int(r.choice([1, 2, 3, 5, 8, 13, 16, 17, 31, 32, 33, 50, 64, 100, 127]))
    if n_q * gqa <= 16:
        n_q = 16 // gqa + 1
    n_kv = int(r.choice([n_q + 5, 129, 640, 1000, 2048, 3000, 4500]))
    mask_kind = str(r.choice(["none", "zeros", "noise", "causal", "causal"]))
    if mask_kind == "causal" and n_kv < n_q:
        mask_kind = "noise"
    q8 = bool(r.rand() < 0.3) and D in (64, 128)
    return dict(D=D, n_q=n_q, n_kv=n_kv, H=H, Hk=Hk, B=B, mask_kind=mask_kind, q8=q8, q_f16=bool(r.rand() < 0.4),
                dst_f16=bool(r.rand() < 0.3), cache_view=bool(r.rand() < 0.4) and not q8, flag=bool(r.rand() < 0.5), seed=seed)


@pytest.mark.parametrize("seed", list(range(48)))
def test_random_shape_packed_tiles(seed):
    c = _case3(seed)
    Q, K, V = synth_qkv(c["D"], c["n_q"], c["n_kv"], c["H"], c["Hk"], n_batch=c["B"], seeds=(seed + 21, seed + 22, seed + 23))
    mask = make_mask(c["mask_kind"], c["n_q"], c["n_kv"])
    flags = pkg().FLAG_CAUSAL if (c["mask_kind"] == "causal" and c["flag"]) else 0
    run_both(Q, K, V, mask, flags=flags, q_f16=c["q_f16"], dst_f16=c["dst_f16"], cache_view=c["cache_view"], q8=c["q8"],
             mask_pad=32 if c["flag"] else None, what=str(c))
    assert pkg().last_dispatch() == "prefill_tcgen05", (pkg().last_dispatch(), c)
