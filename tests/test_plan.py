"""CPU: the host-side plan (b200fa_plan) — which kernel family a shape gets and how its work is cut, for BASELINE.json's configs and
the dispatch boundaries.  No GPU involved: this is the planning code of csrc/b200fa_api.cu run for a 148-SM device."""
import pytest

from __graft_entry__ import load_package

F32, F16, Q8 = 0, 1, 8


def P():
    return load_package()


def plan(*a, **kw):
    return P().plan(*a, **kw)


def test_configs():
    p = P()
    c1 = plan(F32, F16, 128, 1, 1, 1, 256, 1)                       # C1: 1 head, 256 keys
    assert c1.kind == p.PLAN_STREAM and c1.grid == 4                 # 4 chunks of 64 keys: no more CTAs than chunks
    c2 = plan(F32, F16, 128, 1, 32, 1, 4096, 32)                     # C2: 32 units on 148 SMs -> 4 CTAs per unit, clustered
    assert c2.kind == p.PLAN_STREAM and c2.grid == 128 and c2.cluster_k == 4 and c2.kv_div == 1
    c3 = plan(F16, F16, 128, 2048, 32, 1, 2048, 32, flags=1)         # C3: 256 work items >= 148 SMs: no split
    assert c3.kind == p.PLAN_PREFILL and c3.n_splits == 1 and c3.kv_f16_copy_bytes == 0
    c4 = plan(F32, F16, 128, 1, 32, 64, 8192, 8)                     # C4: 512 units, flat stream-K list over all SMs
    assert c4.kind == p.PLAN_STREAM and c4.grid == 148 and c4.cluster_k == 0
    c5 = plan(F32, Q8, 128, 1, 32, 1, 131072, 8)                     # C5: 8 units -> 18 CTAs each (144), too many for a cluster
    assert c5.kind == p.PLAN_STREAM and c5.grid == 144 and c5.cluster_k == 0
    c5g = plan(F32, Q8, 128, 1, 32, 1, 131072 // 8, 8)               # its per-GPU slice at 8 GPUs
    assert c5g.kind == p.PLAN_STREAM and c5g.grid == 144


def test_dispatch_boundaries():
    p = P()
    assert plan(F32, F16, 128, 16, 8, 1, 4096, 8).kind == p.PLAN_STREAM          # 16 rows per KV head: still the stream kernel
    b = plan(F32, F16, 128, 8, 32, 8, 8192, 8)                                    # 8 positions x GQA 4 = 32 rows: one packed tile per KV head,
    assert b.kind == p.PLAN_PREFILL and b.kv_div == 1 and b.n_splits == 2         # 64 items x 2 KV segments
    b = plan(F32, F16, 128, 16, 16, 1, 4096, 2)                                   # 16 x GQA 8 = 128 rows: a full packed tile
    assert b.kind == p.PLAN_PREFILL and b.kv_div == 1
    b = plan(F32, F16, 128, 8, 32, 8, 8192, 8, flags=2)                           # B200FA_FLAG_NO_TCGEN05: virtual heads on the stream kernel
    assert b.kind == p.PLAN_STREAM and b.kv_div == 2
    b = plan(F32, F16, 128, 5, 12, 1, 3000, 2)                                    # 5 x GQA 6 = 30 rows (not a power of two) -> 2 virtual heads of 15 rows
    assert b.kind == p.PLAN_STREAM and b.kv_div == 2
    b = plan(F32, F16, 128, 16, 24, 1, 4096, 2)                                   # 16 x GQA 12 = 192 rows: beyond the virtual heads
    assert b.kind == p.PLAN_ROWS16
    assert plan(F32, F16, 128, 16, 32, 1, 4096, 2).kind == p.PLAN_PREFILL         # 16 x GQA 16 = 256 rows: two packed tiles (one item) per KV head
    assert plan(F32, F16, 128, 17, 8, 1, 4096, 8).kind == p.PLAN_PREFILL          # more than 16 positions: the tile kernel
    assert plan(F32, F16, 128, 17, 8, 1, 4096, 8, flags=2).kind == p.PLAN_ROWS16  # ... unless B200FA_FLAG_NO_TCGEN05
    assert plan(F32, F16, 80, 1, 8, 1, 1000, 8).kind == p.PLAN_STREAM             # padded head sizes
    assert plan(F32, F16, 80, 300, 8, 1, 300, 8).kind == p.PLAN_PREFILL


def test_split_kv_prefill_plan():
    p = P()
    a = plan(F16, F16, 128, 256, 32, 1, 32768, 32)            # 32 items, 256 KV tiles: 9 segments -> 288 units = 2 waves of 148
    assert a.kind == p.PLAN_PREFILL and a.n_splits == 9
    assert a.workspace_bytes >= 9 * 256 * 32 * 132 * 4
    b = plan(F16, F16, 128, 256, 4, 1, 4096, 4)               # 4 items, 32 KV tiles: at most 4 segments of >= 8 tiles
    assert b.n_splits == 4
    c = plan(F16, F16, 128, 256, 4, 1, 1024, 4)               # 8 KV tiles: too short to split
    assert c.n_splits == 1
    d = plan(F16, F16, 128, 2048, 32, 1, 32768, 32)           # enough items: no split
    assert d.n_splits == 1
    e = plan(F16, F16, 96, 256, 4, 1, 4096, 4)                # padded head size: the partial path is 128-wide only
    assert e.kind == p.PLAN_PREFILL and e.n_splits == 1


def test_q8_0_prefill_plan():
    p = P()
    a = plan(F16, Q8, 128, 2048, 32, 1, 2048, 32, flags=1)
    assert a.kind == p.PLAN_PREFILL and a.kv_f16_copy_bytes == 2 * 2048 * 32 * 128 * 2
    b = plan(F16, Q8, 128, 512, 32, 1, 131072, 32)            # copies of 1 GiB each: over the limit -> the 16-row kernel
    assert b.kind == p.PLAN_ROWS16 and b.kv_f16_copy_bytes == 0


def test_plan_matches_workspace_size():
    """b200fa_workspace_size covers the plan of every entry point for the shape, so it is never smaller than this plan."""
    p = P()
    for args in [(F32, F16, 128, 1, 32, 1, 4096, 32, 1), (F16, F16, 128, 256, 32, 1, 32768, 32, 1), (F16, Q8, 128, 2048, 32, 1, 2048, 32, 1),
                 (F32, F16, 128, 8, 32, 8, 8192, 8, 8), (F32, Q8, 128, 1, 32, 1, 131072, 8, 1)]:
        pl = plan(*args)
        assert p.workspace_size(*args) >= pl.workspace_bytes, args


def test_head_sizes_above_128_plan_the_16_row_kernel():
    p = P()
    for D, n_q in [(256, 1), (192, 1), (256, 300), (160, 20)]:
        assert plan(F32, F16, D, n_q, 32, 1, 4096, 8).kind == p.PLAN_ROWS16
    assert plan(F32, Q8, 256, 1, 32, 1, 4096, 8).kind == p.PLAN_ROWS16


def test_plan_rejects_bad_shapes():
    p = P()
    for bad in [dict(D=100), dict(D=264), dict(H=30), dict(n_kv=0), dict(kv=Q8, D=96)]:
        a = dict(q=F32, kv=F16, D=128, n_q=1, H=32, B=1, n_kv=256, Hk=8)
        a.update(bad)
        with pytest.raises(p.B200FAError):
            plan(a["q"], a["kv"], a["D"], a["n_q"], a["H"], a["B"], a["n_kv"], a["Hk"])
