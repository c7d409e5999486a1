"""Multi-GPU partitioning of the attention path.  The reference is single-GPU; these are the two
partitions north_star names, as host logic over the C ABI:

* head-parallel: the unit is a KV head together with its GQA group of query heads (flash-llama.h:128-140 maps
  q head -> kv head by integer division, so a group must stay on one rank).  Rank r owns a contiguous band of KV
  heads; Q columns, K/V and dst columns of that band are local.  No collective.
* sequence-parallel (very long KV): rank r owns a contiguous band of KV rows for every head, computes the
  (O~, m, l) triples of its band (b200fa_flash_attn_partial), one all-gather moves `rows * (D + 2)` floats per rank,
  and every rank merges (b200fa_merge_partials = the reference's fa_reduce algebra, flash_row_float.h:415-472).

The arithmetic lives here so that it can be tested without a GPU (tests/test_sharding.py runs it at world size 2 over
gloo with the CPU oracle standing in for the kernels).  The default compute hooks are the CUDA calls of api.py.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class HeadShard:
    kv_head0: int      # first KV head of this rank
    n_kv_heads: int    # KV heads of this rank
    q_head0: int       # first query head (= kv_head0 * gqa)
    n_q_heads: int     # query heads of this rank (= n_kv_heads * gqa)


def head_shard(n_head: int, n_head_kv: int, rank: int, world: int) -> HeadShard:
    """Contiguous band of KV heads (with their whole GQA groups) for `rank`.  Uneven counts are allowed: the first
    `n_head_kv % world` ranks take one extra head; a rank may own none when world > n_head_kv."""
    if n_head_kv <= 0 or n_head % n_head_kv:
        raise ValueError("n_head must be a positive multiple of n_head_kv")
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    gqa = n_head // n_head_kv
    base, extra = divmod(n_head_kv, world)
    n = base + (1 if rank < extra else 0)
    k0 = rank * base + min(rank, extra)
    return HeadShard(k0, n, k0 * gqa, n * gqa)


@dataclass(frozen=True)
class SeqShard:
    kv_pos0: int   # global position of this rank's first key
    n_local: int   # keys of this rank (0 = empty band)


def seq_shard(n_kv: int, rank: int, world: int, align: int = 64) -> SeqShard:
    """Contiguous band of KV rows for `rank`, band boundaries aligned to `align` keys (the decode kernel streams
    64-key chunks; for q8_0 an even boundary also keeps every band 16-byte aligned)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    blocks = (n_kv + align - 1) // align
    base, extra = divmod(blocks, world)
    b0 = rank * base + min(rank, extra)
    b1 = b0 + base + (1 if rank < extra else 0)
    p0, p1 = min(b0 * align, n_kv), min(b1 * align, n_kv)
    return SeqShard(p0, p1 - p0)


def slice_heads(q, k, v, shard: HeadShard):
    """Views of q [b][head][n_q][D] and k, v [b][kv head][n_kv][..] restricted to the rank's band (no copies)."""
    return (q[:, shard.q_head0:shard.q_head0 + shard.n_q_heads],
            k[:, shard.kv_head0:shard.kv_head0 + shard.n_kv_heads],
            v[:, shard.kv_head0:shard.kv_head0 + shard.n_kv_heads])


def flash_attn_ext_head_parallel(q, k, v, mask, rank: int, world: int, attn_fn=None, **kw):
    """This rank's band of dst: [b][n_q][n_q_heads of the band][D].  `q`, `k`, `v` are the FULL logical tensors (or
    anything that slices like them); only the band is touched.  No communication."""
    if attn_fn is None:
        from .api import flash_attn_ext as attn_fn
    sh = head_shard(q.shape[1], k.shape[1], rank, world)
    if sh.n_kv_heads == 0:
        return None, sh
    ql, kl, vl = slice_heads(q, k, v, sh)
    return attn_fn(ql, kl, vl, mask, **kw), sh


def flash_attn_ext_seq_parallel(q, k_local, v_local, mask_local, n_kv_total: int, rank: int, world: int, group=None,
                                partial_fn=None, merge_fn=None, flags: int = 0, scale=None, exchange=None, **kw):
    """dst [rows][D] on every rank.  `k_local`, `v_local` hold this rank's band seq_shard(n_kv_total, rank, world) of the
    KV rows; `mask_local` (or None) the matching columns of the mask.  One all-gather of the (O~, m, l) triples — over NCCL,
    or, with `exchange` (an api.PeerExchange of all ranks), over peer-mapped memory: NVLink stores from the attention kernel,
    a device-side wait and the merge, in one launch for decode shapes (b200fa_flash_attn_seqpar)."""
    if exchange is not None:
        from .api import flash_attn_seqpar
        sh = seq_shard(n_kv_total, rank, world)
        if k_local.shape[2] != sh.n_local or sh.n_local == 0:
            raise ValueError("the peer-memory path needs a non-empty band per rank")
        return flash_attn_seqpar(q, k_local, v_local, exchange, mask=mask_local, scale=scale, kv_pos0=sh.kv_pos0, n_kv_total=n_kv_total,
                                 flags=flags, **kw)
    import torch
    import torch.distributed as dist
    if partial_fn is None or merge_fn is None:
        from .api import flash_attn_partial, merge_partials
        partial_fn = partial_fn or flash_attn_partial
        merge_fn = merge_fn or merge_partials
    sh = seq_shard(n_kv_total, rank, world)
    if k_local.shape[2] != sh.n_local:
        raise ValueError(f"rank {rank} holds {k_local.shape[2]} keys, its band has {sh.n_local}")
    n_b, n_head, n_q, D = q.shape[0], q.shape[1], q.shape[2], q.shape[3]
    rows = n_b * n_q * n_head
    if sh.n_local > 0:
        part = partial_fn(q, k_local, v_local, mask_local, scale=scale, kv_pos0=sh.kv_pos0, n_kv_total=n_kv_total, flags=flags, **kw)
    else:  # an empty band contributes the neutral triple (0, -inf, 0)
        part = torch.zeros((rows, D + 2), dtype=torch.float32, device=q.device)
        part[:, D] = float("-inf")
    if world == 1:
        return merge_fn(part.view(1, rows, D + 2))
    gathered = torch.empty((world * rows, D + 2), dtype=torch.float32, device=part.device)  # rank-major concatenation
    dist.all_gather_into_tensor(gathered, part.contiguous(), group=group)
    return merge_fn(gathered.view(world, rows, D + 2))
