"""ctypes bindings over include/b200fa.h.

`flash_attn_ext_raw` is the ABI call itself: pointers, ne/nb, flags, workspace, stream — the argument
list of the reference kernel (flash-llama.h:6-32).  `flash_attn_ext` derives ne/nb from torch tensors
laid out the ggml way (torch shape [ne3][ne2][ne1][ne0], any strides) and owns nothing but the call.
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB, build

TYPE_F32, TYPE_F16, TYPE_Q8_0 = 0, 1, 8
FLAG_CAUSAL, FLAG_NO_TCGEN05, FLAG_WORKSPACE_ZEROED = 1, 2, 4
Q8_BLOCK_BYTES, Q8_BLOCK_ELEMS = 34, 32

_lib = None


class B200FAError(RuntimeError):
    def __init__(self, status: int, what: str):
        self.status = status
        super().__init__(f"{what}: {lib().b200fa_status_string(status).decode()} ({status})")


def lib() -> C.CDLL:
    """Load libb200fa.so (building it if the sources are newer).  Fails loudly if it cannot be had."""
    global _lib
    if _lib is None:
        # B200FA_LIB: a tuning build of the same sources (profiles/ tools compare variants in one GPU call)
        path = os.environ.get("B200FA_LIB") or (LIB if os.path.exists(LIB) and os.environ.get("B200FA_NO_REBUILD") else build())
        l = C.CDLL(path)
        i64, vp = C.c_int64, C.c_void_p
        l.b200fa_status_string.restype = C.c_char_p
        l.b200fa_status_string.argtypes = [C.c_int]
        l.b200fa_last_dispatch.restype = C.c_char_p
        l.b200fa_last_launch_count.restype = C.c_int
        l.b200fa_version.restype = C.c_int
        l.b200fa_flash_attn_ext.restype = C.c_int
        l.b200fa_flash_attn_ext.argtypes = [vp] * 5 + [C.c_float] + [C.c_int] * 3 + [i64] * 23 + [C.c_uint32, vp, C.c_size_t, vp]
        l.b200fa_flash_attn_ext2.restype = C.c_int
        l.b200fa_flash_attn_ext2.argtypes = [vp] * 5 + [C.c_float] + [C.c_int] * 3 + [i64] * 23 + [vp, C.c_uint32, vp, C.c_size_t, vp]
        l.b200fa_flash_attn_partial.restype = C.c_int
        l.b200fa_flash_attn_partial.argtypes = [vp] * 5 + [C.c_float] + [C.c_int] * 2 + [i64] * 21 + [C.c_uint32, vp, C.c_size_t, vp]
        l.b200fa_workspace_size.restype = C.c_size_t
        l.b200fa_workspace_size.argtypes = [C.c_int, C.c_int] + [i64] * 7 + [C.c_uint32]
        l.b200fa_workspace_init.restype = C.c_int
        l.b200fa_workspace_init.argtypes = [vp, C.c_size_t, vp]
        l.b200fa_merge_partials.restype = C.c_int
        l.b200fa_merge_partials.argtypes = [vp, C.c_int, i64, i64, vp, C.c_int, vp]
        l.b200fa_quantize_q8_0.restype = C.c_int
        l.b200fa_quantize_q8_0.argtypes = [vp, C.c_int, vp, i64, vp]
        l.b200fa_dequantize_q8_0.restype = C.c_int
        l.b200fa_dequantize_q8_0.argtypes = [vp, vp, i64, vp]
        l.b200fa_xchg_bytes.restype = C.c_size_t
        l.b200fa_xchg_bytes.argtypes = [C.c_int, i64, i64]
        l.b200fa_flash_attn_partial_scatter.restype = C.c_int
        l.b200fa_flash_attn_partial_scatter.argtypes = [vp] * 4 + [C.c_float] + [C.c_int] * 2 + [i64] * 21 + [vp, vp, C.c_int, C.c_int, C.c_uint32, vp, C.c_size_t, vp]
        l.b200fa_flash_attn_seqpar.restype = C.c_int
        l.b200fa_flash_attn_seqpar.argtypes = [vp] * 5 + [C.c_float] + [C.c_int] * 3 + [i64] * 21 + [vp, vp, C.c_int, C.c_int, C.c_uint32, vp, C.c_size_t, vp]
        l.b200fa_merge_partials_wait.restype = C.c_int
        l.b200fa_merge_partials_wait.argtypes = [vp, C.c_int, i64, i64, vp, C.c_int, vp]
        l.b200fa_peer_alloc.restype = C.c_int
        l.b200fa_peer_alloc.argtypes = [C.c_size_t, C.POINTER(vp), C.c_char_p]
        l.b200fa_peer_open.restype = C.c_int
        l.b200fa_peer_open.argtypes = [C.c_char_p, C.POINTER(vp)]
        l.b200fa_peer_close.restype = C.c_int
        l.b200fa_peer_close.argtypes = [vp]
        l.b200fa_peer_free.restype = C.c_int
        l.b200fa_peer_free.argtypes = [vp]
        l.b200fa_peer_set_timeout.restype = C.c_int
        l.b200fa_peer_set_timeout.argtypes = [vp, C.c_int, vp]
        l.b200fa_peer_status.restype = C.c_int
        l.b200fa_peer_status.argtypes = [vp, C.POINTER(C.c_int), vp]
        l.b200fa_peer_reset.restype = C.c_int
        l.b200fa_peer_reset.argtypes = [vp, vp]
        l.b200fa_kv_cache_append.restype = C.c_int
        l.b200fa_kv_cache_append.argtypes = [vp, C.c_int, vp, C.c_int] + [i64] * 12 + [vp]
        l.b200fa_debug_timeline.restype = None
        l.b200fa_debug_timeline.argtypes = [vp]
        _lib = l
    return _lib


def last_dispatch() -> str:
    return lib().b200fa_last_dispatch().decode()


def last_launch_count() -> int:
    return lib().b200fa_last_launch_count()


def _stream_ptr(stream=None) -> int:
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream


def _ne_nb(t, type_: int):
    """torch tensor [ne3][ne2][ne1][ne0] -> (ne, nb) ggml style; q8_0 tensors are uint8 [..., row_bytes]."""
    shape = (1,) * (4 - t.dim()) + tuple(t.shape)
    es = t.element_size()
    strides = tuple(s * es for s in t.stride())
    strides = ((strides[0] * t.shape[0],) * (4 - t.dim()) + strides) if t.dim() < 4 else strides
    ne = tuple(reversed(shape)); nb = tuple(reversed(strides))
    if type_ == TYPE_Q8_0:
        ne = (ne[0] // Q8_BLOCK_BYTES * Q8_BLOCK_ELEMS,) + ne[1:]
        nb = (Q8_BLOCK_BYTES,) + nb[1:]
    return ne, nb


def _type_of(t, explicit=None) -> int:
    import torch
    if explicit is not None:
        return explicit
    return {torch.float32: TYPE_F32, torch.float16: TYPE_F16, torch.uint8: TYPE_Q8_0}[t.dtype]


def workspace_size(q_type, kv_type, ne00, ne01, ne02, ne03, ne11, ne12, ne13, flags=0) -> int:
    return lib().b200fa_workspace_size(q_type, kv_type, ne00, ne01, ne02, ne03, ne11, ne12, ne13, flags)


def _temp_workspace(nbytes, device, stream):
    """Scratch for a call that was given none.  The launch is asynchronous and the wrapper drops the buffer on return: that is
    only safe on the allocating (current) stream, so for any other stream the caching allocator is told about the use."""
    import torch
    ws = Workspace(nbytes, device)
    if stream is not None and stream != torch.cuda.current_stream(device):
        ws.buf.record_stream(stream)
    return ws


class _on_device_of:
    """The C side enqueues on the CURRENT device: make that the device of the tensors for the duration of the call."""

    def __init__(self, t):
        import torch
        self.g = torch.cuda.device(t.device)

    def __enter__(self):
        return self.g.__enter__()

    def __exit__(self, *a):
        return self.g.__exit__(*a)


class Workspace:
    """Caller-owned scratch (the reference cudaMallocs its own per call, flash-matrix.cu:223-224)."""

    def __init__(self, nbytes: int, device=None):
        import torch
        self.nbytes = max(int(nbytes), 256)
        self.buf = torch.zeros(self.nbytes + 256, dtype=torch.uint8, device=device or "cuda")  # zero-filled once
        self.ptr = (self.buf.data_ptr() + 255) // 256 * 256


class ExtParams(C.Structure):
    """b200fa_ext_params: the upstream-ggml score modifiers of b200fa_flash_attn_ext2."""
    _fields_ = [("max_bias", C.c_float), ("logit_softcap", C.c_float),
                ("mask_ne2", C.c_int64), ("mask_ne3", C.c_int64), ("mask_nb2", C.c_int64), ("mask_nb3", C.c_int64)]


class PlanInfo(C.Structure):
    """b200fa_plan_info."""
    _fields_ = [("kind", C.c_int32), ("kv_div", C.c_int32), ("n_splits", C.c_int32), ("grid", C.c_int32), ("cluster_k", C.c_int32),
                ("reserved", C.c_int32), ("kv_f16_copy_bytes", C.c_int64), ("workspace_bytes", C.c_int64)]


PLAN_PREFILL, PLAN_STREAM, PLAN_ROWS16 = 0, 1, 2


def plan(q_type, kv_type, D, n_q, n_head, n_batch, n_kv, n_head_kv, n_batch_kv=None, flags=0, sm_count=148) -> PlanInfo:
    """The host-side plan for a shape (no GPU needed): kernel family, virtual heads, splits, grid, workspace."""
    l = lib()
    l.b200fa_plan.restype = C.c_int
    l.b200fa_plan.argtypes = [C.c_int, C.c_int] + [C.c_int64] * 7 + [C.c_uint32, C.c_int, C.POINTER(PlanInfo)]
    info = PlanInfo()
    rc = l.b200fa_plan(q_type, kv_type, D, n_q, n_head, n_batch, n_kv, n_head_kv, n_batch_kv or n_batch, flags, sm_count, C.byref(info))
    if rc != 0:
        raise B200FAError(rc, "b200fa_plan")
    return info


def flash_attn_ext_raw(q, k, v, mask, dst, scale, q_type, kv_type, dst_type, q_ne, k_ne, ne31, nb31, q_nb, k_nb, v_nb,
                       flags, ws_ptr, ws_bytes, stream_ptr, max_bias: float = 0.0, logit_softcap: float = 0.0, mask_slices=None) -> int:
    """The ABI call.  q,k,v,mask,dst are device addresses (ints); returns the status code.
    mask_slices = (ne32, ne33, nb32, nb33): one mask per head / per batch entry (b200fa_ext_params)."""
    if max_bias != 0.0 or logit_softcap != 0.0 or mask_slices is not None:
        ext = ExtParams(max_bias, logit_softcap, *(mask_slices or (0, 0, 0, 0)))
        return lib().b200fa_flash_attn_ext2(
            q, k, v, mask, dst, scale, q_type, kv_type, dst_type, *q_ne, *k_ne, ne31, nb31,
            q_nb[1], q_nb[2], q_nb[3], k_nb[1], k_nb[2], k_nb[3], v_nb[1], v_nb[2], v_nb[3],
            q_ne[0], q_ne[2], q_ne[1], q_ne[3], C.byref(ext), flags, ws_ptr, ws_bytes, stream_ptr)
    return lib().b200fa_flash_attn_ext(
        q, k, v, mask, dst, scale, q_type, kv_type, dst_type, *q_ne, *k_ne, ne31, nb31,
        q_nb[1], q_nb[2], q_nb[3], k_nb[1], k_nb[2], k_nb[3], v_nb[1], v_nb[2], v_nb[3],
        q_ne[0], q_ne[2], q_ne[1], q_ne[3], flags, ws_ptr, ws_bytes, stream_ptr)


def flash_attn_ext(q, k, v, mask=None, scale=None, dst=None, dst_dtype=None, flags=0, workspace: Workspace | None = None,
                   stream=None, kv_type=None, max_bias: float = 0.0, logit_softcap: float = 0.0):
    """dst[b][q][head][D] = softmax(scale·QKᵀ + mask)·V on the current CUDA device.
    max_bias / logit_softcap != 0 select b200fa_flash_attn_ext2 (ALiBi slopes on the mask, tanh soft-cap of the scores).

    q  : [n_batch][n_head][n_q][D] view (any strides with 16-byte aligned rows), f32 or f16
    k,v: [n_batch_kv][n_head_kv][n_kv][D] f16 views, or uint8 [..][n_kv][D/32*34] for q8_0
    mask: f16 [>= n_q][n_kv] or None — or 4-D [ne33][ne32][>= n_q][n_kv] with ne32 in {1, n_head}, ne33 in {1, n_batch}: one mask
          slice per head and / or per batch entry (upstream ggml's broadcast; goes through b200fa_flash_attn_ext2).
    Returns dst (allocated if not given).
    """
    import torch
    if not q.is_cuda:
        raise B200FAError(-4, "b200fa has no CPU path")
    qt, kt = _type_of(q), _type_of(k, kv_type)
    q_ne, q_nb = _ne_nb(q, qt); k_ne, k_nb = _ne_nb(k, kt); _, v_nb = _ne_nb(v, kt)
    D, n_q, n_head, n_b = q_ne
    if scale is None:
        scale = 1.0 / (D ** 0.5)
    if dst is None:
        dst = torch.empty((n_b, n_q, n_head, D), dtype=dst_dtype or torch.float32, device=q.device)
    dt = _type_of(dst)
    slices = None
    if mask is not None and mask.dim() == 4:
        slices = (mask.shape[1], mask.shape[0], mask.stride(1) * 2, mask.stride(0) * 2)
        m_ptr, ne31, nb31 = mask.data_ptr(), mask.shape[2], mask.stride(2) * 2
    else:
        m_ptr, ne31, nb31 = (mask.data_ptr(), mask.shape[0], mask.stride(0) * 2) if mask is not None else (None, 0, 0)
    if workspace is None:
        ext = ExtParams(max_bias, logit_softcap, *(slices or (0, 0, 0, 0)))
        l = lib()
        l.b200fa_workspace_size_ext2.restype = C.c_size_t
        l.b200fa_workspace_size_ext2.argtypes = [C.c_int, C.c_int] + [C.c_int64] * 7 + [C.c_void_p, C.c_uint32]
        workspace = _temp_workspace(l.b200fa_workspace_size_ext2(qt, kt, *q_ne, k_ne[1], k_ne[2], k_ne[3], C.byref(ext), flags), q.device, stream)
    with _on_device_of(q):
        rc = flash_attn_ext_raw(q.data_ptr(), k.data_ptr(), v.data_ptr(), m_ptr, dst.data_ptr(), scale, qt, kt, dt,
                            q_ne, k_ne, ne31, nb31, q_nb, k_nb, v_nb, flags, workspace.ptr, workspace.nbytes,
                            _stream_ptr(stream), max_bias, logit_softcap, slices)
    if rc != 0:
        raise B200FAError(rc, "b200fa_flash_attn_ext")
    return dst


def flash_attn_partial(q, k, v, mask=None, scale=None, kv_pos0=0, n_kv_total=None, flags=0,
                       workspace: Workspace | None = None, stream=None, kv_type=None, out=None, max_bias: float = 0.0, logit_softcap: float = 0.0):
    """Sequence-split building block: per-row (O~[D], m, l) over this device's KV slice -> f32 [rows][D+2].
    max_bias / logit_softcap: the score modifiers of flash_attn_ext (b200fa_flash_attn_partial2); `mask` holds this slice's columns."""
    import torch
    qt, kt = _type_of(q), _type_of(k, kv_type)
    q_ne, q_nb = _ne_nb(q, qt); k_ne, k_nb = _ne_nb(k, kt); _, v_nb = _ne_nb(v, kt)
    D, n_q, n_head, n_b = q_ne
    if scale is None:
        scale = 1.0 / (D ** 0.5)
    if n_kv_total is None:
        n_kv_total = kv_pos0 + k_ne[1]
    if out is None:
        out = torch.empty((n_b * n_q * n_head, D + 2), dtype=torch.float32, device=q.device)
    if workspace is None:
        workspace = _temp_workspace(workspace_size(qt, kt, *q_ne, k_ne[1], k_ne[2], k_ne[3], flags), q.device, stream)
    m_ptr, ne31, nb31 = (mask.data_ptr(), mask.shape[0], mask.stride(0) * 2) if mask is not None else (None, 0, 0)
    ext = ExtParams(max_bias, logit_softcap, 0, 0, 0, 0)
    l = lib()
    l.b200fa_flash_attn_partial2.restype = C.c_int
    l.b200fa_flash_attn_partial2.argtypes = [C.c_void_p] * 5 + [C.c_float] + [C.c_int] * 2 + [C.c_int64] * 21 + [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p]
    with _on_device_of(q):
        rc = l.b200fa_flash_attn_partial2(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), m_ptr, out.data_ptr(), scale, qt, kt, *q_ne, *k_ne, ne31, nb31,
        q_nb[1], q_nb[2], q_nb[3], k_nb[1], k_nb[2], k_nb[3], v_nb[1], v_nb[2], v_nb[3], kv_pos0, n_kv_total,
        C.byref(ext) if (max_bias != 0.0 or logit_softcap != 0.0) else None, flags, workspace.ptr, workspace.nbytes, _stream_ptr(stream))
    if rc != 0:
        raise B200FAError(rc, "b200fa_flash_attn_partial")
    return out


def merge_partials(partials, dst=None, dst_dtype=None, stream=None):
    """partials f32 [n_parts][rows][D+2] -> dst [rows][D]  (fa_reduce algebra, flash_row_float.h:415-472)."""
    import torch
    n_parts, rows, d2 = partials.shape
    if dst is None:
        dst = torch.empty((rows, d2 - 2), dtype=dst_dtype or torch.float32, device=partials.device)
    rc = lib().b200fa_merge_partials(partials.data_ptr(), n_parts, rows, d2 - 2, dst.data_ptr(), _type_of(dst),
                                     _stream_ptr(stream))
    if rc != 0:
        raise B200FAError(rc, "b200fa_merge_partials")
    return dst


def quantize_q8_0(x, stream=None):
    """f32/f16 [..., D] (contiguous) -> uint8 [..., D/32*34] ggml q8_0 rows."""
    import torch
    x = x.contiguous()
    out = torch.empty(x.shape[:-1] + (x.shape[-1] // Q8_BLOCK_ELEMS * Q8_BLOCK_BYTES,), dtype=torch.uint8, device=x.device)
    rc = lib().b200fa_quantize_q8_0(x.data_ptr(), _type_of(x), out.data_ptr(), x.numel(), _stream_ptr(stream))
    if rc != 0:
        raise B200FAError(rc, "b200fa_quantize_q8_0")
    return out


def dequantize_q8_0(y, stream=None):
    import torch
    y = y.contiguous()
    out = torch.empty(y.shape[:-1] + (y.shape[-1] // Q8_BLOCK_BYTES * Q8_BLOCK_ELEMS,), dtype=torch.float32, device=y.device)
    rc = lib().b200fa_dequantize_q8_0(y.data_ptr(), out.data_ptr(), out.numel(), _stream_ptr(stream))
    if rc != 0:
        raise B200FAError(rc, "b200fa_dequantize_q8_0")
    return out


def kv_cache_append(src, cache, n_past: int, cache_type=None, stream=None):
    """cache[b][head][n_past + i][:] = convert(src[b][i][head][:]).

    src  : [n_batch][n_tokens][n_head_kv][D] f32 or f16 (any strides with contiguous rows) — the layout a projection produces
    cache: [n_batch][n_head_kv][n_kv_max][D] f16 view, or uint8 [..][n_kv_max][D/32*34] for q8_0 (any strides)."""
    st, ct = _type_of(src), _type_of(cache, cache_type)
    n_b, n_tok, n_hk, D = src.shape
    es = src.element_size()
    ces = cache.element_size()
    rc = lib().b200fa_kv_cache_append(
        src.data_ptr(), st, cache.data_ptr(), ct, D, n_tok, n_hk, n_b,
        src.stride(1) * es, src.stride(2) * es, src.stride(0) * es,
        cache.stride(2) * ces, cache.stride(1) * ces, cache.stride(0) * ces, n_past, cache.shape[2], _stream_ptr(stream))
    if rc != 0:
        raise B200FAError(rc, "b200fa_kv_cache_append")
    return cache


class PeerExchange:
    """Exchange buffers for the sequence-split combine over peer-mapped memory (b200fa_flash_attn_partial_scatter /
    b200fa_merge_partials_wait).  `PeerExchange.local(world, rows, D)` emulates `world` ranks on ONE device (tests);
    `PeerExchange.distributed(rows, D, group)` allocates this rank's buffer, exchanges cudaIpc handles through
    torch.distributed and maps every peer's buffer (one process per GPU on one node)."""

    def __init__(self, world, rank, rows, D, own_ptr, peer_ptrs, owned, opened, single_device=False):
        import torch
        self.world, self.rank, self.rows, self.D = world, rank, rows, D
        self.single_device = single_device  # all ranks' buffers live on one device (local()): the fused one-kernel step cannot run
        self.own_ptr, self.peer_ptrs, self._owned, self._opened = own_ptr, list(peer_ptrs), owned, opened
        self.peers_dev = torch.tensor(self.peer_ptrs, dtype=torch.int64, device="cuda")  # device array of pointers

    @staticmethod
    def nbytes(world, rows, D):
        return lib().b200fa_xchg_bytes(world, rows, D)

    @classmethod
    def local(cls, world, rows, D):
        """`world` buffers on the current device; returns one PeerExchange per emulated rank.
        Only the launch-separated protocol works on such an exchange (flash_attn_partial_scatter by every rank FIRST, then
        merge_partials_wait): a kernel that waits for a peer's kernel queued behind it on the same device never sees it arrive,
        so flash_attn_seqpar refuses an emulated exchange with world > 1."""
        ptrs, handles = [], []
        for _ in range(world):
            p = C.c_void_p(); h = C.create_string_buffer(64)
            rc = lib().b200fa_peer_alloc(cls.nbytes(world, rows, D), C.byref(p), h)
            if rc != 0:
                raise B200FAError(rc, "b200fa_peer_alloc")
            ptrs.append(p.value)
        return [cls(world, r, rows, D, ptrs[r], ptrs, owned=[ptrs[r]], opened=[], single_device=True) for r in range(world)]

    @classmethod
    def distributed(cls, rows, D, group=None):
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        p = C.c_void_p(); h = C.create_string_buffer(64)
        rc = lib().b200fa_peer_alloc(cls.nbytes(world, rows, D), C.byref(p), h)
        if rc != 0:
            raise B200FAError(rc, "b200fa_peer_alloc")
        handles = [None] * world
        dist.all_gather_object(handles, bytes(h.raw), group=group)
        ptrs, opened = [], []
        for r in range(world):
            if r == rank:
                ptrs.append(p.value)
            else:
                q = C.c_void_p()
                rc = lib().b200fa_peer_open(handles[r], C.byref(q))
                if rc != 0:
                    raise B200FAError(rc, "b200fa_peer_open")
                ptrs.append(q.value); opened.append(q.value)
        return cls(world, rank, rows, D, p.value, ptrs, owned=[p.value], opened=opened)

    def set_timeout(self, ms: int, stream=None):
        """Timeout of the device-side waits on this rank's buffer (default 4 s); a timed-out wait raises the error flag, never traps."""
        rc = lib().b200fa_peer_set_timeout(self.own_ptr, int(ms), _stream_ptr(stream))
        if rc != 0:
            raise B200FAError(rc, "b200fa_peer_set_timeout")

    def timed_out(self, stream=None) -> bool:
        """True if a wait on this rank's buffer timed out since the last reset (synchronises the stream)."""
        flag = C.c_int(0)
        rc = lib().b200fa_peer_status(self.own_ptr, C.byref(flag), _stream_ptr(stream))
        if rc != 0:
            raise B200FAError(rc, "b200fa_peer_status")
        return bool(flag.value)

    def reset(self, stream=None):
        """Back to step 0 (call on EVERY rank at the same point, e.g. after a step that timed out or returned an error somewhere)."""
        rc = lib().b200fa_peer_reset(self.own_ptr, _stream_ptr(stream))
        if rc != 0:
            raise B200FAError(rc, "b200fa_peer_reset")

    def close(self):
        for q in self._opened:
            lib().b200fa_peer_close(q)
        for q in self._owned:
            lib().b200fa_peer_free(q)
        self._opened, self._owned = [], []


def flash_attn_partial_scatter(q, k, v, xch: PeerExchange, mask=None, scale=None, kv_pos0=0, n_kv_total=None, flags=0,
                               workspace: Workspace | None = None, stream=None, kv_type=None):
    """This rank's (O~, m, l) triples -> its slot of every rank's exchange buffer, over NVLink stores (no NCCL)."""
    qt, kt = _type_of(q), _type_of(k, kv_type)
    q_ne, q_nb = _ne_nb(q, qt); k_ne, k_nb = _ne_nb(k, kt); _, v_nb = _ne_nb(v, kt)
    D, n_q, n_head, n_b = q_ne
    if scale is None:
        scale = 1.0 / (D ** 0.5)
    if n_kv_total is None:
        n_kv_total = kv_pos0 + k_ne[1]
    if workspace is None:
        workspace = _temp_workspace(workspace_size(qt, kt, *q_ne, k_ne[1], k_ne[2], k_ne[3], flags), q.device, stream)
    m_ptr, ne31, nb31 = (mask.data_ptr(), mask.shape[0], mask.stride(0) * 2) if mask is not None else (None, 0, 0)
    with _on_device_of(q):
        rc = lib().b200fa_flash_attn_partial_scatter(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), m_ptr, scale, qt, kt, *q_ne, *k_ne, ne31, nb31,
        q_nb[1], q_nb[2], q_nb[3], k_nb[1], k_nb[2], k_nb[3], v_nb[1], v_nb[2], v_nb[3], kv_pos0, n_kv_total,
        xch.own_ptr, xch.peers_dev.data_ptr(), xch.rank, xch.world, flags, workspace.ptr, workspace.nbytes, _stream_ptr(stream))
    if rc != 0:
        raise B200FAError(rc, "b200fa_flash_attn_partial_scatter")


def merge_partials_wait(xch: PeerExchange, dst=None, dst_dtype=None, stream=None):
    """Waits on the device for all ranks' triples of the current step, then merges them -> dst [rows][D]."""
    import torch
    if dst is None:
        dst = torch.empty((xch.rows, xch.D), dtype=dst_dtype or torch.float32, device="cuda")
    rc = lib().b200fa_merge_partials_wait(xch.own_ptr, xch.world, xch.rows, xch.D, dst.data_ptr(), _type_of(dst), _stream_ptr(stream))
    if rc != 0:
        raise B200FAError(rc, "b200fa_merge_partials_wait")
    return dst


def flash_attn_seqpar(q, k, v, xch: PeerExchange, mask=None, scale=None, kv_pos0=0, n_kv_total=None, flags=0, dst=None, dst_dtype=None,
                      workspace: Workspace | None = None, stream=None, kv_type=None):
    """One sequence-parallel step in one call (one kernel for decode shapes): this rank's KV band -> dst [rows][D] on every rank."""
    import torch
    if xch.single_device and xch.world > 1:
        raise B200FAError(-2, "flash_attn_seqpar on an exchange emulated on one device (PeerExchange.local): the ranks' kernels cannot "
                              "run concurrently there; use flash_attn_partial_scatter + merge_partials_wait")
    qt, kt = _type_of(q), _type_of(k, kv_type)
    q_ne, q_nb = _ne_nb(q, qt); k_ne, k_nb = _ne_nb(k, kt); _, v_nb = _ne_nb(v, kt)
    D, n_q, n_head, n_b = q_ne
    if scale is None:
        scale = 1.0 / (D ** 0.5)
    if n_kv_total is None:
        n_kv_total = kv_pos0 + k_ne[1]
    if dst is None:
        dst = torch.empty((n_b * n_q * n_head, D), dtype=dst_dtype or torch.float32, device=q.device)
    if workspace is None:
        workspace = _temp_workspace(workspace_size(qt, kt, *q_ne, k_ne[1], k_ne[2], k_ne[3], flags), q.device, stream)
    m_ptr, ne31, nb31 = (mask.data_ptr(), mask.shape[0], mask.stride(0) * 2) if mask is not None else (None, 0, 0)
    with _on_device_of(q):
        rc = lib().b200fa_flash_attn_seqpar(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), m_ptr, dst.data_ptr(), scale, qt, kt, _type_of(dst), *q_ne, *k_ne, ne31, nb31,
        q_nb[1], q_nb[2], q_nb[3], k_nb[1], k_nb[2], k_nb[3], v_nb[1], v_nb[2], v_nb[3], kv_pos0, n_kv_total,
        xch.own_ptr, xch.peers_dev.data_ptr(), xch.rank, xch.world, flags, workspace.ptr, workspace.nbytes, _stream_ptr(stream))
    if rc != 0:
        raise B200FAError(rc, "b200fa_flash_attn_seqpar")
    return dst
