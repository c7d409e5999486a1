"""ggml tensor-dump files and capture replay — host side of include/b200fa.h's b200fa_tensor_file_* entries.

The reference replays llama.cpp captures as its fixture test (flash-matrix.cu:66-73: fa-cuda-{q,k,v,mask,qkv}-256.tensor,
loaded by utils.h:110-150).  `read_tensor` / `write_tensor` go through the C ABI (the parsing lives in
csrc/tensor_file.cuh); `replay_capture` runs a capture set through `flash_attn_ext` the way test_llama does and returns
the result beside the captured one.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import api

_NP = {api.TYPE_F32: np.float32, api.TYPE_F16: np.float16}


class TensorInfo(C.Structure):
    _fields_ = [("n_dims", C.c_int32), ("type", C.c_int32), ("ne", C.c_int64 * 4), ("name", C.c_char * 64),
                ("data_offset", C.c_int64), ("data_bytes", C.c_int64)]


def _lib():
    l = api.lib()
    l.b200fa_tensor_file_info.restype = C.c_int
    l.b200fa_tensor_file_info.argtypes = [C.c_char_p, C.POINTER(TensorInfo)]
    l.b200fa_tensor_file_read.restype = C.c_int
    l.b200fa_tensor_file_read.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
    l.b200fa_tensor_file_write.restype = C.c_int
    l.b200fa_tensor_file_write.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_void_p]
    return l


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise api.B200FAError(rc, what)


def tensor_info(path: str) -> TensorInfo:
    info = TensorInfo()
    _check(_lib().b200fa_tensor_file_info(os.fsencode(path), C.byref(info)), f"tensor_file_info({path})")
    return info


def read_tensor(path: str):
    """-> (name, ndarray) with the numpy shape = ne reversed (ne[0] is the fastest dimension), dims as stored."""
    info = tensor_info(path)
    shape = tuple(int(info.ne[i]) for i in range(info.n_dims))[::-1]
    out = np.empty(shape, dtype=_NP[info.type])
    _check(_lib().b200fa_tensor_file_read(os.fsencode(path), out.ctypes.data, out.nbytes), f"tensor_file_read({path})")
    return info.name.decode(), out


def write_tensor(path: str, name: str, array) -> None:
    """array: f32 or f16 ndarray, 1..4 dims, C order; stored with ne = shape reversed."""
    a = np.ascontiguousarray(array)
    if a.dtype == np.float32:
        t = api.TYPE_F32
    elif a.dtype == np.float16:
        t = api.TYPE_F16
    else:
        raise api.B200FAError(-2, f"write_tensor: dtype {a.dtype} is not f32/f16")
    if not 1 <= a.ndim <= 4:
        raise api.B200FAError(-1, f"write_tensor: {a.ndim} dims")
    ne = (C.c_int64 * 4)(*(list(a.shape[::-1]) + [1] * (4 - a.ndim)))
    _check(_lib().b200fa_tensor_file_write(os.fsencode(path), name.encode(), t, a.ndim, ne, a.ctypes.data if a.size else None),
           f"tensor_file_write({path})")


CAPTURE_PARTS = ("q", "k", "v", "mask", "qkv")


def capture_paths(directory: str, tag: str = "256", prefix: str = "fa-cuda") -> dict:
    return {p: os.path.join(directory, f"{prefix}-{p}-{tag}.tensor") for p in CAPTURE_PARTS}


def replay_capture(directory: str, tag: str = "256", prefix: str = "fa-cuda", device=None, flags: int = 0, kv_layout: str = "auto"):
    """Replay one capture set (flash-matrix.cu:66-73) on the GPU.  -> dict(out=, ref=, max_abs=, dispatch=, kv_layout=).

    The reference reads the same five files in TWO layouts: its CPU arm (flash-matrix.cu:88-101) takes k as [head_kv][n_kv][D] and v
    TRANSPOSED as [head_kv][D][n_kv]; its live GPU arm (flash-matrix.cu:141,149) reads both k and v as the ggml cache view
    [n_kv][head_kv][D].  kv_layout picks one: "head_major_vt" (the CPU arm's), "cache_view" (the GPU arm's), or "auto": decided from
    the shapes stored in the file headers (q f32 [head][n_q][D] fixes D and the head count; mask [rows][n_kv] fixes n_kv) — ambiguous
    or inconsistent shapes raise instead of replaying permuted data.
    q f32 [head][n_q][D] · mask f16 [rows >= n_q][n_kv] (padding rows allowed) · qkv f32 [n_q][head][D]."""
    import torch

    dev = torch.device("cuda", 0) if device is None else torch.device(device)
    paths = capture_paths(directory, tag, prefix)
    t = {p: read_tensor(paths[p])[1] for p in CAPTURE_PARTS}
    q, k, v, mask, ref = (t[p] for p in CAPTURE_PARTS)
    q = q.reshape((-1,) + q.shape[-2:]) if q.ndim > 3 else q
    if q.ndim == 2:
        q = q[:, None, :]
    H, n_q, D = q.shape
    n_kv = mask.shape[-1]
    mask = mask.reshape(-1, n_kv)
    if mask.shape[0] < n_q:
        raise api.B200FAError(-1, f"replay_capture: mask has {mask.shape[0]} rows for {n_q} queries")
    if k.size != v.size or k.size % (n_kv * D):
        raise api.B200FAError(-1, f"replay_capture: k/v hold {k.size}/{v.size} elements, not a multiple of n_kv*D = {n_kv}*{D}")
    Hk = k.size // (n_kv * D)
    if Hk < 1 or H % Hk:
        raise api.B200FAError(-1, f"replay_capture: {Hk} kv heads for {H} q heads")
    k3 = k.reshape((-1,) + k.shape[-2:]) if k.ndim >= 3 else k.reshape(-1, n_kv, D)
    v3 = v.reshape((-1,) + v.shape[-2:]) if v.ndim >= 3 else v.reshape(-1, D, n_kv)
    shapes = (tuple(k3.shape), tuple(v3.shape))
    fits = {"head_major_vt": shapes == ((Hk, n_kv, D), (Hk, D, n_kv)), "cache_view": shapes == ((n_kv, Hk, D), (n_kv, Hk, D))}
    if kv_layout == "auto":
        ok = [name for name, f in fits.items() if f]
        if len(ok) != 1:
            raise api.B200FAError(-1, f"replay_capture: k {shapes[0]} / v {shapes[1]} match {ok or 'neither layout'} for H={H}, Hk={Hk}, "
                                      f"n_kv={n_kv}, D={D}: pass kv_layout explicitly")
        kv_layout = ok[0]
    elif kv_layout not in fits:
        raise api.B200FAError(-1, f"replay_capture: unknown kv_layout {kv_layout!r}")
    elif not fits[kv_layout] and not (k3.size == Hk * n_kv * D):
        raise api.B200FAError(-1, f"replay_capture: k {shapes[0]} / v {shapes[1]} do not hold a {kv_layout} tensor")
    rows = (n_q + 31) // 32 * 32  # the kernel contract pads the mask to 32 rows (flash-matrix.cu:127, PADD)
    mpad = np.zeros((rows, n_kv), np.float16)
    mpad[: min(rows, mask.shape[0])] = mask[:rows]
    qd = torch.from_numpy(np.ascontiguousarray(q)).to(dev)[None]                                # [1][H][n_q][D]
    if kv_layout == "head_major_vt":
        kd = torch.from_numpy(np.ascontiguousarray(k.reshape(Hk, n_kv, D))).to(dev)[None]
        vd = torch.from_numpy(np.ascontiguousarray(v.reshape(Hk, D, n_kv))).to(dev).transpose(1, 2).contiguous()[None]  # un-transpose on the device
    else:  # the ggml cache view: no copy, the kernel takes the [kv][head][D] strides as they are
        kd = torch.from_numpy(np.ascontiguousarray(k.reshape(n_kv, Hk, D))).to(dev).permute(1, 0, 2)[None]
        vd = torch.from_numpy(np.ascontiguousarray(v.reshape(n_kv, Hk, D))).to(dev).permute(1, 0, 2)[None]
    md = torch.from_numpy(mpad).to(dev)
    out = api.flash_attn_ext(qd, kd, vd, md, flags=flags)
    o = out.float().cpu().numpy().reshape(n_q, H, D)
    r = ref.reshape(n_q, H, D).astype(np.float32)
    return dict(out=o, ref=r, max_abs=float(np.abs(o - r).max()), dispatch=api.last_dispatch(), kv_layout=kv_layout)
