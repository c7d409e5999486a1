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


def replay_capture(directory: str, tag: str = "256", prefix: str = "fa-cuda", device=None, flags: int = 0):
    """Replay one capture set (flash-matrix.cu:66-73) on the GPU.  Layouts as test_llama reads them (flash-matrix.cu:88-165):
    q f32 [head][n_q][D] · k f16 [head_kv][n_kv][D] · v f16 TRANSPOSED [head_kv][D][n_kv] · mask f16 [rows][n_kv] (first n_q rows
    are used, padding allowed) · qkv f32 [n_q][head][D].  -> dict(out=, ref=, max_abs=, dispatch=)."""
    import torch

    dev = torch.device("cuda", 0) if device is None else torch.device(device)
    paths = capture_paths(directory, tag, prefix)
    t = {p: read_tensor(paths[p])[1] for p in CAPTURE_PARTS}
    q, k, vt, mask, ref = (t[p] for p in CAPTURE_PARTS)
    q = q.reshape((-1,) + q.shape[-2:]) if q.ndim > 3 else q
    if q.ndim == 2:
        q = q[:, None, :]
    H, n_q, D = q.shape
    k = k.reshape(-1, k.shape[-2], k.shape[-1])
    Hk, n_kv, _ = k.shape
    vt = vt.reshape(Hk, D, n_kv)
    mask = mask.reshape(-1, n_kv)
    if mask.shape[0] < n_q:
        raise api.B200FAError(-1, f"replay_capture: mask has {mask.shape[0]} rows for {n_q} queries")
    rows = (n_q + 31) // 32 * 32  # the kernel contract pads the mask to 32 rows (flash-matrix.cu:127, PADD)
    mpad = np.zeros((rows, n_kv), np.float16)
    mpad[: min(rows, mask.shape[0])] = mask[:rows]
    qd = torch.from_numpy(np.ascontiguousarray(q)).to(dev)[None]                                # [1][H][n_q][D]
    kd = torch.from_numpy(np.ascontiguousarray(k)).to(dev)[None]
    vd = torch.from_numpy(np.ascontiguousarray(vt)).to(dev).transpose(1, 2).contiguous()[None]  # un-transpose on the device
    md = torch.from_numpy(mpad).to(dev)
    out = api.flash_attn_ext(qd, kd, vd, md, flags=flags)
    o = out.float().cpu().numpy().reshape(n_q, H, D)
    r = ref.reshape(n_q, H, D).astype(np.float32)
    return dict(out=o, ref=r, max_abs=float(np.abs(o - r).max()), dispatch=api.last_dispatch())
