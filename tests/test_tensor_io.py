"""CPU: ggml tensor-dump reader/writer (b200fa_tensor_file_*; reference loader utils.h:110-150) — byte layout, round trips,
malformed files, the committed golden dump, and (where oracle/_ref is built) the reference's own loader reading our files."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

import oracle
from __graft_entry__ import load_package as pkg

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "fa-cuda-sample-8.tensor")


def _blob(name, arr):
    ne = arr.shape[::-1]
    return struct.pack(f"<ii{len(ne)}ii", arr.ndim, 1 if arr.dtype == np.float16 else 0, *ne, len(name)) + name.encode() + arr.tobytes()


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
@pytest.mark.parametrize("shape", [(7,), (3, 5), (2, 3, 4), (2, 1, 3, 8)])
def test_round_trip_and_layout(tmp_path, dtype, shape):
    P = pkg()
    a = np.random.RandomState(len(shape)).uniform(-1, 1, shape).astype(dtype)
    p = str(tmp_path / "t.tensor")
    P.write_tensor(p, "kqv_out-0", a)
    assert open(p, "rb").read() == _blob("kqv_out-0", a)          # the exact bytes of the format
    name, b = P.read_tensor(p)
    assert name == "kqv_out-0" and b.dtype == a.dtype and b.shape == a.shape and np.array_equal(a, b)
    info = P.tensor_info(p)
    assert info.n_dims == len(shape) and list(info.ne)[: len(shape)] == list(shape[::-1]) and info.data_bytes == a.nbytes
    assert info.data_offset == 4 * (3 + len(shape)) + len("kqv_out-0")


def test_reads_foreign_file(tmp_path):
    """A file produced outside the library (hand-packed) parses to the same tensor."""
    P = pkg()
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    p = tmp_path / "f.tensor"
    p.write_bytes(_blob("q", a))
    name, b = P.read_tensor(str(p))
    assert name == "q" and np.array_equal(a, b)


def test_golden_dump():
    P = pkg()
    name, a = P.read_tensor(GOLDEN)
    assert name == "sample" and a.shape == (2, 8) and a.dtype == np.float16
    assert np.array_equal(a, (np.arange(16, dtype=np.float32) / 8 - 1).astype(np.float16).reshape(2, 8))


def test_malformed(tmp_path):
    P = pkg()
    a = np.ones((4, 4), np.float32)
    good = _blob("x", a)
    cases = {
        "truncated_payload": good[:-5],
        "truncated_header": good[:6],
        "zero_dims": struct.pack("<ii", 0, 0),
        "five_dims": struct.pack("<ii5i", 5, 0, 1, 1, 1, 1, 1),
        "bad_type": struct.pack("<iiii", 1, 7, 4, 1) + b"x" + bytes(16),
        "negative_ne": struct.pack("<iiii", 1, 0, -4, 1) + b"x",
        "huge_name": struct.pack("<iiii", 1, 0, 4, 4000) + b"x" * 100,
    }
    for what, blob in cases.items():
        p = tmp_path / (what + ".tensor")
        p.write_bytes(blob)
        with pytest.raises(P.B200FAError):
            P.read_tensor(str(p))
    with pytest.raises(P.B200FAError):
        P.read_tensor(str(tmp_path / "does_not_exist.tensor"))
    with pytest.raises(P.B200FAError):
        P.write_tensor(str(tmp_path / "n.tensor"), "a_name_longer_than_the_reference_field", a)
    with pytest.raises(P.B200FAError):
        P.write_tensor(str(tmp_path / "n.tensor"), "i", np.ones(4, np.int32))
    with pytest.raises(P.B200FAError):
        P.write_tensor(str(tmp_path / "no_such_dir" / "n.tensor"), "x", a)


def test_status_string_io():
    l = pkg().lib()
    l.b200fa_status_string.restype = C.c_char_p
    assert b"tensor file" in l.b200fa_status_string(-5)


@pytest.mark.skipif(not oracle.ref_host_available(), reason="oracle/_ref not built and /root/reference absent")
@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_reference_loader_reads_our_files(tmp_path, dtype):
    """utils.h:110-150 (load_tensor_from_file) on a file written by b200fa_tensor_file_write: same name, same payload."""
    r = oracle.ref_host()
    if not hasattr(r, "ref_host_load_tensor"):
        pytest.skip("prebuilt oracle/_ref predates the loader shim")
    P = pkg()
    a = np.random.RandomState(5).uniform(-1, 1, (3, 4, 16)).astype(dtype)
    p = str(tmp_path / "r.tensor")
    P.write_tensor(p, "fa-cuda-q-256", a)
    out = np.empty_like(a)
    name = C.create_string_buffer(20)
    ty = C.c_int(-1)
    assert r.ref_host_load_tensor(p.encode(), C.byref(ty), name, out.ctypes.data, out.nbytes) == 0
    assert name.value == b"fa-cuda-q-256" and np.array_equal(out, a)
