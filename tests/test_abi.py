"""CPU tests of the C-ABI boundary: the library loads, exports every symbol the headers
declare, the GPU-free host helpers (varint, Buffer) behave like the reference's, and a
compute call without a GPU fails loudly instead of falling back."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

from lightweight_snappy_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(api.LIB_PATH):
        api.build()
    return api.lib()


def _declared_functions():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text):
            names.add(m.group(1))
    return names


def test_every_declared_symbol_is_exported(L):
    declared = _declared_functions()
    assert {"snappy_compress", "snappy_compress_bst", "snappy_decompress", "parse_to_varint", "varint_to_dim",
            "str_varint_to_dim_", "init_Buffer", "move_current", "reset", "snappy_b200_compress_device",
            "snappy_b200_decompress_device_indexed", "snappy_b200_index_device", "snappy_b200_compress_host",
            "snappy_b200_decompress_host"} <= declared
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/ but not exported"
    assert declared == set(api.SIGNATURES), "api.SIGNATURES and include/*.h disagree"


def test_varint_known_answers(L):
    # the reference's own vectors: src/test_varint.c:27-42
    buf = (C.c_ubyte * 16)()
    for v, want in ((127, b"\x7f"), (227, b"\xe3\x01"), (16384, b"\x80\x80\x01"), (1 << 30, b"\x80\x80\x80\x80\x04"),
                    ((1 << 31) - 1, b"\xff\xff\xff\xff\x07"), (0, b"\x00")):
        n = L.parse_to_varint(v, C.addressof(buf))
        assert bytes(buf[:n]) == want
        assert L.str_varint_to_dim_(C.addressof(buf)) == v
    # values past 2^31-1 wrap like the reference's int accumulator
    n = L.parse_to_varint(1 << 31, C.addressof(buf))
    assert L.str_varint_to_dim_(C.addressof(buf)) == -(1 << 31)


def test_varint_to_dim_reads_file(L, tmp_path):
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    p = tmp_path / "v.bin"
    p.write_bytes(b"\xca\x02rest")  # 330, Presentazione snappy 1/formatSnappyEx.pdf
    f = libc.fopen(str(p).encode(), b"rb")
    assert L.varint_to_dim(f) == 330
    libc.fclose(f)


def test_buffer_cursor_semantics(L):
    class Buffer(C.Structure):
        _fields_ = [("current", C.c_void_p), ("beginning", C.c_void_p), ("bytes_left", C.c_uint)]

    b = Buffer()
    L.init_Buffer(C.byref(b), 100)
    assert b.current == b.beginning and b.bytes_left == 100
    assert bytes((C.c_ubyte * 100).from_address(b.beginning)) == bytes(100)  # calloc'd
    L.move_current(C.byref(b), 30)
    assert b.current == b.beginning + 30 and b.bytes_left == 70
    L.reset(C.byref(b))
    assert b.current == b.beginning and b.bytes_left == 70  # reset keeps bytes_left (reference quirk)
    C.CDLL(None).free(C.c_void_p(b.beginning))


def test_size_helpers(L):
    assert L.snappy_b200_max_compressed_bytes(0) == 0
    assert L.snappy_b200_max_compressed_bytes(65536) == 10 + 65536 + 1010
    assert L.snappy_b200_block_count(65537) == 2
    assert L.snappy_b200_compress_workspace_bytes(1 << 20, 0) >= 16 * api.SLOT_STRIDE


def test_no_gpu_means_loud_failure(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.SnappyError):
        api.snappy_compress(np.arange(1000, dtype=np.uint8))
    with pytest.raises(api.SnappyError):
        api.snappy_decompress(np.frombuffer(b"\x03\x08abc", dtype=np.uint8))
    # the empty stream needs no device (reference: empty input -> empty output)
    assert api.snappy_compress(np.zeros(0, dtype=np.uint8)).size == 0
