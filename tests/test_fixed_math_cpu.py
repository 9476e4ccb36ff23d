"""The byte arithmetic of LiquidFixedLenByteArray entries (liquid_cache_b200/csrc/fixed_math.cuh) on the CPU: the
order-preserving form must sort like the signed integers it encodes, and reading it back must give the little-endian
words again — the same source k_fixed_to_ordered / k_fixed_from_var include."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(ROOT, "build", "tests", "libfixed_math_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC", f"-I{ROOT}",
                        os.path.join(ROOT, "tests", "cpp", "fixed_math_host.cc"), "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(out)


@pytest.mark.parametrize("width", [16, 32])
def test_ordered_form_sorts_numerically_and_reads_back(lib, width):
    rng = np.random.default_rng(width)
    bits = 8 * width
    ints = [int.from_bytes(rng.bytes(width), "little", signed=True) >> int(rng.integers(0, bits - 1)) for _ in range(4000)]
    ints += [0, 1, -1, 2**64, -(2**64), 2**(bits - 1) - 1, -(2**(bits - 1)), 255, 256, -256]
    n = len(ints)
    le = np.frombuffer(b"".join(v.to_bytes(width, "little", signed=True) for v in ints), dtype=np.uint8).copy()
    stored = np.zeros(n * width, dtype=np.uint8)
    lib.fx_to_ordered(le.ctypes.data_as(C.c_void_p), n, width, stored.ctypes.data_as(C.c_void_p))
    rows = [stored[i * width:(i + 1) * width].tobytes() for i in range(n)]
    # big-endian with the sign bit flipped: v + 2^(bits-1) as an unsigned big-endian integer
    assert rows == [(v + (1 << (bits - 1))).to_bytes(width, "big") for v in ints]
    order = sorted(range(n), key=lambda i: rows[i])            # unsigned lexicographic, the byte-view comparison order
    assert [ints[i] for i in order] == sorted(ints)
    back = np.zeros(n * width // 4, dtype=np.uint32)
    lib.fx_from_ordered(stored.ctypes.data_as(C.c_void_p), n, width, back.ctypes.data_as(C.c_void_p))
    assert back.tobytes() == le.tobytes()


@pytest.mark.parametrize("width", [16, 32])
def test_literal_needles(lib, width):
    """the needle of a comparison: LC_LIT_I128 halves sign-extended to the column's width, or the column's own bytes"""
    rng = np.random.default_rng(width + 1)
    bits = 8 * width
    lib.fx_needle.argtypes = [C.c_uint64, C.c_int64, C.c_void_p, C.c_uint32, C.c_void_p]
    for _ in range(2000):
        v = int.from_bytes(rng.bytes(16), "little", signed=True) >> int(rng.integers(0, 127))
        out = (C.c_uint8 * width)()
        lib.fx_needle(v & (2**64 - 1), v >> 64, None, width, out)
        assert bytes(out) == (v + (1 << (bits - 1))).to_bytes(width, "big"), v
        w = int.from_bytes(rng.bytes(width), "little", signed=True) >> int(rng.integers(0, bits - 1))
        le = w.to_bytes(width, "little", signed=True)
        lib.fx_needle(0, 0, le, width, out)
        assert bytes(out) == (w + (1 << (bits - 1))).to_bytes(width, "big"), w
