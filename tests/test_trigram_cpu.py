"""The 256-bit trigram pre-filter (csrc/entry_layout.h trigram_bit) on the CPU: the Python twin the GPU layout test
compares HBM images with (tests/test_gpu_insert_layout.py trigram_bloom) is the header's function, and the filter can never
reject a value that contains the needle — the property that makes skipping on it safe."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(ROOT, "build", "tests", "libtrigram_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC", f"-I{ROOT}",
                        os.path.join(ROOT, "tests", "cpp", "trigram_host.cc"), "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(out)


def bloom(lib, b: bytes):
    out = (C.c_uint64 * 4)()
    lib.tg_bloom(b, len(b), out)
    return [int(x) for x in out]


def python_twin(b: bytes):
    x = 0
    for a, c, d in zip(b, b[1:], b[2:]):
        x |= 1 << ((((a << 16) | (c << 8) | d) * 0x9E3779B1 & 0xFFFFFFFF) >> 24)
    return [(x >> (64 * w)) & (2**64 - 1) for w in range(4)]


def test_header_function_is_the_python_twin_and_never_rejects_a_match(lib):
    rng = np.random.default_rng(8)
    alphabet = b"abcdefghijklmnopqrstuvwxyz0123456789/.:%-_?=&" + bytes(range(200, 256))
    for _ in range(3000):
        n = int(rng.integers(0, 120))
        value = bytes(alphabet[int(i)] for i in rng.integers(0, len(alphabet), size=n))
        vb = bloom(lib, value)
        assert vb == python_twin(value)
        if n >= 1:
            a = int(rng.integers(0, n))
            needle = value[a:a + int(rng.integers(1, 12))]
            nb = bloom(lib, needle)
            assert all((v & m) == m for v, m in zip(vb, nb)), (value, needle)
    assert bloom(lib, b"") == [0, 0, 0, 0] and bloom(lib, b"ab") == [0, 0, 0, 0]  # no trigram below three bytes
