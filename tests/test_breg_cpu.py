"""The register-resident integer scan on the CPU: csrc/breg_math.cuh — which words of a FastLanes chunk each thread of the
warp loads, the W-bit field it cuts out of them at every step, and the mask word each step's ballot is — against a plain
FastLanes unpack (oracle/liquid_oracle.py fl_unpack_chunk, the restatement of fastlanes 0.5.0's unified transposed order)
for every width of 8-, 16-, 32- and 64-bit columns the kernel is instantiated for. Bit b of mask word out_word(step) must be
the row 32 * out_word(step) + b."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import liquid_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(ROOT, "build", "tests", "libbreg_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC", f"-I{ROOT}",
                        os.path.join(ROOT, "tests", "cpp", "breg_host.cc"), "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(out)


@pytest.mark.parametrize("tbits", [8, 16, 32, 64])
def test_every_width_reads_the_rows_fastlanes_stored(lib, tbits):
    U = {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}[tbits]
    rng = np.random.default_rng(tbits)
    for width in range(1, min(32, tbits) + 1):
        vals = rng.integers(0, 1 << width, size=1024, dtype=np.uint64).astype(U)
        packed = O.fl_pack_chunk(vals, width)
        assert np.array_equal(O.fl_unpack_chunk(packed, width), vals)
        raw = np.frombuffer(packed.tobytes(), dtype=np.uint8).copy()
        assert len(raw) == 128 * width
        values = np.zeros((32, 32), dtype=np.uint32)
        out_word = np.zeros(32, dtype=np.uint32)
        assert lib.br_chunk(tbits, width, raw.ctypes.data_as(C.c_void_p), values.ctypes.data_as(C.c_void_p),
                            out_word.ctypes.data_as(C.c_void_p)) == 0
        assert sorted(out_word.tolist()) == list(range(32))
        for step in range(32):
            rows = 32 * int(out_word[step]) + np.arange(32)
            assert np.array_equal(values[step].astype(np.uint64), vals[rows].astype(np.uint64)), (tbits, width, step)
