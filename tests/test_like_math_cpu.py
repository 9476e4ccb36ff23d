"""LIKE '%needle%' evaluated on FSST codes (csrc/like_math.cuh) on the CPU: the per-symbol Shift-And step table, the
code-by-code walk and the 32-codes-at-once walk with composed steps — the arithmetic k_like_steps / like_trip /
like_candidates_warp run on the device — must all say what a plain substring search of the decoded bytes says
(comparisons.rs:325-347: `contains` on the decompressed value), for random symbol tables, escaped bytes, needles of 1..31
bytes and matches that straddle symbols and 32-code blocks."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class SymStep(C.Structure):
    _fields_ = [("A", C.c_uint32), ("B", C.c_uint32), ("H", C.c_uint32), ("L", C.c_uint32)]


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(ROOT, "build", "tests", "liblike_math_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC", f"-I{ROOT}",
                        os.path.join(ROOT, "tests", "cpp", "like_math_host.cc"), "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(out)


def random_table(rng, alphabet):
    """A symbol table in the shape of fsst-rs's: up to 255 symbols of 1..8 bytes over `alphabet`."""
    n = int(rng.integers(8, 255))
    syms = []
    for _ in range(n):
        L = int(rng.choice([1, 1, 2, 2, 3, 4, 5, 8]))
        syms.append(bytes(int(alphabet[i]) for i in rng.integers(0, len(alphabet), size=L)))
    return syms


def compress(syms, text: bytes) -> bytes:
    """Greedy longest match; bytes no symbol starts with are escaped (255, byte) — any valid code stream will do here."""
    by_first = {}
    for c, s in enumerate(syms):
        by_first.setdefault(s[0], []).append((len(s), c, s))
    for v in by_first.values():
        v.sort(reverse=True)
    out, i = bytearray(), 0
    while i < len(text):
        for L, c, s in by_first.get(text[i], []):
            if text[i:i + L] == s:
                out.append(c)
                i += L
                break
        else:
            out += bytes([255, text[i]])
            i += 1
    return bytes(out)


def decompress(syms, codes: bytes) -> bytes:
    out, i = bytearray(), 0
    while i < len(codes):
        if codes[i] == 255:
            out.append(codes[i + 1])
            i += 2
        else:
            out += syms[codes[i]]
            i += 1
    return bytes(out)


def test_step_table_and_both_walks_agree_with_substring_search(lib):
    rng = np.random.default_rng(2024)
    alphabet = np.frombuffer(b"abcdeghgo/.:%-_?=&xyz01", dtype=np.uint8)  # few letters: needles do occur by chance
    rare = np.array([0xC3, 0xA9, 0xFF, 0x00, 0x7F], dtype=np.uint8)       # bytes that end up escaped (0xFF among them)
    checked = hits = 0
    for _ in range(60):
        syms = random_table(rng, alphabet)
        symbols = np.zeros(256, dtype=np.uint64)
        lens = np.zeros(256, dtype=np.uint8)
        for c, s in enumerate(syms):
            symbols[c] = int.from_bytes(s, "little")
            lens[c] = len(s)
        for _ in range(40):
            m = int(rng.integers(1, 32))
            n_text = int(rng.integers(0, 400))
            text = bytearray(int(alphabet[i]) for i in rng.integers(0, len(alphabet), size=n_text))
            for _k in range(int(rng.integers(0, 6))):
                if text:
                    text[int(rng.integers(0, len(text)))] = int(rng.choice(rare))
            needle = bytes(int(alphabet[i]) for i in rng.integers(0, len(alphabet), size=m))
            if rng.random() < 0.5 and len(text) >= m:  # plant it, anywhere (across symbol and block boundaries too)
                at = int(rng.integers(0, len(text) - m + 1))
                text[at:at + m] = needle
            if rng.random() < 0.2:
                needle = needle[:-1] + bytes([int(rng.choice(rare))])  # a needle with a byte that only occurs escaped
            text = bytes(text)
            codes = compress(syms, text)
            assert decompress(syms, codes) == text
            steps = (SymStep * 512)()
            lib.lm_table(symbols.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p), needle, len(needle), steps)
            want = int(needle in text)
            assert lib.lm_walk_seq(steps, codes, len(codes)) == want, (needle, text)
            assert lib.lm_walk_blocks(steps, codes, len(codes)) == want, (needle, text)
            checked += 1
            hits += want
    assert checked == 2400 and 300 < hits < 2100


def test_composition_is_associative_and_has_an_identity(lib):
    """step_then is what the shuffle tree relies on: any bracketing of a run of steps is the same step."""
    rng = np.random.default_rng(7)
    alphabet = np.frombuffer(b"abcab/", dtype=np.uint8)
    syms = random_table(rng, alphabet)
    symbols = np.zeros(256, dtype=np.uint64)
    lens = np.zeros(256, dtype=np.uint8)
    for c, s in enumerate(syms):
        symbols[c] = int.from_bytes(s, "little")
        lens[c] = len(s)
    needle = b"abcab"
    steps = (SymStep * 512)()
    lib.lm_table(symbols.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p), needle, len(needle), steps)
    # the block walk pads the last block with identity steps and splits every value at multiples of 32 codes: values of every
    # length 0..130 exercise all tree shapes; both walks must agree on each prefix of a long code stream
    text = bytes(int(alphabet[i]) for i in rng.integers(0, len(alphabet), size=600))
    codes = compress(syms, text)
    for n in range(0, min(len(codes), 131)):
        assert lib.lm_walk_blocks(steps, codes, n) == lib.lm_walk_seq(steps, codes, n), n
