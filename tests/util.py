"""Shared helpers for the parity tests (CPU oracle vs. the CUDA path through the C ABI)."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc


def logical_equal(a: pa.Array, b: pa.Array) -> bool:
    """Arrow logical equality (what every reference test asserts): same type, same validity, same values
    where valid; payload under nulls is ignored."""
    if isinstance(a, pa.ChunkedArray):
        a = a.combine_chunks()
    if isinstance(b, pa.ChunkedArray):
        b = b.combine_chunks()
    if a.type != b.type or len(a) != len(b):
        return False
    if pa.types.is_dictionary(a.type):
        return a.cast(a.type.value_type).equals(b.cast(b.type.value_type))
    return a.equals(b)


def assert_arrays_equal(got: pa.Array, want: pa.Array, what: str = ""):
    assert got.type == want.type, f"{what}: type {got.type} != {want.type}"
    assert len(got) == len(want), f"{what}: length {len(got)} != {len(want)}"
    if not logical_equal(got, want):
        g, w = got.to_pylist(), want.to_pylist()
        bad = [i for i, (x, y) in enumerate(zip(g, w)) if x != y][:5]
        raise AssertionError(f"{what}: mismatch at rows {bad}: got {[g[i] for i in bad]} want {[w[i] for i in bad]}")


def assert_masks_equal(got: pa.Array, want: pa.Array, what: str = ""):
    """BooleanArray equality on validity and on values where valid."""
    assert len(got) == len(want), f"{what}: length {len(got)} != {len(want)}"
    gv = np.asarray(got.is_valid().to_numpy(zero_copy_only=False), dtype=bool)
    wv = np.asarray(want.is_valid().to_numpy(zero_copy_only=False), dtype=bool)
    gb = np.asarray(got.fill_null(False).to_numpy(zero_copy_only=False), dtype=bool)
    wb = np.asarray(want.fill_null(False).to_numpy(zero_copy_only=False), dtype=bool)
    if not (np.array_equal(gv, wv) and np.array_equal(gb, wb)):
        g, w = got.to_pylist(), want.to_pylist()
        bad = [i for i, (x, y) in enumerate(zip(g, w)) if x != y][:8]
        raise AssertionError(f"{what}: mask mismatch at {bad}: got {[g[i] for i in bad]} want {[w[i] for i in bad]}")


def random_selection(rng, n, p):
    if p >= 1.0:
        return pa.array(np.ones(n, dtype=bool))
    return pa.array(rng.random(n) < p)


def assert_float_bits_equal(got: pa.Array, want: pa.Array, what: str = ""):
    """Float arrays: same type, same validity, and the SAME BITS where valid (NaN payloads and the sign of zero count;
    pyarrow's equals() would call NaN != NaN and -0.0 == +0.0)."""
    assert got.type == want.type, f"{what}: type {got.type} != {want.type}"
    assert len(got) == len(want), f"{what}: length {len(got)} != {len(want)}"
    gv = np.asarray(got.is_valid().to_numpy(zero_copy_only=False), dtype=bool)
    wv = np.asarray(want.is_valid().to_numpy(zero_copy_only=False), dtype=bool)
    assert gv.tolist() == wv.tolist(), f"{what}: validity differs"
    it = np.uint32 if got.type.bit_width == 32 else np.uint64
    g = np.asarray(got.fill_null(0).to_numpy(zero_copy_only=False)).view(it)
    w = np.asarray(want.fill_null(0).to_numpy(zero_copy_only=False)).view(it)
    bad = np.flatnonzero((g != w) & gv)[:5]
    assert len(bad) == 0, f"{what}: bits differ at rows {bad.tolist()}: got {g[bad].tolist()} want {w[bad].tolist()}"
