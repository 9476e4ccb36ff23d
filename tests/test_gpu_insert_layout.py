"""GPU parity of the INSERT side: the sections the device builds vs. the CPU oracle's encode.

What the reference fixes and the oracle restates (so it must match bit for bit):
  u16 keys in first-occurrence order, null rows                utils/mod.rs:52-154
  shared prefix, PrefixKey bytes, fingerprints                 byte_view_array/conversions.rs:260-373, fingerprint.rs:19-26
  CompactOffsets (line fit in f64, residual width)             raw/fsst_buffer.rs:267-359
  frame of reference, bit width, FastLanes packed words        primitive_array.rs:159-206, raw/bit_pack_array.rs:71-124
What it leaves open (third-party, only round-tripped by its tests): the FSST symbol table and therefore the compressed
bytes. Those are checked by decompressing every dictionary value with the entry's own table.

The entry image is read back through lc_entry_image and parsed with the layout of csrc/entry_layout.h.
"""
import struct

import numpy as np
import pyarrow as pa
import pytest

from oracle.liquid_oracle import CompactOffsets, OracleByteViewArray, OracleIntArray
from tests.test_gpu_str import make_strings

pytestmark = pytest.mark.gpu

STR_HDR = struct.Struct("<I4B I I i i I 11I Q Q 4I")  # up to rows_off; 128-byte header, tail is padding
INT_HDR = struct.Struct("<I4B I I Q 6I")


def parse_str_image(img: bytes):
    (magic, arrow_type, has_nulls, has_fp, offset_bytes, n, n_unique, slope, intercept, spl, validity_off, keys_off,
     prefix_keys_off, fp_off, resid_off, shared_prefix_off, fsst_off, fsst_bytes, blob_bytes, null_count, max_value_len,
     uncompressed, table_ptr, head_bytes, sp_end, rows_off, bloom_off) = STR_HDR.unpack_from(img, 0)
    assert magic == 0x3153514C and blob_bytes == len(img)
    h = dict(n=n, n_unique=n_unique, slope=slope, intercept=intercept, null_count=null_count, has_fp=has_fp,
             offset_bytes=offset_bytes, max_value_len=max_value_len, uncompressed=uncompressed, has_nulls=has_nulls)
    h["shared_prefix"] = img[shared_prefix_off:shared_prefix_off + spl]
    h["keys"] = np.frombuffer(img, dtype=np.uint16, count=n, offset=keys_off)
    h["valid"] = (np.unpackbits(np.frombuffer(img, dtype=np.uint8, count=(n + 7) // 8, offset=validity_off),
                                bitorder="little")[:n].astype(bool) if has_nulls else np.ones(n, dtype=bool))
    pk = np.frombuffer(img, dtype=np.uint8, count=8 * n_unique, offset=prefix_keys_off).reshape(n_unique, 8)
    h["prefix_keys"] = [(bytes(r[:7]), int(r[7])) for r in pk]
    h["fps"] = np.frombuffer(img, dtype=np.uint32, count=n_unique, offset=fp_off).tolist() if has_fp else None
    dt = {1: np.int8, 2: np.int16, 4: np.int32}[offset_bytes]
    h["resid"] = np.frombuffer(img, dtype=dt, count=n_unique + 1, offset=resid_off).astype(np.int64)
    h["comp"] = img[fsst_off:fsst_off + fsst_bytes]
    h["blooms"] = None
    if bloom_off:
        # entry_layout.h: the trigram sets are stored as 256 PLANES over the dictionary (plane t, bit i = value i has trigram
        # bit t), ceil(U / 32) words per plane; read back here as one 256-bit row per value, four u64 words each
        pw = (n_unique + 31) // 32
        planes = np.frombuffer(img, dtype="<u4", count=256 * pw, offset=bloom_off).reshape(256, pw)
        bits = np.unpackbits(planes.view(np.uint8).reshape(256, pw * 4), axis=1, bitorder="little")  # [plane, value]
        assert not bits[:, n_unique:].any(), "plane padding beyond the dictionary must be zero"
        rows = np.packbits(bits[:, :n_unique].T, axis=1, bitorder="little")                           # [value, 32 bytes]
        h["blooms"] = rows.view("<u8").reshape(-1).tolist()
    return h


def trigram_bloom(b: bytes) -> list:
    """entry_layout.h trigram_bit: the private 256-bit substring pre-filter stored beside the reference fingerprints,
    as its four u64 words."""
    x = 0
    for a, c, d in zip(b, b[1:], b[2:]):
        x |= 1 << ((((a << 16) | (c << 8) | d) * 0x9E3779B1 & 0xFFFFFFFF) >> 24)
    return [(x >> (64 * w)) & (2**64 - 1) for w in range(4)]


def fsst_decompress(table: bytes, comp: bytes) -> bytes:
    syms = np.frombuffer(table, dtype="<u8", count=256, offset=0)
    lens = np.frombuffer(table, dtype=np.uint8, count=256, offset=2048)
    out = bytearray()
    i = 0
    while i < len(comp):
        c = comp[i]
        if c == 255:
            out.append(comp[i + 1])
            i += 2
        else:
            out += int(syms[c]).to_bytes(8, "little")[: int(lens[c])]
            i += 1
    return bytes(out)


def check_string_entry(cache, arr: pa.Array, hint=None, scope=0):
    la = cache.transcode(arr, hint=hint, compressor_scope=scope)
    h = parse_str_image(la.entry_image())
    want = OracleByteViewArray.from_arrow(arr, build_fingerprints=hint is not None)
    n = len(arr)
    assert h["n"] == n and h["n_unique"] == len(want.uniques)
    want_valid = np.array([k is not None for k in want.keys], dtype=bool)
    assert h["null_count"] == int((~want_valid).sum())
    assert np.array_equal(h["valid"], want_valid)
    want_keys = np.array([k if k is not None else 0 for k in want.keys], dtype=np.uint16)
    assert np.array_equal(h["keys"][want_valid], want_keys[want_valid]), "dictionary keys / first-occurrence order"
    assert h["shared_prefix"] == want.shared_prefix
    assert h["prefix_keys"] == want.prefix_keys
    if hint is not None:
        assert h["fps"] == want.fingerprints
        if want.uniques:
            assert h["blooms"] == [w for u in want.uniques for w in trigram_bloom(u)]
    else:
        assert h["fps"] is None and h["blooms"] is None
    assert h["max_value_len"] == max([len(u) for u in want.uniques], default=0)
    assert h["uncompressed"] == sum(len(u) for u in want.uniques)
    # offsets: slope * i + intercept + resid[i], rebuilt and checked against the compressed stream
    U = h["n_unique"]
    offs = [(h["slope"] * i + h["intercept"] + int(h["resid"][i])) & 0xFFFFFFFF for i in range(U + 1)]
    assert offs[0] == 0 and offs[-1] == len(h["comp"]) and all(a <= b for a, b in zip(offs, offs[1:]))
    table = la.fsst_table()
    for u in range(U):
        assert fsst_decompress(table, h["comp"][offs[u]:offs[u + 1]]) == want.uniques[u], f"unique {u}"
    # CompactOffsets of THESE offsets, as the oracle (reference order of f64 operations) computes them
    co = CompactOffsets.from_offsets(offs)
    assert (h["slope"], h["intercept"]) == (co.slope, co.intercept)
    assert [int(x) for x in h["resid"]] == [int(x) for x in co.residuals]
    assert h["offset_bytes"] == co.offset_bytes
    return la


@pytest.mark.parametrize("typ", [pa.string(), pa.binary(), pa.string_view(), pa.binary_view()])
@pytest.mark.parametrize("seed,n,n_unique,null_p", [(1, 8192, 1900, 0.1), (2, 5000, 37, 0.0), (3, 777, 777, 0.3), (4, 8192, 6000, 0.02)])
def test_string_insert_sections(cache, typ, seed, n, n_unique, null_p):
    rng = np.random.default_rng(seed)
    vals, mask = make_strings(rng, n, n_unique, null_p=null_p, prefix="http://" if seed % 2 else "")
    data = [None if (mask is not None and mask[i]) else (v if typ in (pa.string(), pa.string_view()) else v.encode())
            for i, v in enumerate(vals)]
    arr = pa.array(data, type=typ)
    from liquid_cache_b200 import CacheExpression

    check_string_entry(cache, arr, hint=CacheExpression.SubstringSearch if seed % 2 else None, scope=9000 + seed)


def test_string_insert_dictionary_input(cache):
    rng = np.random.default_rng(11)
    values = pa.array(["", "alpha", "alphabet", None, "beta", "alp"])
    keys = pa.array(rng.integers(0, 6, size=4000).astype(np.uint16), mask=rng.random(4000) < 0.1)
    arr = pa.DictionaryArray.from_arrays(keys, values)
    check_string_entry(cache, arr, scope=9100)


def test_string_insert_edge_cases(cache):
    from liquid_cache_b200 import CacheExpression

    hint = CacheExpression.SubstringSearch
    check_string_entry(cache, pa.array([], type=pa.string()), scope=9200)
    check_string_entry(cache, pa.array([None, None, None], type=pa.string()), hint=hint, scope=9201)
    check_string_entry(cache, pa.array(["same"] * 3000), hint=hint, scope=9202)
    check_string_entry(cache, pa.array(["", "", None, ""]), scope=9203)
    long = [str(i) + "x" * 300 for i in range(1200)]  # no shared prefix, suffix length >= 255 -> len byte 255
    check_string_entry(cache, pa.array(long + long[::-1]), hint=hint, scope=9204)
    check_string_entry(cache, pa.array(["prefix_only", "prefix_only_longer", "prefix_only"]), scope=9205)
    # sliced input with a non-zero offset and a validity bitmap that is not byte aligned
    base = pa.array([None if i % 5 == 0 else f"row{i % 300}" for i in range(4100)])
    check_string_entry(cache, base.slice(3, 4001), hint=hint, scope=9206)


def test_string_insert_rejects_more_than_65536_uniques(cache):
    from liquid_cache_b200._native import UnsupportedType

    arr = pa.array([f"v{i}" for i in range(65537)])
    with pytest.raises(UnsupportedType):
        cache.transcode(arr, compressor_scope=9300)
    ok = pa.array([f"v{i}" for i in range(65536)])
    la = cache.transcode(ok, compressor_scope=9301)
    assert la.to_arrow_array().equals(ok)


@pytest.mark.parametrize("np_dt,lo,hi,n", [(np.int64, -5000, 10**12, 8192), (np.uint64, 10, 16, 6), (np.int32, -7, 8, 3000),
                                           (np.int16, -300, 300, 1024), (np.uint8, 0, 256, 2500), (np.int64, 77, 78, 1500),
                                           (np.uint16, 0, 65536, 5000)])  # W = T = 16: how byte-view LQDA packs its dictionary keys
def test_int_insert_layout_matches_oracle_packing(cache, np_dt, lo, hi, n):
    """Reference value, bit width and the FastLanes words themselves (fastlanes 0.5.0 unified transposed order as the
    oracle restates it) — the device packer and the oracle must agree word for word when there are no nulls."""
    rng = np.random.default_rng(n)
    arr = pa.array(rng.integers(lo, hi, size=n, dtype=np_dt))
    la = cache.transcode(arr)
    img = la.entry_image()
    magic, phys, tbits, bit_width, has_nulls, nn, n_chunks, reference, validity_off, packed_off, blob_bytes, null_count, \
        is_signed, _ = INT_HDR.unpack_from(img, 0)
    assert magic == 0x3149514C and nn == n and has_nulls == 0 and null_count == 0
    want = OracleIntArray.from_arrow(arr)
    assert bit_width == want.bit_width and tbits == np.dtype(np_dt).itemsize * 8
    assert reference == want.reference & ((1 << tbits) - 1)  # raw bits of the native type, zero extended
    words = np.frombuffer(img, dtype=want.packed.dtype, count=len(want.packed), offset=packed_off)
    assert np.array_equal(words, want.packed)
