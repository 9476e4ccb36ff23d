"""The byte-view LQDA layout code of the library (csrc/lqda_layout.h) on the CPU, against images the oracle writes
(oracle/liquid_oracle.py byte_view_to_bytes, the restatement of byte_view_array/serialization.rs:87-325): the reader must
find every section where the oracle put it and refuse damaged images; the writer's offsets for the same entry must be the
reader's — so what the device code still has to get right is only the copies between those offsets and HBM."""
import ctypes as C
import os
import subprocess

import numpy as np
import pyarrow as pa
import pytest

from oracle import liquid_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TYPES = [pa.string(), pa.binary(), pa.string_view(), pa.binary_view(), pa.dictionary(pa.uint16(), pa.string()), pa.dictionary(pa.uint16(), pa.binary())]
FIELDS = ["bt", "n", "U", "ob", "n_resid", "sp_size", "fp_size", "comp_bytes", "nulls_len", "kvals_len", "slope", "intercept", "file_nulls", "uncompressed",
          "comp_off", "knulls_off", "kvals_off", "resid_src", "pk_src", "sp_src", "fp_src"]
LAYOUT = ["fsst_off", "keys_off", "keys_nulls_off", "keys_values_off", "co_off", "pk_off", "sp_off", "fp_off", "total", "fsst_raw_size", "keys_size", "nulls_len",
          "keys_values_len", "co_size", "sp_size", "fp_size"]


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(ROOT, "build", "tests", "liblqda_layout_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC", f"-I{ROOT}",
                        os.path.join(ROOT, "tests", "cpp", "lqda_layout_host.cc"), "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(out)


def parse(lib, image: bytes):
    out = (C.c_int64 * 21)()
    why = C.c_char_p()
    rc = lib.lq_parse(image, C.c_uint64(len(image)), out, C.byref(why))
    return (dict(zip(FIELDS, [int(x) for x in out])), None) if rc == 0 else (None, why.value.decode())


def _build(vals, typ):
    if pa.types.is_dictionary(typ):
        text = pa.types.is_string(typ.value_type)
        return pa.array([None if v is None else (v if text else v.encode()) for v in vals], typ.value_type).dictionary_encode().cast(typ)
    text = pa.types.is_string(typ) or pa.types.is_string_view(typ)
    return pa.array([None if v is None else (v if text else v.encode()) for v in vals], typ)


def _cases():
    rng = np.random.default_rng(77)
    urls = [f"http://host{int(i)}.example.com/{'google' if i % 13 == 0 else 'page'}/{int(i) * 7919 % 1000}" for i in rng.integers(0, 900, 6000)]
    return [["hello_world", "hello_rust", None, "hello_test", "hello_world"], [], [None, None, None], ["", "", None, ""], ["only"], urls,
            [None if rng.random() < 0.1 else u for u in urls[:3000]], [f"{i:05d}" for i in range(2500)], ["x" * 300 + str(i % 40) for i in range(1500)]]


@pytest.mark.parametrize("typ", TYPES, ids=str)
@pytest.mark.parametrize("fp", [False, True], ids=["plain", "fingerprints"])
def test_reader_finds_the_oracles_sections_and_writer_agrees(lib, typ, fp):
    for ci, vals in enumerate(_cases()):
        arr = _build(vals, typ)
        o = O.OracleByteViewArray.from_arrow(arr, build_fingerprints=fp)
        image = O.byte_view_to_bytes(o)
        got, why = parse(lib, image)
        assert got is not None, (ci, why)
        U, n = len(o.uniques), len(arr)
        has_nulls = any(k is None for k in o.keys)
        assert (got["bt"], got["n"], got["U"]) == (O._arrow_byte_type_id(arr.type), n, U)
        assert (got["ob"], got["slope"], got["intercept"], got["n_resid"]) == (o.offsets.offset_bytes, o.offsets.slope, o.offsets.intercept, len(o.offsets.residuals))
        assert got["file_nulls"] == int(has_nulls) and got["uncompressed"] == sum(len(u) for u in o.uniques)
        assert image[got["comp_off"]:got["comp_off"] + got["comp_bytes"]] == o.compressed
        assert image[got["sp_src"]:got["sp_src"] + got["sp_size"]] == o.shared_prefix
        assert image[got["pk_src"]:got["pk_src"] + 8 * U] == b"".join(bytes(p7) + bytes([ln]) for p7, ln in o.prefix_keys)
        fps = o.fingerprints or []
        assert got["fp_size"] == 4 * len(fps) and image[got["fp_src"]:got["fp_src"] + got["fp_size"]] == b"".join(int(f).to_bytes(4, "little") for f in fps)
        ob = got["ob"]
        res = [int.from_bytes(image[got["resid_src"] + i * ob:got["resid_src"] + (i + 1) * ob], "little", signed=True) for i in range(got["n_resid"])]
        assert res == [int(r) for r in o.offsets.residuals]
        keys = np.frombuffer(image, dtype=np.uint16, count=got["kvals_len"] // 2, offset=got["kvals_off"])
        want_keys = np.array([0 if k is None else k for k in o.keys], dtype=np.uint16)
        unpacked = np.concatenate([O.fl_unpack_chunk(keys[c * 1024:(c + 1) * 1024], 16) for c in range((n + 1023) // 1024)])[:n] if n else keys[:0]
        assert np.array_equal(unpacked, want_keys)
        if has_nulls:
            bits = np.unpackbits(np.frombuffer(image, dtype=np.uint8, count=(n + 7) // 8, offset=got["knulls_off"]), bitorder="little")[:n].astype(bool)
            assert bits.tolist() == [k is not None for k in o.keys]
        # the writer's layout for an entry with the same facts is the reader's
        out = (C.c_int64 * 16)()
        lib.lq_layout(n, U, int(has_nulls), int(bool(fps)), ob, len(o.compressed), len(o.shared_prefix), out)
        L = dict(zip(LAYOUT, [int(x) for x in out]))
        assert L["total"] == len(image), (ci, L["total"], len(image))
        assert (L["fsst_off"] + 12, L["keys_values_off"], L["co_off"] + 9, L["pk_off"], L["sp_off"], L["fp_off"]) == \
            (got["comp_off"], got["kvals_off"], got["resid_src"], got["pk_src"], got["sp_src"], got["fp_src"])
        if has_nulls:
            assert L["keys_nulls_off"] == got["knulls_off"]
        sizes = [int.from_bytes(image[16 + 4 * i:20 + 4 * i], "little") for i in range(5)]
        assert sizes == [L["keys_size"], L["co_size"], L["sp_size"], L["fsst_raw_size"], L["fp_size"]]


def test_damaged_images_are_refused(lib):
    arr = _build(_cases()[5], pa.string())
    o = O.OracleByteViewArray.from_arrow(arr, build_fingerprints=True)
    good = O.byte_view_to_bytes(o)
    keys_size, co_size, sp_size, fsst_size, fp_size = (int.from_bytes(good[16 + 4 * i:20 + 4 * i], "little") for i in range(5))
    keys_start = ((40 + fsst_size) + 7) & ~7
    co_start = ((keys_start + keys_size) + 7) & ~7
    ob = good[co_start + 8]
    last = co_start + 9 + len(o.uniques) * ob

    def patched(at, data):
        b = bytearray(good)
        b[at:at + len(data)] = data
        return bytes(b)

    bad = {"short": good[:30], "width": patched(keys_start + 4, bytes([9])), "fsst size": patched(28, (len(good) * 2).to_bytes(4, "little")),
           "residual width": patched(co_start + 8, bytes([3])), "closing offset": patched(last, (2 ** (8 * ob - 1) - 1).to_bytes(ob, "little")),
           "first offset": patched(co_start + 4, (o.offsets.intercept + 3).to_bytes(4, "little", signed=True)), "fingerprints cut": good[:-(fp_size // 2)],
           "byte type": patched(8, (9).to_bytes(2, "little")), "rows": patched(keys_start, (7000).to_bytes(4, "little"))}
    for name, image in bad.items():
        got, why = parse(lib, image)
        assert got is None and why, name
    assert parse(lib, good)[0] is not None
