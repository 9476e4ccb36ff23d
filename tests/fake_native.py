"""A stand-in for liblc_gpu.so built on the CPU oracle, for DRY RUNS of the GPU tests on a machine without a GPU.

Not a product path and not a parity claim: it exists so that the Python half of the GPU tests — the mirror classes in
liquid_cache_b200/cache.py and expr.py (argument marshalling, literal lowering, result decoding) and the tests' own
expectations — runs before a GPU is spent on them. `LC_FAKE_NATIVE=1 python -m pytest tests/test_gpu_zy_*.py -m gpu`
(tests/conftest.py installs it); tests/test_zz_dry_run_cpu.py does that in a subprocess. Calls that need the HBM image of an
entry skip the test. The entry points mimic include/lc_gpu.h as cache.py calls them: addresses arrive as ints, out
parameters as ctypes byref objects or arrays.
"""
import ctypes as C
import decimal

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from liquid_cache_b200 import _native as N
from oracle import liquid_oracle as O

OPS = {0: "=", 1: "!=", 2: "<", 3: "<=", 4: ">", 5: ">=", 6: "like", 7: "not like", 8: "true", 9: "false"}
FORMATS = {pa.int8(): "c", pa.int16(): "s", pa.int32(): "i", pa.int64(): "l", pa.uint8(): "C", pa.uint16(): "S", pa.uint32(): "I", pa.uint64(): "L",
           pa.date32(): "tdD", pa.date64(): "tdm", pa.timestamp("s"): "tss:", pa.timestamp("ms"): "tsm:", pa.timestamp("us"): "tsu:",
           pa.timestamp("ns"): "tsn:", pa.string(): "u", pa.binary(): "z", pa.string_view(): "vu", pa.binary_view(): "vz", pa.float32(): "f",
           pa.float64(): "g", pa.dictionary(pa.uint16(), pa.string()): "S:u", pa.dictionary(pa.uint16(), pa.binary()): "S:z"}
BYTE_TYPES = {0: pa.string(), 1: pa.string_view(), 2: pa.dictionary(pa.uint16(), pa.binary()), 3: pa.dictionary(pa.uint16(), pa.string()),
              4: pa.binary(), 5: pa.binary_view()}


class Fail(Exception):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def _format_of(t):
    if pa.types.is_decimal(t):
        return f"d:{t.precision},{t.scale}" + (",256" if pa.types.is_decimal256(t) else "")
    return FORMATS[t]


def _bits_at(addr, n):
    if not addr:
        return None
    raw = (C.c_uint8 * ((n + 7) // 8)).from_address(addr)
    return pa.array(np.unpackbits(np.frombuffer(raw, dtype=np.uint8), bitorder="little")[:n].astype(bool))


def _write(addr, data: bytes):
    if addr and data:
        C.memmove(addr, data, len(data))


def _set(ref, value):
    if ref is None:
        return
    if hasattr(ref, "_obj"):
        ref._obj.value = value
    else:
        C.c_uint64.from_address(int(ref)).value = value


def _addr(x):
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if hasattr(x, "value"):
        return x.value or 0
    return C.addressof(x)


class Entry:
    def __init__(self, oracle, arr_type, kind):
        self.o, self.type, self.kind = oracle, arr_type, kind  # kind: lc_liquid_type
        self.squeeze = 0
        self.io = None
        self.backing_len = 0
        self.scope = None

    def __len__(self):
        o = self.o
        if hasattr(o, "n"):
            return int(o.n)
        if hasattr(o, "keys"):
            return len(o.keys)
        if hasattr(o, "ints"):
            return int(o.ints.n)
        return len(o)


class FakeLib:
    def __init__(self):
        self.err = b""
        self.entries, self.next_h = {}, 1000
        self.caches, self.scopes, self.scans = {}, {}, {}
        self.reads = self.saved = 0

    # ---- plumbing ----
    def _guard(self, fn):
        try:
            return fn() or 0
        except Fail as e:
            self.err = str(e).encode()
            return e.code
        except (AssertionError, IndexError, ValueError, KeyError, OverflowError) as e:  # what a damaged image trips in the oracle
            self.err = f"{type(e).__name__}: {e}".encode()
            return N.LC_ERR_INVALID if hasattr(N, "LC_ERR_INVALID") else -1

    def lc_last_error(self):
        return self.err

    def lc_version(self):
        return b"fake (CPU oracle)"

    def _entry(self, h):
        e = self.entries.get(int(h))
        if e is None:
            raise Fail(-1, "invalid handle")
        return e

    def _new(self, e):
        self.next_h += 8
        self.entries[self.next_h] = e
        return self.next_h

    # ---- context ----
    def lc_ctx_create(self, device, budget, out):
        _set(out, 1)
        return 0

    def lc_ctx_destroy(self, ctx):
        return None

    def lc_ctx_set_stream(self, ctx, s):
        return 0

    def lc_ctx_synchronize(self, ctx):
        return 0

    def lc_ctx_kernel_timing(self, ctx, on):
        return 0

    def lc_ctx_last_kernel_ms(self, ctx):
        return 0.0

    def lc_ctx_stats(self, ctx, out):
        return 0

    def lc_mask_bytes(self, n):
        return ((int(n) + 7) // 8 + 15) // 16 * 16

    # ---- entries ----
    def _encode(self, arr, hint, scope):
        t = arr.type
        base = t.value_type if pa.types.is_dictionary(t) else t
        is_bytes = pa.types.is_string(base) or pa.types.is_binary(base) or pa.types.is_string_view(base) or pa.types.is_binary_view(base)
        fsst = self.scopes.get(scope) if (is_bytes or pa.types.is_decimal(t)) else None
        try:
            o = O.transcode(arr, fsst, build_fingerprints=(hint == N.HINT_SUBSTRING_SEARCH))
        except AssertionError as e:
            raise Fail(-2, str(e))
        if o is None:
            raise Fail(-2, f"unsupported arrow type {t}")
        kind = {O.OracleIntArray: 1, O.OracleFloatArray: 2, O.OracleDecimalArray: 6, O.OracleByteViewArray: 4}.get(type(o), 3)
        if getattr(o, "fsst", None) is not None and scope not in self.scopes:
            self.scopes[scope] = o.fsst
        e = Entry(o, t, kind)
        e.scope = scope
        return e

    def lc_encode(self, ctx, sch, arr, hint, scope, out):
        def run():
            a = pa.Array._import_from_c(arr, sch)
            _set(out, self._new(self._encode(a, hint, int(scope))))
        return self._guard(run)

    def lc_release(self, ctx, h):
        self.entries.pop(int(h), None)

    def lc_len(self, ctx, h):
        return len(self._entry(h))

    def lc_memory_size(self, ctx, h):
        e = self._entry(h)
        return 1000 if e.squeeze else 2000

    def lc_data_type(self, ctx, h):
        return self._entry(h).kind

    def lc_arrow_format(self, ctx, h, buf, n):
        f = _format_of(self._entry(h).type).encode()
        buf.value = f
        return 0

    def lc_entry_image(self, *a):
        pytest.skip("dry run: the HBM image of an entry exists on the device only")

    lc_entry_fsst_table = lc_entry_image

    # ---- predicates ----
    def _literal(self, e, p):
        t = e.type
        kind = p.lit_kind
        if kind == N.LIT_BYTES:
            raw = getattr(p, "_keepalive", None)
            raw = raw if raw is not None else C.string_at(p.lit_bytes, p.lit_len)
            if pa.types.is_decimal(t):
                u = int.from_bytes(raw, "little", signed=True)
                with decimal.localcontext() as cx:
                    cx.prec = 100
                    return pa.scalar(decimal.Decimal(u).scaleb(-t.scale), t)
            return raw
        if kind == N.LIT_I128:
            u = (p.lit_i64 << 64) | p.lit_u64
            with decimal.localcontext() as cx:
                cx.prec = 100
                return pa.scalar(decimal.Decimal(u).scaleb(-t.scale), t)
        if kind == N.LIT_F64:
            return float(np.array([p.lit_u64], dtype=np.uint64).view(np.float64)[0])
        return int(p.lit_i64) if kind == N.LIT_I64 else int(p.lit_u64)

    def _eval(self, e, p, sel):
        op = OPS.get(p.op)
        is_bytes = e.kind == 4
        if op is None or (not is_bytes and p.op > 5):
            raise Fail(-3, f"operator {p.op} is not supported")
        n = len(e)
        sel = sel if sel is not None else pa.array([True] * n)
        lit = self._literal(e, p) if p.op < 8 else None
        if is_bytes and p.op in (6, 7):
            lit = lit.decode() if isinstance(lit, bytes) else lit
            if not (len(lit) >= 3 and lit[0] == "%" and lit[-1] == "%" and "%" not in lit[1:-1] and "_" not in lit[1:-1]):
                raise Fail(-3, "LIKE pattern is not %x%")
        if e.squeeze:
            before = e.o.io.reads
            m = e.o.try_eval_predicate(op, lit, sel)
            self.reads += e.o.io.reads - before
            return m
        if not is_bytes and not isinstance(lit, (pa.Scalar, float)):
            lit = _int_scalar(lit, e.type)
            if lit is None:
                return _const_compare(op, e, sel, self._literal(e, p))
        return e.o.try_eval_predicate(op, lit, sel)

    def _write_mask(self, m, vals_addr, valid_addr):
        n = len(m)
        valid = np.asarray(m.is_valid().to_numpy(zero_copy_only=False), dtype=bool) if n else np.zeros(0, bool)
        bits = np.asarray(pc.fill_null(m, False).to_numpy(zero_copy_only=False), dtype=bool) if n else np.zeros(0, bool)
        _write(vals_addr, np.packbits(bits & valid, bitorder="little").tobytes())
        if valid_addr:
            _write(valid_addr, np.packbits(valid, bitorder="little").tobytes())
        return n, int((~valid).sum()), int((bits & valid).sum())

    def lc_eval_predicate(self, ctx, h, pred, sel, sel_len, vals, valid, out_len, out_nulls):
        def run():
            e = self._entry(h)
            m = self._eval(e, pred._obj, _bits_at(_addr(sel), len(e)))
            n, nulls, _t = self._write_mask(m, _addr(vals), _addr(valid))
            _set(out_len, n)
            _set(out_nulls, nulls)
        return self._guard(run)

    def lc_eval_predicate_many(self, ctx, handles, n, pred, sels, vals, valid, offs, out_len, out_nulls, out_true):
        def run():
            hs = np.frombuffer((C.c_uint64 * n).from_address(_addr(handles)), dtype=np.uint64)
            of = np.frombuffer((C.c_uint64 * n).from_address(_addr(offs)), dtype=np.uint64) if offs else np.zeros(n, np.uint64)
            for i in range(n):
                e = self._entry(hs[i])
                if e.squeeze == 3:
                    raise Fail(-1, "date-component entries answer through lc_eval_predicate")
                sel_addr = (sels[i] or 0) if sels is not None else 0
                m = self._eval(e, pred._obj, _bits_at(sel_addr, len(e)))
                k, nulls, trues = self._write_mask(m, _addr(vals) + int(of[i]), (_addr(valid) + int(of[i])) if valid else 0)
                for ref, v in ((out_len, k), (out_nulls, nulls), (out_true, trues)):
                    if ref:
                        C.c_uint64.from_address(_addr(ref) + 8 * i).value = v
        return self._guard(run)

    # ---- reads ----
    def _read(self, e, sel):
        if e.squeeze:
            before = e.o.io.reads
            out = e.o.filter(sel) if sel is not None else e.o.to_arrow()
            self.reads += e.o.io.reads - before
            return out
        out = e.o.filter(sel) if sel is not None else e.o.to_arrow()
        if pa.types.is_string_view(out.type) or pa.types.is_binary_view(out.type):
            out = pa.array(out.to_pylist(), out.type)  # pyarrow crashes exporting a view array that came out of a cast
        return out

    def lc_to_arrow(self, ctx, h, sel, sel_len, out_s, out_a):
        def run():
            e = self._entry(h)
            self._read(e, _bits_at(_addr(sel), len(e)))._export_to_c(out_a, out_s)
        return self._guard(run)

    def lc_to_arrow_many(self, ctx, handles, n, sels, out_s, out_a):
        def run():
            hs = np.frombuffer((C.c_uint64 * n).from_address(_addr(handles)), dtype=np.uint64)
            parts = []
            for i in range(n):
                e = self._entry(hs[i])
                if e.squeeze:
                    raise Fail(-1, "squeezed entries answer through lc_to_arrow / lc_eval_predicate")
                if sels is None:
                    sel_addr = 0
                elif isinstance(sels, int):
                    sel_addr = int(C.c_uint64.from_address(sels + 8 * i).value)
                else:
                    sel_addr = sels[i] or 0
                parts.append(self._read(e, _bits_at(sel_addr, len(e))))
            if len({(p.type, self._entry(hs[i]).kind) for i, p in enumerate(parts)}) > 1:
                raise Fail(-1, "entries of different types in one call")
            pa.concat_arrays(parts)._export_to_c(out_a, out_s)
        return self._guard(run)

    # ---- cache index ----
    def lc_cache_insert(self, ctx, eid, sch, arr, hint):
        def run():
            a = pa.Array._import_from_c(arr, sch)
            self.caches[int(eid)] = self._new(self._encode(a, hint, int(eid) & ~0xFFFF))
        return self._guard(run)

    def lc_cache_insert_many(self, ctx, ids, n, schs, arrs, hint):
        def run():
            ids_v = np.frombuffer((C.c_uint64 * n).from_address(_addr(ids)), dtype=np.uint64)
            staged = [(int(ids_v[i]), self._encode(pa.Array._import_from_c(arrs[i], schs[i]), hint, int(ids_v[i]) & ~0xFFFF)) for i in range(n)]
            for eid, e in staged:
                self.caches[eid] = self._new(e)
        return self._guard(run)

    def lc_cache_is_cached(self, ctx, eid):
        return 1 if int(eid) in self.caches else 0

    def lc_cache_handles(self, ctx, ids, n, out):
        def run():
            ids_v = np.frombuffer((C.c_uint64 * n).from_address(_addr(ids)), dtype=np.uint64)
            for i in range(n):
                if int(ids_v[i]) not in self.caches:
                    raise Fail(-5, "entry not cached")
                C.c_uint64.from_address(_addr(out) + 8 * i).value = self.caches[int(ids_v[i])]
        return self._guard(run)

    def lc_cache_retain(self, ctx, eid, out):
        if int(eid) not in self.caches:
            return -5
        _set(out, self._new(self._entry(self.caches[int(eid)])))  # a second handle on the same entry: its own reference
        return 0

    def lc_cache_get(self, ctx, eid, sel, sel_len, out_s, out_a):
        return self.lc_to_arrow(ctx, self.caches[int(eid)], sel, sel_len, out_s, out_a)

    def lc_cache_reset(self, ctx):
        self.caches.clear()
        return 0

    # ---- LQDA ----
    def lc_to_bytes(self, ctx, h, out, cap, out_bytes):
        def run():
            e = self._entry(h)
            if e.squeeze or e.kind == 3:
                raise Fail(-2, "no serialized form")
            img = O.byte_view_to_bytes(e.o) if e.kind == 4 else O.to_bytes(e.o)
            _set(out_bytes, len(img))
            _write(_addr(out), img)
        return self._guard(run)

    def _from_bytes(self, b, fsst):
        logical, physical = int.from_bytes(b[6:8], "little"), int.from_bytes(b[8:10], "little")
        if logical == 4:
            if fsst is None:
                raise Fail(-1, "a byte-view image needs its symbol table")
            t = BYTE_TYPES[physical]
            # what ipc_host.cc checks and the oracle's parser does not: section sizes against the image, key width 16
            keys_size, co_size, sp_size, fsst_size, fp_size = (int.from_bytes(b[16 + 4 * i:20 + 4 * i], "little") for i in range(5))
            cur = ((40 + fsst_size) + 7) & ~7
            if len(b) < cur + keys_size or (int.from_bytes(b[cur:cur + 4], "little") and b[cur + 4] != 16):
                raise Fail(-1, "keys are not bit-packed at width 16")
            cur = ((cur + keys_size) + 7) & ~7
            n_resid = (co_size - 9) // max(b[cur + 8], 1) if co_size else 0
            cur = ((cur + co_size) + 7) & ~7
            cur = ((cur + 8 * max(n_resid - 1, 0)) + 7) & ~7
            cur = ((cur + sp_size) + 7) & ~7
            if len(b) < cur + fp_size:
                raise Fail(-1, "dictionary sections run past the image")
            o = O.byte_view_from_bytes(b, fsst, t)
            limit = max(len(o.uniques), 1)
            if any(k is not None and k >= limit for k in o.keys):
                raise Fail(-1, "a dictionary key is past the end of the dictionary")
            offs = [o.offsets.get_offset(i) for i in range(len(o.offsets.residuals))]
            if offs and (offs[0] != 0 or any(a > b2 for a, b2 in zip(offs, offs[1:])) or offs[-1] > len(o.compressed)):
                raise Fail(-1, "dictionary offsets do not fit the compressed values")
            o.to_arrow()
            return Entry(o, t, 4)
        o = O.read_from_bytes(b)
        t = o.arrow_type
        o.to_arrow()
        return Entry(o, t, logical)

    def lc_from_bytes(self, ctx, data, n, out):
        def run():
            _set(out, self._new(self._from_bytes(C.string_at(_addr(data), n), None)))
        return self._guard(run)

    def lc_from_bytes_scoped(self, ctx, data, n, scope, out):
        def run():
            _set(out, self._new(self._from_bytes(C.string_at(_addr(data), n), self.scopes.get(int(scope)))))
        return self._guard(run)

    def lc_ctx_save_symbol_table(self, ctx, scope, out, cap, out_bytes):
        def run():
            if int(scope) not in self.scopes:
                raise Fail(-5, "no symbol table for this scope")
            img = O.save_symbol_table(self.scopes[int(scope)])
            _set(out_bytes, len(img))
            _write(_addr(out), img)
        return self._guard(run)

    def lc_ctx_load_symbol_table(self, ctx, scope, data, n):
        def run():
            if int(scope) in self.scopes:
                raise Fail(-1, "scope already has a symbol table")
            self.scopes[int(scope)] = O.load_symbol_table(C.string_at(_addr(data), n))
        return self._guard(run)

    # ---- squeeze ----
    def lc_squeeze(self, ctx, h, policy, hint, read, user, out, cap, out_bytes, out_sq):
        def run():
            e = self._entry(h)
            _set(out_bytes, 0)
            _set(out_sq, 0)
            if e.kind != 1 or e.squeeze or hint == N.HINT_NONE:
                return
            field = {v: k for k, v in N.HINT_EXTRACT.items()}.get(hint)
            ohint = ("ExtractDate32", field) if field else "PredicateColumn"
            io = _CallbackIo(read)
            io.expect = (e.o.n, e.o.arrow_type)
            got = O.squeeze_int(e.o, io, ohint, "clamp" if policy == N.SQUEEZE_CLAMP else "quantize")
            if got is None:
                return
            sq, image = got
            _set(out_bytes, len(image))
            if not out:
                return
            _write(_addr(out), image)
            s = Entry(sq, e.type, 1)
            s.squeeze = 3 if field else (1 if policy == N.SQUEEZE_CLAMP else 2)
            s.backing_len = len(image)
            _set(out_sq, self._new(s))
        return self._guard(run)

    def lc_squeezed_info(self, ctx, h, out):
        e = self._entry(h)
        o = e.o
        out[0] = e.squeeze
        out[1] = (o.bit_width or 0) if e.squeeze else 0
        out[2] = O.DATE32_FIELDS.index(o.field) if e.squeeze == 3 else getattr(o, "bucket_width", 0) if e.squeeze == 2 else 0
        out[3] = e.backing_len
        out[4], out[5] = self.reads, self.saved
        return 0

    def lc_squeezed_component(self, ctx, h, lossy, out_s, out_a):
        def run():
            e = self._entry(h)
            if e.squeeze != 3:
                raise Fail(-1, "not a date-component entry")
            (e.o.to_component_array() if lossy else e.o.to_component_date32())._export_to_c(out_a, out_s)
        return self._guard(run)

    # ---- scan pipeline ----
    def lc_scan_begin(self, ctx, n, rows, out):
        r = np.frombuffer((C.c_uint64 * n).from_address(_addr(rows)), dtype=np.uint64).copy()
        self.next_h += 8
        self.scans[self.next_h] = {"rows": r, "sel": [np.ones(int(x), dtype=bool) for x in r]}
        _set(out, self.next_h)
        return 0

    def _scan(self, s):
        return self.scans[_addr(s)]

    def lc_scan_reset(self, s):
        sc = self._scan(s)
        sc["sel"] = [np.ones(int(x), dtype=bool) for x in sc["rows"]]
        return 0

    def lc_scan_set_selection(self, s, batch, bits, n):
        sc = self._scan(s)
        sc["sel"][batch] = np.asarray(_bits_at(_addr(bits), n).to_numpy(zero_copy_only=False), dtype=bool)
        return 0

    def lc_scan_filter(self, s, handles, pred):
        def run():
            sc = self._scan(s)
            n = len(sc["rows"])
            hs = np.frombuffer((C.c_uint64 * n).from_address(_addr(handles)), dtype=np.uint64)
            for i in range(n):
                e = self._entry(hs[i])
                if e.squeeze == 3:
                    raise Fail(-1, "date-component entries answer through lc_eval_predicate")
                sel = pa.array(sc["sel"][i])
                m = pc.fill_null(self._eval(e, pred._obj, sel), False)
                out = np.zeros(len(e), dtype=bool)
                out[sc["sel"][i]] = np.asarray(m.to_numpy(zero_copy_only=False), dtype=bool)
                sc["sel"][i] = out
        return self._guard(run)

    def lc_scan_counts(self, s, out, total):
        sc = self._scan(s)
        tot = 0
        for i, m in enumerate(sc["sel"]):
            C.c_uint64.from_address(_addr(out) + 8 * i).value = int(m.sum())
            tot += int(m.sum())
        _set(total, tot)
        return 0

    def lc_scan_selection(self, s, batch, out):
        _write(_addr(out), np.packbits(self._scan(s)["sel"][batch], bitorder="little").tobytes())
        return 0

    def lc_scan_read(self, s, handles, out_s, out_a):
        def run():
            sc = self._scan(s)
            n = len(sc["rows"])
            hs = np.frombuffer((C.c_uint64 * n).from_address(_addr(handles)), dtype=np.uint64)
            parts = [self._read(self._entry(hs[i]), pa.array(sc["sel"][i])) for i in range(n) if sc["sel"][i].any()]
            if not parts:
                parts = [self._entry(hs[0]).o.to_arrow().slice(0, 0)]
            pa.concat_arrays(parts)._export_to_c(out_a, out_s)
        return self._guard(run)

    def lc_scan_end(self, s):
        self.scans.pop(_addr(s), None)


class _CallbackIo(O.OracleSqueezeIo):
    """SqueezeIoHandler over the C callback cache.py registers (lc_backing_read)."""

    def __init__(self, cb):
        super().__init__()
        self.cb = cb
        self.expect = None

    def read(self, rng=None):
        self.reads += 1
        off, end = rng
        buf = (C.c_uint8 * (end - off))()
        rc = self.cb(None, off, end - off, C.addressof(buf)) if self.cb else 1
        if rc != 0:
            raise Fail(-1, f"reading the backing bytes failed ({rc})")
        image = bytes(buf)
        if self.expect is not None and (O.read_from_bytes(image).n, O.read_from_bytes(image).arrow_type) != self.expect:
            raise Fail(-1, "the backing bytes are not the image this entry was squeezed from")  # squeeze_host.cc hydrate()
        return image


def _int_scalar(v, t):
    try:
        return pa.scalar(int(v), t)
    except (OverflowError, pa.ArrowInvalid, ValueError):
        return None


def _const_compare(op, e, sel, k):
    """a literal outside the column's type: the comparison folds to a constant side (nulls stay null)"""
    arr = e.o.filter(sel)
    info_hi = k > 0
    const = {"=": False, "!=": True, "<": info_hi, "<=": info_hi, ">": not info_hi, ">=": not info_hi}[op]
    return pc.if_else(arr.is_valid(), pa.scalar(const), pa.scalar(None, pa.bool_()))


def install():
    fake = FakeLib()
    N.lib = lambda: fake
    return fake
