"""A pyarrow stand-in for the few cache / scan calls bench_sweep.py makes, so the harness's own logic (query table, literal
resolution, expression lowering, host fallback, parity check, result bookkeeping) can be exercised without a GPU.
TEST DOUBLE ONLY: it evaluates with Arrow on the CPU and has nothing to do with the product path."""
import datetime as dt

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc


class _Insert:
    def __init__(self, cache, eid, arr):
        self.cache, self.eid, self.arr = cache, eid, arr

    def with_squeeze_hint(self, _hint):
        return self

    def run(self):
        self.cache.store[int(self.eid)] = self.arr


class _Stats:
    hbm_bytes_used = 0


def _eval(expr, arr):
    from liquid_cache_b200 import BinaryExpr, CastExpr, LikeExpr

    e = expr.physical_expr()
    if isinstance(e, LikeExpr):
        m = pc.match_like(arr, e.pattern.value)
        return pc.invert(m) if e.negated else m
    assert isinstance(e, BinaryExpr)
    left = e.left
    while isinstance(left, CastExpr):
        left = left.expr
    lit = e.right.value
    if isinstance(lit, dt.date):
        lit = (lit - dt.date(1970, 1, 1)).days
    fn = {"=": pc.equal, "!=": pc.not_equal, ">=": pc.greater_equal, "<=": pc.less_equal, "<": pc.less, ">": pc.greater}[e.op]
    return fn(arr, pa.scalar(lit, arr.type))


class FakeScan:
    def __init__(self, cache, rows):
        self.cache, self.rows = cache, [int(r) for r in rows]
        self.reset()

    def reset(self):
        self.sel = [np.ones(r, dtype=bool) for r in self.rows]

    def filter(self, handles, expr, column_type):
        expr.to_native(column_type)  # the lowering the real call performs; raises if the shape is not pushed down
        for b, h in enumerate(handles):
            m = np.asarray(_eval(expr, self.cache.store[int(h)]).fill_null(False).to_numpy(zero_copy_only=False), dtype=bool)
            self.sel[b] &= m

    def counts(self):
        c = np.array([int(s.sum()) for s in self.sel], dtype=np.uint64)
        return c, int(c.sum())

    def selection(self, b):
        return pa.array(self.sel[b])

    def set_selection(self, b, sel):
        self.sel[b] = np.asarray(sel, dtype=bool).copy()

    def read(self, handles):
        parts = [self.cache.store[int(h)].filter(pa.array(s)) for h, s in zip(handles, self.sel) if s.any()]
        return pa.concat_arrays(parts) if parts else self.cache.store[int(handles[0])].slice(0, 0)

    def read_torch(self, handles, device):
        return self.read(handles)

    def read_torch_borrowed(self, handles, device):
        return None  # the double has no device: the sweep falls back to read_torch

    def close(self):
        pass


class FakeScanWords(FakeScan):
    """The double with the bulk selection transfer of the real scan (lc_scan_store_selections / lc_scan_load_selections):
    uint32 words, LSB first, every batch padded to a multiple of four words — what the sweep's caller-evaluated conjuncts
    (the IN list) go through on a GPU."""

    def _layout(self):
        offs, tot = [], 0
        for r in self.rows:
            offs.append(tot)
            tot += ((r + 31) // 32 + 3) // 4 * 4
        return offs, tot

    def store_selections(self):
        offs, tot = self._layout()
        out = np.zeros(tot, dtype=np.uint32)
        for o, s in zip(offs, self.sel):
            w = np.packbits(s, bitorder="little")
            w = np.concatenate([w, np.zeros((-len(w)) % 4, dtype=np.uint8)]).view(np.uint32)
            out[o:o + len(w)] = w
        return out

    def load_selections(self, words):
        offs, _tot = self._layout()
        words = np.ascontiguousarray(words, dtype=np.uint32)
        for b, (o, r) in enumerate(zip(offs, self.rows)):
            n_w = (r + 31) // 32
            self.sel[b] = np.unpackbits(words[o:o + n_w].view(np.uint8), bitorder="little")[:r].astype(bool)


class FakeCache:
    def __init__(self, bulk_selections=False):
        self.store = {}
        self.bulk_selections = bulk_selections

    def insert(self, eid, arr):
        return _Insert(self, eid, arr)

    def insert_many(self, eids, arrays, hint=None):
        for e, a in zip(eids, arrays):
            self.store[int(e)] = a

    def handles(self, ids):
        return np.asarray(ids, dtype=np.uint64)

    def scan(self, rows):
        return FakeScanWords(self, rows) if self.bulk_selections else FakeScan(self, rows)

    def stats(self):
        return _Stats()
