"""Multi-GPU plumbing: entries shard by EntryID, one process per GPU, and the only exchange on the path is the
gather of the final filtered Arrow batches to rank 0.

torch.distributed is used for what it is good at — rendezvous and NCCL point-to-point over NVLink/NVSwitch (or
gloo on CPU in tests); no collective touches the scan itself, which is embarrassingly parallel over entries.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import pyarrow as pa


def shard_of(entry_id: int, world: int) -> int:
    """Owner rank of an entry. All columns of one (file, row group, batch) land on the same GPU so the selection
    mask never leaves the device between eval_predicate and get (column id, bits 16..31, is ignored)."""
    file_rg = entry_id >> 32
    batch = entry_id & 0xFFFF
    x = (file_rg * 0x9E3779B97F4A7C15 + batch * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 29
    return int(x % world)


def partition_entries(entry_ids: Sequence[int], world: int) -> list[list[int]]:
    parts: list[list[int]] = [[] for _ in range(world)]
    for e in entry_ids:
        parts[shard_of(int(e), world)].append(int(e))
    return parts


def _buffers_of(arr: pa.Array):
    """(validity bytes | None, offsets int32 | None, data bytes) of a Utf8/Binary or primitive array, offset 0."""
    if isinstance(arr, pa.ChunkedArray):
        arr = arr.combine_chunks()
    n = len(arr)
    valid = None
    if arr.null_count:
        valid = np.packbits(np.asarray(arr.is_valid().to_numpy(zero_copy_only=False), dtype=bool), bitorder="little")
    if pa.types.is_string(arr.type) or pa.types.is_binary(arr.type):
        bufs = arr.buffers()
        off = np.frombuffer(bufs[1], dtype=np.int32, count=n + 1 + arr.offset)[arr.offset:].astype(np.int64)
        data = np.frombuffer(bufs[2], dtype=np.uint8) if bufs[2] is not None else np.zeros(0, np.uint8)
        data = data[int(off[0]):int(off[-1])] if n else np.zeros(0, np.uint8)
        return valid, (off - off[0]).astype(np.int32), np.ascontiguousarray(data)
    width = arr.type.bit_width // 8
    bufs = arr.buffers()
    data = np.frombuffer(bufs[1], dtype=np.uint8)[arr.offset * width:(arr.offset + n) * width] if n else np.zeros(0, np.uint8)
    return valid, None, np.ascontiguousarray(data)


def gather_arrow_to_rank0(arr: pa.Array, rank: int, world: int, device=None) -> Optional[pa.Array]:
    """Variable-length gather of per-rank result arrays to rank 0: one small all_gather of sizes, then grouped
    point-to-point transfers (NCCL has no gatherv). Returns the concatenation (rank order) on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return arr
    backend = dist.get_backend()
    dev = device if (backend == "nccl") else torch.device("cpu")
    valid, off, data = _buffers_of(arr)
    n = len(arr)
    sizes = torch.tensor([n, 0 if valid is None else len(valid), 0 if off is None else len(off), len(data)], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(4, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [t.cpu().tolist() for t in all_sizes]

    def to_t(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).copy()).to(dev)

    if rank != 0:
        ops = []
        for a in (valid, off, data):
            if a is not None and len(a):
                ops.append(dist.P2POp(dist.isend, to_t(a), 0))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        return None
    parts = [arr]
    ops, recv = [], []
    for r in range(1, world):
        nr, nv, no, nd = all_sizes[r]
        bufs = []
        for nbytes in (nv, no * 4, nd):
            if nbytes:
                t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                ops.append(dist.P2POp(dist.irecv, t, r))
                bufs.append(t)
            else:
                bufs.append(None)
        recv.append((nr, bufs))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    for nr, (bv, bo, bd) in recv:
        vb = pa.py_buffer(bv.cpu().numpy().tobytes()) if bv is not None else None
        nulls = -1 if bv is not None else 0
        if pa.types.is_string(arr.type) or pa.types.is_binary(arr.type):
            ob = pa.py_buffer(bo.cpu().numpy().tobytes()) if bo is not None else pa.py_buffer(np.zeros(1, np.int32).tobytes())
            db = pa.py_buffer(bd.cpu().numpy().tobytes()) if bd is not None else pa.py_buffer(b"")
            parts.append(pa.Array.from_buffers(arr.type, nr, [vb, ob, db], null_count=nulls))
        else:
            db = pa.py_buffer(bd.cpu().numpy().tobytes()) if bd is not None else pa.py_buffer(b"")
            parts.append(pa.Array.from_buffers(arr.type, nr, [vb, db], null_count=nulls))
    return pa.concat_arrays(parts)


def gather_device_result_to_rank0(values, offsets, validity, rows: int, nulls: int, arrow_type: pa.DataType, rank: int,
                                  world: int) -> Optional[pa.Array]:
    """Same exchange as gather_arrow_to_rank0, but starting from the DEVICE-resident result of
    `Scan.read_torch` (lc_scan_read_device): the filtered values / int32 offsets / validity words go from each GPU's
    HBM to rank 0's HBM over NCCL without visiting the host; only rank 0 copies the concatenation down once.
    `values` u8 tensor, `offsets` i32 tensor (byte types) or None, `validity` u8 tensor or None."""
    import torch
    import torch.distributed as dist

    is_bytes = pa.types.is_string(arrow_type) or pa.types.is_binary(arrow_type)

    def host_array(v, o, b, n, nn):
        vb = pa.py_buffer(b.cpu().numpy().tobytes()) if (b is not None and nn) else None
        data = pa.py_buffer(v.cpu().numpy().tobytes())
        if is_bytes:
            return pa.Array.from_buffers(arrow_type, n, [vb, pa.py_buffer(o.cpu().numpy().tobytes()), data], null_count=nn)
        return pa.Array.from_buffers(arrow_type, n, [vb, data], null_count=nn)

    if world == 1:
        return host_array(values, offsets, validity, rows, nulls)
    dev = values.device
    sizes = torch.tensor([rows, nulls, values.numel(), 0 if validity is None else validity.numel()], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(4, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [t.cpu().tolist() for t in all_sizes]
    if rank != 0:
        ops = [dist.P2POp(dist.isend, t.view(torch.uint8) if t.dtype != torch.uint8 else t, 0)
               for t in (values, offsets, validity) if t is not None and t.numel()]
        for r in (dist.batch_isend_irecv(ops) if ops else []):
            r.wait()
        return None
    parts = [host_array(values, offsets, validity, rows, nulls)]
    ops, recv = [], []
    for r in range(1, world):
        nr, nn, nv, nb = all_sizes[r]
        tv = torch.empty(nv, dtype=torch.uint8, device=dev)
        to = torch.empty((nr + 1) * 4, dtype=torch.uint8, device=dev) if is_bytes else None
        tb = torch.empty(nb, dtype=torch.uint8, device=dev) if nb else None
        for t in (tv, to, tb):
            if t is not None and t.numel():
                ops.append(dist.P2POp(dist.irecv, t, r))
        recv.append((nr, nn, tv, to, tb))
    for r in (dist.batch_isend_irecv(ops) if ops else []):
        r.wait()
    for nr, nn, tv, to, tb in recv:
        parts.append(host_array(tv, to.view(torch.int32) if to is not None else None, tb, nr, nn))
    return pa.concat_arrays(parts)


class DeviceGather:
    """The one exchange of the path, shaped for use INSIDE a timed loop: every rank's filtered batch goes from its HBM into
    rank 0's HBM (into every rank's, it is one all_gather) with ONE collective and ONE host synchronisation per step.

    Each rank owns a fixed-capacity slot `[header 64 B | int32 offsets[rows_cap + 1] | value bytes]` in device memory;
    `lc_scan_read_async` (Scan.read_async) plans and writes the survivors of the filter straight into it without touching the
    host. `exchange()` all-gathers the slots over NCCL (NVLink / NVSwitch), downloads the `world` headers — 64 bytes each —
    and synchronises: that is the only point where the host waits. The gathered batches stay in HBM as `world` Arrow-layout
    parts (what a device consumer reads); `to_arrow()` materialises them on the host for checking, outside any clock.

    The library must run on torch's current stream (`LiquidCache.set_stream(torch.cuda.current_stream().cuda_stream)`) so
    that the collective and the header download are ordered behind the read without an event.

    Capacities are equal on every rank (all_gather needs equal slots) and every rank sees every header, so growing after an
    overflow is decided identically everywhere without another message."""

    HDR = 64

    def __init__(self, arrow_type: pa.DataType, rank: int, world: int, device, rows_cap: int = 1 << 12, values_cap: int = 1 << 18):
        import torch

        self.torch, self.t, self.rank, self.world, self.dev = torch, arrow_type, rank, world, device
        self.is_bytes = pa.types.is_string(arrow_type) or pa.types.is_binary(arrow_type)
        self.width = 0 if self.is_bytes else arrow_type.bit_width // 8
        self.pin = device.type == "cuda"
        self.grows = 0
        self._alloc(rows_cap, values_cap)

    def _alloc(self, rows_cap: int, values_cap: int):
        torch = self.torch
        self.rows_cap = int(rows_cap)
        self.values_cap = int(values_cap) if self.is_bytes else self.rows_cap * self.width
        self.off_at = self.HDR
        self.val_at = self.HDR + (((self.rows_cap + 1) * 4 + 255) // 256 * 256 if self.is_bytes else 0)
        self.slot = (self.val_at + self.values_cap + 255) // 256 * 256
        self.send = torch.zeros(self.slot, dtype=torch.uint8, device=self.dev)
        self.recv = self.send if self.world == 1 else torch.zeros(self.world * self.slot, dtype=torch.uint8, device=self.dev)
        self.h_hdr = torch.zeros((self.world, self.HDR), dtype=torch.uint8, pin_memory=self.pin)
        self.headers = None

    def addresses(self):
        """(d_values, values_cap, d_offsets, rows_cap, d_header): the arguments of Scan.read_async for this rank's slot."""
        base = self.send.data_ptr()
        return base + self.val_at, self.values_cap, (base + self.off_at) if self.is_bytes else 0, self.rows_cap, base

    def exchange(self):
        """All-gather the slots, download the headers, synchronise. Returns [(rows, value_bytes, overflow)] per rank."""
        import torch.distributed as dist

        if self.world > 1:
            dist.all_gather_into_tensor(self.recv, self.send)
        self.h_hdr.copy_(self.recv.view(self.world, self.slot)[:, : self.HDR], non_blocking=True)
        if self.pin:
            self.torch.cuda.current_stream().synchronize()  # the step's one synchronisation
        h = self.h_hdr.numpy()
        u32, u64 = h.view(np.uint32), h.view(np.uint64)
        self.headers = [(int(u64[r, 1]), int(u64[r, 2]) if self.is_bytes else int(u64[r, 1]) * self.width, int(u32[r, 1]))
                        for r in range(self.world)]
        return self.headers

    def overflowed(self) -> bool:
        return any(o for _r, _b, o in self.headers)

    def grow(self):
        """After an overflow: new equal capacities from what the headers say every rank needs (rows are always reported;
        the byte total is only known once the rows fit, so it doubles until then)."""
        rows = max(r for r, _b, _o in self.headers)
        rows_cap = max(self.rows_cap, rows + rows // 4 + 1024)
        values_cap = self.values_cap
        if self.is_bytes:
            known = max((b for _r, b, o in self.headers if o != 1), default=0)
            values_cap = max(values_cap, known + known // 4 + 4096)
            if any(o == 1 for _r, _b, o in self.headers):  # rows did not fit: bytes unknown there, scale with the rows
                values_cap = max(values_cap, self.values_cap * -(-rows_cap // self.rows_cap))
        self.grows += 1
        self._alloc(rows_cap, values_cap)

    def part(self, r: int):
        """(values u8 view, offsets i32 view | None, rows) of rank r's batch in THIS rank's HBM after exchange()."""
        rows, nbytes, _o = self.headers[r]
        slot = self.recv[r * self.slot:(r + 1) * self.slot]
        offs = slot[self.off_at: self.off_at + (rows + 1) * 4].view(self.torch.int32) if self.is_bytes else None
        return slot[self.val_at: self.val_at + nbytes], offs, rows

    def to_arrow(self, ranks=None) -> pa.Array:
        """The gathered batches as one host Arrow array (rank order). For checking: downloads, not part of a timed step."""
        parts = []
        for r in (range(self.world) if ranks is None else ranks):
            v, o, rows = self.part(r)
            data = pa.py_buffer(v.cpu().numpy().tobytes())
            if self.is_bytes:
                parts.append(pa.Array.from_buffers(self.t, rows, [None, pa.py_buffer(o.cpu().numpy().tobytes()), data], null_count=0))
            else:
                parts.append(pa.Array.from_buffers(self.t, rows, [None, data], null_count=0))
        return pa.concat_arrays(parts) if parts else pa.array([], type=self.t)
