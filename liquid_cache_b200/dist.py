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


class StepGather:
    """The gather of `gather_device_result_to_rank0`, shaped for use INSIDE a timed loop: buffers are allocated once and
    grown geometrically, the per-rank sizes travel in one small all_gather, every part lands at its final position in rank
    0's buffers (no per-part host bounce), string offsets are rebased on the device, and rank 0 downloads the finished
    Arrow buffers once into page-locked memory that the returned array references directly.

    One host synchronisation on every rank (the sizes), one more on rank 0 (the result)."""

    def __init__(self, arrow_type: pa.DataType, rank: int, world: int, device):
        import torch

        self.t, self.rank, self.world, self.dev = arrow_type, rank, world, device
        self.is_bytes = pa.types.is_string(arrow_type) or pa.types.is_binary(arrow_type)
        self.width = 0 if self.is_bytes else arrow_type.bit_width // 8
        self.torch = torch
        self.sizes = torch.zeros(2, dtype=torch.int64, device=device)
        self.all_sizes = torch.zeros(2 * world, dtype=torch.int64, device=device)
        self.vals = self.offs = self.offs_packed = None
        self.h_vals = self.h_offs = None
        self.pin = device.type == "cuda"

    def _grow(self, name: str, n: int, dtype, host: bool):
        cur = getattr(self, name)
        if cur is None or cur.numel() < n:
            cap = max(n + n // 2, 1 << 16)
            t = (self.torch.empty(cap, dtype=dtype, pin_memory=self.pin) if host
                 else self.torch.empty(cap, dtype=dtype, device=self.dev))
            setattr(self, name, t)
        return getattr(self, name)

    def gather(self, values, offsets, rows: int) -> Optional[pa.Array]:
        """values: u8 tensor of this rank's value bytes (ints: rows * width bytes); offsets: i32 tensor [rows + 1] for byte
        types, else None. No nulls (the bench columns have none). Returns the concatenation on rank 0."""
        import torch.distributed as dist

        torch = self.torch
        nbytes = int(values.numel())
        if self.world == 1:
            all_sizes = [rows, nbytes]
        else:
            self.sizes[0] = rows
            self.sizes[1] = nbytes
            dist.all_gather_into_tensor(self.all_sizes, self.sizes)
            all_sizes = self.all_sizes.cpu().tolist()  # the one synchronisation every rank pays
        rows_r = [int(all_sizes[2 * r]) for r in range(self.world)]
        bytes_r = [int(all_sizes[2 * r + 1]) for r in range(self.world)]
        if self.rank != 0:
            ops = []
            if nbytes:
                ops.append(dist.P2POp(dist.isend, values, 0))
            if self.is_bytes and rows:
                ops.append(dist.P2POp(dist.isend, offsets[: rows + 1].view(torch.uint8), 0))
            for w in (dist.batch_isend_irecv(ops) if ops else []):
                w.wait()
            return None
        tot_rows, tot_bytes = sum(rows_r), sum(bytes_r)
        vals = self._grow("vals", max(tot_bytes, 1), torch.uint8, False)
        vals[:nbytes].copy_(values, non_blocking=True)
        offs = None
        if self.is_bytes:
            offs = self._grow("offs", tot_rows + self.world + 1, torch.int32, False)
            offs[: rows + 1].copy_(offsets[: rows + 1], non_blocking=True)
        ops, places = [], []
        vb, ob = nbytes, rows + 1
        for r in range(1, self.world):
            if bytes_r[r]:
                ops.append(dist.P2POp(dist.irecv, vals[vb: vb + bytes_r[r]], r))
            if self.is_bytes and rows_r[r]:
                # each part arrives with its own rows + 1 offsets; they are rebased and packed below
                ops.append(dist.P2POp(dist.irecv, offs[ob: ob + rows_r[r] + 1].view(torch.uint8), r))
            places.append((r, vb, ob))
            vb += bytes_r[r]
            ob += rows_r[r] + 1
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        h_vals = self._grow("h_vals", max(tot_bytes, 1), torch.uint8, True)
        if self.is_bytes:
            # pack the offsets: part r's rows start at byte vb_r of the concatenation; its first offset (0) is dropped
            packed = self._grow("offs_packed", tot_rows + 1, torch.int32, False)
            packed[: rows + 1].copy_(offs[: rows + 1])
            at = rows + 1
            for r, vb_r, ob_r in places:
                if rows_r[r]:
                    packed[at: at + rows_r[r]] = offs[ob_r + 1: ob_r + 1 + rows_r[r]] + vb_r
                    at += rows_r[r]
            h_offs = self._grow("h_offs", tot_rows + 1, torch.int32, True)
            h_offs[: tot_rows + 1].copy_(packed[: tot_rows + 1], non_blocking=True)
        h_vals[:tot_bytes].copy_(vals[:tot_bytes], non_blocking=True)
        if self.pin:
            torch.cuda.current_stream().synchronize()
        data = pa.foreign_buffer(h_vals.data_ptr(), tot_bytes, base=h_vals)
        if self.is_bytes:
            ofb = pa.foreign_buffer(h_offs.data_ptr(), (tot_rows + 1) * 4, base=h_offs)
            return pa.Array.from_buffers(self.t, tot_rows, [None, ofb, data], null_count=0)
        return pa.Array.from_buffers(self.t, tot_rows, [None, data], null_count=0)
