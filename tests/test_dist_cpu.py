"""CPU (gloo, world_size 2) tests of the multi-GPU host logic: EntryID sharding and the gather of filtered
Arrow batches to rank 0. The scan itself needs no collective."""
import os
import socket

import numpy as np
import pyarrow as pa
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from liquid_cache_b200.dist import gather_arrow_to_rank0

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        strs = pa.array([None if rng.random() < 0.2 else f"http://r{rank}/{i}" * int(rng.integers(0, 4)) for i in range(50 + 30 * rank)])
        ints = pa.array(rng.integers(-5, 5, size=10 * (rank + 1)), pa.int64())
        empty = pa.array([], pa.string()) if rank == 1 else pa.array(["only-rank0"])
        import decimal

        floats = pa.array([None if rng.random() < 0.3 else float(rng.standard_normal()) for _ in range(20 + 7 * rank)], pa.float64())
        decs = pa.array([decimal.Decimal(int(x)).scaleb(-2) for x in rng.integers(0, 10**6, size=5 + rank)], pa.decimal128(15, 2))
        dates = pa.array(rng.integers(8036, 10556, size=9 * (rank + 1)), pa.int32()).cast(pa.date32())
        out = []
        for a in (strs, ints, empty):
            g = gather_arrow_to_rank0(a, rank, world)
            out.append(None if g is None else g.to_pylist())
        extra_local, extra_out = [], []
        for a in (floats, decs, dates):  # every fixed-width result type of the path travels the same way
            g = gather_arrow_to_rank0(a, rank, world)
            extra_local.append(a.to_pylist())
            extra_out.append(None if g is None else g.to_pylist())
        # the device-buffer flavour (lc_scan_read_device results), here on CPU tensors over gloo
        import torch

        from liquid_cache_b200.dist import _buffers_of, gather_device_result_to_rank0

        for a in (strs, ints):
            valid, off, data = _buffers_of(a)
            v = torch.from_numpy(data.copy())
            o = torch.from_numpy(off.copy()) if off is not None else None
            b = None
            if valid is not None:
                words = np.zeros(4 * ((len(a) + 31) // 32), dtype=np.uint8)
                words[: len(valid)] = valid
                b = torch.from_numpy(words)
            g = gather_device_result_to_rank0(v, o, b, len(a), a.null_count, a.type, rank, world)
            out.append(None if g is None else g.to_pylist())
        q.put((rank, [strs.to_pylist(), ints.to_pylist(), empty.to_pylist(), strs.to_pylist(), ints.to_pylist()] + extra_local, out + extra_out))
    finally:
        dist.destroy_process_group()


def test_gather_to_rank0_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        rank, local, out = q.get(timeout=120)
        res[rank] = (local, out)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k in range(8):
        assert res[0][1][k] == res[0][0][k] + res[1][0][k], k
        assert res[1][1][k] is None


def test_sharding_keeps_all_columns_of_a_batch_together():
    from liquid_cache_b200 import parquet_array_id
    from liquid_cache_b200.dist import partition_entries, shard_of

    ids = [parquet_array_id(f, rg, col, b) for f in range(2) for rg in range(5) for col in range(4) for b in range(32)]
    for world in (1, 2, 4, 8):
        parts = partition_entries(ids, world)
        assert sum(len(p) for p in parts) == len(ids)
        for f in range(2):
            for rg in range(5):
                for b in range(32):
                    owners = {shard_of(parquet_array_id(f, rg, col, b), world) for col in range(4)}
                    assert len(owners) == 1
        if world > 1:
            assert min(len(p) for p in parts) > 0.5 * len(ids) / world


def _fill_slot(g, arr):
    """What lc_scan_read_async does on the device, done with numpy on a CPU slot: header (rows at byte 8, value bytes at 16,
    overflow at 4), int32 offsets and value bytes — or only the header with the overflow code when a capacity is short."""
    from liquid_cache_b200.dist import _buffers_of

    _v, off, data = _buffers_of(arr)
    buf = g.send.numpy()
    d_values, values_cap, d_offsets, rows_cap, _d_hdr = g.addresses()
    base = g.send.data_ptr()
    hdr64, hdr32 = buf[:64].view(np.uint64), buf[:64].view(np.uint32)
    hdr64[:] = 0
    hdr64[1] = len(arr)
    if len(arr) > rows_cap:
        hdr32[1] = 1
        return
    hdr64[2] = len(data)
    if g.is_bytes and len(data) > values_cap:
        hdr32[1] = 2
        return
    if g.is_bytes:
        at = d_offsets - base
        buf[at: at + 4 * (len(arr) + 1)] = (off if len(arr) else np.zeros(1, np.int32)).view(np.uint8)
    at = d_values - base
    buf[at: at + len(data)] = data


def _device_gather_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from liquid_cache_b200.dist import DeviceGather

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cpu")
        gs = DeviceGather(pa.string(), rank, world, dev, rows_cap=64, values_cap=512)
        gi = DeviceGather(pa.int32(), rank, world, dev, rows_cap=64)
        local, got = [], []
        for step in range(4):  # sizes change from step to step, one rank is empty in step 2: slots are reused / grown
            rng = np.random.default_rng(1000 * step + rank)
            n = 0 if (step == 2 and rank == 1) else int(rng.integers(1, 200)) * (step + 1)
            strs = pa.array([f"http://r{rank}/s{step}/{i}" * int(rng.integers(1, 4)) for i in range(n)], pa.string())
            ints = pa.array(rng.integers(-1000, 1000, size=n).astype(np.int32), pa.int32())
            for g, arr in ((gs, strs), (gi, ints)):
                for _ in range(8):
                    _fill_slot(g, arr)
                    g.exchange()
                    if not g.overflowed():
                        break
                    g.grow()  # every rank sees every header: all of them grow to the same capacities
                assert not g.overflowed()
            local.append((strs.to_pylist(), ints.to_pylist()))
            got.append((gs.to_arrow().to_pylist(), gi.to_arrow().to_pylist(), [h[0] for h in gs.headers]))
        q.put((rank, local, got, gs.grows))
    finally:
        dist.destroy_process_group()


def test_device_gather_world2_matches_concatenation():
    """DeviceGather (the exchange bench.py times inside its step): after every step EVERY rank holds the ranks' batches in
    rank order; slots are reused across steps and grown, identically on all ranks, when a capacity was short."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_device_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        rank, local, got, grows = q.get(timeout=120)
        res[rank] = (local, got, grows)
    for p in procs:
        p.join(timeout=60)
    assert res[0][2] == res[1][2] and res[0][2] > 0
    for step in range(4):
        want_s = res[0][0][step][0] + res[1][0][step][0]
        want_i = res[0][0][step][1] + res[1][0][step][1]
        for r in range(2):
            assert res[r][1][step][0] == want_s
            assert res[r][1][step][1] == want_i
            assert res[r][1][step][2] == [len(res[0][0][step][0]), len(res[1][0][step][0])]
