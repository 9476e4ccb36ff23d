"""The arithmetic the squeeze kernels run per value (liquid_cache_b200/csrc/squeeze_math.cuh: calendar conversion, lossy
dates, clamp / quantize codes) compiled for the host and checked against the oracle and Arrow's calendar — the same source
the device kernels include, so this pins everything about k_date_component / k_date_lossy / k_squeeze_map except the
thread indexing."""
import ctypes as C
import os
import subprocess

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle import liquid_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ["Year", "Month", "Day", "DayOfWeek"]


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(ROOT, "build", "tests", "libsqueeze_math_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC", f"-I{ROOT}",
                        os.path.join(ROOT, "tests", "cpp", "squeeze_math_host.cc"), "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_calendar_round_trip_and_components(lib):
    rng = np.random.default_rng(1)
    days = np.concatenate([rng.integers(-2_000_000, 3_000_000, size=200_000), np.arange(-800, 800), [-(2**31), 2**31 - 1, -719_468, -719_469, 0]]).astype(np.int32)
    n = len(days)
    y, m, d = (np.zeros(n, dtype=np.int64) for _ in range(3))
    lib.sq_civil(_p(days), n, _p(y), _p(m), _p(d))
    oy, om, od = O.ymd_from_epoch_days(days.astype(np.int64))
    assert np.array_equal(y, oy) and np.array_equal(m, om) and np.array_equal(d, od)
    assert m.min() == 1 and m.max() == 12 and d.min() == 1 and d.max() == 31
    lib.sq_days_from_civil.restype = C.c_int32
    for i in rng.integers(0, n, size=3000):
        assert lib.sq_days_from_civil(C.c_longlong(int(y[i])), C.c_longlong(int(m[i])), C.c_longlong(int(d[i]))) == int(days[i])
    in64 = days.astype(np.int64)
    for f, name in enumerate(FIELDS):
        out = np.zeros(n, dtype=np.int32)
        lib.sq_date_component(_p(in64), n, f, C.c_longlong(0), _p(out))
        assert np.array_equal(out, O.component_from_days(name, in64).astype(np.int32)), name
    # Arrow's calendar on the years it covers
    sub = days[(days > -700_000) & (days < 2_900_000)]
    arr = pa.array(sub, pa.int32()).cast(pa.date32())
    for f, fn in enumerate((pc.year, pc.month, pc.day, lambda a: pc.day_of_week(a, count_from_zero=True, week_start=7))):
        out = np.zeros(len(sub), dtype=np.int32)
        lib.sq_date_component(_p(sub.astype(np.int64)), len(sub), f, C.c_longlong(0), _p(out))
        assert np.array_equal(out, np.asarray(fn(arr)).astype(np.int32)), f


@pytest.mark.parametrize("unit", ["s", "ms", "us", "ns"])
def test_timestamp_ticks_to_components(lib, unit):
    t = O._TICKS_PER_DAY[unit]
    rng = np.random.default_rng(len(unit))
    span_days = min(2_000_000, (2**62) // t)
    ticks = rng.integers(-span_days * t, span_days * t, size=100_000).astype(np.int64)
    ticks = np.concatenate([ticks, np.array([0, -1, 1, t - 1, t, -t, -t - 1, -t + 1], dtype=np.int64)])
    for f, name in enumerate(FIELDS):
        out = np.zeros(len(ticks), dtype=np.int32)
        lib.sq_date_component(_p(ticks), len(ticks), f, C.c_longlong(t), _p(out))
        want = O.component_from_days(name, O.timestamp_to_days_since_epoch(ticks, unit).astype(np.int64)).astype(np.int32)
        assert np.array_equal(out, want), (unit, name)


def test_lossy_dates_give_the_component_back(lib):
    cases = {0: np.arange(-300, 9999, dtype=np.int32), 1: np.arange(1, 13, dtype=np.int32), 2: np.arange(1, 32, dtype=np.int32), 3: np.arange(0, 7, dtype=np.int32)}
    for f, comp in cases.items():
        out = np.zeros(len(comp), dtype=np.int32)
        lib.sq_lossy_days(_p(comp), len(comp), f, _p(out))
        back = np.zeros(len(comp), dtype=np.int32)
        lib.sq_date_component(_p(out.astype(np.int64)), len(comp), f, C.c_longlong(0), _p(back))
        assert np.array_equal(back, comp), FIELDS[f]
        want = [O.ymd_to_epoch_days(int(c), 1, 1) if f == 0 else O.ymd_to_epoch_days(1970, int(c), 1) if f == 1 else
                O.ymd_to_epoch_days(1970, 1, int(c)) if f == 2 else O.ymd_to_epoch_days(1970, 1, 4) + int(c) for c in comp]
        assert out.tolist() == want, FIELDS[f]
    big = np.array([2**31 - 1, 2**31 - 3, -(2**31)], dtype=np.int32)  # day of week: saturating_add
    out = np.zeros(3, dtype=np.int32)
    lib.sq_lossy_days(_p(big), 3, 3, _p(out))
    assert out.tolist() == [2**31 - 1, 2**31 - 1, -(2**31) + 3]


def test_clamp_and_quantize_codes(lib):
    rng = np.random.default_rng(3)
    offs = np.concatenate([rng.integers(0, 2**40, size=50_000, dtype=np.uint64), np.array([0, 1, 254, 255, 256, 2**32 - 1, 2**32, 2**64 - 1], dtype=np.uint64)])
    out = np.zeros(len(offs), dtype=np.uint64)
    for new_bw in (4, 8, 16, 32):
        sentinel = (1 << new_bw) - 1
        lib.sq_codes(_p(offs), len(offs), 0, C.c_ulonglong(sentinel), C.c_ulonglong(0), _p(out))
        assert np.array_equal(out, np.minimum(offs, np.uint64(sentinel)))          # primitive_array.rs:433-438
        for bw in (1, 7, 512, 2**32, 2**40 + 3):
            lib.sq_codes(_p(offs), len(offs), 1, C.c_ulonglong(sentinel), C.c_ulonglong(bw), _p(out))
            assert np.array_equal(out, np.minimum(offs // np.uint64(bw), np.uint64(sentinel)))   # :472-481
