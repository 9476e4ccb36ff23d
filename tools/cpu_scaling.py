"""How the CPU arm scales on THIS host: cgroup CPU quota / affinity as the container sees them, then the C port of the
reference path (oracle/c, persistent pool) at 1, 2, 4 ... threads on the URL LIKE workload. Run on the GPU box:
    python tools/cpu_scaling.py > gpurun_out/cpu_scaling.json
bench.py's cpu_baseline uses os.cpu_count() threads; this says what that many threads can actually get."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def read(path):
    try:
        return open(path).read().strip()
    except OSError:
        return None


def main():
    import bench_cpu

    info = {
        "os_cpu_count": os.cpu_count(),
        "sched_affinity": len(os.sched_getaffinity(0)),
        "cgroup_cpu_max": read("/sys/fs/cgroup/cpu.max"),
        "cgroup_v1_quota": read("/sys/fs/cgroup/cpu/cpu.cfs_quota_us"),
        "cgroup_v1_period": read("/sys/fs/cgroup/cpu/cpu.cfs_period_us"),
        "cpu_stat": read("/sys/fs/cgroup/cpu.stat"),
        "loadavg": read("/proc/loadavg"),
    }
    workload = sys.argv[1] if len(sys.argv) > 1 else "url_like"
    entries = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    import bench

    params = {"url_like": lambda: None, "shipdate": bench.shipdate_params, "int_filter": bench.int_filter_params}[workload]()
    arm = bench_cpu.CpuArm(workload, entries, 1, params=params)
    CO = arm.CO
    points = []
    t = 1
    maxt = os.cpu_count() or 1
    while True:
        pool = CO.ScanPool(t).bind(arm.entries)
        if arm.entries2 is not None:
            pool.bind_second(arm.entries2, 0, int(arm.params["uid"]))
        needle, op1, l1, op2, l2 = arm.args
        pool.scan(arm.kind, needle, op1, l1, op2, l2)
        reps, t0 = 0, time.perf_counter()
        rows = 0
        while True:
            _m, rows = pool.scan(arm.kind, needle, op1, l1, op2, l2)
            reps += 1
            dt = time.perf_counter() - t0
            if dt > 1.5:
                break
        pool.close()
        points.append({"threads": t, "Mrows_per_s": rows * reps / dt / 1e6, "ms_per_pass": dt / reps * 1e3})
        if t >= maxt:
            break
        t = min(maxt, t * 2)
    info["cpu_stat_after"] = read("/sys/fs/cgroup/cpu.stat")
    base = points[0]["Mrows_per_s"]
    for p in points:
        p["speedup"] = p["Mrows_per_s"] / base
    print(json.dumps({"workload": workload, "entries": entries, "host": info, "scaling": points}, indent=1))


if __name__ == "__main__":
    main()
