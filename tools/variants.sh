#!/usr/bin/env bash
# Several branches in ONE gpurun call. The GPU box receives the working tree only (no .git), and every call pays a fixed
# cost of a couple of GPU-minutes, so candidate branches are exported side by side and built HERE, then run there:
#
#   tools/variants.sh make r2-all r2-trigram-filter ...     # here: git archive -> variants/<branch>/, build its libraries
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/variants.sh run'     # there: GPU tests + bench lines per variant
#   tools/variants.sh clean
#
# `run` writes gpurun_out/variants/<branch>/{pytest.log,bench_*.json,smoke.log} and a one-line verdict per variant in
# gpurun_out/variants/SUMMARY.txt. variants/ is git-ignored (it still travels with the snapshot).
# Environment for `run`: VARIANTS="a b" restricts the set; PYTEST_ARGS (default: -m gpu -q — every failure of a first run in
# one call, no -x); BENCH=0 skips the bench lines;
# PER_VARIANT_TIMEOUT seconds per pytest run (default 600).
set -u
root="$(cd "$(dirname "$0")/.." && pwd)"
cmd="${1:-}"
shift || true
case "$cmd" in
  make)
    [ $# -ge 1 ] || { echo "usage: $0 make <branch>..." >&2; exit 2; }
    for b in "$@"; do
      d="$root/variants/$b"
      rm -rf "$d"
      mkdir -p "$d"
      git -C "$root" archive "$b" | tar -x -C "$d" || { echo "cannot export $b" >&2; exit 1; }
      [ -f "$root/MEASURED_PEAKS.json" ] && cp "$root/MEASURED_PEAKS.json" "$d/"
      git -C "$root" rev-parse "$b" > "$d/VARIANT_COMMIT"
      (cd "$d" && python __graft_entry__.py build > build.log 2>&1) || { echo "$b: build FAILED (see variants/$b/build.log)" >&2; exit 1; }
      rm -rf "$d/build"   # objects stay home, the libraries travel
      echo "$b: built at $(cut -c1-10 "$d/VARIANT_COMMIT")"
    done
    ;;
  run)
    out="$root/gpurun_out/variants"
    mkdir -p "$out"
    : > "$out/SUMMARY.txt"
    set -- ${VARIANTS:-$(ls "$root/variants" 2>/dev/null)}
    for b in "$@"; do
      d="$root/variants/$b"
      [ -d "$d" ] || { echo "$b: no such variant" | tee -a "$out/SUMMARY.txt"; continue; }
      o="$out/$b"
      mkdir -p "$o"
      (
        cd "$d" || exit 1
        timeout "${PER_VARIANT_TIMEOUT:-600}" python -m pytest tests ${PYTEST_ARGS:--m gpu -q} -p no:cacheprovider > "$o/pytest.log" 2>&1
        echo "pytest exit $?" >> "$o/pytest.log"
        timeout 120 python __graft_entry__.py smoke > "$o/smoke.log" 2>&1
        if [ "${BENCH:-1}" != "0" ]; then
          timeout 300 python bench.py --steps 20 --warmup 5 > "$o/bench_url_like.json" 2> "$o/bench_url_like.err"
          timeout 300 python bench.py --workload int_filter --steps 20 --warmup 5 > "$o/bench_int_filter.json" 2> "$o/bench_int_filter.err"
          timeout 300 python bench.py --workload shipdate --steps 10 --warmup 3 > "$o/bench_shipdate.json" 2> "$o/bench_shipdate.err"
          if python bench.py --help | grep -q squeeze; then
            timeout 300 python bench.py --workload squeeze --rows 20000000 --steps 10 --warmup 3 > "$o/bench_squeeze.json" 2> "$o/bench_squeeze.err"
          fi
        fi
      )
      verdict="$(tail -n 2 "$o/pytest.log" | tr '\n' ' ')"
      line="$(python - "$o" <<'PY'
import json, sys, os
o = sys.argv[1]
parts = []
for name in ("bench_url_like", "bench_int_filter", "bench_shipdate", "bench_squeeze"):
    p = os.path.join(o, name + ".json")
    try:
        j = json.loads(open(p).read().strip().splitlines()[-1])
        r = j.get("roofline")
        r = r[0] if isinstance(r, list) else r
        parts.append(f"{name[6:]}: {j['value']:.0f} {j['unit']} ({j['ms_per_step']:.3f} ms, e2e {j.get('e2e', {}).get('value', 0):.0f}, kernel {r.get('kernel_ms', 0):.3f} ms frac {r.get('frac', 0):.3f})")
    except Exception as e:  # missing or failed run
        parts.append(f"{name[6:]}: -")
print("; ".join(parts))
PY
)"
      echo "$b @ $(cut -c1-10 "$d/VARIANT_COMMIT" 2>/dev/null): $verdict| $line" | tee -a "$out/SUMMARY.txt"
    done
    ;;
  clean)
    rm -rf "$root/variants"
    ;;
  *)
    echo "usage: $0 make <branch>... | run | clean" >&2
    exit 2
    ;;
esac
