#!/usr/bin/env bash
# On the GPU box: per-launch durations of the kernels of the bench step (insert kernels filtered out), cold-cache and
# serialised by ncu — shares, not absolutes.   tools/launch_list.sh <tag> [bench args...]
tag="${1:-ll}"; shift || true
out="gpurun_out/$tag"
mkdir -p "$out"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_str_like|k_like_steps|k_scan_plan|k_str_lengths|k_str_decode|k_str_read|k_int_bits|k_int_scan|k_bits|k_and|k_concat|k_build' \
  --launch-skip 12 -c 40 --csv --log-file "$out/launches.csv" python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline "$@" > "$out/launches_bench.log" 2>&1
awk -F'","' 'NR>1{print $5, $9, $NF}' "$out/launches.csv" | tail -24
