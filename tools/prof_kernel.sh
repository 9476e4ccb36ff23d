#!/usr/bin/env bash
# On the GPU box: one `ncu --set full` capture of a kernel of the bench step, details + source pages exported as text.
#   tools/prof_kernel.sh <tag> <kernel regex> <launch-skip> [bench args...]
#   -> gpurun_out/<tag>/prof_<name>.{ncu-rep,details.txt,source.csv,raw.csv}
tag="$1"; rx="$2"; skip="$3"; shift 3
out="gpurun_out/$tag"
name=$(echo "$rx" | tr -c 'A-Za-z0-9_' '_' | cut -c1-40)
mkdir -p "$out"
ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip "$skip" -c 1 -f -o "$out/prof_$name" \
  python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline "$@" > "$out/prof_${name}_bench.log" 2>&1
ncu -i "$out/prof_$name.ncu-rep" --page details > "$out/prof_$name.details.txt" 2>&1
ncu -i "$out/prof_$name.ncu-rep" --page source --csv > "$out/prof_$name.source.csv" 2>&1
ncu -i "$out/prof_$name.ncu-rep" --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum > "$out/prof_$name.raw.csv" 2>&1
tail -2 "$out/prof_$name.raw.csv"
