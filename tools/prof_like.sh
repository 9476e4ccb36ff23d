#!/usr/bin/env bash
# On the GPU box: one `ncu --set full` capture of the shipped k_str_like launch of the bench step (the fourth launch: the
# first is the untimed counter launch, then warm-ups), details + source pages exported as text beside the report.
#   tools/prof_like.sh <tag>      -> gpurun_out/<tag>/prof_str_like.{ncu-rep,details.txt,source.csv}
tag="${1:-prof}"
out="gpurun_out/$tag"
mkdir -p "$out"
ncu --set full --clock-control none --import-source on -k regex:k_str_like --launch-skip 3 -c 1 -f -o "$out/prof_str_like" \
  python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline > "$out/prof_bench.log" 2>&1
ncu -i "$out/prof_str_like.ncu-rep" --page details > "$out/prof_str_like.details.txt" 2>&1
ncu -i "$out/prof_str_like.ncu-rep" --page source --csv > "$out/prof_str_like.source.csv" 2>&1
ncu -i "$out/prof_str_like.ncu-rep" --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum > "$out/prof_str_like.raw.csv" 2>&1
tail -5 "$out/prof_str_like.raw.csv"
