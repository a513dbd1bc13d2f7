#!/bin/bash
# usage (on the GPU box, one GPU, through gpurun):  scripts/gpu_ncu_capture.sh <tag> <kernel regex> [ENV=VALUE ...]
# e.g.  scripts/gpu_ncu_capture.sh fused 'k_ba_fused'
#       scripts/gpu_ncu_capture.sh pairs 'k_schur_pair'
# 1. runs the short bench WITHOUT ncu first (a number printed under a profiler is never a bench value), 2. takes the launch list
# of the same command, 3. takes one `--set full` capture of the kernels matching the regex (after 40 matching launches of warm-up).
# Outputs: gpurun_out/bench_<tag>.json, launches_<tag>.csv, prof_<tag>.ncu-rep  (read here with `ncu -i ... --page raw --csv`).
set -u
tag="$1"; regex="$2"; shift 2
for kv in "$@"; do export "$kv"; done
mkdir -p gpurun_out
cmd="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ekf"
timeout 400 $cmd > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { echo "bench failed"; tail -5 gpurun_out/bench_$tag.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv $cmd > /dev/null 2> gpurun_out/ncu_launches_$tag.err
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s 40 -c 4 -f -o gpurun_out/prof_$tag $cmd > /dev/null 2> gpurun_out/ncu_full_$tag.err
ls -la gpurun_out/ | grep "$tag"
