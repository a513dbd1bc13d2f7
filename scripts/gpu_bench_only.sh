#!/bin/bash
# usage: scripts/gpu_bench_only.sh "<tag>:<env assignments>" ...  - short fused-pass bench only (no parity tests)
for spec in "$@"; do
  tag="${spec%%:*}"; envs="${spec#*:}"
  ( export $envs
    timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-lm --no-ekf > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
    python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open("gpurun_out/bench_%s.json" % tag))
    print("%-14s %6.1f us/step  kernel %6.1f us  frac %.3f" % (tag, d["ms_per_step"] * 1e3, d["roofline"]["kernel_ms"] * 1e3, d["roofline"]["frac"]))
except Exception as e:
    print(tag, "FAILED", e)
PY
  )
done
