#!/bin/bash
# usage (GPU box): scripts/gpu_knob_sweep.sh "<tag>:<bench.py flags / ENV=VAL before ::>" ...   e.g. "m57:--lm-share 57" "rand:PTZBA_BENCH_ID_ORDER=random::--lm-share 57"
mkdir -p gpurun_out
for spec in "$@"; do
  tag="${spec%%:*}"; rest="${spec#*:}"
  envs=""; flags="$rest"
  if [[ "$rest" == *"::"* ]]; then envs="${rest%%::*}"; flags="${rest#*::}"; fi
  ( [ -n "$envs" ] && export $envs
    timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-lm --no-ekf $flags > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
    python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open("gpurun_out/bench_%s.json" % tag))
    print("%-14s %6.1f us/step  kernel %6.1f us  frac %.3f  e2e %.1f us" % (tag, d["ms_per_step"] * 1e3, d["roofline"]["kernel_ms"] * 1e3, d["roofline"]["frac"], d["e2e"]["ms_per_step"] * 1e3))
except Exception as e:
    print(tag, "FAILED", e); print(open("gpurun_out/bench_%s.err" % tag).read()[-800:])
PY
  )
done
