#!/bin/bash
# usage: scripts/gpu_variant_sweep.sh "<tag>:<env assignments>" ...   (run on the GPU box through gpurun)
# runs test_gpu_ba.py parity and a short fused-pass bench for every configuration, prints one summary line each
for spec in "$@"; do
  tag="${spec%%:*}"; envs="${spec#*:}"
  ( export $envs
    t=$(timeout 600 python -m pytest tests/test_gpu_ba.py -q -x 2>&1 | tail -1)
    timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-lm --no-ekf > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
    python - "$tag" "$t" <<'PY'
import json, sys
tag, t = sys.argv[1], sys.argv[2]
try:
    d = json.load(open("gpurun_out/bench_%s.json" % tag))
    print("%-14s %6.1f us/step  kernel %6.1f us  frac %.3f | tests: %s" % (tag, d["ms_per_step"] * 1e3, d["roofline"]["kernel_ms"] * 1e3, d["roofline"]["frac"], t))
except Exception as e:
    print(tag, "FAILED", e, "| tests:", t)
PY
  )
done
