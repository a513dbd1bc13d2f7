#!/bin/bash
# usage (on the GPU box, through gpurun):
#   scripts/gpu_experiment_sweep.sh "<tag>:<env assignments>" ...
# e.g. scripts/gpu_experiment_sweep.sh "base:PTZBA_NOP=1" "ring:PTZBA_FUSED_RING=1" "pairlist:PTZBA_SCHUR_PAIRLIST=1" \
#                                      "both:PTZBA_FUSED_RING=1 PTZBA_SCHUR_PAIRLIST=1"
# For every configuration: the bundle-adjustment parity tests (kernels vs oracle / goldens, solver vs the reference's
# least_squares, configs 2-5 shapes), then a short bench with the LM iteration and the full solve.  One summary line each.
# Everything is wrapped in `timeout`, so an experiment that hangs costs at most its own limit.
mkdir -p gpurun_out
for spec in "$@"; do
  tag="${spec%%:*}"; envs="${spec#*:}"
  ( export $envs
    t=$(timeout 900 python -m pytest tests/test_gpu_ba.py tests/test_gpu_configs.py -q -x -m gpu 2>&1 | tail -1)
    timeout 400 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-ekf > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
    PTZBA_TRACE=1 timeout 200 python scripts/lm_trace.py > gpurun_out/trace_$tag.log 2>&1
    python - "$tag" "$t" <<'PY'
import json, sys
tag, t = sys.argv[1], sys.argv[2]
try:
    d = json.load(open("gpurun_out/bench_%s.json" % tag))
    print("%-10s step %6.1f us  fused kernels %6.1f us  frac %.3f | LM iter %.3f ms  solve %.2f ms nfev %d | tests: %s" % (
        tag, d["ms_per_step"] * 1e3, d["roofline"]["kernel_ms"] * 1e3, d["roofline"]["frac"], d.get("ms_per_lm_iter", float("nan")),
        d.get("solve", {}).get("ms", float("nan")), d.get("solve", {}).get("nfev", -1), t))
except Exception as e:
    print(tag, "FAILED", e, "| tests:", t)
PY
  )
done
