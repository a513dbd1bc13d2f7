#!/bin/bash
# usage (GPU box): scripts/gpu_lib_sweep.sh [--tests] <tag> ...   (tag "default" = csrc/libptzba.so, else csrc/libptzba_<tag>.so,
# built here with PTZBA_BUILD_TAG=<tag> PTZBA_EXTRA_NVCC_FLAGS=... python pan-tilt-zoom-slam_b200/build.py)
# Per library variant: optionally the BA parity tests, then the short fused-pass / LM bench.  One summary line each.
mkdir -p gpurun_out
TESTS=0
if [ "$1" = "--tests" ]; then TESTS=1; shift; fi
for tag in "$@"; do
  ( if [ "$tag" != "default" ]; then export PTZBA_LIBRARY=$PWD/pan-tilt-zoom-slam_b200/csrc/libptzba_$tag.so; fi
    t="-"
    if [ $TESTS = 1 ]; then t=$(timeout 900 python -m pytest tests/test_gpu_ba.py tests/test_gpu_configs.py -q -x -m gpu 2>&1 | tail -1); fi
    timeout 400 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-ekf > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
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
    print(open("gpurun_out/bench_%s.err" % tag).read()[-1500:])
PY
  )
done
