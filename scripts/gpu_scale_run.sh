#!/bin/bash
# usage (GPU box with N GPUs): scripts/gpu_scale_run.sh <N> <workload> [extra bench flags]  -> gpurun_out/bench_<workload>_<N>gpu.json
N=$1; W=$2; shift 2
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --workload $W --steps 50 --warmup 5 --no-ekf "$@" > gpurun_out/bench_${W}_${N}gpu.json 2> gpurun_out/bench_${W}_${N}gpu.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --workload $W --steps 50 --warmup 5 --no-ekf "$@" > gpurun_out/bench_${W}_${N}gpu.json 2> gpurun_out/bench_${W}_${N}gpu.err
fi
python - "$W" "$N" <<'PY'
import json, sys
w, n = sys.argv[1], sys.argv[2]
try:
    d = json.load(open("gpurun_out/bench_%s_%sgpu.json" % (w, n)))
    print("%s N=%s: pass %.1f us (%.2f G obs/s) kernel %.1f us | e2e %.1f us | LM iter %.3f ms | solve %.1f ms | parity %s | solve parity %s" % (
        w, n, d["ms_per_step"] * 1e3, d["value"] / 1e9, d["roofline"]["kernel_ms"] * 1e3, d["e2e"]["ms_per_step"] * 1e3, d.get("ms_per_lm_iter", float("nan")),
        d.get("solve", {}).get("ms", float("nan")), max(d["parity"]["max_rel_diff"].values()) if d.get("parity") else None,
        (d["solve_parity"]["max_angle_diff_rad"], d["solve_parity"]["max_focal_diff_px"]) if d.get("solve_parity") else None))
except Exception as e:
    print(w, n, "FAILED", e); print(open("gpurun_out/bench_%s_%sgpu.err" % (w, n)).read()[-1500:])
PY
