#!/bin/bash
# First GPU call of whoever picks this tree up (one GPU, through gpurun; ~6 minutes): everything DESIGN.md section 0 lists as
# "not validated on hardware", in the order that protects the validated evidence.
#   1. smoke() and the 62 validated GPU tests + the 6 written after the last GPU run of round 2 (they sort last)
#   2. the default bench line (new keys: roofline_dense, lm_phases_us, projection, verbatim_reference_cpu) - WITHOUT ncu
#   3. the launch list and one --set full capture of the final fused kernel (none exists of the 39.5 us build) + its DRAM traffic
#   4. cfg5 on one GPU (not re-measured in round 2)
# Outputs in gpurun_out/: pending_*.log / .json, launches_pending.csv, prof_fused_final.ncu-rep
set -u
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/pending_smoke.log 2>&1; tail -1 gpurun_out/pending_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pending_tests.log 2>&1; tail -3 gpurun_out/pending_tests.log
timeout 600 python bench.py > gpurun_out/pending_bench.json 2> gpurun_out/pending_bench.err || { echo "bench failed"; tail -5 gpurun_out/pending_bench.err; }
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/pending_bench.json"))
    print("step %.1f us  kernel %.1f us  frac %.3f | e2e %.1f us | LM %.3f ms | dense %s | projection %s | clocks %s" % (
        d["ms_per_step"] * 1e3, d["roofline"]["kernel_ms"] * 1e3, d["roofline"]["frac"], d["e2e"]["ms_per_step"] * 1e3, d.get("ms_per_lm_iter", float("nan")),
        d.get("roofline_dense"), (d.get("projection") or {}).get("roofline"), d.get("clocks")))
except Exception as e:
    print("no bench line:", e)
PY
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ekf --no-lm"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_pending.csv $B > /dev/null 2> gpurun_out/pending_ncu1.err
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_ba_fused" -s 40 -c 3 -f -o gpurun_out/prof_fused_final $B > /dev/null 2> gpurun_out/pending_ncu2.err
timeout 900 python bench.py --workload cfg5 --steps 30 --warmup 5 --no-ekf --no-cpu-baseline > gpurun_out/pending_bench_cfg5.json 2> gpurun_out/pending_bench_cfg5.err
ls -la gpurun_out | grep -E "pending|prof_fused_final"
