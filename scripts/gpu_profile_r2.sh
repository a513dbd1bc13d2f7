#!/bin/bash
# Round-2 profile set (one GPU, through gpurun).  Every program runs once WITHOUT ncu first (numbers printed under a profiler are never
# bench values), then under ncu.  Outputs in gpurun_out/: *.json (plain runs), launches_*.csv, prof_*.ncu-rep.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ekf"
timeout 400 $B > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err || { echo "bench failed"; tail -5 gpurun_out/p_bench.err; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bench_r2.csv $B > /dev/null 2> gpurun_out/p_ncu1.err
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_ba_fused" -s 40 -c 3 -f -o gpurun_out/prof_fused_r2 $B > /dev/null 2> gpurun_out/p_ncu2.err
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_schur_pairlist|k_potrf_coop" -s 2 -c 2 -f -o gpurun_out/prof_solver_r2 $B > /dev/null 2> gpurun_out/p_ncu3.err
timeout 300 python scripts/proj_bench.py > gpurun_out/p_proj.json 2> gpurun_out/p_proj.err && \
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_project_grid" -s 6 -c 1 -f -o gpurun_out/prof_proj_r2 python scripts/proj_bench.py > /dev/null 2> gpurun_out/p_ncu4.err
E="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-lm --ekf-seqs 64 --ekf-frames 3"
timeout 400 $E > gpurun_out/p_ekf.json 2> gpurun_out/p_ekf.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_ekf64_r2.csv $E > /dev/null 2> gpurun_out/p_ncu5.err
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_lu_rows_ll|k_chol_ll_update|k_lu_gemm" -s 200 -c 3 -f -o gpurun_out/prof_ekf_r2 $E > /dev/null 2> gpurun_out/p_ncu6.err
ls -la gpurun_out | grep -E "prof_|launches_|p_" 
