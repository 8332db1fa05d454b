#!/bin/bash
# ncu evidence of round 2 (run on the GPU box): launch list of a short bench run + --set full of one step's hot kernels
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_under_ncu_r2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'^(k_extrapolate|k_flux|k_update|k_flux_update)$' --launch-skip 6 --launch-count 5 \
    -f -o gpurun_out/prof_r2 python profiles/run_profile.py 2000 2 > gpurun_out/prof_r2.log 2>&1
python profiles/ncu_summary.py gpurun_out/prof_r2.ncu-rep > gpurun_out/ncu_summary_r2.txt 2>&1
cat gpurun_out/ncu_summary_r2.txt
