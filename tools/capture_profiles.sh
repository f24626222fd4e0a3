#!/bin/bash
# Runs on the GPU box (gpurun): the round's evidence for the headline kernel, every step after its plain run.
#   bash tools/capture_profiles.sh        -> gpurun_out/final_*
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err || exit 1
# launch list of the same command (cold-cache, serialised per-launch times: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/final_launches_run.log 2>&1
# one full capture of the headline kernel
python tools/quick_bench.py --steps 20 > gpurun_out/final_quick.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:solve_standard_tm -s 8 -c 1 -o gpurun_out/final_tm -f \
    python tools/quick_bench.py --steps 6 > gpurun_out/final_tm_ncu.log 2>&1
tail -2 gpurun_out/final_tm_ncu.log
cat gpurun_out/final_quick.log
