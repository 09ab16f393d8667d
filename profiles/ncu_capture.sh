#!/bin/bash
# Profiling recipe used for the numbers under profiles/ (run through gpurun, 1 GPU).
#   launch list : ncu --nvtx --nvtx-include "fac_timed/" --metrics gpu__time_duration.sum --clock-control none \
#                     --csv --log-file gpurun_out/launches_<tag>.csv python bench.py --steps 1 --warmup 3 --precision <p> --no-cpu-baseline
#   full capture: the command below (2 launches of the dominant kernel: one G1 and one G2 of a middle layer)
set -e
P=${1:-bf16x3}
ncu --set full --clock-control none --import-source on -k regex:wn_gemm_tc -s 20 -c 2 -f -o gpurun_out/prof_${P} \
    python bench.py --steps 1 --warmup 3 --precision ${P} --no-cpu-baseline --batch 8 > gpurun_out/prof_${P}.log 2>&1
ncu -i gpurun_out/prof_${P}.ncu-rep --page raw --csv > gpurun_out/prof_${P}_raw.csv
