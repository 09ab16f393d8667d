#!/bin/bash
# Profiling recipe used for the numbers under profiles/ (run through gpurun, 1 GPU; outputs in gpurun_out/).
#   bash profiles/ncu_capture.sh            # everything below
# 1. launch list of the bench's timed region (NVTX range fac_timed), serialised by ncu, clocks untouched
# 2. ncu --set full of two launches of the dominant kernel (one G1 and one G2 of a middle WN layer)
# 3. launch list of one Tacotron2.inference (32 x 690 frames) and ncu --set full of the persistent decoder kernel
P=${1:-bf16x3}
ncu --nvtx --nvtx-include "fac_timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${P}.csv \
    python bench.py --steps 1 --warmup 3 --precision ${P} --no-cpu-baseline --no-ppg2mel > gpurun_out/launches_${P}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wn_gemm_tc -s 20 -c 2 -f -o gpurun_out/prof_${P} \
    python bench.py --steps 1 --warmup 3 --precision ${P} --no-cpu-baseline --no-ppg2mel --batch 8 > gpurun_out/prof_${P}.log 2>&1
ncu -i gpurun_out/prof_${P}.ncu-rep --page raw --csv > gpurun_out/prof_${P}_raw.csv
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tacotron.csv \
    python tools/tacotron_timing.py 32 690 > gpurun_out/launches_tacotron.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:taco_decoder -s 1 -c 1 -f \
    -o gpurun_out/prof_decoder python tools/tacotron_timing.py 8 200 > gpurun_out/prof_decoder.log 2>&1
ncu -i gpurun_out/prof_decoder.ncu-rep --page raw --csv > gpurun_out/prof_decoder_raw.csv
