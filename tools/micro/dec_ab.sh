timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 200 python tools/decoder_cycle_breakdown.py 1 690 8 690 16 690 32 690 2>&1 | grep -v Warning > gpurun_out/r2_decoder_cycle_breakdown.txt
cat gpurun_out/r2_decoder_cycle_breakdown.txt | grep "us/step"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_decoder.json 2> gpurun_out/bench_decoder.err; tail -c 3000 gpurun_out/bench_decoder.json
