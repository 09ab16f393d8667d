timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python __graft_entry__.py 2>&1 | tail -5
timeout 200 python tools/decoder_cycle_breakdown.py 1 690 8 690 16 690 32 690 2>&1 | grep -v Warning > gpurun_out/r2_decoder_cycle_breakdown.txt
grep "us/step" gpurun_out/r2_decoder_cycle_breakdown.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json
timeout 300 python tools/latency_small.py 2>&1 | grep "x " > gpurun_out/r2_latency_small_final.txt; FAC_TC_FUSED=2 timeout 300 python tools/latency_small.py 2>&1 | grep "x " | sed 's/^/flow-step launch: /' >> gpurun_out/r2_latency_small_final.txt; cat gpurun_out/r2_latency_small_final.txt
