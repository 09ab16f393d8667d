timeout 400 python -m pytest tests/test_tacotron_gpu.py -q -x 2>&1 | tail -2
timeout 200 python tools/decoder_cycle_breakdown.py 1 690 8 690 16 690 32 690 2>&1 | grep -E "us/step"
