FAC_TACO_FLAGS=0 timeout 100 python tools/decoder_cycle_breakdown.py 8 690 2>&1 | grep -v Warning
