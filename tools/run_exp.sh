for t in 0 1; do echo "=== GCBF_TWO_STREAMS=$t"; GCBF_TWO_STREAMS=$t python tools/dp_timing.py 2>&1 | grep -E "plain "; done
GCBF_TWO_STREAMS=1 python -m pytest tests/test_parity_gpu.py -q -x 2>&1 | tail -2
