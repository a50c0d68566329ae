for d in 0 1; do echo "=== GCBF_TC_DBG=$d"; GCBF_TC_DBG=$d timeout 200 python tools/gemm_check.py 2>&1 | grep -E "^\[|fwd with"; done
