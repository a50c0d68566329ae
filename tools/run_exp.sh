for c in 0 1; do echo "=== GCBF_TC_2CTA=$c"; GCBF_TC_2CTA=$c timeout 150 python tools/gemm_check.py 2>&1 | grep -E "^\[|FAILED|Error" | cut -c1-250; done
