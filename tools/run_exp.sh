echo "=== 2CTA"; GCBF_TC_2CTA=1 timeout 150 python tools/gemm_check.py 2>&1 | tail -22 | cut -c1-250
echo "exit: $?"
