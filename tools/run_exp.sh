echo "=== 2CTA"; timeout 150 python tools/gemm_check.py 2>&1 | tail -22 | cut -c1-250
echo "exit: $?"
