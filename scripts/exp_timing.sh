VSP_DEBUG_TIMING=1 timeout 300 python bench.py --steps 1 --warmup 3 --ckpts 9 --no-cpu-baseline 2>/dev/null | grep tridiag_fused | tail -4
