python -m pytest tests/test_tl_gpu.py tests/test_lf_gpu.py tests/test_optimizer_gpu.py tests/test_sweep_gpu.py -m gpu -q -x 2>&1 | tail -3
FAST=0 python tools/tl_quick_bench.py 2>&1 | head -1
