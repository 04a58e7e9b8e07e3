python -m pytest tests/test_tds_gpu.py -m gpu -q -x 2>&1 | tail -3
T=8820 python tools/tds_quick_bench.py 2>&1 | tail -2
