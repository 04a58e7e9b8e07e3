FAST=0 python tools/tl_quick_bench.py 2>&1 | head -1
