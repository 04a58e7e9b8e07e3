python -m pytest tests/test_tds_gpu.py -m gpu -q -x 2>&1 | tail -3
T=8820 python tools/tds_quick_bench.py 2>&1 | tail -2
cp areafunctionsynthesis_b200/libafs_b200.so /tmp/lib_keep.so
AFS_EXTRA_DEFS=-DAFS_NZ_RING=2 python -m areafunctionsynthesis_b200.build --force > /dev/null 2>&1
python -m pytest tests/test_tds_gpu.py -m gpu -q -x -k "noise or per_sample or transvelar" 2>&1 | tail -2
cp /tmp/lib_keep.so areafunctionsynthesis_b200/libafs_b200.so
