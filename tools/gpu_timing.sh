cp areafunctionsynthesis_b200/libafs_b200.so /tmp/lib_keep.so
for d in -DAFS_NZ_DECAY_WARP=0; do
AFS_EXTRA_DEFS="$d" python -m areafunctionsynthesis_b200.build --force > /dev/null 2>&1
echo "$d"; T=8820 python tools/tds_quick_bench.py 2>&1 | tail -1
python -m pytest tests/test_tds_gpu.py -m gpu -q -x -k "noise" 2>&1 | tail -1
done
cp /tmp/lib_keep.so areafunctionsynthesis_b200/libafs_b200.so
