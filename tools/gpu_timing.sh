python -m pytest tests/test_tds_gpu.py -m gpu -q -x -k "noise" 2>&1 | tail -1
T=8820 python tools/tds_quick_bench.py 2>&1 | tail -1
cp areafunctionsynthesis_b200/libafs_b200.so /tmp/lib_keep.so
for d in -DAFS_SCAN_THREADS=768 -DAFS_SCAN_THREADS=1024; do
AFS_EXTRA_DEFS="$d" python -m areafunctionsynthesis_b200.build --force > /dev/null 2>&1
echo "$d"; grep -E "tds_noise_scan_kernel" -A2 areafunctionsynthesis_b200/build/build.log | grep -E "Used|spill"
T=8820 python tools/tds_quick_bench.py 2>&1 | tail -1
done
python -m pytest tests/test_tds_gpu.py -m gpu -q -x -k "noise" 2>&1 | tail -1
cp /tmp/lib_keep.so areafunctionsynthesis_b200/libafs_b200.so
