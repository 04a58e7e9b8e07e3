cp areafunctionsynthesis_b200/libafs_b200.so /tmp/lib_keep.so
AFS_EXTRA_DEFS=-DAFS_NZ_NOLOOP python -m areafunctionsynthesis_b200.build --force > /dev/null 2>&1
T=8820 python tools/tds_quick_bench.py 2>&1 | tail -1
cp /tmp/lib_keep.so areafunctionsynthesis_b200/libafs_b200.so
