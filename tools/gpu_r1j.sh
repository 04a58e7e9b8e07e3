set -x
python __graft_entry__.py --smoke 2>&1 | tail -5
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/r1j_bench.json 2> gpurun_out/r1j_bench.err; tail -2 gpurun_out/r1j_bench.err
python bench.py --impl reference > gpurun_out/r1j_bench_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r1j_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r1j_ncu_bench.log 2>&1
T=4410 ncu --set full --clock-control none --import-source on -k regex:tds_batch_kernel --launch-skip 1 -c 1 -o gpurun_out/r1j_tds python tools/tds_quick_bench.py > gpurun_out/r1j_ncu_tds.log 2>&1
T=4410 ncu --set full --clock-control none --import-source on -k regex:tds_batch_kernel --launch-skip 3 -c 1 -o gpurun_out/r1j_tds_noise python tools/tds_quick_bench.py > gpurun_out/r1j_ncu_noise.log 2>&1
FAST=0 ncu --set full --clock-control none --import-source on -k regex:tl_sweep --launch-skip 3 -c 1 -o gpurun_out/r1j_tl python tools/tl_quick_bench.py > gpurun_out/r1j_ncu_tl.log 2>&1
FAST=1 ncu --set full --clock-control none --import-source on -k regex:tl_sweep --launch-skip 3 -c 1 -o gpurun_out/r1j_tl_fast python tools/tl_quick_bench.py > gpurun_out/r1j_ncu_tlf.log 2>&1
# overflow path of the noise-sum ring: ring of 4
AFS_EXTRA_DEFS=-DAFS_NZ_RING=4 python -m areafunctionsynthesis_b200.build --force > /dev/null 2>&1
python -m pytest tests/test_tds_gpu.py -m gpu -q -x -k "noise or per_sample or transvelar" 2>&1 | tail -2
ls -la gpurun_out | tail -12
