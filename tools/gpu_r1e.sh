set -x
python __graft_entry__.py --smoke 2>&1 | tail -5
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py > gpurun_out/r1e_bench.json 2> gpurun_out/r1e_bench.err; tail -2 gpurun_out/r1e_bench.err
python bench.py --impl reference > gpurun_out/r1e_bench_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1e_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r1e_ncu_bench.log 2>&1
FAST=1 ncu --set full --clock-control none --import-source on -k regex:tl_sweep_fast -c 1 -o gpurun_out/r1e_tl_fast python tools/tl_quick_bench.py > gpurun_out/r1e_ncu_tlf.log 2>&1
ls -la gpurun_out | tail -8
