set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py > gpurun_out/r1c_bench.json 2> gpurun_out/r1c_bench.err; tail -c 1500 gpurun_out/r1c_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r1c_ncu_bench.log 2>&1
FAST=1 ncu --set full --clock-control none --import-source on -k regex:tl_sweep_fast -c 1 -o gpurun_out/r1c_tl_fast python tools/tl_quick_bench.py > gpurun_out/r1c_ncu_tlf.log 2>&1
B=4096 T=44100 ncu --set full --clock-control none --import-source on -k regex:tds_batch_kernel -c 1 -o gpurun_out/r1c_tds python tools/tds_quick_bench.py > gpurun_out/r1c_ncu_tds.log 2>&1
ls -la gpurun_out
