set -x
python __graft_entry__.py --smoke 2>&1 | tail -3
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/r1l_bench.json 2> gpurun_out/r1l_bench.err; tail -2 gpurun_out/r1l_bench.err
python bench.py --impl reference > gpurun_out/r1l_bench_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r1l_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r1l_ncu_bench.log 2>&1
ls -la gpurun_out | tail -5
