set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r1h_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r1h_ncu_bench.log 2>&1
B=4096 T=4410 ncu --set full --clock-control none --import-source on -k regex:tds_noise_scan -c 1 -o gpurun_out/r1h_noise_scan python tools/tds_quick_bench.py > gpurun_out/r1h_ncu_scan.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:area_function_raster -c 1 -o gpurun_out/r1h_raster python tools/raster_quick_bench.py > gpurun_out/r1h_ncu_raster.log 2>&1
python bench.py --impl reference > gpurun_out/r1h_bench_reference_arm.json 2>/dev/null
ls -la gpurun_out | tail -6
