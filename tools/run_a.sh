#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_adapt.py tests/test_gpu_stencils.py tests/test_gpu_amr.py tests/test_gpu_fp32.py tests/test_gpu_fish.py tests/test_gpu_poisson.py -q -s 2>&1 | grep -v "OBSTACLE FACTORY" | tail -60) > gpurun_out/ta.log 2>&1
tail -30 gpurun_out/ta.log
(time python bench.py --config amr --steps 5 --warmup 2) > gpurun_out/bench_amr2.json 2> gpurun_out/bench_amr2.err
tail -c 2500 gpurun_out/bench_amr2.json; tail -3 gpurun_out/bench_amr2.err
B="python bench.py --no-parity --no-sweeps --no-amr --no-cpu"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv $B --steps 2 --warmup 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_smooth_tma -s 8 -c 1 -o gpurun_out/r02_smooth -f $B --steps 1 --warmup 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_advdiff_tma -s 2 -c 1 -o gpurun_out/r02_advdiff -f python tools/sweep_bench.py 6 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_prhs_tma -s 2 -c 1 -o gpurun_out/r02_prhs -f python tools/sweep_bench.py 6 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_advdiff_amr -s 2 -c 1 -o gpurun_out/r02_advdiff_amr -f python bench.py --config amr --steps 1 --warmup 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_smooth_amr -s 20 -c 1 -o gpurun_out/r02_smooth_amr -f python bench.py --config amr --steps 1 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches.csv
