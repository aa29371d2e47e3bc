#!/bin/bash
# 1-GPU evidence run: the complete bench line, the ncu launch list and full captures summarised ON the box
# (gpurun brings back at most 64 MiB: only the JSON summaries and one report travel)
mkdir -p gpurun_out /tmp/rep
(timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_amr.py tests/test_gpu_poisson.py tests/test_gpu_fp32.py -q 2>&1 | tail -12) > gpurun_out/tc.log 2>&1
tail -4 gpurun_out/tc.log
(time python bench.py) > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
tail -c 1500 gpurun_out/bench_c.json
B="python bench.py --no-parity --no-sweeps --no-amr --no-cpu"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv $B --steps 2 --warmup 1 > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 2600 --csv --log-file gpurun_out/r02_launches_amr.csv python bench.py --config amr --steps 1 --warmup 1 > /dev/null 2>&1
cap() {  # name kernel-regex skip command...
  local name=$1 re=$2 skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c 1 -o /tmp/rep/$name -f "$@" > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/rep/$name.ncu-rep > gpurun_out/r02_${name}_ncu_summary.json 2> /dev/null
}
cap smooth_tma k_smooth_tma 8 $B --steps 1 --warmup 1
cap advdiff_tma k_advdiff_tma 2 python tools/sweep_bench.py 6
cap prhs_tma k_prhs_tma 2 python tools/sweep_bench.py 6
cap advdiff_amr k_advdiff_amr 2 python bench.py --config amr --steps 1 --warmup 1
cap smooth_amr k_smooth_amr 20 python bench.py --config amr --steps 1 --warmup 1
cap apply_amr k_apply_amr 10 python bench.py --config amr --steps 1 --warmup 1
cp /tmp/rep/smooth_tma.ncu-rep gpurun_out/r02_smooth_tma.ncu-rep
ls -la gpurun_out /tmp/rep
