#!/bin/bash
# multi-GPU measurement session on one box (run through gpurun --gpus 8): correctness first, then the
# strong-scaling bench with its live parity check, threshold variants and device-timestamp traces
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(time timeout 900 python -m pytest tests/test_gpu_multi.py -q 2>&1 | grep -v "OBSTACLE" | tail -25) > gpurun_out/t8_multi.log 2>&1
tail -8 gpurun_out/t8_multi.log
$TR --nproc-per-node 8 --master-port 29501 bench.py --gpus 8 > gpurun_out/s8.json 2> gpurun_out/s8.err
CUP_COARSE_BLOCKS=512 $TR --nproc-per-node 8 --master-port 29502 bench.py --gpus 8 --no-parity > gpurun_out/s8_c512.json 2> gpurun_out/s8_c512.err
CUP_COARSE_BLOCKS=64 $TR --nproc-per-node 8 --master-port 29503 bench.py --gpus 8 --no-parity > gpurun_out/s8_c64.json 2> gpurun_out/s8_c64.err
CUP_STAMP=1 $TR --nproc-per-node 8 --master-port 29504 bench.py --gpus 8 --no-parity --steps 5 > /dev/null 2> gpurun_out/s8_stamp.err
CUP_STAMP=1 CUP_COARSE_BLOCKS=512 $TR --nproc-per-node 8 --master-port 29505 bench.py --gpus 8 --no-parity --steps 5 > /dev/null 2> gpurun_out/s8_c512_stamp.err
$TR --nproc-per-node 8 --master-port 29506 bench.py --gpus 8 --config amr --steps 5 --warmup 2 > gpurun_out/s8_amr.json 2> gpurun_out/s8_amr.err
for f in s8 s8_c512 s8_c64 s8_amr; do echo "== $f"; python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/$f.json") if l.startswith("{")][-1])
    print({k: d.get(k) for k in ("n_gpus", "ms_per_step", "value", "parity", "fingerprint", "phases_ms")})
except Exception as ex:
    print("no line:", ex)
PY
tail -3 gpurun_out/$f.err; done
grep "stamp rank 0\|stamp rank 7" gpurun_out/s8_stamp.err gpurun_out/s8_c512_stamp.err | cut -c1-1500
