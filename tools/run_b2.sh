#!/bin/bash
# last 2-GPU check of the round: multi-rank AMR tests (pois_op split off), then the face-push variants
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 200 python -m pytest tests/test_gpu_multi.py -q -k "amr2-2 or 8-3-None or amr3-4" 2>&1 | tail -4) > gpurun_out/tb2.log 2>&1
tail -3 gpurun_out/tb2.log
$TR --nproc-per-node 2 --master-port 29521 bench.py --gpus 2 > gpurun_out/s2_tma.json 2> gpurun_out/s2_tma.err
CUP_PUSH=1 $TR --nproc-per-node 2 --master-port 29522 bench.py --gpus 2 --no-parity > gpurun_out/s2_p1.json 2> gpurun_out/s2_p1.err
CUP_STAMP=1 $TR --nproc-per-node 2 --master-port 29523 bench.py --gpus 2 --no-parity --steps 5 > /dev/null 2> gpurun_out/s2_tma_stamp.err
for f in s2_tma s2_p1; do python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/$f.json") if l.startswith("{")][-1])
print("$f", d["ms_per_step"], d["parity"].get("rel_err"), d["fingerprint"]["pois_dot_zz"])
PY
done
grep "stamp rank 0" gpurun_out/s2_tma_stamp.err | cut -c1-900
