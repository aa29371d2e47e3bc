#!/bin/bash
# 2-GPU check of the staged NVLink push: the bench with its parity check, then the device-timestamp trace
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29511 bench.py --gpus 2 > gpurun_out/s2.json 2> gpurun_out/s2.err
CUP_STAMP=1 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --no-parity --steps 5 > /dev/null 2> gpurun_out/s2_stamp.err
for v in "CUP_PUSH=1" "CUP_PUSH=0" "CUP_FUSED=0"; do
  env $v $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --no-parity > gpurun_out/s2_v.json 2> gpurun_out/s2_v.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/s2_v.json") if l.startswith("{")][-1])
print("$v", d["ms_per_step"], d["fingerprint"]["pois_dot_zz"])
PY
done
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/s2.json") if l.startswith("{")][-1])
print({k: d.get(k) for k in ("n_gpus", "ms_per_step", "value", "fingerprint")}, d["parity"]["rel_err"])
PY
grep "stamp rank" gpurun_out/s2_stamp.err | cut -c1-1200
