#!/bin/bash
# cfg-3 weak-scaling line on N GPUs of one box -> gpurun_out/g_cfg3_n$N.json
N=$1
if [ "$N" -eq 1 ]; then
  python bench.py 2>gpurun_out/g_last.err | grep "^{" > gpurun_out/g_cfg3_n1.json
  python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep "^{" > gpurun_out/g_ref_n1.json
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 200 --warmup 10 --skip-cpu 2>gpurun_out/g_last.err | grep "^{" > gpurun_out/g_cfg3_n$N.json
fi
python - <<PY
import json
d=json.load(open("gpurun_out/g_cfg3_n$N.json"))
print("N=$N", d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"), d.get("replicas_identical"), (d.get("roofline") or {}).get("frac"), d.get("clocks"))
PY
