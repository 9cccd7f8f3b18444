#!/bin/bash
# final multi-GPU measurement series on N GPUs of one box: bench lines into gpurun_out/f_*_n$N.json
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
run() { out=$1; shift; $TR bench.py --gpus $N "$@" 2>gpurun_out/f_last.err | grep "^{" > gpurun_out/$out || tail -5 gpurun_out/f_last.err; }
run f_cfg3_n$N.json --steps 200 --warmup 10 --skip-cpu
run f_cfg3_strong_n$N.json --steps 200 --warmup 10 --skip-cpu --strong
run f_cfg5_n$N.json --workload cfg5 --steps 100 --warmup 10 --skip-cpu
if [ "$N" -le 2 ]; then
  run f_cfg4_n$N.json --workload cfg4 --steps 100 --warmup 10 --skip-cpu
  $TR tests/mp_peer_check.py 2>&1 | grep -E "MP_PEER_CHECK|Error|error" | head -3
fi
python - <<PY
import json
for w in ["cfg3","cfg3_strong","cfg5","cfg4"]:
    try:
        d=json.load(open(f"gpurun_out/f_{w}_n$N.json"))
        r=d.get("roofline") or {}
        print(w, "N=$N", d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"), d.get("replicas_identical"), r.get("frac"), r.get("nvlink_frac"), d.get("scaling"))
    except Exception as e:
        print(w, "ERR", repr(e)[:100])
PY
