#!/bin/bash
# bench.py on N GPUs with the long-audio leg (NCCL path + peer-memory path)
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 --skip-cpu > gpurun_out/r3b_bench$N.log 2> gpurun_out/r3b_bench$N.err
echo "rc=$?"; tail -3 gpurun_out/r3b_bench$N.err | cut -c1-300
python - <<PY
import json
for ln in open("gpurun_out/r3b_bench$N.log"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", d.get("e2e", {}).get("value"))
        la = d.get("long_audio", {})
        for k in ("round_trip", "round_trip_peer", "blend_step"):
            print(k, json.dumps({a: b for a, b in la.get(k, {}).items() if not isinstance(b, str) or a in ("store", "unavailable")}))
PY
