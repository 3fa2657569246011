#!/bin/bash
# 8-GPU: bench (all legs except cpu) + config 5
timeout 500 python bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2m_bench8.log 2> gpurun_out/r2m_bench8.err
tail -2 gpurun_out/r2m_bench8.err
python -c "
import json; d=json.loads(open('gpurun_out/r2m_bench8.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step']); print(json.dumps(d['e2e'], indent=1)); print(json.dumps(d['long_audio'], indent=1))"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tools/bench_config5.py > gpurun_out/r2m_c5_8.log 2> gpurun_out/r2m_c5_8.err
tail -2 gpurun_out/r2m_c5_8.err; tail -1 gpurun_out/r2m_c5_8.log
