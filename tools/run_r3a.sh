#!/bin/bash
# peer-memory long-clip round trip on N GPUs: bit identity + timing vs the NCCL path
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/dist_gpu_check.py > gpurun_out/r3a_dist$N.log 2>&1
echo "rc=$?"; grep -v "^W\|^\[W\|Warning" gpurun_out/r3a_dist$N.log | tail -25
