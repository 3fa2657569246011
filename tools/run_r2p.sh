#!/bin/bash
B="timeout 120 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e --skip-aligned --skip-long"
k=0
for e in "A2SB_LIB_VARIANT=" "A2SB_LIB_VARIANT=_seam63" "A2SB_LIB_VARIANT=" "A2SB_LIB_VARIANT=_seam63"; do
  env $e $B > gpurun_out/r2p_$k.log 2>&1; echo "== $e"; python tools/parse_bench.py gpurun_out/r2p_$k.log; k=$((k+1))
done
