#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="timeout 120 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e --skip-aligned"
k=0
for e in "A2SB_INV_TMA=0" "A2SB_INV_TMA=2" "A2SB_INV_TMA=0" "A2SB_INV_TMA=2"; do
  env $e $B > gpurun_out/r2e_$k.log 2>&1; echo "== $e"; python tools/parse_bench.py gpurun_out/r2e_$k.log; k=$((k+1))
done
