#!/bin/bash
# round 2: seam-prefetch policy sweep for the two-round 32-frame K1 (n_fft 2048)
B="timeout 120 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e --skip-long --skip-aligned"
for e in "A2SB_SEAM=1" "A2SB_SEAM=3" "A2SB_SEAM=7" "A2SB_SEAM=5" "A2SB_FWD_TILE=16"; do
  env $e $B > gpurun_out/r2u_$e.log 2>&1; echo "== $e"; python tools/parse_bench.py gpurun_out/r2u_$e.log
done
