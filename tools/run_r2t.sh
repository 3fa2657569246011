#!/bin/bash
# round 2: two-round 32-frame K1 for n_fft 2048 -- bit-identity vs 16-frame tiles, GPU tests, headline A/B
python tools/check_fwd_tiles.py 512 1024 2048; echo "tiles rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2t_tests.log
B="timeout 120 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e --skip-long"
k=0
for e in "A2SB_FWD_TILE=32" "A2SB_FWD_TILE=16" "A2SB_FWD_TILE=32" "A2SB_FWD_TILE=16"; do
  env $e $B > gpurun_out/r2t_$k.log 2>&1; echo "== $e"; python tools/parse_bench.py gpurun_out/r2t_$k.log; k=$((k+1))
done
grep -o '"aligned_rows_variant": {[^}]*}' gpurun_out/r2t_0.log
