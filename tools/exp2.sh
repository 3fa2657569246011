B="python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e"
$B > gpurun_out/y_base.log 2>&1
for m in 2 4 6 8; do A2SB_INV_TILE=8 A2SB_INV_M=$m $B > gpurun_out/y_f8_m$m.log 2>&1; done
for m in 1 2 3 4; do A2SB_INV_M=$m $B > gpurun_out/y_f16_m$m.log 2>&1; done
A2SB_INV_TILE=8 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
echo done
