B="python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e"
$B > gpurun_out/z_base.log 2>&1
A2SB_LIB_VARIANT=_pf3 $B > gpurun_out/z_pf3.log 2>&1
for m in 1 3 4 8; do A2SB_INV_M=$m $B > gpurun_out/z_m$m.log 2>&1; done
echo done
