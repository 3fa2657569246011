B="python bench.py --steps 20 --warmup 5 --skip-cpu --skip-e2e"
A2SB_LIB_VARIANT=_ct862 $B > gpurun_out/x_ct862.log 2>&1
for r in 1 2 3 6 9 18 27 54; do A2SB_BENCH_CLIP_LEN=441856 A2SB_FWD_RUN=$r $B > gpurun_out/x_al_run$r.log 2>&1; done
$B > gpurun_out/x_base.log 2>&1
echo done
