# ncu full capture of the forward kernel, carry on (run 3) and off
for v in on off; do
  if [ $v = off ]; then export A2SB_FWD_CARRY=0; else unset A2SB_FWD_CARRY; fi
  ncu --set full --clock-control none --import-source on -k regex:stft_fwd -c 2 -o gpurun_out/prof_k1_$v -f python bench.py --steps 1 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/ncu_k1_$v.log 2>&1
done
echo done
