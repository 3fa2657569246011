ncu --set full --clock-control none --import-source on -k regex:istft_inv -c 1 -o gpurun_out/prof_k2 -f python bench.py --steps 1 --warmup 3 --skip-cpu --skip-e2e --skip-aligned > gpurun_out/ncu_k2.log 2>&1
echo done
