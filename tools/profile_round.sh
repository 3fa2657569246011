#!/bin/bash
# Round profile: plain bench first (numbers that count), then the ncu launch list of the same command, then one
# `--set full` capture of each transform kernel.  Everything lands in gpurun_out/; tools/profile_summarise.py turns
# the reports into the text summaries kept under profiles/.
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/prof_bench.json 2> gpurun_out/prof_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-aligned > gpurun_out/prof_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft_fwd -c 1 -o gpurun_out/prof_full_k1 -f \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-e2e --skip-aligned > gpurun_out/prof_full_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:istft_inv -c 1 -o gpurun_out/prof_full_k2 -f \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-e2e --skip-aligned > gpurun_out/prof_full_k2.log 2>&1
echo done
