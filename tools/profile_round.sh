#!/bin/bash
# Round profile: plain bench first (numbers that count), then the ncu launch list of the same command, then one
# `--set full` capture of each transform kernel, then a metrics pass over the streaming kernels at config-3 size.
# Everything lands in gpurun_out/; tools/profile_summarise.py turns the reports into the summaries kept under profiles/.
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/prof_bench.json 2> gpurun_out/prof_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-aligned --skip-long > gpurun_out/prof_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft_fwd -c 1 -o gpurun_out/prof_full_k1 -f \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-e2e --skip-aligned --skip-long > gpurun_out/prof_full_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:istft_inv -c 1 -o gpurun_out/prof_full_k2 -f \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-e2e --skip-aligned --skip-long > gpurun_out/prof_full_k2.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"segment_|mask_fill|wrap_pad" -c 120 --csv --log-file gpurun_out/prof_segments.csv \
    python tools/bench_segments.py > gpurun_out/prof_segments.log 2>&1
python tools/bench_config5.py > gpurun_out/prof_config5_1gpu.json 2> gpurun_out/prof_config5_1gpu.err
echo done
