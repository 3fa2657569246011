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
# per n_fft family (BASELINE config 4): timings, then one --set full capture of K1 and K2 per family
python tools/bench_nfft.py > gpurun_out/prof_nfft.json 2> gpurun_out/prof_nfft.err
for n in 512 1024 4096; do
  ncu --set full --clock-control none -k regex:stft_fwd --launch-skip 3 -c 1 -o gpurun_out/prof_full_k1_n$n -f \
      python tools/run_one_nfft.py $n > gpurun_out/prof_full_k1_n$n.log 2>&1
  ncu --set full --clock-control none -k regex:istft_inv --launch-skip 3 -c 1 -o gpurun_out/prof_full_k2_n$n -f \
      python tools/run_one_nfft.py $n > gpurun_out/prof_full_k2_n$n.log 2>&1
  for k in k1 k2; do   # keep the raw-metric page only (gpurun_out/ is capped at 64 MiB)
    ncu -i gpurun_out/prof_full_${k}_n$n.ncu-rep --page raw --csv > gpurun_out/prof_full_${k}_n$n.raw.csv 2>/dev/null && rm -f gpurun_out/prof_full_${k}_n$n.ncu-rep
  done
done
python tools/bench_config5.py > gpurun_out/prof_config5_1gpu.json 2> gpurun_out/prof_config5_1gpu.err
echo done
