#!/bin/bash
# ncu DRAM traffic of the two-round 32-frame K1 (n_fft 2048) with the wide pass B, head-only vs head+tail seam prefetch
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for s in 1 3; do
  A2SB_FWD_TILE=32 A2SB_FWD_WIDE=1 A2SB_SEAM=$s timeout 300 ncu --metrics $M --clock-control none -k regex:stft_fwd --launch-skip 3 -c 1 --csv --log-file gpurun_out/r3n_wide_seam$s.csv python tools/run_one_nfft.py 2048 > gpurun_out/r3n_wide_seam$s.log 2>&1
  echo "wide two-round seam=$s rc=$?"
  python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/r3n_wide_seam$s.csv")) if len(r) > 10]
for r in rows[1:]:
    print("  ", r[-3][:30], r[-2], r[-1])
PY
done
