#!/bin/bash
# ncu DRAM / L2 traffic of K1 at n_fft 2048: two-round 32-frame tiles vs 16-frame tiles
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum"
for tile in 32 16; do
  A2SB_FWD_TILE=$tile timeout 300 ncu --metrics $M --clock-control none -k regex:stft_fwd --launch-skip 3 -c 1 --csv --log-file gpurun_out/r2v_tile$tile.csv python tools/run_one_nfft.py 2048 > gpurun_out/r2v_tile$tile.log 2>&1
  echo "tile $tile rc=$?"
  python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2v_tile$tile.csv")) if len(r) > 10]
for r in rows[1:]:
    print(r[-3][:45], r[-2], r[-1])
PY
done
