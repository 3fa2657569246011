for i in 1 2 3 4; do python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e --skip-aligned > gpurun_out/v_$i.log 2>&1; done
python - <<'PY'
import json
for i in range(1,5):
    d=json.loads(open(f'gpurun_out/v_{i}.log').read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]
    print(i, [round(x,3) for x in k["stft_fwd_kernel"]["ms_min_median_max"]], [round(x,3) for x in k["istft_inv_kernel"]["ms_min_median_max"]])
PY
