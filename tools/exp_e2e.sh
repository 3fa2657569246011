for mb in 64 128 256 512; do
  A2SB_E2E_GROUP_MB=$mb python bench.py --steps 10 --warmup 3 --skip-cpu --skip-aligned > gpurun_out/e2e_$mb.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/e2e_$mb.log').read().strip().splitlines()[-1])
print($mb, round(d['e2e']['ms_per_step'],3), round(d['e2e']['value']))
PY
done
