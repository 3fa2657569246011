#!/bin/bash
# round 2: WIDE pass B (128-byte row-segment stores) -- bit-identity, GPU tests, per-n_fft timings, headline A/B
python tools/check_fwd_tiles.py 512 1024 2048; echo "tiles rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2y_tests.log
for e in "A2SB_FWD_WIDE=1" "A2SB_FWD_WIDE=0"; do
echo "== $e"; env $e python - <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
from audio_intelligence_b200 import _capi, _lib
sys.path.insert(0, "tools")
from bench_nfft import med
wav = (0.3 * torch.randn(256, 441000, device="cuda")).clamp_(-1, 1)
for n in (512, 1024):
    k1 = med(lambda: _lib.stft_forward(wav, n, n, n // 4, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25))
    k1p = med(lambda: _lib.stft_forward(wav, n, n, n // 4, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25, row_align=8))
    print(n, "K1 %.3f ms  pitched %.3f" % (k1, k1p))
PY
done
B="timeout 120 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e --skip-long"
for e in "A2SB_FWD_TILE=32 A2SB_FWD_WIDE=1" "A2SB_FWD_TILE=32 A2SB_FWD_WIDE=0" "A2SB_FWD_TILE=16" "A2SB_FWD_TILE=32 A2SB_FWD_WIDE=1 A2SB_SEAM=3"; do
  f=gpurun_out/r2y_$(echo $e | tr ' =' '__').log
  env $e $B > $f 2>&1; echo "== $e"; python tools/parse_bench.py $f; grep -o '"stft_fwd_kernel_ms": [0-9.]*' $f
done
