#!/bin/bash
# n_fft 4096 inverse: DRAM over-fetch (9.6 GB read for 2.7 GB) -- L2 promotion size x tiles per work item
for v in "" "_ld4" "_ld0"; do
for m in 4 2 1 3 8; do
echo "== variant '$v' A2SB_INV_M=$m"; env A2SB_LIB_VARIANT=$v A2SB_INV_M=$m python - <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
from audio_intelligence_b200 import _capi, _lib
sys.path.insert(0, "tools")
from bench_nfft import med
wav = (0.3 * torch.randn(256, 441000, device="cuda")).clamp_(-1, 1)
n = 4096
spec = _lib.stft_forward(wav, n, n, n // 4, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25)
k2 = med(lambda: _lib.istft_inverse(spec, n, n, n // 4, kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True, power=4.0))
print("  K2-4096 %.3f ms" % k2)
PY
done
done
