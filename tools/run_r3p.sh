#!/bin/bash
# K1 n_fft 2048: streaming (st.cs) vs plain stores for the 16-frame and the wide two-round 32-frame kernels
for v in "" "_plain"; do
for g in "A2SB_FWD_TILE=16" "A2SB_FWD_TILE=32 A2SB_FWD_WIDE=1 A2SB_SEAM=1" "A2SB_FWD_TILE=32 A2SB_FWD_WIDE=1 A2SB_SEAM=3"; do
echo "== variant '$v' $g"; env A2SB_LIB_VARIANT=$v $g timeout 120 python - <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
from audio_intelligence_b200 import _capi, _lib
sys.path.insert(0, "tools")
from bench_nfft import med
wav = (0.3 * torch.randn(256, 441000, device="cuda")).clamp_(-1, 1)
out = []
for n in (2048, 1024, 4096):
    k1 = med(lambda: _lib.stft_forward(wav, n, n, n // 4, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25))
    out.append("%d: %.3f" % (n, k1))
print("  K1 ms  " + "   ".join(out))
PY
done
done
