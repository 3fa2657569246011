#!/bin/bash
# K2 at n_fft 2048 / 1024 with 8-frame tiles (two CTAs per SM) now that the 8-frame kernel's work item is two tiles
for e in "A2SB_INV_TILE=16" "A2SB_INV_TILE=8" "A2SB_INV_TILE=8 A2SB_INV_M=3" "A2SB_INV_TILE=8 A2SB_INV_M=4" "A2SB_INV_TILE=8 A2SB_INV_M=1"; do
echo "== $e"; env $e python - <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
from audio_intelligence_b200 import _capi, _lib
sys.path.insert(0, "tools")
from bench_nfft import med
wav = (0.3 * torch.randn(256, 441000, device="cuda")).clamp_(-1, 1)
out = []
for n in (2048, 1024):
    spec = _lib.stft_forward(wav, n, n, n // 4, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25)
    k2 = med(lambda: _lib.istft_inverse(spec, n, n, n // 4, kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True, power=4.0))
    out.append("%d: %.3f" % (n, k2))
    del spec
print("  K2 ms  " + "   ".join(out))
PY
done
