#!/bin/bash
# seam prefetch flavour (plain / evict_last prefetch / evict_last load) x policy x geometry
for v in "" "_pfel" "_pfld"; do
for e in "A2SB_SEAM=1" "A2SB_SEAM=3" "A2SB_SEAM=1 A2SB_FWD_TILE=32" "A2SB_SEAM=3 A2SB_FWD_TILE=32" "A2SB_SEAM=1 A2SB_FWD_RUN=2" "A2SB_SEAM=1 A2SB_FWD_RUN=4"; do
echo "== variant '$v' $e"; env A2SB_LIB_VARIANT=$v $e python - <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
from audio_intelligence_b200 import _capi, _lib
sys.path.insert(0, "tools")
from bench_nfft import med
wav = (0.3 * torch.randn(256, 441000, device="cuda")).clamp_(-1, 1)
out = []
for n in (512, 1024, 2048, 4096):
    k1 = med(lambda: _lib.stft_forward(wav, n, n, n // 4, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25))
    out.append("%d: %.3f" % (n, k1))
print("  K1 ms  " + "   ".join(out))
PY
done
done
