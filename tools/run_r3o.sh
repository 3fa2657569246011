#!/bin/bash
# K1 n_fft 2048: neighbouring CTAs kept in lockstep by a cluster barrier per round (A2SB_FWD_CLUSTER), all geometries
for g in "A2SB_FWD_TILE=16" "A2SB_FWD_TILE=32 A2SB_FWD_WIDE=1 A2SB_SEAM=1" "A2SB_FWD_TILE=32 A2SB_FWD_WIDE=1 A2SB_SEAM=3"; do
for c in 0 2 4; do
echo "== $g cluster=$c"; env $g A2SB_FWD_CLUSTER=$c timeout 120 python - <<'PY'
import sys, os, torch, hashlib
sys.path.insert(0, os.getcwd())
from audio_intelligence_b200 import _capi, _lib
sys.path.insert(0, "tools")
from bench_nfft import med
g = torch.Generator(device="cuda").manual_seed(3)
wav = (0.3 * torch.randn(256, 441000, device="cuda", generator=g)).clamp_(-1, 1)
out = []
for n in (2048, 4096):
    s = _lib.stft_forward(wav, n, n, n // 4, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25)
    torch.cuda.synchronize()
    hsh = hashlib.sha256(s[:2].cpu().numpy().tobytes()).hexdigest()[:8]
    del s
    k1 = med(lambda: _lib.stft_forward(wav, n, n, n // 4, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25))
    out.append("%d: %.3f (%s)" % (n, k1, hsh))
print("  K1 ms  " + "   ".join(out))
PY
done
done
