#!/bin/bash
# K2 with 32-frame tiles for n_fft 512 / 1024 (opt-in): bit identity vs 16-frame tiles + timing
for e in "A2SB_INV_TILE=16" "A2SB_INV_TILE=32" "A2SB_INV_TILE=16" "A2SB_INV_TILE=32"; do
echo "== $e"; env $e python - <<'PY'
import sys, os, torch, hashlib
sys.path.insert(0, os.getcwd())
from audio_intelligence_b200 import _capi, _lib
sys.path.insert(0, "tools")
from bench_nfft import med
g = torch.Generator(device="cuda").manual_seed(3)
wav = (0.3 * torch.randn(256, 441000, device="cuda", generator=g)).clamp_(-1, 1)
out = []
for n in (512, 1024):
    spec = _lib.stft_forward(wav, n, n, n // 4, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25)
    y = _lib.istft_inverse(spec, n, n, n // 4, kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True, power=4.0)
    hsh = hashlib.sha256(y[:4].cpu().numpy().tobytes()).hexdigest()[:10]
    k2 = med(lambda: _lib.istft_inverse(spec, n, n, n // 4, kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True, power=4.0))
    out.append("%d: %.3f (%s)" % (n, k2, hsh))
    del spec, y
print("  K2 ms  " + "   ".join(out))
PY
done
