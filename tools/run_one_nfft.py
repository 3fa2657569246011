"""Runs K1 and K2 a few times for one n_fft (256 x 10 s clips) -- a target for `ncu -k regex:... --launch-skip N -c 1`.
Usage: python tools/run_one_nfft.py N_FFT [ROW_ALIGN]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_intelligence_b200 import _capi, _lib  # noqa: E402

n = int(sys.argv[1])
align = int(sys.argv[2]) if len(sys.argv) > 2 else None
wav = (0.3 * torch.randn(256, 441000, device="cuda")).clamp_(-1, 1)
for _ in range(4):
    spec = _lib.stft_forward(wav, n, n, n // 4, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25, row_align=align)
    back = _lib.istft_inverse(spec, n, n, n // 4, kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True, power=4.0)
torch.cuda.synchronize()
print("ok", tuple(spec.shape), tuple(back.shape))
