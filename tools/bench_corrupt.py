"""Dataset-side path (A2SB/datasets/datasets.py:235-237): forward chain + mask transform, the two-call form (K1, then the
one-pass mask fill) against K1 with the corruption in its epilogue + the rectangle-mask kernel.  256 x 10 s clips, n_fft 2048.
Usage: python tools/bench_corrupt.py   (needs a GPU)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_intelligence_b200 import _capi, _lib  # noqa: E402
from bench_nfft import med  # noqa: E402


def main():
    B, L, n, hop = 256, 441000, 2048, 512
    wav = (0.3 * torch.randn(B, L, device="cuda")).clamp_(-1, 1)
    T = 1 + L // hop
    noise = torch.randn(B, 3, n // 2, T, device="cuda")
    rows, frames, level = (186, n // 2), (0, T), 0.5            # bandwidth extension above 4 kHz
    kw = dict(kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25)

    def two_calls():
        spec = _lib.stft_forward(wav, n, n, hop, **kw)
        return _lib.mask_fill(spec, noise, rows, frames, level)

    def fused():
        clean, corrupted = _lib.stft_forward(wav, n, n, hop, corrupt=dict(noise=noise, rows=rows, frames=frames, level=level), **kw)
        return corrupted, _lib.rect_mask(tuple(clean.shape), clean.device, rows, frames)
    a, b = two_calls(), fused()
    res = {"two_calls_ms": med(two_calls), "fused_ms": med(fused), "identical": bool(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])),
           "k1_alone_ms": med(lambda: _lib.stft_forward(wav, n, n, hop, **kw)),
           "what": "256 x 10 s clips, n_fft 2048, UpsampleMask-shaped rectangle (rows 186.., all frames), level 0.5; outputs: clean, corrupted, mask"}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
