"""BASELINE config 4 sizes: K1/K2 time and achieved GB/s on algorithmic bytes for n_fft in {512, 1024, 2048, 4096}
(hop = n_fft/4, 256 x 10 s clips).  Usage: python tools/bench_nfft.py  (needs a GPU)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_intelligence_b200 import _capi, _lib  # noqa: E402


def med(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]


def main():
    B, L = 256, 441000
    wav = (0.3 * torch.randn(B, L, device="cuda")).clamp_(-1, 1)
    res = {}
    for n in (512, 1024, 2048, 4096):
        hop = n // 4
        T = 1 + L // hop
        spec = _lib.stft_forward(wav, n, n, hop, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25)
        k1 = med(lambda: _lib.stft_forward(wav, n, n, hop, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25))
        k2 = med(lambda: _lib.istft_inverse(spec, n, n, hop, kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True, power=4.0))
        fwd = B * (4 * L + 6 * n * T)
        inv = B * (6 * n * T + 4 * hop * (T - 1))
        res[n] = {"T": T, "row_bytes_mod_32": (T * 4) % 32, "k1_ms": k1, "k1_gbs": fwd / k1 * 1e-6, "k2_ms": k2,
                  "k2_gbs": inv / k2 * 1e-6}
        del spec
        # opt-in 32-byte row pitch (identical values)
        spec = _lib.stft_forward(wav, n, n, hop, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25, row_align=8)
        k1 = med(lambda: _lib.stft_forward(wav, n, n, hop, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25, row_align=8))
        k2 = med(lambda: _lib.istft_inverse(spec, n, n, hop, kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True, power=4.0))
        res[n].update({"pitched_k1_ms": k1, "pitched_k1_gbs": fwd / k1 * 1e-6, "pitched_k2_ms": k2, "pitched_k2_gbs": inv / k2 * 1e-6})
        del spec
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
