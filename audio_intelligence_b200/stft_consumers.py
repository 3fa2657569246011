"""Other consumers of the forward / inverse kernels in the reference repository:

  * ETTA's `STFT` helper module (ETTA/stable_audio_tools/models/adp.py:1510-1590): torch.stft / torch.istft with
    `normalized=True`, returning (magnitude, phase-angle) or (real, imag) per channel;
  * the STFT of auraloss' (multi-resolution) STFT loss (ETTA/stable_audio_tools/training/losses/auraloss.py:363-381):
    hops that do not divide n_fft (50 / 120 / 240), win_length < n_fft, clamped magnitude.

Both run on K1 (and K2 for decode) with two small pointwise kernels for (|X|, angle X) and polar -> complex.
`normalized=True` needs no kernel change: a plan whose window is w / sqrt(n_fft) scales the forward transform by
n_fft^-1/2 and -- because the inverse divides by the overlap-added SQUARED window -- the inverse by n_fft^+1/2, exactly
torch's convention.

Transform lengths outside {512, 1024, 2048, 4096} -- ETTA's default `num_fft=1023`, an odd length chosen to get exactly 512
bins -- go through the any-length kernels of csrc/dft_generic.cuh (a plain O(n_fft^2) DFT per frame from a table of the
roots of unity, torch.stft / torch.istft semantics incl. `length` beyond hop * (frames - 1)): correct and tested against
torch's results, not tuned -- the module is not instantiated by the shipped configs.  The radix kernels are forward-only for
hops that are not a multiple of 4 dividing n_fft.  Inference / evaluation only: no autograd (the reference's loss
differentiates through torch.stft)."""
from __future__ import annotations

from math import floor
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _capi, _lib

OP_COMPLEX_TO_MAG_ANGLE, OP_POLAR_TO_COMPLEX = 4, 5
SUPPORTED_N_FFT = (512, 1024, 2048, 4096)


def _check_n_fft(n_fft: int) -> None:
    if n_fft not in SUPPORTED_N_FFT:
        raise NotImplementedError(f"n_fft={n_fft}: the radix kernels transform power-of-two lengths {SUPPORTED_N_FFT} only "
                                  "(ETTA's STFT helper takes any length through the generic kernels)")


def _nola_check(w, hop: int, n_frames: int, length: int) -> None:
    """torch.istft's check (ATen SpectralOps: "window overlap add min: 1"): the overlap-added squared window must stay above
    1e-11 on the samples that are returned and that a frame reaches."""
    import numpy as np
    n_fft = w.shape[0]
    total = n_fft + hop * (n_frames - 1)
    env = np.zeros(total)
    w2 = w * w
    for t in range(n_frames):
        env[t * hop:t * hop + n_fft] += w2
    a, b = n_fft // 2, min(n_fft // 2 + length, total)
    if b > a and float(np.abs(env[a:b]).min()) < 1e-11:
        raise RuntimeError("window overlap add min: 1")


def _generic_window(window: Tensor, n_fft: int, normalized: bool, device) -> Tensor:
    """The window centre-padded to n_fft like torch.stft (left = (n_fft - win_length) // 2); `normalized=True` rides on it:
    w / sqrt(n_fft) scales the forward transform by n_fft^-1/2 and the inverse (which divides by the overlap-added SQUARED
    window) by n_fft^+1/2 -- torch's convention in both directions."""
    w = torch.zeros(n_fft, dtype=torch.float32)
    left = (n_fft - window.numel()) // 2
    w[left:left + window.numel()] = window.detach().float().cpu()
    if normalized:
        w = w / float(n_fft) ** 0.5
    return w.to(device)


def _pointwise2(op: int, x: Tensor, eps: float = 0.0) -> Tensor:
    """[B, 2, F, T] -> [B, 2, F, T] through a per-element op on the two planes."""
    out = torch.empty_like(x)
    L = _lib.lib()
    n = x[0, 0].numel()
    for b in range(x.shape[0]):
        _capi.check(L, L.a2sb_pointwise(op, x[b].data_ptr(), out[b].data_ptr(), n, 2, 0, 1.0, float(eps), _lib.stream_ptr()))
    return out


def closest_power_2(x: float) -> int:
    """adp.py: closest power of two to x (ties towards the lower one, like the reference's argmin over candidates)."""
    import math
    e = math.log2(x)
    lo, hi = 2 ** math.floor(e), 2 ** math.ceil(e)
    return int(lo if abs(x - lo) <= abs(x - hi) else hi)


class STFT(torch.nn.Module):
    """Reference: ETTA/stable_audio_tools/models/adp.py:1510-1590 (same constructor keywords, encode / decode /
    encode1d / decode1d).  wave [b, c, t] -> (a, b) each [b, c, num_fft/2 + 1, frames]."""

    def __init__(self, num_fft: int = 1023, hop_length: int = 256, window_length: Optional[int] = None,
                 length: Optional[int] = None, use_complex: bool = False):
        super().__init__()
        self.generic = num_fft not in SUPPORTED_N_FFT          # any other length: csrc/dft_generic.cuh
        self.num_fft = num_fft
        self.hop_length = hop_length if hop_length is not None else floor(num_fft // 4)
        self.window_length = window_length if window_length is not None else num_fft
        self.length = length
        self.register_buffer("window", torch.hann_window(self.window_length))
        self.use_complex = use_complex

    def encode(self, wave: Tensor) -> Tuple[Tensor, Tensor]:
        b, c, t = wave.shape
        w = _lib.stage(wave).reshape(b * c, t)
        if self.generic:
            x = _lib.dft_generic_forward(w, self.num_fft, self.hop_length, _generic_window(self.window, self.num_fft, True, w.device))
        else:
            x = _lib.stft_forward(w, self.num_fft, self.window_length, self.hop_length, kind=_capi.KIND_COMPLEX, normalized=True)
        if not self.use_complex:
            x = _pointwise2(OP_COMPLEX_TO_MAG_ANGLE, x)                   # torch.abs, torch.angle (:1550)
        a, bb = x[:, 0].reshape(b, c, *x.shape[2:]), x[:, 1].reshape(b, c, *x.shape[2:])
        dev = wave.device
        return _lib.to_host(a, dev), _lib.to_host(bb, dev)

    def decode(self, stft_a: Tensor, stft_b: Tensor) -> Tensor:
        b, c, f, l = stft_a.shape
        length = self.length if self.length is not None else closest_power_2(l * self.hop_length)     # :1556, :1580
        x = torch.stack([_lib.stage(stft_a).reshape(b * c, f, l), _lib.stage(stft_b).reshape(b * c, f, l)], dim=1)
        if not self.use_complex:
            x = _pointwise2(OP_POLAR_TO_COMPLEX, x)                       # magnitude * cos / sin (:1566-1567)
        if self.generic:
            win = _generic_window(self.window, self.num_fft, True, x.device)
            _nola_check(_generic_window(self.window, self.num_fft, False, "cpu").double().numpy(), self.hop_length, l, int(length))
            y = _lib.dft_generic_inverse(x.contiguous(), self.num_fft, self.hop_length, win, int(length))
            return _lib.to_host(y.reshape(b, c, length), stft_a.device)
        y = _lib.istft_inverse(x, self.num_fft, self.window_length, self.hop_length, kind=_capi.KIND_COMPLEX, has_dc=True,
                               normalized=True)
        if length > y.shape[-1]:
            raise NotImplementedError(f"length={length} beyond hop_length * (frames - 1) = {y.shape[-1]}: the tail torch.istft "
                                      "reconstructs from the last frames' second halves is not produced by the fused kernel")
        return _lib.to_host(y[:, :length].reshape(b, c, length), stft_a.device)

    def encode1d(self, wave: Tensor, stacked: bool = True):
        a, bb = self.encode(wave)
        a, bb = a.reshape(a.shape[0], -1, a.shape[-1]), bb.reshape(bb.shape[0], -1, bb.shape[-1])     # b (c f) l
        return torch.cat((a, bb), dim=1) if stacked else (a, bb)

    def decode1d(self, stft_pair: Tensor) -> Tensor:
        f = self.num_fft // 2 + 1
        a, bb = stft_pair.chunk(chunks=2, dim=1)
        a, bb = a.reshape(a.shape[0], -1, f, a.shape[-1]), bb.reshape(bb.shape[0], -1, f, bb.shape[-1])
        return self.decode(a, bb)


def stft_magnitude(x: Tensor, fft_size: int, hop_size: int, win_length: int, window: Optional[Tensor] = None,
                   eps: float = 1e-8, want_phase: bool = False) -> Tuple[Tensor, Optional[Tensor]]:
    """auraloss STFTLoss.stft (auraloss.py:363-381): x [B, T] -> (sqrt(clamp(re^2 + im^2, min=eps)), angle or None), each
    [B, fft_size/2 + 1, frames].  `window` defaults to torch.hann_window(win_length) (auraloss' default "hann_window")."""
    _check_n_fft(fft_size)
    c = _lib.stft_forward(_lib.stage(x), fft_size, win_length, hop_size, kind=_capi.KIND_COMPLEX, window=window)
    mp = _pointwise2(OP_COMPLEX_TO_MAG_ANGLE, c, eps=eps)
    return _lib.to_host(mp[:, 0], x.device), (_lib.to_host(mp[:, 1], x.device) if want_phase else None)
