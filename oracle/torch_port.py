"""Torch-on-CPU port of the reference's transform chains -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference runs this path as Python glue over torch.stft / torch.istft / torch.linalg.svd on
CPU tensors (A2SB/datasets/datasets.py:231-237, A2SB/A2SB_lightning_module.py:202-203).  The
reference sources cannot travel to the GPU box, so this module restates the same sequence of
library calls (not the reference's code) for two purposes only:
  * bench.py's `cpu_baseline` leg and `--impl reference` arm (kind "port"): the timing a user of
    the reference sees on the box's host cores, MKL threads included;
  * tests/: a second, independent checker next to the numpy oracle (a2sb_oracle.py).
It is pinned against the fixtures in tests/golden/ (tests/test_oracle.py::test_torch_port_*).
Nothing under audio_intelligence_b200/ imports it.
"""
from __future__ import annotations

import torch


def stft_complex(wav: torch.Tensor, n_fft: int, hop: int) -> torch.Tensor:
    """ComplexSpectrogram (A2SB/audio_transforms/transforms.py:83-105) -> complex [F, T].
    torchaudio Spectrogram(power=None) == torch.stft(center, reflect, hann, onesided, unnormalised)."""
    assert wav.dim() == 1, wav.shape
    w = torch.hann_window(n_fft, dtype=wav.dtype)
    return torch.stft(wav, n_fft, hop_length=hop, win_length=n_fft, window=w, center=True, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True)


def forward_chain(wav: torch.Tensor, n_fft: int = 2048, hop: int = 512, power: float = 0.25,
                  eps: float = 1e-9) -> torch.Tensor:
    """Shipped forward chain (A2SB/configs/ensemble_2split_sampling.yaml:105-119) -> [3, n_fft/2, T].
    Keeps the reference's op count: atan2 -> cos/sin, pow on all three channels, clone + index_put."""
    c = stft_complex(wav, n_fft, hop)
    re, im = c.real, c.imag
    mag = torch.sqrt(re * re + im * im)                    # transforms.py:116
    ph = torch.atan2(im, re)                               # transforms.py:117
    msp = torch.stack([mag, torch.cos(ph), torch.sin(ph)])[:, 1:, :]   # :118 + DropDC :219
    a = msp.abs()
    scale = a ** power / (a + eps)                         # transforms.py:199-200 (all channels)
    out = msp.clone()
    out[[0]] = msp[[0]] * scale[[0]]                       # transforms.py:205-206
    return out


def phase_fix_svd(msp: torch.Tensor) -> torch.Tensor:
    """SVDFixMagInstPhase (transforms.py:135-160) done the reference's way: one 2x2 SVD per bin."""
    c, s = msp[1], msp[2]
    R = torch.stack([torch.stack([c, -s], -1), torch.stack([s, c], -1)], -2)      # [..., 2, 2]
    U, _S, Vh = torch.linalg.svd(R)
    d = torch.linalg.det(U @ Vh)
    S1 = torch.stack([torch.ones_like(d), d], -1)
    Rn = U @ torch.diag_embed(S1) @ Vh
    return torch.stack([msp[0], Rn[..., 0, 0], Rn[..., 1, 0]])


def inverse_chain(spec: torch.Tensor, n_fft: int = 2048, hop: int = 512, power: float = 4.0, eps: float = 1e-9,
                  svd_fix: bool = True) -> torch.Tensor:
    """Shipped inverse chain (configs/ensemble_2split_sampling.yaml:63-78) -> wav [hop*(T-1)]."""
    assert spec.dim() == 3, spec.shape
    a = spec.abs()
    scale = a ** power / (a + eps)
    s = spec.clone()
    s[[0]] = spec[[0]] * scale[[0]]
    s = torch.cat((s[..., :1, :] * 0, s), -2)              # AddDC, transforms.py:227-228
    if svd_fix:
        s = phase_fix_svd(s)
    c = torch.complex(s[0] * s[1], s[0] * s[2])            # transforms.py:129-132,183
    w = torch.hann_window(n_fft, dtype=spec.dtype)
    return torch.istft(c, n_fft, hop_length=hop, win_length=n_fft, window=w, center=True, normalized=False,
                       onesided=True, length=None)


def roundtrip(wavs: torch.Tensor, n_fft: int = 2048, hop: int = 512, svd_fix: bool = False) -> torch.Tensor:
    """Per-clip loop, exactly how vocode_stft drives the chain (A2SB_lightning_module.py:97-98)."""
    return torch.stack([inverse_chain(forward_chain(w, n_fft, hop), n_fft, hop, svd_fix=svd_fix) for w in wavs])
