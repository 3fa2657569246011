"""torch.library registration of the four kernels behind the plugin API, with the signatures SURVEY.md section 8b names:

    torch.ops.a2sb.stft_fwd(wav[B, L], n_fft, hop, power, eps)                      -> spec[B, 3, n_fft/2, T]
    torch.ops.a2sb.istft_inv(spec[B, 3, n_fft/2, T], n_fft, hop, power, eps, svd_fix) -> wav[B, hop * (T - 1)]
    torch.ops.a2sb.segment_gather(x[b, c, h, W], win, hop)                          -> segments[(b l), c, h, win]
    torch.ops.a2sb.segment_blend(segments, b, W, win, hop)                          -> [b, c, h, W]

Registered as custom ops (with fake / meta implementations for shape propagation) so that the path is visible to
torch's dispatcher -- FakeTensor tracing, torch.compile graphs that contain it, CUDA-graph capture helpers -- instead of
being an opaque Python call.  The implementations are the same C-ABI launches as the transform-module API
(audio_intelligence_b200._lib); they enqueue on the current torch stream and do not synchronise, so a warmed-up call
(plan tables uploaded) can be captured into a CUDA graph."""
from __future__ import annotations

import torch

from . import _capi, _lib


@torch.library.custom_op("a2sb::stft_fwd", mutates_args=())
def stft_fwd(wav: torch.Tensor, n_fft: int, hop: int, power: float, eps: float) -> torch.Tensor:
    return _lib.stft_forward(wav.contiguous(), n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=power, eps=eps)


@stft_fwd.register_fake
def _(wav, n_fft, hop, power, eps):
    return wav.new_empty((wav.shape[0], 3, n_fft // 2, 1 + wav.shape[1] // hop))


@torch.library.custom_op("a2sb::istft_inv", mutates_args=())
def istft_inv(spec: torch.Tensor, n_fft: int, hop: int, power: float, eps: float, svd_fix: bool) -> torch.Tensor:
    return _lib.istft_inverse(spec.contiguous(), n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=svd_fix,
                              power=power, eps=eps)


@istft_inv.register_fake
def _(spec, n_fft, hop, power, eps, svd_fix):
    return spec.new_empty((spec.shape[0], hop * (spec.shape[-1] - 1)))


@torch.library.custom_op("a2sb::segment_gather", mutates_args=())
def segment_gather(x: torch.Tensor, win: int, hop: int) -> torch.Tensor:
    return _lib.segment_gather(x.contiguous(), win, hop)


@segment_gather.register_fake
def _(x, win, hop):
    b, c, h, w = x.shape
    n = (w - (win - hop)) // hop if w >= win else 0
    return x.new_empty((b * n, c, h, win))


@torch.library.custom_op("a2sb::segment_blend", mutates_args=())
def segment_blend(segs: torch.Tensor, b: int, width: int, win: int, hop: int) -> torch.Tensor:
    return _lib.segment_blend(segs.contiguous(), b, width, win, hop)


@segment_blend.register_fake
def _(segs, b, width, win, hop):
    return segs.new_empty((b, segs.shape[1], segs.shape[2], width))
