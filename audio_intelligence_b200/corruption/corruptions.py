"""Drop-in mirror of A2SB/corruption/corruptions.py (masks + noise fill).

Every mask the reference builds is an axis-aligned rectangle of ones inside a zero tensor.  Here a mask
class only decides the rectangle (`rect(spec) -> (row range, frame range)`) on the host -- drawing its random
numbers with the same torch / numpy calls in the same order as the reference, so a seeded run picks the same
cut-offs and spans -- and ONE kernel (`a2sb_mask_fill` / `a2sb_rect_mask`, csrc/masks.cuh) materialises the
mask and fills the hole with noise, instead of a zeros allocation, a slice assignment and three elementwise
passes.  The fill `x*(1-mask) + mask*randn*level` keeps the reference's fp32 operation order, and the noise is
drawn exactly where the reference draws it -- `torch.randn_like(spec)` on the INPUT tensor's device and
generator -- so a seeded run returns the reference's masks and the reference's filled values bit for bit
(CPU inputs: CPU generator, then staged to the GPU like every other CPU input of this package).

Reference lines: mask_with_noise :14-15, UpsampleMask :18-54, ExtensionMask :57-82, InpaintMask :85-120,
MultinomialInpaintMaskTransform :123-145, TimestampedSegmentInpaintMaskTransform :147-160.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib

Range = tuple  # (begin, end) with python slice semantics


def _materialise(spec: torch.Tensor, rows: Range, frames: Range) -> torch.Tensor:
    """zeros_like(spec) with ones on rows x frames, on spec's device."""
    m = _lib.rect_mask(tuple(spec.shape), _lib.stage(spec).device, rows, frames)
    return m if spec.is_cuda else m.to(spec.device)


def _fill(spec: torch.Tensor, rows: Range, frames: Range, level: float):
    """(spec with the rectangle replaced by noise * level, mask): one kernel."""
    noise = torch.randn_like(spec)                       # the reference's draw (corruptions.py:15), same device
    tag = getattr(spec, "_a2sb_padded", None)
    if tag is not None and spec.is_cuda and spec.dtype == torch.float32 and spec.stride(-1) == 1:
        # K1 produced `spec` as the view of a buffer padded for the segment windowing (transforms.set_segment_padding):
        # keep that layout -- filled tensor and mask come out padded as well, so the sampler does not pad them again
        return _lib.mask_fill_padded(spec, noise.contiguous(), rows, frames, level, tag[1], tag[2])
    filled, m = _lib.mask_fill(_lib.stage(spec), _lib.stage(noise), rows, frames, level)
    return (filled, m) if spec.is_cuda else (filled.to(spec.device), m.to(spec.device))


def mask_with_noise(x, mask, noise_level):
    """x * (1 - mask) + mask * randn_like(x) * noise_level for an arbitrary mask tensor."""
    noise = torch.randn_like(x)
    m = mask if mask.shape == x.shape else mask.expand_as(x)
    out = _lib.mask_with_noise(_lib.stage(x), _lib.stage(m), _lib.stage(noise), noise_level)
    return out if x.is_cuda else out.to(x.device)


class _RectangleMask:
    """A mask transform whose mask is one rectangle; subclasses implement `rect`."""

    def rect(self, spec: torch.Tensor) -> tuple[Range, Range]:
        raise NotImplementedError

    def __call__(self, spec: torch.Tensor) -> torch.Tensor:
        rows, frames = self.rect(spec)
        return _materialise(spec, rows, frames)


class UpsampleMask(_RectangleMask):
    """Bandwidth extension: every row from a randomly drawn cut-off bin upwards is masked."""

    def __init__(self, min_cutoff_freq: int, max_cutoff_freq: int, sampling_rate: int, dc_dropped: bool = True):
        self.min_cutoff_freq, self.max_cutoff_freq = min_cutoff_freq, max_cutoff_freq
        self.sampling_rate, self.dc_dropped = sampling_rate, dc_dropped

    @staticmethod
    def cutoff_row(n_rows: int, min_cutoff_freq, max_cutoff_freq, sampling_rate, dc_dropped=True) -> int:
        """Bin of the cut-off: n_fft = 2*rows (DC dropped) or 2*(rows-1); one torch.randint draw in [lo, hi)."""
        bins = 2 * n_rows if dc_dropped else 2 * (n_rows - 1)
        sr = float(sampling_rate)
        lo = int(bins * min_cutoff_freq / sr)
        hi = max(min(int(bins * max_cutoff_freq / sr), n_rows), lo + 1)
        return int(torch.randint(low=lo, high=hi, size=[1])[0])

    @staticmethod
    def get_upsample_mask(spec: torch.Tensor, min_cutoff_freq: int, max_cutoff_freq: int, sampling_rate: int,
                          dc_dropped=True):
        _c, n_rows, n_frames = spec.shape
        first = UpsampleMask.cutoff_row(n_rows, min_cutoff_freq, max_cutoff_freq, sampling_rate, dc_dropped)
        return _materialise(spec, (first, n_rows), (0, n_frames))

    def rect(self, spec):
        _c, n_rows, n_frames = spec.shape
        first = self.cutoff_row(n_rows, self.min_cutoff_freq, self.max_cutoff_freq, self.sampling_rate, self.dc_dropped)
        return (first, n_rows), (0, n_frames)


class ExtensionMask(_RectangleMask):
    """Extension: everything to the right (or, on a coin flip, to the left) of a random frame is masked."""

    def __init__(self, min_edge_distance=32):
        self.min_edge_distance = min_edge_distance

    @staticmethod
    def _frames(n_frames: int, margin: int) -> Range:
        pivot = int(torch.randint(low=margin, high=n_frames - margin, size=[1])[0])
        return (pivot, n_frames) if torch.randn(1) > 0 else (0, pivot)

    @staticmethod
    def get_extension_mask(spec: torch.Tensor, min_edge_distance: int):
        _c, n_rows, n_frames = spec.shape
        return _materialise(spec, (0, n_rows), ExtensionMask._frames(n_frames, min_edge_distance))

    def rect(self, spec):
        _c, n_rows, n_frames = spec.shape
        return (0, n_rows), self._frames(n_frames, self.min_edge_distance)


class InpaintMask(_RectangleMask):
    """Inpainting: a span covering a random fraction of the frames, centred or randomly placed."""

    def __init__(self, min_inpainting_frac: float, max_inpainting_frac: float, is_random: bool):
        assert 0.0 <= min_inpainting_frac <= max_inpainting_frac <= 1.0
        self.min_inpainting_frac, self.max_inpainting_frac = min_inpainting_frac, max_inpainting_frac
        self.is_random = is_random

    @staticmethod
    def _frames(n_frames: int, lo_frac, hi_frac, is_random) -> Range:
        frac = np.random.rand() * (hi_frac - lo_frac) + lo_frac            # first numpy draw: the length
        if frac == 0:
            return (0, 0)
        start = np.random.rand() * (1.0 - frac) if is_random else 0.5 - frac / 2.0   # second draw only if random
        return int(start * n_frames), int((start + frac) * n_frames)

    @staticmethod
    def get_inpainting_mask(spec: torch.Tensor, min_inpainting_frac, max_inpainting_frac, is_random):
        _c, n_rows, n_frames = spec.shape
        return _materialise(spec, (0, n_rows), InpaintMask._frames(n_frames, min_inpainting_frac, max_inpainting_frac, is_random))

    def rect(self, spec):
        _c, n_rows, n_frames = spec.shape
        return (0, n_rows), self._frames(n_frames, self.min_inpainting_frac, self.max_inpainting_frac, self.is_random)


class MultinomialInpaintMaskTransform:
    """Pick one of the three mask kinds by torch.multinomial, then mask and fill: (filled spec, mask)."""

    def __init__(self, p_upsample_mask=0.5, p_extension_mask=0.5, p_inpaint_mask=0.0, fill_noise_level=0.5,
                 sampling_rate=22050, upsample_mask_kwargs={}, inpainting_mask_kwargs={}):
        self.mask_fns = [UpsampleMask(sampling_rate=sampling_rate, **upsample_mask_kwargs), ExtensionMask(),
                         InpaintMask(**inpainting_mask_kwargs)]
        self.mask_multinomial_probs = torch.Tensor([p_upsample_mask, p_extension_mask, p_inpaint_mask])
        self.fill_noise_level = fill_noise_level

    def __call__(self, spec: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        kind = self.mask_fns[torch.multinomial(self.mask_multinomial_probs, 1)]
        rows, frames = kind.rect(spec)
        return _fill(spec, rows, frames, self.fill_noise_level)


class TimestampedSegmentInpaintMaskTransform:
    """Mask the frames between two timestamps (seconds) and fill them with noise: (filled spec, mask)."""

    def __init__(self, start_time=0.5, end_time=1.0, hop_length=512, sampling_rate=44100, fill_noise_level=0.5):
        frames_per_second = sampling_rate / hop_length
        self.start_idx = int(frames_per_second * start_time)
        self.end_idx = int(frames_per_second * end_time)
        self.fill_noise_level = fill_noise_level

    def __call__(self, spec: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        return _fill(spec, (0, spec.shape[-2]), (self.start_idx, self.end_idx), self.fill_noise_level)
