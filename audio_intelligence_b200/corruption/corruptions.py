"""Drop-in mirror of A2SB/corruption/corruptions.py (masks + noise fill).

Same class names, constructor keywords, random draws (the same torch / numpy calls in the same
order, so a seeded run picks the same cut-offs and spans as the reference) and return values.
Every mask the reference builds is an axis-aligned rectangle; here it is produced, and the holes
filled with noise, by ONE kernel (`a2sb_mask_fill`, csrc/masks.cuh) instead of a zeros allocation,
a slice assignment and three elementwise passes.  The fill `x*(1-mask) + mask*randn*level` keeps the
reference's fp32 operation order, and the noise is drawn exactly where the reference draws it --
`torch.randn_like(spec)` on the INPUT tensor's device and generator -- so a seeded run returns the
reference's masks and the reference's filled values bit for bit (CPU inputs: CPU generator, then staged
to the GPU like every other CPU input of this package).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib


def mask_with_noise(x, mask, noise_level):
    """Reference: corruptions.py:14-15."""
    noise = torch.randn_like(x)
    xs, ms = _lib.stage(x), _lib.stage(mask.expand_as(x) if mask.shape != x.shape else mask)
    out = _lib.mask_with_noise(xs, ms, _lib.stage(noise), noise_level)
    return out if x.is_cuda else out.to(x.device)


def _rect(spec, rows_range, cols_range):
    out = _lib.rect_mask(tuple(spec.shape), _lib.stage(spec).device, rows_range, cols_range)
    return out if spec.is_cuda else out.to(spec.device)


class UpsampleMask:
    """Reference: corruptions.py:18-54 (bandwidth extension: rows [cutoff, h) are masked)."""

    def __init__(self, min_cutoff_freq: int, max_cutoff_freq: int, sampling_rate: int, dc_dropped: bool = True):
        self.min_cutoff_freq = min_cutoff_freq
        self.max_cutoff_freq = max_cutoff_freq
        self.sampling_rate = sampling_rate
        self.dc_dropped = dc_dropped

    @staticmethod
    def cutoff_row(h: int, min_cutoff_freq, max_cutoff_freq, sampling_rate, dc_dropped=True) -> int:
        """The integer arithmetic and the single torch.randint draw of corruptions.py:38-48."""
        n_fft = h * 2 if dc_dropped else (h - 1) * 2
        low = int(n_fft * min_cutoff_freq / float(sampling_rate))
        high = min(int(n_fft * max_cutoff_freq / float(sampling_rate)), h)
        high = max(high, low + 1)  # make sure high > low
        return int(torch.randint(low=low, high=high, size=[1])[0])

    @staticmethod
    def get_upsample_mask(spec: torch.Tensor, min_cutoff_freq: int, max_cutoff_freq: int, sampling_rate: int,
                          dc_dropped=True):
        c, h, l = spec.shape
        cutoff = UpsampleMask.cutoff_row(h, min_cutoff_freq, max_cutoff_freq, sampling_rate, dc_dropped)
        return _rect(spec, (cutoff, h), (0, l))

    def rect(self, spec):
        c, h, l = spec.shape
        return (self.cutoff_row(h, self.min_cutoff_freq, self.max_cutoff_freq, self.sampling_rate, self.dc_dropped), h), (0, l)

    def __call__(self, spec: torch.Tensor):
        return self.get_upsample_mask(spec, self.min_cutoff_freq, self.max_cutoff_freq, self.sampling_rate, self.dc_dropped)


class ExtensionMask:
    """Reference: corruptions.py:57-82."""

    def __init__(self, min_edge_distance=32):
        self.min_edge_distance = min_edge_distance

    @staticmethod
    def _span(l: int, min_edge_distance: int):
        start = int(torch.randint(low=min_edge_distance, high=l - min_edge_distance, size=[1])[0])
        if torch.randn(1) > 0:  # to the right
            return start, l
        return 0, start          # to the left

    @staticmethod
    def get_extension_mask(spec: torch.Tensor, min_edge_distance: int):
        c, h, l = spec.shape
        return _rect(spec, (0, h), ExtensionMask._span(l, min_edge_distance))

    def rect(self, spec):
        c, h, l = spec.shape
        return (0, h), self._span(l, self.min_edge_distance)

    def __call__(self, spec: torch.Tensor):
        return self.get_extension_mask(spec, self.min_edge_distance)


class InpaintMask:
    """Reference: corruptions.py:85-120."""

    def __init__(self, min_inpainting_frac: float, max_inpainting_frac: float, is_random: bool):
        assert 0.0 <= min_inpainting_frac <= max_inpainting_frac <= 1.0
        self.min_inpainting_frac = min_inpainting_frac
        self.max_inpainting_frac = max_inpainting_frac
        self.is_random = is_random

    @staticmethod
    def _span(w: int, min_inpainting_frac, max_inpainting_frac, is_random):
        random_variable_for_length = np.random.rand()
        inpainting_frac = random_variable_for_length * (max_inpainting_frac - min_inpainting_frac) + min_inpainting_frac
        if inpainting_frac == 0:
            return 0, 0
        if not is_random:
            inpainting_start_frac = 0.5 - inpainting_frac / 2.0
        else:
            inpainting_start_frac = np.random.rand() * (1.0 - inpainting_frac)
        return int(inpainting_start_frac * w), int((inpainting_start_frac + inpainting_frac) * w)

    @staticmethod
    def get_inpainting_mask(spec: torch.Tensor, min_inpainting_frac, max_inpainting_frac, is_random):
        c, h, w = spec.shape
        return _rect(spec, (0, h), InpaintMask._span(w, min_inpainting_frac, max_inpainting_frac, is_random))

    def rect(self, spec):
        c, h, w = spec.shape
        return (0, h), self._span(w, self.min_inpainting_frac, self.max_inpainting_frac, self.is_random)

    def __call__(self, spec: torch.Tensor):
        return self.get_inpainting_mask(spec, self.min_inpainting_frac, self.max_inpainting_frac, self.is_random)


def _fill(spec, rows_range, cols_range, level):
    noise = torch.randn_like(spec)
    out, mask = _lib.mask_fill(_lib.stage(spec), _lib.stage(noise), rows_range, cols_range, level)
    if not spec.is_cuda:
        out, mask = out.to(spec.device), mask.to(spec.device)
    return out, mask


class MultinomialInpaintMaskTransform:
    """Reference: corruptions.py:123-145."""

    def __init__(self, p_upsample_mask=0.5, p_extension_mask=0.5, p_inpaint_mask=0.0, fill_noise_level=0.5,
                 sampling_rate=22050, upsample_mask_kwargs={}, inpainting_mask_kwargs={}):
        self.mask_fns = [UpsampleMask(sampling_rate=sampling_rate, **upsample_mask_kwargs), ExtensionMask(),
                         InpaintMask(**inpainting_mask_kwargs)]
        self.mask_multinomial_probs = torch.Tensor([p_upsample_mask, p_extension_mask, p_inpaint_mask])
        self.fill_noise_level = fill_noise_level

    def __call__(self, spec: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        mask_fn = self.mask_fns[torch.multinomial(self.mask_multinomial_probs, 1)]
        rows_range, cols_range = mask_fn.rect(spec)
        return _fill(spec, rows_range, cols_range, self.fill_noise_level)


class TimestampedSegmentInpaintMaskTransform:
    """Reference: corruptions.py:147-160."""

    def __init__(self, start_time=0.5, end_time=1.0, hop_length=512, sampling_rate=44100, fill_noise_level=0.5):
        self.start_idx = int(sampling_rate / hop_length * start_time)
        self.end_idx = int(sampling_rate / hop_length * end_time)
        self.fill_noise_level = fill_noise_level

    def __call__(self, spec: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        return _fill(spec, (0, spec.shape[-2]), (self.start_idx, self.end_idx), self.fill_noise_level)
