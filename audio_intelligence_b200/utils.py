"""Drop-in mirror of the integer helper of A2SB/utils.py that sits next to the blend path
(SURVEY.md section 8a row B4), plus the window clamp its caller applies."""
from __future__ import annotations

import torch

from . import _lib


def find_middle_of_zero_segments(binary_array: torch.Tensor) -> torch.Tensor:
    """Reference: utils.py:54-81.  Middle indices ((start + end) / 2, truncated) of the runs of zeros of a
    1-D 0/1 tensor, computed by one single-CTA kernel (csrc/masks.cuh) instead of diff / nonzero / cat."""
    if not torch.is_tensor(binary_array) or binary_array.ndim != 1:
        raise ValueError("Input must be a 1D tensor.")
    centres, _ = _lib.zero_segment_windows(_lib.stage(binary_array), 1)
    return centres if binary_array.is_cuda else centres.to(binary_array.device)


def zero_segment_windows(binary_array: torch.Tensor, win_length: int) -> list[tuple[int, int]]:
    """The loop head of the fast-inpaint sampler (A2SB_lightning_module.py:160-174): for every zero run of
    `binary_array` (= 1 - mask row) the window [l, r) of `win_length` columns centred on it and shifted
    inside [0, len).  Raises AssertionError like the reference when a window cannot be placed."""
    if not torch.is_tensor(binary_array) or binary_array.ndim != 1:
        raise ValueError("Input must be a 1D tensor.")
    _, lr = _lib.zero_segment_windows(_lib.stage(binary_array), win_length)
    out = []
    for l_idx, r_idx in lr.tolist():
        assert r_idx - l_idx == win_length
        assert l_idx >= 0
        assert r_idx <= binary_array.shape[-1]
        out.append((l_idx, r_idx))
    return out
