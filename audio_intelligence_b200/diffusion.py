"""Drop-in mirror of the segment windowing / blend half of A2SB/diffusion.py (lines 27-87).

`multidiffusion_pad_inputs`, `get_multidiffusion_vf` and `multidiffusion_unpad_outputs` keep the
reference's names, argument meaning and results (bit-identical for the same network outputs), but
the im2col `nn.Unfold` + einops copy becomes one gather kernel (K3) and the 2*num_hops Python-level
slice updates plus the final divide become one blend kernel (K4) -- see csrc/segments.cuh.
Callers: A2SB/A2SB_lightning_module.py:115-116,129-131,145,158-159,179.
The Schroedinger-bridge schedule (`Diffusion`, diffusion.py:90-168) is on the network side of the
boundary and is out of scope here.
"""
from __future__ import annotations

from math import ceil

import torch

from . import _lib


def multidiffusion_pad_inputs(input, win_length, hop_length, padding_constant=None):
    """Reference: diffusion.py:67-83.  Pads the frame axis by copying the HEAD of the signal (times 0
    plus `padding_constant` if given).  Like the reference, a pad longer than the input is
    truncated to the input's width (the reference slices `input[..., :to_pad]`)."""
    _b, _c, _h, width = input.shape
    if width <= win_length:  # no hops
        to_pad = win_length - width
    else:
        pad_to = ceil((width - win_length) / hop_length) * hop_length + win_length
        to_pad = pad_to - width
    x = _lib.stage(input)
    if to_pad > 0:
        out = _lib.wrap_pad(x, width + min(to_pad, width), padding_constant)
    else:
        out = x.clone()
    return out if input.is_cuda else out.to(input.device)


def multidiffusion_unpad_outputs(output, original_width: int):
    """Reference: diffusion.py:86-87."""
    return output[..., :original_width]


def get_multidiffusion_vf(vf_model, x_t, t_emb, win_length=256, hop_length=128, batch_size=16):
    """Reference: diffusion.py:27-64.

    t_emb should be b x emb_dim with all embeddings for the same time step.  Segments are ordered
    "(b l)" exactly like the reference's rearrange; the network is evaluated on torch.chunk-sized
    mini-batches (same chunk boundaries as the reference); overlapping outputs are summed in
    ascending segment order and divided by the overlap count."""
    b_size, _num_channels, _win_height, seq_len = x_t.shape
    num_hops = (seq_len - (win_length - hop_length)) // hop_length
    x = _lib.stage(x_t)
    segs = _lib.segment_gather(x, win_length, hop_length)          # [(b l), c, h, w]
    num_chunks = ceil(segs.shape[0] / batch_size)
    seg_chunks = torch.chunk(segs, num_chunks)
    t_emb_rpt = t_emb.repeat(num_hops, 1)
    t_emb_chunked = torch.chunk(t_emb_rpt, num_chunks)
    vfields = torch.empty_like(segs)
    row = 0
    for b_chunk_idx in range(num_chunks):
        chunk = seg_chunks[b_chunk_idx]
        out = vf_model(chunk if x_t.is_cuda else chunk.to(x_t.device), t_emb_chunked[b_chunk_idx])
        vfields[row:row + out.shape[0]].copy_(out)
        row += out.shape[0]
    out = _lib.segment_blend(vfields, b_size, seq_len, win_length, hop_length)
    return out if x_t.is_cuda else out.to(x_t.device)
