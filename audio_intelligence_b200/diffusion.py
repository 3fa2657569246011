"""Drop-in mirror of the segment windowing / blend half of A2SB/diffusion.py (lines 27-87).

`multidiffusion_pad_inputs`, `get_multidiffusion_vf` and `multidiffusion_unpad_outputs` keep the
reference's names, argument meaning and results (bit-identical for the same network outputs), but
the im2col `nn.Unfold` + einops copy becomes one gather kernel (K3) and the 2*num_hops Python-level
slice updates plus the final divide become one blend kernel (K4) -- see csrc/segments.cuh.
Callers: A2SB/A2SB_lightning_module.py:115-116,129-131,145,158-159,179.
`Diffusion` restates the closed-form Schroedinger-bridge schedule scalars (diffusion.py:90-168) and
`ddpm_sample` the reverse sampling loop of A2SB_lightning_module.py:103-146:
per step, the segment gather, the network on torch.chunk mini-batches, and then ONE kernel (K4s,
csrc/segments.cuh) for blend + get_pred_x0 + mask merge + p_posterior + re-imposition of the known
region, instead of ~10 elementwise passes.  Results are bit-identical to the reference's torch ops
for the same network outputs and noise tensors.
"""
from __future__ import annotations

import ctypes as C
from math import ceil

import torch

from . import _capi, _lib


def multidiffusion_pad_inputs(input, win_length, hop_length, padding_constant=None):
    """Reference: diffusion.py:67-83.  Pads the frame axis by copying the HEAD of the signal (times 0
    plus `padding_constant` if given).  Like the reference, a pad longer than the input is
    truncated to the input's width (the reference slices `input[..., :to_pad]`)."""
    _b, _c, _h, width = input.shape
    ready = _lib.padded_buffer_of(input, win_length, hop_length, padding_constant) if input.is_cuda else None
    if ready is not None:
        # a kernel of this package (K1 / the mask fill, transforms.set_segment_padding) already wrote the padding behind
        # the view: no copy.  (The reference returns a fresh tensor; its callers only read it, A2SB_lightning_module.py:115-146.)
        return ready
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


def compute_gaussian_product_coef(sigma1, sigma2):
    """Reference: diffusion.py:91-99.  p1 = N(x_t | x_0, sigma1^2), p2 = N(x_t | x_1, sigma2^2) ->
    p1 * p2 = N(x_t | coef1 x_0 + coef2 x_1, var)."""
    s1, s2 = sigma1 ** 2, sigma2 ** 2
    denom = s1 + s2
    return s2 / denom, s1 / denom, (s1 * s2) / denom


class Diffusion(torch.nn.Module):
    """Reference: diffusion.py:101-168 -- the symmetric quadratic beta schedule of the bridge
    (t = 0 clean data, t = 1 corrupted posterior) and the scalars derived from it.  Only host-side scalar
    math lives here; the tensor arithmetic that uses these scalars is fused into K4s."""

    def __init__(self, beta_min=1e-4, beta_max=0.3):
        super().__init__()
        self.beta_min = beta_min
        self.beta_max = beta_max

    def get_beta_t(self, t):
        return (t ** 2 if t <= 0.5 else (1 - t) ** 2) * self.beta_max

    def get_int_beta_0_t(self, t):
        """Integral of beta over [0, t] for a tensor of times in [0, 1]."""
        third = 1 / 3 * self.beta_max
        whole = 2 * self.beta_max * (0.5 ** 3) / 3
        out = t.clone()
        upper = t > 0.5
        out[upper] = whole - third * ((1 - t[upper]) ** 3)
        out[~upper] = third * (t[~upper] ** 3)
        return out

    def get_std_fwd(self, t):
        return torch.sqrt(self.get_int_beta_0_t(t))

    def get_std_rev(self, t):
        return torch.sqrt(self.get_int_beta_0_t(1 - t))

    def get_std_t(self, t):
        _c1, _c2, var = compute_gaussian_product_coef(self.get_std_fwd(t), self.get_std_rev(t))
        return torch.sqrt(var)

    def posterior_coefs(self, t_prev, t):
        """(mu_x0, mu_xt, var) of p(x_{t_prev} | x_t, x_0)  (p_posterior, diffusion.py:153-158)."""
        assert t_prev < t
        std_t = self.get_std_fwd(t)
        std_t_prev = self.get_std_fwd(t_prev)
        std_delta = (std_t ** 2 - std_t_prev ** 2).sqrt()
        return compute_gaussian_product_coef(std_t_prev, std_delta)

    def q_sample(self, t, x_0, x_1, ot_ode=False):
        """Reference: diffusion.py:137-151 (training-side sample of q(x_t | x_0, x_1); plain torch)."""
        coef1, coef2, var = compute_gaussian_product_coef(self.get_std_fwd(t), self.get_std_rev(t))
        while len(coef1.shape) < len(x_0.shape):
            coef1, coef2, var = coef1[:, None], coef2[:, None], var[:, None]
        x_t = coef1 * x_0 + coef2 * x_1
        if not ot_ode:
            x_t += torch.sqrt(var) * torch.randn_like(x_t)
        return x_t.detach()

    def p_posterior(self, t_prev, t, x_t, x_0, ot_ode=False):
        """Reference: diffusion.py:153-163 (stand-alone form; `ddpm_sample` below uses the fused kernel)."""
        mu_x0, mu_xt, var = self.posterior_coefs(t_prev, t)
        x_t_prev = mu_x0 * x_0 + mu_xt * x_t
        if not ot_ode and t_prev > 0:
            x_t_prev = x_t_prev + var.sqrt() * torch.randn_like(x_t_prev)
        return x_t_prev

    def get_pred_x0(self, t, x_t, net_out):
        """Reference: diffusion.py:165-168."""
        return x_t - self.get_std_fwd(t) * net_out


def _scalar(v) -> float:
    return float(v.reshape(-1)[0].item()) if torch.is_tensor(v) else float(v)


@torch.no_grad()
def ddpm_sample(vf_model, ddpm: Diffusion, x_1, t_steps, t_to_emb, mask=None, mask_pred_x0=True, win_length=256,
                hop_length=256, batch_size=16, use_ot_ode=True, get_vf_model=None, outputs_to_cpu=False, history="all"):
    """Reference: A2SB_lightning_module.py:103-146 (`A2SBModel.ddpm_sample`) as a free function: `vf_model`
    (or `get_vf_model(t) -> model`, :83-86), `ddpm`, `t_to_emb` and `use_ot_ode` are the attributes the method
    reads from `self`.  x_1: [b, c, h, w]; t_steps: [1, n_steps + 1] descending times; returns the list of
    per-step pred_x0 (un-padded).

    The reference copies every step's pred_x0 to the host synchronously (`pred_x0.cpu()`, :133 -- 3.8 GB per step
    at the 1 h configuration) although its callers only read the last one (:179, :202).  Here:
      outputs_to_cpu=False (default)  the history stays on the device, no copy, no synchronisation;
      outputs_to_cpu="async"          each step's pred_x0 is copied into pinned host memory on a side stream, under the
                                      next step's kernels; one synchronisation at the end; returns CPU tensors;
      outputs_to_cpu=True             the reference's blocking per-step copy.
      history="last"                  keep (and return a 1-element list with) only the final pred_x0."""
    assert hop_length <= win_length
    assert history in ("all", "last")
    if not x_1.is_cuda and not outputs_to_cpu:
        outputs_to_cpu = True          # module convention: results live on the input's device
    n_steps = t_steps.shape[1] - 1
    original_width = x_1.shape[-1]
    dev_in = x_1.device
    # (pad_inputs stages CPU inputs itself, and hands out an already padded buffer without copying -- do not stage first)
    x_1 = _lib.stage(multidiffusion_pad_inputs(x_1, win_length, hop_length))
    if mask is not None:
        mask = _lib.stage(multidiffusion_pad_inputs(mask, win_length, hop_length))
        if mask.shape != x_1.shape:
            mask = mask.expand_as(x_1).contiguous()
    x_t = x_1.clone()
    b_size, _c, _h, seq_len = x_1.shape
    num_hops = (seq_len - (win_length - hop_length)) // hop_length
    all_pred_x0s = []
    L = _lib.lib()
    copy_stream = torch.cuda.Stream(device=x_1.device) if outputs_to_cpu == "async" else None
    for t_idx in range(n_steps):
        t = t_steps[:, t_idx]
        t_prev = t_steps[:, t_idx + 1]
        t_emb = t_to_emb(t).repeat(x_1.shape[0], 1)
        model = get_vf_model(t[0].item()) if get_vf_model is not None else vf_model
        # --- get_multidiffusion_vf up to the network outputs (diffusion.py:33-50)
        segs = _lib.segment_gather(x_t, win_length, hop_length)
        num_chunks = ceil(segs.shape[0] / batch_size)
        seg_chunks = torch.chunk(segs, num_chunks)
        t_emb_chunked = torch.chunk(t_emb.repeat(num_hops, 1), num_chunks)
        vfields = torch.empty_like(segs)
        row = 0
        for k in range(num_chunks):
            out = model(seg_chunks[k], t_emb_chunked[k])
            vfields[row:row + out.shape[0]].copy_(out)
            row += out.shape[0]
        # --- schedule scalars (host) and noise (drawn in the reference's order)
        mu_x0, mu_xt, var = ddpm.posterior_coefs(t_prev, t)
        noise_post = torch.randn_like(x_t) if (not use_ot_ode and bool(t_prev > 0)) else None
        noise_mask = torch.randn_like(x_1) if (mask is not None and not use_ot_ode) else None
        pred_x0 = torch.empty_like(x_t)
        x_next = torch.empty_like(x_t)
        a = _capi.StepArgs(x_t.data_ptr(), x_1.data_ptr(), mask.data_ptr() if mask is not None else None,
                           noise_post.data_ptr() if noise_post is not None else None,
                           noise_mask.data_ptr() if noise_mask is not None else None, pred_x0.data_ptr(),
                           x_next.data_ptr(), _scalar(ddpm.get_std_fwd(t)), _scalar(mu_x0), _scalar(mu_xt),
                           _scalar(var.sqrt()), _scalar(ddpm.get_std_t(t_prev)) if noise_mask is not None else 0.0,
                           int(bool(mask_pred_x0)))
        _capi.check(L, L.a2sb_segment_blend_step(vfields.data_ptr(), C.byref(a), b_size, x_t.shape[1] * x_t.shape[2],
                                                 seq_len, win_length, hop_length, _lib.stream_ptr()))
        if history == "all" or t_idx == n_steps - 1:
            if outputs_to_cpu == "async":
                host = torch.empty(pred_x0.shape, dtype=pred_x0.dtype, pin_memory=True)
                copy_stream.wait_stream(torch.cuda.current_stream(x_1.device))
                with torch.cuda.stream(copy_stream):
                    host.copy_(pred_x0, non_blocking=True)
                pred_x0.record_stream(copy_stream)
                all_pred_x0s.append(host)
            else:
                all_pred_x0s.append(pred_x0.cpu() if outputs_to_cpu else pred_x0)
        x_t = x_next
    if copy_stream is not None:
        copy_stream.synchronize()
    del dev_in
    return [multidiffusion_unpad_outputs(pred, original_width) for pred in all_pred_x0s]


@torch.no_grad()
def fast_inpaint_ddpm_sample(vf_model, ddpm: Diffusion, x_1, t_steps, t_to_emb, mask=None, mask_pred_x0=True,
                             win_length=256, hop_length=256, batch_size=16, use_ot_ode=True, get_vf_model=None):
    """Reference: A2SB_lightning_module.py:149-180 (`A2SBModel.fast_inpaint_ddpm_sample`): assumes every masked
    stretch is shorter than `win_length` and sufficiently separated; samples ONE window per hole and pastes the
    final pred_x0 back.  Window placement (centre of each zero run of 1 - mask[0, 0, 0], shifted inside the padded
    width) comes from the device kernel behind `utils.zero_segment_windows`; everything stays on the device and
    only the last pred_x0 of each window's sampling run is kept.  Returns `[x_1]` like the reference."""
    from .utils import zero_segment_windows
    original_width = x_1.shape[-1]
    dev_in = x_1.device
    x_1 = multidiffusion_pad_inputs(_lib.stage(x_1), win_length, hop_length)            # always a copy (:156-157)
    mask = multidiffusion_pad_inputs(_lib.stage(mask), win_length, hop_length, padding_constant=0)
    for l_idx, r_idx in zero_segment_windows(1 - mask[0, 0, 0], win_length):
        curr_x_1 = x_1[:, :, :, l_idx:r_idx].contiguous()
        curr_mask = mask[:, :, :, l_idx:r_idx].contiguous()
        new_x_0 = ddpm_sample(vf_model, ddpm, curr_x_1, t_steps, t_to_emb, mask=curr_mask, mask_pred_x0=mask_pred_x0,
                              win_length=win_length, hop_length=hop_length, batch_size=batch_size, use_ot_ode=use_ot_ode,
                              get_vf_model=get_vf_model, outputs_to_cpu=False, history="last")
        x_1[:, :, :, l_idx:r_idx] = new_x_0[-1]
    x_1 = multidiffusion_unpad_outputs(x_1, original_width)
    return [x_1 if dev_in.type == "cuda" else x_1.to(dev_in)]
