"""Multi-GPU sharding of the spectral-transform path: integer planners + halo exchange.

The reference runs this path on one GPU, batch 1 (A2SB/A2SB_lightning_module.py:185); there is no
reference counterpart.  Batches shard by clip with no communication.  ONE long clip (BASELINE
config 3) shards by contiguous frame / segment ranges, one process per GPU:

  forward  rank g owns frames [t_g, t_{g+1}) and the samples [t_g*hop, t_{g+1}*hop); it needs a halo of
           n_fft/2 samples on the left and n_fft/2 - hop on the right from its neighbours
  inverse  rank g owns the same frames and the output samples [t_g*hop, t_{g+1}*hop) (clipped to
           hop*(T-1)); it needs 1 halo frame on the left and 2 on the right (n_fft = 4*hop)
  blend    rank g owns segments [k_g, k_{g+1}); it needs the first win-hop columns of its right neighbour
           to cut its last segments and the last ceil(win/hop)-1 network outputs of its left neighbour
           to blend its first columns (summed in ascending segment order -> bit-identical)

Halos travel by neighbour send/recv (`torch.distributed.batch_isend_irecv`, NCCL over NVLink on the
GPU box, gloo in the CPU tests); results are collected with `all_gather`.  The compute calls go to the
CUDA library through `audio_intelligence_b200._lib`; tests inject a CPU backend (the oracle) to check
the planner and the exchange logic without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from math import ceil
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

# ------------------------------------------------------------------------------------------------
# integer planners (pure, bit-exact)
# ------------------------------------------------------------------------------------------------


def split_range(n: int, world: int, align: int = 1) -> List[Tuple[int, int]]:
    """Contiguous, balanced, `align`-aligned cut of [0, n) into `world` ranges (trailing ranges may be empty)."""
    units = ceil(n / align)
    cuts = [min(n, ((units * g) // world) * align) for g in range(world)] + [n]
    return [(cuts[g], cuts[g + 1]) for g in range(world)]


@dataclass(frozen=True)
class ForwardShard:
    t0: int          # first owned frame
    t1: int          # one past the last owned frame
    own0: int        # owned samples [own0, own1)
    own1: int
    need0: int       # samples the owned frames touch after reflection, [need0, need1)
    need1: int


def forward_shard(length: int, n_fft: int, hop: int, world: int, rank: int, frame_align: int = 16) -> ForwardShard:
    """Frames [t0, t1) of T = 1 + L // hop and the sample window they need (reflect padding folded in)."""
    T = 1 + length // hop
    t0, t1 = split_range(T, world, frame_align)[rank]
    own0, own1 = min(t0 * hop, length), (length if t1 >= T else min(t1 * hop, length))
    if t1 <= t0:
        return ForwardShard(t0, t1, own0, own0, own0, own0)
    lo, hi = t0 * hop - n_fft // 2, (t1 - 1) * hop + n_fft // 2 - 1
    need0, need1 = max(lo, 0), min(hi, length - 1)
    if lo < 0:
        need1 = max(need1, -lo)                       # reflected head: x[-i] = x[i]
    if hi >= length:
        need0 = min(need0, 2 * (length - 1) - hi)     # reflected tail
    return ForwardShard(t0, t1, own0, own1, need0, need1 + 1)


@dataclass(frozen=True)
class InverseShard:
    t0: int          # owned frames [t0, t1)
    t1: int
    out0: int        # owned trimmed output samples [out0, out0 + out_n)
    out_n: int
    f0: int          # frames needed [f0, f1)
    f1: int


def inverse_shard(n_frames: int, n_fft: int, hop: int, world: int, rank: int, frame_align: int = 16) -> InverseShard:
    T = n_frames
    total = hop * (T - 1)
    t0, t1 = split_range(T, world, frame_align)[rank]
    out0, out1 = min(t0 * hop, total), (total if t1 >= T else min(t1 * hop, total))
    if out1 <= out0:
        return InverseShard(t0, t1, out0, 0, t0, t0)
    rov = n_fft // hop
    hop_begin = (out0 + n_fft // 2) // hop
    hop_end = (out1 + n_fft // 2 + hop - 1) // hop
    f0, f1 = max(hop_begin - (rov - 1), 0), min(hop_end, T)
    return InverseShard(t0, t1, out0, out1 - out0, f0, f1)


@dataclass(frozen=True)
class BlendShard:
    k0: int          # owned segments [k0, k1)
    k1: int
    col0: int        # owned output columns [col0, col1)
    col1: int
    in0: int         # input columns needed to cut the owned segments, [in0, in1)
    in1: int
    left_halo: int   # number of the left neighbour's last segments that overlap the owned columns


def blend_shard(width: int, win: int, hop: int, world: int, rank: int) -> BlendShard:
    L = (width - (win - hop)) // hop                  # A2SB/diffusion.py:33
    k0, k1 = split_range(L, world)[rank]
    col0 = k0 * hop
    col1 = width if k1 >= L else k1 * hop
    in0, in1 = k0 * hop, (k1 - 1) * hop + win if k1 > k0 else k0 * hop
    left = min(ceil(win / hop) - 1, k0) if k1 > k0 else 0
    return BlendShard(k0, k1, col0, col1, in0, in1, left)


# ------------------------------------------------------------------------------------------------
# halo exchange
# ------------------------------------------------------------------------------------------------


def _neighbour_exchange(send_left: Optional[torch.Tensor], send_right: Optional[torch.Tensor],
                        recv_left: Optional[torch.Tensor], recv_right: Optional[torch.Tensor], rank: int, world: int):
    """One grouped neighbour exchange along the rank chain (no wrap-around)."""
    ops = []
    if rank > 0:
        if send_left is not None and send_left.numel():
            ops.append(dist.P2POp(dist.isend, send_left.contiguous(), rank - 1))
        if recv_left is not None and recv_left.numel():
            ops.append(dist.P2POp(dist.irecv, recv_left, rank - 1))
    if rank < world - 1:
        if send_right is not None and send_right.numel():
            ops.append(dist.P2POp(dist.isend, send_right.contiguous(), rank + 1))
        if recv_right is not None and recv_right.numel():
            ops.append(dist.P2POp(dist.irecv, recv_right, rank + 1))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()


def _cuda_forward(local, n_fft, hop, total_len, sample_first, t_range):
    from . import _capi, _lib
    return _lib.stft_forward(local, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25,
                             total_len=total_len, sample_first=sample_first, t_range=t_range)


def _cuda_inverse(local, n_fft, hop, n_frames, spec_t_first, out_range):
    from . import _capi, _lib
    return _lib.istft_inverse(local, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True,
                              power=4.0, n_frames=n_frames, spec_t_first=spec_t_first, out_range=out_range)


def sharded_forward(owned: torch.Tensor, length: int, n_fft: int, hop: int, rank: int, world: int,
                    compute: Callable = _cuda_forward) -> torch.Tensor:
    """owned: [B, own1-own0] samples of forward_shard(...).  Returns the owned frames [B, 3, n_fft/2, t1-t0].
    Assumes every rank owns at least n_fft/2 samples (true for anything worth sharding)."""
    sh = forward_shard(length, n_fft, hop, world, rank)
    left = [forward_shard(length, n_fft, hop, world, r) for r in range(world)]
    B = owned.shape[0]
    nl, nr = sh.own0 - sh.need0, sh.need1 - sh.own1            # halo widths I need
    lneed = left[rank - 1].need1 - left[rank - 1].own1 if rank > 0 else 0      # what my left neighbour needs of me
    rneed = left[rank + 1].own0 - left[rank + 1].need0 if rank < world - 1 else 0
    hl = owned.new_empty((B, max(nl, 0)))
    hr = owned.new_empty((B, max(nr, 0)))
    _neighbour_exchange(owned[:, :lneed] if lneed > 0 else None, owned[:, owned.shape[1] - rneed:] if rneed > 0 else None,
                        hl if nl > 0 else None, hr if nr > 0 else None, rank, world)
    local = torch.cat([hl, owned, hr], dim=1).contiguous()
    if sh.t1 <= sh.t0:
        return owned.new_empty((B, 3, n_fft // 2, 0))
    return compute(local, n_fft, hop, length, sh.need0, (sh.t0, sh.t1))


def sharded_inverse(owned: torch.Tensor, n_frames: int, n_fft: int, hop: int, rank: int, world: int,
                    compute: Callable = _cuda_inverse) -> torch.Tensor:
    """owned: [B, 3, rows, t1-t0] frames of inverse_shard(...).  Returns the owned samples [B, out_n]."""
    shards = [inverse_shard(n_frames, n_fft, hop, world, r) for r in range(world)]
    sh = shards[rank]
    nl, nr = max(sh.t0 - sh.f0, 0), max(sh.f1 - sh.t1, 0)
    lneed = max(shards[rank - 1].f1 - shards[rank - 1].t1, 0) if rank > 0 else 0
    rneed = max(shards[rank + 1].t0 - shards[rank + 1].f0, 0) if rank < world - 1 else 0
    shp = list(owned.shape)
    hl = owned.new_empty(shp[:-1] + [nl])
    hr = owned.new_empty(shp[:-1] + [nr])
    _neighbour_exchange(owned[..., :lneed] if lneed > 0 else None, owned[..., owned.shape[-1] - rneed:] if rneed > 0 else None,
                        hl if nl > 0 else None, hr if nr > 0 else None, rank, world)
    if sh.out_n == 0:
        return owned.new_empty((shp[0], 0))
    local = torch.cat([hl, owned, hr], dim=-1).contiguous()
    lo = sh.t0 - nl
    return compute(local, n_fft, hop, n_frames, lo, (sh.out0, sh.out_n))


def gather_concat(part: torch.Tensor, sizes: List[int], world: int, dim: int = -1) -> torch.Tensor:
    """all_gather of unequal pieces along `dim` (padded to the largest piece, then trimmed)."""
    if world == 1:
        return part
    mx = max(sizes)
    pad = list(part.shape)
    pad[dim] = mx - part.shape[dim]
    buf = torch.cat([part, part.new_zeros(pad)], dim=dim).contiguous() if pad[dim] else part.contiguous()
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    return torch.cat([o.narrow(dim if dim >= 0 else o.dim() + dim, 0, s) for o, s in zip(outs, sizes)], dim=dim)


def sharded_multidiffusion_vf(vf_model, x_owned: torch.Tensor, t_emb: torch.Tensor, width: int, win: int, hop: int,
                              batch_size: int, rank: int, world: int, gather=None, blend=None) -> torch.Tensor:
    """get_multidiffusion_vf (A2SB/diffusion.py:27-64) over a frame axis sharded by blend_shard(...).
    x_owned: [b, c, h, col1-col0] owned columns of the padded input.  Returns the owned output columns."""
    from . import _lib
    gather = gather or _lib.segment_gather
    blend = blend or _lib.segment_blend
    shards = [blend_shard(width, win, hop, world, r) for r in range(world)]
    sh = shards[rank]
    b = x_owned.shape[0]
    # 1. right halo of input columns (win - hop columns of the right neighbour's head)
    nr = max(sh.in1 - sh.col1, 0)
    lneed = max(shards[rank - 1].in1 - shards[rank - 1].col1, 0) if rank > 0 else 0
    hr = x_owned.new_empty(list(x_owned.shape[:-1]) + [nr])
    _neighbour_exchange(x_owned[..., :lneed] if lneed > 0 else None, None, None, hr if nr > 0 else None, rank, world)
    nk = sh.k1 - sh.k0
    local = torch.cat([x_owned, hr], dim=-1)[..., : max(sh.in1 - sh.in0, 0)].contiguous()
    # 2. segments + network on torch.chunk-sized mini-batches (reference chunking, diffusion.py:43-50)
    segs = gather(local, win, hop) if nk > 0 else x_owned.new_empty((0,) + tuple(x_owned.shape[1:-1]) + (win,))
    outs = torch.empty_like(segs)
    if segs.shape[0]:
        n_chunks = ceil(segs.shape[0] / batch_size)
        row = 0
        for ch, te in zip(torch.chunk(segs, n_chunks), torch.chunk(t_emb.repeat(nk, 1), n_chunks)):
            o = vf_model(ch, te)
            outs[row:row + o.shape[0]].copy_(o)
            row += o.shape[0]
    # 3. left halo of network outputs: the left neighbour's last segments overlap my first columns
    nh = sh.left_halo
    rsend = shards[rank + 1].left_halo if rank < world - 1 else 0
    v = outs.reshape((b, nk) + tuple(outs.shape[1:])) if nk else outs.reshape((b, 0) + tuple(outs.shape[1:]))
    hl = outs.new_empty((b, nh) + tuple(outs.shape[1:]))
    _neighbour_exchange(None, v[:, nk - rsend:] if rsend > 0 else None, hl if nh > 0 else None, None, rank, world)
    # 4. blend [halo segments + own segments] over columns starting at (k0 - nh)*hop, keep the owned columns
    allseg = torch.cat([hl, v], dim=1).reshape((b * (nh + nk),) + tuple(outs.shape[1:])).contiguous()
    w_local = (nh + nk - 1) * hop + win if (nh + nk) else 0
    if w_local == 0:
        return x_owned.new_empty(list(x_owned.shape[:-1]) + [0])
    full = blend(allseg, b, w_local, win, hop)
    off = sh.col0 - (sh.k0 - nh) * hop
    return full[..., off:off + (sh.col1 - sh.col0)].contiguous()
