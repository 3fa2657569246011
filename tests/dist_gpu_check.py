"""torchrun --nproc-per-node N tests/dist_gpu_check.py : the sharded long-clip path on N GPUs (NCCL halo
exchange + all_gather, CUDA kernels) must equal the unsharded CUDA result bit for bit.  Prints one line
with the device-timed sharded round trip of a 10-minute clip."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_intelligence_b200 import _capi, _lib, diffusion as D, sharding as S  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_fft, hop = 2048, 512
    L = 44100 * 600 + 77
    g = torch.Generator(device=dev).manual_seed(1234)          # same seed on every rank: same clip
    wav = (0.3 * torch.randn(1, L, generator=g, device=dev)).clamp_(-1, 1)
    T = 1 + L // hop
    # unsharded reference on every rank
    spec_ref = _lib.stft_forward(wav, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25)
    y_ref = _lib.istft_inverse(spec_ref, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True, power=4.0)
    fs = S.forward_shard(L, n_fft, hop, world, rank)
    owned = wav[:, fs.own0:fs.own1].contiguous()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for it in range(3):
        dist.barrier(device_ids=[local]); torch.cuda.synchronize()
        ev[0].record()
        spec = S.sharded_forward(owned, L, n_fft, hop, rank, world)
        y = S.sharded_inverse(spec, T, n_fft, hop, rank, world)
        ysizes = [S.inverse_shard(T, n_fft, hop, world, r).out_n for r in range(world)]
        full_y = S.gather_concat(y, ysizes, world)
        ev[1].record(); torch.cuda.synchronize()
    ms = torch.tensor([ev[0].elapsed_time(ev[1])], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    assert torch.equal(spec, spec_ref[..., fs.t0:fs.t1]), "sharded forward differs from the unsharded result"
    assert torch.equal(full_y, y_ref), "sharded inverse differs from the unsharded result"
    # segment blend, sharded along the frame axis (identity + affine stub network)
    xp = D.multidiffusion_pad_inputs(spec_ref[:, :, :256], 256, 128)
    W = xp.shape[-1]
    t_emb = torch.zeros(1, 4, device=dev)
    ref = D.get_multidiffusion_vf(lambda a, t: a * 2 + 0.1, xp, t_emb, 256, 128, 16)
    bs = S.blend_shard(W, 256, 128, world, rank)
    out = S.sharded_multidiffusion_vf(lambda a, t: a * 2 + 0.1, xp[..., bs.col0:bs.col1].contiguous(), t_emb, W, 256, 128, 16,
                                      rank, world)
    bsizes = [S.blend_shard(W, 256, 128, world, r).col1 - S.blend_shard(W, 256, 128, world, r).col0 for r in range(world)]
    full = S.gather_concat(out, bsizes, world)
    assert torch.equal(full, ref), "sharded blend differs from the unsharded result"
    # peer-memory round trip: halos pulled from the neighbours' symmetric buffers, K2 storing its output into every GPU's
    # result buffer (peer stores, then the multicast store when the fabric has one) -- no collective on the data path
    peer_ms = {}
    Lbig = int(os.environ.get("A2SB_DIST_LONG", "0"))      # e.g. 158760000: also time the variants on a 1 h clip
    cases = [("nccl", None, "fused", 2, L), ("peers", False, "fused", 2, L), ("multicast", True, "fused", 2, L), ("ce", None, "ce", 2, L),
             ("ce4", None, "ce", 4, L), ("overlap4", True, "fused", 4, L)]
    if Lbig:
        cases += [(f"{m}@1h/{rd}", mc, gm, rd, Lbig) for m, mc, gm in (("nccl", None, "fused"), ("multicast", True, "fused"), ("overlap", True, "fused"))
                  for rd in (2, 4, 8)]
    for mode, mc, gm, rounds_, Lc in cases:
        big = Lc != L
        if big:
            gb = torch.Generator(device=dev).manual_seed(99)
            src = (0.3 * torch.randn(1, Lc, generator=gb, device=dev)).clamp_(-1, 1)
        else:
            src = wav
        if mode.startswith("nccl"):
            rt = S.LongClipRoundTrip(Lc, n_fft, hop, rank, world, dev, rounds=rounds_)
        else:
            rt = S.PeerLongClipRoundTrip(Lc, n_fft, hop, rank, world, dev, rounds=rounds_, multicast=mc, gather=gm,
                                         overlap=mode.startswith("overlap"))
            if mc and not rt.multicast:
                peer_ms[mode] = None
                continue
        for c in range(rounds_):
            sh = rt.mine[c]
            rt.owned_wav(c).copy_(src[:, sh.own0:sh.own1])
        if not mode.startswith("nccl"):
            rt.ready()
        best = 1e9
        for it in range(5):
            dist.barrier(device_ids=[local]); torch.cuda.synchronize()
            ev[0].record()
            if mode.startswith("nccl"):
                rt.exchange_wav_allgather()
            else:
                rt.pull_halos()
            y2 = rt.run()
            ev[1].record(); torch.cuda.synchronize()
            t_ = torch.tensor([ev[0].elapsed_time(ev[1])], device=dev)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            best = min(best, float(t_))
        if not big:
            assert torch.equal(y2, y_ref), f"{mode} round trip differs from the unsharded result"
        elif rank == 0:
            sp1 = _lib.stft_forward(src, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25)
            y1 = _lib.istft_inverse(sp1, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True, power=4.0)
            assert torch.equal(y2, y1), f"{mode} round trip differs from the unsharded result"
            del sp1, y1
        peer_ms[mode] = round(best, 4)
        del rt, y2, src
        torch.cuda.empty_cache()
    if rank == 0:
        print("long-clip round trip, 600 s (and 1 h), ms (max over ranks, best of 5):", peer_ms, flush=True)
    if rank == 0:
        print(f"dist_gpu_check ok: world {world}, 600 s clip sharded round trip {float(ms):.3f} ms "
              f"({600.0 / (float(ms) * 1e-3):.0f} audio-s/s incl. halo exchange and all_gather)", flush=True)
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
