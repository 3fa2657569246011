"""Multi-GPU sharding (audio_intelligence_b200/sharding.py): integer planners, and the halo exchange
+ gather logic run as a real world_size-2 (and 3) `gloo` job on CPU with the oracle injected as the
compute backend.  The CUDA backend of the same code path is exercised by tests/test_parity_gpu.py
(sharded ranges bit-identical) and by `bench.py --gpus N`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import a2sb_oracle as O
from audio_intelligence_b200 import sharding as S


def test_split_range_properties():
    for n in (0, 1, 15, 16, 17, 862, 310079):
        for world in (1, 2, 3, 4, 8):
            for align in (1, 16):
                cuts = S.split_range(n, world, align)
                assert cuts[0][0] == 0 and cuts[-1][1] == n
                assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
                assert all(lo <= hi for lo, hi in cuts)
                assert all(lo % align == 0 for lo, hi in cuts if lo < n)


@pytest.mark.parametrize("L,n_fft,world", [(441000, 2048, 2), (441000, 2048, 8), (158760000, 2048, 8), (70001, 1024, 3)])
def test_forward_and_inverse_shards_cover_everything(L, n_fft, world):
    hop = n_fft // 4
    T = 1 + L // hop
    fs = [S.forward_shard(L, n_fft, hop, world, r) for r in range(world)]
    assert fs[0].t0 == 0 and fs[-1].t1 == T and fs[0].own0 == 0 and fs[-1].own1 == L
    for a, b in zip(fs, fs[1:]):
        assert a.t1 == b.t0 and a.own1 == b.own0
    for sh in fs:
        if sh.t1 > sh.t0:
            # every sample a frame of this shard touches (after reflection) lies in [need0, need1)
            lo, hi = sh.t0 * hop - n_fft // 2, (sh.t1 - 1) * hop + n_fft // 2 - 1
            idx = O.reflect_index(np.array([lo, min(hi, lo + 5000), hi, max(lo, hi - 5000)]), L)
            assert idx.min() >= sh.need0 and idx.max() < sh.need1
            # halos come from the direct neighbours only
            assert sh.own0 - sh.need0 <= n_fft // 2 and sh.need1 - sh.own1 <= n_fft // 2
    inv = [S.inverse_shard(T, n_fft, hop, world, r) for r in range(world)]
    assert inv[0].out0 == 0 and inv[-1].out0 + inv[-1].out_n == hop * (T - 1)
    for a, b in zip(inv, inv[1:]):
        assert a.out0 + a.out_n == b.out0
    for sh in inv:
        assert sh.t0 - sh.f0 <= 3 and sh.f1 - sh.t1 <= 3 and sh.out0 % hop == 0


def test_blend_shard_known_answers():
    # BASELINE config 3: 310,144 padded frames -> 2422 segments of 256 hopped by 128
    sh = [S.blend_shard(310144, 256, 128, 8, r) for r in range(8)]
    assert sh[0].k0 == 0 and sh[-1].k1 == 2422 and sh[-1].col1 == 310144
    assert all(a.k1 == b.k0 and a.col1 == b.col0 for a, b in zip(sh, sh[1:]))
    assert all(s.in1 - s.col1 == 128 for s in sh[:-1]) and sh[-1].in1 == 310144
    assert sh[0].left_halo == 0 and all(s.left_halo == 1 for s in sh[1:])


# ---------------------------------------------------------------------------------- gloo job


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cpu_forward(local, n_fft, hop, total_len, sample_first, t_range):
    """Oracle-backed stand-in for the CUDA call: frames [t0, t1) from a local sample window."""
    t0, t1 = t_range
    outs = []
    for w in local.numpy():
        idx = (np.arange(t0, t1)[:, None] * hop + np.arange(n_fft)[None, :]) - n_fft // 2
        src = O.reflect_index(idx, total_len) - sample_first
        assert src.min() >= 0 and src.max() < w.shape[0], "halo exchange delivered too few samples"
        frames = w[src] * O.padded_window(n_fft, n_fft, None, np.float32)[None, :]
        c = np.fft.rfft(frames.astype(np.float64), axis=1).T.astype(np.complex64)
        outs.append(O.power_scale(O.drop_dc(O.complex_to_mag_phase(c)), 0.25, [0], 1e-9))
    return torch.from_numpy(np.stack(outs))


def _cpu_inverse(local, n_fft, hop, n_frames, spec_t_first, out_range):
    o0, on = out_range
    outs = []
    for s in local.numpy():
        T = n_frames
        full = np.zeros((3, n_fft // 2, T), np.float32)
        full[1] = 1.0
        lo, hi = max(spec_t_first, 0), min(spec_t_first + s.shape[-1], T)      # the local buffer may overhang [0, T)
        full[..., lo:hi] = s[..., lo - spec_t_first:hi - spec_t_first]
        hop_begin, hop_end = (o0 + n_fft // 2) // hop, (o0 + on + n_fft // 2 + hop - 1) // hop
        assert spec_t_first <= max(hop_begin - 3, 0) and spec_t_first + s.shape[-1] >= min(hop_end, T), "frame halo too small"
        outs.append(O.inverse_chain(full, n_fft, hop)[o0:o0 + on])
    return torch.from_numpy(np.stack(outs))


def _worker(rank, world, port, L, n_fft, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hop = n_fft // 4
        wav = np.stack([O.synth_noise(L, 1000 + i) for i in range(2)])
        fs = S.forward_shard(L, n_fft, hop, world, rank)
        owned = torch.from_numpy(wav[:, fs.own0:fs.own1].copy())
        spec = S.sharded_forward(owned, L, n_fft, hop, rank, world, compute=_cpu_forward)
        T = 1 + L // hop
        sizes = [max(S.forward_shard(L, n_fft, hop, world, r).t1 - S.forward_shard(L, n_fft, hop, world, r).t0, 0) for r in range(world)]
        full_spec = S.gather_concat(spec, sizes, world)
        y = S.sharded_inverse(spec, T, n_fft, hop, rank, world, compute=_cpu_inverse)
        ysizes = [S.inverse_shard(T, n_fft, hop, world, r).out_n for r in range(world)]
        full_y = S.gather_concat(y, ysizes, world)
        # segment blend over the gathered spectrogram, sharded along the frame axis
        win, bhop = 64, 32
        xp = torch.from_numpy(O.multidiffusion_pad_inputs(full_spec.numpy()[:1, :, :8], win, bhop))
        W = xp.shape[-1]
        bs = S.blend_shard(W, win, bhop, world, rank)
        net = lambda a, t: a * 1.7 - 0.3 + t[:, :1, None, None]
        gather = lambda x, w, h: torch.from_numpy(O.segment_gather(x.numpy(), w, h))
        blend = lambda sg, b, w_, w, h: torch.from_numpy(O.segment_blend(sg.numpy(), b, w_, w, h))
        t_emb = torch.full((1, 4), 0.25)
        out = S.sharded_multidiffusion_vf(net, xp[..., bs.col0:bs.col1].contiguous(), t_emb, W, win, bhop, 5, rank, world,
                                          gather=gather, blend=blend)
        bsizes = [S.blend_shard(W, win, bhop, world, r).col1 - S.blend_shard(W, win, bhop, world, r).col0 for r in range(world)]
        full_out = S.gather_concat(out, bsizes, world)
        if rank == 0:
            q.put((full_spec.numpy(), full_y.numpy(), xp.numpy(), full_out.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,L,n_fft", [(2, 20011, 512), (3, 41000, 1024)])
def test_sharded_path_equals_unsharded(world, L, n_fft):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, L, n_fft, q)) for r in range(world)]
    for p in procs:
        p.start()
    spec, y, xp, blended = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    hop = n_fft // 4
    for i in range(2):
        wav = O.synth_noise(L, 1000 + i)
        ref = O.forward_chain(wav, n_fft, hop)
        np.testing.assert_array_equal(spec[i], ref)                       # same arithmetic, sharded index math
        np.testing.assert_array_equal(y[i], O.inverse_chain(ref, n_fft, hop))
    t = np.full((1, 4), 0.25, np.float32)
    net = lambda a, e: a * np.float32(1.7) - np.float32(0.3) + e[:, :1, None, None]
    np.testing.assert_array_equal(blended, O.get_multidiffusion_vf(net, xp, t, 64, 32, 5))


# ---------------------------------------------------------------------------------- pre-padded buffers (fast path)


def _cpu_forward_into(wav_local, spec_buf, col_off, n_fft, hop, total_len, sample_first, t_range):
    spec_buf[..., col_off:col_off + (t_range[1] - t_range[0])] = _cpu_forward(wav_local, n_fft, hop, total_len, sample_first, t_range)


def _cpu_inverse_into(spec_local, out, n_fft, hop, n_frames, spec_t_first, out_range):
    out[:, :out_range[1]] = _cpu_inverse(spec_local, n_fft, hop, n_frames, spec_t_first, out_range)


def _worker_prepadded(rank, world, port, L, n_fft, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hop = n_fft // 4
        wav = O.synth_noise(L, 1000)
        ys = []
        for rounds in (1, 3):
            rt = S.LongClipRoundTrip(L, n_fft, hop, rank, world, "cpu", rounds=rounds, fwd_into=_cpu_forward_into,
                                     inv_into=_cpu_inverse_into)
            rt.spec.fill_(float("nan"))
            for c in range(rounds):
                rt.wav[c].fill_(float("nan"))               # any sample that is not delivered shows up in the result
                sh = rt.mine[c]
                rt.owned_wav(c).copy_(torch.from_numpy(wav[None, sh.own0:sh.own1].copy()))
            rt.exchange_wav() if rounds == 1 else rt.exchange_wav_allgather()
            ys.append(rt.run().clone())
        assert torch.equal(ys[0], ys[1])
        y = ys[1]
        # blend: two steps on the owned columns of a padded random input, affine network stub
        win, bhop, width = 64, 32, 64 + 32 * 37
        x_full = np.random.default_rng(5).standard_normal((1, 3, 8, width)).astype(np.float32)
        gather = lambda x, segs, w, h: segs.copy_(torch.from_numpy(O.segment_gather(x.numpy(), w, h)))

        def blend_window(segs, out, b, W, w, h, off, cnt):
            out[..., :cnt] = torch.from_numpy(O.segment_blend(segs.numpy(), b, W, w, h))[..., off:off + cnt]
        sb = S.ShardedBlend(3, 8, width, win, bhop, rank, world, "cpu", gather_into=gather, blend_window=blend_window)
        sb.owned_x.copy_(torch.from_numpy(x_full[..., sb.sh.col0:sb.sh.col1].copy()))
        sb.prime()
        net = lambda a, t: a * 1.7 - 0.3 + t[:, :1, None, None]
        t_emb = torch.full((1, 4), 0.25)
        outs = []
        for _ in range(3):
            sb.step(net, t_emb, batch_size=5)
            sb.swap()
            sizes = [sh.col1 - sh.col0 for sh in sb.shards]
            outs.append(S.gather_concat(sb.owned_x.contiguous(), sizes, world))
        if rank == 0:
            q.put((y.numpy(), x_full, [o.numpy() for o in outs]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,L,n_fft", [(2, 20011, 512), (3, 41000, 1024)])
def test_prepadded_buffers_equal_unsharded(world, L, n_fft):
    """LongClipRoundTrip / ShardedBlend: halos received into the edges of pre-padded buffers, one all_gather_into_tensor."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_prepadded, args=(r, world, port, L, n_fft, q)) for r in range(world)]
    for p in procs:
        p.start()
    y, x_full, outs = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    hop = n_fft // 4
    wav = O.synth_noise(L, 1000)
    np.testing.assert_array_equal(y[0], O.inverse_chain(O.forward_chain(wav, n_fft, hop), n_fft, hop))
    t = np.full((1, 4), 0.25, np.float32)
    net = lambda a, e: a * np.float32(1.7) - np.float32(0.3) + e[:, :1, None, None]
    ref = x_full
    for o in outs:
        ref = O.get_multidiffusion_vf(net, ref, t, 64, 32, 5)
        np.testing.assert_array_equal(o, ref)


# ------------------------------------------------------------------------------------------------
# peer-memory round trip (PeerLongClipRoundTrip): host logic with a single-process stand-in for symmetric memory
# ------------------------------------------------------------------------------------------------


class _FakeSymm:
    """torch.distributed._symmetric_memory for ONE process playing every rank in turn: empty() allocates a normal tensor,
    rendezvous() registers it under (current rank, allocation index) and returns a handle whose get_buffer() views the
    registered tensor of any rank, whose buffer_ptrs are (rank, allocation) tokens and whose barrier() does nothing."""

    def __init__(self, world):
        self.world, self.rank, self.bufs, self.count = world, 0, {}, {}

    def empty(self, n, dtype=None, device=None):
        return torch.full((int(n),), float("nan"), dtype=dtype)

    def rendezvous(self, t, group):
        idx = self.count.get(self.rank, 0)
        self.count[self.rank] = idx + 1
        self.bufs[(self.rank, idx)] = t
        outer = self

        class H:
            multicast_ptr = 0
            buffer_ptrs = [r * 1_000_000_000_000 + idx * 1_000_000_000 for r in range(outer.world)]   # byte "addresses": rank, allocation

            def get_buffer(self_, rank, sizes, dtype, off=0):
                n = int(np.prod(sizes))
                return outer.bufs[(rank, idx)][off: off + n].view(*sizes)

            def barrier(self_):
                pass
        return H()


@pytest.mark.parametrize("gather", ["fused", "ce"])
@pytest.mark.parametrize("world,rounds", [(2, 1), (3, 2), (4, 3)])
def test_peer_round_trip_host_logic(world, rounds, gather):
    """Every rank of a PeerLongClipRoundTrip, played by one process over a fake symmetric memory and the oracle as K1 / K2:
    the halo pull plan reads the right samples out of the right neighbour buffers, the (fused or pushed) gather lands every
    piece at its offset of every rank's result, and each rank's result equals the unsharded round trip bit for bit."""
    n_fft, hop, L = 512, 128, 128 * 16 * 3 * world * rounds + 12345
    wav = O.synth_noise(L, 4242)
    ref = _cpu_inverse(_cpu_forward(torch.from_numpy(wav[None]), n_fft, hop, L, 0, (0, 1 + L // hop)), n_fft, hop, 1 + L // hop, 0,
                       (0, hop * (L // hop)))
    fake = _FakeSymm(world)

    def inv_mirrored(spec_local, out, n_fft_, hop_, n_frames, t_first, out_range, mirrors, multicast):
        _cpu_inverse_into(spec_local, out, n_fft_, hop_, n_frames, t_first, out_range)
        assert not multicast
        for m in mirrors or []:                            # "peer stores": the token names (rank, allocation) + a byte offset
            r, rest = divmod(int(m), 1_000_000_000_000)
            idx, off = divmod(rest, 1_000_000_000)
            fake.bufs[(r, idx)][off // 4: off // 4 + out.shape[1]].copy_(out[0])
    rts = []
    for r in range(world):
        fake.rank = r
        rts.append(S.PeerLongClipRoundTrip(L, n_fft, hop, r, world, "cpu", rounds=rounds, gather=gather, fwd_into=_cpu_forward_into,
                                           inv_mirrored=inv_mirrored, symm=fake))
    for rt in rts:
        rt.spec.fill_(float("nan"))
        for c in range(rounds):
            sh = rt.mine[c]
            rt.owned_wav(c).copy_(torch.from_numpy(wav[None, sh.own0:sh.own1].copy()))
    for rt in rts:
        rt.pull_halos()
    outs = [rt.run() for rt in rts]
    for y in outs:
        assert torch.equal(y, ref)
