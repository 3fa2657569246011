#!/usr/bin/env python
"""bench.py -- STFT+iSTFT audio-seconds per second @44.1 kHz (BASELINE.json metric) on N B200s.

A step = one pass of the hot path over one batch: K1 (wav -> A2SB spectrogram) then K2
(spectrogram -> wav) over BASELINE config 2: 256 x 10 s mono clips, n_fft 2048 / hop 512, fp32,
per GPU (weak scaling: clips are sharded contiguously across ranks, no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one JSON line on rank 0)
  python bench.py --impl reference [...]                         the reference's CPU path (port)
  torchrun ... bench.py --gpus N ...                             one rank per GPU (driver launch)

`value`   : device-timed (CUDA events, max over ranks), inputs resident in HBM.
`e2e`     : same metric through the C ABI's host-buffer entry point (pinned host wav in, wav out;
            H2D and D2H inside the timed region).
`roofline`: slower of K1/K2, algorithmic bytes / mean CUDA-event duration vs MEASURED_PEAKS.json.
`cpu_baseline`: oracle/torch_port.py (the reference's torch.stft/istft call sequence) on the host.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 44100
N_FFT, HOP = 2048, 512
CLIP_LEN = int(os.environ.get("A2SB_BENCH_CLIP_LEN", 10 * SR))   # experiments only; the bench config is 10 s
METRIC = "STFT+iSTFT audio-sec/sec @44.1kHz"
UNIT = "audio-s/s"


def workload_config(clips: int, world: int) -> dict:
    T = 1 + CLIP_LEN // HOP
    return {
        "workload": "BASELINE configs[1]: batched round trip, 256 x 10 s mono clips @44.1 kHz, n_fft=2048 hop=512, "
                    "A2SB shipped forward chain (STFT->mag/cos/sin->drop DC->power .25) + inverse chain "
                    "(power 4->add DC->phase fix->complex->iSTFT)",
        "clips_per_gpu": clips, "clip_len": CLIP_LEN, "n_fft": N_FFT, "hop": HOP, "frames": T,
        "global_clips": clips * world,
        "l2": "inputs larger than L2 (0.45 GB wav + 2.7 GB spectrogram per step vs 126 MB L2): no flush",
        "parallelism": f"clips sharded contiguously over {world} rank(s); no data-path collective",
    }


def algorithmic_bytes(clips: int) -> tuple[int, int]:
    """forward 4L + 6NT, inverse 6NT + 4H(T-1) bytes per clip."""
    T = 1 + CLIP_LEN // HOP
    fwd = 4 * CLIP_LEN + 6 * N_FFT * T
    inv = 6 * N_FFT * T + 4 * HOP * (T - 1)
    return fwd * clips, inv * clips


# --------------------------------------------------------------------------------------- clocks


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML, ~every 20 ms)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None
        self._t = threading.Thread(target=self._run, daemon=True)

    _NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
              0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
              0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def sample(self):
        if self._h is None:
            return
        try:
            self.samples.append(float(self._nv.nvmlDeviceGetClockInfo(self._h, self._nv.NVML_CLOCK_SM)))
            bits = int(self._nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
            for b, n in self._NAMES.items():
                if bits & b and n != "gpu_idle":
                    self.reasons.add(n)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self.sample()
            self._stop.wait(0.02)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self.sample()
        self._stop.set()
        self._t.join()

    def summary(self) -> dict:
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this rank (and the pinned host buffers it is about to allocate: first touch) to the NUMA node its GPU hangs
    off, so that the e2e leg's H2D/D2H traffic of the 8 ranks does not cross the socket interconnect.  Best effort:
    returns the node, or None when the topology cannot be read."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


# --------------------------------------------------------------------------------------- CPU arm


def find_reference_dir():
    """The reference's A2SB sources, when they are reachable at run time: $A2SB_REFERENCE, the driver's install location
    baseline/_ref/A2SB, or /root/reference/A2SB (build container only -- the GPU box has neither unless the driver put
    one there).  Returns None when absent; the CPU arm then runs the port and says so."""
    for d in (os.environ.get("A2SB_REFERENCE"), os.path.join(ROOT, "baseline", "_ref", "A2SB"),
              os.path.join(ROOT, "baseline", "_ref"), "/root/reference/A2SB"):
        if d and os.path.isfile(os.path.join(d, "audio_transforms", "transforms.py")):
            return d
    return None


def cpu_chain(svd_fix: bool):
    """(kind, describe, roundtrip(wavs) -> wavs): the reference's own transforms.py driven exactly like vocode_stft
    (A2SB_lightning_module.py:89-100: per-clip loop over apply_audio_transforms) when its sources are present
    ("reference"), else oracle/torch_port.py, the same sequence of torch library calls ("port")."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    ref = find_reference_dir()
    if ref is not None:
        try:
            import make_golden
            make_golden.REF = ref
            T, _D, _U, _C = make_golden.import_reference()
            fwd, inv, inv_nosvd = make_golden.chains(T, N_FFT, HOP)
            chain_inv = inv if svd_fix else inv_nosvd

            def rt(wavs):
                return torch.stack([T.apply_audio_transforms(T.apply_audio_transforms(w, fwd)[0], chain_inv)[0] for w in wavs])
            return "reference", f"unmodified {ref}/audio_transforms/transforms.py (apply_audio_transforms, per-clip loop)", rt
        except Exception as e:                      # e.g. torchaudio missing on the box
            sys.stderr.write(f"bench.py: reference at {ref} not importable ({e!r}); using the port\n")
    import torch_port
    return ("port", "oracle/torch_port.py = the reference's torch.stft/torch.istft call sequence (reference sources not "
            "present on this box)", lambda wavs: torch_port.roundtrip(wavs, N_FFT, HOP, svd_fix=svd_fix))


def cpu_inputs(n_clips: int):
    import torch
    g = torch.Generator().manual_seed(1000)
    return (0.3 * torch.randn(n_clips, CLIP_LEN, generator=g)).clamp_(-1, 1)


def cpu_time_pass(rt, wav, reps: int = 1):
    """Best-of-`reps` seconds for one pass of `rt` over `wav` (inputs generated and MKL warmed up by the caller)."""
    import torch
    best, out = float("inf"), None
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            out = rt(wav)
            best = min(best, time.perf_counter() - t0)
    return best, out


def cpu_baseline_leg(gpu_out_fn=None) -> dict:
    """The `cpu_baseline` object of our arm's line.  The throughput figure is produced by the SAME code as the
    `--impl reference` arm -- bench.py --impl reference in a fresh process (no CUDA context, no leftover threads) -- so the
    two numbers of a round are one measurement procedure; the SVD-fix variant and the parity check run in this process."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=900)
    ref_line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    res = dict(ref_line["cpu_baseline"])
    res["ms_per_step"], res["clips_per_step"] = ref_line["ms_per_step"], ref_line["clips_per_step"]
    kind, desc, rt = cpu_chain(False)
    _, _, rt_svd = cpu_chain(True)
    wav = cpu_inputs(3)
    with torch.no_grad():
        out = rt(wav[:2])
        rt_svd(wav[:1])
    t_svd, _ = cpu_time_pass(rt_svd, wav[:3])
    res["with_svd_fix"] = {"value": 30.0 / t_svd, "sample": "3 clips, shipped inverse chain incl. per-bin 2x2 SVD"}
    res["torch"] = torch.__version__
    if gpu_out_fn is not None:                      # parity of the measured path against the CPU chain
        import numpy as np
        got = gpu_out_fn(wav[:2])
        ref = out[:2].double().numpy()
        err = got.astype(np.float64) - ref
        res["parity_snr_db_vs_cpu"] = float(10 * np.log10((ref * ref).sum() / max((err * err).sum(), 1e-300)))
    return res


def run_reference(args) -> None:
    """The reference's own CPU implementation of the path on this box's host cores: every step is one pass over a fixed
    sample of the workload's clips; inputs, thread count and the warm-up are set up ONCE outside the timed loop."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))       # the workload our arm runs at this N
    n = int(os.environ.get("A2SB_BENCH_REF_CLIPS", "32"))
    kind, desc, rt = cpu_chain(False)
    wav = cpu_inputs(n)
    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            rt(wav)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rt(wav)
        dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = 10.0 * n / dt
    sample = (f"each step = one pass over {n} of the {args.clips * world} clips of the workload: {desc}, no "
              f"SVDFixMagInstPhase (the faster of the reference's two shipped inverse chains), {cores} MKL threads")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": dt * 1e3, "clips_per_step": n,
            "audio_s_per_step": 10.0 * n, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.clips, world),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)




# --------------------------------------------------------------------------------------- plugin-shaped e2e


def e2e_plugin_leg(dev, clips: int) -> dict:
    """The round trip through the reference-facing Python API exactly as the reference's call sites use it: CPU fp32
    tensors in, CPU tensors out (A2SB/datasets/datasets.py:235 hands `apply_audio_transforms` a CPU waveform and keeps the
    CPU spectrogram; A2SB_lightning_module.py:202 hands `vocode_stft` a CPU spectrogram and takes a CPU waveform).  Pageable
    host memory, synchronous semantics; the spectrogram crosses PCIe in both directions (10.6 MB per clip each way), which
    `a2sb_roundtrip_host` (the `e2e` key) avoids by keeping it in HBM.
      config 1: one 10 s clip, wall-clock latency (median of 30 after warm-up);
      config 2: `clips` clips -- per-clip loop like vocode_stft (:97-98), and one batched call (leading batch dimension)."""
    import torch
    from audio_intelligence_b200.audio_transforms import transforms as T
    fwd = [T.ComplexSpectrogram(N_FFT, N_FFT, HOP), T.ComplexToMagInstPhase(), T.SpectrogramDropDCTerm(),
           T.PowerScaleSpectrogram(0.25, [0])]
    inv = [T.PowerScaleSpectrogram(4, [0]), T.SpectrogramAddDCTerm(), T.SVDFixMagInstPhase(), T.MagInstPhaseToComplex(),
           T.InverseComplexSpectrogram(N_FFT, N_FFT, HOP)]
    g = torch.Generator().manual_seed(1000)
    wavs = (0.3 * torch.randn(clips, CLIP_LEN, generator=g)).clamp_(-1, 1)

    def one(w):
        spec, _ = T.apply_audio_transforms(w, fwd)
        assert not spec.is_cuda
        y, _ = T.apply_audio_transforms(spec, inv)
        assert not y.is_cuda
        return y

    ts = []
    for i in range(35):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        y1 = one(wavs[0])
        ts.append(time.perf_counter() - t0)
    ts = ts[5:]
    rec = {"config1_single_clip": {"median_ms": statistics.median(ts) * 1e3, "min_ms": min(ts) * 1e3,
                                   "audio_s_per_s": 10.0 / statistics.median(ts)}}
    n_loop = min(clips, 64)
    for _ in range(2):
        t0 = time.perf_counter()
        for w in wavs[:n_loop]:
            one(w)
        t_loop = time.perf_counter() - t0
    for _ in range(2):
        t0 = time.perf_counter()
        yb = one(wavs)
        t_batch = time.perf_counter() - t0
    spec_bytes = 3 * (N_FFT // 2) * (1 + CLIP_LEN // HOP) * 4
    rec["config2_per_clip_loop"] = {"clips": n_loop, "audio_s_per_s": 10.0 * n_loop / t_loop, "ms_per_clip": t_loop / n_loop * 1e3}
    rec["config2_batched_call"] = {"clips": clips, "audio_s_per_s": 10.0 * clips / t_batch, "ms": t_batch * 1e3,
                                   "matches_single_clip_call": bool(torch.equal(yb[0], y1))}
    rec["bytes_per_clip"] = {"h2d": CLIP_LEN * 4 + spec_bytes, "d2h": spec_bytes + HOP * (CLIP_LEN // HOP) * 4}
    rec["api"] = ("audio_transforms.transforms.apply_audio_transforms(forward chain) then (inverse chain), pageable CPU "
                  "tensors in and out, one fused kernel per chain")
    return rec


# --------------------------------------------------------------------------------------- long audio (config 3)


def long_audio_leg(rank: int, world: int, local: int, dev, reps: int = 5) -> dict | None:
    """BASELINE configs[2]: ONE 1 h clip (L = 158,760,000) strong-scaled over the ranks by contiguous frame / segment
    ranges (audio_intelligence_b200/sharding.py): (a) the transform round trip = sample-halo exchange, K1, frame-halo
    exchange, K2, all-gather of the waveform; (b) one segment-blend step at config-3 size ([1, 3, 1024, 310144], 2422
    segments of 256 hopped by 128, identity network stub working in place) = column-halo exchange, K3, segment-halo
    exchange, K4 on the owned columns.  Device-timed (CUDA events, max over ranks); rank 0 also runs the unsharded
    computation on its own GPU for the speed-up and checks the sharded result against it bit for bit."""
    import torch
    import torch.distributed as dist
    from audio_intelligence_b200 import _capi, _lib, sharding as S

    L = int(os.environ.get("A2SB_BENCH_LONG_LEN", 158_760_000))
    T = 1 + L // HOP

    def sync():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def timed(fn, marks: int):
        """fn(ev) records `marks` events after its phases; returns per-phase ms (max over ranks) of the best rep by total."""
        best = None
        for it in range(reps + 2):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(marks + 1)]
            sync()
            ev[0].record()
            fn(ev)
            torch.cuda.synchronize()
            ms = torch.tensor([ev[i].elapsed_time(ev[i + 1]) for i in range(marks)] + [ev[0].elapsed_time(ev[marks])],
                              device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if it >= 2 and (best is None or float(ms[-1]) < float(best[-1])):
                best = ms
        return [float(v) for v in best]

    rec: dict = {"clip_s": L / SR, "samples": L, "frames": T, "world": world}
    # ---- (a) transform round trip
    rounds = 1 if world == 1 else int(os.environ.get("A2SB_BENCH_LONG_ROUNDS", "2"))
    rt = S.LongClipRoundTrip(L, N_FFT, HOP, rank, world, dev, rounds=rounds)
    for c in range(rounds):                                               # config 3 is generated per shard
        g = torch.Generator(device=dev).manual_seed(1000 + c * world + rank)
        rt.owned_wav(c).copy_((0.3 * torch.randn(rt.owned_wav(c).shape, generator=g, device=dev)).clamp_(-1, 1))
    final = torch.empty(rt.pieces * rt.out_max, dtype=torch.float32, device=dev)

    halo = rt.exchange_wav if os.environ.get("A2SB_BENCH_LONG_HALO", "allgather") == "p2p" else rt.exchange_wav_allgather

    def roundtrip(ev):
        halo(); ev[1].record()
        rt.run(final); ev[2].record()

    def roundtrip_nogather(ev):
        halo(); ev[1].record()
        rt.run(None, gather=False); ev[2].record()
    h_ms, r_ms, tot = timed(roundtrip, 2)
    _h2, _r2, tot_ng = timed(roundtrip_nogather, 2)
    halo(); rt.run(final)
    rec["round_trip"] = {"ms": tot, "halo_exchange_ms": h_ms, "transform_and_gather_ms": r_ms, "without_gather_ms": tot_ng,
                         "rounds": rounds, "audio_s_per_s": (L / SR) / (tot * 1e-3),
                         "layout": f"{rt.pieces} equal pieces, piece j on rank j % {world} in round j // {world} (block-cyclic)",
                         "halo": "ONE all_gather_into_tensor of every piece's edge samples (n_fft/2 + 3 hop per side) for all rounds; "
                                 "the inverse transform's 3 halo frames are recomputed by K1, not exchanged",
                         "gather": "per round one asynchronous all_gather_into_tensor of the equal-sized pieces straight into the "
                                   "result buffer, overlapped with the next round's kernels"}
    # the whole clip on every rank: the anchor's input, and the source of the peer path's pieces
    if world > 1:
        parts = []
        for c in range(rounds):
            sizes = [rt.shards[c * world + r].own1 - rt.shards[c * world + r].own0 for r in range(world)]
            parts.append(S.gather_concat(rt.owned_wav(c).contiguous(), sizes, world))
        full_wav = torch.cat(parts, dim=1)
        del parts
    else:
        full_wav = rt.owned_wav(0)
    # ---- (a') the same round trip with no collective on the data path: halos pulled out of the neighbours' peer-mapped
    # buffers, K2 storing its output into every GPU's result buffer (multicast store when the fabric has one)
    peer = None
    if world > 1 and os.environ.get("A2SB_BENCH_LONG_PEER", "1") != "0":
        try:
            # >= 4 GPUs: K2's fused gather is NVLink-ingress bound (every GPU receives 7/8 of the result), so K1 of the next of 4
            # rounds runs under it on the other half of the SMs (measured at 8 GPUs: 1.20 ms -> 1.17 ms; at 2 GPUs it loses)
            p_rounds, p_overlap = (4, True) if world >= 4 else (rounds, False)
            peer = S.PeerLongClipRoundTrip(L, N_FFT, HOP, rank, world, dev, rounds=p_rounds, overlap=p_overlap)
            for c in range(p_rounds):
                sh = peer.mine[c]
                peer.owned_wav(c).copy_(full_wav[:, sh.own0:sh.own1])
            peer.ready()

            def roundtrip_peer(ev):
                peer.pull_halos(); ev[1].record()
                peer.run(); ev[2].record()
            ph_ms, pr_ms, ptot = timed(roundtrip_peer, 2)
            peer.pull_halos(); peer.run()
            rec["round_trip_peer"] = {
                "ms": ptot, "halo_pull_ms": ph_ms, "transform_and_fused_gather_ms": pr_ms, "rounds": p_rounds,
                "k1_under_k2_on_half_the_sms": p_overlap,
                "audio_s_per_s": (L / SR) / (ptot * 1e-3), "store": "multimem.st (NVSwitch multicast)" if peer.multicast else "peer stores",
                "what": "symmetric memory (torch.distributed._symmetric_memory for allocation + handle exchange): one-sided peer copies "
                        "for the sample halos, K2 (a2sb_istft_inverse_mirrored) writes every output vector into the result buffer of "
                        "every GPU from inside the kernel, one symmetric-memory barrier at the end; no NCCL call on the data path"}
        except Exception as e:  # symmetric memory not available on this box / torch build
            rec["round_trip_peer"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
            peer = None
    # unsharded anchor + bit identity (below)
    if rank == 0:
        def unsharded(ev):
            sp = _lib.stft_forward(full_wav, N_FFT, N_FFT, HOP, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25, eps=1e-9)
            ev[1].record()
            unsharded.y = _lib.istft_inverse(sp, N_FFT, N_FFT, HOP, kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True,
                                             power=4.0, eps=1e-9)
            ev[2].record()
    else:
        def unsharded(ev):
            ev[1].record(); ev[2].record()
    u_f, u_i, u_tot = timed(unsharded, 2)
    if rank == 0:
        rec["round_trip"]["one_gpu_ms"] = u_tot
        rec["round_trip"]["speedup_vs_one_gpu"] = u_tot / tot
        rec["round_trip"]["speedup_without_gather"] = u_tot / tot_ng
        rec["round_trip"]["bit_identical_to_unsharded"] = bool(torch.equal(final[: rt.total_out], unsharded.y[0]))
        if peer is not None:
            rec["round_trip_peer"]["speedup_vs_one_gpu"] = u_tot / rec["round_trip_peer"]["ms"]
            rec["round_trip_peer"]["bit_identical_to_unsharded"] = bool(torch.equal(peer.final[: peer.total_out], unsharded.y[0]))
        del unsharded.y
    del rt, final, full_wav, peer
    torch.cuda.empty_cache()
    # ---- (b) blend step at config-3 size
    C_, H_, W_, WIN, BHOP = 3, N_FFT // 2, 310144, 256, 128
    sb = S.ShardedBlend(C_, H_, W_, WIN, BHOP, rank, world, dev)
    g = torch.Generator(device=dev).manual_seed(2000 + rank)
    sb.owned_x.copy_(torch.randn(sb.owned_x.shape, generator=g, device=dev))
    x0 = sb.owned_x.clone()
    sb.prime()
    ident = lambda a, t: a
    t_emb = torch.zeros(1, 4, device=dev)

    def blend_step(ev):
        sb.step(ident, t_emb, network_in_place=True)
        sb.swap()
        ev[1].record()
    (b_ms, _tot) = timed(blend_step, 1)
    same = torch.tensor([int(torch.equal(sb.owned_x, x0))], device=dev)     # identity network: blend(gather(x)) == x exactly
    if world > 1:
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
    n_seg = (W_ - (WIN - BHOP)) // BHOP
    alg = 4 * C_ * H_ * (2 * W_ + 2 * WIN * n_seg)                           # gather + blend bytes per step
    rec["blend_step"] = {"ms": b_ms, "segments": n_seg, "algorithmic_bytes": alg, "aggregate_gbs": alg / b_ms * 1e-6,
                         "identity_bit_exact": bool(int(same.item())),
                         "halo": "ONE grouped neighbour send/recv per step: 1 network-output segment (3 MB) each way; the 128 halo columns of the "
                                 "next state are blended locally from it"}
    sizes = [sh.col1 - sh.col0 for sh in sb.shards]
    full_x = S.gather_concat(x0, sizes, world) if world > 1 else x0
    if rank == 0:
        def unsharded_blend(ev):
            segs = _lib.segment_gather(full_x, WIN, BHOP)
            unsharded_blend.y = _lib.segment_blend(segs, 1, W_, WIN, BHOP)
            ev[1].record()
    else:
        def unsharded_blend(ev):
            ev[1].record()
    (ub_ms, _t) = timed(unsharded_blend, 1)
    if rank == 0:
        rec["blend_step"]["one_gpu_ms"] = ub_ms
        rec["blend_step"]["speedup_vs_one_gpu"] = ub_ms / b_ms
        rec["blend_step"]["unsharded_identity_bit_exact"] = bool(torch.equal(unsharded_blend.y, full_x))
        del unsharded_blend.y
    del sb, x0, full_x
    torch.cuda.empty_cache()
    return rec if rank == 0 else None


# --------------------------------------------------------------------------------------- GPU arm


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from audio_intelligence_b200 import _capi, _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the A2SB B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    B = args.clips
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    wav = (0.3 * torch.randn(B, CLIP_LEN, generator=g, device=dev)).clamp_(-1, 1)

    def k1(w):
        return _lib.stft_forward(w, N_FFT, N_FFT, HOP, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25, eps=1e-9)

    def k2(s):
        return _lib.istft_inverse(s, N_FFT, N_FFT, HOP, kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True,
                                  power=4.0, eps=1e-9)

    # warm-up with exactly the statement pattern of the timed loop, so the caching allocator already owns every
    # block the loop cycles through (a first-use cudaMalloc of the 2.7 GB spectrogram costs several ms)
    for _ in range(max(args.warmup, 3)):
        spec = k1(wav)
        out = k2(spec)
    barrier()
    K = args.steps
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K + 1)]
    n0 = _lib.launch_count()
    with ClockSampler(physical_gpu_index(local)) as clk:
        ev[0].record()
        for i in range(K):
            spec = k1(wav)
            ev[2 * i + 1].record()
            out = k2(spec)
            ev[2 * i + 2].record()
        clk.sample()
        barrier()
    launches = _lib.launch_count() - n0
    total_ms = ev[0].elapsed_time(ev[2 * K])
    k1_ms = [ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(K)]
    k2_ms = [ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(K)]
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / K
    audio_s_per_step = 10.0 * B * world
    value = audio_s_per_step / (ms_per_step * 1e-3)

    # ---- secondary (not the headline): the same round trip with the opt-in 32-byte row pitch of the spectrogram
    # (audio_transforms.transforms.set_row_alignment(8) / a2sb_fwd_args.out_pitch): identical values, rows padded
    # from T to a multiple of 8 frames, so K1 writes whole sectors only.
    aligned = None
    if rank == 0 and not args.skip_aligned:
        def k1a(w):
            return _lib.stft_forward(w, N_FFT, N_FFT, HOP, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25, eps=1e-9,
                                     row_align=8)
        for _ in range(3):          # same statement pattern as the timed loop (allocator steady state)
            spec_a = k1a(wav)
            out_a = k2(spec_a)
        torch.cuda.synchronize()
        Ka = max(3, min(K, 20))
        eva = [torch.cuda.Event(enable_timing=True) for _ in range(2 * Ka + 1)]
        eva[0].record()
        for i in range(Ka):
            spec_a = k1a(wav)
            eva[2 * i + 1].record()
            out_a = k2(spec_a)
            eva[2 * i + 2].record()
        torch.cuda.synchronize()
        a1 = statistics.median(eva[2 * i].elapsed_time(eva[2 * i + 1]) for i in range(Ka))
        a2 = statistics.median(eva[2 * i + 1].elapsed_time(eva[2 * i + 2]) for i in range(Ka))
        aligned = {"row_pitch_frames": int(spec_a.stride(2)), "stft_fwd_kernel_ms": a1, "istft_inv_kernel_ms": a2,
                   "bit_identical_to_contiguous": bool(torch.equal(out_a, out) and torch.equal(spec_a, spec)),
                   "note": "opt-in layout, not the headline: spectrogram rows padded to a multiple of 8 frames"}
        del spec_a, out_a

    # ---- e2e: host buffers through the C ABI (a2sb_roundtrip_host), copies inside the timed region
    e2e = None
    if not args.skip_e2e:
        h_in = torch.empty(B, CLIP_LEN, dtype=torch.float32).pin_memory()
        h_in.copy_(wav)
        out_len = HOP * (CLIP_LEN // HOP)
        h_out = torch.empty(B, out_len, dtype=torch.float32).pin_memory()
        for _ in range(2):
            _lib.roundtrip_host(h_in, h_out, N_FFT, HOP)
        Ke = max(3, min(K, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            _lib.roundtrip_host(h_in, h_out, N_FFT, HOP)          # returns after the D2H has completed
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / Ke], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": audio_s_per_step / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(h_in.numel() * 4 * world), "d2h_bytes_per_step": int(h_out.numel() * 4 * world),
               "ms_per_step": float(dt.item()) * 1e3, "steps": Ke,
               "api": "a2sb_roundtrip_host (C ABI): pinned host wav -> H2D -> K1 -> K2 -> D2H -> pinned host wav; "
                      "spectrogram stays in HBM (it feeds the on-device network)",
               "numa_node_rank0": numa_node}
        same = torch.equal(h_out[:2], out[:2].cpu())
        e2e["matches_device_path"] = bool(same)
        # PCIe floor of the same step on the same ranks at the same time: the step's H2D and D2H bytes as two single
        # concurrent pinned copies per rank (no kernels).  e2e_over_floor = how far the pipelined call is from it.
        d_in = torch.empty_like(h_in, device=dev)
        d_out = torch.empty_like(h_out, device=dev)
        s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        best = float("inf")
        for it in range(5):
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(s_up):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_dn):
                h_out.copy_(d_out, non_blocking=True)
            s_up.synchronize(); s_dn.synchronize()
            barrier()
            dtp = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(dtp, op=dist.ReduceOp.MAX)
            if it >= 1:
                best = min(best, float(dtp.item()))
        e2e["pcie_probe"] = {"floor_ms": best * 1e3, "h2d_gbs_per_gpu": h_in.numel() * 4 / best * 1e-9,
                             "d2h_gbs_per_gpu": h_out.numel() * 4 / best * 1e-9,
                             "aggregate_gbs_per_direction": h_in.numel() * 4 * world / best * 1e-9,
                             "what": "per rank one pinned H2D of the step's input and one pinned D2H of its output, concurrently, "
                                     "all ranks at the same time; max over ranks, best of 4"}
        e2e["e2e_over_floor"] = e2e["ms_per_step"] / (best * 1e3)
        # The same call with the wav FILE content on both sides (16-bit PCM in, PCM_16 out: SURVEY 8f rank 4, the wav edges):
        # decode fused into K1's load, encode into K2's store, half the PCIe bytes each way.  Reported beside `e2e`, whose
        # float32 buffers are what the reference's transform API takes; the reference arm's tensors are float32 as well.
        try:
            p_in = torch.empty(B, CLIP_LEN, dtype=torch.int16).pin_memory()
            p_in.copy_((wav * 32767.0).round().clamp_(-32768, 32767).to(torch.int16))
            p_out = torch.empty(B, out_len, dtype=torch.int16).pin_memory()
            for _ in range(2):
                _lib.roundtrip_host(p_in, p_out, N_FFT, HOP)
            barrier()
            t0 = time.perf_counter()
            for _ in range(Ke):
                _lib.roundtrip_host(p_in, p_out, N_FFT, HOP)
            barrier()
            dtq = torch.tensor([(time.perf_counter() - t0) / Ke], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(dtq, op=dist.ReduceOp.MAX)
            # check: decode -> float kernels -> libsndfile rule on the host == the fused PCM path (first two clips)
            ref = _lib.istft_inverse(_lib.stft_forward((p_in[:2].to(dev).float() / 32768.0), N_FFT, N_FFT, HOP, kind=_capi.KIND_MAGPHASE,
                                                       drop_dc=True, power=0.25),
                                     N_FFT, N_FFT, HOP, kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True, power=4.0)
            sc = (ref * 2147483648.0)
            q = torch.where(sc >= 2147483647.0, torch.full_like(sc, 32767.0),
                            torch.where(sc <= -2147483648.0, torch.full_like(sc, -32768.0), torch.floor(torch.round(sc.double()) / 65536.0).float()))
            e2e["pcm16"] = {"value": audio_s_per_step / float(dtq.item()), "unit": UNIT, "ms_per_step": float(dtq.item()) * 1e3,
                            "h2d_bytes_per_step": int(p_in.numel() * 2 * world), "d2h_bytes_per_step": int(p_out.numel() * 2 * world),
                            "matches_decode_transform_encode": bool(torch.equal(q.to(torch.int16).cpu(), p_out[:2])),
                            "api": "a2sb_roundtrip_host_pcm16 (C ABI): pinned int16 PCM -> H2D -> K1 (decode fused) -> K2 (PCM_16 encode "
                                   "fused, libsndfile rule) -> D2H -> pinned int16 PCM"}
            del p_in, p_out
        except Exception as e:
            e2e["pcm16"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        del h_in, h_out, d_in, d_out

    seg_rec = None
    if rank == 0 and world == 1 and not args.skip_long:
        # streaming kernels of the blend / mask path at config-3 size (tools/bench_segments.py), CUDA-event medians
        del spec, out
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_segments
        seg_rec = {k: {"ms": v["ms"], "algorithmic_bytes": v["bytes"], "gbs": v["gbs"], "frac_of_measured_peak": None}
                   for k, v in bench_segments.measure_kernels(dev).items()}
        torch.cuda.empty_cache()
        spec = out = None
    plugin_rec = None
    if rank == 0 and world == 1 and not args.skip_e2e:
        plugin_rec = e2e_plugin_leg(dev, B)
    long_rec = None
    if not args.skip_long:
        spec = out = None
        torch.cuda.empty_cache()
        long_rec = long_audio_leg(rank, world, local, dev)

    if rank == 0:
        fwd_b, inv_b = algorithmic_bytes(B)
        peaks, peak_src = None, "fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = float(json.load(fh)["hbm_gbs"])
                peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            peaks = 6650.0
            peak_src = "fallback (B200_PROFILING.md, 6.65 TB/s)"
        kern = {"stft_fwd_kernel": (fwd_b, statistics.mean(k1_ms)), "istft_inv_kernel": (inv_b, statistics.mean(k2_ms))}
        traffic = {}
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh)
        except Exception:
            pass
        dom = max(kern, key=lambda k: kern[k][1])
        per = {k: {"algorithmic_bytes": b, "ms": ms, "gbs": b / ms * 1e-6, "frac": b / ms * 1e-6 / peaks,
                   "traffic": traffic.get(k)} for k, (b, ms) in kern.items()}
        for k, v in (("stft_fwd_kernel", k1_ms), ("istft_inv_kernel", k2_ms)):
            per[k]["ms_min_median_max"] = [min(v), statistics.median(v), max(v)]
        roof = {"bound": "hbm", "kernel": dom, "achieved": per[dom]["gbs"], "peak": peaks, "unit": "GB/s",
                "frac": per[dom]["frac"], "traffic": per[dom]["traffic"], "peak_source": peak_src,
                "frac_of_nominal_8TBs": per[dom]["gbs"] / 8000.0,
                "round_trip_gbs": (fwd_b + inv_b) / ms_per_step * 1e-6, "kernels": per}
        if aligned is not None:
            aligned["stft_fwd_kernel_gbs"] = fwd_b / aligned["stft_fwd_kernel_ms"] * 1e-6
            aligned["istft_inv_kernel_gbs"] = inv_b / aligned["istft_inv_kernel_ms"] * 1e-6
            roof["aligned_rows_variant"] = aligned
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(B, world), "clocks": clk.summary(),
                "gpu_launches": int(launches), "roofline": roof}
        if e2e is not None:
            line["e2e"] = e2e
        if seg_rec is not None:
            for v in seg_rec.values():
                v["frac_of_measured_peak"] = v["gbs"] / peaks
            line["segment_kernels"] = {"size": "config 3: [1, 3, 1024, 310144] frames, 2422 segments of 256 hopped by 128",
                                       "kernels": seg_rec}
        if plugin_rec is not None:
            line["e2e_plugin"] = plugin_rec
        if long_rec is not None:
            line["long_audio"] = long_rec
        if world == 1 and not args.skip_cpu:
            def gpu_out(w_cpu):
                return k2(k1(w_cpu.to(dev))).cpu().numpy()
            line["cpu_baseline"] = cpu_baseline_leg(gpu_out)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=256, help="clips per GPU (BASELINE config 2: 256)")
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg (profiling runs)")
    ap.add_argument("--skip-e2e", action="store_true", help="omit the host-buffer e2e leg (profiling runs)")
    ap.add_argument("--skip-aligned", action="store_true", help="omit the secondary row-pitched measurement")
    ap.add_argument("--skip-long", action="store_true", help="omit the sharded 1 h clip leg (BASELINE configs[2])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"), *sys.argv]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
