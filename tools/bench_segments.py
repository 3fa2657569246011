"""Streaming kernels of the blend path at BASELINE config 3 size (1 h of audio: [1, 3, 1024, 310144] frames, 2422
segments of 256 @ hop 128): achieved HBM GB/s on algorithmic bytes.  Also a pinned-memory PCIe
probe for the e2e figure.  Usage: python tools/bench_segments.py  (needs a GPU)."""
import ctypes as C
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_intelligence_b200 import _capi, _lib, diffusion as D  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]


def measure_kernels(dev, reps=10):
    """CUDA-event medians of the streaming kernels at config-3 size -> {kernel: {ms, bytes (algorithmic), gbs}}."""
    W0, win, hop = 310079, 256, 128
    x = torch.randn(1, 3, 1024, W0, device=dev)
    res = {}
    xp = D.multidiffusion_pad_inputs(x, win, hop)
    W = xp.shape[-1]
    n_seg = (W - (win - hop)) // hop
    plane = 4 * 3 * 1024
    res["wrap_pad"] = {"ms": timed(lambda: D.multidiffusion_pad_inputs(x, win, hop), reps), "bytes": plane * (W0 + W)}
    segs = _lib.segment_gather(xp, win, hop)
    res["segment_gather"] = {"ms": timed(lambda: _lib.segment_gather(xp, win, hop)), "bytes": plane * (W + win * n_seg)}
    res["segment_blend"] = {"ms": timed(lambda: _lib.segment_blend(segs, 1, W, win, hop)), "bytes": plane * (win * n_seg + W)}
    mask = torch.zeros_like(xp)
    mask[..., 1000:2000] = 1
    pred, nxt, x1 = torch.empty_like(xp), torch.empty_like(xp), torch.randn_like(xp)
    L = _lib.lib()

    def step():
        a = _capi.StepArgs(xp.data_ptr(), x1.data_ptr(), mask.data_ptr(), None, None, pred.data_ptr(), nxt.data_ptr(),
                           0.1, 0.4, 0.6, 0.0, 0.0, 1)
        _capi.check(L, L.a2sb_segment_blend_step(segs.data_ptr(), C.byref(a), 1, 3 * 1024, W, win, hop, _lib.stream_ptr()))
    res["segment_blend_step"] = {"ms": timed(step), "bytes": plane * (win * n_seg + 3 * W + 2 * W)}
    noise = torch.randn_like(xp)
    res["mask_fill"] = {"ms": timed(lambda: _lib.mask_fill(xp, noise, (185, 1024), (0, W), 0.5)), "bytes": plane * 4 * W}
    # fused padding variant: row-pitched input -> filled tensor + mask at the padded width (mask_fill_padded_kernel)
    noise0 = noise[..., :W0].contiguous()
    res["mask_fill_padded"] = {"ms": timed(lambda: _lib.mask_fill_padded(xp[..., :W0], noise0, (185, 1024), (0, W0), 0.5, win, hop), reps),
                               "bytes": plane * (2 * W0 + 2 * W)}
    for k, v in res.items():
        v["gbs"] = v["bytes"] / v["ms"] * 1e-6
    return res


def main():
    dev = torch.device("cuda")
    res = measure_kernels(dev)
    # PCIe probe (pinned): what bounds bench.py's e2e figure
    n = 451584000
    h = torch.empty(n // 4, dtype=torch.float32).pin_memory()
    d = torch.empty(n // 4, dtype=torch.float32, device=dev)
    d2 = torch.empty_like(d)
    h2 = torch.empty_like(h).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def wall(fn, reps=5):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    t_h2d = wall(lambda: d.copy_(h, non_blocking=True))
    t_d2h = wall(lambda: h2.copy_(d2, non_blocking=True))

    def both():
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    t_both = wall(both)
    res["pcie"] = {"bytes_each_way": n, "h2d_gbs": n / t_h2d * 1e-9, "d2h_gbs": n / t_d2h * 1e-9,
                   "concurrent_ms": t_both * 1e3, "concurrent_gbs_each_way": n / t_both * 1e-9}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
