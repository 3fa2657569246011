"""BASELINE config 1: one 10 s mono clip, forward + inverse chain through the reference-facing Python API
(apply_audio_transforms), device-resident and from/to CPU tensors.  Wall-clock per round trip, median of 200.
Usage: python tools/bench_config1.py  (needs a GPU)."""
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_intelligence_b200.audio_transforms import transforms as T  # noqa: E402


def main():
    n_fft, hop = 2048, 512
    fwd = [T.ComplexSpectrogram(n_fft, n_fft, hop), T.ComplexToMagInstPhase(), T.SpectrogramDropDCTerm(),
           T.PowerScaleSpectrogram(0.25, [0])]
    inv = [T.PowerScaleSpectrogram(4, [0]), T.SpectrogramAddDCTerm(), T.SVDFixMagInstPhase(),
           T.MagInstPhaseToComplex(), T.InverseComplexSpectrogram(n_fft, n_fft, hop)]
    wav_cpu = (0.3 * torch.randn(441000)).clamp_(-1, 1)
    wav = wav_cpu.cuda()
    res = {}
    for name, x in (("device_resident", wav), ("cpu_tensors", wav_cpu)):
        ts = []
        for i in range(220):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            spec, _ = T.apply_audio_transforms(x, fwd)
            y, _ = T.apply_audio_transforms(spec, inv)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        ts = ts[20:]
        res[name] = {"median_us": statistics.median(ts) * 1e6, "min_us": min(ts) * 1e6,
                     "audio_s_per_s": 10.0 / statistics.median(ts)}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    k1, k2 = [], []
    for i in range(50):
        ev[0].record()
        spec, _ = T.apply_audio_transforms(wav, fwd)
        ev[1].record()
        y, _ = T.apply_audio_transforms(spec, inv)
        ev[2].record()
        torch.cuda.synchronize()
        k1.append(ev[0].elapsed_time(ev[1]) * 1e3)
        k2.append(ev[1].elapsed_time(ev[2]) * 1e3)
    res["device_time_us"] = {"forward": statistics.median(k1), "inverse": statistics.median(k2)}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
