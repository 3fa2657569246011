"""BASELINE configs[4]: the A2SB inference loop's transform I/O at batch 64 per GPU with a random-init network stub,
data-parallel over the GPUs of one box (torchrun --nproc-per-node N tools/bench_config5.py, or plain python for N = 1).

Per rank and batch: forward chain (K1) -> bandwidth-extension mask + noise fill -> ddpm_sample (n_steps bridge steps:
segment gather K3, network stub per torch.chunk mini-batch, fused blend + sampler step K4s) -> inverse chain (K2) on the
last prediction.  Two layouts are timed: "contiguous" (the reference's tensors) and "padded" (transforms.
set_segment_padding(256, 128): K1 emits the wrap-padded, hop-aligned width, the mask fill keeps it, the sampler starts
without its two wrap-pad passes, K2 reads the padded prediction in place); and two network stubs: a random-init 3->3
channel 3x3 convolution (cuDNN -- library code, the stand-in for the UNet) and `identity` (what is left is this package's
own work: the "transform I/O" of the loop).  Device-timed, max over ranks; one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_intelligence_b200 import _lib, diffusion as D  # noqa: E402
from audio_intelligence_b200.audio_transforms import transforms as T  # noqa: E402
from audio_intelligence_b200.corruption import corruptions as CO  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, n_fft, hop, sr = int(os.environ.get("A2SB_C5_BATCH", 64)), 2048, 512, 44100
    n_steps = int(os.environ.get("A2SB_C5_STEPS", 50))
    reps = int(os.environ.get("A2SB_C5_REPS", 3))
    g = torch.Generator(device=dev).manual_seed(3000 + rank)
    wav = (0.3 * torch.randn(B, 10 * sr, generator=g, device=dev)).clamp_(-1, 1)
    fwd = [T.ComplexSpectrogram(n_fft, n_fft, hop), T.ComplexToMagInstPhase(), T.SpectrogramDropDCTerm(), T.PowerScaleSpectrogram(0.25, [0])]
    inv = [T.PowerScaleSpectrogram(4, [0]), T.SpectrogramAddDCTerm(), T.SVDFixMagInstPhase(), T.MagInstPhaseToComplex(),
           T.InverseComplexSpectrogram(n_fft, n_fft, hop)]
    conv = torch.nn.Conv2d(3, 3, 3, padding=1).to(dev).requires_grad_(False)
    torch.nn.init.normal_(conv.weight, std=0.05, generator=torch.Generator(device=dev).manual_seed(1))
    nets = {"conv3x3_stub": lambda x, te: conv(x) + te[:, :1, None, None], "identity": lambda x, te: x}
    t_to_emb = lambda t: torch.stack([t, t * t], dim=1).to(dev)
    ts = torch.linspace(1.0, 0.05, n_steps)[None]                     # A2SB_lightning_module.py:193
    ddpm = D.Diffusion()
    first_row = 185                                                   # 4 kHz cut-off at n_fft 2048 / 44.1 kHz

    def sync():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def loop(net):
        x0, _ = T.apply_audio_transforms(wav, fwd)
        x1, mask = CO._fill(x0, (first_row, x0.shape[-2]), (0, x0.shape[-1]), 0.5)
        preds = D.ddpm_sample(net, ddpm, x1, ts, t_to_emb, mask=mask, win_length=256, hop_length=128, batch_size=16,
                              use_ot_ode=True, history="last")
        y, _ = T.apply_audio_transforms(preds[-1], inv)
        return y

    res = {"config": f"batch {B} x 10 s per GPU, n_fft 2048 / hop 512, {n_steps - 1} sampler steps, windows 256/128, chunks of 16, "
                     f"{world} GPU(s) data-parallel", "world": world}
    outs = {}
    for layout in ("contiguous", "padded"):
        T.set_segment_padding(256, 128) if layout == "padded" else T.set_segment_padding(None)
        for name, net in nets.items():
            best, launches = None, 0
            for it in range(reps + 1):
                torch.manual_seed(7)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n0 = _lib.launch_count()
                sync()
                e0.record()
                y = loop(net)
                e1.record()
                torch.cuda.synchronize()
                launches = _lib.launch_count() - n0
                ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                if it >= 1:
                    best = float(ms) if best is None else min(best, float(ms))
            outs[(layout, name)] = y
            res[f"{layout}/{name}"] = {"ms_per_batch": best, "audio_s_per_s_per_gpu": 10.0 * B / (best * 1e-3),
                                       "audio_s_per_s_total": 10.0 * B * world / (best * 1e-3), "launches_of_this_library": launches}
    T.set_segment_padding(None)
    res["padded_bit_identical_to_contiguous"] = bool(all(torch.equal(outs[("padded", n)], outs[("contiguous", n)]) for n in nets))
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
