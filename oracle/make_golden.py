"""Generate tests/golden/*.npz|json by running the UNMODIFIED reference on seeded inputs.

Run in the build container, where the reference is mounted read-only at /root/reference:

    python oracle/make_golden.py            # rewrites tests/golden/

The reference modules (A2SB/audio_transforms/transforms.py, A2SB/diffusion.py, A2SB/utils.py,
A2SB/corruption/corruptions.py) are imported as they are; the only shim is a stub `jsonargparse`
module, which transforms.py imports at module top (transforms.py:16,22) but only uses for
isinstance checks.  The fixtures pin the oracle (tests/test_oracle.py) and, on the GPU box where
/root/reference does not exist, the CUDA path (tests/test_parity_gpu.py).
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("A2SB_REFERENCE", "/root/reference/A2SB")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    ja = types.ModuleType("jsonargparse")
    ns = types.ModuleType("jsonargparse._namespace")

    class Namespace:  # only used for isinstance/type checks (transforms.py:27,67)
        pass

    ja.Namespace = ns.Namespace = Namespace
    ja._namespace = ns
    sys.modules.setdefault("jsonargparse", ja)
    sys.modules.setdefault("jsonargparse._namespace", ns)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import audio_transforms.transforms as T  # noqa: E402
    import diffusion as D  # noqa: E402
    import utils as U  # noqa: E402
    from corruption import corruptions as C  # noqa: E402
    return T, D, U, C


def chains(T, n_fft, hop):
    fwd = [T.ComplexSpectrogram(n_fft, n_fft, hop), T.ComplexToMagInstPhase(), T.SpectrogramDropDCTerm(),
           T.PowerScaleSpectrogram(0.25, [0])]          # configs/ensemble_2split_sampling.yaml:105-119
    inv = [T.PowerScaleSpectrogram(4, [0]), T.SpectrogramAddDCTerm(), T.SVDFixMagInstPhase(),
           T.MagInstPhaseToComplex(), T.InverseComplexSpectrogram(n_fft, n_fft, hop)]   # :63-78
    inv_nosvd = [t for t in inv if not isinstance(t, T.SVDFixMagInstPhase)]
    return fwd, inv, inv_nosvd


def main():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import a2sb_oracle as O

    T, D, U, C = import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    meta = {"torch": torch.__version__, "reference": REF}

    # ---- transform chains, all four n_fft of BASELINE config 4 -----------------------------------
    for n_fft in (512, 1024, 2048, 4096):
        hop = n_fft // 4
        L = 5 * n_fft + 37 * 4 + 3          # not a multiple of hop: exercises T = 1 + L // hop
        wav = O.synth_noise(L, 1000 + n_fft)
        wav[: n_fft // 8] *= 0.0            # a silent stretch: exercises mag -> 0 handling
        fwd, inv, inv_nosvd = chains(T, n_fft, hop)
        w = torch.from_numpy(wav)
        cplx = fwd[0](w)                                            # [2, F, T] view
        spec, _ = T.apply_audio_transforms(w, fwd)                  # [3, n_fft/2, T]
        wav_inv, _ = T.apply_audio_transforms(spec, inv)
        wav_inv_nosvd, _ = T.apply_audio_transforms(spec, inv_nosvd)
        g = torch.Generator().manual_seed(7 + n_fft)
        pert = spec + 0.05 * torch.randn(spec.shape, generator=g)   # network-like: phases off the circle
        wav_pert, _ = T.apply_audio_transforms(pert, inv)
        wav_cplx = T.InverseComplexSpectrogram(n_fft, n_fft, hop)(cplx)
        np.savez_compressed(
            os.path.join(OUT, f"chain_n{n_fft}.npz"),
            wav=wav, complex_spec=cplx.contiguous().numpy(), spec=spec.contiguous().numpy(),
            wav_inv=wav_inv.numpy(), wav_inv_nosvd=wav_inv_nosvd.numpy(), spec_pert=pert.numpy(),
            wav_pert=wav_pert.numpy(), wav_cplx=wav_cplx.numpy())
        meta[f"chain_n{n_fft}"] = {"L": L, "hop": hop, "T": int(spec.shape[-1]), "out_len": int(wav_inv.shape[0])}

    # ---- tonal clip at the shipped parameters (sparse spectrum, many near-zero bins) ---------------
    n_fft, hop = 2048, 512
    wav = O.synth_tonal(6 * 2048)
    fwd, inv, _ = chains(T, n_fft, hop)
    spec, _ = T.apply_audio_transforms(torch.from_numpy(wav), fwd)
    wav_inv, _ = T.apply_audio_transforms(spec, inv)
    np.savez_compressed(os.path.join(OUT, "chain_tonal.npz"), wav=wav, spec=spec.contiguous().numpy(),
                        wav_inv=wav_inv.numpy())

    # ---- standalone ops ---------------------------------------------------------------------------
    g = torch.Generator().manual_seed(11)
    msp = torch.randn([3, 40, 9], generator=g)
    msp[1, 0, 0] = 0.0; msp[2, 0, 0] = 0.0                # degenerate pair -> (1, 0)
    msp[1, 0, 1] = -3.0; msp[2, 0, 1] = 0.0               # -> (-1, 0)
    np.savez_compressed(
        os.path.join(OUT, "ops.npz"), msp=msp.numpy(),
        svd_fix=T.SVDFixMagInstPhase()(msp).numpy(),
        to_complex=T.MagInstPhaseToComplex()(msp).numpy(),
        to_magphase=T.ComplexToMagInstPhase()(msp[:2]).numpy(),
        pow_half_all=T.PowerScaleSpectrogram(0.5)(msp).numpy(),
        pow_quarter_c0=T.PowerScaleSpectrogram(0.25, [0])(msp).numpy(),
        pow_four_c0=T.PowerScaleSpectrogram(4, [0])(msp).numpy(),
        add_dc=T.SpectrogramAddDCTerm()(msp).numpy(),
        drop_dc=T.SpectrogramDropDCTerm()(msp).contiguous().numpy())

    # ---- segment windowing / blend ----------------------------------------------------------------
    g = torch.Generator().manual_seed(5)
    x = torch.randn([2, 3, 8, 300], generator=g)
    win, hop_s = 64, 32
    xp = D.multidiffusion_pad_inputs(x, win, hop_s)
    xp_const = D.multidiffusion_pad_inputs(x, win, hop_s, padding_constant=0)
    t_emb = torch.zeros([2, 4])
    ident = D.get_multidiffusion_vf(lambda a, t: a, xp, t_emb, win, hop_s, batch_size=5)
    affine = D.get_multidiffusion_vf(lambda a, t: a * 2 + 0.1, xp, t_emb, win, hop_s, batch_size=5)
    # segment-index-dependent "network": catches ordering mistakes in the (b l) axis
    def ramp(a, t):
        return a + torch.arange(a.shape[0], dtype=a.dtype).view(-1, 1, 1, 1) * 0.001
    ramp_out = D.get_multidiffusion_vf(ramp, xp, t_emb, win, hop_s, batch_size=1000)
    g3 = torch.Generator().manual_seed(6)
    x3 = torch.randn([1, 3, 4, 130], generator=g3)       # 3-way overlap: win 48, hop 16
    xp3 = D.multidiffusion_pad_inputs(x3, 48, 16)
    noisy = D.get_multidiffusion_vf(lambda a, t: a * 1.7 - 0.3, xp3, torch.zeros([1, 4]), 48, 16, batch_size=3)
    np.savez_compressed(os.path.join(OUT, "blend.npz"), x=x.numpy(), xp=xp.numpy(), xp_const=xp_const.numpy(),
                        ident=ident.numpy(), affine=affine.numpy(), ramp=ramp_out.numpy(), x3=x3.numpy(),
                        xp3=xp3.numpy(), noisy3=noisy.numpy())

    # ---- integer known answers --------------------------------------------------------------------
    ka = {"pad_widths": {}, "frames": {}, "short_input": {}}
    for w_in in (862, 257, 385, 256, 100, 128, 129, 310079):
        ka["pad_widths"][str(w_in)] = int(D.multidiffusion_pad_inputs(torch.zeros([1, 1, 1, w_in]), 256, 128).shape[-1])
    for n_fft in (512, 1024, 2048, 4096):
        hop = n_fft // 4
        s = T.ComplexSpectrogram(n_fft, n_fft, hop)(torch.zeros(441000))
        y = T.InverseComplexSpectrogram(n_fft, n_fft, hop)(s)
        ka["frames"][str(n_fft)] = [int(s.shape[-1]), int(y.shape[0])]
    mask_row = torch.ones(896)
    for a, b in ((86, 103), (318, 344), (800, 896)):
        mask_row[a:b] = 0
    mids = [int(v) for v in U.find_middle_of_zero_segments(mask_row)]
    ka["zero_segment_centres"] = mids
    ka["upsample_first_row"] = {}
    for n_fft in (512, 1024, 2048, 4096):
        # min == max cutoff makes the reference's randint deterministic (corruptions.py:38-46)
        m = C.UpsampleMask.get_upsample_mask(torch.zeros([3, n_fft // 2, 4]), 4000, 4000, 44100)
        ka["upsample_first_row"][str(n_fft)] = int(torch.nonzero(m[0, :, 0])[0, 0])
    seg = C.TimestampedSegmentInpaintMaskTransform if hasattr(C, "TimestampedSegmentInpaintMaskTransform") else None
    ka["inpaint_frames_1.0_1.2"] = [int(44100 / 512 * 1.0), int(44100 / 512 * 1.2)]
    wins = []
    for c in mids:                       # A2SB_lightning_module.py:161-174
        l, r = int(c - 256 / 2), int(c + 256 / 2)
        if l < 0:
            r -= l; l = 0
        if r > 896:
            l -= (r - 896); r = 896
        wins.append([l, r])
    ka["inpaint_windows"] = wins
    ka["envelope"] = {}
    w = torch.hann_window(2048)
    env = torch.zeros(2048 + 512 * 20)
    for t in range(21):
        env[t * 512:t * 512 + 2048] += w * w
    ka["envelope"] = {"first_kept": float(env[1024]), "interior": float(env[4096])}
    with open(os.path.join(OUT, "known_answers.json"), "w") as fh:
        json.dump({"meta": meta, **ka}, fh, indent=1, sort_keys=True)
    print("wrote", sorted(os.listdir(OUT)))


def mask_cases():
    """(name, constructor kwargs) of the corruption transforms the mask fixtures cover."""
    return [
        ("multinomial", dict(p_upsample_mask=0.4, p_extension_mask=0.3, p_inpaint_mask=0.3, fill_noise_level=0.5,
                             sampling_rate=44100, upsample_mask_kwargs=dict(min_cutoff_freq=2000, max_cutoff_freq=8000),
                             inpainting_mask_kwargs=dict(min_inpainting_frac=0.05, max_inpainting_frac=0.4, is_random=True))),
        ("timestamped", dict(start_time=0.1, end_time=0.35, hop_length=512, sampling_rate=44100, fill_noise_level=0.5)),
    ]


def make_masks():
    """tests/golden/masks.npz: masks and noise-filled spectrograms of the UNMODIFIED reference corruption
    transforms (corruption/corruptions.py) for seeded runs on CPU tensors, plus zero-segment centres
    (utils.py:54-81) of random rows."""
    T, D, U, C = import_reference()
    out = {}
    spec = (torch.arange(3 * 64 * 80, dtype=torch.float32).reshape(3, 64, 80) % 17 - 8) / 4
    out["spec"] = spec.numpy()
    for name, kw in mask_cases():
        cls = C.MultinomialInpaintMaskTransform if name == "multinomial" else C.TimestampedSegmentInpaintMaskTransform
        for seed in range(6):
            torch.manual_seed(seed)
            np.random.seed(seed)
            filled, mask = cls(**kw)(spec.clone())
            out[f"{name}_{seed}_mask"] = mask.numpy().astype(np.uint8)
            out[f"{name}_{seed}_filled"] = filled.numpy()
    for seed in range(6):
        torch.manual_seed(100 + seed)
        m = C.UpsampleMask.get_upsample_mask(torch.zeros(3, 128, 5), 1000, 9000, 44100)
        out[f"upsample_{seed}"] = m.numpy().astype(np.uint8)
        m = C.ExtensionMask.get_extension_mask(torch.zeros(3, 4, 200), 32)
        out[f"extension_{seed}"] = m.numpy().astype(np.uint8)
        np.random.seed(100 + seed)
        m = C.InpaintMask.get_inpainting_mask(torch.zeros(3, 4, 200), 0.1, 0.5, seed % 2 == 0)
        out[f"inpaint_{seed}"] = m.numpy().astype(np.uint8)
        g = torch.Generator().manual_seed(200 + seed)
        row = (torch.rand(int(torch.randint(1, 3000, [1], generator=g)), generator=g) < [0.05, 0.5, 0.95][seed % 3]).float()
        out[f"zero_row_{seed}"] = row.numpy().astype(np.uint8)
        out[f"zero_mid_{seed}"] = U.find_middle_of_zero_segments(row).numpy().astype(np.int32)
    torch.manual_seed(0)
    out["rng_probe"] = torch.randn(64).numpy()      # lets a test tell whether this host's CPU generator matches
    np.savez_compressed(os.path.join(OUT, "masks.npz"), **out)
    print("wrote masks.npz with", len(out), "arrays")


def sampler_setup():
    """Inputs shared by the generator and the tests: deterministic stub network, time embedding, schedule."""
    g = torch.Generator().manual_seed(77)
    x_1 = torch.randn(2, 3, 4, 150, generator=g)
    mask = torch.zeros(2, 3, 4, 150)
    mask[..., 50:75] = 1
    mask[..., 144:] = 1
    t_steps = torch.linspace(1.0, 0.0, 5)[None]

    def t_to_emb(t):
        return torch.stack([t, t * t], dim=1)

    def net(x, t_emb):
        pos = torch.arange(x.shape[-1], dtype=x.dtype, device=x.device)
        return x * 0.8 + t_emb[:, :1, None, None].to(x.device) * 0.1 + pos * 0.01

    return x_1, mask, t_steps, t_to_emb, net


def reference_ddpm_sample(D, ddpm, net, t_to_emb, x_1, t_steps, mask, mask_pred_x0, win_length, hop_length, batch_size,
                          use_ot_ode=True):
    """The loop of A2SBModel.ddpm_sample (A2SB_lightning_module.py:103-146) restated around the reference's own
    diffusion.py functions (the Lightning module itself cannot be imported here: no `lightning`)."""
    n_steps = t_steps.shape[1] - 1
    original_width = x_1.shape[-1]
    x_1 = D.multidiffusion_pad_inputs(x_1, win_length, hop_length)
    mask = D.multidiffusion_pad_inputs(mask, win_length, hop_length)
    x_t = x_1.clone()
    outs, states = [], []
    for t_idx in range(n_steps):
        t_emb = t_to_emb(t_steps[:, t_idx]).repeat(x_1.shape[0], 1)
        t, t_prev = t_steps[:, t_idx], t_steps[:, t_idx + 1]
        vf = D.get_multidiffusion_vf(net, x_t, t_emb, win_length=win_length, hop_length=hop_length, batch_size=batch_size)
        pred_x0 = ddpm.get_pred_x0(t_steps[:, t_idx], x_t, vf)
        if mask is not None and mask_pred_x0:
            pred_x0 = pred_x0 * mask + (1 - mask) * x_1
        outs.append(pred_x0.cpu())
        x_t = ddpm.p_posterior(t_prev, t, x_t, pred_x0, ot_ode=use_ot_ode)
        if mask is not None:
            xt_true = x_1
            if not use_ot_ode:
                xt_true = xt_true + ddpm.get_std_t(t_prev) * torch.randn_like(xt_true)
            x_t = (1. - mask) * xt_true + mask * x_t
        states.append(x_t.clone())
    return [D.multidiffusion_unpad_outputs(p, original_width) for p in outs], states


def make_sampler():
    """tests/golden/sampler.npz: schedule scalars of the reference's Diffusion and a 4-step ot-ode sampling run
    (stub network, multidiffusion windows 64/32) -- pins K4s and the host-side schedule."""
    T, D, U, C = import_reference()
    ddpm = D.Diffusion()
    out = {}
    ts = torch.tensor([0.0, 1e-3, 0.1, 0.25, 0.4999, 0.5, 0.5001, 0.75, 0.9, 0.999, 1.0])
    out["t"] = ts.numpy()
    out["int_beta"] = ddpm.get_int_beta_0_t(ts).numpy()
    out["std_fwd"] = ddpm.get_std_fwd(ts).numpy()
    out["std_rev"] = ddpm.get_std_rev(ts).numpy()
    out["std_t"] = ddpm.get_std_t(ts).numpy()
    x_1, mask, t_steps, t_to_emb, net = sampler_setup()
    out["x_1"], out["mask"], out["t_steps"] = x_1.numpy(), mask.numpy().astype(np.uint8), t_steps.numpy()
    for tag, mp in (("mp1", True), ("mp0", False)):
        preds, states = reference_ddpm_sample(D, ddpm, net, t_to_emb, x_1, t_steps, mask, mp, 64, 32, 4)
        out[f"pred_{tag}"] = torch.stack(preds).numpy()
        out[f"state_{tag}"] = torch.stack(states).numpy()
    post = [torch.stack(D.compute_gaussian_product_coef(ddpm.get_std_fwd(t_steps[:, i + 1]),
                                                        (ddpm.get_std_fwd(t_steps[:, i]) ** 2 - ddpm.get_std_fwd(t_steps[:, i + 1]) ** 2).sqrt()))
            for i in range(t_steps.shape[1] - 1)]
    out["posterior_coefs"] = torch.stack(post).numpy()
    np.savez_compressed(os.path.join(OUT, "sampler.npz"), **out)
    print("wrote sampler.npz with", len(out), "arrays")


def make_fast_inpaint():
    """tests/golden/fast_inpaint.npz: A2SBModel.fast_inpaint_ddpm_sample (A2SB_lightning_module.py:149-180) restated
    around the reference's own diffusion.py / utils.py functions and `reference_ddpm_sample` above: two holes shorter
    than the window, one of them near the right edge (exercises the window shift), width 150 padded to 160."""
    T, D, U, C = import_reference()
    ddpm = D.Diffusion()
    g = torch.Generator().manual_seed(91)
    x_1 = torch.randn(1, 3, 4, 150, generator=g)
    mask = torch.zeros(1, 3, 4, 150)
    mask[..., 40:52] = 1
    mask[..., 137:146] = 1
    _x, _m, t_steps, t_to_emb, net = sampler_setup()
    win, hop, bs = 32, 32, 4
    original_width = x_1.shape[-1]
    x = D.multidiffusion_pad_inputs(x_1.clone(), win, hop)
    m = D.multidiffusion_pad_inputs(mask, win, hop, padding_constant=0)
    windows = []
    for center_idx in U.find_middle_of_zero_segments(1 - m[0, 0, 0]):
        l_idx, r_idx = int(center_idx - win / 2), int(center_idx + win / 2)
        if l_idx < 0:
            r_idx -= l_idx
            l_idx = 0
        if r_idx > x.shape[-1]:
            l_idx -= (r_idx - x.shape[-1])
            r_idx = x.shape[-1]
        assert r_idx - l_idx == win and l_idx >= 0 and r_idx <= x.shape[-1]
        windows.append((l_idx, r_idx))
        preds, _ = reference_ddpm_sample(D, ddpm, net, t_to_emb, x[:, :, :, l_idx:r_idx], t_steps, m[:, :, :, l_idx:r_idx],
                                         True, win, hop, bs)
        x[:, :, :, l_idx:r_idx] = preds[-1]
    out = {"x_1": x_1.numpy(), "mask": mask.numpy().astype(np.uint8), "t_steps": t_steps.numpy(),
           "windows": np.asarray(windows, np.int32), "result": D.multidiffusion_unpad_outputs(x, original_width).numpy()}
    np.savez_compressed(os.path.join(OUT, "fast_inpaint.npz"), **out)
    print("wrote fast_inpaint.npz; windows", windows)


def make_consumers():
    """tests/golden/consumers.npz: the torch.stft / torch.istft calls of the other STFT users in the reference repository,
    with the arguments of their call sites (the modules themselves need einops_exts / librosa, absent here):
    ETTA STFT.encode / decode (adp.py:1536-1586, normalized=True) and auraloss STFTLoss.stft (auraloss.py:363-381)."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import a2sb_oracle as O
    out = {}
    # ETTA STFT, power-of-two num_fft (the default 1023 is not supported, see stft_consumers.py)
    n_fft, hop, t = 1024, 256, 2 ** 13
    wave = torch.from_numpy(np.stack([O.synth_noise(t, 21), O.synth_tonal(t)]))[None]            # [1, 2, t]
    win = torch.hann_window(n_fft)
    st = torch.stft(wave[0], n_fft=n_fft, hop_length=hop, win_length=n_fft, window=win, return_complex=True, normalized=True)
    out["etta_wave"] = wave.numpy()
    out["etta_real"], out["etta_imag"] = st.real.numpy(), st.imag.numpy()
    out["etta_mag"], out["etta_phase"] = torch.abs(st).numpy(), torch.angle(st).numpy()
    out["etta_decode"] = torch.istft(st, n_fft=n_fft, hop_length=hop, win_length=n_fft, window=win, length=t, normalized=True).numpy()
    # ETTA STFT with its DEFAULT num_fft = 1023 (odd length: 512 bins), hop 256, length = closest power of two to frames * hop
    n_odd = 1023
    win_o = torch.hann_window(n_odd)
    so = torch.stft(wave[0], n_fft=n_odd, hop_length=hop, win_length=n_odd, window=win_o, return_complex=True, normalized=True)
    out["etta1023_real"], out["etta1023_imag"] = so.real.numpy(), so.imag.numpy()
    out["etta1023_mag"], out["etta1023_phase"] = torch.abs(so).numpy(), torch.angle(so).numpy()
    out["etta1023_decode"] = torch.istft(so, n_fft=n_odd, hop_length=hop, win_length=n_odd, window=win_o, length=t, normalized=True).numpy()
    # auraloss multi-resolution STFT (fft sizes / hops / window lengths of the reference's default)
    x = torch.from_numpy(np.stack([O.synth_noise(7000, 22) + O.synth_tonal(7000)]))
    out["aura_x"] = x.numpy()
    for fs, hs, wl in ((1024, 120, 600), (2048, 240, 1200), (512, 50, 240)):
        xs = torch.stft(x, fs, hs, wl, torch.hann_window(wl), return_complex=True)
        out[f"aura_mag_{fs}"] = torch.sqrt(torch.clamp(xs.real ** 2 + xs.imag ** 2, min=1e-8)).numpy()
        out[f"aura_phs_{fs}"] = torch.angle(xs).numpy()
    np.savez_compressed(os.path.join(OUT, "consumers.npz"), **out)
    print("wrote consumers.npz with", len(out), "arrays")


def make_griffinlim():
    """tests/golden/griffinlim.npz: the reference's MagInstPhaseToGriffinLim (128 iterations) and a 4-iteration run
    of its `griffinlim` on a small seeded spectrogram (n_fft 512, hop 128)."""
    T, D, U, C = import_reference()
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import a2sb_oracle as O
    wav = torch.from_numpy(O.synth_tonal(23 * 128) + 0.05 * O.synth_noise(23 * 128, 5))
    msp = T.ComplexToMagInstPhase()(T.ComplexSpectrogram(512, 512, 128)(wav))
    out = {"msp": msp.numpy()}
    torch.manual_seed(0)
    out["gl128"] = T.MagInstPhaseToGriffinLim(512, 512, 128)(msp).numpy()
    torch.manual_seed(1)
    out["gl4"] = T.griffinlim(msp[0], None, None, window=torch.hann_window(512), n_fft=512, hop_length=128, win_length=512,
                              power=1, n_iter=4, momentum=.99, length=None, rand_init=True).numpy()
    out["gl4_init"] = T.griffinlim(msp[0], msp[1], msp[2], window=torch.hann_window(512), n_fft=512, hop_length=128,
                                   win_length=512, power=1, n_iter=4, momentum=.99, length=None, rand_init=False).numpy()
    np.savez_compressed(os.path.join(OUT, "griffinlim.npz"), **out)
    print("wrote griffinlim.npz")


if __name__ == "__main__":
    if "--gl" in sys.argv:
        make_griffinlim()
    elif "--sampler" in sys.argv:
        make_sampler()
    elif "--fast-inpaint" in sys.argv:
        make_fast_inpaint()
    elif "--consumers" in sys.argv:
        make_consumers()
    elif "--masks" in sys.argv:
        make_masks()
    else:
        main()
        make_masks()
        make_sampler()
        make_fast_inpaint()
        make_consumers()
        make_griffinlim()
