"""GPU parity: the sm_100a kernels, called through the C ABI (ctypes) and the reference-facing
Python API, against (a) the fixtures produced by the unmodified reference (tests/golden), (b) the
CPU oracle on seeded inputs, (c) size-independent properties at BASELINE.json's full sizes.

Tolerances (north_star): magnitude max rel err <= 1e-4 (relative to max(|b|, 1e-3 max|b|));
waveform parity and DC-retaining round trip >= 100 dB SNR; framing / segment indexing bit-exact.
"""
import numpy as np
import pytest

import a2sb_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu

NFFTS = (512, 1024, 2048, 4096)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    from audio_intelligence_b200 import _lib
    assert _lib.lib().a2sb_is_device_build() == 1       # the CUDA library, never an emulation
    return torch


@pytest.fixture(scope="module")
def T(torch_cuda):
    from audio_intelligence_b200.audio_transforms import transforms
    return transforms


@pytest.fixture(scope="module")
def D(torch_cuda):
    from audio_intelligence_b200 import diffusion
    return diffusion


def chains(T, n_fft, hop):
    fwd = [T.ComplexSpectrogram(n_fft, n_fft, hop), T.ComplexToMagInstPhase(), T.SpectrogramDropDCTerm(),
           T.PowerScaleSpectrogram(0.25, [0])]
    inv = [T.PowerScaleSpectrogram(4, [0]), T.SpectrogramAddDCTerm(), T.SVDFixMagInstPhase(),
           T.MagInstPhaseToComplex(), T.InverseComplexSpectrogram(n_fft, n_fft, hop)]
    return fwd, inv


def to_np(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------ golden fixtures


@pytest.mark.parametrize("n_fft", NFFTS)
def test_forward_chain_vs_reference_fixture(torch_cuda, T, n_fft):
    torch = torch_cuda
    g = load_golden(f"chain_n{n_fft}.npz")
    hop = n_fft // 4
    fwd, _ = chains(T, n_fft, hop)
    from audio_intelligence_b200 import _lib
    n0 = _lib.launch_count()
    spec, mask = T.apply_audio_transforms(torch.from_numpy(g["wav"]).cuda(), fwd)
    assert _lib.launch_count() - n0 == 1                  # the whole forward chain is ONE kernel
    assert mask is None and spec.is_cuda and spec.is_contiguous()
    spec = to_np(spec)
    assert spec.shape == g["spec"].shape
    assert O.mag_rel_err(g["spec"][0] ** 4, spec[0] ** 4) <= 1e-4
    weighted, strong = O.phase_err(g["spec"], spec)
    assert weighted <= 1e-6 and strong <= 1e-5         # complex error vs the peak; phases of bins >= 1% of it
    assert np.abs((spec[1] ** 2 + spec[2] ** 2) - 1).max() <= 1e-5
    c = to_np(T.ComplexSpectrogram(n_fft, n_fft, hop)(torch.from_numpy(g["wav"]).cuda()))
    assert c.shape == g["complex_spec"].shape
    assert np.abs(c - g["complex_spec"]).max() <= 2e-6 * np.abs(g["complex_spec"]).max()


@pytest.mark.parametrize("n_fft", NFFTS)
def test_inverse_chain_vs_reference_fixture(torch_cuda, T, n_fft):
    torch = torch_cuda
    g = load_golden(f"chain_n{n_fft}.npz")
    hop = n_fft // 4
    _, inv = chains(T, n_fft, hop)
    inv_nosvd = [t for t in inv if not isinstance(t, T.SVDFixMagInstPhase)]
    from audio_intelligence_b200 import _lib
    n0 = _lib.launch_count()
    y, _ = T.apply_audio_transforms(torch.from_numpy(g["spec"]).cuda(), inv)
    assert _lib.launch_count() - n0 == 1                  # the whole inverse chain is ONE kernel
    assert y.shape == g["wav_inv"].shape
    assert O.snr_db(g["wav_inv"], to_np(y)) >= 100
    y, _ = T.apply_audio_transforms(torch.from_numpy(g["spec_pert"]).cuda(), inv)
    assert O.snr_db(g["wav_pert"], to_np(y)) >= 100
    y, _ = T.apply_audio_transforms(torch.from_numpy(g["spec"]).cuda(), inv_nosvd)
    assert O.snr_db(g["wav_inv_nosvd"], to_np(y)) >= 100
    y = T.InverseComplexSpectrogram(n_fft, n_fft, hop)(torch.from_numpy(g["complex_spec"]).cuda())
    assert O.snr_db(g["wav_cplx"], to_np(y)) >= 100


def test_tonal_fixture(torch_cuda, T):
    torch = torch_cuda
    g = load_golden("chain_tonal.npz")
    fwd, inv = chains(T, 2048, 512)
    spec, _ = T.apply_audio_transforms(torch.from_numpy(g["wav"]).cuda(), fwd)
    assert O.mag_rel_err(g["spec"][0] ** 4, to_np(spec)[0] ** 4) <= 1e-4
    y, _ = T.apply_audio_transforms(torch.from_numpy(g["spec"]).cuda(), inv)
    assert O.snr_db(g["wav_inv"], to_np(y)) >= 100


def test_unfused_ops_match_fused_chain_and_fixture(torch_cuda, T):
    """Each op stand-alone (list-indexable, A2SB_lightning_module.py:501) and the op-by-op chain."""
    torch = torch_cuda
    g = load_golden("ops.npz")
    msp = torch.from_numpy(g["msp"]).cuda()
    np.testing.assert_allclose(to_np(T.SVDFixMagInstPhase()(msp)), g["svd_fix"], atol=5e-6)
    np.testing.assert_array_equal(to_np(T.MagInstPhaseToComplex()(msp)), g["to_complex"])
    np.testing.assert_allclose(to_np(T.ComplexToMagInstPhase()(msp[:2])), g["to_magphase"], atol=2e-6)
    np.testing.assert_allclose(to_np(T.PowerScaleSpectrogram(0.5)(msp)), g["pow_half_all"], rtol=5e-6, atol=1e-7)
    np.testing.assert_allclose(to_np(T.PowerScaleSpectrogram(0.25, [0])(msp)), g["pow_quarter_c0"], rtol=5e-6, atol=1e-7)
    np.testing.assert_allclose(to_np(T.PowerScaleSpectrogram(4, [0])(msp)), g["pow_four_c0"], rtol=5e-6, atol=1e-7)
    np.testing.assert_array_equal(to_np(T.SpectrogramAddDCTerm()(msp)), g["add_dc"])
    np.testing.assert_array_equal(to_np(T.SpectrogramDropDCTerm()(msp)), g["drop_dc"])
    # op-by-op == fused, to fp32 rounding of the differently ordered arithmetic
    gg = load_golden("chain_n1024.npz")
    fwd, inv = chains(T, 1024, 256)
    x = torch.from_numpy(gg["wav"]).cuda()
    fused, _ = T.apply_audio_transforms(x, fwd)
    step = x
    for op in fwd:
        step = op(step)
    assert O.mag_rel_err(to_np(fused)[0] ** 4, to_np(step)[0] ** 4) <= 1e-5
    yf, _ = T.apply_audio_transforms(fused, inv)
    ys = fused
    for op in inv:
        ys = op(ys)
    assert O.snr_db(to_np(yf), to_np(ys)) >= 110


# ------------------------------------------------------------------------------ oracle, seeded inputs


@pytest.mark.parametrize("n_fft,L", [(512, 257), (512, 1000), (512, 5000), (1024, 513), (1024, 44100),
                                     (2048, 1025), (2048, 2048), (2048, 44100 * 2 + 17), (2048, 16 * 512 * 5)])
def test_ragged_lengths_vs_oracle(torch_cuda, T, n_fft, L):
    torch = torch_cuda
    hop = n_fft // 4
    fwd, inv = chains(T, n_fft, hop)
    wav = np.stack([O.synth_noise(L, 1000 + i) for i in range(3)])
    spec, _ = T.apply_audio_transforms(torch.from_numpy(wav).cuda(), fwd)       # batched: [B, 3, rows, T]
    assert tuple(spec.shape) == (3, 3, n_fft // 2, O.num_frames(L, hop))
    s = to_np(spec)
    assert not np.isnan(s).any()
    y, _ = T.apply_audio_transforms(spec, inv)
    assert tuple(y.shape) == (3, O.istft_length(s.shape[-1], hop))
    yn = to_np(y)
    for i in range(3):
        ref = O.forward_chain(wav[i], n_fft, hop)
        assert O.mag_rel_err(ref[0] ** 4, s[i, 0] ** 4) <= 1e-4
        assert O.snr_db(O.inverse_chain(s[i], n_fft, hop), yn[i]) >= 100
    # unbatched call == row of the batched call, bit for bit
    one, _ = T.apply_audio_transforms(torch.from_numpy(wav[1]).cuda(), fwd)
    np.testing.assert_array_equal(to_np(one), s[1])


@pytest.mark.parametrize("n_fft", NFFTS)
def test_dc_retaining_round_trip_snr(torch_cuda, T, n_fft):
    """STFT -> mag/phase -> power .25 -> power 4 -> phase fix -> complex -> iSTFT, DC row kept."""
    torch = torch_cuda
    hop = n_fft // 4
    wav = O.synth_noise(44100 * 3, 7)
    fwd = [T.ComplexSpectrogram(n_fft, n_fft, hop), T.ComplexToMagInstPhase(), T.PowerScaleSpectrogram(0.25, [0])]
    inv = [T.PowerScaleSpectrogram(4, [0]), T.SVDFixMagInstPhase(), T.MagInstPhaseToComplex(),
           T.InverseComplexSpectrogram(n_fft, n_fft, hop)]
    spec, _ = T.apply_audio_transforms(torch.from_numpy(wav).cuda(), fwd)
    assert spec.shape[1] == n_fft // 2 + 1
    y, _ = T.apply_audio_transforms(spec, inv)
    yn = to_np(y)
    assert O.snr_db(wav[: yn.shape[0]], yn) >= 100
    # full A2SB chain (DC dropped) is lossy by design; it must be lossy by the SAME amount as the oracle
    fwd2, inv2 = chains(T, n_fft, hop)
    s2, _ = T.apply_audio_transforms(torch.from_numpy(wav).cuda(), fwd2)
    y2 = to_np(T.apply_audio_transforms(s2, inv2)[0])
    ref = O.inverse_chain(O.forward_chain(wav, n_fft, hop), n_fft, hop)
    assert abs(O.snr_db(wav[: y2.shape[0]], y2) - O.snr_db(wav[: ref.shape[0]], ref)) <= 0.1


def test_edge_signals(torch_cuda, T):
    torch = torch_cuda
    n_fft, hop, L = 2048, 512, 30000
    fwd, inv = chains(T, n_fft, hop)
    z = torch.zeros(L, device="cuda")
    spec, _ = T.apply_audio_transforms(z, fwd)
    s = to_np(spec)
    assert (s[0] == 0).all() and (s[1] == 1).all() and (s[2] == 0).all()
    assert (to_np(T.apply_audio_transforms(spec, inv)[0]) == 0).all()
    for pos in (0, L - 1):
        imp = np.zeros(L, np.float32)
        imp[pos] = 1.0
        s = to_np(T.apply_audio_transforms(torch.from_numpy(imp).cuda(), fwd)[0])
        ref = O.forward_chain(imp, n_fft, hop)
        assert O.mag_rel_err(ref[0] ** 4, s[0] ** 4) <= 1e-4
    # NaN/Inf in row 0 propagate into the re-created DC bin exactly like `spec[..., :1, :] * 0`
    sp = np.array(O.forward_chain(O.synth_noise(L, 3), n_fft, hop))
    sp[0, 0, 5] = np.inf
    y = to_np(T.apply_audio_transforms(torch.from_numpy(sp).cuda(), inv)[0])
    ref = O.inverse_chain(sp, n_fft, hop)
    assert np.array_equal(np.isnan(y), np.isnan(ref)) and np.isnan(y).any()


def test_cpu_tensors_are_staged_and_returned_on_cpu(torch_cuda, T):
    torch = torch_cuda
    fwd, inv = chains(T, 1024, 256)
    wav = torch.from_numpy(O.synth_noise(20000, 5))
    spec, _ = T.apply_audio_transforms(wav, fwd)
    assert not spec.is_cuda
    y, _ = T.apply_audio_transforms(spec, inv)
    assert not y.is_cuda and y.shape[0] == 256 * (spec.shape[-1] - 1)
    ref = O.inverse_chain(spec.numpy(), 1024, 256)
    assert O.snr_db(ref, y.numpy()) >= 100


def test_error_behaviour(torch_cuda, T):
    torch = torch_cuda
    with pytest.raises(RuntimeError, match="Padding size should be less than"):
        T.ComplexSpectrogram(2048, 2048, 512)(torch.zeros(1024, device="cuda"))
    with pytest.raises(AssertionError):
        T.ComplexSpectrogram(2048, 2048, 512)(torch.zeros(1, 1, 4096, device="cuda"))
    with pytest.raises(AssertionError):
        T.InverseComplexSpectrogram(2048, 2048, 512)(torch.zeros(1025, 9, device="cuda"))
    with pytest.raises(RuntimeError):
        T.ComplexSpectrogram(1000, 1000, 250)(torch.zeros(4096, device="cuda"))
    ns = T.Namespace(class_path="audio_intelligence_b200.audio_transforms.transforms.PowerScaleSpectrogram",
                     init_args=T.Namespace(power=4, channels=[0]))
    out, _ = T.apply_audio_transforms(torch.ones(3, 4, 5, device="cuda") * 2, [ns])
    assert abs(float(out[0, 0, 0]) - 16.0) < 1e-4 and float(out[1, 0, 0]) == 2.0


def test_sharded_ranges_bit_identical(torch_cuda):
    """Frame / output ranges computed from local windows equal the unsharded result bit for bit."""
    torch = torch_cuda
    from audio_intelligence_b200 import _capi, _lib
    n_fft, hop, L = 2048, 512, 44100 * 4
    wav = torch.from_numpy(O.synth_noise(L, 9)).cuda()[None]
    full = _lib.stft_forward(wav, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25)
    Tn = full.shape[-1]
    t0, t1 = 37, 201
    lo, hi = max(t0 * hop - n_fft // 2, 0), min((t1 - 1) * hop + n_fft // 2, L)
    part = _lib.stft_forward(wav[:, lo:hi].contiguous(), n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, drop_dc=True,
                             power=0.25, total_len=L, sample_first=lo, t_range=(t0, t1))
    assert torch.equal(part, full[..., t0:t1])
    y = _lib.istft_inverse(full, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True, power=4.0)
    o0, on = 40 * hop, 150 * hop
    f_lo = max((o0 + n_fft // 2) // hop - 3, 0)
    f_hi = min((o0 + on + n_fft // 2 + hop - 1) // hop, Tn)
    yp = _lib.istft_inverse(full[..., f_lo:f_hi].contiguous(), n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE,
                            has_dc=False, phase_fix=True, power=4.0, n_frames=Tn, spec_t_first=f_lo,
                            out_range=(o0, on))
    assert torch.equal(yp, y[:, o0:o0 + on])


# ------------------------------------------------------------------------------ segments / blend


def test_segments_bit_exact_vs_fixture(torch_cuda, D):
    torch = torch_cuda
    g = load_golden("blend.npz")
    x = torch.from_numpy(g["x"]).cuda()
    xp = D.multidiffusion_pad_inputs(x, 64, 32)
    np.testing.assert_array_equal(to_np(xp), g["xp"])
    np.testing.assert_array_equal(to_np(D.multidiffusion_pad_inputs(x, 64, 32, padding_constant=0)), g["xp_const"])
    t = torch.zeros(2, 4, device="cuda")
    np.testing.assert_array_equal(to_np(D.get_multidiffusion_vf(lambda a, e: a, xp, t, 64, 32, 5)), g["ident"])
    np.testing.assert_array_equal(to_np(D.get_multidiffusion_vf(lambda a, e: a * 2 + 0.1, xp, t, 64, 32, 5)), g["affine"])

    def ramp(a, e):
        return a + torch.arange(a.shape[0], dtype=a.dtype, device=a.device).view(-1, 1, 1, 1) * 0.001
    np.testing.assert_array_equal(to_np(D.get_multidiffusion_vf(ramp, xp, t, 64, 32, 1000)), g["ramp"])
    xp3 = torch.from_numpy(g["xp3"]).cuda()
    out = D.get_multidiffusion_vf(lambda a, e: a * 1.7 - 0.3, xp3, torch.zeros(1, 4, device="cuda"), 48, 16, 3)
    np.testing.assert_array_equal(to_np(out), g["noisy3"])
    np.testing.assert_array_equal(to_np(D.multidiffusion_unpad_outputs(xp, 300)), g["xp"][..., :300])


def test_segments_pad_widths(torch_cuda, D, known_answers):
    torch = torch_cuda
    for w, padded in known_answers["pad_widths"].items():
        if int(w) > 10000:
            continue
        out = D.multidiffusion_pad_inputs(torch.zeros(1, 1, 2, int(w), device="cuda"), 256, 128)
        assert out.shape[-1] == padded


def test_segments_t_emb_chunking_matches_reference_contract(torch_cuda, D):
    """The network sees torch.chunk-sized mini-batches with matching t_emb rows (diffusion.py:43-50)."""
    torch = torch_cuda
    x = torch.randn(2, 3, 8, 128 * 9 + 128, device="cuda")
    seen = []

    def net(a, e):
        seen.append((a.shape[0], e.shape[0]))
        return a
    t = torch.randn(2, 6, device="cuda")
    out = D.get_multidiffusion_vf(net, x, t, 256, 128, 16)
    assert torch.equal(out, x)
    n = 2 * ((x.shape[-1] - 128) // 128)
    chunks = -(-n // 16)
    per = -(-n // chunks)
    assert [s[0] for s in seen] == [min(per, n - i) for i in range(0, n, per)]
    assert all(a == b for a, b in seen)


# ------------------------------------------------------------------------------ full sizes (properties)


def test_config2_full_size_properties(torch_cuda):
    """BASELINE config 2: 256 x 10 s clips, n_fft 2048 / hop 512.  Checked through size-independent
    properties: clip 0 and clip 255 against the oracle; linearity of the complex STFT; the
    DC-retaining round trip >= 100 dB on every clip; batched == per-clip bit for bit."""
    torch = torch_cuda
    from audio_intelligence_b200 import _capi, _lib
    n_fft, hop, L, B = 2048, 512, 441000, 256
    g = torch.Generator(device="cuda").manual_seed(1000)
    wav = (0.3 * torch.randn(B, L, generator=g, device="cuda")).clamp_(-1, 1)
    spec = _lib.stft_forward(wav, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25)
    assert tuple(spec.shape) == (B, 3, 1024, 862)
    for i in (0, 255):
        ref = O.forward_chain(to_np(wav[i]), n_fft, hop)
        assert O.mag_rel_err(ref[0] ** 4, to_np(spec[i, 0]) ** 4) <= 1e-4
    one = _lib.stft_forward(wav[100:101], n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25)
    assert torch.equal(one[0], spec[100])
    y = _lib.istft_inverse(spec, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True, power=4.0)
    assert tuple(y.shape) == (B, 440832)
    assert O.snr_db(O.inverse_chain(to_np(spec[255]), n_fft, hop), to_np(y[255])) >= 100
    del spec, y
    # DC-retaining round trip on all clips (SNR per clip, computed on the device in fp64)
    s = _lib.stft_forward(wav, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, drop_dc=False, power=0.25)
    y = _lib.istft_inverse(s, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, has_dc=True, phase_fix=True, power=4.0)
    ref = wav[:, : y.shape[1]].double()
    err = (y.double() - ref)
    snr = 10 * torch.log10((ref * ref).sum(1) / (err * err).sum(1))
    assert float(snr.min()) >= 100.0, float(snr.min())
    del s, y, ref, err
    # linearity of the complex transform: STFT(a + 2b) == STFT(a) + 2 STFT(b)
    a, b = wav[:4], wav[4:8]
    ca = _lib.stft_forward(a, n_fft, n_fft, hop, kind=_capi.KIND_COMPLEX)
    cb = _lib.stft_forward(b, n_fft, n_fft, hop, kind=_capi.KIND_COMPLEX)
    cab = _lib.stft_forward((a + 2 * b).contiguous(), n_fft, n_fft, hop, kind=_capi.KIND_COMPLEX)
    scale = float(cab.abs().max())
    assert float((cab - (ca + 2 * cb)).abs().max()) <= 2e-6 * scale


def test_config3_hour_long_blend_identity(torch_cuda, D, known_answers):
    """BASELINE config 3 geometry: 1 h of audio = 310,079 frames -> padded 310,144 -> 2422 segments.
    Identity and affine stubs: the blend returns its input bit-exactly (SURVEY 8c pin 4)."""
    torch = torch_cuda
    W = 310079
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(1, 3, 1024, W, generator=g, device="cuda")
    xp = D.multidiffusion_pad_inputs(x, 256, 128)
    assert xp.shape[-1] == known_answers["pad_widths"]["310079"] == 310144
    assert torch.equal(xp[..., :W], x) and torch.equal(xp[..., W:], x[..., : 310144 - W])
    calls = []

    def net(a, e):
        calls.append(a.shape[0])
        return a
    t = torch.zeros(1, 4, device="cuda")
    out = D.get_multidiffusion_vf(net, xp, t, 256, 128, 16)
    assert sum(calls) == 2422
    assert torch.equal(out, xp)
    del out
    out = D.get_multidiffusion_vf(lambda a, e: a * 2 + 0.1, xp, t, 256, 128, 4096)
    # every column is covered once or twice by identical values v: (v+v)/2 == v and v/1 == v exactly
    assert torch.equal(out, xp * 2 + 0.1)
    assert torch.equal(D.multidiffusion_unpad_outputs(out, W), (x * 2 + 0.1))


# ------------------------------------------------------------------ corruption masks / zero segments (rows M1, B4)


def _mask_cases():
    return [
        ("multinomial", dict(p_upsample_mask=0.4, p_extension_mask=0.3, p_inpaint_mask=0.3, fill_noise_level=0.5,
                             sampling_rate=44100, upsample_mask_kwargs=dict(min_cutoff_freq=2000, max_cutoff_freq=8000),
                             inpainting_mask_kwargs=dict(min_inpainting_frac=0.05, max_inpainting_frac=0.4, is_random=True))),
        ("timestamped", dict(start_time=0.1, end_time=0.35, hop_length=512, sampling_rate=44100, fill_noise_level=0.5)),
    ]   # == oracle/make_golden.py::mask_cases


def test_corruption_transforms_vs_reference_fixture(torch_cuda):
    """Seeded CPU-tensor runs of the mirrored corruption transforms reproduce the reference's masks exactly and --
    when this host's CPU generator produces the stream the fixtures were drawn from -- its noise-filled values too."""
    torch = torch_cuda
    from audio_intelligence_b200 import _lib
    from audio_intelligence_b200.corruption import corruptions as C
    g = load_golden("masks.npz")
    spec = torch.from_numpy(g["spec"])
    torch.manual_seed(0)
    same_rng = np.array_equal(torch.randn(64).numpy(), g["rng_probe"])
    for name, kw in _mask_cases():
        cls = C.MultinomialInpaintMaskTransform if name == "multinomial" else C.TimestampedSegmentInpaintMaskTransform
        for seed in range(6):
            torch.manual_seed(seed)
            np.random.seed(seed)
            n0 = _lib.launch_count()
            filled, mask = cls(**kw)(spec.clone())
            assert _lib.launch_count() - n0 == 1                 # mask + fill: ONE kernel
            assert not filled.is_cuda and not mask.is_cuda       # returned where the input lives
            m = g[f"{name}_{seed}_mask"].astype(np.float32)
            assert np.array_equal(mask.numpy(), m)
            keep = m == 0
            assert np.array_equal(filled.numpy()[keep], g["spec"][keep])
            if same_rng:
                assert np.array_equal(filled.numpy(), g[f"{name}_{seed}_filled"])
    for seed in range(6):
        torch.manual_seed(100 + seed)
        m = C.UpsampleMask.get_upsample_mask(torch.zeros(3, 128, 5), 1000, 9000, 44100)
        assert np.array_equal(m.numpy(), g[f"upsample_{seed}"].astype(np.float32))
        m = C.ExtensionMask.get_extension_mask(torch.zeros(3, 4, 200), 32)
        assert np.array_equal(m.numpy(), g[f"extension_{seed}"].astype(np.float32))
        np.random.seed(100 + seed)
        m = C.InpaintMask.get_inpainting_mask(torch.zeros(3, 4, 200), 0.1, 0.5, seed % 2 == 0)
        assert np.array_equal(m.numpy(), g[f"inpaint_{seed}"].astype(np.float32))


def test_mask_with_noise_bit_exact_on_device(torch_cuda, known_answers):
    """CUDA inputs: same generator call as the reference would make on the device, fp32 op order of corruptions.py:15."""
    torch = torch_cuda
    from audio_intelligence_b200.corruption import corruptions as C
    x = torch.randn(3, 1024, 862, device="cuda")
    t = C.TimestampedSegmentInpaintMaskTransform(1.0, 1.2, 512, 44100, 0.5)
    assert [t.start_idx, t.end_idx] == known_answers["inpaint_frames_1.0_1.2"]
    torch.manual_seed(5)
    filled, mask = t(x)
    torch.manual_seed(5)
    noise = torch.randn_like(x)
    ref_mask = torch.zeros_like(x)
    ref_mask[:, :, t.start_idx:t.end_idx] = 1
    assert torch.equal(mask, ref_mask)
    assert torch.equal(filled, x * (1 - ref_mask) + ref_mask * noise * 0.5)
    torch.manual_seed(6)
    out = C.mask_with_noise(x, ref_mask, 0.25)
    torch.manual_seed(6)
    assert torch.equal(out, x * (1 - ref_mask) + ref_mask * torch.randn_like(x) * 0.25)
    for n_fft, row in known_answers["upsample_first_row"].items():
        m = C.UpsampleMask.get_upsample_mask(torch.zeros(3, int(n_fft) // 2, 4, device="cuda"), 4000, 4000, 44100)
        assert int(torch.nonzero(m[0, :, 0])[0, 0]) == row and m.is_cuda


def test_zero_segments_vs_reference_fixture(torch_cuda, known_answers):
    torch = torch_cuda
    from audio_intelligence_b200 import utils as U
    g = load_golden("masks.npz")
    for seed in range(6):
        row = torch.from_numpy(g[f"zero_row_{seed}"].astype(np.float32)).cuda()
        mids = U.find_middle_of_zero_segments(row)
        assert mids.is_cuda and mids.dtype == torch.int32
        assert mids.cpu().tolist() == g[f"zero_mid_{seed}"].tolist()
    mask_row = torch.ones(896)
    for a, b in ((86, 103), (318, 344), (800, 896)):
        mask_row[a:b] = 0
    assert U.find_middle_of_zero_segments(mask_row).tolist() == known_answers["zero_segment_centres"]
    assert [list(w) for w in U.zero_segment_windows(mask_row.cuda(), 256)] == known_answers["inpaint_windows"]
    with pytest.raises(ValueError):
        U.find_middle_of_zero_segments(torch.zeros(2, 3))
    # config 3 width: one hole per minute of the padded hour
    row = torch.ones(310144, device="cuda")
    want = []
    for k in range(60):
        a = 1000 + k * 5000
        row[a:a + 40 + k] = 0
        want.append(int((a + a + 40 + k - 1) / 2))
    assert U.find_middle_of_zero_segments(row).cpu().tolist() == want


@pytest.mark.parametrize("n_fft", NFFTS)
def test_config4_mask_and_reconstruction_sweep(torch_cuda, T, n_fft, known_answers):
    """BASELINE config 4: bandwidth-extension (4 kHz cut-off) + inpainting masks on the forward output, then the
    inverse chain, for every n_fft; each stage against the oracle on the same tensors."""
    torch = torch_cuda
    from audio_intelligence_b200.corruption import corruptions as C
    hop, sr = n_fft // 4, 44100
    wav = O.synth_noise(3 * sr, 1000)
    fwd, inv = chains(T, n_fft, hop)
    spec, _ = T.apply_audio_transforms(torch.from_numpy(wav).cuda(), fwd)
    ref_spec = O.forward_chain(wav, n_fft, hop)
    assert O.mag_rel_err(ref_spec[0] ** 4, to_np(spec)[0] ** 4) <= 1e-4
    torch.manual_seed(n_fft)
    bwe = C.MultinomialInpaintMaskTransform(1.0, 0.0, 0.0, 0.5, sr, dict(min_cutoff_freq=4000, max_cutoff_freq=4000),
                                            dict(min_inpainting_frac=0.1, max_inpainting_frac=0.2, is_random=False))
    x, m1 = bwe(spec)
    first = known_answers["upsample_first_row"][str(n_fft)]
    assert not m1[:, :first].any() and m1[:, first:].all()
    masks = [m1]
    for t0, t1 in ((1.0, 1.2), (2.0, 2.5)):
        tr = C.TimestampedSegmentInpaintMaskTransform(t0, t1, hop, sr, 0.5)
        assert (tr.start_idx, tr.end_idx) == O.inpaint_frames(t0, t1, hop, sr)
        x, m = tr(x)
        assert m[:, :, tr.start_idx:tr.end_idx].all() and m.sum().item() == 3 * (n_fft // 2) * (tr.end_idx - tr.start_idx)
        masks.append(m)
    total = torch.stack(masks).sum(0).clamp(0, 1)             # apply_audio_transforms' mask accumulation (:75-79)
    keep = to_np(total) == 0
    assert np.array_equal(to_np(x)[keep], to_np(spec)[keep])  # untouched outside the masks, bit for bit
    y, _ = T.apply_audio_transforms(x, inv)
    want = O.inverse_chain(to_np(x), n_fft, hop)
    assert y.shape == want.shape and O.snr_db(want, to_np(y)) >= 100


# ------------------------------------------------------------------ bridge sampler step


def _torch_sampler_stubs(torch):
    def t_to_emb(t):
        return torch.stack([t, t * t], dim=1)

    def net(x, t_emb):
        pos = torch.arange(x.shape[-1], dtype=x.dtype, device=x.device)
        return x * 0.8 + t_emb[:, :1, None, None].to(x.device) * 0.1 + pos * 0.01
    return t_to_emb, net   # == oracle/make_golden.py::sampler_setup


def test_diffusion_schedule_vs_reference_fixture(torch_cuda, D):
    torch = torch_cuda
    g = load_golden("sampler.npz")
    ddpm = D.Diffusion()
    t = torch.from_numpy(g["t"])
    assert np.array_equal(ddpm.get_int_beta_0_t(t).numpy(), g["int_beta"])
    assert np.array_equal(ddpm.get_std_fwd(t).numpy(), g["std_fwd"])
    assert np.array_equal(ddpm.get_std_rev(t).numpy(), g["std_rev"])
    np.testing.assert_array_equal(ddpm.get_std_t(t).numpy(), g["std_t"])
    ts = torch.from_numpy(g["t_steps"])
    for i in range(ts.shape[1] - 1):
        got = torch.stack(ddpm.posterior_coefs(ts[:, i + 1], ts[:, i])).numpy()
        assert np.array_equal(got, g["posterior_coefs"][i])


@pytest.mark.parametrize("tag,mp", [("mp1", True), ("mp0", False)])
def test_ddpm_sample_vs_reference_fixture(torch_cuda, D, tag, mp):
    """4-step ot-ode sampling run of the reference (stub network, windows 64/32): bit-identical pred_x0 per step,
    one gather + one fused blend/step kernel per iteration."""
    torch = torch_cuda
    from audio_intelligence_b200 import _lib
    g = load_golden("sampler.npz")
    t_to_emb, net = _torch_sampler_stubs(torch)
    x_1, mask, ts = torch.from_numpy(g["x_1"]), torch.from_numpy(g["mask"].astype(np.float32)), torch.from_numpy(g["t_steps"])
    n0 = _lib.launch_count()
    preds = D.ddpm_sample(net, D.Diffusion(), x_1, ts, t_to_emb, mask=mask, mask_pred_x0=mp, win_length=64, hop_length=32,
                          batch_size=4, use_ot_ode=True)
    assert _lib.launch_count() - n0 == 2 + 2 * 4              # two wrap-pads, then (gather, blend+step) x 4 steps
    assert len(preds) == 4 and all(not p.is_cuda and p.shape == x_1.shape for p in preds)
    assert np.array_equal(torch.stack(preds).numpy(), g[f"pred_{tag}"])


def test_ddpm_sample_history_modes(torch_cuda, D):
    """Device inputs: the default keeps the history on the device (no per-step D2H); "async" returns pinned
    host copies made under the following steps; history="last" keeps only the final pred_x0.  All bit-identical."""
    torch = torch_cuda
    g = load_golden("sampler.npz")
    t_to_emb, net = _torch_sampler_stubs(torch)
    x_1 = torch.from_numpy(g["x_1"]).cuda()
    mask = torch.from_numpy(g["mask"].astype(np.float32)).cuda()
    ts = torch.from_numpy(g["t_steps"])
    kw = dict(mask=mask, win_length=64, hop_length=32, batch_size=4, use_ot_ode=True)
    dev = D.ddpm_sample(net, D.Diffusion(), x_1, ts, t_to_emb, **kw)
    assert len(dev) == 4 and all(p.is_cuda for p in dev)
    assert np.array_equal(torch.stack(dev).cpu().numpy(), g["pred_mp1"])
    host = D.ddpm_sample(net, D.Diffusion(), x_1, ts, t_to_emb, outputs_to_cpu="async", **kw)
    assert all((not p.is_cuda) for p in host) and np.array_equal(torch.stack(host).numpy(), g["pred_mp1"])
    last = D.ddpm_sample(net, D.Diffusion(), x_1, ts, t_to_emb, history="last", **kw)
    assert len(last) == 1 and torch.equal(last[0], dev[-1])


def test_fast_inpaint_ddpm_sample_vs_reference_fixture(torch_cuda, D):
    """A2SB_lightning_module.py:149-180: one sampling run per hole, pasted back; bit-identical to the reference-generated
    fixture, for host and device inputs."""
    torch = torch_cuda
    g = load_golden("fast_inpaint.npz")
    t_to_emb, net = _torch_sampler_stubs(torch)
    x_1, mask, ts = torch.from_numpy(g["x_1"]), torch.from_numpy(g["mask"].astype(np.float32)), torch.from_numpy(g["t_steps"])
    out = D.fast_inpaint_ddpm_sample(net, D.Diffusion(), x_1, ts, t_to_emb, mask=mask, win_length=32, hop_length=32, batch_size=4)
    assert isinstance(out, list) and len(out) == 1 and not out[0].is_cuda and out[0].shape == x_1.shape
    assert np.array_equal(out[0].numpy(), g["result"])
    out_d = D.fast_inpaint_ddpm_sample(net, D.Diffusion(), x_1.cuda(), ts, t_to_emb, mask=mask.cuda(), win_length=32,
                                       hop_length=32, batch_size=4)
    assert out_d[0].is_cuda and np.array_equal(out_d[0].cpu().numpy(), g["result"])
    assert np.array_equal(x_1.numpy(), g["x_1"])             # the input is not modified (the reference clones, :156)


def test_ddpm_sample_with_noise_bit_exact_vs_torch_on_device(torch_cuda, D):
    """use_ot_ode=False: the same generator calls in the same order as the reference loop, checked against the
    reference's expression evaluated with torch ops on the device."""
    torch = torch_cuda
    g = load_golden("sampler.npz")
    t_to_emb, net = _torch_sampler_stubs(torch)
    x_1 = torch.from_numpy(g["x_1"]).cuda()
    mask = torch.from_numpy(g["mask"].astype(np.float32)).cuda()
    ts = torch.from_numpy(g["t_steps"])
    ddpm = D.Diffusion()
    torch.manual_seed(11)
    got = D.ddpm_sample(net, ddpm, x_1, ts, t_to_emb, mask=mask, win_length=64, hop_length=32, batch_size=4,
                        use_ot_ode=False, outputs_to_cpu=False)
    torch.manual_seed(11)
    x1p = D.multidiffusion_pad_inputs(x_1, 64, 32)
    mp = D.multidiffusion_pad_inputs(mask, 64, 32)
    x_t = x1p.clone()
    for i in range(4):
        t, tp = ts[:, i], ts[:, i + 1]
        vf = D.get_multidiffusion_vf(net, x_t, t_to_emb(t).repeat(2, 1), 64, 32, 4)
        pred = x_t - ddpm.get_std_fwd(t).cuda() * vf
        pred = pred * mp + (1 - mp) * x1p
        assert torch.equal(got[i], pred[..., :x_1.shape[-1]])
        mu_x0, mu_xt, var = (c.cuda() for c in ddpm.posterior_coefs(tp, t))
        x_prev = mu_x0 * pred + mu_xt * x_t
        if tp > 0:
            x_prev = x_prev + var.sqrt() * torch.randn_like(x_prev)
        xt_true = x1p + ddpm.get_std_t(tp).cuda() * torch.randn_like(x1p)
        x_t = (1. - mp) * xt_true + mp * x_prev


def test_row_aligned_layout_is_a_view_with_identical_values(torch_cuda, T):
    """Opt-in 32-byte row pitch: same numbers as the contiguous layout, forward and inverse, fused chains."""
    torch = torch_cuda
    wav = torch.from_numpy(np.stack([O.synth_noise(44100, 1000 + i) for i in range(3)])).cuda()
    fwd, inv = chains(T, 2048, 512)
    spec, _ = T.apply_audio_transforms(wav, fwd)
    y, _ = T.apply_audio_transforms(spec, inv)
    T.set_row_alignment(8)
    try:
        spec_a, _ = T.apply_audio_transforms(wav, fwd)
        assert not spec_a.is_contiguous() and spec_a.stride(2) % 8 == 0 and spec_a.shape == spec.shape
        assert torch.equal(spec_a, spec)
        y_a, _ = T.apply_audio_transforms(spec_a, inv)
        assert torch.equal(y_a, y)
        s1, _ = T.apply_audio_transforms(wav[0], fwd)            # unbatched, like the reference's call sites
        y1, _ = T.apply_audio_transforms(s1, inv)
        assert torch.equal(s1, spec[0]) and torch.equal(y1, y[0])
    finally:
        T.set_row_alignment(None)
    with pytest.raises(ValueError):
        T.set_row_alignment(6)


# ------------------------------------------------------------------ Griffin-Lim


def test_griffinlim_vs_reference_fixture(torch_cuda, T):
    """K2 -> K1 -> phase update per iteration.  4 iterations: >= 100 dB vs the reference; 128 iterations of a chaotic
    fixed-point iteration amplify fp32 rounding differences, so there the gate is the quality of the result
    (spectral convergence equal to the reference's within 0.5 dB), not sample-wise equality."""
    torch = torch_cuda
    g = load_golden("griffinlim.npz")
    msp = torch.from_numpy(g["msp"])
    win = torch.hann_window(512)
    torch.manual_seed(1)
    y4 = T.griffinlim(msp[0], None, None, window=win, n_fft=512, hop_length=128, win_length=512, power=1, n_iter=4,
                      momentum=.99, length=None, rand_init=True)
    assert y4.shape == g["gl4"].shape and not y4.is_cuda
    assert O.snr_db(g["gl4"], y4.numpy()) >= 100
    y4i = T.griffinlim(msp[0].cuda(), msp[1].cuda(), msp[2].cuda(), window=win, n_fft=512, hop_length=128, win_length=512,
                       power=1, n_iter=4, momentum=.99, length=None, rand_init=False)
    assert y4i.is_cuda and O.snr_db(g["gl4_init"], to_np(y4i)) >= 100
    torch.manual_seed(0)
    y = T.MagInstPhaseToGriffinLim(512, 512, 128)(msp)
    assert y.shape == g["gl128"].shape

    def convergence_db(wav):       # || |STFT(y)| - mag || / || mag ||, the quantity Griffin-Lim minimises
        m = np.abs(O.stft_complex(np.asarray(wav, np.float32), 512, 128))
        return 20 * np.log10(np.linalg.norm(m - g["msp"][0]) / np.linalg.norm(g["msp"][0]))
    assert abs(convergence_db(y.numpy()) - convergence_db(g["gl128"])) <= 0.5
    with pytest.raises(ValueError):
        T.griffinlim(msp[0], None, None, window=win, n_fft=512, hop_length=128, win_length=512, power=1, n_iter=1,
                     momentum=1.5, length=None, rand_init=True)


def test_config3_hour_long_transform_round_trip(torch_cuda):
    """Maximum size (A2SB/modelcard.md:50: 1 hour): L = 158,760,000 samples -> T = 310,079 frames on ONE GPU.
    Frame count / length exact, a mid-hour frame range bit-identical to the same range computed from its local
    window (the sharded entry points), oracle parity on that range, DC-retaining round trip >= 100 dB over the hour."""
    torch = torch_cuda
    from audio_intelligence_b200 import _capi, _lib
    n_fft, hop, L = 2048, 512, 3600 * 44100
    g = torch.Generator(device="cuda").manual_seed(1000)
    wav = (0.3 * torch.randn(1, L, generator=g, device="cuda")).clamp_(-1, 1)
    spec = _lib.stft_forward(wav, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, drop_dc=False, power=0.25)
    T = 1 + L // hop
    assert tuple(spec.shape) == (1, 3, 1025, T) and T == 310079 == O.num_frames(L, hop)
    t0, t1 = 150000, 150064
    lo, hi = t0 * hop - n_fft // 2, (t1 - 1) * hop + n_fft // 2
    part = _lib.stft_forward(wav[:, lo:hi].contiguous(), n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, drop_dc=False,
                             power=0.25, total_len=L, sample_first=lo, t_range=(t0, t1))
    assert torch.equal(part, spec[..., t0:t1])
    # oracle on the same local window: frames t0..t1 of the hour are frames 2..2+64 of a clip starting 2 hops earlier
    w = to_np(wav[0, lo - hop * 2: hi + hop * 2])
    ref = O.complex_to_mag_phase(O.stft_complex(w, n_fft, hop))
    ref[0] = O.power_scale(ref[:1], 0.25, None)[0]
    k = (lo + n_fft // 2 - (lo - hop * 2)) // hop        # frame index of t0 inside the local clip
    assert O.mag_rel_err(ref[0][:, k:k + 64] ** 4, to_np(part)[0, 0] ** 4) <= 1e-4
    y = _lib.istft_inverse(spec, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, has_dc=True, phase_fix=True, power=4.0)
    assert tuple(y.shape) == (1, hop * (T - 1)) == (1, O.istft_length(T, hop))
    ref_w = wav[0, : y.shape[1]].double()
    err = y[0].double() - ref_w
    snr = 10 * torch.log10((ref_w * ref_w).sum() / (err * err).sum())
    assert float(snr) >= 100.0, float(snr)


def test_config5_batch64_sampler_loop_with_network_stub(torch_cuda, T, D):
    """BASELINE config 5 per GPU: batch 64 of 10 s clips, forward transform -> mask -> 3 bridge sampling steps with a
    random-init 3->3 channel 3x3 conv stub (multidiffusion windows 256/128, batch_size 16, ot-ode) -> inverse transform.
    Every sampler step is checked against the reference's expressions evaluated with torch ops on the same tensors."""
    torch = torch_cuda
    from audio_intelligence_b200 import _lib
    B, n_fft, hop = 64, 2048, 512
    g = torch.Generator(device="cuda").manual_seed(2000)
    wav = (0.3 * torch.randn(B, 441000, generator=g, device="cuda")).clamp_(-1, 1)
    fwd, inv = chains(T, n_fft, hop)
    x0, _ = T.apply_audio_transforms(wav, fwd)                              # [64, 3, 1024, 862]
    assert tuple(x0.shape) == (B, 3, 1024, 862)
    mask = torch.zeros_like(x0)
    mask[:, :, 185:, :] = 1                                                  # 4 kHz bandwidth-extension mask
    torch.manual_seed(7)
    x1 = x0 * (1 - mask) + mask * torch.randn_like(x0) * 0.5
    conv = torch.nn.Conv2d(3, 3, 3, padding=1).cuda().requires_grad_(False)
    torch.nn.init.normal_(conv.weight, std=0.05, generator=torch.Generator(device="cuda").manual_seed(1))

    def net(x, t_emb):
        return conv(x) + t_emb[:, :1, None, None]

    def t_to_emb(t):
        return torch.stack([t, t * t], dim=1).cuda()
    ts = torch.linspace(1.0, 0.0, 4)[None]
    ddpm = D.Diffusion()
    n0 = _lib.launch_count()
    preds = D.ddpm_sample(net, ddpm, x1, ts, t_to_emb, mask=mask, win_length=256, hop_length=128, batch_size=16,
                          use_ot_ode=True, outputs_to_cpu=False)
    assert _lib.launch_count() - n0 == 2 + 2 * 3 and len(preds) == 3
    # the same three steps with the reference's expressions (torch ops), from the same inputs
    x1p, mp = D.multidiffusion_pad_inputs(x1, 256, 128), D.multidiffusion_pad_inputs(mask, 256, 128)
    assert x1p.shape[-1] == 896
    x_t = x1p.clone()
    for i in range(3):
        t, tp = ts[:, i], ts[:, i + 1]
        vf = D.get_multidiffusion_vf(net, x_t, t_to_emb(t).repeat(B, 1), 256, 128, 16)
        pred = x_t - ddpm.get_std_fwd(t).cuda() * vf
        pred = pred * mp + (1 - mp) * x1p
        assert torch.equal(preds[i], pred[..., :862])
        mu_x0, mu_xt, _ = (c.cuda() for c in ddpm.posterior_coefs(tp, t))
        x_t = (1. - mp) * x1p + mp * (mu_x0 * pred + mu_xt * x_t)
        del vf
    y, _ = T.apply_audio_transforms(preds[-1], inv)
    assert tuple(y.shape) == (B, 440832) and bool(torch.isfinite(y).all())


def test_window_shorter_than_n_fft(torch_cuda, T):
    """ComplexSpectrogram(n_fft=1024, win_length=800, hop_length=256): the centre-padded window, forward and inverse."""
    torch = torch_cuda
    n_fft, win, hop = 1024, 800, 256
    wav = O.synth_noise(9000, 11)
    c = T.ComplexSpectrogram(n_fft, win, hop)(torch.from_numpy(wav).cuda())
    refc = O.stft_complex(wav, n_fft, hop, win_length=win)
    ref = np.stack([refc.real, refc.imag]).astype(np.float32)
    assert tuple(c.shape) == ref.shape and np.abs(to_np(c) - ref).max() <= 2e-6 * np.abs(ref).max()
    y = T.InverseComplexSpectrogram(n_fft, win, hop)(torch.from_numpy(ref).cuda())
    yr = O.istft_complex(refc, n_fft, hop, win_length=win)
    assert tuple(y.shape) == yr.shape and O.snr_db(yr, to_np(y)) >= 100


def test_roundtrip_host_matches_device_path(torch_cuda, T, monkeypatch):
    """a2sb_roundtrip_host (pinned host buffers, internal clip groups of unequal size on three streams) returns exactly
    what the device-resident forward + inverse chain returns, clip for clip."""
    torch = torch_cuda
    from audio_intelligence_b200 import _capi, _lib
    monkeypatch.setenv("A2SB_E2E_GROUP_MB", "8")        # 7 clips per group at this clip length -> 7 groups of 6 or 7
    n_fft, hop, B, L = 2048, 512, 43, 44100
    wav = torch.from_numpy(np.stack([O.synth_noise(L, 1000 + i) for i in range(B)]))
    h_in = wav.pin_memory()
    Tn = 1 + L // hop
    h_out = torch.full((B, hop * (Tn - 1)), float("nan")).pin_memory()
    h_spec = torch.full((B, 3, n_fft // 2, Tn), float("nan")).pin_memory()
    _lib.roundtrip_host(h_in, h_out, n_fft, hop, spec_pinned=h_spec)
    spec = _lib.stft_forward(wav.cuda(), n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25)
    back = _lib.istft_inverse(spec, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, has_dc=False, power=4.0, phase_fix=True)
    assert torch.equal(h_spec, spec.cpu())
    assert torch.equal(h_out, back.cpu())


def test_segment_padding_pipeline_is_bit_identical_and_skips_the_pad_launches(torch_cuda, T, D):
    """transforms.set_segment_padding(256, 128): K1 emits the wrap-padded, hop-aligned width (view tagged with its buffer),
    the corruption transform keeps the layout (filled tensor AND mask padded), multidiffusion_pad_inputs hands the buffers
    out without a launch, and K2 reads the padded sampler output in place.  Everything equals the contiguous pipeline."""
    torch = torch_cuda
    from audio_intelligence_b200 import _lib
    from audio_intelligence_b200.corruption import corruptions as CO
    B, n_fft, hop = 4, 2048, 512
    g = torch.Generator(device="cuda").manual_seed(31)
    wav = (0.3 * torch.randn(B, 441000, generator=g, device="cuda")).clamp_(-1, 1)
    fwd, inv = chains(T, n_fft, hop)
    ref, _ = T.apply_audio_transforms(wav, fwd)
    T.set_segment_padding(256, 128)
    try:
        x0, _ = T.apply_audio_transforms(wav, fwd)
        assert tuple(x0.shape) == tuple(ref.shape) and x0.stride(-2) == 896 and torch.equal(x0, ref)
        buf = _lib.padded_buffer_of(x0, 256, 128, None)
        assert buf is not None and torch.equal(buf, D.multidiffusion_pad_inputs(ref, 256, 128))
        n0 = _lib.launch_count()
        assert D.multidiffusion_pad_inputs(x0, 256, 128) is buf and _lib.launch_count() == n0      # no kernel
        # corruption on the padded view: same values / same mask as on the contiguous tensor, layout kept
        torch.manual_seed(5)
        fill_ref, m_ref = CO._fill(ref, (185, 1024), (0, 862), 0.5)
        torch.manual_seed(5)
        fill_pad, m_pad = CO._fill(x0, (185, 1024), (0, 862), 0.5)
        assert torch.equal(fill_pad, fill_ref) and torch.equal(m_pad, m_ref) and fill_pad.stride(-2) == 896
        assert torch.equal(_lib.padded_buffer_of(fill_pad, 256, 128, None), D.multidiffusion_pad_inputs(fill_ref, 256, 128))
        assert torch.equal(_lib.padded_buffer_of(m_pad, 256, 128, None), D.multidiffusion_pad_inputs(m_ref, 256, 128))
        # sampler: two launches fewer, same predictions; inverse reads the padded prediction in place
        conv = torch.nn.Conv2d(3, 3, 3, padding=1).cuda().requires_grad_(False)
        net = lambda x, t_emb: conv(x) + t_emb[:, :1, None, None]
        t_to_emb = lambda t: torch.stack([t, t * t], dim=1).cuda()
        ts = torch.linspace(1.0, 0.0, 3)[None]
        kw = dict(win_length=256, hop_length=128, batch_size=16, use_ot_ode=True)
        n0 = _lib.launch_count()
        p_ref = D.ddpm_sample(net, D.Diffusion(), fill_ref, ts, t_to_emb, mask=m_ref, **kw)
        n_ref = _lib.launch_count() - n0
        n0 = _lib.launch_count()
        p_pad = D.ddpm_sample(net, D.Diffusion(), fill_pad, ts, t_to_emb, mask=m_pad, **kw)
        assert _lib.launch_count() - n0 == n_ref - 2
        assert all(torch.equal(a, b) for a, b in zip(p_pad, p_ref))
        y_ref, _ = T.apply_audio_transforms(p_ref[-1], inv)
        y_pad, _ = T.apply_audio_transforms(p_pad[-1], inv)          # [..., :862] view of a [.., 896] buffer: read in place
        assert not p_pad[-1].is_contiguous() and torch.equal(y_pad, y_ref)
    finally:
        T.set_segment_padding(None)


# ------------------------------------------------------------------ other STFT consumers


def test_etta_stft_helper_vs_reference_call_fixture(torch_cuda):
    """ETTA STFT (adp.py:1510-1590) with normalized=True on K1 / K2: encode (mag/angle and real/imag) and decode vs the
    reference's torch.stft / torch.istft calls (tests/golden/consumers.npz, oracle/make_golden.py::make_consumers)."""
    torch = torch_cuda
    from audio_intelligence_b200 import stft_consumers as SC
    g = load_golden("consumers.npz")
    wave = torch.from_numpy(g["etta_wave"])
    peak = float(np.abs(g["etta_mag"]).max())
    st = SC.STFT(num_fft=1024, hop_length=256, use_complex=True)
    re, im = st.encode(wave)
    assert tuple(re.shape) == (1, 2, 513, g["etta_real"].shape[-1]) and not re.is_cuda
    assert np.abs(re[0].numpy() - g["etta_real"]).max() <= 2e-6 * peak and np.abs(im[0].numpy() - g["etta_imag"]).max() <= 2e-6 * peak
    sp = SC.STFT(num_fft=1024, hop_length=256)
    mag, ph = sp.encode(wave.cuda())
    assert mag.is_cuda and np.abs(mag[0].cpu().numpy() - g["etta_mag"]).max() <= 2e-6 * peak
    big = g["etta_mag"] > 1e-2 * peak                       # the angle of a bin far below the peak is fp32 noise in torch too
    dphi = np.angle(np.exp(1j * (ph[0].cpu().numpy() - g["etta_phase"])))
    assert np.abs(dphi[big]).max() <= 1e-4
    y = sp.decode(mag, ph)                                  # length = closest_power_2(frames * hop) = t
    assert tuple(y.shape) == (1, 2, g["etta_wave"].shape[-1])
    assert O.snr_db(g["etta_decode"], y[0].cpu().numpy()) >= 100
    assert O.snr_db(g["etta_decode"], st.decode(re, im)[0].numpy()) >= 100
    e1 = sp.encode1d(wave.cuda())
    assert tuple(e1.shape) == (1, 2 * 2 * 513, mag.shape[-1]) and O.snr_db(g["etta_decode"], sp.decode1d(e1)[0].cpu().numpy()) >= 100
    # the reference's DEFAULT num_fft = 1023 (odd: 512 bins) through the any-length kernels, vs torch.stft / torch.istft
    sd = SC.STFT()
    assert sd.num_fft == 1023 and sd.generic
    mag, ph = sd.encode(wave.cuda())
    pk = float(np.abs(g["etta1023_mag"]).max())
    assert tuple(mag.shape) == (1, 2, 512, g["etta1023_mag"].shape[-1])
    assert np.abs(mag[0].cpu().numpy() - g["etta1023_mag"]).max() <= 2e-6 * pk
    big = g["etta1023_mag"] > 1e-2 * pk
    dphi = np.angle(np.exp(1j * (ph[0].cpu().numpy() - g["etta1023_phase"])))
    assert np.abs(dphi[big]).max() <= 1e-4
    y = sd.decode(mag, ph)                                  # length 8192 > hop * (frames - 1) + 1: the tail torch reconstructs too
    assert tuple(y.shape) == (1, 2, 8192) and O.snr_db(g["etta1023_decode"], y[0].cpu().numpy()) >= 100
    sc = SC.STFT(use_complex=True)
    re, im = sc.encode(wave)
    assert np.abs(re[0].numpy() - g["etta1023_real"]).max() <= 2e-6 * pk and np.abs(im[0].numpy() - g["etta1023_imag"]).max() <= 2e-6 * pk
    assert O.snr_db(g["etta1023_decode"], sc.decode(re, im)[0].numpy()) >= 100


@pytest.mark.parametrize("fs,hs,wl", [(1024, 120, 600), (2048, 240, 1200), (512, 50, 240)])
def test_auraloss_stft_vs_reference_call_fixture(torch_cuda, fs, hs, wl):
    """auraloss STFTLoss.stft (auraloss.py:363-381): hops that do not divide n_fft, win_length < n_fft, clamped magnitude."""
    torch = torch_cuda
    from audio_intelligence_b200 import stft_consumers as SC
    g = load_golden("consumers.npz")
    x = torch.from_numpy(g["aura_x"])
    mag, phs = SC.stft_magnitude(x.cuda(), fs, hs, wl, want_phase=True)
    want = g[f"aura_mag_{fs}"]
    assert tuple(mag.shape) == want.shape and np.abs(mag.cpu().numpy() - want).max() <= 2e-6 * want.max()
    big = want > 1e-2 * want.max()
    dphi = np.angle(np.exp(1j * (phs.cpu().numpy() - g[f"aura_phs_{fs}"])))
    assert np.abs(dphi[big]).max() <= 1e-4
    mag_cpu, none = SC.stft_magnitude(x, fs, hs, wl, window=torch.hann_window(wl))
    assert none is None and not mag_cpu.is_cuda and torch.equal(mag_cpu, mag.cpu())
    with pytest.raises(RuntimeError, match="inverse transform needs"):
        from audio_intelligence_b200 import _capi, _lib
        _lib.istft_inverse(torch.zeros(1, 2, fs // 2 + 1, 9, device="cuda"), fs, wl, hs, kind=_capi.KIND_COMPLEX)


def test_torch_library_ops_match_the_module_api_and_capture_into_a_cuda_graph(torch_cuda, T, D):
    """torch.ops.a2sb.* (audio_intelligence_b200/ops.py): same results as the transform-module API, shape propagation
    through FakeTensor, and a warmed-up forward + inverse pair replays from a CUDA graph."""
    torch = torch_cuda
    import audio_intelligence_b200.ops  # noqa: F401  (registers the ops)
    g = torch.Generator(device="cuda").manual_seed(3)
    wav = (0.3 * torch.randn(3, 60000, generator=g, device="cuda")).clamp_(-1, 1)
    fwd, inv = chains(T, 2048, 512)
    spec_ref, _ = T.apply_audio_transforms(wav, fwd)
    y_ref, _ = T.apply_audio_transforms(spec_ref, inv)
    spec = torch.ops.a2sb.stft_fwd(wav, 2048, 512, 0.25, 1e-9)
    y = torch.ops.a2sb.istft_inv(spec, 2048, 512, 4.0, 1e-9, True)
    assert torch.equal(spec, spec_ref) and torch.equal(y, y_ref)
    xp = D.multidiffusion_pad_inputs(spec, 64, 32)
    segs = torch.ops.a2sb.segment_gather(xp, 64, 32)
    assert torch.equal(torch.ops.a2sb.segment_blend(segs, 3, xp.shape[-1], 64, 32), xp)
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        fw = torch.empty(3, 60000, device="cuda")
        fs = torch.ops.a2sb.stft_fwd(fw, 2048, 512, 0.25, 1e-9)
        assert tuple(fs.shape) == tuple(spec.shape)
        assert tuple(torch.ops.a2sb.istft_inv(fs, 2048, 512, 4.0, 1e-9, True).shape) == tuple(y.shape)
    graph = torch.cuda.CUDAGraph()
    static_in = wav.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            torch.ops.a2sb.istft_inv(torch.ops.a2sb.stft_fwd(static_in, 2048, 512, 0.25, 1e-9), 2048, 512, 4.0, 1e-9, True)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(graph):
        static_out = torch.ops.a2sb.istft_inv(torch.ops.a2sb.stft_fwd(static_in, 2048, 512, 0.25, 1e-9), 2048, 512, 4.0, 1e-9, True)
    static_in.copy_(wav.flip(0))
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_out, y_ref.flip(0))


# ------------------------------------------------------------------------------ round 2: wav edges, tile geometries, mirrors


@pytest.mark.parametrize("n_fft", NFFTS)
def test_pcm16_edges(torch_cuda, n_fft):
    """16-bit PCM ingest fused into K1's load is bit-identical to K1 on the decoded float32 samples (oracle.pcm16_decode:
    sample / 32768, what librosa.load hands the reference); PCM egress fused into K2's store equals the oracle's restatement
    of libsndfile's float -> PCM_16 rule (pcm16_encode) applied to K2's own float output, incl. the saturating branches."""
    torch = torch_cuda
    from audio_intelligence_b200 import _capi, _lib
    hop = n_fft // 4
    g = np.random.default_rng(n_fft)
    pcm = g.integers(-32768, 32768, size=(3, 44100 + 2 * hop + 6), dtype=np.int64).astype(np.int16)
    pcm[1, 1000:1000 + 6 * n_fft] = 0
    pcm[0, :4] = [-32768, 32767, 0, -1]
    dec = O.pcm16_decode(pcm)
    kw = dict(kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25)
    a = _lib.stft_forward(torch.from_numpy(pcm).cuda(), n_fft, n_fft, hop, **kw)
    b = _lib.stft_forward(torch.from_numpy(dec).cuda(), n_fft, n_fft, hop, **kw)
    assert torch.equal(a, b)
    if n_fft == 2048:          # the transform-module API takes the PCM tensor as it is (CPU or GPU) and returns float32
        from audio_intelligence_b200.audio_transforms import transforms as TT
        fwd, _ = chains(TT, n_fft, hop)
        s1, _m = TT.apply_audio_transforms(torch.from_numpy(pcm[0]), fwd)
        assert not s1.is_cuda and s1.dtype == torch.float32 and torch.equal(s1, b[0].cpu())
        s2, _m = TT.apply_audio_transforms(torch.from_numpy(pcm[0]).cuda(), fwd[:2])      # a chain K1's PCM form is not built for
        s3, _m = TT.apply_audio_transforms(torch.from_numpy(dec[0]).cuda(), fwd[:2])
        assert torch.equal(s2, s3)
    spec = b.clone()
    spec[0, 0] *= 1.3
    ikw = dict(kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True, power=4.0)
    y = _lib.istft_inverse(spec, n_fft, n_fft, hop, **ikw)
    q = _lib.istft_inverse(spec, n_fft, n_fft, hop, pcm16=True, **ikw)
    assert q.dtype == torch.int16
    qn = to_np(q)
    assert np.array_equal(qn, O.pcm16_encode(to_np(y)))
    assert (qn == 32767).any() and (qn == -32768).any()


def test_roundtrip_host_pcm16_matches_device_path(torch_cuda):
    torch = torch_cuda
    from audio_intelligence_b200 import _capi, _lib
    n_fft, hop, B, L = 2048, 512, 19, 44100
    g = np.random.default_rng(5)
    pcm = torch.from_numpy(g.integers(-20000, 20000, size=(B, L), dtype=np.int64).astype(np.int16))
    h_in = pcm.pin_memory()
    Tn = 1 + L // hop
    h_out = torch.zeros((B, hop * (Tn - 1)), dtype=torch.int16).pin_memory()
    _lib.roundtrip_host(h_in, h_out, n_fft, hop)
    spec = _lib.stft_forward(pcm.cuda(), n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25)
    back = _lib.istft_inverse(spec, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, has_dc=False, power=4.0, phase_fix=True, pcm16=True)
    assert torch.equal(h_out, back.cpu())
    assert int(back.abs().max()) > 1000          # (the shipped chain drops the DC bin of every frame: lossy by design)


def test_mirrored_inverse_writes_every_buffer(torch_cuda):
    """a2sb_istft_inverse_mirrored in peer mode with the 'peers' on the same GPU: every mirror receives exactly the samples
    of the primary output (the multi-GPU use is tests/dist_gpu_check.py / bench.py long_audio.round_trip_peer)."""
    torch = torch_cuda
    from audio_intelligence_b200 import _capi, _lib
    n_fft, hop = 2048, 512
    wav = torch.from_numpy(np.stack([O.synth_noise(50000 + 13, 7), O.synth_noise(50000 + 13, 8)])).cuda()
    spec = _lib.stft_forward(wav, n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25)
    ikw = dict(kind=_capi.KIND_MAGPHASE, has_dc=False, phase_fix=True, power=4.0)
    ref = _lib.istft_inverse(spec, n_fft, n_fft, hop, **ikw)
    mirrors = [torch.zeros_like(ref) for _ in range(3)]
    out = _lib.istft_inverse(spec, n_fft, n_fft, hop, mirrors=[m.data_ptr() for m in mirrors], **ikw)
    assert torch.equal(out, ref) and all(torch.equal(m, ref) for m in mirrors)


@pytest.mark.parametrize("n_fft,hop", [(512, 128), (1024, 256), (1024, 120), (4096, 1024), (4096, 512), (2048, 512)])
def test_forward_tile_geometries_agree_with_the_oracle(torch_cuda, n_fft, hop):
    """Round-2 forward geometries (32-frame wide tiles for n_fft 512 / 1024, two rounds for n_fft 4096) on ragged lengths,
    hops other than n_fft / 4 and digital silence, against the oracle."""
    torch = torch_cuda
    from audio_intelligence_b200 import _capi, _lib
    for L in (n_fft // 2 + 1, 5 * n_fft + 7, 44100 + 3):
        wav = O.synth_noise(L, L % 1000)
        if L > 4 * n_fft:
            wav[n_fft: 3 * n_fft] = 0.0
        spec = to_np(_lib.stft_forward(torch.from_numpy(wav[None]).cuda(), n_fft, n_fft, hop, kind=_capi.KIND_MAGPHASE,
                                       drop_dc=True, power=0.25))[0]
        ref = O.forward_chain(wav, n_fft, hop)
        assert spec.shape == ref.shape
        assert O.mag_rel_err(ref[0] ** 4, spec[0] ** 4) <= 1e-4
        weighted, strong = O.phase_err(ref, spec)
        assert weighted <= 1e-6 and strong <= 1e-5


@pytest.mark.parametrize("n_fft", NFFTS)
def test_corruption_epilogue_equals_the_two_call_path(torch_cuda, T, n_fft):
    """apply_audio_transforms_with_corruption (K1 writing the clean and the corrupted spectrogram from one pass, SURVEY 8f
    rank 2) returns exactly what the reference-shaped two calls return -- target, transformed, mask -- for the same seed,
    with fewer passes over the spectrogram; every mask kind, 1-D and batched audio."""
    torch = torch_cuda
    from audio_intelligence_b200 import _lib
    from audio_intelligence_b200.corruption import corruptions as C
    hop = n_fft // 4
    fwd, _ = chains(T, n_fft, hop)
    wav = torch.from_numpy(np.stack([O.synth_noise(2 * 44100 + 77, 31), O.synth_noise(2 * 44100 + 77, 32)])).cuda()
    wav[1, 5000:5000 + 4 * n_fft] = 0.0
    augs = [C.TimestampedSegmentInpaintMaskTransform(0.5, 1.0, hop, 44100, 0.5),
            C.MultinomialInpaintMaskTransform(0.4, 0.3, 0.3, 0.5, 44100, dict(min_cutoff_freq=2000, max_cutoff_freq=9000),
                                              dict(min_inpainting_frac=0.1, max_inpainting_frac=0.4, is_random=True))]
    for aug in augs:
        # (the reference's InpaintMask unpacks a 3-D shape: the multinomial transform takes un-batched spectrograms only)
        for x in ((wav[0], wav) if aug is augs[0] else (wav[0], wav[1])):
            for seed in range(3):
                torch.manual_seed(seed); np.random.seed(seed)
                target, _m = T.apply_audio_transforms(x, fwd)
                transformed, mask = T.apply_audio_transforms(target, [aug])
                torch.manual_seed(seed); np.random.seed(seed)
                n0 = _lib.launch_count()
                t2, tr2, m2 = T.apply_audio_transforms_with_corruption(x, fwd, [aug])
                assert _lib.launch_count() - n0 == 2               # K1 with the epilogue + the rectangle mask
                assert torch.equal(t2, target) and torch.equal(m2, mask)
                assert torch.equal(tr2, transformed)
                assert torch.equal(torch.signbit(tr2), torch.signbit(transformed))
    # a chain the epilogue is not built for falls back to the two calls
    t3, tr3, m3 = T.apply_audio_transforms_with_corruption(wav[0].cpu(), fwd, [augs[0]])
    assert not t3.is_cuda and tr3.shape == t3.shape and m3.shape == t3.shape
