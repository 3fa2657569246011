"""The oracle (oracle/a2sb_oracle.py) against the golden fixtures produced by the unmodified
reference (oracle/make_golden.py), and against the live reference when it is mounted."""
import os

import numpy as np
import pytest

import a2sb_oracle as O
from conftest import load_golden

NFFTS = (512, 1024, 2048, 4096)


@pytest.mark.parametrize("n_fft", NFFTS)
def test_forward_chain_matches_reference(n_fft):
    g = load_golden(f"chain_n{n_fft}.npz")
    hop = n_fft // 4
    spec = O.forward_chain(g["wav"], n_fft, hop)
    assert spec.shape == g["spec"].shape
    # magnitude gate of north_star: max rel err <= 1e-4 (relative to max(|b|, 1e-3 max|b|))
    assert O.mag_rel_err(g["spec"][0] ** 4, spec[0] ** 4) <= 1e-4
    mag = g["spec"][0] ** 4
    big = mag > 1e-4 * mag.max()
    assert np.abs(spec[1:] - g["spec"][1:])[:, big].max() <= 2e-4   # phase of well-conditioned bins
    c = O.stft_complex(g["wav"], n_fft, hop)
    ref = g["complex_spec"][0] + 1j * g["complex_spec"][1]
    assert np.abs(c - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("n_fft", NFFTS)
def test_inverse_chain_matches_reference(n_fft):
    g = load_golden(f"chain_n{n_fft}.npz")
    hop = n_fft // 4
    for key, kw, src in (("wav_inv", {}, "spec"), ("wav_inv_nosvd", {"svd_fix": False}, "spec"),
                         ("wav_pert", {}, "spec_pert")):
        wav = O.inverse_chain(g[src], n_fft, hop, **kw)
        assert wav.shape == g[key].shape
        assert O.snr_db(g[key], wav) >= 100.0, key
    c = g["complex_spec"][0] + 1j * g["complex_spec"][1]
    assert O.snr_db(g["wav_cplx"], O.istft_complex(c, n_fft, hop)) >= 100.0


def test_tonal_chain():
    g = load_golden("chain_tonal.npz")
    spec = O.forward_chain(g["wav"])
    assert O.mag_rel_err(g["spec"][0] ** 4, spec[0] ** 4) <= 1e-4
    assert O.snr_db(g["wav_inv"], O.inverse_chain(g["spec"])) >= 100.0


def test_standalone_ops():
    g = load_golden("ops.npz")
    msp = g["msp"]
    np.testing.assert_allclose(O.phase_fix(msp), g["svd_fix"], atol=5e-6)
    assert tuple(O.phase_fix(msp)[1:, 0, 0]) == (1.0, 0.0)          # degenerate -> (1, 0)
    assert tuple(O.phase_fix(msp)[1:, 0, 1]) == (-1.0, 0.0)
    c = O.mag_phase_to_complex(msp)
    np.testing.assert_array_equal(np.stack([c.real, c.imag]), g["to_complex"])
    np.testing.assert_allclose(O.complex_to_mag_phase(msp[0] + 1j * msp[1]), g["to_magphase"], atol=2e-6)
    np.testing.assert_allclose(O.power_scale(msp, 0.5), g["pow_half_all"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(O.power_scale(msp, 0.25, [0]), g["pow_quarter_c0"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(O.power_scale(msp, 4, [0]), g["pow_four_c0"], rtol=2e-6, atol=1e-7)
    np.testing.assert_array_equal(O.add_dc(msp), g["add_dc"])
    np.testing.assert_array_equal(O.drop_dc(msp), g["drop_dc"])
    assert O.power_scale(np.zeros((1, 2, 2), np.float32), 0.25)[0, 0, 0] == 0.0


def test_segment_blend_bit_exact():
    g = load_golden("blend.npz")
    x, win, hop = g["x"], 64, 32
    np.testing.assert_array_equal(O.multidiffusion_pad_inputs(x, win, hop), g["xp"])
    np.testing.assert_array_equal(O.multidiffusion_pad_inputs(x, win, hop, 0), g["xp_const"])
    xp = g["xp"]
    t = np.zeros((2, 4), np.float32)
    np.testing.assert_array_equal(O.get_multidiffusion_vf(lambda a, e: a, xp, t, win, hop, 5), g["ident"])
    np.testing.assert_array_equal(g["ident"], xp)                  # identity network returns its input
    two, tenth = np.float32(2), np.float32(0.1)
    np.testing.assert_array_equal(O.get_multidiffusion_vf(lambda a, e: a * two + tenth, xp, t, win, hop, 5), g["affine"])
    ramp = lambda a, e: a + (np.arange(a.shape[0], dtype=np.float32) * np.float32(0.001)).reshape(-1, 1, 1, 1)
    np.testing.assert_array_equal(O.get_multidiffusion_vf(ramp, xp, t, win, hop, 1000), g["ramp"])
    f = lambda a, e: a * np.float32(1.7) - np.float32(0.3)
    np.testing.assert_array_equal(O.get_multidiffusion_vf(f, g["xp3"], np.zeros((1, 4), np.float32), 48, 16, 3), g["noisy3"])


def test_known_answers(known_answers):
    ka = known_answers
    for w, padded in ka["pad_widths"].items():
        assert O.multidiffusion_pad_width(int(w), 256, 128) == padded
    assert ka["pad_widths"]["862"] == 896 and ka["pad_widths"]["310079"] == 310144 and ka["pad_widths"]["100"] == 200
    assert O.num_hops(310144, 256, 128) == 2422
    for n_fft, (T, out_len) in ka["frames"].items():
        hop = int(n_fft) // 4
        assert O.num_frames(441000, hop) == T and O.istft_length(T, hop) == out_len
    assert ka["frames"]["2048"] == [862, 440832]
    row = np.ones(896)
    for a, b in ((86, 103), (318, 344), (800, 896)):
        row[a:b] = 0
    mids = O.find_middle_of_zero_segments(row)
    assert mids == ka["zero_segment_centres"] == [94, 330, 847]
    assert [list(O.inpaint_window(c, 256, 896)) for c in mids] == ka["inpaint_windows"]
    for n_fft, r in ka["upsample_first_row"].items():
        assert O.upsample_mask_first_row(int(n_fft) // 2, 4000) == r
    assert ka["upsample_first_row"]["2048"] == 185
    assert list(O.inpaint_frames(1.0, 1.2)) == ka["inpaint_frames_1.0_1.2"] == [86, 103]
    w2 = O.hann_window(2048) ** 2
    assert abs(sum(w2[1024 - 512 * m] for m in range(0, 3)) - ka["envelope"]["first_kept"]) < 1e-6
    assert abs(sum(w2[m * 512] for m in range(4)) - ka["envelope"]["interior"]) < 1e-6


def test_fp64_restatement_is_self_consistent():
    """stft -> istft in float64 reproduces the interior of the signal to ~1e-13 (bounds the
    reference's own fp32 error)."""
    wav = O.synth_noise(9000, 3).astype(np.float64)
    s = O.stft_complex(wav, 1024, 256, dtype=np.float64)
    y = O.istft_complex(s, 1024, 256, dtype=np.float64)
    assert np.abs(y - wav[: y.shape[0]]).max() < 1e-12


def test_short_input_raises_like_torch():
    with pytest.raises(RuntimeError, match="Padding size"):
        O.stft_complex(np.zeros(1024, np.float32), 2048, 512)


@pytest.mark.skipif(not os.path.isdir("/root/reference/A2SB"), reason="reference not mounted")
def test_oracle_against_live_reference():
    import torch
    import make_golden
    T, D, U, C = make_golden.import_reference()
    wav = O.synth_noise(7000, 77)
    fwd, inv, _ = make_golden.chains(T, 1024, 256)
    spec, _ = T.apply_audio_transforms(torch.from_numpy(wav), fwd)
    mine = O.forward_chain(wav, 1024, 256)
    assert O.mag_rel_err(spec[0].numpy() ** 4, mine[0] ** 4) <= 1e-4
    ref_wav, _ = T.apply_audio_transforms(spec, inv)
    assert O.snr_db(ref_wav.numpy(), O.inverse_chain(spec.numpy(), 1024, 256)) >= 100.0


@pytest.mark.parametrize("n_fft", NFFTS)
def test_torch_port_matches_reference_fixture(n_fft):
    """oracle/torch_port.py (bench.py's CPU baseline) reproduces the reference's outputs."""
    import torch
    import torch_port as P
    g = load_golden(f"chain_n{n_fft}.npz")
    hop = n_fft // 4
    spec = P.forward_chain(torch.from_numpy(g["wav"]), n_fft, hop).numpy()
    assert spec.shape == g["spec"].shape
    np.testing.assert_allclose(spec, g["spec"], rtol=1e-5, atol=1e-5)
    for key, kw, src in (("wav_inv", {}, "spec"), ("wav_inv_nosvd", {"svd_fix": False}, "spec"),
                         ("wav_pert", {}, "spec_pert")):
        wav = P.inverse_chain(torch.from_numpy(g[src]), n_fft, hop, **kw).numpy()
        assert wav.shape == g[key].shape and O.snr_db(g[key], wav) >= 120.0, key


def test_mask_fixtures_from_reference():
    """tests/golden/masks.npz (reference corruption transforms + find_middle_of_zero_segments, seeded)."""
    g = load_golden("masks.npz")
    spec = g["spec"]
    for seed in range(6):
        row = g[f"zero_row_{seed}"].astype(np.float32)
        assert O.find_middle_of_zero_segments(row) == g[f"zero_mid_{seed}"].tolist()
        m = g[f"timestamped_{seed}_mask"].astype(np.float32)
        assert np.array_equal(m, O.rect_mask(spec.shape, (0, 64), O.inpaint_frames(0.1, 0.35)))
        for name in ("timestamped", "multinomial"):
            m = g[f"{name}_{seed}_mask"].astype(np.float32)
            filled = g[f"{name}_{seed}_filled"]
            # every mask is one rectangle; outside it the input is untouched, inside it noise * level
            rows = np.where(m[0].any(axis=1))[0]
            cols = np.where(m[0].any(axis=0))[0]
            if rows.size:
                assert np.array_equal(m, O.rect_mask(spec.shape, (rows[0], rows[-1] + 1), (cols[0], cols[-1] + 1)))
            noise = np.where(m == 1, filled / np.float32(0.5), 0).astype(np.float32)
            assert np.array_equal(O.mask_with_noise(spec, m, noise, 0.5), filled)
        up = g[f"upsample_{seed}"]
        first = int(np.nonzero(up[0, :, 0])[0][0])
        assert O.upsample_mask_first_row(128, 1000) <= first < max(O.upsample_mask_first_row(128, 9000), first + 1)
        assert up[:, first:, :].all() and not up[:, :first, :].any()


def _np_sampler_stubs():
    def t_to_emb(t):
        t = np.asarray(t, np.float32)
        return np.stack([t, t * t], axis=1)

    def net(x, t_emb):
        pos = np.arange(x.shape[-1], dtype=np.float32)
        return (((x * np.float32(0.8)).astype(np.float32) + (t_emb[:, :1, None, None] * np.float32(0.1)).astype(np.float32)).astype(np.float32)
                + (pos * np.float32(0.01)).astype(np.float32)).astype(np.float32)
    return t_to_emb, net


def test_bridge_schedule_and_sampler_vs_reference_fixture():
    """tests/golden/sampler.npz: the reference's Diffusion scalars and a 4-step ot-ode ddpm_sample run."""
    g = load_golden("sampler.npz")
    t = g["t"]
    assert np.array_equal(O.int_beta_0_t(t), g["int_beta"])
    assert np.array_equal(O.std_fwd(t), g["std_fwd"])
    assert np.array_equal(O.std_rev(t), g["std_rev"])
    np.testing.assert_array_equal(O.std_t(t), g["std_t"])        # NaN at t = 0 and 1 (0/0) on both sides
    ts = g["t_steps"]
    for i in range(ts.shape[1] - 1):
        want = g["posterior_coefs"][i][:, 0]
        got = [c[0] for c in O.posterior_coefs(ts[:, i + 1], ts[:, i])]
        assert np.array_equal(np.array(got, np.float32), want)
    t_to_emb, net = _np_sampler_stubs()
    for tag, mp in (("mp1", True), ("mp0", False)):
        preds, states = O.ddpm_sample(net, t_to_emb, g["x_1"], ts, g["mask"].astype(np.float32), mp, 64, 32, 4)
        assert np.array_equal(np.stack(preds), g[f"pred_{tag}"])
        assert np.array_equal(np.stack(states), g[f"state_{tag}"])


def test_fast_inpaint_sampler_vs_reference_fixture():
    """tests/golden/fast_inpaint.npz: fast_inpaint_ddpm_sample restated around the reference's own functions."""
    g = load_golden("fast_inpaint.npz")
    t_to_emb, net = _np_sampler_stubs()
    x, windows = O.fast_inpaint_ddpm_sample(net, t_to_emb, g["x_1"], g["t_steps"], g["mask"].astype(np.float32), True, 32, 32, 4)
    assert windows == [tuple(w) for w in g["windows"].tolist()] == [(29, 61), (125, 157)]
    assert np.array_equal(x, g["result"])


def test_short_window_padding_matches_torch():
    """win_length < n_fft: torch.stft / torch.istft centre-pad the window to n_fft (torch/functional.py:508 `stft`);
    the oracle's `padded_window` restates that and is pinned here against torch itself."""
    torch = pytest.importorskip("torch")
    n_fft, win, hop = 1024, 800, 256
    wav = O.synth_noise(9000, 11)
    w = torch.hann_window(win)
    ref = torch.stft(torch.from_numpy(wav), n_fft, hop_length=hop, win_length=win, window=w, center=True, pad_mode="reflect",
                     normalized=False, onesided=True, return_complex=True).numpy()
    got = O.stft_complex(wav, n_fft, hop, win_length=win)
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    y_ref = torch.istft(torch.from_numpy(ref), n_fft, hop_length=hop, win_length=win, window=w).numpy()
    y = O.istft_complex(ref, n_fft, hop, win_length=win)
    assert y.shape == y_ref.shape and O.snr_db(y_ref, y) >= 100
