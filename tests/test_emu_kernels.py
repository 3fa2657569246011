"""The real kernel sources (audio_intelligence_b200/csrc), compiled by g++ against the CUDA
execution-model emulator in tests/emu, checked against the reference's golden fixtures and the
oracle.  This is the no-GPU half of the parity suite: it proves index math, barrier structure,
shared-memory layout and the host-side validation of the C ABI; tests/test_parity_gpu.py repeats
the same checks on the sm_100a build."""
import numpy as np
import pytest

import a2sb_oracle as O
from conftest import load_golden

NFFTS = (512, 1024, 2048, 4096)


@pytest.fixture(scope="module")
def plans(emu):
    ps = {n: emu.plan(n, n // 4) for n in NFFTS}
    yield ps
    for p in ps.values():
        emu.destroy(p)


@pytest.mark.parametrize("n_fft", NFFTS)
def test_forward_matches_reference_fixture(emu, plans, n_fft):
    g = load_golden(f"chain_n{n_fft}.npz")
    hop = n_fft // 4
    spec = emu.forward(plans[n_fft], g["wav"][None], n_fft, hop)[0]
    assert spec.shape == g["spec"].shape
    assert O.mag_rel_err(g["spec"][0] ** 4, spec[0] ** 4) <= 1e-4
    weighted, strong = O.phase_err(g["spec"], spec)
    assert weighted <= 1e-6 and strong <= 1e-5         # complex error vs the peak; phases of bins >= 1% of it
    assert np.abs((spec[1] ** 2 + spec[2] ** 2) - 1).max() <= 1e-5          # unit phasors everywhere
    c = emu.forward(plans[n_fft], g["wav"][None], n_fft, hop, kind=0, drop_dc=0, power_on=0)[0]
    assert c.shape == g["complex_spec"].shape
    assert np.abs(c - g["complex_spec"]).max() <= 2e-6 * np.abs(g["complex_spec"]).max()


@pytest.mark.parametrize("n_fft", NFFTS)
def test_inverse_matches_reference_fixture(emu, plans, n_fft):
    g = load_golden(f"chain_n{n_fft}.npz")
    hop = n_fft // 4
    p = plans[n_fft]
    assert O.snr_db(g["wav_inv"], emu.inverse(p, g["spec"][None], n_fft, hop)[0]) >= 100
    assert O.snr_db(g["wav_pert"], emu.inverse(p, g["spec_pert"][None], n_fft, hop)[0]) >= 100
    assert O.snr_db(g["wav_inv_nosvd"], emu.inverse(p, g["spec"][None], n_fft, hop, phase_fix=0)[0]) >= 100
    w = emu.inverse(p, g["complex_spec"][None], n_fft, hop, kind=0, has_dc=1, phase_fix=0, power_on=0)[0]
    assert w.shape == g["wav_cplx"].shape and O.snr_db(g["wav_cplx"], w) >= 100


def test_tonal_fixture(emu, plans):
    g = load_golden("chain_tonal.npz")
    spec = emu.forward(plans[2048], g["wav"][None], 2048, 512)[0]
    assert O.mag_rel_err(g["spec"][0] ** 4, spec[0] ** 4) <= 1e-4
    assert O.snr_db(g["wav_inv"], emu.inverse(plans[2048], g["spec"][None], 2048, 512)[0]) >= 100


@pytest.mark.parametrize("L", [257, 1000, 1024, 1027, 5000, 16 * 256 * 3 + 1])
def test_ragged_lengths_and_batch(emu, plans, L):
    """T = 1 + L // hop for lengths around tile and hop boundaries; batch of 3 unequal seeds."""
    n_fft, hop = 512, 128
    wav = np.stack([O.synth_noise(L, 1000 + i) for i in range(3)])
    spec = emu.forward(plans[n_fft], wav, n_fft, hop)
    assert spec.shape == (3, 3, n_fft // 2, 1 + L // hop)
    assert not np.isnan(spec).any()
    for i in range(3):
        ref = O.forward_chain(wav[i], n_fft, hop)
        assert O.mag_rel_err(ref[0] ** 4, spec[i, 0] ** 4) <= 1e-4
    y = emu.inverse(plans[n_fft], spec, n_fft, hop)
    assert y.shape == (3, hop * (L // hop)) and not np.isnan(y).any()
    for i in range(3):
        assert O.snr_db(O.inverse_chain(spec[i], n_fft, hop), y[i]) >= 100


def test_dc_retaining_round_trip_snr(emu, plans):
    """STFT -> mag/phase -> power .25 -> power 4 -> complex -> iSTFT with the DC row kept: >= 100 dB."""
    n_fft, hop = 1024, 256
    wav = O.synth_noise(9000, 42)[None]
    spec = emu.forward(plans[n_fft], wav, n_fft, hop, drop_dc=0)
    y = emu.inverse(plans[n_fft], spec, n_fft, hop, has_dc=1)[0]
    assert O.snr_db(wav[0, : y.shape[0]], y) >= 100


def test_edge_signals(emu, plans):
    n_fft, hop = 512, 128
    L = 3000
    zeros = np.zeros((1, L), np.float32)
    spec = emu.forward(plans[n_fft], zeros, n_fft, hop)[0]
    assert (spec[0] == 0).all() and (spec[1] == 1).all() and (spec[2] == 0).all()   # atan2(0,0)=0 -> (1,0)
    assert (emu.inverse(plans[n_fft], spec[None], n_fft, hop) == 0).all()
    for pos in (0, L - 1):
        imp = zeros.copy()
        imp[0, pos] = 1.0
        s = emu.forward(plans[n_fft], imp, n_fft, hop)[0]
        ref = O.forward_chain(imp[0], n_fft, hop)
        assert O.mag_rel_err(ref[0] ** 4, s[0] ** 4) <= 1e-4
    dc = np.full((1, L), 0.5, np.float32)
    s = emu.forward(plans[n_fft], dc, n_fft, hop, drop_dc=0)[0]
    ref = O.power_scale(O.complex_to_mag_phase(O.stft_complex(dc[0], n_fft, hop)), 0.25, [0])
    assert O.mag_rel_err(ref[0] ** 4, s[0] ** 4) <= 1e-4


@pytest.mark.parametrize("n_fft", NFFTS)
def test_sharded_ranges_are_bit_identical(emu, plans, n_fft):
    """A frame range / output range computed from a local window equals the unsharded result
    bit for bit (the contract the multi-GPU planner relies on) -- for every kernel family (32-frame wide tiles, runs of two
    tiles, two rounds, pair-split inverse), ranges that start and end inside tiles, a pitched output and a batch of two."""
    hop = n_fft // 4
    L = 70 * hop + n_fft + 57
    wav = np.stack([O.synth_noise(L, 9), O.synth_noise(L, 10)])
    full = emu.forward(plans[n_fft], wav, n_fft, hop)
    T = full.shape[-1]
    for t0, t1 in ((37, 59), (0, 33), (40, T)):
        lo = max(t0 * hop - n_fft // 2, 0)
        hi = min((t1 - 1) * hop + n_fft // 2, L)
        part = emu.forward(plans[n_fft], wav[:, lo:hi], n_fft, hop, t_range=(t0, t1), sample_first=lo, total_len=L)
        np.testing.assert_array_equal(part, full[..., t0:t1])
    pitched = emu.forward(plans[n_fft], wav[:, :hi], n_fft, hop, t_range=(3, 41), sample_first=0, total_len=L, pitch=48)
    np.testing.assert_array_equal(pitched[..., :38], full[..., 3:41])
    y = emu.inverse(plans[n_fft], full, n_fft, hop)
    for o0, on in ((40 * hop, 17 * hop), (0, 9 * hop), (33 * hop, (T - 1 - 33) * hop)):
        f_lo = max((o0 + n_fft // 2) // hop - 3, 0)
        f_hi = min((o0 + on + n_fft // 2 + hop - 1) // hop, T)
        yp = emu.inverse(plans[n_fft], np.ascontiguousarray(full[..., f_lo:f_hi]), n_fft, hop, n_frames=T,
                         spec_t_first=f_lo, out_range=(o0, on))
        np.testing.assert_array_equal(yp, y[:, o0:o0 + on])


def test_standalone_ops(emu):
    g = load_golden("ops.npz")
    msp = g["msp"]
    np.testing.assert_allclose(emu.pointwise(2, msp, 3), g["svd_fix"], atol=5e-6)
    np.testing.assert_array_equal(emu.pointwise(1, msp, 2), g["to_complex"])
    np.testing.assert_allclose(emu.pointwise(0, msp[:2], 3), g["to_magphase"], atol=2e-6)
    np.testing.assert_allclose(emu.pointwise(3, msp, 3, 0xFFFFFFFF, 0.5), g["pow_half_all"], rtol=5e-6, atol=1e-7)
    np.testing.assert_allclose(emu.pointwise(3, msp, 3, 1, 0.25), g["pow_quarter_c0"], rtol=5e-6, atol=1e-7)
    np.testing.assert_allclose(emu.pointwise(3, msp, 3, 1, 4.0), g["pow_four_c0"], rtol=5e-6, atol=1e-7)


def test_segments_bit_exact(emu):
    g = load_golden("blend.npz")
    x, xp = g["x"], g["xp"]
    np.testing.assert_array_equal(emu.wrap_pad(x, xp.shape[-1]), xp)
    np.testing.assert_array_equal(emu.wrap_pad(x, xp.shape[-1], 0.0), g["xp_const"])
    for w_in, w_out in ((37, 51), (37, 52), (40, 52), (7, 8)):       # scalar and 128-bit store paths, odd source rows
        y = np.arange(2 * 3 * w_in, dtype=np.float32).reshape(2, 3, w_in)
        np.testing.assert_array_equal(emu.wrap_pad(y, w_out), np.concatenate([y, y[..., :w_out - w_in]], -1))
    segs = emu.gather(xp, 64, 32)
    np.testing.assert_array_equal(segs, O.segment_gather(xp, 64, 32))
    np.testing.assert_array_equal(emu.blend(segs, 2, xp.shape[-1], 64, 32), g["ident"])
    np.testing.assert_array_equal(emu.blend(segs * np.float32(2) + np.float32(0.1), 2, xp.shape[-1], 64, 32), g["affine"])
    ramp = segs + (np.arange(segs.shape[0], dtype=np.float32) * np.float32(0.001)).reshape(-1, 1, 1, 1)
    np.testing.assert_array_equal(emu.blend(ramp, 2, xp.shape[-1], 64, 32), g["ramp"])
    s3 = emu.gather(g["xp3"], 48, 16)
    np.testing.assert_array_equal(emu.blend(s3 * np.float32(1.7) - np.float32(0.3), 1, g["xp3"].shape[-1], 48, 16),
                                  g["noisy3"])
    # unaligned geometry takes the scalar kernels
    xo = np.random.default_rng(1).standard_normal((1, 2, 3, 61)).astype(np.float32)
    so = emu.gather(xo, 21, 10)
    np.testing.assert_array_equal(so, O.segment_gather(xo, 21, 10))
    np.testing.assert_array_equal(emu.blend(so, 1, 61, 21, 10), O.segment_blend(so, 1, 61, 21, 10))


def test_error_behaviour(emu, plans):
    """Same conditions the reference raises through torch: short input (reflect pad), bad ranges."""
    with pytest.raises(emu.capi.A2SBError, match="Padding size should be less than"):
        emu.forward(plans[2048], np.zeros((1, 1024), np.float32), 2048, 512)
    with pytest.raises(emu.capi.A2SBError):
        emu.plan(1000, 250)
    with pytest.raises(emu.capi.A2SBError):
        emu.plan(1024, 301)                     # odd hop: not even a forward-only plan
    p300 = emu.plan(1024, 300)                  # even hop that does not divide n_fft: forward-only plan (round 2)
    try:
        with pytest.raises(emu.capi.A2SBError, match="inverse transform needs"):
            emu.inverse(p300, np.zeros((1, 3, 512, 9), np.float32), 1024, 300)
    finally:
        emu.destroy(p300)
    # a rectangular window of 1 sample violates NOLA exactly like torch.istft's check
    w = np.zeros(512, np.float32)
    w[0] = 1.0
    p = emu.plan(512, 128, window=w)
    try:
        with pytest.raises(emu.capi.A2SBError, match="window overlap add min"):
            emu.inverse(p, np.zeros((1, 3, 256, 9), np.float32), 512, 128)
    finally:
        emu.destroy(p)


@pytest.mark.parametrize("n_fft,hop", [(512, 64), (512, 256), (1024, 128), (2048, 1024), (4096, 2048), (4096, 512)])
def test_hops_other_than_quarter_window(emu, n_fft, hop):
    """hop != n_fft/4 takes the run-time-hop overlap-add path (A2SB itself always uses n_fft/4)."""
    L = 20 * hop + 37
    wav = O.synth_noise(L, 7)
    p = emu.plan(n_fft, hop)
    try:
        refc = O.stft_complex(wav, n_fft, hop)
        ref = np.stack([refc.real, refc.imag]).astype(np.float32)
        if (n_fft, hop) == (4096, 2048):
            # the forward kernel stages a tile's input span in shared memory: 7 * 2048 + 4096 samples next to the 128 KB
            # exchange do not fit 227 KB -- rejected with a message (the inverse kernel below has no such limit)
            with pytest.raises(emu.capi.A2SBError, match="does not fit the forward kernel"):
                emu.forward(p, wav[None], n_fft, hop, kind=0, drop_dc=0, power_on=0)
            y = emu.inverse(p, ref[None], n_fft, hop, kind=0, has_dc=1, phase_fix=0, power_on=0)[0]
            assert O.snr_db(O.istft_complex(refc, n_fft, hop), y) >= 100
            return
        c = emu.forward(p, wav[None], n_fft, hop, kind=0, drop_dc=0, power_on=0)[0]
        assert c.shape == ref.shape and np.abs(c - ref).max() <= 2e-6 * np.abs(ref).max()
        y = emu.inverse(p, ref[None], n_fft, hop, kind=0, has_dc=1, phase_fix=0, power_on=0)[0]
        yr = O.istft_complex(refc, n_fft, hop)
        assert y.shape == yr.shape and O.snr_db(yr, y) >= 100
        spec = emu.forward(p, wav[None], n_fft, hop)
        assert O.snr_db(O.inverse_chain(spec[0], n_fft, hop), emu.inverse(p, spec, n_fft, hop)[0]) >= 100
    finally:
        emu.destroy(p)


# ------------------------------------------------------------------ masks / zero segments (rows M1, B4)


def test_zero_segment_windows_known_answer(emu, known_answers):
    row = np.ones(896, np.float32)
    for a, b in ((86, 103), (318, 344), (800, 896)):
        row[a:b] = 0
    centres, lr, k = emu.zero_segment_windows(row, 256)
    assert k == 3 and centres.tolist() == known_answers["zero_segment_centres"] == [94, 330, 847]
    assert lr.tolist() == known_answers["inpaint_windows"]


@pytest.mark.parametrize("seed", range(6))
def test_zero_segment_windows_random_rows(emu, seed):
    """Random run structure incl. runs touching both ends, length-1 runs, all-zero / all-one rows, odd windows."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 5000))
    row = (rng.random(n) < rng.choice([0.02, 0.5, 0.98])).astype(np.float32)
    if seed == 4:
        row[:] = 0
    if seed == 5:
        row[:] = 1
    win = int(rng.integers(1, max(2, n)))
    centres, lr, k = emu.zero_segment_windows(row, win)
    want = O.find_middle_of_zero_segments(row)
    assert k == len(want) and centres.tolist() == want
    assert lr.tolist() == [list(O.inpaint_window(c, win, n)) for c in want]


@pytest.mark.parametrize("width", [57, 64])          # scalar and 128-bit paths
def test_masks_and_noise_fill_bit_exact(emu, width):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 3, 40, width)).astype(np.float32)
    noise = rng.standard_normal(x.shape).astype(np.float32)
    # bounds follow python slice semantics like the reference's mask[:, a:b, c:d] = 1 (negative = from the end)
    for rows_range, cols_range in (((7, 40), (0, width)), ((0, 40), (11, 30)), ((0, 40), (0, 0)), ((5, 900), (-3, 21)),
                                   ((-12, -2), (-30, -4))):
        m = np.zeros_like(x)
        m[:, :, rows_range[0]:rows_range[1], cols_range[0]:cols_range[1]] = 1
        want = (x * (1 - m) + m * noise * np.float32(0.5)).astype(np.float32)
        assert np.array_equal(emu.rect_mask(x.shape, rows_range, cols_range), m)
        out, mask = emu.mask_fill(x, noise, rows_range, cols_range, 0.5)
        assert np.array_equal(mask, m) and np.array_equal(out, want)
        assert np.array_equal(emu.mask_with_noise(x, m, noise, 0.5), want)
        out2, none = emu.mask_fill(x, noise, rows_range, cols_range, 0.5, want_mask=False)
        assert none is None and np.array_equal(out2, want)


def test_zero_segment_fixtures_from_reference(emu):
    g = load_golden("masks.npz")
    for seed in range(6):
        row = g[f"zero_row_{seed}"].astype(np.float32)
        centres, _, k = emu.zero_segment_windows(row, 16)
        assert k == len(g[f"zero_mid_{seed}"]) and centres.tolist() == g[f"zero_mid_{seed}"].tolist()


@pytest.mark.parametrize("W,win,hop", [(160, 64, 32), (96, 64, 64), (131, 50, 27)])
def test_blend_step_bit_exact_vs_oracle(emu, W, win, hop):
    """K4s = blend + get_pred_x0 + mask merge + p_posterior + known-region re-imposition (vector and scalar paths)."""
    rng = np.random.default_rng(W)
    b, c, h = 2, 3, 4
    W = ((W - win) // hop) * hop + win if W > win else win        # widths the pad produces
    L = O.num_hops(W, win, hop)
    segs = rng.standard_normal((b * L, c, h, win)).astype(np.float32)
    x_t, x_1, n1, n2 = (rng.standard_normal((b, c, h, W)).astype(np.float32) for _ in range(4))
    mask = (rng.random((b, c, h, W)) < 0.3).astype(np.float32)
    vf = O.segment_blend(segs, b, W, win, hop)
    sc = dict(std_fwd_t=0.123, mu_x0=0.4, mu_xt=0.6, sd_post=0.05, std_sb=0.07)
    for kw in (dict(mask=None), dict(mask=mask), dict(mask=mask, mask_pred_x0=False),
               dict(mask=mask, noise_post=n1, noise_mask=n2), dict(mask=None, noise_post=n1)):
        kw = {**sc, **kw}
        want = O.sampler_step(vf, x_t, x_1, **kw)
        got = emu.blend_step(segs, b, W, win, hop, x_t, x_1, **kw)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


def test_row_pitched_output_and_input(emu, plans):
    """out_pitch > T: same values in a pitched buffer, padding untouched; K2 reads the pitched buffer in place."""
    g = load_golden("chain_n1024.npz")
    n_fft, hop = 1024, 256
    ref = emu.forward(plans[n_fft], g["wav"][None], n_fft, hop)
    T = ref.shape[-1]
    pitch = -(-T // 8) * 8 + 8
    out = emu.forward(plans[n_fft], g["wav"][None], n_fft, hop, pitch=pitch)
    assert out.shape[-1] == pitch and np.array_equal(out[..., :T], ref) and np.isnan(out[..., T:]).all()
    y_ref = emu.inverse(plans[n_fft], ref, n_fft, hop)
    y = emu.inverse(plans[n_fft], np.nan_to_num(out, nan=7.0), n_fft, hop, n_frames=T)
    assert np.array_equal(y, y_ref)


def test_window_shorter_than_n_fft(emu):
    """win_length < n_fft (the classes take it as a separate argument, transforms.py:84-96): centre-padded window."""
    n_fft, win, hop = 1024, 800, 256
    wav = O.synth_noise(9000, 11)
    p = emu.plan(n_fft, hop, win_length=win)
    try:
        c = emu.forward(p, wav[None], n_fft, hop, kind=0, drop_dc=0, power_on=0)[0]
        refc = O.stft_complex(wav, n_fft, hop, win_length=win)
        ref = np.stack([refc.real, refc.imag]).astype(np.float32)
        assert c.shape == ref.shape and np.abs(c - ref).max() <= 2e-6 * np.abs(ref).max()
        y = emu.inverse(p, ref[None], n_fft, hop, kind=0, has_dc=1, phase_fix=0, power_on=0)[0]
        yr = O.istft_complex(refc, n_fft, hop, win_length=win)
        assert y.shape == yr.shape and O.snr_db(yr, y) >= 100
    finally:
        emu.destroy(p)


# ------------------------------------------------------------------ fused padding / windowed blend (round 2)


@pytest.mark.parametrize("n_fft,win,hop_s", [(512, 64, 32), (1024, 32, 32), (2048, 40, 16), (4096, 40, 16)])
def test_forward_emits_the_segment_padding(emu, n_fft, win, hop_s):
    """a2sb_fwd_args.wrap_cols: K1 writes the first frames again behind column T -- the result equals
    multidiffusion_pad_inputs (diffusion.py:67-83) of the contiguous spectrogram, bit for bit."""
    hop = n_fft // 4
    L = 47 * hop + 5
    wav = np.stack([O.synth_noise(L, 3), O.synth_tonal(L)])
    wav[1, :n_fft] = 0                                     # digital silence at the head: the careful path is re-emitted too
    p = emu.plan(n_fft, hop)
    try:
        ref = emu.forward(p, wav, n_fft, hop)
        T = ref.shape[-1]
        want = O.multidiffusion_pad_inputs(ref, win, hop_s)
        got = emu.forward(p, wav, n_fft, hop, pitch=want.shape[-1], wrap_cols=want.shape[-1] - T)
        assert want.shape[-1] > T and np.array_equal(got, want)
        with pytest.raises(emu.capi.A2SBError):             # the padding needs room in the rows
            emu.forward(p, wav, n_fft, hop, pitch=T + 1, wrap_cols=2)
    finally:
        emu.destroy(p)


@pytest.mark.parametrize("width,pitch", [(150, 160), (151, 153)])       # 64-bit and scalar paths
def test_mask_fill_padded_equals_fill_then_pad(emu, width, pitch):
    rng = np.random.default_rng(11)
    win, hop_s = 64, 32
    xbuf = rng.standard_normal((2, 3, 12, pitch)).astype(np.float32)
    x = xbuf[..., :width]
    noise = rng.standard_normal(x.shape).astype(np.float32)
    filled, m = emu.mask_fill(np.ascontiguousarray(x), noise, (4, 12), (30, 75), 0.5)
    want_x, want_m = O.multidiffusion_pad_inputs(filled, win, hop_s), O.multidiffusion_pad_inputs(m, win, hop_s)
    got_x, got_m = emu.mask_fill_padded(xbuf, width, noise, (4, 12), (30, 75), 0.5, want_x.shape[-1])
    assert np.array_equal(got_x, want_x) and np.array_equal(got_m, want_m)


def test_blend_window_equals_slice_of_full_blend(emu):
    rng = np.random.default_rng(12)
    win, hop_s, n = 64, 32, 9
    segs = rng.standard_normal((n, 3, 8, win)).astype(np.float32)
    W = (n - 1) * hop_s + win
    full = emu.blend(segs, 1, W, win, hop_s)
    for off, cnt, pitch in ((0, W, W), (32, 128, 160), (20, 50, 57)):      # 128-bit and scalar paths
        got = emu.blend_window(segs, 1, W, win, hop_s, off, cnt, pitch)
        assert np.array_equal(got[..., :cnt], full[..., off:off + cnt])
        assert np.isnan(got[..., cnt:]).all()


def test_other_stft_consumers_vs_reference_call_fixture(emu):
    """ETTA's normalized STFT helper and auraloss' STFT (tests/golden/consumers.npz): a plan whose window is
    hann / sqrt(n_fft) IS torch's normalized=True in both directions; forward-only plans take hops that do not divide n_fft;
    pointwise ops 4 / 5 are (|X|, angle X) and polar -> complex."""
    g = load_golden("consumers.npz")
    n_fft, hop = 1024, 256
    w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)).astype(np.float32) * np.float32(n_fft ** -0.5)
    p = emu.plan(n_fft, hop, window=w)
    try:
        wave = g["etta_wave"][0]
        c = emu.forward(p, wave, n_fft, hop, kind=0, drop_dc=0, power_on=0)
        peak = np.abs(g["etta_mag"]).max()
        assert np.abs(c[:, 0] - g["etta_real"]).max() <= 2e-6 * peak and np.abs(c[:, 1] - g["etta_imag"]).max() <= 2e-6 * peak
        y = emu.inverse(p, c, n_fft, hop, kind=0, has_dc=1, phase_fix=0, power_on=0)
        assert y.shape == g["etta_decode"].shape and O.snr_db(g["etta_decode"], y) >= 100
        mp = emu.pointwise(4, c[0], 2, eps=0.0)
        assert np.abs(mp[0] - g["etta_mag"][0]).max() <= 2e-6 * peak
        back = emu.pointwise(5, mp, 2)
        assert np.abs(back - c[0]).max() <= 4e-6 * peak
    finally:
        emu.destroy(p)
    x = g["aura_x"]
    for fs, hs, wl in ((1024, 120, 600), (512, 50, 240)):
        win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(wl) / wl)).astype(np.float32)
        p = emu.plan(fs, hs, win_length=wl, window=win)
        try:
            c = emu.forward(p, x, fs, hs, kind=0, drop_dc=0, power_on=0)
            mag = emu.pointwise(4, c[0], 2, eps=1e-8)[0]
            want = g[f"aura_mag_{fs}"][0]
            assert mag.shape == want.shape and np.abs(mag - want).max() <= 2e-6 * want.max()
            with pytest.raises(emu.capi.A2SBError, match="inverse transform needs"):
                emu.inverse(p, c, fs, hs, kind=0, has_dc=1, phase_fix=0, power_on=0)
        finally:
            emu.destroy(p)


@pytest.mark.parametrize("n_fft", NFFTS)
def test_pcm16_edges(emu, plans, n_fft):
    """16-bit PCM ingest fused into K1's load: bit-identical to K1 on the decoded float32 samples (the decode's 2^-15 is a
    power of two carried by the window).  PCM egress fused into K2's store: the oracle's restatement of libsndfile's
    float -> PCM_16 rule applied to K2's own float output, sample for sample (full-scale + clipping cases included)."""
    hop = n_fft // 4
    g = np.random.default_rng(n_fft)
    pcm = g.integers(-32768, 32768, size=(2, 9 * n_fft + 2 * hop + 6), dtype=np.int64).astype(np.int16)
    pcm[1, 100:100 + 3 * n_fft] = 0                                   # digital silence: the careful path
    pcm[0, :8] = [-32768, 32767, 0, 1, -1, 16384, -16384, 255]
    dec = O.pcm16_decode(pcm)
    assert dec.dtype == np.float32 and dec.min() == -1.0 and dec.max() < 1.0
    a = emu.forward(plans[n_fft], pcm, n_fft, hop)
    b = emu.forward(plans[n_fft], dec, n_fft, hop)
    assert np.array_equal(a, b)
    spec = b.copy()
    spec[0, 0] *= 1.3                                                 # louder than full scale: the saturating branches
    y = emu.inverse(plans[n_fft], spec, n_fft, hop)
    q = emu.inverse(plans[n_fft], spec, n_fft, hop, pcm16=True)
    assert q.dtype == np.int16 and q.shape == y.shape
    assert np.array_equal(q, O.pcm16_encode(y))
    assert (q == 32767).any() and (q == -32768).any()


@pytest.mark.parametrize("n_fft", NFFTS)
def test_corruption_epilogue(emu, plans, n_fft):
    """K1 with the corruption in its epilogue: the clean output is K1's, the corrupted one equals the oracle's
    mask_with_noise(rect mask) of it -- including the sign of zeros outside the mask, which the reference's expression takes
    from the noise sample -- for masks in the interior, at the edges, empty and negative-indexed."""
    hop = n_fft // 4
    wav = np.stack([O.synth_noise(6 * n_fft + 3 * hop + 5, 77), O.synth_noise(6 * n_fft + 3 * hop + 5, 78)])
    wav[1, hop: hop + 3 * n_fft] = 0.0                         # digital silence: the careful path, and exact zeros
    clean0 = emu.forward(plans[n_fft], wav, n_fft, hop)
    g = np.random.default_rng(n_fft)
    noise = g.standard_normal(clean0.shape).astype(np.float32)
    rows, T = clean0.shape[-2:]
    for rr, cc in (((rows // 3, rows), (0, T)), ((0, rows), (5, 11)), ((7, 7), (0, T)), ((-9, rows), (-6, -1)), ((0, rows), (0, T))):
        clean, corrupted = emu.forward(plans[n_fft], wav, n_fft, hop, corrupt=(noise, rr, cc, 0.5))
        assert np.array_equal(clean, clean0)
        mask = O.rect_mask(clean0.shape, rr, cc)
        ref = O.mask_with_noise(clean0, mask, noise, 0.5)
        assert np.array_equal(corrupted, ref)
        assert np.array_equal(np.signbit(corrupted), np.signbit(ref))


@pytest.mark.parametrize("n_fft,hop,win", [(1023, 256, 1023), (96, 24, 80), (45, 7, 45)])
def test_any_length_transform_vs_oracle_and_reference_fixture(emu, n_fft, hop, win):
    """The any-length kernels (ETTA's STFT helper defaults to num_fft = 1023): forward and inverse against the oracle's
    torch.stft / torch.istft restatement for odd, even non-power-of-two and tiny lengths; n_fft 1023 also against the
    reference-generated fixture (torch.stft / istft with normalized=True, length = 8192 > hop * (frames - 1))."""
    L = 8192 if n_fft == 1023 else 40 * hop + 11
    g = load_golden("consumers.npz")
    wav = g["etta_wave"][0] if n_fft == 1023 else np.stack([O.synth_noise(L, 5), O.synth_tonal(L)])
    n = np.arange(win)
    window = (0.5 - 0.5 * np.cos(2 * np.pi * n / win)).astype(np.float32)          # torch.hann_window (periodic)
    wpad = O.padded_window(n_fft, win, window.astype(np.float64), np.float64)
    wn = (wpad / np.sqrt(n_fft)).astype(np.float32)                                # normalized=True rides on the window
    spec = emu.dft_generic_forward(wav, n_fft, hop, wn)
    for b in range(wav.shape[0]):
        ref = O.stft_any_length(wav[b], n_fft, hop, window, normalized=True)
        got = spec[b, 0] + 1j * spec[b, 1]
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    length = L if n_fft == 1023 else hop * (spec.shape[-1] - 1) + 3
    y = emu.dft_generic_inverse(spec, n_fft, hop, wn, length)
    for b in range(wav.shape[0]):
        ref = O.istft_any_length(spec[b, 0] + 1j * spec[b, 1], n_fft, hop, window, length, normalized=True)
        assert O.snr_db(ref, y[b]) >= 100
    if n_fft == 1023:
        assert np.abs(spec[:, 0] - g["etta1023_real"]).max() <= 2e-6 * np.abs(g["etta1023_real"]).max()
        assert np.abs(spec[:, 1] - g["etta1023_imag"]).max() <= 2e-6 * np.abs(g["etta1023_real"]).max()
        yr = emu.dft_generic_inverse(np.stack([g["etta1023_real"], g["etta1023_imag"]], 1), n_fft, hop, wn, 8192)
        assert all(O.snr_db(g["etta1023_decode"][b], yr[b]) >= 100 for b in range(2))


def test_any_length_transform_random_geometries(emu):
    """Seeded random (n_fft, hop, win_length, length) -- odd / even / prime lengths, hops that do not divide n_fft, windows
    shorter than n_fft, output lengths short of and beyond hop * (frames - 1) -- against the oracle's torch.stft / istft
    restatement, without normalisation."""
    rng = np.random.default_rng(20261019)
    for _ in range(6):
        n_fft = int(rng.integers(9, 200))
        hop = int(rng.integers(1, max(2, n_fft // 2)))
        win = int(rng.integers(max(2, n_fft // 2), n_fft + 1))
        L = int(rng.integers(n_fft, 6 * n_fft))
        wav = O.synth_noise(L, int(rng.integers(1 << 30)))[None]
        window = (0.54 - 0.46 * np.cos(2 * np.pi * (np.arange(win) + 0.5) / win)).astype(np.float32)      # strictly positive
        wpad = O.padded_window(n_fft, win, window.astype(np.float64), np.float64).astype(np.float32)
        spec = emu.dft_generic_forward(wav, n_fft, hop, wpad)
        ref = O.stft_any_length(wav[0], n_fft, hop, window)
        got = spec[0, 0] + 1j * spec[0, 1]
        assert got.shape == ref.shape, (n_fft, hop, win, L)
        assert np.abs(got - ref).max() <= 3e-6 * max(np.abs(ref).max(), 1e-6), (n_fft, hop, win, L)
        T = spec.shape[-1]
        natural = n_fft + hop * (T - 1) - 2 * (n_fft // 2)
        for length in (max(1, natural - 3), natural + n_fft // 3):
            y = emu.dft_generic_inverse(spec, n_fft, hop, wpad, length)[0]
            yr = O.istft_any_length(got, n_fft, hop, window, length)
            covered = np.arange(length) + n_fft // 2 < n_fft + hop * (T - 1)
            ok = covered & (np.abs(yr) < 1e3)               # (a window shorter than the hop leaves samples no frame reaches)
            assert np.abs(y[ok] - yr[ok]).max() <= 2e-4 * max(np.abs(yr[ok]).max(), 1e-6), (n_fft, hop, win, L, length)
            assert np.all(y[~covered] == 0)
