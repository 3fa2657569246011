"""CPU oracle for the A2SB spectral-transform hot path -- TEST INFRASTRUCTURE ONLY.

This module is a plain numpy restatement of what the reference computes on this path.  It is the
checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; nothing under
audio_intelligence_b200/ imports it and the product path never routes through it.

Where the arithmetic lives.  The reference's own code (paths relative to the reference root,
A2SB/...) is thin glue over third-party kernels that are NOT in the reference tree:
  * torchaudio.transforms.Spectrogram / InverseSpectrogram -> torch.stft / torch.istft
    (torch 2.2.2 + unpinned torchaudio per A2SB/DockerFile:1,9; torch/torchaudio 2.11.0 here);
  * torch.linalg.svd for SVDFixMagInstPhase; torch.nn.Unfold for the segment windowing.
Their published algorithms are restated below (reflect pad -> frame -> window -> rFFT; irFFT ->
window -> overlap-add / window-envelope -> centre trim; polar projection of a 2x2 matrix; im2col).

Pinning.  The reference ships no tests or golden vectors for this path, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/make_golden.py imports the
unmodified reference modules from /root/reference/A2SB (with a 3-line jsonargparse stub), runs them
on seeded inputs and commits the results under tests/golden/; tests/test_oracle.py checks every
function here against those fixtures (and against the live reference when it is mounted).

All functions take/return numpy arrays.  `dtype=np.float32` reproduces the reference's precision
class; `dtype=np.float64` is the high-precision restatement used to separate "our error" from
"torch's own fp32 error".
"""
from __future__ import annotations

from math import ceil

import numpy as np

# --------------------------------------------------------------------------------------------------
# index arithmetic (bit-exact)
# --------------------------------------------------------------------------------------------------


def num_frames(length: int, hop: int) -> int:
    """torch.stft(center=True): T = 1 + L // hop  (A2SB/audio_transforms/transforms.py:91-96,103)."""
    return 1 + length // hop


def istft_length(n_frames: int, hop: int) -> int:
    """torch.istft(length=None, center=True): hop * (T - 1)  (transforms.py:171-175,184)."""
    return hop * (n_frames - 1)


def reflect_index(i: np.ndarray, length: int) -> np.ndarray:
    """Source index of F.pad(mode='reflect') for padded position i - pad (no edge repeat)."""
    i = np.abs(i)
    return np.where(i >= length, 2 * (length - 1) - i, i)


def hann_window(win_length: int, dtype=np.float64) -> np.ndarray:
    """torch.hann_window(win_length) (periodic): 0.5 - 0.5 cos(2 pi n / win_length)."""
    n = np.arange(win_length, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)).astype(dtype)


def padded_window(n_fft: int, win_length: int, window: np.ndarray | None = None, dtype=np.float64) -> np.ndarray:
    """torch.stft centre-pads a short window to n_fft: left = (n_fft - win_length) // 2."""
    w = hann_window(win_length, dtype) if window is None else np.asarray(window, dtype=dtype)
    out = np.zeros(n_fft, dtype=dtype)
    left = (n_fft - win_length) // 2
    out[left:left + win_length] = w
    return out


# --------------------------------------------------------------------------------------------------
# forward chain  (transforms.py:83-118, 187-219)
# --------------------------------------------------------------------------------------------------


def stft_complex(wav: np.ndarray, n_fft: int, hop: int, win_length: int | None = None, window=None,
                 dtype=np.float32) -> np.ndarray:
    """ComplexSpectrogram.__call__ (transforms.py:98-105): complex [n_fft/2+1, T].

    X[k, t] = sum_n w[n] * xp[t*hop + n] * exp(-2 pi i k n / n_fft), xp = reflect_pad(x, n_fft/2).
    """
    wav = np.asarray(wav, dtype=dtype)
    assert wav.ndim == 1, wav.shape  # transforms.py:102
    L = wav.shape[0]
    if L <= n_fft // 2:
        raise RuntimeError("Padding size should be less than the corresponding input dimension")
    win_length = n_fft if win_length is None else win_length
    w = padded_window(n_fft, win_length, window, dtype)
    T = num_frames(L, hop)
    idx = (np.arange(T)[:, None] * hop + np.arange(n_fft)[None, :]) - n_fft // 2
    frames = wav[reflect_index(idx, L)] * w[None, :]
    cdt = np.complex64 if dtype == np.float32 else np.complex128
    return np.fft.rfft(frames.astype(np.float64), axis=1).T.astype(cdt)


def stft_any_length(wav: np.ndarray, n_fft: int, hop: int, window: np.ndarray, normalized: bool = False) -> np.ndarray:
    """torch.stft(center=True, pad_mode="reflect", onesided) for ANY n_fft (ETTA's STFT helper uses 1023,
    ETTA/stable_audio_tools/models/adp.py:1510-1550): complex [n_fft//2 + 1, T], T = 1 + (L + 2 (n_fft//2) - n_fft) // hop."""
    x = np.asarray(wav, np.float64)
    L, pad = x.shape[0], n_fft // 2
    w = padded_window(n_fft, len(window), np.asarray(window, np.float64), np.float64)
    T = 1 + (L + 2 * pad - n_fft) // hop
    idx = (np.arange(T)[:, None] * hop + np.arange(n_fft)[None, :]) - pad
    frames = x[reflect_index(idx, L)] * w[None, :]
    X = np.fft.rfft(frames, n=n_fft, axis=1).T
    return X / np.sqrt(n_fft) if normalized else X


def istft_any_length(spec: np.ndarray, n_fft: int, hop: int, window: np.ndarray, length: int, normalized: bool = False) -> np.ndarray:
    """torch.istft(center=True, length=length) for ANY n_fft (adp.py:1570-1586): overlap-added irfft(X) * w over the
    overlap-added w^2, trimmed by n_fft//2 at the head; zero past the last frame."""
    K, T = spec.shape
    assert K == n_fft // 2 + 1
    w = padded_window(n_fft, len(window), np.asarray(window, np.float64), np.float64)
    X = np.asarray(spec, np.complex128) * (np.sqrt(n_fft) if normalized else 1.0)
    frames = np.fft.irfft(X.T, n=n_fft, axis=1) * w[None, :]
    total = n_fft + hop * (T - 1)
    y, env = np.zeros(max(total, n_fft // 2 + length)), np.zeros(max(total, n_fft // 2 + length))
    for t in range(T):
        y[t * hop:t * hop + n_fft] += frames[t]
        env[t * hop:t * hop + n_fft] += w * w
    a = n_fft // 2
    y, env = y[a:a + length], env[a:a + length]
    return np.where(env > 0, y / np.where(env > 0, env, 1.0), 0.0)


def complex_to_mag_phase(spec: np.ndarray) -> np.ndarray:
    """ComplexToMagInstPhase (transforms.py:108-118): [mag, cos(atan2), sin(atan2)] as [3, F, T]."""
    re, im = spec.real, spec.imag
    dt = re.dtype
    mag = np.sqrt(re * re + im * im)
    ph = np.arctan2(im, re)
    return np.stack([mag, np.cos(ph), np.sin(ph)]).astype(dt)


def drop_dc(spec: np.ndarray) -> np.ndarray:
    """SpectrogramDropDCTerm (transforms.py:214-219)."""
    return spec[..., 1:, :]


def add_dc(spec: np.ndarray) -> np.ndarray:
    """SpectrogramAddDCTerm (transforms.py:222-228): prepend row0 * 0."""
    return np.concatenate([spec[..., :1, :] * 0, spec], axis=-2)


def power_scale(spec: np.ndarray, power: float, channels=None, eps: float = 1e-9) -> np.ndarray:
    """PowerScaleSpectrogram (transforms.py:187-207): spec * |spec|^p / (|spec| + eps) on `channels`."""
    dt = spec.dtype
    a = np.abs(spec)
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = (a ** dt.type(power) / (a + dt.type(eps))).astype(dt)
    if channels is None:
        return (spec * scale).astype(dt)
    out = spec.copy()
    ch = list(channels)
    out[ch] = spec[ch] * scale[ch]
    return out


def forward_chain(wav: np.ndarray, n_fft=2048, hop=512, power=0.25, eps=1e-9, dtype=np.float32) -> np.ndarray:
    """The shipped forward chain (configs/ensemble_2split_sampling.yaml:105-119) -> [3, n_fft/2, T]."""
    s = stft_complex(wav, n_fft, hop, dtype=dtype)
    return power_scale(drop_dc(complex_to_mag_phase(s)), power, [0], eps)


# --------------------------------------------------------------------------------------------------
# inverse chain  (transforms.py:121-184, 187-228)
# --------------------------------------------------------------------------------------------------


def phase_fix(msp: np.ndarray) -> np.ndarray:
    """SVDFixMagInstPhase (transforms.py:135-160) in closed form.

    R = [[c, -s], [s, c]] = sqrt(c^2+s^2) * Rot(theta); U diag(1, det(U Vh)) Vh is the polar factor
    Rot(theta), so the first column is (c, s)/sqrt(c^2+s^2); the zero matrix maps to the identity,
    i.e. (0, 0) -> (1, 0).
    """
    dt = msp.dtype
    c, s = msp[1].astype(np.float64), msp[2].astype(np.float64)
    n = np.sqrt(c * c + s * s)
    with np.errstate(divide="ignore", invalid="ignore"):
        cn = np.where(n > 0, c / n, 1.0)
        sn = np.where(n > 0, s / n, 0.0)
    return np.stack([msp[0], cn.astype(dt), sn.astype(dt)])


def mag_phase_to_complex(msp: np.ndarray) -> np.ndarray:
    """MagInstPhaseToComplex (transforms.py:121-132) -> complex [F, T]."""
    cdt = np.complex64 if msp.dtype == np.float32 else np.complex128
    return (msp[0] * msp[1] + 1j * (msp[0] * msp[2])).astype(cdt)


def istft_complex(spec: np.ndarray, n_fft: int, hop: int, win_length: int | None = None, window=None,
                  dtype=np.float32) -> np.ndarray:
    """InverseComplexSpectrogram.__call__ (transforms.py:177-184) via torch.istft semantics.

    f_t = irfft(X[:, t], n_fft) * w (imaginary parts of the DC and Nyquist bins are ignored);
    y = OLA(f_t) / OLA(w^2), trimmed to [n_fft/2, n_fft/2 + hop*(T-1)).  Raises like torch when the
    trimmed envelope dips below 1e-11.
    """
    F, T = spec.shape
    assert F == n_fft // 2 + 1, spec.shape
    win_length = n_fft if win_length is None else win_length
    w = padded_window(n_fft, win_length, window, np.float64)
    frames = np.fft.irfft(spec.T.astype(np.complex128), n=n_fft, axis=1) * w[None, :]
    total = n_fft + hop * (T - 1)
    y = np.zeros(total, dtype=np.float64)
    env = np.zeros(total, dtype=np.float64)
    w2 = w * w
    for t in range(T):
        y[t * hop:t * hop + n_fft] += frames[t]
        env[t * hop:t * hop + n_fft] += w2
    a, b = n_fft // 2, n_fft // 2 + istft_length(T, hop)
    y, env = y[a:b], env[a:b]
    if env.size and np.abs(env).min() < 1e-11:
        raise RuntimeError("window overlap add min: 1")
    return (y / env).astype(dtype)


def inverse_chain(spec: np.ndarray, n_fft=2048, hop=512, power=4.0, eps=1e-9, svd_fix=True,
                  dtype=np.float32) -> np.ndarray:
    """The shipped inverse chain (configs/ensemble_2split_sampling.yaml:63-78) -> wav [hop*(T-1)]."""
    s = np.asarray(spec, dtype=dtype)
    assert s.ndim == 3
    s = add_dc(power_scale(s, power, [0], eps))
    if svd_fix:
        s = phase_fix(s)
    return istft_complex(mag_phase_to_complex(s), n_fft, hop, dtype=dtype)


# --------------------------------------------------------------------------------------------------
# segment windowing / blend  (A2SB/diffusion.py:27-87)
# --------------------------------------------------------------------------------------------------


def multidiffusion_pad_width(width: int, win_length: int, hop_length: int) -> int:
    """Padded width produced by multidiffusion_pad_inputs (diffusion.py:67-83), including the
    reference's short-input behaviour: the pad is a slice of the head, so it cannot exceed `width`."""
    if width <= win_length:
        to_pad = win_length - width
    else:
        to_pad = ceil((width - win_length) / hop_length) * hop_length + win_length - width
    return width + min(to_pad, width) if to_pad > 0 else width


def multidiffusion_pad_inputs(x: np.ndarray, win_length: int, hop_length: int, padding_constant=None) -> np.ndarray:
    """diffusion.py:67-83: pad the last axis by copying the head of the signal (or a constant)."""
    width = x.shape[-1]
    if width <= win_length:
        to_pad = win_length - width
    else:
        to_pad = ceil((width - win_length) / hop_length) * hop_length + win_length - width
    if to_pad > 0:
        pad = x[..., :to_pad]
        if padding_constant is not None:
            pad = pad * 0 + padding_constant
        return np.concatenate([x, pad], axis=-1)
    return x.copy()


def multidiffusion_unpad_outputs(x: np.ndarray, original_width: int) -> np.ndarray:
    """diffusion.py:86-87."""
    return x[..., :original_width]


def num_hops(width: int, win_length: int, hop_length: int) -> int:
    """diffusion.py:33."""
    return (width - (win_length - hop_length)) // hop_length


def segment_gather(x: np.ndarray, win_length: int, hop_length: int) -> np.ndarray:
    """diffusion.py:35-42: Unfold([h, win], stride=hop) + "b (c h w) l -> (b l) c h w"."""
    b = x.shape[0]
    L = num_hops(x.shape[-1], win_length, hop_length)
    segs = [x[bi, ..., l * hop_length:l * hop_length + win_length] for bi in range(b) for l in range(L)]
    return np.stack(segs) if segs else np.zeros((0,) + x.shape[1:-1] + (win_length,), x.dtype)


def segment_blend(segs: np.ndarray, batch: int, width: int, win_length: int, hop_length: int) -> np.ndarray:
    """diffusion.py:51-64: sequential `vf_t[l:r] += seg_l; counts[l:r] += 1` then vf_t / counts."""
    L = num_hops(width, win_length, hop_length)
    v = segs.reshape((batch, L) + segs.shape[1:])
    out = np.zeros((batch,) + segs.shape[1:-1] + (width,), dtype=segs.dtype)
    counts = np.zeros_like(out)
    for l in range(L):
        out[..., l * hop_length:l * hop_length + win_length] += v[:, l]
        counts[..., l * hop_length:l * hop_length + win_length] += 1
    with np.errstate(divide="ignore", invalid="ignore"):
        return out / counts


def get_multidiffusion_vf(vf_model, x_t: np.ndarray, t_emb: np.ndarray, win_length=256, hop_length=128,
                          batch_size=16) -> np.ndarray:
    """diffusion.py:27-64 with torch.chunk's chunk sizing (ceil(n / num_chunks) rows per chunk)."""
    b = x_t.shape[0]
    L = num_hops(x_t.shape[-1], win_length, hop_length)
    segs = segment_gather(x_t, win_length, hop_length)
    n = segs.shape[0]
    n_chunks = ceil(n / batch_size)
    rows = ceil(n / n_chunks)
    t_rpt = np.tile(t_emb, (L, 1))
    outs = [vf_model(segs[i:i + rows], t_rpt[i:i + rows]) for i in range(0, n, rows)]
    return segment_blend(np.concatenate(outs, 0), b, x_t.shape[-1], win_length, hop_length)


# --------------------------------------------------------------------------------------------------
# adjacent integer logic
# --------------------------------------------------------------------------------------------------


def find_middle_of_zero_segments(row: np.ndarray) -> list[int]:
    """A2SB/utils.py:54-81: centres ((start+end)/2 truncated) of the zero runs of a 1-D array."""
    is_zero = np.concatenate([[0], (np.asarray(row) == 0).astype(np.int8), [0]])
    d = np.diff(is_zero)
    starts, ends = np.where(d == 1)[0], np.where(d == -1)[0]
    return [int((s + (e - 1)) / 2) for s, e in zip(starts, ends)]  # inclusive end


def inpaint_window(center: int, win_length: int, width: int) -> tuple[int, int]:
    """A2SB/A2SB_lightning_module.py:161-174: [l, r) of length win_length clamped into [0, width]."""
    l, r = int(center - win_length / 2), int(center + win_length / 2)
    if l < 0:
        r -= l
        l = 0
    if r > width:
        l -= r - width
        r = width
    return l, r


def upsample_mask_first_row(n_rows: int, cutoff_hz: float, sr: int = 44100) -> int:
    """A2SB/corruption/corruptions.py:26-51: n_fft = 2*rows (DC dropped); first masked row int(n_fft*f/sr)."""
    return int(2 * n_rows * cutoff_hz / sr)


def inpaint_frames(t0: float, t1: float, hop: int = 512, sr: int = 44100) -> tuple[int, int]:
    """A2SB/corruption/corruptions.py:147-160: frames [int(sr/hop*t0), int(sr/hop*t1))."""
    return int(sr / hop * t0), int(sr / hop * t1)


# --------------------------------------------------------------------------------------------------
# synthetic inputs and error metrics
# --------------------------------------------------------------------------------------------------


def pcm16_decode(pcm: np.ndarray) -> np.ndarray:
    """16-bit PCM -> float32 as libsndfile / soundfile.read(dtype='float32') / librosa.load hand it to the reference
    (A2SB/datasets/datasets.py:231): sample / 32768, exact in fp32 (libsndfile pcm.c, s2f_array: normfact = 1 / 0x8000)."""
    return (np.asarray(pcm, np.int16).astype(np.float32) / np.float32(32768.0)).astype(np.float32)


def pcm16_encode(x: np.ndarray) -> np.ndarray:
    """float32 -> 16-bit PCM as soundfile.write stores the reconstructed audio in a WAV file
    (A2SB/inference/A2SB_inpaint_dataset.py:126; default subtype PCM_16, files opened with SFC_SET_CLIPPING):
    libsndfile pcm.c, f2s_clip_array with normalisation: scaled = x * 2^31 in fp32; >= 2^31 - 1 -> 0x7FFF,
    <= -2^31 -> -0x8000, else lrintf(scaled) >> 16.
    PARITY UNPINNED: soundfile / libsndfile are not installed in the build container and /root/reference holds neither
    them nor a PCM fixture; this restates the published libsndfile source."""
    sc = np.asarray(x, np.float32) * np.float32(2147483648.0)
    out = np.zeros(sc.shape, np.int64)
    hi, lo = sc >= np.float32(2147483647.0), sc <= np.float32(-2147483648.0)
    mid = ~(hi | lo) & ~np.isnan(sc)
    out[mid] = np.rint(sc[mid].astype(np.float64)).astype(np.int64) >> 16     # rint: ties to even, like lrintf
    out[hi], out[lo] = 0x7FFF, -0x8000
    return out.astype(np.int16)


def synth_noise(length: int, seed: int) -> np.ndarray:
    """Broadband: 0.3 * N(0,1) clamped to [-1, 1] (numpy PCG64 stream; seed = 1000 + clip index)."""
    g = np.random.default_rng(seed)
    return np.clip(0.3 * g.standard_normal(length), -1.0, 1.0).astype(np.float32)


def synth_tonal(length: int, sr: int = 44100) -> np.ndarray:
    t = np.arange(length, dtype=np.float64) / sr
    return (0.3 * np.sin(2 * np.pi * 440.0 * t) + 0.2 * np.sin(2 * np.pi * 3000.0 * t + 1.0)).astype(np.float32)


def snr_db(ref: np.ndarray, test: np.ndarray) -> float:
    ref = np.asarray(ref, np.float64)
    err = np.asarray(test, np.float64) - ref
    den = float(np.sum(err * err))
    if den == 0.0:
        return float("inf")
    return 10.0 * np.log10(float(np.sum(ref * ref)) / den)


def mag_rel_err(ref_mag: np.ndarray, test_mag: np.ndarray) -> float:
    """max |a-b| / max(|b|, 1e-3 max|b|)  (SURVEY.md section 8d parity gate)."""
    ref = np.asarray(ref_mag, np.float64)
    floor = 1e-3 * np.abs(ref).max() if ref.size else 0.0
    den = np.maximum(np.abs(ref), max(floor, 1e-30))
    return float((np.abs(np.asarray(test_mag, np.float64) - ref) / den).max()) if ref.size else 0.0


def phase_err(ref_spec: np.ndarray, test_spec: np.ndarray, power: float = 0.25) -> tuple[float, float]:
    """Phase-channel parity of an A2SB spectrogram [3, rows, T] (channels mag^power, cos, sin).

    Returns (weighted, strong):
      weighted = max |d(cos, sin)| * |X| / max|X|   -- the error of the complex value, relative to the
                 spectrum peak (a bin of magnitude 1e-4 max carries 1e4 x the phase noise of a peak bin, in
                 torch's own fp32 path as much as in ours, so the raw phase difference is not a fixed gate);
      strong   = max |d(cos, sin)| over bins with |X| >= 1e-2 max|X|   (<= 1e-5).
    """
    mag = np.abs(np.asarray(ref_spec[0], np.float64)) ** (1.0 / power)
    peak = mag.max() if mag.size else 1.0
    d = np.abs(np.asarray(test_spec[1:], np.float64) - np.asarray(ref_spec[1:], np.float64))
    weighted = float((d * (mag / peak)).max()) if mag.size else 0.0
    strong_mask = mag >= 1e-2 * peak
    strong = float(d[:, strong_mask].max()) if strong_mask.any() else 0.0
    return weighted, strong


# --------------------------------------------------------------------------------------------------
# corruption masks and noise fill  (A2SB/corruption/corruptions.py:14-51,120-160)
# --------------------------------------------------------------------------------------------------


def rect_mask(shape, rows_range, cols_range) -> np.ndarray:
    """Every reference mask is `zeros; mask[:, r0:r1, c0:c1] = 1` (python slice semantics)."""
    m = np.zeros(shape, np.float32)
    m[..., slice(*rows_range), slice(*cols_range)] = 1
    return m


def mask_with_noise(x: np.ndarray, mask: np.ndarray, noise: np.ndarray, level: float) -> np.ndarray:
    """corruptions.py:14-15 with the noise tensor made explicit: x*(1-mask) + mask*noise*level in fp32."""
    x, mask, noise = (np.asarray(a, np.float32) for a in (x, mask, noise))
    return (x * (np.float32(1) - mask) + mask * noise * np.float32(level)).astype(np.float32)


# --------------------------------------------------------------------------------------------------
# bridge schedule and the reverse sampling step  (A2SB/diffusion.py:91-168,
# A2SB/A2SB_lightning_module.py:103-146).  fp32 throughout, evaluated in
# the reference's operation order so that results are bit-identical to its torch ops.
# --------------------------------------------------------------------------------------------------

_F = np.float32


def int_beta_0_t(t, beta_max: float = 0.3) -> np.ndarray:
    """Diffusion.get_int_beta_0_t (diffusion.py:115-123); x**3 as x*x*x like torch's pow(x, 3)."""
    t = np.asarray(t, _F)
    third = _F(1 / 3 * beta_max)
    whole = _F(2 * beta_max * (0.5 ** 3) / 3)
    one = _F(1)
    u = one - t
    return np.where(t > _F(0.5), whole - third * (u * u * u), third * (t * t * t)).astype(_F)


def std_fwd(t, beta_max: float = 0.3) -> np.ndarray:
    return np.sqrt(int_beta_0_t(t, beta_max)).astype(_F)


def std_rev(t, beta_max: float = 0.3) -> np.ndarray:
    return np.sqrt(int_beta_0_t(_F(1) - np.asarray(t, _F), beta_max)).astype(_F)


def gaussian_product_coef(s1, s2):
    """compute_gaussian_product_coef (diffusion.py:91-99)."""
    a, b = (np.asarray(s1, _F) ** 2).astype(_F), (np.asarray(s2, _F) ** 2).astype(_F)
    den = (a + b).astype(_F)
    return (b / den).astype(_F), (a / den).astype(_F), ((a * b).astype(_F) / den).astype(_F)


def std_t(t, beta_max: float = 0.3) -> np.ndarray:
    return np.sqrt(gaussian_product_coef(std_fwd(t, beta_max), std_rev(t, beta_max))[2]).astype(_F)


def posterior_coefs(t_prev, t, beta_max: float = 0.3):
    """(mu_x0, mu_xt, var) of Diffusion.p_posterior (diffusion.py:153-158)."""
    st, sp = std_fwd(t, beta_max), std_fwd(t_prev, beta_max)
    delta = np.sqrt(((st ** 2).astype(_F) - (sp ** 2).astype(_F)).astype(_F)).astype(_F)
    return gaussian_product_coef(sp, delta)


def sampler_step(vf, x_t, x_1, mask, std_fwd_t, mu_x0, mu_xt, sd_post=0.0, std_sb=0.0, noise_post=None,
                 noise_mask=None, mask_pred_x0=True):
    """One iteration of ddpm_sample after the blend (A2SB_lightning_module.py:132-144): (pred_x0, x_next)."""
    vf, x_t, x_1 = (np.asarray(a, _F) for a in (vf, x_t, x_1))
    pred = (x_t - (_F(std_fwd_t) * vf).astype(_F)).astype(_F)
    if mask is not None:
        m = np.asarray(mask, _F)
        om = (_F(1) - m).astype(_F)
        if mask_pred_x0:
            pred = ((pred * m).astype(_F) + (om * x_1).astype(_F)).astype(_F)
    xp = ((_F(mu_x0) * pred).astype(_F) + (_F(mu_xt) * x_t).astype(_F)).astype(_F)
    if noise_post is not None:
        xp = (xp + (_F(sd_post) * np.asarray(noise_post, _F)).astype(_F)).astype(_F)
    xn = xp
    if mask is not None:
        xtrue = x_1
        if noise_mask is not None:
            xtrue = (xtrue + (_F(std_sb) * np.asarray(noise_mask, _F)).astype(_F)).astype(_F)
        xn = ((om * xtrue).astype(_F) + (m * xp).astype(_F)).astype(_F)
    return pred, xn


def ddpm_sample(net, t_to_emb, x_1, t_steps, mask, mask_pred_x0=True, win_length=256, hop_length=256, batch_size=16):
    """ddpm_sample with use_ot_ode=True (A2SB_lightning_module.py:103-146): per-step pred_x0 and states."""
    t_steps = np.asarray(t_steps, _F)
    n_steps = t_steps.shape[1] - 1
    W = x_1.shape[-1]
    x_1 = multidiffusion_pad_inputs(np.asarray(x_1, _F), win_length, hop_length)
    mask = multidiffusion_pad_inputs(np.asarray(mask, _F), win_length, hop_length)
    x_t = x_1.copy()
    preds, states = [], []
    for i in range(n_steps):
        t, tp = t_steps[:, i], t_steps[:, i + 1]
        t_emb = np.repeat(t_to_emb(t), x_1.shape[0], axis=0)
        vf = get_multidiffusion_vf(net, x_t, t_emb, win_length, hop_length, batch_size)
        mu_x0, mu_xt, _var = posterior_coefs(tp, t)
        pred, x_t = sampler_step(vf, x_t, x_1, mask, std_fwd(t)[0], mu_x0[0], mu_xt[0], mask_pred_x0=mask_pred_x0)
        preds.append(multidiffusion_unpad_outputs(pred, W))
        states.append(x_t)
    return preds, states


def fast_inpaint_ddpm_sample(net, t_to_emb, x_1, t_steps, mask, mask_pred_x0=True, win_length=256, hop_length=256,
                             batch_size=16):
    """fast_inpaint_ddpm_sample (A2SB_lightning_module.py:149-180) with use_ot_ode=True: one sampling run per hole
    on the window centred on it (shifted inside the padded width), final pred_x0 pasted back.  Returns (x, windows)."""
    W = x_1.shape[-1]
    x = multidiffusion_pad_inputs(np.asarray(x_1, _F).copy(), win_length, hop_length)
    m = multidiffusion_pad_inputs(np.asarray(mask, _F), win_length, hop_length, padding_constant=0)
    windows = []
    for c in find_middle_of_zero_segments(1 - m[0, 0, 0]):
        l, r = int(c - win_length / 2), int(c + win_length / 2)       # :162-163
        if l < 0:
            r -= l
            l = 0
        if r > x.shape[-1]:
            l -= r - x.shape[-1]
            r = x.shape[-1]
        assert r - l == win_length and l >= 0 and r <= x.shape[-1]
        windows.append((l, r))
        preds, _ = ddpm_sample(net, t_to_emb, x[..., l:r], t_steps, m[..., l:r], mask_pred_x0, win_length, hop_length,
                               batch_size)
        x[..., l:r] = preds[-1]
    return multidiffusion_unpad_outputs(x, W), windows
