"""Loader and torch-tensor front-end of liba2sb_b200.so (the sm_100a CUDA library).

There is NO CPU fallback: if the library is missing, was not built by nvcc for sm_100a, or CUDA is
unavailable, every entry point raises.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

from . import _capi
from ._capi import A2SBError  # noqa: F401  (re-exported)

_HERE = os.path.dirname(os.path.abspath(__file__))
# A2SB_LIB_VARIANT selects an experiment build made by `python -m audio_intelligence_b200.build --suffix=_x -D...`
LIB_PATH = os.path.join(_HERE, "liba2sb_b200%s.so" % os.environ.get("A2SB_LIB_VARIANT", ""))
_lock = threading.Lock()
_lib = None
_plans: dict = {}


def lib() -> C.CDLL:
    """The bound CUDA library; raises if it is not present (build with audio_intelligence_b200.build)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} not found: build it with `python -m audio_intelligence_b200.build` "
                        "(nvcc, sm_100a). The A2SB B200 path has no CPU fallback.")
                handle = _capi.bind(C.CDLL(LIB_PATH))
                if handle.a2sb_is_device_build() != 1:
                    raise RuntimeError(f"{LIB_PATH} is not a CUDA (sm_100a) build")
                _lib = handle
    return _lib


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("audio_intelligence_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def launch_count() -> int:
    return int(lib().a2sb_launch_count())


def tma_launch_count() -> int:
    """Inverse-kernel launches so far that took the TMA box-ring variant (shipped chain, n_fft <= 2048)."""
    return int(lib().a2sb_tma_launch_count())


def stream_ptr() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def stage(x: torch.Tensor, keep_pitch: bool = False) -> torch.Tensor:
    """fp32, contiguous, on the current CUDA device (CPU tensors are copied over; never computed on).
    keep_pitch: leave a row-pitched view (unit stride along frames) as it is -- K2 reads it in place."""
    require_cuda()
    if x.dtype != torch.float32:
        x = x.float()
    if not x.is_cuda:
        x = x.cuda(non_blocking=True)
    if keep_pitch and x.dim() >= 3 and x.stride(-1) == 1:
        return x
    return x.contiguous()


PINNED_RETURN_LIMIT = 256 << 20     # bytes; larger results go back through a pageable tensor


def to_host(result: torch.Tensor, device) -> torch.Tensor:
    """Result of a kernel for a caller that handed in a CPU tensor (the reference's call sites do:
    A2SB/datasets/datasets.py:235, A2SB_lightning_module.py:202): copied into PINNED host memory (torch's caching host
    allocator) with one DMA transfer and returned as an ordinary CPU tensor -- no pageable bounce buffer, and a tensor
    that is handed back to this package later (spectrogram -> inverse chain) uploads at full PCIe speed."""
    if result.device == device:
        return result
    if device.type != "cpu" or result.numel() * result.element_size() > PINNED_RETURN_LIMIT or not result.is_cuda:
        return result.to(device)
    src = result if result.is_contiguous() else result.contiguous()
    host = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
    host.copy_(src, non_blocking=True)
    torch.cuda.current_stream(src.device).synchronize()
    return host


def get_plan(n_fft: int, win_length: int, hop_length: int, normalized: bool = False,
             window: torch.Tensor | None = None) -> C.c_void_p:
    """Plan cache keyed by (device, n_fft, win_length, hop, normalized, window).  The default window is
    torch.hann_window(win_length), i.e. exactly the tensor the reference hands to torch.stft (transforms.py:91-96).
    normalized=True (torch.stft / torch.istft `normalized`, used by ETTA's STFT helper, adp.py:1543,1583): the window is
    scaled by n_fft^-1/2, which scales the forward transform by n_fft^-1/2 and -- the inverse divides by the overlap-added
    squared window -- the inverse by n_fft^+1/2."""
    require_cuda()
    wkey = None
    if window is not None:
        window = window.detach().to("cpu", torch.float32).contiguous()
        if window.numel() != int(win_length):
            raise ValueError(f"window has {window.numel()} samples, win_length is {win_length}")
        wkey = hash(window.numpy().tobytes())
    key = (torch.cuda.current_device(), int(n_fft), int(win_length), int(hop_length), bool(normalized), wkey)
    p = _plans.get(key)
    if p is None:
        with _lock:
            p = _plans.get(key)
            if p is None:
                w = (window if window is not None else torch.hann_window(int(win_length), dtype=torch.float32)).contiguous()
                if normalized:
                    w = (w * float(n_fft) ** -0.5).contiguous()
                h = C.c_void_p()
                L = lib()
                _capi.check(L, L.a2sb_plan_create(C.byref(h), int(n_fft), int(win_length), int(hop_length), w.data_ptr()))
                _plans[key] = p = h
    return p


def stft_forward(wav: torch.Tensor, n_fft: int, win_length: int, hop_length: int, *, kind: int, drop_dc: bool = False,
                 power: float | None = None, eps: float = 1e-9, total_len: int | None = None, sample_first: int = 0,
                 t_range: tuple[int, int] | None = None, row_align: int | None = None,
                 out: torch.Tensor | None = None, pad_segments: tuple[int, int] | None = None, normalized: bool = False,
                 window: torch.Tensor | None = None, corrupt: dict | None = None):
    """wav [B, n_local] (cuda fp32) -> [B, C, rows, T] via K1.

    row_align=None returns a contiguous tensor like the reference.  row_align=k (k a multiple of 8) stores the rows
    with a pitch rounded up to a multiple of k frames and returns the [..., :T] view of that buffer: every row is
    then 32-byte aligned, K1 writes whole sectors only and `istft_inverse` reads the view in place.
    pad_segments=(win, hop): the rows get the width multidiffusion_pad_inputs(., win, hop) would pad them to (a multiple
    of hop: 512-byte aligned rows for the shipped 256 / 128) and K1 also writes the padding (the head frames again); the
    returned [..., :T] view carries the padded buffer in `._a2sb_padded`, which multidiffusion_pad_inputs hands out instead
    of launching its own copy.
    corrupt=dict(noise=randn_like(spec), rows=(r0, r1), frames=(c0, c1), level=l): the corruption epilogue
    (a2sb_stft_forward_corrupt) -- returns (clean, corrupted) from ONE pass."""
    L = lib()
    plan = get_plan(n_fft, win_length, hop_length, normalized, window)
    B, n_local = wav.shape
    total = n_local if total_len is None else int(total_len)
    T = 1 + total // hop_length
    t0, t1 = (0, T) if t_range is None else t_range
    ch = 2 if kind == _capi.KIND_COMPLEX else 3
    rows = n_fft // 2 + 1 - (1 if (kind == _capi.KIND_MAGPHASE and drop_dc) else 0)
    n_t = max(t1 - t0, 0)
    pitch = n_t if not row_align else -(-n_t // int(row_align)) * int(row_align)
    wrap = 0
    if pad_segments is not None and (t0, t1) == (0, T) and n_t > 0:
        wrap = segment_pad_width(n_t, *pad_segments) - n_t
        pitch = n_t + wrap
    if out is None:
        out = torch.empty((B, ch, rows, pitch), dtype=torch.float32, device=wav.device)
    elif tuple(out.shape) != (B, ch, rows, pitch) or not out.is_contiguous() or out.dtype != torch.float32:
        raise ValueError(f"out must be a contiguous fp32 tensor of shape {(B, ch, rows, pitch)}")
    a = _capi.FwdArgs(wav.data_ptr(), B, total, wav.stride(0) if B > 1 else n_local, sample_first, n_local, t0, t1,
                      out.data_ptr(), pitch, kind, int(bool(drop_dc)), int(power is not None),
                      float(power if power is not None else 1.0), float(eps), stream_ptr(), wrap)
    if corrupt is not None:
        if wav.dtype != torch.float32 or pitch != n_t or kind != _capi.KIND_MAGPHASE:
            raise ValueError("corruption epilogue: float32 samples, contiguous mag/phase output")
        noise = corrupt["noise"]
        if tuple(noise.shape) != (B, ch, rows, n_t) or not noise.is_contiguous() or noise.dtype != torch.float32 or not noise.is_cuda:
            raise ValueError(f"noise must be a contiguous cuda fp32 tensor of shape {(B, ch, rows, n_t)}")
        out2 = torch.empty_like(out)
        (r0, r1), (c0, c1) = corrupt["rows"], corrupt["frames"]
        ca = _capi.CorruptArgs(out2.data_ptr(), noise.data_ptr(), 0, int(r0), int(r1), int(c0), int(c1), float(corrupt["level"]))
        _capi.check(L, L.a2sb_stft_forward_corrupt(plan, C.byref(a), C.byref(ca)))
        return out, out2
    if wav.dtype == torch.int16:      # 16-bit PCM ingest: decode fused into K1's load (a2sb_stft_forward_pcm16)
        _capi.check(L, L.a2sb_stft_forward_pcm16(plan, C.byref(a)))
    elif wav.dtype == torch.float32:
        _capi.check(L, L.a2sb_stft_forward(plan, C.byref(a)))
    else:
        raise TypeError(f"waveform must be float32 or int16 PCM, got {wav.dtype}")
    if pitch == n_t:
        return out
    view = out[..., :n_t]
    if pad_segments is not None and (t0, t1) == (0, T):
        view._a2sb_padded = (out, int(pad_segments[0]), int(pad_segments[1]), None)
    return view


def segment_pad_width(width: int, win: int, hop: int) -> int:
    """Width multidiffusion_pad_inputs (A2SB/diffusion.py:67-83) pads `width` columns to -- including the reference's
    truncation when the pad is longer than the input (it slices input[..., :to_pad])."""
    if width <= win:
        to_pad = win - width
    else:
        to_pad = -(-(width - win) // hop) * hop + win - width
    return width + min(to_pad, width)


def padded_buffer_of(x: torch.Tensor, win: int, hop: int, const) -> torch.Tensor | None:
    """The wrap-padded buffer `x` is the [..., :W] view of, if a kernel of this package already produced it for (win, hop)."""
    tag = getattr(x, "_a2sb_padded", None)
    if tag is None:
        return None
    buf, twin, thop, tconst = tag
    if (twin, thop, tconst) != (int(win), int(hop), const) or buf.dim() != x.dim() or x.data_ptr() != buf.data_ptr():
        return None
    if tuple(buf.shape[:-1]) != tuple(x.shape[:-1]) or buf.shape[-1] != segment_pad_width(x.shape[-1], win, hop):
        return None
    if x.stride()[:-1] != buf.stride()[:-1] or x.stride(-1) != 1:
        return None
    return buf


def istft_inverse(spec: torch.Tensor, n_fft: int, win_length: int, hop_length: int, *, kind: int, has_dc: bool = True,
                  phase_fix: bool = False, power: float | None = None, eps: float = 1e-9,
                  n_frames: int | None = None, spec_t_first: int = 0,
                  out_range: tuple[int, int] | None = None, out: torch.Tensor | None = None, normalized: bool = False,
                  window: torch.Tensor | None = None, mirrors: list[int] | None = None,
                  multicast: bool = False, pcm16: bool = False) -> torch.Tensor:
    """spec [B, C, rows, spec_T] (cuda fp32) -> wav [B, n_out] via K2.  `spec` is either contiguous or the
    [..., :T] view of a row-pitched buffer made by stft_forward(row_align=...), which is read in place.
    mirrors: device addresses (peer-mapped buffers on other GPUs, each the counterpart of out[0, 0]) that receive the
    same samples from inside the kernel -- the fused gather of a sharded result; multicast=True: ONE multicast address,
    written with multimem.st (then `out` itself is only written through the multicast binding).
    pcm16=True: the waveform is returned as int16 PCM (libsndfile's float -> PCM_16 rule, a2sb_istft_inverse_pcm16)."""
    L = lib()
    plan = get_plan(n_fft, win_length, hop_length, normalized, window)
    B, _, _, spec_T = spec.shape
    T = spec_T if n_frames is None else int(n_frames)
    if not spec.is_contiguous():
        pitch = spec.stride(2)
        rows_ = spec.shape[2]
        pitched = (spec.stride(3) == 1 and pitch >= spec_T and spec.stride(1) == rows_ * pitch and
                   (B == 1 or spec.stride(0) == spec.shape[1] * rows_ * pitch) and n_frames is None)
        if pitched:
            spec_T = pitch          # frames per row present in the buffer; n_frames = T of them are valid
        else:
            spec = spec.contiguous()
    total = hop_length * (T - 1)
    o0, on = (0, total) if out_range is None else out_range
    odt = torch.int16 if pcm16 else torch.float32
    if out is None:
        out = torch.empty((B, max(on, 0)), dtype=odt, device=spec.device)
    elif tuple(out.shape) != (B, max(on, 0)) or not out.is_contiguous() or out.dtype != odt:
        raise ValueError(f"out must be a contiguous {odt} tensor of shape {(B, max(on, 0))}")
    a = _capi.InvArgs(spec.data_ptr(), B, T, spec_T, spec_t_first, kind, int(bool(has_dc)), int(bool(phase_fix)),
                      int(power is not None), float(power if power is not None else 1.0), float(eps),
                      out.data_ptr(), max(on, 0), o0, on, stream_ptr())
    if pcm16:
        _capi.check(L, L.a2sb_istft_inverse_pcm16(plan, C.byref(a)))
    elif mirrors:
        arr = (C.c_void_p * len(mirrors))(*[int(m) for m in mirrors])
        _capi.check(L, L.a2sb_istft_inverse_mirrored(plan, C.byref(a), _capi.MIRROR_MULTICAST if multicast else _capi.MIRROR_PEERS,
                                                     len(mirrors), arr))
    else:
        _capi.check(L, L.a2sb_istft_inverse(plan, C.byref(a)))
    return out


def pointwise(op: int, x: torch.Tensor, out_channels: int, chan_mask: int = 0xFFFFFFFF, power: float = 1.0,
              eps: float = 1e-9) -> torch.Tensor:
    """x [C, ...] -> [out_channels, ...] (standalone per-bin ops)."""
    L = lib()
    n = x[0].numel()
    out = torch.empty((out_channels,) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
    _capi.check(L, L.a2sb_pointwise(op, x.data_ptr(), out.data_ptr(), n, x.shape[0], chan_mask & 0xFFFFFFFF,
                                    float(power), float(eps), stream_ptr()))
    return out


def wrap_pad(x: torch.Tensor, out_width: int, const: float | None) -> torch.Tensor:
    L = lib()
    W = x.shape[-1]
    out = torch.empty(tuple(x.shape[:-1]) + (out_width,), dtype=torch.float32, device=x.device)
    _capi.check(L, L.a2sb_wrap_pad(x.data_ptr(), out.data_ptr(), x.numel() // max(W, 1), W, out_width,
                                   int(const is not None), float(const or 0.0), stream_ptr()))
    return out


def segment_gather(x: torch.Tensor, win: int, hop: int) -> torch.Tensor:
    L = lib()
    b, c, h, W = x.shape
    nh = (W - (win - hop)) // hop if W >= win else 0
    out = torch.empty((b * nh, c, h, win), dtype=torch.float32, device=x.device)
    _capi.check(L, L.a2sb_segment_gather(x.data_ptr(), out.data_ptr(), b, c * h, W, win, hop, stream_ptr()))
    return out


def segment_blend(segs: torch.Tensor, b: int, W: int, win: int, hop: int) -> torch.Tensor:
    L = lib()
    _, c, h, _ = segs.shape
    out = torch.empty((b, c, h, W), dtype=torch.float32, device=segs.device)
    _capi.check(L, L.a2sb_segment_blend(segs.data_ptr(), out.data_ptr(), b, c * h, W, win, hop, stream_ptr()))
    return out


def segment_gather_into(x: torch.Tensor, out: torch.Tensor, win: int, hop: int) -> None:
    """K3 into a caller-owned segment buffer `out` [(b * L), c, h, win] (contiguous; e.g. the tail of a buffer whose
    head receives halo segments from the neighbouring rank)."""
    L = lib()
    b, c, h, W = x.shape
    assert x.is_contiguous() and out.is_contiguous()
    _capi.check(L, L.a2sb_segment_gather(x.data_ptr(), out.data_ptr(), b, c * h, W, win, hop, stream_ptr()))


def segment_blend_window(segs: torch.Tensor, out: torch.Tensor, b: int, W: int, win: int, hop: int, col_off: int,
                         col_cnt: int) -> None:
    """K4 restricted to output columns [col_off, col_off + col_cnt), written to out[..., :col_cnt] of a caller-owned
    buffer `out` [b, c, h, pitch] (pitch = out.shape[-1] >= col_cnt)."""
    L = lib()
    _, c, h, _ = segs.shape
    assert out.is_contiguous() and segs.is_contiguous()
    _capi.check(L, L.a2sb_segment_blend_window(segs.data_ptr(), out.data_ptr(), b, c * h, W, win, hop, col_off, col_cnt,
                                               out.shape[-1], stream_ptr()))


def dft_generic_forward(wav: torch.Tensor, n_fft: int, hop: int, window: torch.Tensor) -> torch.Tensor:
    """wav [B, L] (cuda fp32), window [n_fft] (cuda fp32, already centre-padded / normalised) -> [B, 2, n_fft//2+1, T] (re, im)
    with torch.stft(center=True, reflect) framing, for ANY n_fft (a2sb_dft_generic_forward: plain DFT, not a hot path)."""
    L = lib()
    B, n = wav.shape
    T = 1 + (n + 2 * (n_fft // 2) - n_fft) // hop
    out = torch.empty((B, 2, n_fft // 2 + 1, T), dtype=torch.float32, device=wav.device)
    _capi.check(L, L.a2sb_dft_generic_forward(wav.data_ptr(), B, n, wav.stride(0) if B > 1 else n, n_fft, hop, window.data_ptr(),
                                              out.data_ptr(), stream_ptr()))
    return out


def dft_generic_inverse(spec: torch.Tensor, n_fft: int, hop: int, window: torch.Tensor, out_len: int) -> torch.Tensor:
    """spec [B, 2, n_fft//2+1, T] -> [B, out_len]: torch.istft(center=True, length=out_len) for ANY n_fft."""
    L = lib()
    B, _, K, T = spec.shape
    frames = torch.empty((B, T, n_fft), dtype=torch.float32, device=spec.device)
    out = torch.empty((B, out_len), dtype=torch.float32, device=spec.device)
    _capi.check(L, L.a2sb_dft_generic_inverse(spec.data_ptr(), B, T, n_fft, hop, window.data_ptr(), frames.data_ptr(), out.data_ptr(),
                                              out_len, stream_ptr()))
    return out


def set_grid_limit(forward_ctas: int = 0, inverse_ctas: int = 0) -> None:
    """Cap the persistent grids of K1 / K2 (0 = every SM): a2sb_set_grid_limit."""
    L = lib()
    _capi.check(L, L.a2sb_set_grid_limit(int(forward_ctas), int(inverse_ctas)))


def sm_count() -> int:
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count


def roundtrip_host(wav_pinned: torch.Tensor, out_pinned: torch.Tensor, n_fft: int, hop_length: int, *,
                   power_fwd: float = 0.25, power_inv: float = 4.0, eps: float = 1e-9, phase_fix: bool = True,
                   spec_pinned: torch.Tensor | None = None) -> None:
    """Host-buffer round trip (bench.py `e2e`): H2D, K1, K2, D2H pipelined inside the library.  int16 buffers on both sides
    select the 16-bit PCM edges (a2sb_roundtrip_host_pcm16)."""
    require_cuda()
    L = lib()
    plan = get_plan(n_fft, n_fft, hop_length)
    B, n = wav_pinned.shape
    if wav_pinned.dtype != out_pinned.dtype or wav_pinned.dtype not in (torch.float32, torch.int16):
        raise TypeError("roundtrip_host: both host buffers must be float32, or both int16 PCM")
    fn = L.a2sb_roundtrip_host_pcm16 if wav_pinned.dtype == torch.int16 else L.a2sb_roundtrip_host
    _capi.check(L, fn(plan, wav_pinned.data_ptr(), B, n, out_pinned.data_ptr(),
                                         spec_pinned.data_ptr() if spec_pinned is not None else None,
                                         float(power_fwd), float(power_inv), float(eps), int(bool(phase_fix))))


def rect_mask(shape, device, rows_range: tuple[int, int], cols_range: tuple[int, int]) -> torch.Tensor:
    """[..., rows, width] mask of ones on rows_range x cols_range (python slice bounds), zeros elsewhere."""
    L = lib()
    *lead, rows, width = shape
    slices = 1
    for d in lead:
        slices *= int(d)
    out = torch.empty(tuple(shape), dtype=torch.float32, device=device)
    _capi.check(L, L.a2sb_rect_mask(out.data_ptr(), slices, rows, width, rows_range[0], rows_range[1], cols_range[0],
                                    cols_range[1], stream_ptr()))
    return out


def mask_with_noise(x: torch.Tensor, mask: torch.Tensor, noise: torch.Tensor, level: float) -> torch.Tensor:
    L = lib()
    out = torch.empty_like(x)
    _capi.check(L, L.a2sb_mask_with_noise(x.data_ptr(), mask.data_ptr(), noise.data_ptr(), out.data_ptr(), x.numel(),
                                          float(level), stream_ptr()))
    return out


def mask_fill(x: torch.Tensor, noise: torch.Tensor, rows_range: tuple[int, int], cols_range: tuple[int, int],
              level: float, want_mask: bool = True) -> tuple[torch.Tensor, torch.Tensor | None]:
    """x [..., rows, width]: rectangle mask + noise fill in one kernel -> (filled, mask)."""
    L = lib()
    *lead, rows, width = x.shape
    slices = 1
    for d in lead:
        slices *= int(d)
    out = torch.empty_like(x)
    mask = torch.empty_like(x) if want_mask else None
    _capi.check(L, L.a2sb_mask_fill(x.data_ptr(), noise.data_ptr(), out.data_ptr(), mask.data_ptr() if want_mask else None,
                                    slices, rows, width, rows_range[0], rows_range[1], cols_range[0], cols_range[1],
                                    float(level), stream_ptr()))
    return out, mask


def mask_fill_padded(x: torch.Tensor, noise: torch.Tensor, rows_range: tuple[int, int], cols_range: tuple[int, int],
                     level: float, win: int, hop: int) -> tuple[torch.Tensor, torch.Tensor]:
    """mask_fill for a row-pitched x [..., rows, width] (unit stride along frames), writing the filled tensor and the mask
    with the width multidiffusion_pad_inputs(., win, hop) pads to, padding included; returns the two [..., :width] views,
    each carrying its padded buffer (see stft_forward)."""
    L = lib()
    *lead, rows, width = x.shape
    slices = 1
    for d in lead:
        slices *= int(d)
    pitch = x.stride(-2)
    assert x.stride(-1) == 1 and all(x.stride(i) == x.stride(i + 1) * x.shape[i + 1] for i in range(x.dim() - 3, -1, -1) if x.dim() > 2) \
        or x.is_contiguous(), "x must be contiguous or row-pitched"
    out_w = segment_pad_width(width, win, hop)
    out = torch.empty(tuple(lead) + (rows, out_w), dtype=torch.float32, device=x.device)
    mask = torch.empty_like(out)
    _capi.check(L, L.a2sb_mask_fill_padded(x.data_ptr(), pitch, noise.data_ptr(), out.data_ptr(), mask.data_ptr(), slices, rows, width,
                                           out_w, rows_range[0], rows_range[1], cols_range[0], cols_range[1], float(level),
                                           stream_ptr()))
    vo, vm = out[..., :width], mask[..., :width]
    vo._a2sb_padded = (out, int(win), int(hop), None)
    vm._a2sb_padded = (mask, int(win), int(hop), None)
    return vo, vm


def zero_segment_windows(row: torch.Tensor, win_length: int) -> tuple[torch.Tensor, torch.Tensor]:
    """row [n] (cuda fp32) -> (centres int32 [k], windows int32 [k, 2]) of its zero runs."""
    L = lib()
    n = row.numel()
    max_out = n // 2 + 1                       # zero runs are separated by at least one non-zero
    centres = torch.empty(max_out, dtype=torch.int32, device=row.device)
    lr = torch.empty((max_out, 2), dtype=torch.int32, device=row.device)
    count = torch.zeros(1, dtype=torch.int32, device=row.device)
    _capi.check(L, L.a2sb_zero_segment_windows(row.data_ptr(), n, int(win_length), centres.data_ptr(), lr.data_ptr(),
                                               count.data_ptr(), max_out, stream_ptr()))
    k = int(count.item())
    return centres[:k], lr[:k]


def griffinlim_update(rebuilt: torch.Tensor, tprev: torch.Tensor | None, mag: torch.Tensor, product: torch.Tensor,
                      momentum: float) -> None:
    """rebuilt/tprev/product [B, 2, F, T] (re, im planes), mag [B, F, T]: one Griffin-Lim phase update into `product`."""
    L = lib()
    B = mag.shape[0]
    _capi.check(L, L.a2sb_griffinlim_update(rebuilt.data_ptr(), tprev.data_ptr() if tprev is not None else None,
                                            mag.data_ptr(), product.data_ptr(), B, mag.numel() // max(B, 1),
                                            float(momentum), stream_ptr()))
