"""Test-only helpers: build and drive the CPU emulation of the kernel sources (tests/emu).

The emulation library is the SAME C-ABI and kernel source as liba2sb_b200.so, compiled with g++
and -DA2SB_EMU so that CUDA threads become OS threads and device pointers are host pointers.  It
exists so index math and barrier structure can be checked in a container without a GPU; it is
never loaded by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "audio_intelligence_b200", "csrc")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "liba2sb_emu.so")


def build_emu(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(EMU_DIR, "cuda_emu.h"),
                                                               os.path.join(ROOT, "include", "a2sb_b200.h")]
    newest = max(os.path.getmtime(s) for s in srcs)
    if not force and os.path.exists(EMU_LIB) and os.path.getmtime(EMU_LIB) >= newest:
        return EMU_LIB
    cmd = ["g++", "-x", "c++", "-std=c++20", "-O2", "-DA2SB_EMU", "-DA2SB_INST_ALL", "-I", EMU_DIR, "-I", CSRC,
           "-shared", "-fPIC", "-pthread", "-o", EMU_LIB,
           os.path.join(CSRC, "a2sb_api.cu"), os.path.join(CSRC, "inst.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + r.stderr[-4000:])
    return EMU_LIB


class Emu:
    """numpy front-end over the emulation library (host pointers everywhere)."""

    def __init__(self):
        import sys
        sys.path.insert(0, ROOT)
        from audio_intelligence_b200 import _capi
        self.capi = _capi
        self.lib = _capi.bind(C.CDLL(build_emu()))
        assert self.lib.a2sb_is_device_build() == 0

    def plan(self, n_fft, hop, win_length=None, window=None):
        win_length = n_fft if win_length is None else win_length
        h = C.c_void_p()
        wp = None
        if window is not None:
            window = np.ascontiguousarray(window, np.float32)
            wp = window.ctypes.data
        self.capi.check(self.lib, self.lib.a2sb_plan_create(C.byref(h), n_fft, win_length, hop, wp))
        return h

    def destroy(self, plan):
        self.lib.a2sb_plan_destroy(plan)

    def forward(self, plan, wav, n_fft, hop, kind=1, drop_dc=1, power=0.25, eps=1e-9, power_on=1,
                t_range=None, sample_first=0, total_len=None, pitch=0, wrap_cols=0, corrupt=None):
        pcm = np.asarray(wav).dtype == np.int16          # 16-bit PCM ingest (a2sb_stft_forward_pcm16)
        wav = np.ascontiguousarray(wav, np.int16 if pcm else np.float32)
        B, n_local = wav.shape
        L = n_local if total_len is None else total_len
        T = 1 + L // hop
        t0, t1 = (0, T) if t_range is None else t_range
        ch = 2 if kind == 0 else 3
        rows = n_fft // 2 + 1 if kind == 0 else n_fft // 2 + 1 - drop_dc
        out = np.full((B, ch, rows, pitch if pitch else t1 - t0), np.nan, np.float32)
        a = self.capi.FwdArgs(wav.ctypes.data, B, L, n_local, sample_first, n_local, t0, t1, out.ctypes.data, pitch, kind,
                              drop_dc, power_on, power, eps, None, wrap_cols)
        if corrupt is not None:      # (noise, rows range, frames range, level): the corruption epilogue
            noise, (r0, r1), (c0, c1), level = corrupt
            noise = np.ascontiguousarray(noise, np.float32)
            assert noise.shape == out.shape
            out2 = np.full_like(out, np.nan)
            ca = self.capi.CorruptArgs(out2.ctypes.data, noise.ctypes.data, 0, r0, r1, c0, c1, level)
            self.capi.check(self.lib, self.lib.a2sb_stft_forward_corrupt(plan, C.byref(a), C.byref(ca)))
            return out, out2
        self.capi.check(self.lib, (self.lib.a2sb_stft_forward_pcm16 if pcm else self.lib.a2sb_stft_forward)(plan, C.byref(a)))
        return out

    def inverse(self, plan, spec, n_fft, hop, kind=1, has_dc=0, phase_fix=1, power=4.0, eps=1e-9, power_on=1,
                n_frames=None, spec_t_first=0, out_range=None, pcm16=False):
        spec = np.ascontiguousarray(spec, np.float32)
        B, _, _, spec_T = spec.shape
        T = spec_T if n_frames is None else n_frames
        total = hop * (T - 1)
        o0, on = (0, total) if out_range is None else out_range
        out = np.full((B, on), -12345, np.int16) if pcm16 else np.full((B, on), np.nan, np.float32)
        a = self.capi.InvArgs(spec.ctypes.data, B, T, spec_T, spec_t_first, kind, has_dc, phase_fix, power_on, power,
                              eps, out.ctypes.data, on, o0, on, None)
        self.capi.check(self.lib, (self.lib.a2sb_istft_inverse_pcm16 if pcm16 else self.lib.a2sb_istft_inverse)(plan, C.byref(a)))
        return out

    def dft_generic_forward(self, wav, n_fft, hop, window):
        wav = np.ascontiguousarray(wav, np.float32)
        window = np.ascontiguousarray(window, np.float32)
        B, n = wav.shape
        T = 1 + (n + 2 * (n_fft // 2) - n_fft) // hop
        out = np.full((B, 2, n_fft // 2 + 1, T), np.nan, np.float32)
        self.capi.check(self.lib, self.lib.a2sb_dft_generic_forward(wav.ctypes.data, B, n, n, n_fft, hop, window.ctypes.data,
                                                                    out.ctypes.data, None))
        return out

    def dft_generic_inverse(self, spec, n_fft, hop, window, out_len):
        spec = np.ascontiguousarray(spec, np.float32)
        window = np.ascontiguousarray(window, np.float32)
        B, _, K, T = spec.shape
        frames = np.full((B, T, n_fft), np.nan, np.float32)
        out = np.full((B, out_len), np.nan, np.float32)
        self.capi.check(self.lib, self.lib.a2sb_dft_generic_inverse(spec.ctypes.data, B, T, n_fft, hop, window.ctypes.data,
                                                                    frames.ctypes.data, out.ctypes.data, out_len, None))
        return out

    def pointwise(self, op, x, out_channels, channels_mask=0xFFFFFFFF, power=1.0, eps=1e-9):
        x = np.ascontiguousarray(x, np.float32)
        n = int(np.prod(x.shape[1:]))
        out = np.full((out_channels,) + x.shape[1:], np.nan, np.float32)
        self.capi.check(self.lib, self.lib.a2sb_pointwise(op, x.ctypes.data, out.ctypes.data, n, x.shape[0],
                                                          channels_mask, power, eps, None))
        return out

    def wrap_pad(self, x, out_width, const=None):
        x = np.ascontiguousarray(x, np.float32)
        W = x.shape[-1]
        nrows = x.size // W
        out = np.full(x.shape[:-1] + (out_width,), np.nan, np.float32)
        self.capi.check(self.lib, self.lib.a2sb_wrap_pad(x.ctypes.data, out.ctypes.data, nrows, W, out_width,
                                                         0 if const is None else 1, 0.0 if const is None else const,
                                                         None))
        return out

    def blend_window(self, segs, b, W, win, hop, col_off, col_cnt, pitch):
        segs = np.ascontiguousarray(segs, np.float32)
        _, c, h, _ = segs.shape
        out = np.full((b, c, h, pitch), np.nan, np.float32)
        self.capi.check(self.lib, self.lib.a2sb_segment_blend_window(segs.ctypes.data, out.ctypes.data, b, c * h, W, win, hop,
                                                                     col_off, col_cnt, pitch, None))
        return out

    def mask_fill_padded(self, x_buf, width, noise, rows_range, cols_range, level, out_width):
        """x_buf [..., rows, pitch] with `width` valid columns per row."""
        x_buf = np.ascontiguousarray(x_buf, np.float32)
        noise = np.ascontiguousarray(noise, np.float32)
        *lead, rows, pitch = x_buf.shape
        slices = int(np.prod(lead)) if lead else 1
        out = np.full(tuple(lead) + (rows, out_width), np.nan, np.float32)
        mask = np.full_like(out, np.nan)
        self.capi.check(self.lib, self.lib.a2sb_mask_fill_padded(x_buf.ctypes.data, pitch, noise.ctypes.data, out.ctypes.data,
                                                                 mask.ctypes.data, slices, rows, width, out_width, rows_range[0],
                                                                 rows_range[1], cols_range[0], cols_range[1], level, None))
        return out, mask

    def gather(self, x, win, hop):
        x = np.ascontiguousarray(x, np.float32)
        b, c, h, W = x.shape
        L = (W - (win - hop)) // hop
        out = np.full((b * L, c, h, win), np.nan, np.float32)
        self.capi.check(self.lib, self.lib.a2sb_segment_gather(x.ctypes.data, out.ctypes.data, b, c * h, W, win, hop, None))
        return out

    def blend(self, segs, b, W, win, hop):
        segs = np.ascontiguousarray(segs, np.float32)
        _, c, h, _ = segs.shape
        out = np.full((b, c, h, W), np.nan, np.float32)
        self.capi.check(self.lib, self.lib.a2sb_segment_blend(segs.ctypes.data, out.ctypes.data, b, c * h, W, win, hop, None))
        return out

    def rect_mask(self, shape, rows_range, cols_range):
        out = np.full(shape, np.nan, np.float32)
        slices = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
        self.capi.check(self.lib, self.lib.a2sb_rect_mask(out.ctypes.data, slices, shape[-2], shape[-1], rows_range[0],
                                                          rows_range[1], cols_range[0], cols_range[1], None))
        return out

    def mask_with_noise(self, x, mask, noise, level):
        x, mask, noise = (np.ascontiguousarray(a, np.float32) for a in (x, mask, noise))
        out = np.full(x.shape, np.nan, np.float32)
        self.capi.check(self.lib, self.lib.a2sb_mask_with_noise(x.ctypes.data, mask.ctypes.data, noise.ctypes.data,
                                                                out.ctypes.data, x.size, level, None))
        return out

    def mask_fill(self, x, noise, rows_range, cols_range, level, want_mask=True):
        x, noise = (np.ascontiguousarray(a, np.float32) for a in (x, noise))
        out = np.full(x.shape, np.nan, np.float32)
        mask = np.full(x.shape, np.nan, np.float32) if want_mask else None
        slices = int(np.prod(x.shape[:-2])) if x.ndim > 2 else 1
        self.capi.check(self.lib, self.lib.a2sb_mask_fill(x.ctypes.data, noise.ctypes.data, out.ctypes.data,
                                                          mask.ctypes.data if want_mask else None, slices, x.shape[-2],
                                                          x.shape[-1], rows_range[0], rows_range[1], cols_range[0],
                                                          cols_range[1], level, None))
        return out, mask

    def zero_segment_windows(self, row, win_length, max_out=None):
        row = np.ascontiguousarray(row, np.float32)
        max_out = row.size // 2 + 1 if max_out is None else max_out
        centres = np.full(max_out, -1, np.int32)
        lr = np.full((max_out, 2), -1, np.int32)
        count = np.zeros(1, np.int32)
        self.capi.check(self.lib, self.lib.a2sb_zero_segment_windows(row.ctypes.data, row.size, win_length,
                                                                     centres.ctypes.data, lr.ctypes.data,
                                                                     count.ctypes.data, max_out, None))
        k = int(count[0])
        return centres[:min(k, max_out)], lr[:min(k, max_out)], k

    def blend_step(self, segs, b, W, win, hop, x_t, x_1, mask, std_fwd_t, mu_x0, mu_xt, sd_post=0.0, std_sb=0.0,
                   noise_post=None, noise_mask=None, mask_pred_x0=True):
        f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        segs, x_t, x_1, mask, noise_post, noise_mask = map(f, (segs, x_t, x_1, mask, noise_post, noise_mask))
        _, c, h, _ = segs.shape
        pred = np.full(x_t.shape, np.nan, np.float32)
        nxt = np.full(x_t.shape, np.nan, np.float32)
        p = lambda a: None if a is None else a.ctypes.data
        a = self.capi.StepArgs(p(x_t), p(x_1), p(mask), p(noise_post), p(noise_mask), p(pred), p(nxt), std_fwd_t, mu_x0,
                               mu_xt, sd_post, std_sb, int(mask_pred_x0))
        self.capi.check(self.lib, self.lib.a2sb_segment_blend_step(segs.ctypes.data, C.byref(a), b, c * h, W, win, hop, None))
        return pred, nxt
