/* a2sb_b200.h -- C ABI of liba2sb_b200.so: the B200-native spectral-transform hot path of A2SB.
 *
 * The reference (NVIDIA/audio-intelligence, A2SB/) has no FFI for this path: the boundary is the
 * Python transform-module API.  This header is the C-ABI a binding for that API calls into; the
 * Python mirror of the reference interface (audio_intelligence_b200/audio_transforms/transforms.py,
 * audio_intelligence_b200/diffusion.py) binds it with ctypes.  Each entry point cites the reference
 * interface it replaces (paths relative to the reference root).
 *
 * Conventions: plain pointers and sizes only; `stream` is a cudaStream_t passed as void* (NULL =
 * legacy default stream); `d_` pointers are device memory on the current CUDA device, `h_`
 * pointers are host memory; every function returns A2SB_OK or a negative error code and
 * a2sb_last_error() returns the thread-local message.  Nothing here falls back to the CPU.
 */
#ifndef A2SB_B200_H_
#define A2SB_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A2SB_OK 0
#define A2SB_ERR_INVALID (-1)   /* bad argument / unsupported parameter combination            */
#define A2SB_ERR_CUDA (-2)      /* CUDA runtime error (message has the cudaGetErrorString text) */
#define A2SB_ERR_NOLA (-3)      /* istft: window overlap-add envelope < 1e-11 (torch.istft check)*/

#define A2SB_KIND_COMPLEX 0     /* [2, n_fft/2+1, T]  (re, im)   -- ComplexSpectrogram layout    */
#define A2SB_KIND_MAGPHASE 1    /* [3, rows, T]       (mag, cos, sin) -- ComplexToMagInstPhase   */

#define A2SB_OP_COMPLEX_TO_MAGPHASE 0
#define A2SB_OP_MAGPHASE_TO_COMPLEX 1
#define A2SB_OP_PHASE_FIX 2
#define A2SB_OP_POWER_SCALE 3

typedef struct a2sb_plan a2sb_plan;

const char* a2sb_last_error(void);
int a2sb_version(void);
/* 1 when built by nvcc for sm_100a, 0 for the CPU emulation build used only by tests/. */
int a2sb_is_device_build(void);

/* Plan = (n_fft, win_length, hop_length) + device tables (window, twiddles, envelope).
 * Mirrors the constructor of ComplexSpectrogram / InverseComplexSpectrogram
 * (A2SB/audio_transforms/transforms.py:83-96,163-175: torchaudio Spectrogram with
 * window_fn=torch.hann_window, center=True, pad_mode="reflect", onesided, not normalized).
 * h_window: win_length floats (e.g. torch.hann_window(win_length)) or NULL for a periodic Hann
 * evaluated in double precision.  Supported: n_fft in {512,1024,2048,4096}, win_length <= n_fft, hop_length even.
 * The inverse transform (and a2sb_roundtrip_host) additionally needs hop_length % 4 == 0 and n_fft % hop_length == 0;
 * other hops give a forward-only plan (the STFT of ETTA's auraloss uses 50 / 120 / 240, auraloss.py:363-372).
 * torch's `normalized=True` (ETTA adp.py:1543,1583) is a window pre-scaled by n_fft^-1/2: it scales the forward transform by
 * n_fft^-1/2 and, through the squared-window envelope, the inverse by n_fft^+1/2. */
int a2sb_plan_create(a2sb_plan** plan, int n_fft, int win_length, int hop_length, const float* h_window);
int a2sb_plan_destroy(a2sb_plan* plan);

/* Frame count of the forward transform, T = 1 + len / hop (torch.stft, center=True). */
int64_t a2sb_num_frames(int64_t len, int hop_length);
/* Output length of the inverse transform, hop * (T - 1) (torch.istft with length=None). */
int64_t a2sb_istft_length(int64_t n_frames, int hop_length);

typedef struct a2sb_fwd_args {
    const float* d_wav;      /* [batch][wav_stride]; sample `sample_first + i` of each clip at [i]  */
    int64_t batch;
    int64_t len;             /* global clip length L                                              */
    int64_t wav_stride;      /* elements between consecutive clips                                */
    int64_t sample_first;    /* global index of d_wav[b][0] (0 unless the clip is sharded)        */
    int64_t n_local;         /* samples present per clip in d_wav (== len unless sharded)         */
    int64_t t_begin, t_end;  /* global frame range to compute; [0, T) unless sharded              */
    float* d_out;            /* [batch][C][rows][out_pitch], frames fastest                       */
    int64_t out_pitch;       /* elements between consecutive rows of d_out; 0 = t_end - t_begin (contiguous).
                                A pitch that is a multiple of 8 makes every row 32-byte aligned: K1 then
                                writes whole sectors only (1.4x faster than with T*4 % 32 != 0)          */
    int out_kind;            /* A2SB_KIND_*                                                       */
    int drop_dc;             /* MAGPHASE only: rows = bins 1..n_fft/2  (SpectrogramDropDCTerm)    */
    int power_on;            /* MAGPHASE only: PowerScaleSpectrogram on channel 0                 */
    float power, eps;
    void* stream;
    int64_t wrap_cols;       /* > 0: multidiffusion_pad_inputs fused (A2SB/diffusion.py:67-83): the first wrap_cols frames are
                                also written at columns T .. T + wrap_cols - 1 of every row (the padding the segment
                                windowing appends by copying the head).  Needs the whole clip in one launch
                                (t_begin = 0, t_end = T) and out_pitch >= T + wrap_cols                              */
} a2sb_fwd_args;

/* K1. wav -> spectrogram.  Replaces ComplexSpectrogram.__call__ (transforms.py:98-105) and, with
 * out_kind = MAGPHASE, the fused chain ComplexToMagInstPhase (:108-118) ->
 * SpectrogramDropDCTerm (:214-219) -> PowerScaleSpectrogram(power, channels=[0]) (:187-207). */
int a2sb_stft_forward(a2sb_plan* plan, const a2sb_fwd_args* args);

/* K1 with the corruption of A2SB/corruption/corruptions.py in its epilogue (call site
 * A2SB/datasets/datasets.py:235-237: stft_target = forward chain(audio); stft_transformed, mask = mask transform(stft_target)).
 * The clean spectrogram goes to args->d_out as usual; d_out_corrupt (same geometry) receives
 * x * (1 - mask) + mask * noise * level (mask_with_noise, corruptions.py:14-15, in the reference's fp32 operation order) for the
 * rectangle mask rows [row0, row1) x frames [col0, col1) of every [rows][T] slice -- tensor coordinates, python slice semantics,
 * what UpsampleMask / ExtensionMask / InpaintMask / TimestampedSegmentInpaintMaskTransform build (:18-160).  d_noise is
 * torch.randn_like(spec) of the caller's generator: [batch][3][rows][noise_pitch] (0 = frames per row), read only where the mask
 * is one (and where a spectrogram value is a zero, whose sign the expression takes from the noise).  The mask tensor itself is a
 * rectangle: a2sb_rect_mask writes it at streaming speed.  Shipped chain on float32 samples, no wrap padding. */
typedef struct a2sb_corrupt_args {
    float* d_out_corrupt;
    const float* d_noise;
    int64_t noise_pitch;
    int64_t row0, row1, col0, col1;
    float level;
} a2sb_corrupt_args;
int a2sb_stft_forward_corrupt(a2sb_plan* plan, const a2sb_fwd_args* args, const a2sb_corrupt_args* corrupt);

/* K1 on 16-bit PCM (the wav edge): args->d_wav points at int16 samples ([batch][wav_stride], strides and
 * counts in samples) -- the file content librosa.load / soundfile decode to float32 by an exact division by 32768
 * (A2SB/datasets/datasets.py:231-234).  The decode is fused into the kernel's load (the 2^-15 rides on the window), so the
 * result is bit-identical to a2sb_stft_forward on the decoded float32 samples and the input side moves half the bytes.
 * Shipped chain only (MAGPHASE, drop_dc, power 0.25), default tile geometry. */
int a2sb_stft_forward_pcm16(a2sb_plan* plan, const a2sb_fwd_args* args);

typedef struct a2sb_inv_args {
    const float* d_spec;     /* [batch][C][rows][spec_T], frames fastest                           */
    int64_t batch;
    int64_t n_frames;        /* global frame count T                                              */
    int64_t spec_T;          /* frames per row present in d_spec (== n_frames unless sharded)     */
    int64_t spec_t_first;    /* global frame index of column 0 (0 unless sharded)                 */
    int in_kind;             /* A2SB_KIND_*                                                       */
    int has_dc;              /* MAGPHASE only: 1 rows = bins 0..n_fft/2; 0 rows = bins 1..n_fft/2 and
                                the DC bin is re-created as 0*row0 (SpectrogramAddDCTerm :222-228) */
    int phase_fix;           /* MAGPHASE only: SVDFixMagInstPhase (:135-160) in closed form       */
    int power_on;            /* MAGPHASE only: PowerScaleSpectrogram(power, [0]) applied first    */
    float power, eps;
    float* d_wav;            /* [batch][wav_stride]; trimmed sample `out_first + i` at [i]        */
    int64_t wav_stride;
    int64_t out_first;       /* first trimmed output sample produced (0 unless sharded)           */
    int64_t out_count;       /* samples produced per clip (hop*(T-1) unless sharded)              */
    void* stream;
} a2sb_inv_args;

/* K2. spectrogram -> wav.  Replaces InverseComplexSpectrogram.__call__ (transforms.py:177-184) and,
 * with in_kind = MAGPHASE, the fused chain PowerScaleSpectrogram (:187-207) ->
 * SpectrogramAddDCTerm (:222-228) -> SVDFixMagInstPhase (:135-160) -> MagInstPhaseToComplex
 * (:121-132).  out_first/out_count must be multiples of hop_length when sharded. */
int a2sb_istft_inverse(a2sb_plan* plan, const a2sb_inv_args* args);

/* K2 with a fused gather ("direct st.global into the neighbour's buffer" instead of a collective after the
 * kernel; the reference has no multi-GPU form of this path -- A2SB/A2SB_lightning_module.py:185 asserts batch 1 on one GPU).
 * Every output vector is stored at the same offset of additional buffers that live on OTHER GPUs of the box:
 *   A2SB_MIRROR_PEERS      d_mirrors[0..n) are peer-mapped device pointers (cudaIpc / symmetric memory); d_wav is written too;
 *   A2SB_MIRROR_MULTICAST  d_mirrors[0] is a multicast address bound to every GPU's buffer (this GPU's included): ONE
 *                          multimem.st per vector, replicated by NVSwitch; d_wav itself is NOT written.
 * d_mirrors is a HOST array; each entry corresponds to args->d_wav (same offset into the sharded result) and must have its
 * alignment mod 16.  Shipped chain only (MAGPHASE rows 1.., power 4, phase fix).  The caller orders the peers' reads after
 * the kernel (a symmetric-memory barrier / any collective on the same stream). */
/* K2 writing 16-bit PCM (the other wav edge): args->d_wav points at int16 samples.  Conversion rule = libsndfile's float ->
 * PCM_16 with clipping, which is what soundfile.write does to the reconstructed audio
 * (A2SB/inference/A2SB_inpaint_dataset.py:126): lrintf(x * 2^31), saturated, >> 16.  (A2SB_lightning_module.py:204-205 writes
 * float32 WAVs with scipy instead: that edge needs no conversion.)  Shipped chain only. */
int a2sb_istft_inverse_pcm16(a2sb_plan* plan, const a2sb_inv_args* args);

#define A2SB_MIRROR_PEERS 1
#define A2SB_MIRROR_MULTICAST 2
int a2sb_istft_inverse_mirrored(a2sb_plan* plan, const a2sb_inv_args* args, int mode, int n_mirrors, float* const* d_mirrors);

/* Caps the persistent grids of K1 / K2 (CTAs; 0 = default: every SM).  Process-wide.  With both capped to half the SMs a caller
 * can run K1 of one piece and K2 of the previous piece concurrently on two streams (no reference counterpart). */
int a2sb_set_grid_limit(int max_ctas_forward, int max_ctas_inverse);

/* STFT / iSTFT for ANY transform length (ETTA's STFT helper defaults to num_fft = 1023,
 * ETTA/stable_audio_tools/models/adp.py:1510-1590): torch.stft(center=True, pad_mode="reflect", onesided) ->
 * d_spec [batch][2][n_fft/2 + 1][T] (re, im planes), T = 1 + (len + 2 (n_fft/2) - n_fft) / hop; and torch.istft(center=True,
 * length = out_len): overlap-added frames / overlap-added squared window, trimmed by n_fft/2 at the head, zero where no frame
 * reaches.  d_window: n_fft device floats (the window centre-padded to n_fft; `normalized=True` = window / sqrt(n_fft), both
 * directions).  d_frames: scratch of batch * n_frames * n_fft floats.  A plain O(n_fft^2) DFT per frame: a convenience path, not
 * a hot one; the radix kernels (plans) cover n_fft in {512, 1024, 2048, 4096}. */
int a2sb_dft_generic_forward(const float* d_wav, int64_t batch, int64_t len, int64_t wav_stride, int n_fft, int hop_length,
                             const float* d_window, float* d_spec, void* stream);
int a2sb_dft_generic_inverse(const float* d_spec, int64_t batch, int64_t n_frames, int n_fft, int hop_length,
                             const float* d_window, float* d_frames, float* d_out, int64_t out_len, void* stream);

/* Standalone per-bin ops on contiguous [C][n] tensors (transforms.py:108-160,187-207).
 * chan_mask: bit c set = channel c is scaled (POWER_SCALE only; `channels=None` -> all bits). */
int a2sb_pointwise(int op, const float* d_in, float* d_out, int64_t n, int channels, uint32_t chan_mask,
                   float power, float eps, void* stream);

/* One phase update of Griffin-Lim (the loop body of `griffinlim`, A2SB/audio_transforms/transforms.py:351-362):
 * angles = rebuilt - momentum*tprev; angles /= |angles| + 1e-16; product = mag * angles.  rebuilt, tprev, product:
 * [batch][2][n] (re, im planes); mag: [batch][n]; d_tprev may be NULL (first iteration).  SURVEY.md 8f, rank 3. */
int a2sb_griffinlim_update(const float* d_rebuilt, const float* d_tprev, const float* d_mag, float* d_product,
                           int64_t batch, int64_t n, float momentum, void* stream);

/* multidiffusion_pad_inputs (A2SB/diffusion.py:67-83): d_out[row][w] = w < width ? d_in[row][w]
 * : d_in[row][w - width] (head copy), or pad_const when use_const.  out_width - width <= width. */
int a2sb_wrap_pad(const float* d_in, float* d_out, int64_t nrows, int64_t width, int64_t out_width,
                  int use_const, float pad_const, void* stream);

/* K3. Segment windowing of get_multidiffusion_vf (A2SB/diffusion.py:35-42: nn.Unfold +
 * "b (c h w) l -> (b l) c h w").  d_x [batch][rows][width] -> d_seg [(batch*L)][rows][win],
 * L = (width - (win - hop)) / hop. */
int a2sb_segment_gather(const float* d_x, float* d_seg, int64_t batch, int64_t rows, int64_t width, int win,
                        int hop, void* stream);

/* K4. Overlap blend of get_multidiffusion_vf (A2SB/diffusion.py:52-64): ascending-segment sum /
 * overlap count.  d_seg [(batch*L)][rows][win] -> d_out [batch][rows][width]. */
int a2sb_segment_blend(const float* d_seg, float* d_out, int64_t batch, int64_t rows, int64_t width, int win,
                       int hop, void* stream);
/* Same blend, but only output columns [col_off, col_off + col_cnt) are produced, into rows of `out_pitch` floats
 * (d_out [batch][rows][out_pitch], column col_off first).  Used when the frame axis of one long clip is sharded by
 * segment ranges over several GPUs: a rank blends [left-halo segments + its own] and keeps
 * its owned columns, written straight into the pre-padded state buffer of the next sampling step.  The reference has
 * no counterpart (A2SB_lightning_module.py:185 asserts batch 1 on one GPU); values equal a2sb_segment_blend's. */
int a2sb_segment_blend_window(const float* d_seg, float* d_out, int64_t batch, int64_t rows, int64_t width, int win,
                              int hop, int64_t col_off, int64_t col_cnt, int64_t out_pitch, void* stream);

typedef struct a2sb_step_args {
    const float* d_x_t;        /* [batch][rows][width] current state                                  */
    const float* d_x_1;        /* corrupted input (same shape)                                        */
    const float* d_mask;       /* inpainting mask, or NULL                                            */
    const float* d_noise_post; /* randn_like for p_posterior, or NULL (ot_ode, or t_prev == 0)        */
    const float* d_noise_mask; /* randn_like for the known region, or NULL (ot_ode, or no mask)       */
    float* d_pred_x0;          /* out: pred_x0 after the mask merge                                   */
    float* d_x_next;           /* out: x_t for the next step                                          */
    float std_fwd_t;           /* Diffusion.get_std_fwd(t)                 (diffusion.py:125-126)     */
    float mu_x0, mu_xt;        /* compute_gaussian_product_coef(std_t_prev, std_delta) (:91-99,153-158) */
    float sd_post;             /* sqrt(var) of the same product                                       */
    float std_sb;              /* Diffusion.get_std_t(t_prev)              (:131-135)                 */
    int mask_pred_x0;
} a2sb_step_args;

/* K4s. Overlap blend (K4) fused with one reverse step of the bridge sampler: get_pred_x0, mask merge,
 * p_posterior and the re-imposition of the known region (A2SB/A2SB_lightning_module.py:127-144,
 * A2SB/diffusion.py:153-168), fp32 operation order of the reference.  SURVEY.md section 8f, rank 1. */
int a2sb_segment_blend_step(const float* d_seg, const a2sb_step_args* args, int64_t batch, int64_t rows, int64_t width,
                            int win, int hop, void* stream);

/* M1. Corruption masks and noise fill (A2SB/corruption/corruptions.py).  Every mask the reference builds
 * is an axis-aligned rectangle of ones: rows [row0, row1) x frames [col0, col1) of each [rows][width] slice
 * (UpsampleMask :26-51 rows [cutoff, rows); ExtensionMask :60-79, InpaintMask :90-117 and
 * TimestampedSegmentInpaintMaskTransform :147-160 frame ranges).  Bounds follow python slice semantics. */
int a2sb_rect_mask(float* d_mask, int64_t slices, int64_t rows, int64_t width, int64_t row0, int64_t row1,
                   int64_t col0, int64_t col1, void* stream);

/* mask_with_noise (corruptions.py:14-15) for an arbitrary mask tensor: out = x*(1-mask) + mask*noise*level,
 * fp32 operations in the reference's order (bit-identical for the same noise tensor). */
int a2sb_mask_with_noise(const float* d_x, const float* d_mask, const float* d_noise, float* d_out, int64_t n,
                         float level, void* stream);

/* Rectangle mask + mask_with_noise in one pass (MultinomialInpaintMaskTransform.__call__ :132-145,
 * TimestampedSegmentInpaintMaskTransform.__call__ :153-160).  d_mask may be NULL. */
int a2sb_mask_fill(const float* d_x, const float* d_noise, float* d_out, float* d_mask, int64_t slices, int64_t rows,
                   int64_t width, int64_t row0, int64_t row1, int64_t col0, int64_t col1, float level, void* stream);

/* M1 + B1 fused: the same rectangle mask + noise fill, reading a row-pitched x (K1's aligned output, in_pitch elements
 * between rows) and writing the filled spectrogram AND the mask with rows of out_width columns whose tail replicates
 * the head: exactly what multidiffusion_pad_inputs (A2SB/diffusion.py:67-83; A2SB_lightning_module.py:115-116) appends
 * to both tensors before the sampling loop.  d_noise is the contiguous [slices][rows][width] tensor of torch.randn_like. */
int a2sb_mask_fill_padded(const float* d_x, int64_t in_pitch, const float* d_noise, float* d_out, float* d_mask,
                          int64_t slices, int64_t rows, int64_t width, int64_t out_width, int64_t row0, int64_t row1,
                          int64_t col0, int64_t col1, float level, void* stream);

/* B4. find_middle_of_zero_segments (A2SB/utils.py:54-81) and the window clamp of the fast-inpaint sampler
 * (A2SB/A2SB_lightning_module.py:161-174), on the device.  d_row: n values (the reference passes
 * 1 - mask[0,0,0]).  d_count[0] = segments found; the first min(count, max_out) centres go to d_centres and
 * their windows [l, r) (length win_length, shifted inside [0, n]) to d_lr[k][2]. */
int a2sb_zero_segment_windows(const float* d_row, int64_t n, int win_length, int32_t* d_centres, int32_t* d_lr,
                              int32_t* d_count, int max_out, void* stream);

/* End-to-end host-buffer round trip (wav -> A2SB spectrogram -> wav) used by bench.py's `e2e`
 * figure: host->device copies, K1, K2 and device->host copies, pipelined over clip groups on
 * internal streams.  h_spec may be NULL (spectrogram stays on the device).  Host buffers should be
 * page-locked for the copies to overlap.  Returns after everything has completed. */
int a2sb_roundtrip_host(a2sb_plan* plan, const float* h_wav, int64_t batch, int64_t len, float* h_wav_out,
                        float* h_spec, float power_fwd, float power_inv, float eps, int phase_fix);

/* The same round trip with 16-bit PCM host buffers on both sides (a2sb_stft_forward_pcm16 / a2sb_istft_inverse_pcm16):
 * the wav file content goes in, the PCM_16 file content comes out, half the PCIe bytes each way. */
int a2sb_roundtrip_host_pcm16(a2sb_plan* plan, const int16_t* h_pcm, int64_t batch, int64_t len, int16_t* h_pcm_out,
                              float* h_spec, float power_fwd, float power_inv, float eps, int phase_fix);

/* Launch bookkeeping for bench.py: number of kernels launched by this library since load. */
int64_t a2sb_launch_count(void);
/* Number of a2sb_istft_inverse launches so far that took the TMA box-ring variant of the inverse kernel (shipped chain,
 * n_fft <= 2048): lets tests and bench.py assert which variant produced a result. */
int64_t a2sb_tma_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* A2SB_B200_H_ */
