// ThreadSanitizer driver (test infrastructure): runs the kernel sources on the CUDA emulator, where every CUDA thread
// is an OS thread, so that unsynchronised shared-memory accesses between "CUDA threads" show up as data races.
// Build: g++ -std=c++20 -O1 -g -fsanitize=thread -DA2SB_EMU -DA2SB_INST_ALL -I tests/emu -I audio_intelligence_b200/csrc \
//        -x c++ audio_intelligence_b200/csrc/a2sb_api.cu audio_intelligence_b200/csrc/inst.cu tests/emu/tsan_driver.cpp -pthread
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/a2sb_b200.h"

static int run(int n_fft, long long L, int batch) {
    const int hop = n_fft / 4;
    a2sb_plan* plan = nullptr;
    if (a2sb_plan_create(&plan, n_fft, n_fft, hop, nullptr)) { std::printf("plan: %s\n", a2sb_last_error()); return 1; }
    const long long T = a2sb_num_frames(L, hop), rows = n_fft / 2, out_len = a2sb_istft_length(T, hop);
    std::vector<float> wav((size_t)batch * L), spec((size_t)batch * 3 * rows * T), out((size_t)batch * out_len);
    unsigned s = 12345u;
    for (auto& v : wav) { s = s * 1664525u + 1013904223u; v = ((s >> 8) & 0xffff) / 65536.0f - 0.5f; }
    a2sb_fwd_args fa{};
    fa.d_wav = wav.data(); fa.batch = batch; fa.len = L; fa.wav_stride = L; fa.n_local = L; fa.t_begin = 0; fa.t_end = T;
    fa.d_out = spec.data(); fa.out_kind = A2SB_KIND_MAGPHASE; fa.drop_dc = 1; fa.power_on = 1; fa.power = 0.25f; fa.eps = 1e-9f;
    if (a2sb_stft_forward(plan, &fa)) { std::printf("fwd: %s\n", a2sb_last_error()); return 1; }
    a2sb_inv_args ia{};
    ia.d_spec = spec.data(); ia.batch = batch; ia.n_frames = T; ia.spec_T = T; ia.in_kind = A2SB_KIND_MAGPHASE; ia.has_dc = 0;
    ia.phase_fix = 1; ia.power_on = 1; ia.power = 4.0f; ia.eps = 1e-9f; ia.d_wav = out.data(); ia.wav_stride = out_len;
    ia.out_first = 0; ia.out_count = out_len;
    if (a2sb_istft_inverse(plan, &ia)) { std::printf("inv: %s\n", a2sb_last_error()); return 1; }
    // the careful path: complex output / complex input, every bin through the scalar emitters
    std::vector<float> cspec((size_t)batch * 2 * (rows + 1) * T), out2((size_t)batch * out_len);
    fa.d_out = cspec.data(); fa.out_kind = A2SB_KIND_COMPLEX; fa.drop_dc = 0; fa.power_on = 0;
    if (a2sb_stft_forward(plan, &fa)) { std::printf("fwd complex: %s\n", a2sb_last_error()); return 1; }
    ia.d_spec = cspec.data(); ia.in_kind = A2SB_KIND_COMPLEX; ia.has_dc = 1; ia.phase_fix = 0; ia.power_on = 0; ia.d_wav = out2.data();
    if (a2sb_istft_inverse(plan, &ia)) { std::printf("inv complex: %s\n", a2sb_last_error()); return 1; }
    double e = 0, r = 0, d = 0;
    for (long long i = 0; i < out_len; ++i) d += (out2[i] - wav[i]) * (double)(out2[i] - wav[i]);
    for (long long i = 0; i < out_len; ++i) { e += out[i] * out[i]; r += wav[i] * wav[i]; }
    std::printf("n_fft %d: T %lld, out energy ratio %.3f, complex round trip err %.2e\n", n_fft, T, e / r, std::sqrt(d / r));
    a2sb_plan_destroy(plan);
    return 0;
}

int main(int argc, char** argv) {
    const int n_fft = argc > 1 ? std::atoi(argv[1]) : 512;
    return run(n_fft, 20LL * (n_fft / 4) + 37, 1);
}
