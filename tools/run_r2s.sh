#!/bin/bash
# round 2, 32-frame forward tiles for n_fft 512 / 1024: GPU tests + per-n_fft timings (F = 32 default vs A2SB_FWD_TILE=16)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2s_tests.log
timeout 300 python tools/bench_nfft.py > gpurun_out/r2s_nfft_f32.json 2> gpurun_out/r2s_nfft_f32.err; echo "nfft rc=$?"
A2SB_FWD_TILE=16 timeout 300 python tools/bench_nfft.py > gpurun_out/r2s_nfft_f16.json 2> gpurun_out/r2s_nfft_f16.err; echo "nfft16 rc=$?"
python - <<'PY'
import json
for tag in ("f32", "f16"):
    try:
        r = json.load(open(f"gpurun_out/r2s_nfft_{tag}.json"))
        for n, v in r.items():
            print(tag, n, "T", v["T"], "K1 %.3f (%.0f)  K2 %.3f (%.0f) | pitched K1 %.3f K2 %.3f" % (v["k1_ms"], v["k1_gbs"], v["k2_ms"], v["k2_gbs"], v["pitched_k1_ms"], v["pitched_k2_ms"]))
    except Exception as e:
        print(tag, "failed", e)
PY
