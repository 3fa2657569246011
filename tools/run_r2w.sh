#!/bin/bash
# round 2: n_fft 4096 forward in two rounds (16-frame tiles) -- bit-identity vs the 8-frame kernel, GPU tests, per-n_fft timings
python tools/check_fwd_tiles.py 4096; echo "tiles(16 = two-round, 8 via second run) rc=$?"
A2SB_CHECK_TILES="8,16" python tools/check_fwd_tiles.py 4096; echo "rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2w_tests.log
timeout 300 python tools/bench_nfft.py > gpurun_out/r2w_nfft.json 2> gpurun_out/r2w_nfft.err; echo "nfft rc=$?"
python - <<'PY'
import json
r = json.load(open("gpurun_out/r2w_nfft.json"))
for n, v in r.items():
    print(n, "T", v["T"], "K1 %.3f (%.0f)  K2 %.3f (%.0f) | pitched K1 %.3f K2 %.3f" % (v["k1_ms"], v["k1_gbs"], v["k2_ms"], v["k2_gbs"], v["pitched_k1_ms"], v["pitched_k2_ms"]))
PY
