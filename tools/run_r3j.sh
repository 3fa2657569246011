#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3j_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r3j_tests.log
timeout 300 python tools/bench_nfft.py > gpurun_out/r3j_nfft.json 2> gpurun_out/r3j_nfft.err; echo "nfft rc=$?"
python - <<'PY'
import json
r = json.load(open("gpurun_out/r3j_nfft.json"))
for n, v in r.items():
    print(n, "T", v["T"], "K1 %.3f (%.0f)  K2 %.3f (%.0f) | pitched K1 %.3f K2 %.3f" % (v["k1_ms"], v["k1_gbs"], v["k2_ms"], v["k2_gbs"], v["pitched_k1_ms"], v["pitched_k2_ms"]))
PY
