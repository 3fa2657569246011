#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3c_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r3c_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-long > gpurun_out/r3c_bench.log 2> gpurun_out/r3c_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for ln in open("gpurun_out/r3c_bench.log"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("value", d["value"], "ms", d["ms_per_step"])
        e = d["e2e"]; print("e2e", e["value"], e["ms_per_step"], "floor", e["pcie_probe"]["floor_ms"]); print("pcm16", e.get("pcm16"))
PY
