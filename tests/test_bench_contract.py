"""bench.py's reference arm (the CPU leg the driver runs beside ours) without a GPU: one JSON line with the
contract's keys.  The GPU arm is exercised on the GPU box by the driver itself."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, A2SB_BENCH_REF_CLIPS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("STFT+iSTFT") and d["vs_baseline"] is None and d["n_gpus"] == 1
    # "reference" when the reference's own transforms.py is reachable (this container), "port" on the GPU box
    want_kind = "reference" if os.path.isfile("/root/reference/A2SB/audio_transforms/transforms.py") else "port"
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    # the line reports what was actually run: measured ms per step of `clips_per_step` clips (no extrapolation)
    assert d["clips_per_step"] == 1 and abs(d["value"] - 10.0 * d["clips_per_step"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["global_clips"] == 256


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
