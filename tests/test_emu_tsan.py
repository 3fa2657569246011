"""Barrier structure of the kernels under ThreadSanitizer: the kernel sources run on the CUDA emulator (one OS
thread per CUDA thread), so a shared-memory access that is not ordered by __syncthreads / __syncwarp / a named
barrier is a data race TSan reports.  A build with the block barrier compiled out proves the detector fires."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "audio_intelligence_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")


def build(tmp, name, extra):
    exe = os.path.join(tmp, name)
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-fsanitize=thread", "-DA2SB_EMU", "-DA2SB_INST_ALL", *extra, "-I", EMU, "-I", CSRC,
           "-x", "c++", os.path.join(CSRC, "a2sb_api.cu"), os.path.join(CSRC, "inst.cu"), os.path.join(EMU, "tsan_driver.cpp"),
           "-pthread", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "tsan" in (r.stderr or "").lower():
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


@pytest.fixture(scope="module")
def tmp(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    return str(tmp_path_factory.mktemp("tsan"))


def run(exe, n_fft, **extra_env):
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66", **extra_env)
    r = subprocess.run([exe, str(n_fft)], capture_output=True, text=True, env=env, timeout=600)
    if "FATAL: ThreadSanitizer" in r.stderr:        # the runtime could not start here (e.g. ASLR layout): not a kernel finding
        pytest.skip("ThreadSanitizer runtime failed to initialise: " + r.stderr.strip().splitlines()[0])
    return r


def test_kernels_are_race_free_on_the_emulator(tmp):
    exe = build(tmp, "drv", [])
    for n_fft in (512, 1024, 2048, 4096):
        r = run(exe, n_fft)
        assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr, (n_fft, r.stderr[-1500:])
        assert f"n_fft {n_fft}" in r.stdout


@pytest.mark.parametrize("mode", ["1"])      # (mode 2 = the plain kernel + tensor-map L2 prefetches: no new shared-memory traffic)
def test_tma_variants_of_the_inverse_kernel_are_race_free(tmp, mode):
    """A2SB_INV_TMA=1: the box ring + job queue (mbarrier-ordered slots, in-place expansion / transform of the exchange by
    whichever warp takes the job); 2: register loads + tensor-map prefetch.  The emulator models mbarriers with a mutex and
    a condition variable, so a slot read that is not ordered after its box's arrival -- or a refill issued before every lane
    has finished reading -- is a data race TSan reports."""
    exe = build(tmp, "drv", [])
    for n_fft in (512, 2048):
        r = run(exe, n_fft, A2SB_INV_TMA=mode)
        assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr, (n_fft, r.stderr[-1500:])


@pytest.mark.parametrize("wide", ["1"])
def test_two_round_forward_kernel_is_race_free(tmp, wide):
    """n_fft 2048 with 32-frame tiles in two rounds (opt-in): the exchange is rewritten between the rounds, the input span
    is read by both, and (wide) the two warps of a residue class swap half-spectra through the class's exchange region
    behind 64-thread named barriers."""
    exe = build(tmp, "drv", [])
    r = run(exe, 2048, A2SB_FWD_TILE="32", A2SB_FWD_WIDE=wide)
    assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr, r.stderr[-1500:]


def test_detector_fires_without_block_barriers(tmp):
    exe = build(tmp, "drv_nosync", ["-DA2SB_EMU_NO_SYNC"])
    r = run(exe, 512)
    assert r.returncode != 0 and "ThreadSanitizer: data race" in r.stderr
