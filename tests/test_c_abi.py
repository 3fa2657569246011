"""The C-ABI boundary without a GPU: include/a2sb_b200.h, the ctypes prototypes and the built CUDA library agree
on the exported symbols; only device-free entry points are called."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "a2sb_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(a2sb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_ctypes_prototypes_agree():
    from audio_intelligence_b200 import _capi
    syms = declared_symbols()
    assert len(syms) >= 18
    assert sorted(_capi.PROTOTYPES) == syms


def test_cuda_library_loads_and_exports_every_declared_symbol():
    from audio_intelligence_b200 import _capi, build
    path = build.build()                                    # nvcc cross-compiles sm_100a without a GPU
    lib = _capi.bind(C.CDLL(path))                          # raises AttributeError on a missing symbol
    for name in declared_symbols():
        assert getattr(lib, name) is not None
    assert lib.a2sb_is_device_build() == 1
    assert lib.a2sb_version() > 0
    assert lib.a2sb_num_frames(441000, 512) == 862          # torch.stft: 1 + L // hop
    assert lib.a2sb_istft_length(862, 512) == 440832        # hop * (T - 1)
    assert isinstance(lib.a2sb_last_error(), bytes)


def test_sass_uses_packed_fp32_and_tma():
    """The forward kernel family is really sm_100a code: packed FFMA2 butterflies and the TMA bulk copy (UBLKCP)."""
    import shutil
    import subprocess
    from audio_intelligence_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "FFMA2" in out and "UBLKCP" in out
