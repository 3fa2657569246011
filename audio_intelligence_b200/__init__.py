"""audio_intelligence_b200 -- B200-native (sm_100a) spectral-transform hot path of A2SB.

Drop-in mirrors of the reference's Python transform-module API:
  audio_intelligence_b200.audio_transforms.transforms   (A2SB/audio_transforms/transforms.py)
  audio_intelligence_b200.diffusion                     (A2SB/diffusion.py, segment windowing/blend)
backed by hand-written CUDA kernels behind the C ABI in include/a2sb_b200.h.
"""
__version__ = "0.1.0"
