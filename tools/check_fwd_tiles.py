"""Forward kernel: the tile geometries of one n_fft must give bit-identical spectrograms (A2SB_FWD_TILE is read when a plan is
created, so each geometry runs in its own process).  Usage: python tools/check_fwd_tiles.py [n_fft ...]   (needs a GPU)"""
import hashlib
import os
import subprocess
import sys

CHILD = r"""
import sys, hashlib, torch
sys.path.insert(0, %r)
from audio_intelligence_b200 import _capi, _lib
n = int(sys.argv[1])
g = torch.Generator(device="cuda").manual_seed(7)
wav = (0.3 * torch.randn(5, 44100 * 3 + 17, device="cuda", generator=g)).clamp_(-1, 1)
wav[1, 5000:30000] = 0.0          # digital silence: the careful path
out = []
for kw in (dict(kind=_capi.KIND_MAGPHASE, drop_dc=True, power=0.25), dict(kind=_capi.KIND_COMPLEX),
           dict(kind=_capi.KIND_MAGPHASE, drop_dc=False, power=None)):
    s = _lib.stft_forward(wav, n, n, n // 4, **kw)
    out.append(hashlib.sha256(s.cpu().numpy().tobytes()).hexdigest()[:16])
print(" ".join(out))
"""


def main():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ok = True
    for n in [int(a) for a in sys.argv[1:]] or [512, 1024, 2048]:
        res = {}
        tiles = tuple(os.environ.get("A2SB_CHECK_TILES", "16,32").split(","))
        for tile in tiles:
            r = subprocess.run([sys.executable, "-c", CHILD % root, str(n)], capture_output=True, text=True,
                               env=dict(os.environ, A2SB_FWD_TILE=tile))
            res[tile] = r.stdout.strip() or r.stderr.strip()[-300:]
        a, b = tiles
        same = res[a] == res[b]
        ok &= same
        print(f"n_fft {n}: tile {a} {res[a]} | tile {b} {res[b]} -> {'bit-identical' if same else 'DIFFERENT'}")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
