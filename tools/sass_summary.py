"""SASS opcode summary of the headline kernels in liba2sb_b200.so (no GPU needed: cuobjdump -sass).
Usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt
Per kernel: instruction count, packed-fp32 / MUFU / shuffle counts, TMA opcodes, and the widths of global and shared
loads / stores -- the evidence that the code is Blackwell-native (FFMA2 packed fp32, UBLKCP / UTMALDG TMA) on record."""
import collections
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audio_intelligence_b200", "liba2sb_b200.so")
WANT = [
    ("stft_fwd_kernel<1024,32,32,16,FAST=1>  (K1, n_fft 2048, shipped chain)", "stft_fwd_kernelILi1024ELi32ELi32ELi16ELi1E"),
    ("istft_inv_kernel<1024,32,32,16,FAST=1,TMA=0>  (K2, n_fft 2048, shipped chain)", "istft_inv_kernelILi1024ELi32ELi32ELi16ELi1ELi0E"),
    ("istft_inv_kernel<1024,32,32,16,FAST=1,TMA=1>  (K2 experiment: TMA box ring, opt-in)", "istft_inv_kernelILi1024ELi32ELi32ELi16ELi1ELi1E"),
    ("istft_inv_kernel<1024,32,32,16,FAST=1,TMA=2>  (K2 experiment: tensor-map L2 prefetch, opt-in)", "istft_inv_kernelILi1024ELi32ELi32ELi16ELi1ELi2E"),
    ("segment_gather_kernel<4,true>  (K3)", "segment_gather_kernelILi4ELb1E"),
    ("segment_blend_kernel<4,true>  (K4)", "segment_blend_kernelILi4ELb1E"),
    ("segment_blend_step_kernel<4,true>  (K4s)", "segment_blend_step_kernelILi4ELb1E"),
    ("mask_fill_kernel<4,true>", "mask_fill_kernelILi4ELb1E"),
]
GROUPS = [
    ("packed fp32 (FFMA2/FADD2/FMUL2)", r"^(FFMA2|FADD2|FMUL2)"),
    ("scalar fp32 (FFMA/FADD/FMUL)", r"^(FFMA|FADD|FMUL)(\.|$)"),
    ("MUFU", r"^MUFU"),
    ("SHFL", r"^SHFL"),
    ("TMA bulk copy (UBLKCP)", r"^UBLKCP"),
    ("TMA tensor load (UTMALDG)", r"^UTMALDG"),
    ("TMA tensor prefetch (UTMAPF)", r"^UTMAPF"),
    ("mbarrier (SYNCS)", r"^SYNCS"),
    ("LDG.32", r"^LDG\.E(?!\.(64|128))"), ("LDG.64", r"^LDG\.E\.64"), ("LDG.128", r"^LDG\.E\.128"),
    ("STG.32", r"^STG\.E(?!\.(64|128))"), ("STG.64", r"^STG\.E\.64"), ("STG.128", r"^STG\.E\.128"),
    ("LDS.32", r"^LDS(?!\.(64|128))"), ("LDS.64", r"^LDS\.64"), ("LDS.128", r"^LDS\.128"),
    ("STS.32", r"^STS(?!\.(64|128))"), ("STS.64", r"^STS\.64"), ("STS.128", r"^STS\.128"),
    ("BAR", r"^BAR"),
]


def main():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    txt = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            funcs[cur].append(m.group(1))
    print("# SASS opcode summary of liba2sb_b200.so (cuobjdump -sass, sm_100a); counts are static instructions")
    for title, key in WANT:
        names = [n for n in funcs if key in n]
        if not names:
            print(f"\n## {title}\n  (not found)")
            continue
        ops = funcs[names[0]]
        print(f"\n## {title}\n  static instructions: {len(ops)}")
        for label, pat in GROUPS:
            n = sum(1 for o in ops if re.search(pat, o))
            if n:
                print(f"  {label:34s} {n}")
        top = collections.Counter(o.split(".")[0] for o in ops).most_common(8)
        print("  top opcodes: " + ", ".join(f"{k} {v}" for k, v in top))


if __name__ == "__main__":
    main()
