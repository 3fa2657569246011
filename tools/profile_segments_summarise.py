"""gpurun_out/prof_segments.csv (ncu metrics pass over tools/bench_segments.py, see tools/profile_round.sh) ->
profiles/rNN_segments_ncu.txt: per streaming kernel the median launch of the pass -- duration, DRAM bytes read / written,
DRAM throughput, against the algorithmic bytes of SURVEY.md section 8d.  usage: python tools/profile_segments_summarise.py r02"""
import collections
import csv
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
plane, W0, W, win, nseg = 4 * 3 * 1024, 310079, 310144, 256, 2422
ALG = {"wrap_pad_kernel": plane * (W0 + W), "segment_gather_kernel": plane * (W + win * nseg),
       "segment_blend_kernel": plane * (win * nseg + W), "segment_blend_step_kernel": plane * (win * nseg + 5 * W),
       "mask_fill_kernel": plane * 4 * W, "mask_fill_padded_kernel": plane * (2 * W0 + 2 * W)}
rows = list(csv.reader(l for l in open(os.path.join(ROOT, "gpurun_out", "prof_segments.csv")) if l.startswith('"')))
hdr = rows[0]
ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.defaultdict(lambda: collections.defaultdict(dict))
for r in rows[1:]:
    name = r[ik].replace("void ", "").split("<")[0]
    per[name][r[iid]][r[im]] = float(r[iv].replace(",", ""))
lines = ["# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput... --clock-control none,",
         "# python tools/bench_segments.py: streaming kernels at config-3 size ([1, 3, 1024, 310144] frames, 2422 segments of 256 @ 128).",
         "# median launch per kernel; times under ncu are serialised / cold-cache -- bench.py's `segment_kernels` (CUDA events) is the number",
         "# that counts; this file is the traffic evidence: DRAM bytes vs algorithmic bytes.",
         f"{'kernel':28s} {'launches':>8s} {'us':>9s} {'dram rd GB':>11s} {'dram wr GB':>11s} {'algorithmic GB':>15s} {'traffic/alg':>11s} {'dram %peak':>10s}"]
for name, launches in per.items():
    ls = [m for m in launches.values() if "gpu__time_duration.sum" in m]
    ls.sort(key=lambda m: m["gpu__time_duration.sum"])
    m = ls[len(ls) // 2]
    rd, wr = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
    alg = ALG.get(name)
    lines.append(f"{name:28s} {len(ls):8d} {m['gpu__time_duration.sum'] / 1e3:9.1f} {rd / 1e9:11.3f} {wr / 1e9:11.3f} "
                 f"{(alg or 0) / 1e9:15.3f} {((rd + wr) / alg if alg else 0):11.3f} "
                 f"{m.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0):10.1f}")
open(os.path.join(ROOT, "profiles", f"{tag}_segments_ncu.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
