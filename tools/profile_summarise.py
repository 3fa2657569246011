"""Turn the reports of tools/profile_round.sh (gpurun_out/) into the committed summaries under profiles/.
usage: python tools/profile_summarise.py r01"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, GP = os.path.join(ROOT, "profiles"), os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
           "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
           "lts__t_sectors_srcunit_tex_op_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def raw(rep):
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    bench = json.loads(open(os.path.join(GP, "prof_bench.json")).read().strip().splitlines()[-1])
    alg = {k: v["algorithmic_bytes"] for k, v in bench["roofline"]["kernels"].items()}
    traffic, lines = {}, []
    lines.append(f"# ncu --set full --clock-control none --import-source on, bench.py --steps 1 --warmup 3 --skip-cpu --skip-e2e "
                 f"--skip-aligned (256 x 10 s clips, n_fft 2048 / hop 512)")
    lines.append("# per-launch values; times under ncu are cold-cache and serialised -- the bench line below is the number that counts")
    lines.append("# bench (no profiler, same box): " + json.dumps({k: round(v["ms"], 4) for k, v in bench["roofline"]["kernels"].items()}))
    for key, rep in (("stft_fwd_kernel", "prof_full_k1.ncu-rep"), ("istft_inv_kernel", "prof_full_k2.ncu-rep")):
        hdr, units, rows = raw(os.path.join(GP, rep))
        r = rows[0]
        lines.append("")
        lines.append("## " + r[hdr.index("Kernel Name")])
        for m in METRICS:
            if m in hdr:
                lines.append(f"{m:80s} {r[hdr.index(m)]:>16s} {units[hdr.index(m)]}")
        rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        traffic[key] = int(rd + wr)
        lines.append(f"dram traffic (read+write) per launch: {(rd + wr) / 1e9:.3f} GB; algorithmic {alg[key] / 1e9:.3f} GB")
        src = subprocess.run(["ncu", "-i", os.path.join(GP, rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
        tmp = os.path.join(GP, rep + ".src.csv")
        open(tmp, "w").write(src)
        summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_src_summary.py"), tmp, "25"], capture_output=True, text=True).stdout
        open(os.path.join(OUT, f"{tag}_{key.replace('_kernel', '')}_source_summary.txt"), "w").write(summ)
    open(os.path.join(OUT, f"{tag}_ncu_full_summary.txt"), "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
    # launch list: keep our kernels' names readable, collapse torch's template names
    rows = list(csv.reader(l for l in open(os.path.join(GP, "prof_launches.csv")) if l.startswith('"')))
    hdr = rows[0]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    with open(os.path.join(OUT, f"{tag}_launches.csv"), "w") as fh:
        fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-aligned\n")
        fh.write("id,kernel,grid,block,gpu__time_duration.sum[ns]\n")
        tot, ours = 0.0, 0.0
        for r in rows[1:]:
            name = r[ik]
            short = name.split("<")[0].replace("void ", "") if "a2sb" not in name and "stft" not in name else name.replace("void ", "")
            t = float(r[iv].replace(",", ""))
            tot += t
            if "stft" in name:
                ours += t
            fh.write(f'{r[0]},"{short[:90]}",{r[ig]},{r[ib]},{t:.0f}\n')
        fh.write(f"# total {tot / 1e6:.3f} ms, transform kernels {ours / 1e6:.3f} ms ({100 * ours / tot:.1f} %)\n")
    json.dump(bench, open(os.path.join(OUT, f"{tag}_bench_line.json"), "w"), indent=1)
    families(tag)
    print("\n".join(lines[:12]))


def families(tag):
    """profiles/<tag>_nfft_families.txt: CUDA-event timings per n_fft (tools/bench_nfft.py) + the ncu --set full key metrics of
    K1 and K2 of every family (n_fft 2048 comes from the bench captures)."""
    path = os.path.join(GP, "prof_nfft.json")
    if not os.path.exists(path):
        return
    res = json.load(open(path))
    out = ["# BASELINE configs[3]: 256 x 10 s clips, hop = n_fft / 4, contiguous reference layout unless 'pitched' (opt-in 32-byte row pitch)",
           "# CUDA-event medians (tools/bench_nfft.py); GB/s on algorithmic bytes (SURVEY 8d)",
           "n_fft     T  K1 ms (GB/s)   K2 ms (GB/s)   K1 pitched  K2 pitched"]
    for n, v in res.items():
        out.append(f"{n:>5s} {v['T']:5d}  {v['k1_ms']:.3f} ({v['k1_gbs']:.0f})   {v['k2_ms']:.3f} ({v['k2_gbs']:.0f})   "
                   f"{v['pitched_k1_ms']:.3f}       {v['pitched_k2_ms']:.3f}")
    keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
            "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]
    for n in ("512", "1024", "4096"):
        for k in ("k1", "k2"):
            rep = os.path.join(GP, f"prof_full_{k}_n{n}.raw.csv")
            if not os.path.exists(rep):
                continue
            hdr, units, rows = raw(rep)
            if not rows:
                continue
            r = rows[0]
            out.append("")
            out.append(f"## n_fft {n} {k.upper()}: " + r[hdr.index("Kernel Name")])
            for m in keep:
                if m in hdr:
                    out.append(f"{m:72s} {r[hdr.index(m)]:>16s} {units[hdr.index(m)]}")
    open(os.path.join(OUT, f"{tag}_nfft_families.txt"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
