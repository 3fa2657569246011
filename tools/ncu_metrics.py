"""Print selected raw metrics of every kernel in an .ncu-rep (reads it through `ncu -i ... --page raw --csv`)."""
import csv, io, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "smsp__inst_executed_op_global_st.sum", "smsp__inst_executed_op_global_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "lts__t_sectors_srcunit_tex_op_read.sum"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    extra = [a for a in sys.argv[1:] if not a.endswith(".ncu-rep")]
    for r in rows[2:]:
        print("==", rep, r[hdr.index("Kernel Name")][:70])
        for w in WANT:
            if w in hdr:
                print(f"  {w:75s} {r[hdr.index(w)]:>18s} {rows[1][hdr.index(w)]}")
