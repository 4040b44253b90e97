"""One line per profiled launch from an .ncu-rep (needs ncu on PATH): duration, DRAM bytes, hit rates, throughput %."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("lts__t_sectors_srcunit_tex_op_read.sum", "l2_rd_sectors_from_l1"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__occupancy_limit_shared_mem", "lim_smem"), ("launch__occupancy_limit_registers", "lim_regs"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "stall_lg"),
        ("l1tex__data_pipe_lsu_wavefronts.sum", "l1_wavefronts")]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
for r in rows[2:]:
    out = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.split("(")[0][-60:]
        else:
            try:
                v = f"{float(v):.4g}{units[i] if n in ('dram_rd', 'dram_wr') else ''}"
            except ValueError:
                pass
        out.append(f"{n}={v}")
    print("  ".join(out))
