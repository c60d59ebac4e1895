"""Print selected raw metrics of an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv, subprocess, sys
rep = sys.argv[1]
keys = sys.argv[2:] or ["gpu__time_duration.sum", "gpc__cycles_elapsed.max", "sm__pipe_tensor", "l1tex__data_pipe_tc_wavefronts_mem_shared", "l1tex__m_xbar2l1tex_read_bytes.sum",
                        "lts__t_sectors_srcunit_tex.avg.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct",
                        "launch__registers_per_thread", "smsp__warp_issue_stalled", "l1tex__data_bank", "sm__throughput.avg.pct", "l1tex__throughput.avg.pct",
                        "smsp__inst_executed.sum ", "sm__inst_executed_pipe_tc", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "tma", "sm__cycles_active.avg"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")])
    for h, u, v in zip(hdr, units, r):
        if any(k in h + " " for k in keys):
            print(f"  {h} [{u}] = {v}")
