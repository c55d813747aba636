"""Selected raw metrics per kernel launch of an `ncu --set full` report (memory-bound kernels: bytes, hit rates, issue activity).
Usage: python profiles/ncu_pick.py report.ncu-rep "<how the capture was taken>" > profiles/<name>.txt"""
import csv, io, subprocess, sys
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
print(sys.argv[2] if len(sys.argv) > 2 else "")
for r in rows[2:]:
    print()
    print(r[col["Kernel Name"]][:150])
    for k in KEEP:
        if k in col:
            print("  %-80s %s %s" % (k, r[col[k]], units[col[k]]))
