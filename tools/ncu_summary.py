"""Key roofline metrics per captured kernel from an `ncu --set full` report (read on the CPU box).
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_full.txt"""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = [hdr.index(w) for w in WANT if w in hdr]
for r in rows[2:]:
    print("-" * 100)
    for i in idx:
        print("  %-72s %-16s %s" % (hdr[i], units[i], r[i][:110]))
