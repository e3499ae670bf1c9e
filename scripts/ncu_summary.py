"""Summarise an .ncu-rep (read here, no GPU): one line per launch with the metrics the roofline uses.
    python scripts/ncu_summary.py gpurun_out/r01_conv.ncu-rep [more.ncu-rep ...]"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_uniform", "uni"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("lts__t_bytes.sum", "l2_bytes"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__cycles_active.avg", "cyc"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%")]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}")
    for row in rows[2:]:
        parts = [row[hdr.index("Kernel Name")][:48]]
        for key, short in WANT:
            cols = [i for i, h in enumerate(hdr) if h == key or h.endswith("." + key)]      # some columns carry a section prefix
            if cols:
                i = cols[0]
                parts.append(f"{short}={row[i]}{units[i]}")
        print("  ".join(parts))
