"""Summarises an .ncu-rep (read here on the CPU box with `ncu -i`): one line per captured launch with the
metrics the profiling recipe names.  usage: ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__t_bytes.sum", "l2_MB"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
]


def to_num(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    return x * scale.get(unit, 1.0)


for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {}
    for i, h in enumerate(hdr):
        for k, _ in KEYS:
            if h.endswith(k) and k not in ix:
                ix[k] = i
    kn = hdr.index("Kernel Name")
    print("## %s" % rep)
    print("| kernel | " + " | ".join(n for _, n in KEYS) + " |")
    print("|---|" + "---|" * len(KEYS))
    for r in data:
        cells = []
        for k, _ in KEYS:
            if k in ix:
                v = to_num(r[ix[k]], units[ix[k]])
                cells.append("%.1f" % v if isinstance(v, float) else str(v))
            else:
                cells.append("-")
        print("| %s | %s |" % (r[kn].split("(")[0][-40:], " | ".join(cells)))
