"""Per-kernel table from an .ncu-rep: duration, DRAM bytes, achieved DRAM GB/s, L2 hit rate, issue utilisation.
usage: python profiles/tools/ncu_kernels_table.py report.ncu-rep > profiles/<name>.txt"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
def col(r, name):
    try:
        return float(r[h.index(name)].replace(",", ""))
    except Exception:
        return float("nan")
units = dict(zip(h, rows[1]))
def scale(name):   # bytes -> MB regardless of the unit ncu picked
    u = units.get(name, "")
    return {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
def tscale():
    return {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units.get("gpu__time_duration.sum", "us"), 1.0)
print("%-62s %9s %9s %9s %9s %7s %7s %6s %5s" % ("kernel", "us", "rd MB", "wr MB", "GB/s", "dram%", "L2hit%", "issue%", "regs"))
for r in rows[2:]:
    name = r[h.index("Kernel Name")].replace("effdet::", "")[:62]
    us = col(r, "gpu__time_duration.sum") * tscale()
    rd = col(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum")
    wr = col(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum")
    print("%-62s %9.2f %9.3f %9.3f %9.1f %7.1f %7.1f %6.1f %5d" % (
        name, us, rd, wr, (rd + wr) / us * 1e3 if us else 0,
        col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), col(r, "lts__t_sector_hit_rate.pct"),
        col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), int(col(r, "launch__registers_per_thread"))))
