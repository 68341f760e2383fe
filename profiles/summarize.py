"""Summarises ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python profiles/summarize.py launches gpurun_out/r1_launches.csv > profiles/r1_launches_summary.txt
  python profiles/summarize.py full gpurun_out/r1_prof_conv_tc.ncu-rep > profiles/r1_ncu_conv_tc.txt
"""
import collections
import csv
import re
import subprocess
import sys

RAW = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
       "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
       "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
       "launch__waves_per_multiprocessor", "smsp__cycles_active.avg",
       "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        short = re.sub(r"<.*", "", row["Kernel Name"].split("(")[0]).replace("effdet::", "").replace("void ", "")
        agg[short][0] += 1
        agg[short][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("# launches %d, summed device time %.1f us" % (sum(v[0] for v in agg.values()), tot))
    print("%-36s %7s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-36s %7d %12.1f %10.2f %7.3f" % (k, v[0], v[1], v[1] / v[0], v[1] / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in RAW if w in hdr]
    print("# ncu --set full --clock-control none --import-source on  (%s)" % path)
    for r in rows[2:]:
        print("-" * 100)
        for w, i in idx:
            print("%-70s %s %s" % (w, r[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
