"""profiles/traffic.json from an ncu --set full capture taken with EFFDET_PROFILE_KINDS (engine.Plan.profile brackets
one launch per op of those kinds, in plan order):

  EFFDET_PROFILE_KINDS=dwconv,conv1x1_tc EFFDET_BENCH_NO_CPU=1 ncu --profile-from-start off --set full \
      --clock-control none -o gpurun_out/X python bench.py --no-sub-records --steps 2 --warmup 1
  python profiles/tools/traffic_from_ncu.py gpurun_out/X.ncu-rep d0_train_b32 profiles/X_ncu.txt dwconv=dwconv_tma_kernel conv1x1_tc=conv_tc_kernel

Writes, per workload and kind, the mean dram__bytes_read.sum + dram__bytes_write.sum per launch."""
import csv, io, json, os, subprocess, sys
rep, workload, source = sys.argv[1], sys.argv[2], sys.argv[3]     # source: the committed summary of this capture
kinds = dict(a.split("=") for a in sys.argv[4:])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
def mb(r, name):
    u = units[h.index(name)]
    return float(r[h.index(name)].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = os.path.join(root, "profiles", "traffic.json")
out = json.load(open(path))
entry = {}
for kind, kname in kinds.items():
    sel = [r for r in rows[2:] if kname in r[h.index("Kernel Name")]]
    tot = sum(mb(r, "dram__bytes_read.sum") + mb(r, "dram__bytes_write.sum") for r in sel)
    us = sum(float(r[h.index("gpu__time_duration.sum")].replace(",", "")) for r in sel)
    entry[kind] = {"launches": len(sel), "bytes_per_launch": int(tot / max(len(sel), 1)),
                   "source": source}
    print(kind, entry[kind])
out[workload] = entry
json.dump(out, open(path, "w"), indent=1)
