# HEAD evidence (run on the GPU box; only small text summaries are left in gpurun_out/):
#  (a) launch list of the default bench command, (b) DRAM bytes of EVERY launch of the two dominant kernel kinds
#  (-> profiles/traffic.json), (c) ncu --set full of the heaviest launches of both kinds (first 12 in plan order)
set -x
export EFFDET_BENCH_NO_CPU=1
B="python bench.py --no-sub-records --steps 2 --warmup 1"
$B > gpurun_out/r2x_plain.json 2> gpurun_out/r2x_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file /tmp/r2x_launches.csv $B > gpurun_out/r2x_ncu_a.log 2>&1
python profiles/tools/ncu_durations.py /tmp/r2x_launches.csv > gpurun_out/r2x_launches_summary.txt
EFFDET_PROFILE_KINDS=dwconv,conv1x1_tc ncu --profile-from-start off --clock-control none \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread \
  -o /tmp/r2x_traffic $B > gpurun_out/r2x_ncu_b.log 2>&1
python profiles/tools/ncu_kernels_table.py /tmp/r2x_traffic.ncu-rep > gpurun_out/r2x_traffic_table.txt
cp profiles/traffic.json /tmp/traffic_before.json
python profiles/tools/traffic_from_ncu.py /tmp/r2x_traffic.ncu-rep d0_train_b32 profiles/r2_dominant_kernels_traffic.txt dwconv=dwconv_tma_kernel conv1x1_tc=conv_tc_kernel > gpurun_out/r2x_traffic_entry.txt
cp profiles/traffic.json gpurun_out/r2x_traffic.json
EFFDET_PROFILE_KINDS=dwconv,conv1x1_tc ncu --profile-from-start off --set full --clock-control none -c 12 -o /tmp/r2x_full $B > gpurun_out/r2x_ncu_c.log 2>&1
python profiles/tools/ncu_sum.py /tmp/r2x_full.ncu-rep > gpurun_out/r2x_full_summary.txt
ls -la gpurun_out/r2x* /tmp/*.ncu-rep
