# Final round-2 evidence (HEAD): default bench line, reference arm, launch list of the default command (short form),
# ncu of the detection tail after the score-bitmask change.  Only small text files are left in gpurun_out/.
set -x
python bench.py > gpurun_out/r2F_bench.json 2> gpurun_out/r2F_bench.err || exit 1
python bench.py --impl reference > gpurun_out/r2F_ref.json 2> gpurun_out/r2F_ref.err
export EFFDET_BENCH_NO_CPU=1
B="python bench.py --no-sub-records --steps 2 --warmup 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file /tmp/r2F_launches.csv $B > gpurun_out/r2F_ncu_a.log 2>&1
python profiles/tools/ncu_durations.py /tmp/r2F_launches.csv > gpurun_out/r2F_launches_summary.txt
T="python bench.py --workload d2_infer_b64 --steps 2 --warmup 1"
ncu --set full --clock-control none -k regex:"boxes_kernel|scan_scores_kernel|offsets_kernel|sort_nms|merge_topk" -s 14 -c 7 -o /tmp/r2F_tail $T > gpurun_out/r2F_ncu_d.log 2>&1
python profiles/tools/ncu_kernels_table.py /tmp/r2F_tail.ncu-rep > gpurun_out/r2F_tail_ncu.txt
ls -la gpurun_out/r2F*
