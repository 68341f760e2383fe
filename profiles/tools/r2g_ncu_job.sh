set -x
export EFFDET_BENCH_NO_CPU=1
B="python bench.py --no-sub-records --steps 2 --warmup 1"
$B > gpurun_out/r2g_plain.json 2> gpurun_out/r2g_plain.err || exit 1
# (a) launch list of the default bench command (short form)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2g_launches.csv $B > gpurun_out/r2g_ncu_a.log 2>&1
# (b) every depthwise launch of one replayed step
ncu --set full --clock-control none --import-source on -k regex:dwconv_tma_kernel -s 96 -c 48 -o gpurun_out/r2g_dwconv $B > gpurun_out/r2g_ncu_b.log 2>&1
# (c) losses / optimizer kernels of the training step
ncu --set full --clock-control none -k regex:"focal_kernel|smooth_l1_kernel|sgd_momentum_kernel|anchor_targets_kernel|overlap_kernel" -s 10 -c 10 -o gpurun_out/r2g_loss_sgd $B > gpurun_out/r2g_ncu_c.log 2>&1
# (d) detection tail (D2, batch 64, 90 classes, ~5000 candidates per image)
T="python bench.py --workload d2_infer_b64 --steps 2 --warmup 1"
$T > gpurun_out/r2g_plain_tail.json 2> gpurun_out/r2g_plain_tail.err || exit 1
ncu --set full --clock-control none -k regex:"boxes_kernel|scan_scores_kernel|offsets_kernel|sort_nms|merge_topk" -s 14 -c 14 -o gpurun_out/r2g_tail $T > gpurun_out/r2g_ncu_d.log 2>&1
ls -la gpurun_out/r2g*
