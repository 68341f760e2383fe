# ncu captures of the detection-tail, target-assignment, loss and optimizer kernels (round 2 evidence)
set -x
export EFFDET_BENCH_NO_CPU=1
B="python bench.py --no-sub-records --steps 2 --warmup 1"
$B > gpurun_out/r2o_plain.json 2> gpurun_out/r2o_plain.err || exit 1
ncu --set full --clock-control none -k regex:"focal_kernel|smooth_l1_kernel|sgd_momentum_kernel|anchor_targets_kernel|overlap_kernel" -s 10 -c 10 -o gpurun_out/r2o_loss_sgd $B > gpurun_out/r2o_ncu_c.log 2>&1
T="python bench.py --workload d2_infer_b64 --steps 2 --warmup 1"
$T > gpurun_out/r2o_plain_tail.json 2> gpurun_out/r2o_plain_tail.err || exit 1
ncu --set full --clock-control none -k regex:"boxes_kernel|scan_scores_kernel|offsets_kernel|sort_nms|merge_topk" -s 14 -c 14 -o gpurun_out/r2o_tail $T > gpurun_out/r2o_ncu_d.log 2>&1
ls -la gpurun_out/r2o*
