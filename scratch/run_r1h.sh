set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1h_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r1h_gpu_tests.log
EFFDET_DUMP_OPS=gpurun_out/r1h_ops_d4.json python bench.py --workload d4_train_b8 --steps 10 --warmup 3 > gpurun_out/r1h_bench_d4.json 2> gpurun_out/r1h_bench_d4.err; echo "bench d4 rc=$?"
cat gpurun_out/r1h_bench_d4.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bn_act_bwd|scale_shift_act|colreduce|se_bwd_reduce|dw_dgrad|se_apply" -c 400 --csv --log-file gpurun_out/r1h_d4_bn_launches.csv python bench.py --workload d4_train_b8 --steps 1 --warmup 1 > gpurun_out/r1h_ncu.log 2>&1; echo "ncu rc=$?"
