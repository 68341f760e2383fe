mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1q_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r1q_gpu_tests.log
EFFDET_DUMP_OPS=gpurun_out/r1q_ops.json python bench.py > gpurun_out/r1q_bench.json 2> gpurun_out/r1q_bench.err; echo "bench rc=$?"
EFFDET_BN_CLUSTER=0 python bench.py > gpurun_out/r1q_bench_nocluster.json 2> gpurun_out/r1q_bench_nocluster.err; echo "bench nc rc=$?"
EFFDET_DUMP_OPS=gpurun_out/r1q_ops_d4.json python bench.py --workload d4_train_b8 --steps 10 --warmup 3 > gpurun_out/r1q_bench_d4.json 2> gpurun_out/r1q_bench_d4.err; echo "bench d4 rc=$?"
