set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1j_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -12 gpurun_out/r1j_gpu_tests.log
EFFDET_DUMP_OPS=gpurun_out/r1j_ops_d4.json python bench.py --workload d4_train_b8 --steps 10 --warmup 3 > gpurun_out/r1j_bench_d4.json 2> gpurun_out/r1j_bench_d4.err; echo "bench d4 rc=$?"
cat gpurun_out/r1j_bench_d4.json
EFFDET_DUMP_OPS=gpurun_out/r1j_ops.json python bench.py > gpurun_out/r1j_bench.json 2> gpurun_out/r1j_bench.err; echo "bench rc=$?"
cat gpurun_out/r1j_bench.json
python bench.py --workload d0_infer_b1 --steps 50 --warmup 5 > gpurun_out/r1j_bench_b1.json 2> gpurun_out/r1j_bench_b1.err; echo "bench b1 rc=$?"
cat gpurun_out/r1j_bench_b1.json
EFFDET_SE_CLUSTER=0 python bench.py --workload d0_infer_b1 --steps 50 --warmup 5 > gpurun_out/r1j_bench_b1_nocluster.json 2> gpurun_out/r1j_bench_b1_nocluster.err; echo "bench b1 nc rc=$?"
cat gpurun_out/r1j_bench_b1_nocluster.json
