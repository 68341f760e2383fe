set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1i_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r1i_gpu_tests.log
EFFDET_DUMP_OPS=gpurun_out/r1i_ops.json python bench.py > gpurun_out/r1i_bench.json 2> gpurun_out/r1i_bench.err; echo "bench rc=$?"
cat gpurun_out/r1i_bench.json
EFFDET_DUMP_OPS=gpurun_out/r1i_ops_d4.json python bench.py --workload d4_train_b8 --steps 10 --warmup 3 > gpurun_out/r1i_bench_d4.json 2> gpurun_out/r1i_bench_d4.err; echo "bench d4 rc=$?"
cat gpurun_out/r1i_bench_d4.json
python bench.py --workload d2_infer_b64 --steps 10 --warmup 3 > gpurun_out/r1i_bench_d2.json 2> gpurun_out/r1i_bench_d2.err; echo "bench d2 rc=$?"
cat gpurun_out/r1i_bench_d2.json
