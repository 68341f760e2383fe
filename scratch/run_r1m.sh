mkdir -p gpurun_out
python scratch/bench_bn.py b2 b5 > gpurun_out/r1m_bn.log 2>&1; cat gpurun_out/r1m_bn.log
python -m pytest tests -m gpu -x -q > gpurun_out/r1m_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r1m_gpu_tests.log
EFFDET_DUMP_OPS=gpurun_out/r1m_ops_d4.json python bench.py --workload d4_train_b8 --steps 10 --warmup 3 > gpurun_out/r1m_bench_d4.json 2> gpurun_out/r1m_bench_d4.err; echo "bench d4 rc=$?"
EFFDET_DUMP_OPS=gpurun_out/r1m_ops.json python bench.py > gpurun_out/r1m_bench.json 2> gpurun_out/r1m_bench.err; echo "bench rc=$?"
