mkdir -p gpurun_out
python bench.py --workload d4_train_b8 --steps 10 --warmup 3 > gpurun_out/r1w_bench_d4.json 2> gpurun_out/r1w_bench_d4.err; echo "d4 rc=$?"
python bench.py --workload d0_infer_b32 --steps 20 --warmup 5 > gpurun_out/r1w_bench_b32.json 2> gpurun_out/r1w_bench_b32.err; echo "b32 rc=$?"
