mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1u_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r1u_gpu_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r1u_bench.json 2> gpurun_out/r1u_bench.err; echo "bench rc=$?"
