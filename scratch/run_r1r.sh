mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1r_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r1r_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r1r_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r1r_smoke.log
python bench.py > gpurun_out/r1r_bench.json 2> gpurun_out/r1r_bench.err; echo "bench rc=$?"
python bench.py --workload d0_infer_b1 --steps 50 > gpurun_out/r1r_bench_b1.json 2> gpurun_out/r1r_bench_b1.err; echo "bench b1 rc=$?"
