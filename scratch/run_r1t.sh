mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1t_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r1t_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r1t_bench.json 2> gpurun_out/r1t_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r1t_bench.json; tail -3 gpurun_out/r1t_bench.err
