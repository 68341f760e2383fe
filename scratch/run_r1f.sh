set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1f_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r1f_gpu_tests.log
EFFDET_DUMP_OPS=gpurun_out/r1f_ops.json python bench.py > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err; echo "bench rc=$?"
cat gpurun_out/r1f_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1f_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r1f_ncu1.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:dwconv_tma_kernel -s 96 -c 48 -o gpurun_out/r1f_dwconv python bench.py --steps 1 --warmup 1 > gpurun_out/r1f_ncu2.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out/*.ncu-rep
