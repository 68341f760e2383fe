set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_forward.py -m gpu -x -q -k "dwconv_and_se or network_forward" > gpurun_out/r1k_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r1k_tests.log
python scratch/bench_bn.py b2 b5 > gpurun_out/r1k_bn.log 2>&1; cat gpurun_out/r1k_bn.log
ncu --set full --clock-control none --import-source on -k regex:bn_act_bwd_ --launch-skip 6 -c 2 -o gpurun_out/r1k_bn python scratch/bench_bn.py b2 > gpurun_out/r1k_ncu.log 2>&1; echo "ncu rc=$?"
