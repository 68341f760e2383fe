mkdir -p gpurun_out
python scratch/bench_bn.py b2 b3 b5 b7 > gpurun_out/r1l_bn.log 2>&1; cat gpurun_out/r1l_bn.log
python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "bn_train or full_training" > gpurun_out/r1l_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r1l_tests.log
