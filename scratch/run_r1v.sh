mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1v_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r1v_gpu_tests.log
