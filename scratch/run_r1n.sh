mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2900 --csv --log-file gpurun_out/r1n_d4_launches.csv python bench.py --workload d4_train_b8 --steps 1 --warmup 1 > gpurun_out/r1n_ncu.log 2>&1; echo "ncu rc=$?"
