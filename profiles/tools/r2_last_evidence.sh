# Last round-2 evidence at HEAD (6 GPU-minutes left): launch lists of the default bench command (short form) and of
# one D4 training step.  Only small text summaries are left in gpurun_out/.
set -x
export EFFDET_BENCH_NO_CPU=1
B="python bench.py --no-sub-records --steps 2 --warmup 1"
timeout 140 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file /tmp/r2L_launches.csv $B > gpurun_out/r2L_ncu_a.log 2>&1
python profiles/tools/ncu_durations.py /tmp/r2L_launches.csv > gpurun_out/r2L_launches_d0_train.txt
python profiles/summarize.py launches /tmp/r2L_launches.csv > gpurun_out/r2L_launches_d0_train_shares.txt
D="python bench.py --workload d4_train_b8 --no-sub-records --steps 1 --warmup 1"
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file /tmp/r2L_d4.csv $D > gpurun_out/r2L_ncu_b.log 2>&1
python profiles/tools/ncu_durations.py /tmp/r2L_d4.csv > gpurun_out/r2L_launches_d4_train.txt
python profiles/summarize.py launches /tmp/r2L_d4.csv > gpurun_out/r2L_launches_d4_train_shares.txt
ls -la gpurun_out/r2L*
