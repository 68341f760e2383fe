mkdir -p gpurun_out
for w in d0_train_b32 d4_train_b8 d0_infer_b1 d0_infer_b32 d2_infer_b64 d6_infer_b16 d6_infer_b16_1280; do
  st=20; [ $w = d4_train_b8 ] && st=10; [ $w = d6_infer_b16 ] && st=8; [ $w = d6_infer_b16_1280 ] && st=8; [ $w = d0_infer_b1 ] && st=50
  EFFDET_DUMP_OPS=gpurun_out/r1p_ops_$w.json python bench.py --workload $w --steps $st --warmup 5 > gpurun_out/r1p_bench_$w.json 2> gpurun_out/r1p_bench_$w.err; echo "$w rc=$?"
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1p_bench_reference.json 2> gpurun_out/r1p_bench_reference.err; echo "reference rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r1p_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r1p_ncu1.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:dwconv_tma_kernel -s 98 -c 3 -o gpurun_out/r1p_dwconv python bench.py --steps 1 --warmup 1 > gpurun_out/r1p_ncu2.log 2>&1; echo "ncu2 rc=$?"
