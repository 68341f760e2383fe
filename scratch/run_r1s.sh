mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r1s_bench_n2.json 2> gpurun_out/r1s_bench_n2.err; echo "n2 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --workload d4_train_b8 > gpurun_out/r1s_bench_d4_n2.json 2> gpurun_out/r1s_bench_d4_n2.err; echo "d4 n2 rc=$?"
