python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_v2.json 2> gpurun_out/r2_bench_n8_v2.err
tail -c 1800 gpurun_out/r2_bench_n8_v2.json; tail -3 gpurun_out/r2_bench_n8_v2.err
