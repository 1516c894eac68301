python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_v2.json 2> gpurun_out/r2_bench_n2_v2.err
tail -c 2500 gpurun_out/r2_bench_n2_v2.json; tail -5 gpurun_out/r2_bench_n2_v2.err
