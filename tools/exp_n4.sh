python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_n4_v2.json 2> gpurun_out/r2_bench_n4_v2.err
tail -c 900 gpurun_out/r2_bench_n4_v2.json; tail -2 gpurun_out/r2_bench_n4_v2.err
python bench.py --impl reference --gpus 4 --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_n4.json 2>/dev/null; tail -c 600 gpurun_out/r2_bench_ref_n4.json
