nvidia-smi topo -m > gpurun_out/r2_topo_n8.txt 2>&1; lscpu | grep -i "numa\|socket\|model name" >> gpurun_out/r2_topo_n8.txt; numactl -H >> gpurun_out/r2_topo_n8.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_v4.json 2> gpurun_out/r2_bench_n8_v4.err
tail -c 600 gpurun_out/r2_bench_n8_v4.json; tail -2 gpurun_out/r2_bench_n8_v4.err
