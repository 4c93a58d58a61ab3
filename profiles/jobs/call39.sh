mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/bench_r01c_c2_4gpu.json 2> gpurun_out/bench_4gpu.err
tail -c 700 gpurun_out/bench_r01c_c2_4gpu.json; tail -3 gpurun_out/bench_4gpu.err
