# round 2, job J (8 GPUs): bench line under torchrun at N = 8 (c5 = the BASELINE 1M-frame job, 2^17 frames per rank),
# reference arm, product-level multi-device host path vs the bare-copy ceiling at 1 / 2 / 4 / 8 devices
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02j_bench_8gpu.json 2> gpurun_out/r02j_bench_8gpu.err; tail -c 300 gpurun_out/r02j_bench_8gpu.json; tail -3 gpurun_out/r02j_bench_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02j_bench_4gpu.json 2> gpurun_out/r02j_bench_4gpu.err; tail -c 200 gpurun_out/r02j_bench_4gpu.json
python profiles/multi_e2e.py > gpurun_out/r02j_multi_e2e.jsonl 2>&1; cat gpurun_out/r02j_multi_e2e.jsonl
nvidia-smi topo -m > gpurun_out/r02j_topo.txt 2>&1; lscpu | egrep 'Model name|Socket|NUMA|^CPU\(s\)' >> gpurun_out/r02j_topo.txt
