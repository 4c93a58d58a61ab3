# round 2, final 8-GPU evidence: bench line under torchrun at N = 8 (c5 = the BASELINE 1M-frame job, 2^17 frames per rank),
# the reference arm under torchrun, product-level multi-device host path vs the bare-copy ceiling at 1 / 2 / 4 / 8 devices
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02final_bench_8gpu.json 2> gpurun_out/r02final_bench_8gpu.err; tail -c 300 gpurun_out/r02final_bench_8gpu.json; tail -2 gpurun_out/r02final_bench_8gpu.err
python profiles/multi_e2e.py > gpurun_out/r02final_multi_e2e.jsonl 2>&1; cat gpurun_out/r02final_multi_e2e.jsonl
