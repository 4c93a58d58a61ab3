# round 2, job F (2 GPUs): the bench line under torchrun, the reference arm under torchrun, multi-device API test
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02f_bench_2gpu.json 2> gpurun_out/r02f_bench_2gpu.err; tail -c 400 gpurun_out/r02f_bench_2gpu.json; tail -5 gpurun_out/r02f_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --impl reference > gpurun_out/r02f_ref_2gpu.json 2> gpurun_out/r02f_ref_2gpu.err; tail -c 300 gpurun_out/r02f_ref_2gpu.json
(timeout 600 python -m pytest tests -m gpu -x -q -k "multi_device or ring or pair") > gpurun_out/r02f_pytest.txt 2>&1; tail -3 gpurun_out/r02f_pytest.txt
python profiles/pair_time.py > gpurun_out/r02f_pair.jsonl 2>&1; cat gpurun_out/r02f_pair.jsonl
python profiles/multi_e2e.py > gpurun_out/r02f_multi_e2e.jsonl 2>&1; cat gpurun_out/r02f_multi_e2e.jsonl
