mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/c21_pytest.txt 2>&1
tail -3 gpurun_out/c21_pytest.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_r01c_c2_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 900 gpurun_out/bench_r01c_c2_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2>> gpurun_out/bench_2gpu.err
tail -c 300 gpurun_out/bench_ref_2gpu.json
