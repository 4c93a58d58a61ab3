set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/c1_pytest.txt 2>&1
tail -5 gpurun_out/c1_pytest.txt
python bench.py > gpurun_out/c1_bench_c2.json 2> gpurun_out/c1_bench_c2.err
tail -c 600 gpurun_out/c1_bench_c2.json
python profiles/pcie_probe.py > gpurun_out/c1_pcie.txt 2>&1
cat gpurun_out/c1_pcie.txt
python profiles/quick_time.py c2 others > gpurun_out/c1_quick.txt 2>&1
cat gpurun_out/c1_quick.txt
