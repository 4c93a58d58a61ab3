mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/c57_pytest.txt 2>&1
tail -3 gpurun_out/c57_pytest.txt
python profiles/quick_time.py others > gpurun_out/c57_quick.txt 2>&1
cat gpurun_out/c57_quick.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline | tail -c 500
