mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/c36_pytest.txt 2>&1
tail -3 gpurun_out/c36_pytest.txt
python profiles/quick_time.py c3 wide > gpurun_out/c36_quick.txt 2>&1
cat gpurun_out/c36_quick.txt
