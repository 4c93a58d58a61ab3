mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/c22_pytest.txt 2>&1
tail -4 gpurun_out/c22_pytest.txt
python profiles/quick_time.py others > gpurun_out/c22_quick.txt 2>&1
cat gpurun_out/c22_quick.txt
