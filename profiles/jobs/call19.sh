mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -k "16bit or packed16 or c2 or bitrev or nfft20") > gpurun_out/c19_pytest.txt 2>&1
tail -3 gpurun_out/c19_pytest.txt
python profiles/quick_time.py c2 c4 > gpurun_out/c19_quick.txt 2>&1
cat gpurun_out/c19_quick.txt
