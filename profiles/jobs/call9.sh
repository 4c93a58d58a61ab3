mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -k "lane64 or multipass or widths or pair or natural") > gpurun_out/c9_pytest.txt 2>&1
tail -3 gpurun_out/c9_pytest.txt
python profiles/quick_time.py c3 > gpurun_out/c9_quick.txt 2>&1
cat gpurun_out/c9_quick.txt
