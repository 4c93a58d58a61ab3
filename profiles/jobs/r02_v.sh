# round 2, job V: bounds tests hoisted in the fast64 / c2 store loops; fused pair after the DIT store change
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02v_pytest.txt 2>&1; tail -4 gpurun_out/r02v_pytest.txt
python profiles/quick_time.py c2 c3 c4 c5 > gpurun_out/r02v_times.txt 2>&1; cat gpurun_out/r02v_times.txt
python profiles/pair_time.py > gpurun_out/r02v_pair.jsonl 2>&1; cat gpurun_out/r02v_pair.jsonl
