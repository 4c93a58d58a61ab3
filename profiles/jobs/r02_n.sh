# round 2, job N: strided-pass scheduling (balanced round-robin deal for packed-16, contiguous ranges for 8-byte elements)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02n_pytest.txt 2>&1; tail -3 gpurun_out/r02n_pytest.txt
python - > gpurun_out/r02n_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for d in (0, 1):
    q.time_plan(256, steps=20, direction=d, NFFT=20, DATA_WIDTH=16, FORMAT=0)
    q.time_plan(2048, steps=20, direction=d, NFFT=17, DATA_WIDTH=16, FORMAT=0)
    q.time_plan(4096, steps=20, direction=d, NFFT=16, DATA_WIDTH=16, FORMAT=0)
    q.time_plan(16384, steps=20, direction=d, NFFT=14, DATA_WIDTH=16, FORMAT=0)
    q.time_plan(512, steps=20, direction=d, NFFT=18, DATA_WIDTH=18, FORMAT=0)
    q.time_plan(8192, steps=20, direction=d, NFFT=14, DATA_WIDTH=18, FORMAT=0)
q.time_plan(4096, steps=20, NFFT=16, DATA_WIDTH=24, FORMAT=1)
PY
cat gpurun_out/r02n_times.txt
