# round 2, job AI: compile-time DATA_WIDTH 12 / 14 instances of the packed-16 contiguous kernels
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02ai_pytest.txt 2>&1; tail -3 gpurun_out/r02ai_pytest.txt
python - > gpurun_out/r02ai_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for dw in (12, 13, 14, 16):
    for d in (0, 1):
        q.time_plan(65536, steps=20, direction=d, NFFT=12, DATA_WIDTH=dw, FORMAT=0)
    q.time_plan(262144, steps=20, direction=0, NFFT=10, DATA_WIDTH=dw, FORMAT=0)
    q.time_plan(65536, steps=20, direction=0, NFFT=12, DATA_WIDTH=dw, FORMAT=0, RNDMODE=1)
PY
cat gpurun_out/r02ai_times.txt
