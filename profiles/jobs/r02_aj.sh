# round 2, job AJ: compile-time DATA_WIDTH 12 / 14 in the one-pass and TMA strided packed-16 kernels
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02aj_pytest.txt 2>&1; tail -3 gpurun_out/r02aj_pytest.txt
python - > gpurun_out/r02aj_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for dw in (12, 13, 14):
    for d in (0, 1):
        q.time_plan(32768, steps=20, direction=d, NFFT=13, DATA_WIDTH=dw, FORMAT=0)
        q.time_plan(16384, steps=20, direction=d, NFFT=14, DATA_WIDTH=dw, FORMAT=0)
        q.time_plan(4096, steps=20, direction=d, NFFT=16, DATA_WIDTH=dw, FORMAT=0)
        q.time_plan(256, steps=20, direction=d, NFFT=20, DATA_WIDTH=dw, FORMAT=0)
PY
cat gpurun_out/r02aj_times.txt
