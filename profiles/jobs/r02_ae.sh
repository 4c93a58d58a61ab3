# round 2, job AE: packed-16 G = 4 TMA strided pass with a three-tile landing ring (two frames of lead)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02ae_pytest.txt 2>&1; tail -3 gpurun_out/r02ae_pytest.txt
python - > gpurun_out/r02ae_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for d in (0, 1):
    q.time_plan(16384, steps=20, direction=d, NFFT=14, DATA_WIDTH=16, FORMAT=0)
    q.time_plan(8192, steps=20, direction=d, NFFT=15, DATA_WIDTH=16, FORMAT=0)
    q.time_plan(4096, steps=20, direction=d, NFFT=16, DATA_WIDTH=16, FORMAT=0)
    q.time_plan(256, steps=20, direction=d, NFFT=20, DATA_WIDTH=16, FORMAT=0)
PY
cat gpurun_out/r02ae_times.txt
