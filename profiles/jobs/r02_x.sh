# round 2, job X: 32-bit-lane contiguous DIF kernels with the tile landed by one bulk TMA copy
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02x_pytest.txt 2>&1; tail -4 gpurun_out/r02x_pytest.txt
python - > gpurun_out/r02x_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for n in (9, 10, 11, 12):
    q.time_plan(32768 << (12 - n), steps=20, direction=0, NFFT=n, DATA_WIDTH=18, FORMAT=0)
q.time_plan(65536, steps=20, direction=0, NFFT=12, DATA_WIDTH=16, FORMAT=1)
q.time_plan(65536, steps=20, direction=0, NFFT=12, DATA_WIDTH=16, FORMAT=0, TWDL_WIDTH=18)
q.time_plan(2048, steps=20, direction=0, NFFT=16, DATA_WIDTH=18, FORMAT=0)
q.time_plan(512, steps=20, direction=0, NFFT=18, DATA_WIDTH=18, FORMAT=0)
q.time_plan(32768, steps=20, direction=0, NFFT=12, DATA_WIDTH=20, FORMAT=0, RNDMODE=1)
PY
cat gpurun_out/r02x_times.txt
