# round 2, job U: packed-16 DIT kernels with the whole-tile test hoisted out of the store / prefetch loops
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02u_pytest.txt 2>&1; tail -4 gpurun_out/r02u_pytest.txt
python - > gpurun_out/r02u_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for n in (8, 9, 10, 11, 12):
    q.time_plan(65536 << (12 - n), steps=20, direction=1, NFFT=n, DATA_WIDTH=16, FORMAT=0)
q.time_plan(65536, steps=20, direction=1, NFFT=12, DATA_WIDTH=12, FORMAT=0)
q.time_plan(65536, steps=20, direction=1, NFFT=12, DATA_WIDTH=16, FORMAT=0, RNDMODE=1)
q.time_plan(4096, steps=20, direction=1, NFFT=16, DATA_WIDTH=16, FORMAT=0)
q.time_plan(256, steps=20, direction=1, NFFT=20, DATA_WIDTH=16, FORMAT=0)
q.time_plan(65536, steps=20, direction=0, NFFT=12, DATA_WIDTH=16, FORMAT=0)
sys.argv = ["x"]
exec(open("profiles/pair_time.py").read())
PY
cat gpurun_out/r02u_times.txt
