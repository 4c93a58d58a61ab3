# round 2, job T: three CTAs per SM for the 2^9..2^11-point packed-16 kernels (middle-round twiddles in shared memory)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02t_pytest.txt 2>&1; tail -4 gpurun_out/r02t_pytest.txt
python - > gpurun_out/r02t_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for d in (0, 1):
    for n in (9, 10, 11, 12):
        q.time_plan(65536 << (12 - n), steps=20, direction=d, NFFT=n, DATA_WIDTH=16, FORMAT=0)
    q.time_plan(2048, steps=20, direction=d, NFFT=17, DATA_WIDTH=16, FORMAT=0)
    q.time_plan(16384, steps=20, direction=d, NFFT=14, DATA_WIDTH=16, FORMAT=0)
    q.time_plan(262144, steps=20, direction=d, NFFT=10, DATA_WIDTH=12, FORMAT=0, RNDMODE=1)
sys.argv = ["x"]; import pair_time
PY
cat gpurun_out/r02t_times.txt
