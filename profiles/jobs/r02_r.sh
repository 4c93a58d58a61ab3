# round 2, job R: 32-bit-lane strided pass (G = 8) with 2-D TMA load + store vs the cp.async / STG form
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02r_pytest.txt 2>&1; tail -4 gpurun_out/r02r_pytest.txt
python - > gpurun_out/r02r_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for env in ("0", "1"):
    os.environ["INTFFT_STRIDED_TMA"] = env
    print("INTFFT_STRIDED_TMA =", env)
    q.time_plan(4096, steps=20, NFFT=16, DATA_WIDTH=24, FORMAT=1)
    q.time_plan(4096, steps=20, direction=1, NFFT=16, DATA_WIDTH=24, FORMAT=1)
    for d in (0, 1):
        q.time_plan(512, steps=20, direction=d, NFFT=18, DATA_WIDTH=18, FORMAT=0)
        q.time_plan(1024, steps=20, direction=d, NFFT=17, DATA_WIDTH=14, FORMAT=1)
        q.time_plan(2048, steps=20, direction=d, NFFT=16, DATA_WIDTH=16, FORMAT=1)
        q.time_plan(8192, steps=20, direction=d, NFFT=14, DATA_WIDTH=18, FORMAT=0)
        q.time_plan(2048, steps=20, direction=d, NFFT=16, DATA_WIDTH=18, FORMAT=0)
PY
cat gpurun_out/r02r_times.txt
