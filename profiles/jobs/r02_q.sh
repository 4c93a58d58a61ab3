# round 2, job Q: TMA-staged strided pass for G = 4 as well (NFFT 14..16, 16 rows x 1 KB)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02q_pytest.txt 2>&1; tail -4 gpurun_out/r02q_pytest.txt
python - > gpurun_out/r02q_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for env in ("0", "1"):
    os.environ["INTFFT_STRIDED_TMA"] = env
    print("INTFFT_STRIDED_TMA =", env)
    for d in (0, 1):
        q.time_plan(16384, steps=20, direction=d, NFFT=14, DATA_WIDTH=16, FORMAT=0)
        q.time_plan(8192, steps=20, direction=d, NFFT=15, DATA_WIDTH=16, FORMAT=0)
        q.time_plan(4096, steps=20, direction=d, NFFT=16, DATA_WIDTH=16, FORMAT=0)
        q.time_plan(4096, steps=20, direction=d, NFFT=16, DATA_WIDTH=12, FORMAT=0, RNDMODE=1)
PY
cat gpurun_out/r02q_times.txt
