# round 2, job AF: one-pass 16384-point packed-16 kernel vs the strided-4 + contiguous-10 schedule
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02af_pytest.txt 2>&1; tail -5 gpurun_out/r02af_pytest.txt
python - > gpurun_out/r02af_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for env in ("1", None):
    if env: os.environ["INTFFT_N14_TWO_PASS"] = env
    else: os.environ.pop("INTFFT_N14_TWO_PASS", None)
    print("INTFFT_N14_TWO_PASS =", env)
    for d in (0, 1):
        q.time_plan(16384, steps=20, direction=d, NFFT=14, DATA_WIDTH=16, FORMAT=0)
        q.time_plan(16384, steps=20, direction=d, NFFT=14, DATA_WIDTH=12, FORMAT=0)
        q.time_plan(16384, steps=20, direction=d, NFFT=14, DATA_WIDTH=16, FORMAT=0, RNDMODE=1)
        q.time_plan(32768, steps=20, direction=d, NFFT=13, DATA_WIDTH=16, FORMAT=0)
PY
cat gpurun_out/r02af_times.txt
