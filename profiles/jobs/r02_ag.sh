# round 2, job AG: 2^17 (and 2^18 DIT) packed-16 plans as strided-4 + one-pass kernel
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02ag_pytest.txt 2>&1; tail -3 gpurun_out/r02ag_pytest.txt
python - > gpurun_out/r02ag_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for env in ("1", None):
    if env: os.environ["INTFFT_N13_TWO_PASS"] = env
    else: os.environ.pop("INTFFT_N13_TWO_PASS", None)
    print("INTFFT_N13_TWO_PASS =", env, "(set: the old 8 + 9 / 8 + 10 splits)")
    for d in (0, 1):
        q.time_plan(2048, steps=20, direction=d, NFFT=17, DATA_WIDTH=16, FORMAT=0)
        q.time_plan(1024, steps=20, direction=d, NFFT=18, DATA_WIDTH=16, FORMAT=0)
PY
cat gpurun_out/r02ag_times.txt
