# round 2, job AB: pre-shifted twiddles + product halves for the 32-bit-lane DIF TRUNCATE kernels
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02ab_pytest.txt 2>&1; tail -3 gpurun_out/r02ab_pytest.txt
python - > gpurun_out/r02ab_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for env in ("1", None):
    if env: os.environ["INTFFT_NO_PRESHIFT"] = env
    else: os.environ.pop("INTFFT_NO_PRESHIFT", None)
    print("INTFFT_NO_PRESHIFT =", env)
    for n, b in ((8, 1 << 19), (10, 1 << 17), (12, 32768), (13, 16384), (16, 2048), (18, 512)):
        q.time_plan(b, steps=20, direction=0, NFFT=n, DATA_WIDTH=18, FORMAT=0)
    q.time_plan(32768, steps=20, direction=0, NFFT=12, DATA_WIDTH=24, FORMAT=0)
    q.time_plan(131072, steps=10, direction=1, NFFT=13, DATA_WIDTH=18, FORMAT=0)
PY
cat gpurun_out/r02ab_times.txt
