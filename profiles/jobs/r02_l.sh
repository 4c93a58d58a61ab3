# round 2, job L: fast64_kernel input prefetch (cp.async landing area) — parity of the 64-bit-lane plans + c3 timing
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -k "lane64 or wide or baseline or c3 or 64") > gpurun_out/r02l_pytest.txt 2>&1; tail -3 gpurun_out/r02l_pytest.txt
python - > gpurun_out/r02l_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for env in ("1", None):
    if env: os.environ["INTFFT_F64_NO_PREFETCH"] = env
    else: os.environ.pop("INTFFT_F64_NO_PREFETCH", None)
    print("INTFFT_F64_NO_PREFETCH =", env)
    q.time_plan(4096, steps=20, NFFT=16, DATA_WIDTH=24, FORMAT=1)
    q.time_plan(4096, steps=20, direction=1, NFFT=16, DATA_WIDTH=24, FORMAT=1)
    q.time_plan(65536, steps=20, NFFT=12, DATA_WIDTH=24, FORMAT=1)
    q.time_plan(65536, steps=20, direction=1, NFFT=12, DATA_WIDTH=24, FORMAT=1)
    q.time_plan(1048576, steps=20, NFFT=8, DATA_WIDTH=30, FORMAT=1)
PY
cat gpurun_out/r02l_times.txt
