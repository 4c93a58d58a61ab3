# round 2, job G: parity after packed STAGE-12 twiddles / pre-shift in every DIT 32-bit-lane kernel; timings
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r02g_pytest.txt 2>&1; tail -4 gpurun_out/r02g_pytest.txt
python - > gpurun_out/r02g_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for env in ("1", None):
    if env: os.environ["INTFFT_NO_PRESHIFT"] = env
    else: os.environ.pop("INTFFT_NO_PRESHIFT", None)
    print("INTFFT_NO_PRESHIFT =", env)
    q.time_plan(131072, steps=20, direction=1, NFFT=13, DATA_WIDTH=18, FORMAT=0)
    q.time_plan(65536, steps=20, direction=1, NFFT=12, DATA_WIDTH=18, FORMAT=0)
    q.time_plan(262144, steps=20, direction=1, NFFT=10, DATA_WIDTH=18, FORMAT=0)
    q.time_plan(4096, steps=20, direction=1, NFFT=16, DATA_WIDTH=18, FORMAT=0)
    q.time_plan(1024, steps=20, direction=1, NFFT=17, DATA_WIDTH=18, FORMAT=0)
PY
cat gpurun_out/r02g_times.txt
