# round 2, job I: one-pass 8192-point packed-16 kernel — parity + timing vs the two-pass schedule
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r02i_pytest.txt 2>&1; tail -4 gpurun_out/r02i_pytest.txt
python - > gpurun_out/r02i_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for env in ("1", None):
    if env: os.environ["INTFFT_N13_TWO_PASS"] = env
    else: os.environ.pop("INTFFT_N13_TWO_PASS", None)
    print("INTFFT_N13_TWO_PASS =", env)
    for d in (0, 1):
        q.time_plan(32768, steps=20, direction=d, NFFT=13, DATA_WIDTH=16, FORMAT=0)
        q.time_plan(32768, steps=20, direction=d, NFFT=13, DATA_WIDTH=12, FORMAT=0)
        q.time_plan(32768, steps=20, direction=d, NFFT=13, DATA_WIDTH=16, FORMAT=0, RNDMODE=1)
PY
cat gpurun_out/r02i_times.txt
