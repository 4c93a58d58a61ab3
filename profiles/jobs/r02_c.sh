# round 2, job C: parity after the trpl18 / pre-shift changes; c5 with and without pre-shifted twiddles
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r02c_pytest.txt 2>&1; tail -3 gpurun_out/r02c_pytest.txt
python - > gpurun_out/r02c_c5.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for env in ("1", None):
    if env: os.environ["INTFFT_NO_PRESHIFT"] = env
    else: os.environ.pop("INTFFT_NO_PRESHIFT", None)
    print("INTFFT_NO_PRESHIFT =", env)
    q.time_plan(131072, steps=20, direction=1, NFFT=13, DATA_WIDTH=18, FORMAT=0)
    q.time_plan(131072, steps=20, direction=1, NFFT=13, DATA_WIDTH=17, FORMAT=0, TWDL_WIDTH=18)
    q.time_plan(131072, steps=20, direction=1, NFFT=13, DATA_WIDTH=20, FORMAT=0, TWDL_WIDTH=24)
PY
cat gpurun_out/r02c_c5.txt
