# round 2, job P: packed-16 strided pass with the column block moved by 2-D TMA (load + store) vs the cp.async / STG form
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02p_pytest.txt 2>&1; tail -4 gpurun_out/r02p_pytest.txt
python - > gpurun_out/r02p_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for env in ("0", "1"):
    os.environ["INTFFT_STRIDED_TMA"] = env
    print("INTFFT_STRIDED_TMA =", env)
    for d in (0, 1):
        q.time_plan(256, steps=20, direction=d, NFFT=20, DATA_WIDTH=16, FORMAT=0)
        q.time_plan(2048, steps=20, direction=d, NFFT=17, DATA_WIDTH=16, FORMAT=0)
        q.time_plan(512, steps=20, direction=d, NFFT=19, DATA_WIDTH=12, FORMAT=0)
        q.time_plan(1024, steps=20, direction=d, NFFT=18, DATA_WIDTH=16, FORMAT=0, RNDMODE=1)
PY
cat gpurun_out/r02p_times.txt
