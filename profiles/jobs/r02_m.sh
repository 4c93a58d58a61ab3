# round 2, job M: on-device Taylor twiddles (NFFT >= 17 strided passes) — parity + c4 timing with / without
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02m_pytest.txt 2>&1; tail -4 gpurun_out/r02m_pytest.txt
python - > gpurun_out/r02m_times.txt 2>&1 <<'PY'
import sys, os, time
sys.path.insert(0, "profiles")
import quick_time as q
import intfftk_b200 as ib
for env in ("99", None):
    if env: os.environ["INTFFT_TAYLOR_MIN_NFFT"] = env
    else: os.environ.pop("INTFFT_TAYLOR_MIN_NFFT", None)
    print("INTFFT_TAYLOR_MIN_NFFT =", env, "(99 = tables only, unset = device Taylor from NFFT 17)")
    t0 = time.perf_counter(); c = ib.Core(ib.Generics(NFFT=20, DATA_WIDTH=16, FORMAT=0), 256, 0); t1 = time.perf_counter(); c.close()
    print(f"plan creation NFFT=20: {(t1 - t0) * 1e3:.1f} ms")
    for d in (0, 1):
        q.time_plan(256, steps=20, direction=d, NFFT=20, DATA_WIDTH=16, FORMAT=0)
        q.time_plan(2048, steps=20, direction=d, NFFT=17, DATA_WIDTH=16, FORMAT=0)
        q.time_plan(512, steps=20, direction=d, NFFT=18, DATA_WIDTH=18, FORMAT=0)
PY
cat gpurun_out/r02m_times.txt
