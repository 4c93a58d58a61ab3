# round 2, job AA: bulk-TMA input for the two-round 32-bit-lane DIF kernels too (2^5..2^8 points)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02aa_pytest.txt 2>&1; tail -3 gpurun_out/r02aa_pytest.txt
python - > gpurun_out/r02aa_times.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "profiles")
import quick_time as q
for n in (5, 6, 7, 8):
    q.time_plan(1 << (27 - n), steps=20, direction=0, NFFT=n, DATA_WIDTH=18, FORMAT=0)
q.time_plan(1 << 19, steps=20, direction=0, NFFT=8, DATA_WIDTH=16, FORMAT=1)
q.time_plan(16384, steps=20, direction=0, NFFT=13, DATA_WIDTH=18, FORMAT=0)
PY
cat gpurun_out/r02aa_times.txt
