mkdir -p gpurun_out
python profiles/sanitize_plans.py > gpurun_out/r02_san_plain.txt 2>&1; tail -2 gpurun_out/r02_san_plain.txt
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python profiles/sanitize_plans.py > gpurun_out/r02_san_racecheck.txt 2>&1; tail -15 gpurun_out/r02_san_racecheck.txt
timeout 900 compute-sanitizer --tool memcheck python profiles/sanitize_plans.py > gpurun_out/r02_san_memcheck.txt 2>&1; tail -6 gpurun_out/r02_san_memcheck.txt
