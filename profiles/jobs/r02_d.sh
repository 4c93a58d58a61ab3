# round 2, job D: full GPU test suite + the new bench line (all BASELINE configs in one run) + reference arm
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r02d_pytest.txt 2>&1; tail -4 gpurun_out/r02d_pytest.txt
(time python bench.py --steps 20 --warmup 5) > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -c 1500 gpurun_out/r02d_bench.json; tail -5 gpurun_out/r02d_bench.err
(time python bench.py --impl reference --steps 20 --warmup 5) > gpurun_out/r02d_bench_ref.json 2> gpurun_out/r02d_bench_ref.err; tail -c 600 gpurun_out/r02d_bench_ref.json; tail -4 gpurun_out/r02d_bench_ref.err
