mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/final_pytest.txt 2>&1; tail -2 gpurun_out/final_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_r01c_c2.json 2> gpurun_out/final_bench_c2.err; tail -c 300 gpurun_out/bench_r01c_c2.json
python bench.py --config c5 > gpurun_out/bench_r01c_c5.json 2> gpurun_out/final_bench_c5.err; tail -c 300 gpurun_out/bench_r01c_c5.json
