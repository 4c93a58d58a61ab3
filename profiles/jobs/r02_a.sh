# round 2, job A: GPU tests after the exec refactor + group-size sweep for the two-pass plans
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r02a_pytest.txt 2>&1; tail -3 gpurun_out/r02a_pytest.txt
python profiles/group_sweep.py > gpurun_out/r02a_group_sweep.jsonl 2> gpurun_out/r02a_group_sweep.err
cat gpurun_out/r02a_group_sweep.jsonl; tail -3 gpurun_out/r02a_group_sweep.err
