# round 2, job E: fused pair kernel — parity (pair tests) and timing against the two-launch form; bench line
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q -k "pair or reentrant or mirror") > gpurun_out/r02e_pytest.txt 2>&1; tail -4 gpurun_out/r02e_pytest.txt
python profiles/pair_time.py > gpurun_out/r02e_pair.jsonl 2>&1; cat gpurun_out/r02e_pair.jsonl
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench.json').read().strip().splitlines()[0])
print('c2', round(d['ms_per_step'],4), round(d['roofline']['frac'],3))
for k,v in d['also'].items(): print(k, round(v['ms_per_step'],3), round(v['frac'],3))
PY
