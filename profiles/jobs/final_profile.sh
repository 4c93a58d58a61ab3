# round-1c evidence: ncu --set full of every BASELINE config's kernels (summarised on the box: the report
# itself is too large to bring back), the launch list of the bench command, the bench lines of all configs
# and the reference arm
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:'fast|strided|n13' -f -o /tmp/prof_r01c_all python profiles/prof_plan.py c2 c3 c4 c5 c5u > gpurun_out/fp_ncu.log 2>&1
tail -2 gpurun_out/fp_ncu.log
python profiles/summarize_ncu.py /tmp/prof_r01c_all.ncu-rep gpurun_out/r01c_all.ncu_summary.txt > /dev/null 2>&1
ls -la /tmp/prof_r01c_all.ncu-rep gpurun_out/
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c2_bench_r01c.launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/fp_bench_under_ncu.log 2>&1
for c in c2 c3 c4 c5 c5u; do
  python bench.py --config $c > gpurun_out/bench_r01c_$c.json 2> gpurun_out/bench_r01c_$c.err
  tail -c 200 gpurun_out/bench_r01c_$c.json
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r01c_reference.json 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
python profiles/sweep_nfft.py > gpurun_out/sweep_nfft_r01c.jsonl 2> gpurun_out/sweep.err
