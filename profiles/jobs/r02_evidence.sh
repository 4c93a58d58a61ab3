# round-2 evidence (re-run after the last kernel change): ncu --set full of every BASELINE config's kernels (summary
# made on the box AND the report brought back), the launch list of the bench command, the default bench line, the
# reference arm, smoke(), the per-NFFT sweep timed and once more under ncu for per-launch DRAM bytes.
TAG=${1:-r02}
mkdir -p gpurun_out
# (the report stays on the box: with ten launches it exceeds what gpurun copies back; its summaries travel)
ncu --set full --clock-control none -k regex:'fast|strided|n13' -f -o /tmp/${TAG}_all python profiles/prof_plan.py c2 c3 c4 c5 c5u > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
python profiles/summarize_ncu.py /tmp/${TAG}_all.ncu-rep gpurun_out/${TAG}_all.ncu_summary.txt > /dev/null 2>&1
ncu -i /tmp/${TAG}_all.ncu-rep --page source --csv --print-source sass > /tmp/${TAG}_src.csv 2>/dev/null
python profiles/ncu_source_hot.py /tmp/${TAG}_src.csv 12 > gpurun_out/${TAG}_all.ncu_hot.txt 2>&1
ls -la /tmp/${TAG}_all.ncu-rep gpurun_out/${TAG}_all.ncu_summary.txt gpurun_out/${TAG}_all.ncu_hot.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_bench.launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 300 gpurun_out/${TAG}_bench_reference.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; tail -2 gpurun_out/${TAG}_smoke.txt
python profiles/sweep_nfft.py > gpurun_out/${TAG}_sweep_nfft.jsonl 2> gpurun_out/${TAG}_sweep.err
SWEEP_ONCE=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'fast|strided|n13|tile' --csv --log-file gpurun_out/${TAG}_sweep_nfft.ncu.csv python profiles/sweep_nfft.py > /dev/null 2>&1
wc -l gpurun_out/${TAG}_sweep_nfft.jsonl gpurun_out/${TAG}_sweep_nfft.ncu.csv
