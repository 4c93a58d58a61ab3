mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/sw_pytest.txt 2>&1
tail -2 gpurun_out/sw_pytest.txt
python profiles/sweep_nfft.py > gpurun_out/sweep_nfft_r01c.jsonl 2> gpurun_out/sweep.err
tail -3 gpurun_out/sweep.err
SWEEP_ONCE=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'fast|strided|n13|tile' --csv --log-file gpurun_out/sweep_nfft_r01c.ncu.csv python profiles/sweep_nfft.py > /dev/null 2>&1
wc -l gpurun_out/sweep_nfft_r01c.jsonl gpurun_out/sweep_nfft_r01c.ncu.csv
