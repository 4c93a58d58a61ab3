mkdir -p gpurun_out
python profiles/quick_time.py tiny > gpurun_out/c42_quick.txt 2>&1
cat gpurun_out/c42_quick.txt
