mkdir -p gpurun_out
python profiles/quick_time.py f12 > gpurun_out/c14_quick.txt 2>&1
cat gpurun_out/c14_quick.txt
